// XTC (Gromacs xdr3dfcoord) frames decoded on the GPU, so that only compressed bytes cross PCIe.
//
// Replaces, for .xtc input, what read_first_x/read_next_x hand to the frame loop
// (g_ri3Dc.c:392,428): the decoder is Gromacs library code (not in the reference tree; format
// restated in host/trajio.c, PARITY WITH GROMACS UNPINNED, checked bit for bit against that host
// decoder).  The host only parses the 92-byte frame headers (gcb_xtc_scan); the payloads are
// uploaded as they are.  One warp per frame walks the control bits of the stream and drops
// checkpoints; one thread per (frame, 256-atom segment) then decodes coordinates (xtc_device.cuh).  Output is
// the same x[frame][atom][3] f32 layout the binning kernels read, with the decoder's f32 product
// (float)int * fl(1/precision); gcb_counter_add_xtc bins straight out of the segment decoder instead.
#include "gcb_internal.h"
#include "xtc_device.cuh"
#include <cuda_runtime.h>
#include <algorithm>
#include <vector>

using namespace xtcdev;

namespace {

constexpr uint32_t XTC_MAGIC = 1995;

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline float bef(const uint8_t *p) { uint32_t w = be32(p); float f; memcpy(&f, &w, 4); return f; }

// random-access read of n <= 25 bits at bit position bp (MSB first)
__device__ __forceinline__ unsigned peek_bits(const uint32_t *w, unsigned nwords, unsigned bp, int n, bool &over)
{
  const unsigned i = bp >> 5;
  if (i >= nwords) { over = true; return 0u; }
  const unsigned w0 = __byte_perm(__ldg(w + i), 0, 0x0123);
  const unsigned w1 = i + 1 < nwords ? __byte_perm(__ldg(w + i + 1), 0, 0x0123) : 0u;
  const unsigned long long both = ((unsigned long long)w0 << 32) | w1;
  return (unsigned)(both >> (64 - (bp & 31) - n)) & ((1u << n) - 1u);
}

// status[f]: 0 ok, 1 damaged stream.  nseg[f] = checkpoints written.
// One WARP per frame.  Most groups leave run and smallidx alone (flag bit 0) and then all have the same length, so
// the flag bits of the next groups sit at known, evenly spaced positions: lane j fetches the flags of groups j,
// j + 32, ... (SKIM_DEPTH independent loads in flight per lane, a handful of cache lines for the whole warp),
// ballots find the first flagged group, and the warp jumps over the unflagged prefix in one step -- up to a whole
// 256-atom segment per step; a flagged group is handled on its own.  The walk is a chain of dependent loads, so
// its time is (steps per frame) x (load latency) whatever the number of frames: the fewer steps, the better.
constexpr int SKIM_WARPS = 4;
constexpr int SKIM_DEPTH = 4;
// ceil(atoms / g) for the group sizes that can occur (g = 1 .. 11) and small atom counts (< 4096), without a division
// ceil(2^16 / g) for g = 1 .. 11: a 5-bit run field can name up to 10 small atoms behind the full one (the encoder stops at 8)
__constant__ unsigned c_rcp16[12] = {0u, 65536u, 32768u, 21846u, 16384u, 13108u, 10923u, 9363u, 8192u, 7282u, 6554u, 5958u};
__device__ __forceinline__ unsigned ceil_div_small(unsigned atoms, unsigned g)
{
  return ((atoms + g - 1u) * c_rcp16[g]) >> 16;
}

__global__ void __launch_bounds__(SKIM_WARPS * 32)
xtc_skim_kernel(const uint8_t *__restrict__ buf, const FrameDev *__restrict__ meta, long nframes, int natoms,
                Checkpoint *__restrict__ ckpt, int *__restrict__ nseg, int *__restrict__ status)
{
  const long f = (long)blockIdx.x * SKIM_WARPS + (threadIdx.x >> 5);
  const unsigned lane = threadIdx.x & 31;
  if (f >= nframes) return;
  const FrameDev m = meta[f];
  Checkpoint *out = ckpt + (size_t)f * max_segments(natoms);
  if (natoms <= 9) {
    if (lane == 0) { out[0] = Checkpoint{0u, 0u, 0, 0}; nseg[f] = 1; status[f] = 0; }
    return;
  }
  const uint32_t *w = reinterpret_cast<const uint32_t *>(buf + m.offset);
  const unsigned nwords = (m.nbytes + 3) / 4;
  const unsigned total_bits = m.nbytes * 8u;                 // < 2^31 (gcb_xtc_scan): every position below fits 32 bits
  const FrameCoding fc = frame_coding(m);
  const unsigned fullbits = fc.bitsize ? (unsigned)fc.bitsize : fc.bitsizeint[0] + fc.bitsizeint[1] + fc.bitsizeint[2];
  int smallidx = m.smallidx, run = 0, k = 0;
  unsigned i = 0, next_mark = 0, bp = 0;
  bool bad = false, over = false;
  constexpr unsigned NG = 32u * SKIM_DEPTH;                  // groups looked at per step
  while (i < (unsigned)natoms) {                             // all lanes carry the same state
    const unsigned g = 1u + (unsigned)(run / 3);             // atoms per unflagged group (1 .. 11)
    const unsigned glen = fullbits + 1u + (unsigned)(run / 3) * (unsigned)smallidx;
    const unsigned rem = (unsigned)natoms - i;
    const unsigned nmax = rem >= NG * g ? NG : (g == 1u ? rem : (rem + g - 1u) / g);       // a division only near the end
    // the walk only moves forward: pull the lines it is about to read (1 .. 5 KB ahead) into L1 now, so that the
    // dependent loads of the coming steps are L1 hits
    {
      const unsigned ahead = ((bp >> 3) & ~127u) + 128u * (lane + 8u);
      if (ahead < m.nbytes) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const uint8_t *>(w) + ahead));
    }
    // flag bit of group j sits at bp + fullbits + j * glen; a position at or past the end of the payload (a damaged
    // stream, caught below) reads as "flagged"
    const unsigned p0 = bp + fullbits + lane * glen;
    unsigned word[SKIM_DEPTH];
#pragma unroll
    for (int d = 0; d < SKIM_DEPTH; d++) {                   // all loads first: they overlap
      const unsigned p = p0 + 32u * d * glen;
      word[d] = (lane + 32u * d < nmax && p < total_bits) ? __ldg(w + (p >> 5)) : 0xffffffffu;
    }
    unsigned j0 = nmax;
#pragma unroll
    for (int d = SKIM_DEPTH - 1; d >= 0; d--) {
      const unsigned p = p0 + 32u * d * glen;
      const bool flag = lane + 32u * d < nmax && ((__byte_perm(word[d], 0, 0x0123) >> (31u - (p & 31u))) & 1u) != 0u;
      const unsigned flags = __ballot_sync(0xffffffffu, flag);
      if (flags) j0 = 32u * d + (unsigned)(__ffs((int)flags) - 1);
    }
    // Groups 0 .. j0-1 are unflagged and group j0 (if there is one) begins like them: all start in the state
    // (smallidx, run) at known positions.  Every group that is the first to start at or behind a segment mark gets
    // a checkpoint: the first one ceil((next_mark - i) / g) groups on (at most SEG_ATOMS atoms away), the following
    // ones every ceil(SEG_ATOMS / g) groups.
    {
      const unsigned jend = j0 < nmax ? j0 + 1u : j0;
      const unsigned stride = ceil_div_small(SEG_ATOMS, g);
      unsigned j = next_mark <= i ? 0u : ceil_div_small(next_mark - i, g);
      while (j < jend) {
        if (lane == 0) out[k] = Checkpoint{bp + j * glen, i + j * g, smallidx, run};
        k++;
        next_mark = i + j * g + SEG_ATOMS;
        j += stride;
      }
    }
    bp += j0 * glen;
    if (bp > total_bits) { bad = true; break; }
    i += j0 * g;
    if (j0 < nmax) {                                                          // a group that changes run / smallidx
      const unsigned q = bp + fullbits;
      if (q + 6u > total_bits) { bad = true; break; }
      const unsigned six = peek_bits(w, nwords, q, 6, over);
      if (!(six & 32u)) { bad = true; break; }
      run = (int)(six & 31u);
      const int is_smaller = run % 3 - 1;
      run -= run % 3;
      i += 1u + (unsigned)(run / 3);
      bp = q + 6u + (unsigned)(run / 3) * (unsigned)smallidx;
      if (bp > total_bits) { bad = true; break; }
      smallidx += is_smaller;
    }
    if (smallidx < FIRSTIDX || smallidx > LASTIDX || i > (unsigned)natoms || over) { bad = true; break; }
  }
  if (lane == 0) { nseg[f] = k; status[f] = bad ? 1 : 0; }
}

struct StoreEmit {
  float *frame;                      // x[f]
  __device__ __forceinline__ void operator()(int atom, float x, float y, float z) const
  {
    float *o = frame + 3 * (size_t)atom;
    o[0] = x; o[1] = y; o[2] = z;
  }
};

__global__ void __launch_bounds__(128)
xtc_decode_kernel(const uint8_t *__restrict__ buf, const FrameDev *__restrict__ meta, long nframes, int natoms,
                  const Checkpoint *__restrict__ ckpt, const int *__restrict__ nseg, float *__restrict__ x,
                  int *__restrict__ status)
{
  const int msegs = max_segments(natoms);
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long f = t / msegs;
  const int sg = (int)(t - f * msegs);
  if (f >= nframes || status[f] != 0 || sg >= nseg[f]) return;
  const FrameDev m = meta[f];
  if (natoms <= 9) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(buf + m.offset);
    float *o = x + (size_t)f * natoms * 3;
    for (int i = 0; i < natoms * 3; i++) o[i] = __uint_as_float(__byte_perm(__ldg(w + i), 0, 0x0123));
    return;
  }
  const Checkpoint cp = ckpt[(size_t)f * msegs + sg];
  const int i_end = sg + 1 < nseg[f] ? (int)ckpt[(size_t)f * msegs + sg + 1].atom : natoms;
  StoreEmit emit{x + (size_t)f * natoms * 3};
  decode_segment(buf, m, cp, i_end, emit);
}

}  // namespace

extern "C" {

// Parse frame headers; `frames` may be NULL to count.  Stops at the first incomplete frame.
int gcb_xtc_scan(const uint8_t *buf, size_t len, gcb_xtc_frame *frames, long cap, long *nframes_out, size_t *consumed)
{
  if (!buf || !nframes_out) return gcb::fail(GCB_ERR_ARG, "gcb_xtc_scan: null argument");
  size_t off = 0;
  long n = 0;
  while (len - off >= 56) {
    const uint8_t *p = buf + off;
    if (be32(p) != XTC_MAGIC) return gcb::fail(GCB_ERR_FORMAT, "not an xtc frame at byte %zu (magic %u)", off, be32(p));
    const int natoms = (int)be32(p + 4);
    if (natoms < 0 || (int)be32(p + 52) != natoms) return gcb::fail(GCB_ERR_FORMAT, "damaged xtc frame header at byte %zu", off);
    size_t total;
    gcb_xtc_frame fr;
    memset(&fr, 0, sizeof fr);
    fr.natoms = natoms;
    fr.step = (int)be32(p + 8);
    fr.time = bef(p + 12);
    for (int i = 0; i < 9; i++) fr.box[i] = bef(p + 16 + 4 * i);
    if (natoms <= 9) {
      total = 56 + (size_t)natoms * 12;
      if (len - off < total) break;
      fr.payload_offset = off + 56;
      fr.payload_bytes = 0;
    } else {
      if (len - off < 92) break;
      fr.precision = bef(p + 56);
      for (int d = 0; d < 3; d++) { fr.minint[d] = (int)be32(p + 60 + 4 * d); fr.maxint[d] = (int)be32(p + 72 + 4 * d); }
      fr.smallidx = (int)be32(p + 84);
      const size_t nbytes = be32(p + 88);
      total = 92 + ((nbytes + 3) & ~(size_t)3);
      if (len - off < total) break;
      if (fr.smallidx < FIRSTIDX || fr.smallidx > LASTIDX || !(fr.precision > 0) || nbytes >= ((size_t)1 << 28))   // (bit positions stay below 2^31)
        return gcb::fail(GCB_ERR_FORMAT, "damaged xtc frame header at byte %zu", off);
      for (int d = 0; d < 3; d++)
        if (fr.maxint[d] < fr.minint[d]) return gcb::fail(GCB_ERR_FORMAT, "damaged xtc frame header at byte %zu", off);
      fr.payload_offset = off + 92;
      fr.payload_bytes = (uint32_t)nbytes;
    }
    if (frames) {
      if (n >= cap) break;
      frames[n] = fr;
    }
    n++;
    off += total;
  }
  *nframes_out = n;
  if (consumed) *consumed = off;
  return GCB_OK;
}

// Stage 1 of decoding `nframes` frames (all with `natoms` atoms) whose payloads sit in the DEVICE copy `buf_dev` of
// the bytes the frames were scanned from: frame table upload + skim (checkpoints, verdicts) on `stream`.
int gcb_xtc_skim_on_stream_(const uint8_t *buf_dev, const gcb_xtc_frame *frames, long nframes, int natoms, void *meta_dev,
                            int *status_dev, void *stream, void *meta_host_pinned, void *work_dev)
{
  // with a pinned staging table the call does not wait for the stream
  std::vector<FrameDev> local;
  FrameDev *meta = static_cast<FrameDev *>(meta_host_pinned);
  if (!meta) { local.resize((size_t)nframes); meta = local.data(); }
  for (long f = 0; f < nframes; f++) {
    const gcb_xtc_frame &s = frames[f];
    if (s.natoms != natoms) return gcb::fail(GCB_ERR_FORMAT, "xtc frame %ld has %d atoms, expected %d", f, s.natoms, natoms);
    FrameDev &m = meta[(size_t)f];
    m.offset = s.payload_offset; m.nbytes = s.payload_bytes; m.prec = s.precision; m.smallidx = s.smallidx;
    for (int d = 0; d < 3; d++) { m.minint[d] = s.minint[d]; m.sizeint[d] = (unsigned)(s.maxint[d] - s.minint[d]) + 1u; }
  }
  cudaStream_t st = (cudaStream_t)stream;
  GCB_CUDA(cudaMemcpyAsync(meta_dev, meta, sizeof(FrameDev) * (size_t)nframes, cudaMemcpyHostToDevice, st));
  if (!meta_host_pinned) GCB_CUDA(cudaStreamSynchronize(st));       // `local` is pageable and goes out of scope
  int *nseg = static_cast<int *>(work_dev);
  Checkpoint *ckpt = reinterpret_cast<Checkpoint *>(static_cast<char *>(work_dev) + work_ckpt_offset(nframes));
  xtc_skim_kernel<<<(unsigned)((nframes + SKIM_WARPS - 1) / SKIM_WARPS), SKIM_WARPS * 32, 0, st>>>(buf_dev, (const FrameDev *)meta_dev, nframes, natoms, ckpt, nseg, status_dev);
  GCB_CUDA(cudaGetLastError());
  return GCB_OK;
}

// Stage 2: one thread per (frame, segment) decodes to x_dev[nframes][natoms][3]
int gcb_xtc_segments_on_stream_(const uint8_t *buf_dev, long nframes, int natoms, float *x_dev, const void *meta_dev,
                                int *status_dev, void *stream, const void *work_dev)
{
  const int *nseg = static_cast<const int *>(work_dev);
  const Checkpoint *ckpt = reinterpret_cast<const Checkpoint *>(static_cast<const char *>(work_dev) + work_ckpt_offset(nframes));
  const long long nthreads = (long long)nframes * max_segments(natoms);
  xtc_decode_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(buf_dev, (const FrameDev *)meta_dev, nframes, natoms, ckpt,
                                                                                       nseg, x_dev, status_dev);
  GCB_CUDA(cudaGetLastError());
  return GCB_OK;
}

int gcb_xtc_decode_on_stream_(const uint8_t *buf_dev, const gcb_xtc_frame *frames, long nframes, int natoms,
                              float *x_dev, void *meta_dev, int *status_dev, void *stream, void *meta_host_pinned,
                              void *work_dev)
{
  int rc = gcb_xtc_skim_on_stream_(buf_dev, frames, nframes, natoms, meta_dev, status_dev, stream, meta_host_pinned, work_dev);
  if (rc) return rc;
  return gcb_xtc_segments_on_stream_(buf_dev, nframes, natoms, x_dev, meta_dev, status_dev, stream, work_dev);
}

size_t gcb_xtc_meta_bytes_(long nframes) { return sizeof(FrameDev) * (size_t)nframes; }
size_t gcb_xtc_work_bytes_(long nframes, int natoms)
{
  return work_ckpt_offset(nframes) + sizeof(Checkpoint) * (size_t)nframes * (size_t)max_segments(natoms);
}

// Host bytes -> device frames, synchronous (tests, benchmarks, tools that keep decoded frames on the GPU).
int gcb_xtc_decode_device(const uint8_t *buf, size_t len, const gcb_xtc_frame *frames, long nframes, int natoms,
                          float *x_dev, int device)
{
  if (!buf || !frames || nframes < 0 || natoms <= 0 || !x_dev) return gcb::fail(GCB_ERR_ARG, "gcb_xtc_decode_device: bad argument");
  if (gcb_device_count() <= 0) return gcb::fail(GCB_ERR_CUDA, "no CUDA device: libgridcount_b200 has no CPU fallback");
  if (nframes == 0) return GCB_OK;
  GCB_CUDA(cudaSetDevice(device));
  // upload only the span the frames cover
  size_t lo = frames[0].payload_offset, hi = 0;
  for (long f = 0; f < nframes; f++) {
    const size_t b = frames[f].payload_offset, n = frames[f].natoms <= 9 ? (size_t)frames[f].natoms * 12 : ((size_t)frames[f].payload_bytes + 3) & ~(size_t)3;
    if (b + n > len) return gcb::fail(GCB_ERR_ARG, "xtc frame %ld lies outside the buffer", f);
    lo = std::min(lo, b); hi = std::max(hi, b + n);
  }
  lo &= ~(size_t)3;
  std::vector<gcb_xtc_frame> rel(frames, frames + nframes);
  for (auto &r : rel) r.payload_offset -= lo;
  uint8_t *d_buf = nullptr; void *d_meta = nullptr, *d_work = nullptr; int *d_status = nullptr;
  GCB_CUDA(cudaMalloc(&d_buf, hi - lo + 16));
  GCB_CUDA(cudaMalloc(&d_meta, gcb_xtc_meta_bytes_(nframes)));
  GCB_CUDA(cudaMalloc(&d_work, gcb_xtc_work_bytes_(nframes, natoms)));
  GCB_CUDA(cudaMalloc(&d_status, sizeof(int) * (size_t)nframes));
  int rc = GCB_OK;
  if (cudaMemcpy(d_buf, buf + lo, hi - lo, cudaMemcpyHostToDevice) != cudaSuccess) rc = gcb::fail(GCB_ERR_CUDA, "H2D of the xtc bytes failed");
  if (!rc) rc = gcb_xtc_decode_on_stream_(d_buf, rel.data(), nframes, natoms, x_dev, d_meta, d_status, nullptr, nullptr, d_work);
  std::vector<int> status((size_t)nframes, 0);
  if (!rc && cudaMemcpy(status.data(), d_status, sizeof(int) * (size_t)nframes, cudaMemcpyDeviceToHost) != cudaSuccess)
    rc = gcb::fail(GCB_ERR_CUDA, "xtc decode kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_buf); cudaFree(d_meta); cudaFree(d_status); cudaFree(d_work);
  if (rc) return rc;
  for (long f = 0; f < nframes; f++)
    if (status[(size_t)f]) return gcb::fail(GCB_ERR_FORMAT, "xtc frame %ld: damaged compressed stream", f);
  return GCB_OK;
}

}  // extern "C"
