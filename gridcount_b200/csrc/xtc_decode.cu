// XTC (Gromacs xdr3dfcoord) frames decoded on the GPU, so that only compressed bytes cross PCIe.
//
// Replaces, for .xtc input, what read_first_x/read_next_x hand to the frame loop
// (g_ri3Dc.c:392,428): the decoder is Gromacs library code (not in the reference tree; format
// restated in host/trajio.c, PARITY WITH GROMACS UNPINNED, checked bit for bit against that host
// decoder).  The host only parses the 92-byte frame headers (gcb_xtc_scan); the payloads are
// uploaded as they are.  One thread per frame walks the control bits of the stream and drops
// checkpoints; one thread per (frame, 256-atom segment) then decodes coordinates.  Output is the same
// x[frame][atom][3] f32 layout the binning kernels read, with the decoder's f32 product
// (float)int * fl(1/precision).
#include "gcb_internal.h"
#include <cuda_runtime.h>
#include <algorithm>
#include <vector>

namespace {

__constant__ double c_inv_magic[73];          // 1.0 / magicints[i] (0 for the unused head), filled on first use
__constant__ int c_magicints[73] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812,
    1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015,
    65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122,
    1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216};
constexpr int FIRSTIDX = 9, LASTIDX = 72;
constexpr uint32_t XTC_MAGIC = 1995;

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline float bef(const uint8_t *p) { uint32_t w = be32(p); float f; memcpy(&f, &w, 4); return f; }

struct FrameDev {                  // what the kernel needs of a frame
  unsigned long long offset;       // payload offset in the uploaded buffer (4-byte aligned)
  unsigned nbytes;                 // compressed bytes (0: natoms <= 9, payload is 3*natoms big-endian floats)
  float prec;
  int minint[3];
  unsigned sizeint[3];
  int smallidx;
};

// MSB-first bit reader over big-endian 32-bit words
struct BitReader {
  const uint32_t *w;
  unsigned nwords, pos;
  unsigned long long acc;          // unread bits, left aligned
  int have;
  bool over;
  __device__ __forceinline__ void init(const uint8_t *p, unsigned nbytes)
  {
    w = reinterpret_cast<const uint32_t *>(p); nwords = (nbytes + 3) / 4; pos = 0; acc = 0; have = 0; over = false;
  }
  __device__ __forceinline__ unsigned read(int n)      // 0 <= n <= 32
  {
    if (n == 0) return 0u;
    if (have < n) {
      unsigned word = 0;
      if (pos < nwords) word = __byte_perm(__ldg(w + pos), 0, 0x0123); else over = true;
      pos++;
      acc |= (unsigned long long)word << (32 - have);
      have += 32;
    }
    const unsigned v = (unsigned)(acc >> (64 - n));
    acc <<= n;
    have -= n;
    return v;
  }
};

// v / s for v < 2^53, s < 2^24 without the 64-bit division routine: the double product is off by at most one
__device__ __forceinline__ unsigned long long div_by(unsigned long long v, unsigned s, double inv_s, unsigned &rem)
{
  unsigned long long q = __double2ull_rd(__dmul_rn((double)v, inv_s));
  long long r = (long long)(v - q * s);
  if (r < 0) { q--; r += s; }
  else if (r >= (long long)s) { q++; r -= s; }
  rem = (unsigned)r;
  return q;
}

// receiveints: `nbits` bits hold one number, low byte first; split it over the mixed radix (s0, s1, s2).
// inv1, inv2 = 1.0/sizes[1], 1.0/sizes[2].
__device__ __forceinline__ void read_triple(BitReader &br, int nbits, const unsigned sizes[3], double inv1, double inv2, int out[3])
{
  if (nbits <= 64) {
    unsigned long long v = 0;
    int shift = 0, nb = nbits;
    while (nb > 32) { v |= (unsigned long long)__byte_perm(br.read(32), 0, 0x0123) << shift; shift += 32; nb -= 32; }
    while (nb > 8) { v |= (unsigned long long)br.read(8) << shift; shift += 8; nb -= 8; }
    if (nb > 0) v |= (unsigned long long)br.read(nb) << shift;
    if (nbits <= 52) {
      unsigned r2, r1;
      v = div_by(v, sizes[2], inv2, r2);
      v = div_by(v, sizes[1], inv1, r1);
      out[2] = (int)r2; out[1] = (int)r1; out[0] = (int)v;
    } else {
      out[2] = (int)(v % sizes[2]); v /= sizes[2];
      out[1] = (int)(v % sizes[1]); v /= sizes[1];
      out[0] = (int)v;
    }
  } else {
    unsigned __int128 v = 0;
    int shift = 0, nb = nbits;
    while (nb > 8) { v |= (unsigned __int128)br.read(8) << shift; shift += 8; nb -= 8; }
    if (nb > 0) v |= (unsigned __int128)br.read(nb) << shift;
    out[2] = (int)(v % sizes[2]); v /= sizes[2];
    out[1] = (int)(v % sizes[1]); v /= sizes[1];
    out[0] = (int)v;
  }
}

__device__ __forceinline__ int bits_for(unsigned size)
{
  unsigned num = 1;
  int nbits = 0;
  while (size >= num && nbits < 32) { nbits++; num <<= 1; }
  return nbits;
}
__device__ __forceinline__ int bits_for_triple(const unsigned s[3])
{
  unsigned __int128 t = (unsigned __int128)s[0] * s[1] * s[2];
  int nbytes = 0, nbits = 0;
  while (t >> 8) { t >>= 8; nbytes++; }
  const unsigned top = (unsigned)t;
  unsigned num = 1;
  while (top >= num) { nbits++; num *= 2; }
  return nbits + nbytes * 8;
}

// ---------------------------------------------------------------------------
// A frame's bit stream is sequential, but only its CONTROL is: where a group of atoms starts, the
// current smallidx and run length.  The coordinates of a group never depend on an earlier group
// (each starts with a full triple).  So decoding is two kernels:
//   xtc_skim_kernel    one warp per frame walks the control bits only (skip the full triple,
//                      read the flag / run field, skip the run) and drops a checkpoint
//                      {bit position, atom index, smallidx, run} every SEG_ATOMS atoms;
//   xtc_decode_kernel  one thread per (frame, segment) decodes from its checkpoint to the next.
// The skim is ~20 instructions per group and the only part bound by per-thread latency; the
// divisions, float conversions and stores run with hundreds of thousands of threads.
// ---------------------------------------------------------------------------
constexpr int SEG_ATOMS = 256;

struct Checkpoint { unsigned bitpos, atom; int smallidx, run; };

__host__ __device__ inline int max_segments(int natoms) { return natoms / SEG_ATOMS + 2; }

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

struct FrameCoding { int bitsize; unsigned bitsizeint[3]; };
__device__ __forceinline__ FrameCoding frame_coding(const FrameDev &m)
{
  FrameCoding c;
  c.bitsizeint[0] = c.bitsizeint[1] = c.bitsizeint[2] = 0;
  if ((m.sizeint[0] | m.sizeint[1] | m.sizeint[2]) > 0xffffffu) {
    for (int d = 0; d < 3; d++) c.bitsizeint[d] = (unsigned)bits_for(m.sizeint[d]);
    c.bitsize = 0;
  } else {
    c.bitsize = bits_for_triple(m.sizeint);
  }
  return c;
}

// status[f]: 0 ok, 1 damaged stream.  nseg[f] = checkpoints written.
// One WARP per frame.  Most groups leave run and smallidx alone (flag bit 0) and then all have the same
// length, so the flag bits of the next 32 groups sit at known, closely spaced positions: lane j fetches
// the flag of group j (one or two cache lines for the whole warp), a ballot finds the first flagged group,
// and the warp jumps over the unflagged prefix in one step; a flagged group is handled on its own.
constexpr int SKIM_WARPS = 4;
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
  const unsigned long long total_bits = (unsigned long long)m.nbytes * 8ull;
  const FrameCoding fc = frame_coding(m);
  const unsigned fullbits = fc.bitsize ? (unsigned)fc.bitsize : fc.bitsizeint[0] + fc.bitsizeint[1] + fc.bitsizeint[2];
  int smallidx = m.smallidx, run = 0, k = 0;
  unsigned i = 0, next_mark = 0;
  unsigned long long bp = 0;
  bool bad = false, over = false;
  while (i < (unsigned)natoms) {                                              // all lanes carry the same state
    if (i >= next_mark) {
      if (lane == 0) out[k] = Checkpoint{(unsigned)bp, i, smallidx, run};
      k++;
      next_mark = i + SEG_ATOMS;
    }
    const unsigned g = 1u + (unsigned)(run / 3);                              // atoms per unflagged group
    const unsigned long long glen = fullbits + 1ull + (unsigned long long)(run / 3) * (unsigned)smallidx;
    unsigned nmax = min(32u, ((unsigned)natoms - i + g - 1u) / g);
    nmax = min(nmax, (next_mark - i + g - 1u) / g);                           // land on the group that starts a segment
    const unsigned long long p = bp + fullbits + (unsigned long long)lane * glen;
    bool flag = false;
    if (lane < nmax) {
      if (p >= total_bits) flag = true;                                       // runs off the end: stop there, caught below
      else flag = ((__byte_perm(__ldg(w + (unsigned)(p >> 5)), 0, 0x0123) >> (31u - ((unsigned)p & 31u))) & 1u) != 0u;
    }
    const unsigned flags = __ballot_sync(0xffffffffu, flag);
    const unsigned j0 = flags ? (unsigned)(__ffs((int)flags) - 1) : nmax;
    bp += (unsigned long long)j0 * glen;
    i += j0 * g;
    if (j0 < nmax) {                                                          // a group that changes run / smallidx
      const unsigned long long q = bp + fullbits;
      if (q + 6 > total_bits) { bad = true; break; }
      const unsigned six = peek_bits(w, nwords, (unsigned)q, 6, over);
      if (!(six & 32u)) { bad = true; break; }
      run = (int)(six & 31u);
      const int is_smaller = run % 3 - 1;
      run -= run % 3;
      i += 1u + (unsigned)(run / 3);
      bp = q + 6 + (unsigned long long)(run / 3) * (unsigned)smallidx;
      smallidx += is_smaller;
    }
    if (smallidx < FIRSTIDX || smallidx > LASTIDX || i > (unsigned)natoms || bp > total_bits || over) { bad = true; break; }
  }
  if (lane == 0) { nseg[f] = k; status[f] = bad ? 1 : 0; }
}

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
  float *o = x + ((size_t)f * natoms + cp.atom) * 3;
  BitReader br;
  br.init(buf + m.offset, m.nbytes);
  br.pos = cp.bitpos >> 5;
  if (cp.bitpos & 31u) br.read((int)(cp.bitpos & 31u));          // drop the bits before the checkpoint
  const FrameCoding fc = frame_coding(m);
  int smallidx = cp.smallidx;
  int smaller = c_magicints[max(FIRSTIDX, smallidx - 1)] / 2;
  int smallnum = c_magicints[smallidx] / 2;
  unsigned sizesmall[3];
  sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)c_magicints[smallidx];
  const float inv = __fdiv_rn(1.0f, m.prec);
  const double inv_s1 = 1.0 / (double)m.sizeint[1], inv_s2 = 1.0 / (double)m.sizeint[2];
  double inv_small = c_inv_magic[smallidx];
  int run = cp.run, i = (int)cp.atom;
  while (i < i_end) {
    int cur[3], prev[3], is_smaller = 0;
    if (fc.bitsize == 0) { for (int d = 0; d < 3; d++) cur[d] = (int)br.read((int)fc.bitsizeint[d]); }
    else read_triple(br, fc.bitsize, m.sizeint, inv_s1, inv_s2, cur);
    i++;
    for (int d = 0; d < 3; d++) { cur[d] += m.minint[d]; prev[d] = cur[d]; }
    if (br.read(1) == 1u) {
      run = (int)br.read(5);
      is_smaller = run % 3;
      run -= is_smaller;
      is_smaller--;
    }
    if (run > 0) {
      for (int k = 0; k < run; k += 3) {
        read_triple(br, smallidx, sizesmall, inv_small, inv_small, cur);
        i++;
        for (int d = 0; d < 3; d++) cur[d] += prev[d] - smallnum;
        if (k == 0) {                    // the first two atoms of a run were swapped by the encoder
          for (int d = 0; d < 3; d++) { const int tmp = cur[d]; cur[d] = prev[d]; prev[d] = tmp; }
          for (int d = 0; d < 3; d++) *o++ = __fmul_rn((float)prev[d], inv);
        } else {
          for (int d = 0; d < 3; d++) prev[d] = cur[d];
        }
        for (int d = 0; d < 3; d++) *o++ = __fmul_rn((float)cur[d], inv);
      }
    } else {
      for (int d = 0; d < 3; d++) *o++ = __fmul_rn((float)cur[d], inv);
    }
    smallidx += is_smaller;                 // the skim kernel has validated the whole control sequence
    if (is_smaller < 0) { smallnum = smaller; smaller = smallidx > FIRSTIDX ? c_magicints[smallidx - 1] / 2 : 0; }
    else if (is_smaller > 0) { smaller = smallnum; smallnum = c_magicints[smallidx] / 2; }
    sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)c_magicints[smallidx];
    inv_small = c_inv_magic[smallidx];
  }
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
      if (fr.smallidx < FIRSTIDX || fr.smallidx > LASTIDX || !(fr.precision > 0))
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

// Decode `nframes` frames (all with `natoms` atoms) whose payloads sit in the DEVICE copy `buf_dev` of the bytes
// the frames were scanned from.  x_dev receives x[nframes][natoms][3].  Enqueued on `stream` (cudaStream_t).
int gcb_xtc_decode_on_stream_(const uint8_t *buf_dev, const gcb_xtc_frame *frames, long nframes, int natoms,
                              float *x_dev, void *meta_dev, int *status_dev, void *stream, void *meta_host_pinned,
                              void *work_dev)
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
  {
    static bool table_done[64] = {false};
    int dev = 0;
    GCB_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !table_done[dev]) {
      static const int mi[73] = {
          0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812,
          1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015,
          65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122,
          1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216};
      double inv[73];
      for (int i = 0; i < 73; i++) inv[i] = mi[i] ? 1.0 / (double)mi[i] : 0.0;
      GCB_CUDA(cudaMemcpyToSymbol(c_inv_magic, inv, sizeof inv));
      table_done[dev] = true;
    }
  }
  GCB_CUDA(cudaMemcpyAsync(meta_dev, meta, sizeof(FrameDev) * (size_t)nframes, cudaMemcpyHostToDevice, st));
  if (!meta_host_pinned) GCB_CUDA(cudaStreamSynchronize(st));       // `local` is pageable and goes out of scope
  // work_dev: [nframes] segment counts, then [nframes][max_segments] checkpoints
  int *nseg = static_cast<int *>(work_dev);
  Checkpoint *ckpt = reinterpret_cast<Checkpoint *>(static_cast<char *>(work_dev) + (((size_t)nframes * sizeof(int) + 15) & ~(size_t)15));
  xtc_skim_kernel<<<(unsigned)((nframes + SKIM_WARPS - 1) / SKIM_WARPS), SKIM_WARPS * 32, 0, st>>>(buf_dev, (const FrameDev *)meta_dev, nframes, natoms, ckpt, nseg, status_dev);
  GCB_CUDA(cudaGetLastError());
  const long long nthreads = (long long)nframes * max_segments(natoms);
  xtc_decode_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(buf_dev, (const FrameDev *)meta_dev, nframes, natoms, ckpt, nseg,
                                                                     x_dev, status_dev);
  GCB_CUDA(cudaGetLastError());
  return GCB_OK;
}

size_t gcb_xtc_meta_bytes_(long nframes) { return sizeof(FrameDev) * (size_t)nframes; }
size_t gcb_xtc_work_bytes_(long nframes, int natoms)
{
  return (((size_t)nframes * sizeof(int) + 15) & ~(size_t)15) + sizeof(Checkpoint) * (size_t)nframes * (size_t)max_segments(natoms);
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
