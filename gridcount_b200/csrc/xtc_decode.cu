// XTC (Gromacs xdr3dfcoord) frames decoded on the GPU, so that only compressed bytes cross PCIe.
//
// Replaces, for .xtc input, what read_first_x/read_next_x hand to the frame loop
// (g_ri3Dc.c:392,428): the decoder is Gromacs library code (not in the reference tree; format
// restated in host/trajio.c, PARITY WITH GROMACS UNPINNED, checked bit for bit against that host
// decoder).  The host only parses the 92-byte frame headers (gcb_xtc_scan); the payloads are
// uploaded as they are and one thread decodes one frame: a frame's bit stream is strictly
// sequential, frames are independent, and a batch holds thousands of them.  Output is the same
// x[frame][atom][3] f32 layout the binning kernels read, with the decoder's f32 product
// (float)int * fl(1/precision).
#include "gcb_internal.h"
#include <cuda_runtime.h>
#include <algorithm>
#include <vector>

namespace {

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

// receiveints: `nbits` bits hold one number, low byte first; split it over the mixed radix (s0, s1, s2)
__device__ __forceinline__ void read_triple(BitReader &br, int nbits, const unsigned sizes[3], int out[3])
{
  if (nbits <= 64) {
    unsigned long long v = 0;
    int shift = 0, nb = nbits;
    while (nb > 32) { v |= (unsigned long long)__byte_perm(br.read(32), 0, 0x0123) << shift; shift += 32; nb -= 32; }
    while (nb > 8) { v |= (unsigned long long)br.read(8) << shift; shift += 8; nb -= 8; }
    if (nb > 0) v |= (unsigned long long)br.read(nb) << shift;
    if (nbits <= 32) {
      unsigned u = (unsigned)v;
      out[2] = (int)(u % sizes[2]); u /= sizes[2];
      out[1] = (int)(u % sizes[1]); u /= sizes[1];
      out[0] = (int)u;
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

// status[f]: 0 ok, 1 damaged stream
__global__ void __launch_bounds__(64)
xtc_decode_kernel(const uint8_t *__restrict__ buf, const FrameDev *__restrict__ meta, long nframes, int natoms,
                  float *__restrict__ x, int *__restrict__ status)
{
  const long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nframes) return;
  const FrameDev m = meta[f];
  float *o = x + (size_t)f * natoms * 3;
  if (natoms <= 9) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(buf + m.offset);
    for (int i = 0; i < natoms * 3; i++) o[i] = __uint_as_float(__byte_perm(__ldg(w + i), 0, 0x0123));
    status[f] = 0;
    return;
  }
  BitReader br;
  br.init(buf + m.offset, m.nbytes);
  unsigned bitsizeint[3] = {0, 0, 0};
  int bitsize;
  if ((m.sizeint[0] | m.sizeint[1] | m.sizeint[2]) > 0xffffffu) {
    for (int d = 0; d < 3; d++) bitsizeint[d] = (unsigned)bits_for(m.sizeint[d]);
    bitsize = 0;
  } else {
    bitsize = bits_for_triple(m.sizeint);
  }
  int smallidx = m.smallidx;
  int smaller = c_magicints[max(FIRSTIDX, smallidx - 1)] / 2;
  int smallnum = c_magicints[smallidx] / 2;
  unsigned sizesmall[3];
  sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)c_magicints[smallidx];
  const float inv = __fdiv_rn(1.0f, m.prec);
  int run = 0, i = 0;
  bool bad = false;
  while (i < natoms && !bad) {
    int cur[3], prev[3], is_smaller = 0;
    if (bitsize == 0) { for (int d = 0; d < 3; d++) cur[d] = (int)br.read((int)bitsizeint[d]); }
    else read_triple(br, bitsize, m.sizeint, cur);
    i++;
    for (int d = 0; d < 3; d++) { cur[d] += m.minint[d]; prev[d] = cur[d]; }
    if (br.read(1) == 1u) {
      run = (int)br.read(5);
      is_smaller = run % 3;
      run -= is_smaller;
      is_smaller--;
    }
    if (run > 0) {
      if (i + run / 3 > natoms) { bad = true; break; }
      for (int k = 0; k < run; k += 3) {
        read_triple(br, smallidx, sizesmall, cur);
        i++;
        for (int d = 0; d < 3; d++) cur[d] += prev[d] - smallnum;
        if (k == 0) {                    // the first two atoms of a run were swapped by the encoder
          for (int d = 0; d < 3; d++) { const int t = cur[d]; cur[d] = prev[d]; prev[d] = t; }
          for (int d = 0; d < 3; d++) *o++ = __fmul_rn((float)prev[d], inv);
        } else {
          for (int d = 0; d < 3; d++) prev[d] = cur[d];
        }
        for (int d = 0; d < 3; d++) *o++ = __fmul_rn((float)cur[d], inv);
      }
    } else {
      for (int d = 0; d < 3; d++) *o++ = __fmul_rn((float)cur[d], inv);
    }
    smallidx += is_smaller;
    if (smallidx < FIRSTIDX || smallidx > LASTIDX) { bad = true; break; }
    if (is_smaller < 0) { smallnum = smaller; smaller = smallidx > FIRSTIDX ? c_magicints[smallidx - 1] / 2 : 0; }
    else if (is_smaller > 0) { smaller = smallnum; smallnum = c_magicints[smallidx] / 2; }
    sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)c_magicints[smallidx];
    if (br.over) bad = true;
  }
  status[f] = bad ? 1 : 0;
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
                              float *x_dev, void *meta_dev, int *status_dev, void *stream, void *meta_host_pinned)
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
  xtc_decode_kernel<<<(unsigned)((nframes + 63) / 64), 64, 0, st>>>(buf_dev, (const FrameDev *)meta_dev, nframes, natoms, x_dev, status_dev);
  GCB_CUDA(cudaGetLastError());
  return GCB_OK;
}

size_t gcb_xtc_meta_bytes_(long nframes) { return sizeof(FrameDev) * (size_t)nframes; }

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
  uint8_t *d_buf = nullptr; void *d_meta = nullptr; int *d_status = nullptr;
  GCB_CUDA(cudaMalloc(&d_buf, hi - lo + 16));
  GCB_CUDA(cudaMalloc(&d_meta, gcb_xtc_meta_bytes_(nframes)));
  GCB_CUDA(cudaMalloc(&d_status, sizeof(int) * (size_t)nframes));
  int rc = GCB_OK;
  if (cudaMemcpy(d_buf, buf + lo, hi - lo, cudaMemcpyHostToDevice) != cudaSuccess) rc = gcb::fail(GCB_ERR_CUDA, "H2D of the xtc bytes failed");
  if (!rc) rc = gcb_xtc_decode_on_stream_(d_buf, rel.data(), nframes, natoms, x_dev, d_meta, d_status, nullptr, nullptr);
  std::vector<int> status((size_t)nframes, 0);
  if (!rc && cudaMemcpy(status.data(), d_status, sizeof(int) * (size_t)nframes, cudaMemcpyDeviceToHost) != cudaSuccess)
    rc = gcb::fail(GCB_ERR_CUDA, "xtc decode kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_buf); cudaFree(d_meta); cudaFree(d_status);
  if (rc) return rc;
  for (long f = 0; f < nframes; f++)
    if (status[(size_t)f]) return gcb::fail(GCB_ERR_FORMAT, "xtc frame %ld: damaged compressed stream", f);
  return GCB_OK;
}

}  // extern "C"
