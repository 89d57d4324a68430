// Device side of the XTC (Gromacs xdr3dfcoord) decoder, shared by xtc_decode.cu (frames -> x[frame][atom][3]) and
// binning.cu (frames -> bins, no coordinates stored).  The format is restated in host/trajio.c; PARITY WITH GROMACS
// UNPINNED (the library is not in the reference tree), checked bit for bit against that host decoder.
//
// A frame's bit stream is sequential, but only its CONTROL is: where a group of atoms starts, the current smallidx
// and run length.  The coordinates of a group never depend on an earlier group (each starts with a full triple).
// So decoding is two kernels:
//   xtc_skim_kernel      one warp per frame walks the control bits only and drops a checkpoint
//                        {bit position, atom index, smallidx, run} every SEG_ATOMS atoms;
//   a segment kernel     one thread per (frame, segment) decodes from its checkpoint to the next and hands every
//                        atom's integer coordinates times fl(1/precision) to an emitter (store, or bin).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace xtcdev {

static __constant__ int c_magicints[73] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512,
    645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007,
    32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561,
    832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983,
    13316085, 16777216};
// 1.0 / magicints[i]: constant-folded, correctly rounded doubles (0 for the unused head)
static __constant__ double c_inv_magic[73] = {
    0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0 / 8.0, 1.0 / 10.0, 1.0 / 12.0, 1.0 / 16.0, 1.0 / 20.0, 1.0 /
    25.0, 1.0 / 32.0, 1.0 / 40.0, 1.0 / 50.0, 1.0 / 64.0, 1.0 / 80.0, 1.0 / 101.0, 1.0 / 128.0, 1.0 / 161.0, 1.0 /
    203.0, 1.0 / 256.0, 1.0 / 322.0, 1.0 / 406.0, 1.0 / 512.0, 1.0 / 645.0, 1.0 / 812.0, 1.0 / 1024.0, 1.0 /
    1290.0, 1.0 / 1625.0, 1.0 / 2048.0, 1.0 / 2580.0, 1.0 / 3250.0, 1.0 / 4096.0, 1.0 / 5060.0, 1.0 / 6501.0, 1.0
    / 8192.0, 1.0 / 10321.0, 1.0 / 13003.0, 1.0 / 16384.0, 1.0 / 20642.0, 1.0 / 26007.0, 1.0 / 32768.0, 1.0 /
    41285.0, 1.0 / 52015.0, 1.0 / 65536.0, 1.0 / 82570.0, 1.0 / 104031.0, 1.0 / 131072.0, 1.0 / 165140.0, 1.0 /
    208063.0, 1.0 / 262144.0, 1.0 / 330280.0, 1.0 / 416127.0, 1.0 / 524287.0, 1.0 / 660561.0, 1.0 / 832255.0, 1.0
    / 1048576.0, 1.0 / 1321122.0, 1.0 / 1664510.0, 1.0 / 2097152.0, 1.0 / 2642245.0, 1.0 / 3329021.0, 1.0 /
    4194304.0, 1.0 / 5284491.0, 1.0 / 6658042.0, 1.0 / 8388607.0, 1.0 / 10568983.0, 1.0 / 13316085.0, 1.0 /
    16777216.0};
constexpr int FIRSTIDX = 9, LASTIDX = 72;
#ifndef GCB_XTC_SEG_ATOMS
#define GCB_XTC_SEG_ATOMS 64
#endif
constexpr int SEG_ATOMS = GCB_XTC_SEG_ATOMS;

struct FrameDev {                  // what the kernels need of a frame
  unsigned long long offset;       // payload offset in the uploaded buffer (4-byte aligned)
  unsigned nbytes;                 // compressed bytes (0: natoms <= 9, payload is 3*natoms big-endian floats)
  float prec;
  int minint[3];
  unsigned sizeint[3];
  int smallidx;
};

struct Checkpoint { unsigned bitpos, atom; int smallidx, run; };

__host__ __device__ inline int max_segments(int natoms) { return natoms / SEG_ATOMS + 2; }

// layout of the work buffer: [nframes] segment counts, then [nframes][max_segments] checkpoints
__host__ __device__ inline size_t work_ckpt_offset(long nframes) { return ((size_t)nframes * sizeof(int) + 15) & ~(size_t)15; }

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

// Decode the atoms [cp.atom, i_end) of one frame from the checkpoint `cp` and hand them, in atom order, to
// emit(atom index in the frame, x, y, z) with the decoder's f32 product (float)int * fl(1/precision).
// (natoms > 9; the skim kernel has validated the whole control sequence.)
template <class EMIT>
__device__ __forceinline__ void decode_segment(const uint8_t *__restrict__ buf, const FrameDev &m, const Checkpoint &cp,
                                               int i_end, EMIT &emit)
{
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
        for (int d = 0; d < 3; d++) cur[d] += prev[d] - smallnum;
        if (k == 0) {                    // the first two atoms of a run were swapped by the encoder
          for (int d = 0; d < 3; d++) { const int tmp = cur[d]; cur[d] = prev[d]; prev[d] = tmp; }
          emit(i++, __fmul_rn((float)prev[0], inv), __fmul_rn((float)prev[1], inv), __fmul_rn((float)prev[2], inv));
        } else {
          for (int d = 0; d < 3; d++) prev[d] = cur[d];
        }
        emit(i++, __fmul_rn((float)cur[0], inv), __fmul_rn((float)cur[1], inv), __fmul_rn((float)cur[2], inv));
      }
    } else {
      emit(i++, __fmul_rn((float)cur[0], inv), __fmul_rn((float)cur[1], inv), __fmul_rn((float)cur[2], inv));
    }
    smallidx += is_smaller;
    if (is_smaller < 0) { smallnum = smaller; smaller = smallidx > FIRSTIDX ? c_magicints[smallidx - 1] / 2 : 0; }
    else if (is_smaller > 0) { smaller = smallnum; smallnum = c_magicints[smallidx] / 2; }
    sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)c_magicints[smallidx];
    inv_small = c_inv_magic[smallidx];
  }
}

}  // namespace xtcdev
