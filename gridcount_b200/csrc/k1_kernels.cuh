// K1 device code: the float -> bin conversion (g_ri3Dc.c:144-147), the gather kernel, the XTC decode-and-bin kernel
// and the streaming kernel with its sinks (where a hit goes: private shared-memory grids, the uint32 grid, packed
// L2-resident counters).  Included by binning.cu only; sm_100a, built with -fmad=false.
#pragma once
#include "gcb_internal.h"
#include "xtc_device.cuh"
#include <cuda_runtime.h>
#include <cstdint>

namespace k1 {

struct GridParams {
  float a[3], b[3], delta[3], fmx[3];
  int my, mz;
};

enum { MODE_COUNT = 0, MODE_FLOAT = 1 };

// ---------------------------------------------------------------------------
// float -> bin, g_ri3Dc.c:144-147.  Returns the flat z-fastest cell, -1 when
// the atom fails a bounds test (u < 0 or x - b >= 0, two separate f32
// subtractions as in the reference), -2 when it passes all three but a
// quotient is >= mx or NaN (the reference would then write out of bounds).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int cell_of(const GridParams &P, float x, float y, float z)
{
  const float ux = __fsub_rn(x, P.a[0]), uy = __fsub_rn(y, P.a[1]), uz = __fsub_rn(z, P.a[2]);
  const bool outside = (ux < 0.0f) | (__fsub_rn(x, P.b[0]) >= 0.0f) |
                       (uy < 0.0f) | (__fsub_rn(y, P.b[1]) >= 0.0f) |
                       (uz < 0.0f) | (__fsub_rn(z, P.b[2]) >= 0.0f);
  if (outside) return -1;
  const float qx = __fdiv_rn(ux, P.delta[0]), qy = __fdiv_rn(uy, P.delta[1]), qz = __fdiv_rn(uz, P.delta[2]);
  const bool bad = !(qx < P.fmx[0]) | !(qy < P.fmx[1]) | !(qz < P.fmx[2]);
  if (bad) return -2;
  // (int)floor((double)q) == round-down conversion for 0 <= q < 2^31
  return (__float2int_rd(qx) * P.my + __float2int_rd(qy)) * P.mz + __float2int_rd(qz);
}

// count.c:88-92: single-image wrap, `< 0` below and `> L` above
__device__ __forceinline__ float wrap1(float v, float L)
{
  if (v < 0.0f) v = __fadd_rn(v, L);
  if (v > L) v = __fsub_rn(v, L);
  return v;
}

template <int MODE>
__device__ __forceinline__ void deposit(uint32_t *__restrict__ counts, float *__restrict__ acc, int cell, float w)
{
  if (MODE == MODE_COUNT) {
    atomicAdd(counts + cell, 1u);            // RED.E.ADD (no return value used)
  } else {
    atomicAdd(acc + cell, w);                // same-weight float adds commute: exact within one weight run
    atomicAdd(counts + cell, 1u);
  }
}

// One atomic per (warp, frame) on the per-frame hit counters.
__device__ __forceinline__ void count_hits(unsigned long long *__restrict__ hits, bool hit, long long f)
{
  unsigned todo = __ballot_sync(0xffffffffu, hit);
  const int lane = threadIdx.x & 31;
  while (todo) {
    const int leader = __ffs(todo) - 1;
    const long long fl = __shfl_sync(0xffffffffu, f, leader);
    const unsigned same = __ballot_sync(0xffffffffu, hit && f == fl);
    if (lane == leader) atomicAdd(hits + fl, (unsigned long long)__popc(same));
    todo &= ~same;
  }
}

// ---------------------------------------------------------------------------
// K1a  gather kernel: one lane per selected atom-frame e in [e_begin, e_end),
// direct LDG of x[f][gindex[i]].  Used for sparse index groups, unaligned
// inputs and the tail the streaming kernel leaves.
// ---------------------------------------------------------------------------
constexpr int GATHER_THREADS = 256;
constexpr int GATHER_ITEMS = 8;

template <int MODE, bool PBC>
__global__ void __launch_bounds__(GATHER_THREADS)
bin_gather_kernel(const float *__restrict__ x, int natoms, const int *__restrict__ gindex, unsigned gnx,
                  long long e_begin, long long e_end, GridParams P, uint32_t *__restrict__ counts,
                  float *__restrict__ acc, float w, const float *__restrict__ box3,
                  unsigned long long *__restrict__ hits, unsigned long long *__restrict__ diag)
{
  const long long chunk = (long long)GATHER_THREADS * GATHER_ITEMS;
  const long long nchunks = (e_end - e_begin + chunk - 1) / chunk;
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const long long e0 = e_begin + c * chunk;
    const long long f0 = e0 / gnx;                       // one 64-bit division per chunk
    const unsigned r0 = (unsigned)(e0 - f0 * gnx);
#pragma unroll
    for (int it = 0; it < GATHER_ITEMS; it++) {
      const long long e = e0 + it * GATHER_THREADS + threadIdx.x;
      const bool valid = e < e_end;
      unsigned rel = r0 + it * GATHER_THREADS + threadIdx.x;
      long long f = f0;
      if (rel >= gnx) { const unsigned q = rel / gnx; rel -= q * gnx; f += q; }
      int cell = -1;
      if (valid) {
        const int idx = __ldg(gindex + rel);
        const float *p = x + ((size_t)f * natoms + idx) * 3;
        float px = __ldg(p), py = __ldg(p + 1), pz = __ldg(p + 2);
        if (PBC) {
          const float *L = box3 + 3 * f;
          px = wrap1(px, __ldg(L)); py = wrap1(py, __ldg(L + 1)); pz = wrap1(pz, __ldg(L + 2));
        }
        cell = cell_of(P, px, py, pz);
        if (cell >= 0) deposit<MODE>(counts, acc, cell, w);
        else if (cell == -2) atomicAdd(diag, 1ull);
      }
      count_hits(hits, cell >= 0, f);
    }
  }
}

// ---------------------------------------------------------------------------
// K1x  XTC frames binned straight out of the segment decoder (xtc_device.cuh): one thread per (frame, 256-atom
// segment) decodes its atoms and bins the selected ones at once -- the reference's read_next_x + gridcount() for
// .xtc input (g_ri3Dc.c:425-428) without the decoded coordinates ever being stored.  Same f32 products as the
// decoder that writes x, same exact conversion as the gather kernel, so counts equal the two-step path bit for bit.
// Selection: arithmetic for groups g0 + i*step; otherwise the multiplicity rank[atom + 1] - rank[atom].
// ---------------------------------------------------------------------------
template <bool PBC, bool AP>
struct XtcBinEmit {
  const GridParams &P;
  uint32_t *__restrict__ counts;
  const unsigned *__restrict__ rank;
  unsigned long long *__restrict__ diag;
  int next_sel, step, sel_end;       // AP: the next selected atom, the group's stride, one past its last atom
  float L0, L1, L2;                  // PBC: this frame's box
  unsigned nh;
  __device__ __forceinline__ void operator()(int atom, float x, float y, float z)
  {
    unsigned mult = 1u;
    if (AP) {
      if (atom != next_sel) return;
      next_sel += step;
      if (next_sel >= sel_end) next_sel = 0x7fffffff;
    } else {
      mult = __ldg(rank + atom + 1) - __ldg(rank + atom);
      if (!mult) return;
    }
    if (PBC) { x = wrap1(x, L0); y = wrap1(y, L1); z = wrap1(z, L2); }
    const int cell = cell_of(P, x, y, z);
    if (cell >= 0) { atomicAdd(counts + cell, mult); nh += mult; }
    else if (cell == -2) atomicAdd(diag, (unsigned long long)mult);
  }
};

template <bool PBC, bool AP>
__global__ void __launch_bounds__(128)
xtc_bin_kernel(const uint8_t *__restrict__ buf, const xtcdev::FrameDev *__restrict__ meta, long nframes, int natoms,
               const xtcdev::Checkpoint *__restrict__ ckpt, const int *__restrict__ nseg, const int *__restrict__ status,
               GridParams P, uint32_t *__restrict__ counts, int g0, int step, int gnx, const unsigned *__restrict__ rank,
               const float *__restrict__ box3, unsigned long long *__restrict__ hits, unsigned long long *__restrict__ diag)
{
  const int msegs = xtcdev::max_segments(natoms);
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long f = t / msegs;
  const int sg = (int)(t - f * msegs);
  if (f >= nframes || status[f] != 0 || sg >= nseg[f]) return;      // (a damaged frame bins nothing: the call fails)
  const xtcdev::FrameDev m = meta[f];
  XtcBinEmit<PBC, AP> emit{P, counts, rank, diag, 0x7fffffff, step, g0 + gnx * step, 0.f, 0.f, 0.f, 0u};
  if (PBC) { emit.L0 = __ldg(box3 + 3 * f); emit.L1 = __ldg(box3 + 3 * f + 1); emit.L2 = __ldg(box3 + 3 * f + 2); }
  int a0 = 0, a1 = natoms;
  xtcdev::Checkpoint cp{0u, 0u, 0, 0};
  if (natoms > 9) {
    cp = ckpt[(size_t)f * msegs + sg];
    a0 = (int)cp.atom;
    a1 = sg + 1 < nseg[f] ? (int)ckpt[(size_t)f * msegs + sg + 1].atom : natoms;
  }
  if (AP) {
    const int k = a0 <= g0 ? 0 : (a0 - g0 + step - 1) / step;
    if (k < gnx) emit.next_sel = g0 + k * step;
  }
  if (natoms <= 9) {                 // tiny frames are stored as plain big-endian floats
    const uint32_t *w = reinterpret_cast<const uint32_t *>(buf + m.offset);
    for (int i = 0; i < natoms; i++)
      emit(i, __uint_as_float(__byte_perm(__ldg(w + 3 * i), 0, 0x0123)), __uint_as_float(__byte_perm(__ldg(w + 3 * i + 1), 0, 0x0123)),
           __uint_as_float(__byte_perm(__ldg(w + 3 * i + 2), 0, 0x0123)));
  } else {
    xtcdev::decode_segment(buf, m, cp, a1, emit);
  }
  if (emit.nh) atomicAdd(hits + f, (unsigned long long)emit.nh);
}

// ---------------------------------------------------------------------------
// K1b  streaming kernel.  Persistent, warp-specialised CTAs: one producer warp
// pulls fixed-size tiles of the raw frame array into a shared-memory ring with
// TMA bulk copies (cp.async.bulk + mbarrier, full/empty barrier pairs, no
// __syncthreads in the loop); the consumer warps (four, eight atoms per thread) pick their selected atoms
// out of the tile through the sorted index group, convert to bins and fire
// one RED per hit.  HBM sees only full-line streaming reads (evict-first);
// the gather happens in shared memory (stride 3 or 9 words: conflict free).
//
// Tiles are TILE_ATOMS atoms of the FLAT array (frame boundaries fall inside
// tiles); TILE_ATOMS*3 floats is a multiple of both 3 and 4, so atoms never
// straddle tiles and every tile starts 16-byte aligned.
// ---------------------------------------------------------------------------
#ifndef GCB_CONSUMER_WARPS
#define GCB_CONSUMER_WARPS 4
#endif
#ifndef GCB_STAGES
#define GCB_STAGES 3
#endif
#ifndef GCB_UNROLL
#define GCB_UNROLL 8
#endif
#ifndef GCB_PRIV_WARPS
#define GCB_PRIV_WARPS 16
#endif
#ifndef GCB_PRIV_UNROLL
#define GCB_PRIV_UNROLL 2
#endif
constexpr int CONSUMER_WARPS = GCB_CONSUMER_WARPS;
constexpr int CONSUMER_THREADS = CONSUMER_WARPS * 32;
constexpr int STREAM_THREADS = CONSUMER_THREADS + 32;     // + producer warp
#ifndef GCB_TILE_ATOMS
#define GCB_TILE_ATOMS 3072
#endif
constexpr int TILE_ATOMS = GCB_TILE_ATOMS;                // 3072: 1024 selected per tile for 3-site water
static_assert(TILE_ATOMS % 4 == 0, "tiles must start 16-byte aligned");
constexpr int TILE_FLOATS = TILE_ATOMS * 3;                  // 36 KiB per tile
constexpr int STAGES = GCB_STAGES;
constexpr int UNROLL = GCB_UNROLL;

struct TileMeta {
  long long f0;          // frame of the tile's first atom
  int r0;                // atom index (within frame f0) of the tile's first atom
  unsigned rel0;         // rank[r0]: index-group position of the first selected atom, relative to frame f0
  unsigned nsel;         // selected atom-frames inside the tile
  // arithmetic-progression groups: selected slot j sits at tile atom  (j < nA ? baseA : baseB) + j*step
  unsigned nA;           // selected atoms of frame f0 inside the tile (the rest belong to frame f0+1)
  int baseA, baseB;
};

// Index group of the form g0 + i*step, i < gnx (every water oxygen, all atoms, ...): the selected
// atoms of a tile follow from arithmetic, no index or rank loads in the inner loop.
struct ApGroup { int g0, step, tile_atoms; };

template <int NST, int TCAP_ATOMS = TILE_ATOMS>
struct StreamSmem {
  float tile[NST][TCAP_ATOMS * 3];
  unsigned long long full[NST], empty[NST];
  TileMeta meta[NST];
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// TMA 1D bulk copy global -> shared, completion signalled on the mbarrier; frames are read once: evict-first in L2
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar,
                                            unsigned long long policy)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// Fast float -> bin for the streaming kernel.  Same result as cell_of() for every input, with no
// division in the common case.  The real quotient R = u/delta is bracketed by two products,
//   p_lo = u * fl((1/delta)(1 - 2^-21))  <  R (1 - 2^-24)   and   p_hi = u * fl((1/delta)(1 + 2^-21))  >  R (1 + 2^-24)
// (three f32 roundings of 2^-24 each cannot eat the 2^-21 margin), and the reference's IEEE quotient
// fl(R) lies within 2^-24 of R, hence strictly between them.  floor() of each product is read from
// the mantissa after a round-DOWN addition of 1.5*2^23.  When both floors agree (`sure`), that is
// floor(fl(R)).  Otherwise -- an atom within a few ulp of a cell boundary, ~1e-4 of them -- the
// caller redoes the atom with cell_of_exact().  The host sets the reciprocals to NaN (nothing is
// ever `sure`) for geometries where an in-bounds atom could bin to mx (gcb_edge_vulnerable) or with
// more than 2^21 cells along a dimension, so the fast path needs no range check on the indices.
struct GridParamsFast {
  float a[3], b[3], delta[3], inv_lo[3], inv_hi[3], fmx[3];
  int my, mz;
  int bias;                                          // 0x4B400000 * (my*mz + mz + 1), mod 2^32
};

struct FastCell { int cell; bool inside, sure; };

__device__ __forceinline__ FastCell cell_of_fast(const GridParamsFast &P, float x, float y, float z)
{
  FastCell o;
  // x - a < 0 <=> x < a and x - b >= 0 <=> x >= b: a difference of two finite floats is zero
  // only when they are equal (subnormals are kept, -ftz=false), and its sign is the comparison's
  o.inside = !((x < P.a[0]) | (x >= P.b[0]) | (y < P.a[1]) | (y >= P.b[1]) | (z < P.a[2]) | (z >= P.b[2]));
  const float MAGIC = 12582912.0f;                   // 1.5 * 2^23
  const float ux = __fsub_rn(x, P.a[0]), uy = __fsub_rn(y, P.a[1]), uz = __fsub_rn(z, P.a[2]);
  const float lx = __fadd_rd(__fmul_rn(ux, P.inv_lo[0]), MAGIC), hx = __fadd_rd(__fmul_rn(ux, P.inv_hi[0]), MAGIC);
  const float ly = __fadd_rd(__fmul_rn(uy, P.inv_lo[1]), MAGIC), hy = __fadd_rd(__fmul_rn(uy, P.inv_hi[1]), MAGIC);
  const float lz = __fadd_rd(__fmul_rn(uz, P.inv_lo[2]), MAGIC), hz = __fadd_rd(__fmul_rn(uz, P.inv_hi[2]), MAGIC);
  o.sure = (lx == hx) & (ly == hy) & (lz == hz);
  o.cell = (__float_as_int(lx) * P.my + __float_as_int(ly)) * P.mz + __float_as_int(lz) - P.bias;
  return o;
}

// the reference's own operations (g_ri3Dc.c:145-147) for an atom already known to be inside [a, b).
// Out of line and fed scalars, so the rare path costs the hot loop neither registers nor a stack copy.
__device__ __noinline__ int cell_exact_scalar(float x, float y, float z, float a0, float a1, float a2, float d0,
                                              float d1, float d2, float f0, float f1, float f2, int my, int mz)
{
  const float qx = __fdiv_rn(__fsub_rn(x, a0), d0);
  const float qy = __fdiv_rn(__fsub_rn(y, a1), d1);
  const float qz = __fdiv_rn(__fsub_rn(z, a2), d2);
  if (!(qx < f0) | !(qy < f1) | !(qz < f2)) return -2;      // >= mx, or NaN
  return (__float2int_rd(qx) * my + __float2int_rd(qy)) * mz + __float2int_rd(qz);
}
__device__ __forceinline__ int cell_of_exact(const GridParamsFast &P, float x, float y, float z)
{
  return cell_exact_scalar(x, y, z, P.a[0], P.a[1], P.a[2], P.delta[0], P.delta[1], P.delta[2], P.fmx[0], P.fmx[1],
                           P.fmx[2], P.my, P.mz);
}

// what the consumers need to know about tile t of the flat frame array (the producer publishes it with the tile)
template <bool AP>
__device__ __forceinline__ TileMeta make_tile_meta(long long t, int tile_atoms, int natoms, unsigned gnx,
                                                   const unsigned *__restrict__ rank, const ApGroup &G)
{
  const long long g0 = t * tile_atoms, g1 = g0 + tile_atoms;
  const long long f0 = g0 / natoms, f1 = g1 / natoms;
  const int r0 = (int)(g0 - f0 * natoms), r1 = (int)(g1 - f1 * natoms);
  TileMeta m;
  m.f0 = f0; m.r0 = r0;
  if (AP) {
    // frame f0 contributes atoms [r0, endA), frame f0+1 atoms [0, endB)   (tile_atoms <= natoms)
    const int endA = min(natoms, r0 + tile_atoms), endB = r0 + tile_atoms - natoms;
    const int iA0 = r0 <= G.g0 ? 0 : min((int)gnx, (r0 - G.g0 + G.step - 1) / G.step);
    const int iA1 = endA <= G.g0 ? 0 : min((int)gnx, (endA - G.g0 + G.step - 1) / G.step);
    const int nB = endB <= G.g0 ? 0 : min((int)gnx, (endB - G.g0 + G.step - 1) / G.step);
    m.nA = (unsigned)(iA1 - iA0);
    m.nsel = m.nA + (unsigned)nB;
    m.baseA = G.g0 + iA0 * G.step - r0;
    m.baseB = G.g0 + (natoms - r0) - (int)m.nA * G.step;
    m.rel0 = 0;
  } else {
    const unsigned rel0 = __ldg(rank + r0);
    m.rel0 = rel0;
    m.nsel = (unsigned)((f1 - f0) * (long long)gnx + (long long)__ldg(rank + r1) - (long long)rel0);
    m.nA = 0; m.baseA = 0; m.baseB = 0;
  }
  return m;
}

// ---------------------------------------------------------------------------
// Where a hit goes.
//
// DirectSink: one RED per hit into the global uint32 grid (L2-resident when the grid is small).  On B200
// that is bounded by what the L2 sustains for the mix "coordinate stream + one scattered RED per unit"
// (gcb_selftest_l2 measures it: ~120 G units/s for one atom in three, ~175 for every atom), not by HBM.
// PackedSink: the same for grids whose uint32 cells exceed the L2 (below).
//
// Tried and removed (numbers in DESIGN.md section 3): routing hits as keys to shared-memory sized buckets or
// L2-sized slabs and counting them in a second kernel (round 1: 92 G/s against 111 on the nanopore shape,
// 72 against 82+ on 400^3), and packed counters with returning atomics (37 G/s on 400^3).
// ---------------------------------------------------------------------------
// Every sink also fixes the shape of the streaming kernel it is compiled into: consumer warps per CTA, atoms per
// thread and loop step, tile capacity and ring depth.
template <int MODE>
struct DirectSink {
  static constexpr int NSTAGES = STAGES, CWARPS = CONSUMER_WARPS, UNROLL_ = UNROLL, TILE_CAP = TILE_ATOMS;
  static constexpr bool PRIVATE = false;
  uint32_t *counts; float *acc; float w;
  __host__ __device__ __forceinline__ size_t extra_smem() const { return 0; }
  // `smem` is the sink's share of the dynamic shared memory (unused here)
  __device__ __forceinline__ void init(void *, unsigned) const {}
  __device__ __forceinline__ void hit(void *, int cell) const { deposit<MODE>(counts, acc, cell, w); }
  __device__ __forceinline__ void tile_hits(void *, unsigned) const {}
  __device__ __forceinline__ bool finish(void *, unsigned) const { return false; }
  __device__ __forceinline__ void redo_hit(int) const {}
};

// PackedSink: L2-resident atomics for grids whose uint32 counters do not fit the L2 but whose packed counters do
// (BASELINE config 3: 400^3 cells = 256 MB of uint32, 64 MB as bytes; config 4: 500^3 = 500 MB, 62.5 MB as nibbles).
// 32/W counters of W bits share a word; a hit is ONE NON-RETURNING add (RED) of 1 << W*(cell mod 32/W) on the word,
// so it runs at the L2's RED rate and nothing comes back to the SM.  Overflow of a W-bit lane is not looked for per
// hit but once per round: every carry out of a lane lowers the word's digit sum (base 2^W) by at least 2^W - 1 and
// nothing can raise it, so "sum of all digits == hits the round reported" holds iff no lane overflowed and no RED
// was lost (packed_verify_kernel).  A round that fails is thrown away and redone by the direct uint32 kernel.
template <int W>
struct PackedSink {
  static constexpr int NSTAGES = STAGES, CWARPS = CONSUMER_WARPS, UNROLL_ = UNROLL, TILE_CAP = TILE_ATOMS;
  static constexpr bool PRIVATE = false;
  uint32_t *pk;                            // packed counters, all-zero between rounds
  __host__ __device__ __forceinline__ size_t extra_smem() const { return 0; }
  __device__ __forceinline__ void init(void *, unsigned) const {}
  __device__ __forceinline__ void hit(void *, int cell) const
  {
    constexpr unsigned PER = 32u / W, SH = W == 8 ? 2u : 3u;
    atomicAdd(pk + ((unsigned)cell >> SH), 1u << ((unsigned)W * ((unsigned)cell & (PER - 1u))));      // RED.E.ADD
  }
  __device__ __forceinline__ void tile_hits(void *, unsigned) const {}
  __device__ __forceinline__ bool finish(void *, unsigned) const { return false; }
  __device__ __forceinline__ void redo_hit(int) const {}
};

// PrivateSink: the shared-memory privatisation for grids that fit one CTA's shared memory as W-bit counters next
// to the tile ring: 8 bits up to ~118 000 cells (a typical pore analysis: 40 x 40 x 60), 4 bits up to ~236 000 (the
// 60^3 grid of BASELINE config 1).  Every CTA (one per SM) keeps its OWN copy of the whole grid in shared memory; a
// hit is one non-returning shared-memory atomic, and the L2 sees nothing but the coordinate stream.  Between
// launches the private grids rest in global memory (loaded at the start of a launch, stored at its end: 2 x ~110 KB
// per CTA, L2-resident), and they are folded into the uint32 grid only when somebody reads it or when a lane could
// fill up (flush_private): one streaming pass over the copies per ~32 (8-bit) or ~1.5 (4-bit) hits per private cell
// instead of one L2 atomic per hit.
// Exactness does not rest on the density being tame: at the end of a launch the CTA compares the digit sum of its
// private grid with the hits it holds (every carry out of a lane lowers the digit sum); on a mismatch it leaves the
// stored copy as it was, bins its tiles of this launch again with plain uint32 REDs (hits and diagnostics were
// already counted), and raises a flag that sends the next launches straight to the uint32 grid (back-off).
template <int W> __device__ __forceinline__ unsigned digit_sum(uint32_t w);

struct PrivState {
  int direct;                              // this launch bins into the uint32 grid (backing off)
  int skip, backoff, failed;
  unsigned long long launches, redone;
};

template <int W>
struct PrivateSink {
  // full-size tiles (1024 selected atoms of 3-site water per tile, two per thread of 16 consumer warps), 110 KB of ring
  static constexpr int NSTAGES = 3, CWARPS = GCB_PRIV_WARPS, UNROLL_ = GCB_PRIV_UNROLL, TILE_CAP = TILE_ATOMS;
  static constexpr bool PRIVATE = true;
  static constexpr unsigned PER = 32u / W, SH = W == 8 ? 2u : 3u;
  uint32_t *counts;                        // the uint32 grid (direct mode, redo)
  uint32_t *stage;                         // [CTA][nwords] the private grids between launches
  unsigned long long *stage_hits;          // [CTA] hits held by stage[CTA]
  PrivState *st;
  unsigned nwords;
  struct Head { unsigned long long hits; int direct, redo; };        // in front of the private grid in shared memory
  __host__ __device__ __forceinline__ size_t extra_smem() const { return sizeof(Head) + (size_t)nwords * 4; }
  static __device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(CWARPS * 32) : "memory"); }
  __device__ __forceinline__ void init(void *smem, unsigned ctid) const
  {
    Head *h = static_cast<Head *>(smem);
    uint32_t *g = reinterpret_cast<uint32_t *>(h + 1);
    const int direct = *(volatile int *)&st->direct;
    if (ctid == 0) { h->hits = direct ? 0ull : stage_hits[blockIdx.x]; h->direct = direct; h->redo = 0; }
    if (!direct) {
      const uint32_t *src = stage + (size_t)blockIdx.x * nwords;
      for (unsigned i = ctid; i < nwords; i += CWARPS * 32) g[i] = src[i];
    }
    consumer_barrier();
  }
  __device__ __forceinline__ void hit(void *smem, int cell) const
  {
    const Head *h = static_cast<const Head *>(smem);
    if (h->direct) { atomicAdd(counts + cell, 1u); return; }
    uint32_t *g = reinterpret_cast<uint32_t *>(const_cast<Head *>(h) + 1);
    atomicAdd(g + ((unsigned)cell >> SH), 1u << ((unsigned)W * ((unsigned)cell & (PER - 1u))));   // shared-memory atomic, no return value
  }
  __device__ __forceinline__ void tile_hits(void *smem, unsigned n) const    // lane 0 of a warp: hits of its share of a tile
  {
    Head *h = static_cast<Head *>(smem);
    if (n && !h->direct) atomicAdd(&h->hits, (unsigned long long)n);
  }
  // -> true when this CTA has to bin its tiles again into the uint32 grid (uniform over the CTA's consumers)
  __device__ __forceinline__ bool finish(void *smem, unsigned ctid) const
  {
    Head *h = static_cast<Head *>(smem);
    if (h->direct) return false;
    uint32_t *g = reinterpret_cast<uint32_t *>(h + 1);
    __shared__ unsigned long long digits;
    if (ctid == 0) digits = 0;
    consumer_barrier();
    unsigned long long d = 0;
    for (unsigned i = ctid; i < nwords; i += CWARPS * 32) {
      d += digit_sum<W>(g[i]);
    }
    for (int o = 16; o > 0; o >>= 1) d += __shfl_down_sync(0xffffffffu, d, o);
    if ((ctid & 31u) == 0 && d) atomicAdd(&digits, d);
    consumer_barrier();
    const bool ok = digits == h->hits;
    if (ok) {
      uint32_t *dst = stage + (size_t)blockIdx.x * nwords;
      for (unsigned i = ctid; i < nwords; i += CWARPS * 32) dst[i] = g[i];
      if (ctid == 0) stage_hits[blockIdx.x] = h->hits;
    } else if (ctid == 0) {
      atomicExch(&st->failed, 1);
    }
    return !ok;
  }
  __device__ __forceinline__ void redo_hit(int cell) const { atomicAdd(counts + cell, 1u); }
};

// device-resident state of the packed path: the host queues the same kernels every round, they decide here
struct PackedState {
  int alarm;                               // this round is binned by the direct uint32 kernel
  int dirty;                               // the packed pass of this round ran and failed: its side effects are taken back
  int skip;                                // rounds still to be binned directly (back-off after a failed round)
  int backoff;
  unsigned ticket, pad;
  unsigned long long digits, expect;       // digit sum of the packed counters / hits the packed pass reported
  unsigned long long diag_snap[2];
  unsigned long long rounds, redone;       // statistics: rounds queued, rounds that the direct kernel had to bin
};

// GUARD: 0 = always run; 1 = run only if *run_if != 0; 2 = run only if *run_if == 0 (the two passes of a packed
// round, decided on the device; separate instantiations, because the test costs the plain kernel a spill)
template <class SINK, bool PBC, bool AP, int GUARD = 0>
__global__ void __launch_bounds__((SINK::CWARPS + 1) * 32)
bin_stream_kernel(const float *__restrict__ x, int natoms, const int *__restrict__ gindex, unsigned gnx,
                  const unsigned *__restrict__ rank, long long ntiles, ApGroup G, GridParamsFast P, SINK sink,
                  const float *__restrict__ box3, unsigned long long *__restrict__ hits,
                  unsigned long long *__restrict__ diag, const int *__restrict__ run_if)
{
  constexpr int NST = SINK::NSTAGES, CTHREADS = SINK::CWARPS * 32, UNR = SINK::UNROLL_;
  using Smem = StreamSmem<NST, SINK::TILE_CAP>;
  if constexpr (GUARD == 1) { if (*run_if == 0) return; }
  if constexpr (GUARD == 2) { if (*run_if != 0) return; }
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem &S = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x;

  if (tid == 0) {
    for (int s = 0; s < NST; s++) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], SINK::CWARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid < 32) {
    // ---------------- producer warp (one lane) ----------------
    if (tid == 0) {
      unsigned long long policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      int it = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int stage = it % NST;
        if (it >= NST) mbar_wait(&S.empty[stage], (unsigned)(it / NST - 1) & 1u);
        const int tile_atoms = AP ? G.tile_atoms : TILE_ATOMS;
        const TileMeta m = make_tile_meta<AP>(t, tile_atoms, natoms, gnx, rank, G);
        S.meta[stage] = m;                         // published by the mbarrier's release/acquire
        const unsigned bytes = (unsigned)tile_atoms * 12u;
        mbar_expect_tx(&S.full[stage], bytes);
        tma_load_1d(S.tile[stage], x + (size_t)t * ((size_t)tile_atoms * 3), bytes, &S.full[stage], policy);
      }
    }
    return;
  }

  // ---------------- consumer warps ----------------
  const unsigned ctid = (unsigned)tid - 32u;
  const int lane = tid & 31;
  void *const sink_smem = smem_raw + ((sizeof(Smem) + 15) & ~(size_t)15);
  sink.init(sink_smem, ctid);
  int it = 0;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int stage = it % NST;
    mbar_wait(&S.full[stage], (unsigned)(it / NST) & 1u);
    const TileMeta m = S.meta[stage];
    const uint32_t tile = smem_addr(S.tile[stage]);
    unsigned c0 = 0, c1 = 0;            // hits in frame f0 / f0+1 (later frames go straight to global)
    if (AP) {
      const uint32_t addrA = tile + 12u * (uint32_t)m.baseA, addrB = tile + 12u * (uint32_t)m.baseB;
      const uint32_t step12 = 12u * (uint32_t)G.step;
      for (unsigned base = ctid; base < m.nsel; base += CTHREADS * UNR) {
        float px[UNR], py[UNR], pz[UNR];
        bool valid[UNR], inB[UNR];
#pragma unroll
        for (int k = 0; k < UNR; k++) {
          const unsigned j = base + k * CTHREADS;
          valid[k] = j < m.nsel;
          inB[k] = j >= m.nA;
          const uint32_t p = valid[k] ? (inB[k] ? addrB : addrA) + j * step12 : tile;
          px[k] = lds_f32(p); py[k] = lds_f32(p + 4); pz[k] = lds_f32(p + 8);
        }
        unsigned rare = 0;
#pragma unroll
        for (int k = 0; k < UNR; k++) {
          if (PBC) {
            const float *L = box3 + 3 * (m.f0 + (inB[k] ? 1 : 0));
            px[k] = wrap1(px[k], __ldg(L)); py[k] = wrap1(py[k], __ldg(L + 1)); pz[k] = wrap1(pz[k], __ldg(L + 2));
          }
          const FastCell fc = cell_of_fast(P, px[k], py[k], pz[k]);
          const bool go = valid[k] & fc.inside & fc.sure;
          if (go) sink.hit(sink_smem, fc.cell);
          c0 += (go & !inB[k]) ? 1u : 0u;
          c1 += (go & inB[k]) ? 1u : 0u;
          rare |= (valid[k] & fc.inside & !fc.sure) ? (1u << k) : 0u;
        }
        if (rare) {                                  // atoms within a few ulp of a cell boundary
#pragma unroll
          for (int k = 0; k < UNR; k++) {
            if (!((rare >> k) & 1u)) continue;
            const int cell = cell_of_exact(P, px[k], py[k], pz[k]);
            if (cell >= 0) {
              sink.hit(sink_smem, cell);
              if (inB[k]) c1++; else c0++;
            } else {
              atomicAdd(diag, 1ull);
            }
          }
        }
      }
    } else
    for (unsigned base = 0; base < m.nsel; base += CTHREADS * UNR) {
      int local[UNR];
      unsigned df[UNR];
      bool valid[UNR];
#pragma unroll
      for (int k = 0; k < UNR; k++) {
        const unsigned s = base + k * CTHREADS + ctid;
        valid[k] = s < m.nsel;
        unsigned rel = m.rel0 + s;
        df[k] = 0;
        if (rel >= gnx) { df[k] = rel / gnx; rel -= df[k] * gnx; }
        const int idx = valid[k] ? __ldg(gindex + rel) : m.r0;
        local[k] = (int)((long long)df[k] * natoms + idx - m.r0);
      }
      float px[UNR], py[UNR], pz[UNR];
#pragma unroll
      for (int k = 0; k < UNR; k++) {
        const uint32_t p = tile + 12u * (uint32_t)(valid[k] ? local[k] : 0);
        px[k] = lds_f32(p); py[k] = lds_f32(p + 4); pz[k] = lds_f32(p + 8);
      }
#pragma unroll
      for (int k = 0; k < UNR; k++) {
        if (!valid[k]) continue;
        if (PBC) {
          const float *L = box3 + 3 * (m.f0 + df[k]);
          px[k] = wrap1(px[k], __ldg(L)); py[k] = wrap1(py[k], __ldg(L + 1)); pz[k] = wrap1(pz[k], __ldg(L + 2));
        }
        const FastCell fc = cell_of_fast(P, px[k], py[k], pz[k]);
        const int cell = !fc.inside ? -1 : fc.sure ? fc.cell : cell_of_exact(P, px[k], py[k], pz[k]);
        if (cell >= 0) {
          sink.hit(sink_smem, cell);
          if (df[k] == 0) c0++;
          else if (df[k] == 1) c1++;
          else atomicAdd(hits + m.f0 + df[k], 1ull);
        } else if (cell == -2) {
          atomicAdd(diag, 1ull);
        }
      }
    }
    // Per-frame hit counters: one fire-and-forget global atomic per (warp, frame).  (Adding the warps of a CTA up in
    // shared memory first, so that a large frame's counter sees one atomic per tile instead of eight, was measured:
    // the returning shared-memory atomic and fence it needs sit on every tile's critical path and HALVE the kernel,
    // 52.7 vs 101 G/s on C3.)
    c0 = __reduce_add_sync(0xffffffffu, (c0 & 0xffffu) | (c1 << 16));          // both sums in one reduction (<= 32 * UNR each)
    if (lane == 0) {
      c1 = c0 >> 16; c0 &= 0xffffu;
      if (c0) atomicAdd(hits + m.f0, (unsigned long long)c0);
      if (c1) atomicAdd(hits + m.f0 + 1, (unsigned long long)c1);
      sink.tile_hits(sink_smem, c0 + c1);
      mbar_arrive(&S.empty[stage]);     // this warp is done with the stage (its LDS results are in registers)
    }
  }
  const bool redo = sink.finish(sink_smem, ctid);
  if constexpr (SINK::PRIVATE && AP) {
    // the private grid of this CTA overflowed a lane (atoms piled onto a few cells): its tiles again, straight from
    // global memory into the uint32 grid; hits and diagnostics were counted by the pass above
    if (redo) {
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const TileMeta m = make_tile_meta<true>(t, G.tile_atoms, natoms, gnx, rank, G);
        const float *tile = x + (size_t)t * ((size_t)G.tile_atoms * 3);
        for (unsigned j = ctid; j < m.nsel; j += CTHREADS) {
          const bool inB = j >= m.nA;
          const float *p = tile + 3 * ((long long)(inB ? m.baseB : m.baseA) + (long long)j * G.step);
          float px = __ldg(p), py = __ldg(p + 1), pz = __ldg(p + 2);
          if (PBC) {
            const float *L = box3 + 3 * (m.f0 + (inB ? 1 : 0));
            px = wrap1(px, __ldg(L)); py = wrap1(py, __ldg(L + 1)); pz = wrap1(pz, __ldg(L + 2));
          }
          const FastCell fc = cell_of_fast(P, px, py, pz);
          if (!fc.inside) continue;
          const int cell = fc.sure ? fc.cell : cell_of_exact(P, px, py, pz);
          if (cell >= 0) sink.redo_hit(cell);
        }
      }
    }
  }
}

}  // namespace k1
