// K1: per-frame binning of index-group atom positions into the occupancy grid
// (replaces gridcount(), g_ri3Dc.c:139-152, and the atom loop g_ri3Dc.c:425-426),
// K2: density normalisation (g_ri3Dc.c:451-460) + file-order transpose
//     (grid3_serialise, xdr_grid.c:113-126),
// K4: counter-based synthetic frames, and the gcb_counter host object that
//     stages frames through pinned buffers on CUDA streams.
//
// sm_100a only.  Built with -fmad=false; the float->bin conversion uses
// explicit round-to-nearest intrinsics (__fsub_rn, __fdiv_rn) in the
// reference's operation order so that bin indices match bit for bit.
//
// Data layout in HBM
//   frames   x[frame][atom][3] f32, exactly the host rvec layout (12 B/atom)
//   counts   u32 [mx][my][mz], z fastest (the reference's contiguous block)
//   acc      f32 same shape, only materialised when frame weights change
//   gindex   i32 sorted copy of the index group, rank[] = prefix count per atom
#include "gcb_internal.h"
#include "xtc_device.cuh"
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <new>
#include <vector>

namespace {

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



// ---------------------------------------------------------------------------
// n sequential f32 additions of w to v (see gcb::seq_add_f32)
// ---------------------------------------------------------------------------
__device__ float seq_add_dev(float v, float w, unsigned long long n)
{
  // From zero, with a weight whose significand is short (1, 2, 0.5, 0.25 ...: m * 2^e with n * m < 2^24), every
  // partial sum j * w is exactly representable, so the chain is the exact product (the common weight-1 case).
  if (v == 0.0f && w > 0.0f && n > 0 && n < (1ull << 24) && (__float_as_uint(w) >> 23) != 0u) {
    unsigned m = (__float_as_uint(w) & 0x7fffffu) | 0x800000u;
    m >>= __ffs((int)m) - 1;
    const float p = __fmul_rn((float)n, w);
    if (n * (unsigned long long)m < (1ull << 24) && p < 3.0e38f) return p;
  }
  while (n > 0) {
    const float v1 = __fadd_rn(v, w);
    --n;
    if (v1 == v || v1 != v1) return v1;
    v = v1;
    if (n < 2) continue;
    const float v2 = __fadd_rn(v, w), v3 = __fadd_rn(v2, w);
    unsigned b1 = __float_as_uint(v), b2 = __float_as_uint(v2), b3 = __float_as_uint(v3);
    if (v > 0.0f && w > 0.0f && (b1 >> 23) == (b2 >> 23) && (b2 >> 23) == (b3 >> 23) && b3 > b2) {
      const unsigned inc = b3 - b2, last = b2 | 0x7FFFFFu;
      --n;
      unsigned long long k = (unsigned long long)(last - b2) / inc;
      if (k > n) k = n;
      b2 += (unsigned)(k * inc);
      n -= k;
      v = __uint_as_float(b2);
    }
  }
  return v;
}

// ---------------------------------------------------------------------------
// The three small kernels around the packed pass of a round (see PackedSink).
// ---------------------------------------------------------------------------
__global__ void packed_begin_kernel(PackedState *__restrict__ st, const unsigned long long *__restrict__ diag)
{
  st->alarm = st->skip > 0 ? 1 : 0;         // backing off: the packed pass returns at once, the direct kernel bins the round
  st->dirty = 0;
  st->ticket = 0;
  st->digits = 0; st->expect = 0;
  st->diag_snap[0] = diag[0]; st->diag_snap[1] = diag[1];
  st->rounds++;
}

template <int W>
__device__ __forceinline__ unsigned digit_sum(uint32_t w)
{
  if (W == 8) return __dp4a(w, 0x01010101u, 0u);
  return __dp4a(w & 0x0F0F0F0Fu, 0x01010101u, __dp4a((w >> 4) & 0x0F0F0F0Fu, 0x01010101u, 0u));
}

// digit sum of all packed counters against the hits the packed pass reported for the round's frames; the last
// block to finish writes the verdict and runs the back-off (a failed round is followed by `backoff` rounds that
// go straight to the direct kernel; the back-off doubles on every failure and halves on every success)
template <int W>
__global__ void __launch_bounds__(256)
packed_verify_kernel(const uint4 *__restrict__ pk, size_t nquads, const unsigned long long *__restrict__ hits, long nhits,
                     PackedState *__restrict__ st)
{
  __shared__ unsigned long long red[2][8];
  if (st->alarm) {                                   // skipped round
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->skip--; st->redone++; }
    return;
  }
  unsigned long long dsum = 0, hsum = 0;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = pk[q];
    dsum += digit_sum<W>(v.x) + digit_sum<W>(v.y) + digit_sum<W>(v.z) + digit_sum<W>(v.w);
  }
  for (long f = (long)blockIdx.x * blockDim.x + threadIdx.x; f < nhits; f += (long)gridDim.x * blockDim.x) hsum += hits[f];
  for (int o = 16; o > 0; o >>= 1) { dsum += __shfl_down_sync(0xffffffffu, dsum, o); hsum += __shfl_down_sync(0xffffffffu, hsum, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = dsum; red[1][threadIdx.x >> 5] = hsum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) { dsum += red[0][w]; hsum += red[1][w]; }
    if (dsum) atomicAdd(&st->digits, dsum);
    if (hsum) atomicAdd(&st->expect, hsum);
    __threadfence();
    if (atomicAdd(&st->ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      const unsigned long long d = *(volatile unsigned long long *)&st->digits, e = *(volatile unsigned long long *)&st->expect;
      if (d != e) {
        st->alarm = 1; st->dirty = 1; st->redone++;
        st->backoff = st->backoff < 1 ? 1 : (st->backoff >= 32 ? 64 : 2 * st->backoff);
        st->skip = st->backoff;
      } else {
        st->backoff >>= 1;
      }
    }
  }
}

// End of the packed pass.  Verified: the packed counters are added to the uint32 grid (every word is owned by one
// thread, nothing else touches the grid meanwhile) and zeroed.  Failed: they are zeroed only, and what the pass
// added to the per-frame hit counters and the diagnostics is taken back, because the guarded uint32 kernel behind
// this one redoes the round.  Skipped round: nothing to do.
template <int W>
__global__ void __launch_bounds__(256)
packed_apply_kernel(uint32_t *__restrict__ pk, size_t nwords, uint32_t *__restrict__ counts, size_t ncells,
                    const PackedState *__restrict__ st, unsigned long long *__restrict__ hits, long nhits,
                    unsigned long long *__restrict__ diag)
{
  const bool bad = st->alarm != 0;
  if (bad && !st->dirty) return;
  constexpr int PER = 32 / W;                      // cells per word
  constexpr uint32_t M = (1u << W) - 1u;
  // one thread per packed word: a warp reads 128 contiguous bytes of packed counters and updates 32 * PER
  // contiguous uint32 cells (16 or 32 bytes per lane)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t w = pk[i];
    if (!w) continue;
    pk[i] = 0u;
    if (bad) continue;
    const size_t c0 = i * PER;
    if (c0 + PER <= ncells) {
#pragma unroll
      for (int h = 0; h < PER / 4; h++) {
        const uint32_t part = w >> (4 * W * h);
        if (W == 4 && !(part & 0xFFFFu)) continue;
        uint4 *dst = reinterpret_cast<uint4 *>(counts + c0) + h;
        uint4 t = *dst;
        t.x += part & M; t.y += (part >> W) & M; t.z += (part >> (2 * W)) & M; t.w += (part >> (3 * W)) & M;
        *dst = t;
      }
    } else {
      for (int b = 0; b < PER; b++) if (c0 + b < ncells) counts[c0 + b] += (w >> (W * b)) & M;
    }
  }
  if (bad) {
    for (long f = (long)blockIdx.x * blockDim.x + threadIdx.x; f < nhits; f += (long)gridDim.x * blockDim.x) hits[f] = 0;
    if (blockIdx.x == 0 && threadIdx.x < 2) diag[threadIdx.x] = st->diag_snap[threadIdx.x];
  }
}

// The 4-bit rounds of a grid like 500^3 (125 M cells) would each read and write the whole 500 MB uint32 grid to add
// at most 15 per cell.  They go through a BYTE grid instead (a quarter of the traffic): verified nibbles are added
// to bytes for up to 16 rounds (16 x 15 <= 255: no byte can overflow), then the bytes are folded into the uint32
// grid in one pass (flush_mid; also before anything reads the grid).  Same contract as packed_apply_kernel otherwise.
__global__ void __launch_bounds__(256)
packed_apply_mid_kernel(uint32_t *__restrict__ pk, size_t nwords, uint2 *__restrict__ mid, const PackedState *__restrict__ st,
                        unsigned long long *__restrict__ hits, long nhits, unsigned long long *__restrict__ diag)
{
  const bool bad = st->alarm != 0;
  if (bad && !st->dirty) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t w = pk[i];
    if (!w) continue;
    pk[i] = 0u;
    if (bad) continue;
    // nibbles 0..3 -> the four bytes of .x, nibbles 4..7 -> .y (cells 8i .. 8i+7 are bytes 8i .. 8i+7, little endian)
    const uint32_t lo = w & 0xffffu, hi = w >> 16;
    uint2 m = mid[i];
    m.x += (lo & 0xfu) | ((lo & 0xf0u) << 4) | ((lo & 0xf00u) << 8) | ((lo & 0xf000u) << 12);
    m.y += (hi & 0xfu) | ((hi & 0xf0u) << 4) | ((hi & 0xf00u) << 8) | ((hi & 0xf000u) << 12);
    mid[i] = m;
  }
  if (bad) {
    for (long f = (long)blockIdx.x * blockDim.x + threadIdx.x; f < nhits; f += (long)gridDim.x * blockDim.x) hits[f] = 0;
    if (blockIdx.x == 0 && threadIdx.x < 2) diag[threadIdx.x] = st->diag_snap[threadIdx.x];
  }
}

__global__ void __launch_bounds__(256)
mid_flush_kernel(uint2 *__restrict__ mid, size_t nwords, uint32_t *__restrict__ counts, size_t ncells)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) {
    const uint2 m = mid[i];
    if (!(m.x | m.y)) continue;
    mid[i] = make_uint2(0u, 0u);
    const size_t c0 = i * 8;
    if (c0 + 8 <= ncells) {
      uint4 *dst = reinterpret_cast<uint4 *>(counts + c0);
      uint4 a = dst[0], b = dst[1];
      a.x += m.x & 0xffu; a.y += (m.x >> 8) & 0xffu; a.z += (m.x >> 16) & 0xffu; a.w += m.x >> 24;
      b.x += m.y & 0xffu; b.y += (m.y >> 8) & 0xffu; b.z += (m.y >> 16) & 0xffu; b.w += m.y >> 24;
      dst[0] = a; dst[1] = b;
    } else {
      for (int l = 0; l < 8; l++) if (c0 + l < ncells) counts[c0 + l] += ((l < 4 ? m.x : m.y) >> (8 * (l & 3))) & 0xffu;
    }
  }
}

// In front of every private launch: turn the verdicts of the launches before into this launch's mode.
__global__ void private_begin_kernel(PrivState *__restrict__ st)
{
  if (st->failed) {                                  // some CTA overflowed a lane in the launch before: back off
    st->backoff = st->backoff < 1 ? 1 : (st->backoff >= 32 ? 64 : 2 * st->backoff);
    st->skip = st->backoff;
    st->failed = 0;
    st->redone++;
  } else if (st->skip == 0 && st->backoff > 0) {
    st->backoff >>= 1;
  }
  st->direct = st->skip > 0 ? 1 : 0;
  if (st->skip > 0) st->skip--;
  st->launches++;
}

// Fold the private grids into the uint32 grid and clear them: one thread per word position, the CTAs' copies of a
// word are `nwords` apart (coalesced across the warp); nothing else touches the grid meanwhile (stream order).
template <int W>
__global__ void __launch_bounds__(256)
private_flush_kernel(uint32_t *__restrict__ stage, unsigned long long *__restrict__ stage_hits, unsigned nwords, int nctas,
                     uint32_t *__restrict__ counts, size_t ncells)
{
  constexpr unsigned PER = 32u / W, LANE = (1u << W) - 1u;
  constexpr int DEPTH = 8;                           // copies read per step: independent loads in flight
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x) {
    unsigned sum[PER];
#pragma unroll
    for (unsigned l = 0; l < PER; l++) sum[l] = 0;
    for (int b0 = 0; b0 < nctas; b0 += DEPTH) {
      uint32_t w[DEPTH];
#pragma unroll
      for (int q = 0; q < DEPTH; q++) w[q] = b0 + q < nctas ? stage[(size_t)(b0 + q) * nwords + i] : 0u;
#pragma unroll
      for (int q = 0; q < DEPTH; q++) {
        if (!w[q]) continue;
        stage[(size_t)(b0 + q) * nwords + i] = 0u;
#pragma unroll
        for (unsigned l = 0; l < PER; l++) sum[l] += (w[q] >> (W * l)) & LANE;
      }
    }
#pragma unroll
    for (unsigned l = 0; l < PER; l++) {
      const size_t cell = (size_t)i * PER + l;
      if (sum[l] && cell < ncells) counts[cell] += sum[l];
    }
  }
  if (blockIdx.x == 0) for (int b = threadIdx.x; b < nctas; b += blockDim.x) stage_hits[b] = 0ull;
}

// Saturation guard of the uint32 cells.  The reference accumulates in f32, which stops moving after at most
// ~2^26 additions of one weight, so a cell count clamped at 2^27 (or above) gives the same occupancy and density
// bit for bit; a WRAPPED count would not.  The host keeps an upper bound on what any cell can hold (what was
// binned since the last clamp) and runs this pass before the bound could reach 2^32: once per ~4.2e9 atom-frames.
__global__ void clamp_counts_kernel(uint32_t *__restrict__ cnt, size_t n, uint32_t limit, unsigned long long *__restrict__ flag)
{
  bool any = false;
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x)
    if (cnt[c] > limit) { cnt[c] = limit; any = true; }
  if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) atomicOr(flag, 1ull);
}

// fold a finished weight run into the f32 accumulator: acc = acc (+)= w, runcnt times
__global__ void fold_run_kernel(uint32_t *__restrict__ runcnt, uint32_t *__restrict__ total,
                                float *__restrict__ acc, float w, size_t n)
{
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x) {
    const uint32_t k = runcnt[c];
    if (k) {
      acc[c] = seq_add_dev(acc[c], w, k);
      total[c] += k;
      runcnt[c] = 0;
    }
  }
}

// K2: occupancy (the reference's f32 `grid` before normalisation) and density
//     grid / ((double)(float)dV * T_tot), g_ri3Dc.c:451-460; denom == 0 -> occupancy only
__global__ void density_kernel(const uint32_t *__restrict__ runcnt, const float *__restrict__ acc, float w,
                               int have_run, double denom, float *__restrict__ out, size_t n)
{
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x) {
    float v = acc ? acc[c] : 0.0f;
    if (have_run) {
      const uint32_t k = runcnt[c];
      if (k) v = seq_add_dev(v, w, k);
    }
    out[c] = denom != 0.0 ? (float)__ddiv_rn((double)v, denom) : v;
  }
}

// z-fastest [mx][my][mz] -> x-fastest [mz][my][mx]   (grid3_serialise, xdr_grid.c:113-126)
__global__ void zfast_to_xfast_kernel(const float *__restrict__ in, float *__restrict__ out, int mx, int my, int mz)
{
  __shared__ float tile[32][33];
  const int j = blockIdx.z;
  const int i0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r, k = k0 + threadIdx.x;
    if (i < mx && k < mz) tile[r][threadIdx.x] = in[((size_t)i * my + j) * mz + k];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, i = i0 + threadIdx.x;
    if (i < mx && k < mz) out[((size_t)k * my + j) * mx + i] = tile[threadIdx.x][r];
  }
}

// The same reordering with the file's encoding fused in: EMIT 1 = XDR big-endian floats (grid3_serialise +
// xdr_float, xdr_grid.c:113-126), EMIT 2 = gOpenMol plt cells `norm * grid` in host byte order (plt.c:69-75).
template <int EMIT>
__global__ void zfast_to_file_kernel(const float *__restrict__ in, uint32_t *__restrict__ out, int mx, int my, int mz, float scale)
{
  __shared__ float tile[32][33];
  const int j = blockIdx.z;
  const int i0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r, k = k0 + threadIdx.x;
    if (i < mx && k < mz) tile[r][threadIdx.x] = in[((size_t)i * my + j) * mz + k];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, i = i0 + threadIdx.x;
    if (i < mx && k < mz) {
      const float v = tile[threadIdx.x][r];
      out[((size_t)k * my + j) * mx + i] = EMIT == 1 ? __byte_perm(__float_as_uint(v), 0u, 0x0123)
                                                     : __float_as_uint(__fmul_rn(scale, v));
    }
  }
}

// ---------------------------------------------------------------------------
// K4: counter-based synthetic frames; bit-identical to oracle/gc_oracle.c
// ---------------------------------------------------------------------------
__host__ __device__ inline unsigned long long splitmix64(unsigned long long z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned gen_u24(unsigned long long seed, unsigned long long frame,
                                            unsigned long long atom, unsigned long long lane)
{
  const unsigned long long key = seed ^ (frame * 0xD1342543DE82EF95ull) ^ (atom * 0xA24BAED4963EE407ull) ^
                                 (lane * 0x9FB21C651E98DF25ull);
  return (unsigned)(splitmix64(key) >> 40);
}
__device__ __forceinline__ float unit24(unsigned u) { return __fmul_rn(__uint2float_rn(u), 5.9604644775390625e-08f); }

__global__ void gen_frames_kernel(gcb_gen p, long long frame0, long long nframes, int natoms, float *__restrict__ x)
{
  const long long total = nframes * natoms;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long fl = g / natoms;
    const unsigned long long fr = (unsigned long long)(frame0 + fl);
    const unsigned long long i = (unsigned long long)(g - fl * natoms);
    float L[3] = {p.L[0], p.L[1], p.L[2]};
    if (p.box_period > 0) {
      const long long ph = (frame0 + fl) % p.box_period, half = p.box_period / 2;
      const float hh = __fmul_rn(0.5f, (float)half);
      const float tri = __fdiv_rn(__fsub_rn((float)llabs(ph - half), hh), hh);
      const float s = __fadd_rn(1.0f, __fmul_rn(p.box_amp, tri));
      for (int d = 0; d < 3; d++) L[d] = __fmul_rn(p.L[d], s);
    }
    float r[3];
    if (p.mode == GCB_GEN_CLUSTER) {
      const unsigned site = gen_u24(p.seed, fr, i, 3) % (unsigned)p.nsites;
      const float spread = __fmul_rn(p.sigma, 1.7320508f);
      for (int d = 0; d < 3; d++) {
        const float c = __fmul_rn(unit24(gen_u24(p.seed ^ 0x5151515151515151ull, 0, site, d)), L[d]);
        const unsigned s4 = gen_u24(p.seed, fr, i, 4 + d) + gen_u24(p.seed, fr, i, 8 + d) +
                            gen_u24(p.seed, fr, i, 12 + d) + gen_u24(p.seed, fr, i, 16 + d);
        const float off = __fmul_rn(__fsub_rn(__fmul_rn(__uint2float_rn(s4), 5.9604644775390625e-08f), 2.0f), spread);
        float v = __fadd_rn(c, off);
        if (v < 0.0f) v = __fadd_rn(v, L[d]);
        if (v >= L[d]) v = __fsub_rn(v, L[d]);
        r[d] = v;
      }
    } else {
      for (int d = 0; d < 3; d++) r[d] = __fmul_rn(unit24(gen_u24(p.seed, fr, i, d)), L[d]);
      if (p.mode == GCB_GEN_PORE) {
        const float dx = __fsub_rn(r[0], __fmul_rn(0.5f, L[0])), dy = __fsub_rn(r[1], __fmul_rn(0.5f, L[1]));
        const float r2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (r[2] >= p.slab_lo && r[2] < p.slab_hi && r2 > __fmul_rn(p.pore_r, p.pore_r)) {
          const float thick = __fsub_rn(p.slab_hi, p.slab_lo);
          float z = __fmul_rn(unit24(gen_u24(p.seed, fr, i, 3)), __fsub_rn(L[2], thick));
          if (z >= p.slab_lo) z = __fadd_rn(z, thick);
          r[2] = z;
        }
      }
      if (p.mode == GCB_GEN_QUANTISED)
        for (int d = 0; d < 3; d++) r[d] = __fmul_rn(rintf(__fmul_rn(r[d], 1000.0f)), 0.001f);
    }
    float *o = x + (size_t)g * 3;
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
  }
}

// ---------------------------------------------------------------------------
// host object
// ---------------------------------------------------------------------------
struct Slot {
  cudaStream_t copy = nullptr;
  cudaEvent_t copied = nullptr, consumed = nullptr;
  float *h_x = nullptr;               // pinned staging [cap][natoms][3]
  float *d_x = nullptr;               // device staging
  unsigned long long *d_hits = nullptr, *h_hits = nullptr;
  float *d_box3 = nullptr, *h_box3 = nullptr;
  long cap = 0;                       // frames
  long pending = 0;                   // frames whose hit counts have not been collected yet
  size_t frame0 = 0;                  // global index of the first pending frame
};

struct DevHits {                      // hit counters of an add_device_frames call awaiting collection
  unsigned long long *d = nullptr;    // inside one of the counter's hit arenas
  float *d_box3 = nullptr;
  long n = 0;
  size_t frame0 = 0;
};

}  // namespace

constexpr int GCB_XTC_SETS = 6;

struct gcb_counter {
  gcb_tgrid tg{};
  GridParams P{};
  GridParamsFast PF{};
  gcb_counter_opts opts{};
  int natoms = 0, gnx = 0;
  size_t ncells = 0;
  int sm_count = 148;
  bool stream_ok = false;             // index group dense enough for the streaming kernel
  bool ap_ok = false;                 // index group is g0 + i*step: arithmetic selection inside the tile
  ApGroup ap{0, 1, TILE_ATOMS};
  // packed (W-bit lanes, L2-resident) path
  int packed_bits = 0;                // 0: not available for this grid; 8 or 4
  bool packed_auto = false;           // AUTO takes it: uint32 grid beyond L2, packed counters inside it
  long long packed_units = 0;         // atom-frames per round
  uint32_t *d_pk = nullptr;           // packed counters, always all-zero between rounds
  PackedState *d_pstate = nullptr;
  PackedState *h_pstate = nullptr;    // pinned mirror, refreshed (asynchronously) behind every call's rounds
  int packed_pause = 0;               // calls still to be binned without the packed pass (see launch_packed)
  uint64_t packed_rounds = 0, packed_redone = 0;
  // privatised path (PrivateSink): per-CTA W-bit copies of the grid, resting in d_priv between launches
  int priv_bits = 0;                  // 0: grid too large for one CTA's shared memory; 8 or 4
  unsigned priv_words = 0;            // words per private grid
  int priv_ctas = 0;                  // CTAs (private grids) of every private launch
  uint32_t *d_priv = nullptr;         // [priv_ctas][priv_words]
  unsigned long long *d_priv_hits = nullptr;   // [priv_ctas]
  PrivState *d_privstate = nullptr;
  bool priv_dirty = false;            // d_priv holds hits that are not in d_runcnt yet
  unsigned long long priv_bound = 0;  // upper bound of the hits any one private grid holds
  unsigned long long priv_limit = 0;  // flush before priv_bound could exceed this
  long long priv_min_units = 0;       // AUTO takes the private path for launches of at least this many atom-frames
  uint64_t priv_launches = 0, priv_redone = 0;
  ApGroup ap_priv{0, 1, TILE_ATOMS};
  uint2 *d_mid = nullptr;             // 4-bit packed path: byte grid between the packed rounds and the uint32 grid
  int mid_rounds = 0;                 // rounds folded into d_mid since it was last flushed (<= 16)
  bool l2_window = false;             // persisting-L2 window over d_pk set on the compute stream
  bool saturated = false;             // some cell was clamped (diag[3])
  // xtc path: GCB_XTC_SETS sets of {compressed bytes, frame table, status, decoded frames}, copy stream + events
  struct XtcSet {
    uint8_t *d_buf = nullptr; size_t buf_cap = 0;
    void *d_meta = nullptr, *h_meta = nullptr, *d_work = nullptr; int *d_status = nullptr, *h_status = nullptr; long frames_cap = 0;
    float *d_x = nullptr; size_t x_cap = 0;
    cudaStream_t stream = nullptr;                      // H2D of the compressed bytes + frame table + skim
    cudaEvent_t skimmed = nullptr, consumed = nullptr;  // skim done (set stream) / segment kernels done (compute stream)
    long nframes = 0; long first = 0;
  } xtc[GCB_XTC_SETS];
  // multi-GPU statistics scratch
  unsigned long long *d_scr = nullptr, *h_scr = nullptr;
  size_t scr_words = 0;
  cudaStream_t comm_stream = nullptr; // rank-table exchange, beside the compute stream
  cudaEvent_t reduce_done = nullptr;  // behind the D2H of the gathered statistics
  cudaEvent_t scr_up = nullptr, grp_done = nullptr;   // scratch uploaded / NCCL group finished
  bool reduce_pending = false;        // host half of the last all-reduce still to do (finish_reduce)
  size_t pend_total = 0, pend_off = 0, pend_ndev = 0, pend_nlocal = 0, pend_blk = 0;
  std::vector<size_t> pend_frames;    // frames each rank contributed to the pending reduce
  gcb_ctl *ctl = nullptr;             // shared-memory control plane between the ranks of this node (may be absent)

  std::vector<int> h_gindex;          // sorted copy of the index group
  int *d_gindex = nullptr;
  unsigned *d_rank = nullptr;         // rank[r] = number of group entries < r, r in [0, natoms]
  uint32_t *d_runcnt = nullptr, *d_total = nullptr;
  float *d_acc = nullptr;
  bool acc_used = false;              // d_acc / d_total hold something since the last reset (they stay allocated)
  unsigned long long *d_diag = nullptr;
  float *d_out = nullptr, *d_out2 = nullptr;   // density / transpose outputs

  // saturation guard (clamp_counts_kernel): upper bounds on any cell of d_runcnt / d_total
  unsigned long long bound_run = 0, bound_total = 0;
  unsigned long long clamp_at = 1ull << 27, clamp_ceiling = 0xffffffffull;
  bool have_run = false;              // d_runcnt holds hits of weight w_cur not yet folded into d_acc
  float w_cur = 0;
  bool uniform = true, any = false;
  float w_first = 0;

  cudaStream_t compute = nullptr;
  std::vector<Slot> slots;
  int next_slot = 0;
  std::vector<DevHits> devhits;
  // per-frame hit counters of device-resident batches are carved out of a few large device buffers (no allocation
  // per call in steady state, however many calls lie between two collections); `used` goes back to zero whenever
  // no batch is waiting for collection (later launches that reuse the memory are ordered on the compute stream)
  struct HitArena { unsigned long long *d; size_t cap, used; };
  std::vector<HitArena> hit_arenas;

  std::vector<uint64_t> hits_per_frame;
  std::vector<float> weight_per_frame;
  double sumP = 0;
  size_t sumP_frames = 0;
  uint64_t hits_total = 0, edge_overflow = 0;
  int64_t launches = 0;
  void *nccl = nullptr;               // the communicator: an ncclComm_t (nccl_reduce.cpp) or an in-process rank (local_reduce.cu)
  const gcb_reduce_ops *ops = nullptr;
  int nranks = 1, rank = 0;
  std::vector<cudaEvent_t> events;
  // gcb_counter_profile: an event pair around every launch of the streaming kernel (K1's dominant kernel)
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev;
  size_t prof_used = 0;
  std::vector<int> prof_kind;         // per event pair: GCB_PROF_*
  std::vector<double> prof_units;     // per event pair: atom-frames the launch covered
  bool reduced = false;               // statistics below are global after gcb_counter_allreduce
  uint64_t global_frames = 0;
};

namespace {

inline bool same_bits(float p, float q) { uint32_t x, y; memcpy(&x, &p, 4); memcpy(&y, &q, 4); return x == y; }

unsigned rank_host(const gcb_counter *c, int r)
{
  return (unsigned)(std::lower_bound(c->h_gindex.begin(), c->h_gindex.end(), r) - c->h_gindex.begin());
}

int flush_private(gcb_counter *c);

int ensure_acc(gcb_counter *c)
{
  c->acc_used = true;
  if (c->d_acc) return GCB_OK;
  GCB_CUDA(cudaMalloc(&c->d_acc, c->ncells * sizeof(float)));
  GCB_CUDA(cudaMalloc(&c->d_total, c->ncells * sizeof(uint32_t)));
  GCB_CUDA(cudaMemsetAsync(c->d_acc, 0, c->ncells * sizeof(float), c->compute));
  GCB_CUDA(cudaMemsetAsync(c->d_total, 0, c->ncells * sizeof(uint32_t), c->compute));
  return GCB_OK;
}

int elementwise_grid(const gcb_counter *c, size_t n)
{
  const size_t b = (n + 255) / 256, cap = (size_t)c->sm_count * 8;
  return (int)std::max<size_t>(1, std::min(b, cap));
}

// make sure `units` more hits fit into every cell of `cnt` (bound = what the fullest cell may hold now)
int ensure_headroom(gcb_counter *c, uint32_t *cnt, unsigned long long &bound, unsigned long long units)
{
  if (bound + units <= c->clamp_ceiling) return GCB_OK;
  if (bound > c->clamp_at) {
    if (cnt == c->d_runcnt) { int rc = flush_private(c); if (rc) return rc; }
    clamp_counts_kernel<<<elementwise_grid(c, c->ncells), 256, 0, c->compute>>>(cnt, c->ncells, (uint32_t)c->clamp_at, c->d_diag + 3);
    c->launches++;
    GCB_CUDA(cudaGetLastError());
    bound = c->clamp_at;
  }
  if (bound + units > c->clamp_ceiling)
    return gcb::fail(GCB_ERR_RANGE, "a single launch of %llu atom-frames cannot be guarded against uint32 overflow", units);
  return GCB_OK;
}

int fold_pending(gcb_counter *c)
{
  if (!c->have_run) return GCB_OK;
  int rc = flush_private(c);
  if (rc) return rc;
  rc = ensure_acc(c);
  if (rc) return rc;
  rc = ensure_headroom(c, c->d_total, c->bound_total, c->bound_run);
  if (rc) return rc;
  c->bound_total += c->bound_run;
  c->bound_run = 0;
  fold_run_kernel<<<elementwise_grid(c, c->ncells), 256, 0, c->compute>>>(c->d_runcnt, c->d_total, c->d_acc,
                                                                           c->w_cur, c->ncells);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  c->have_run = false;
  return GCB_OK;
}

// gcb_counter_profile: an event pair on the compute stream around one launch (or a short run of launches)
int prof_begin(gcb_counter *c)
{
  if (!c->prof) return GCB_OK;
  if (c->prof_used + 2 > c->prof_ev.size()) {
    c->prof_ev.resize(c->prof_used + 2, nullptr);
    GCB_CUDA(cudaEventCreate(&c->prof_ev[c->prof_used]));
    GCB_CUDA(cudaEventCreate(&c->prof_ev[c->prof_used + 1]));
  }
  GCB_CUDA(cudaEventRecord(c->prof_ev[c->prof_used], c->compute));
  return GCB_OK;
}
int prof_end(gcb_counter *c, int kind, double units)
{
  if (!c->prof) return GCB_OK;
  GCB_CUDA(cudaEventRecord(c->prof_ev[c->prof_used + 1], c->compute));
  c->prof_used += 2;
  c->prof_kind.push_back(kind);
  c->prof_units.push_back(units);
  return GCB_OK;
}

template <class SINK>
size_t stream_smem_bytes() { return (sizeof(StreamSmem<SINK::NSTAGES, SINK::TILE_CAP>) + 15) & ~(size_t)15; }

template <class SINK, bool PBC, bool AP, int GUARD = 0>
int launch_stream(gcb_counter *c, const float *d_x, long long ntiles, const SINK &sink, int grid_cap,
                  const float *d_box3, unsigned long long *d_hits, int *grid_out, const int *run_if = nullptr,
                  const ApGroup *tiles = nullptr)
{
  static size_t attr_done[64] = {0};             // per instantiation and per device (the attribute is per context)
  const size_t smem = stream_smem_bytes<SINK>() + sink.extra_smem();
  const int dv = c->opts.device & 63;
  if (attr_done[dv] < smem) {
    GCB_CUDA(cudaFuncSetAttribute(bin_stream_kernel<SINK, PBC, AP, GUARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dv] = smem;
  }
  const ApGroup &G = tiles ? *tiles : c->ap;
  const int per_sm = std::max(1, (int)((227 * 1024) / (smem + 1024)));
  long long grid = std::min<long long>(ntiles, (long long)c->sm_count * per_sm);
  if (grid_cap > 0) grid = std::min<long long>(grid, grid_cap);
  const int kind = SINK::PRIVATE ? GCB_PROF_PRIVATE : GUARD == 2 ? GCB_PROF_PACKED : GUARD == 1 ? GCB_PROF_REDO : GCB_PROF_DIRECT;
  int rc = prof_begin(c);
  if (rc) return rc;
  bin_stream_kernel<SINK, PBC, AP, GUARD><<<(unsigned)grid, (SINK::CWARPS + 1) * 32, smem, c->compute>>>(
      d_x, c->natoms, c->d_gindex, (unsigned)c->gnx, c->d_rank, ntiles, G, c->PF, sink, d_box3, d_hits, c->d_diag, run_if);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  rc = prof_end(c, kind, (double)ntiles * (AP ? G.tile_atoms : TILE_ATOMS) / (double)c->natoms * (double)c->gnx);
  if (rc) return rc;
  if (grid_out) *grid_out = (int)grid;
  return GCB_OK;
}

// the part of a batch the streaming kernel leaves: fewer than one tile of atoms at the end
template <int MODE, bool PBC>
int launch_gather(gcb_counter *c, const float *d_x, long long e_begin, long long e_end, uint32_t *cnt, float w,
                  const float *d_box3, unsigned long long *d_hits)
{
  if (e_begin >= e_end) return GCB_OK;
  const long long chunk = (long long)GATHER_THREADS * GATHER_ITEMS;
  const long long nchunks = (e_end - e_begin + chunk - 1) / chunk;
  const long long grid = std::min<long long>(nchunks, (long long)c->sm_count * 8);
  int rc = prof_begin(c);
  if (rc) return rc;
  bin_gather_kernel<MODE, PBC><<<(unsigned)grid, GATHER_THREADS, 0, c->compute>>>(
      d_x, c->natoms, c->d_gindex, (unsigned)c->gnx, e_begin, e_end, c->P, cnt, c->d_acc, w, d_box3, d_hits, c->d_diag);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  return prof_end(c, GCB_PROF_OTHER, (double)(e_end - e_begin));
}

// Packed path: rounds of ~packed_units atom-frames; per round the streaming kernel into the packed counters
// (returns at once while backing off), verify + apply, the guarded uint32 kernel (returns at once unless the round
// failed or is skipped), and the gather kernel for the sub-tile tail (straight into the uint32 grid).
template <int W, bool PBC>
int launch_packed(gcb_counter *c, const float *d_x, long nframes, const float *d_box3, unsigned long long *d_hits)
{
  const bool ap = c->ap_ok && c->opts.kernel != GCB_KERNEL_STREAM_INDEXED;
  const int tile_atoms = ap ? c->ap.tile_atoms : TILE_ATOMS;
  // frames per round: as many as packed_units allows, the call cut into equal rounds, a multiple of 4 frames
  // (every round starts 16-byte aligned)
  long fpr = (long)std::max<long long>(4, (c->packed_units / std::max(c->gnx, 1)) & ~3LL);
  {
    const long nrounds = (nframes + fpr - 1) / fpr;
    fpr = std::max<long>(4, (((nframes + nrounds - 1) / nrounds) + 3) & ~3L);
  }
  const size_t nwords = (c->ncells + (32 / W) - 1) / (32 / W), nquads = (nwords + 3) / 4;
  if (!c->d_pk) {
    GCB_CUDA(cudaMalloc(&c->d_pk, nquads * 16));
    GCB_CUDA(cudaMemsetAsync(c->d_pk, 0, nquads * 16, c->compute));
    GCB_CUDA(cudaMalloc(&c->d_pstate, sizeof(PackedState)));
    GCB_CUDA(cudaMemsetAsync(c->d_pstate, 0, sizeof(PackedState), c->compute));
    if (W == 4 && !getenv("GCB_NO_MID")) {
      GCB_CUDA(cudaMalloc(&c->d_mid, nwords * sizeof(uint2)));
      GCB_CUDA(cudaMemsetAsync(c->d_mid, 0, nwords * sizeof(uint2), c->compute));
    }
    GCB_CUDA(cudaHostAlloc((void **)&c->h_pstate, sizeof(PackedState), cudaHostAllocDefault));
    memset(c->h_pstate, 0, sizeof(PackedState));
    // Keep the packed counters L2-resident next to the coordinate stream: a persisting-L2 set-aside with an
    // access-policy window over them on the compute stream.  Measured (tools/l2_window.cu, profiles/l2_window_r02.txt):
    // one RED per 36 streamed bytes into a 64 MB window runs at 83 G/s without it (the stream evicts the window,
    // whatever evict-first hints the loads carry) and at 123 G/s with it -- the rate of a 4 MB window.
    if (!getenv("GCB_NO_L2_PERSIST") && nquads * 16 > ((size_t)32 << 20)) {
      cudaDeviceProp prop;
      GCB_CUDA(cudaGetDeviceProperties(&prop, c->opts.device));
      const size_t want = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, nquads * 16 + ((size_t)2 << 20));
      if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
        cudaStreamAttrValue av = {};
        av.accessPolicyWindow.base_ptr = c->d_pk;
        av.accessPolicyWindow.num_bytes = std::min<size_t>(nquads * 16, (size_t)prop.accessPolicyMaxWindowSize);
        av.accessPolicyWindow.hitRatio = std::min(1.0f, (float)((double)want / (double)(nquads * 16)));
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(c->compute, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess) c->l2_window = true;
      }
      cudaGetLastError();
    }
  }
  PackedState *st = c->d_pstate;
  const int aux_grid = (int)std::max<size_t>(1, std::min<size_t>((nquads + 255) / 256, (size_t)c->sm_count * 8));
  const int apply_grid = (int)std::max<size_t>(1, std::min<size_t>((nwords + 255) / 256, (size_t)c->sm_count * 16));
  for (long f0 = 0; f0 < nframes; f0 += fpr) {
    const long n = std::min(fpr, nframes - f0);
    const float *xs = d_x + (size_t)f0 * c->natoms * 3;
    const float *bs = d_box3 ? d_box3 + 3 * f0 : nullptr;
    const long long total_e = (long long)n * c->gnx, total_atoms = (long long)n * c->natoms;
    const long long ntiles = total_atoms / tile_atoms;
    long long e_begin = 0;
    if (ntiles > 0) {
      packed_begin_kernel<<<1, 1, 0, c->compute>>>(st, c->d_diag);
      PackedSink<W> sink;
      sink.pk = c->d_pk;
      int rc = ap ? launch_stream<PackedSink<W>, PBC, true, 2>(c, xs, ntiles, sink, 0, bs, d_hits + f0, nullptr, &st->alarm)
                  : launch_stream<PackedSink<W>, PBC, false, 2>(c, xs, ntiles, sink, 0, bs, d_hits + f0, nullptr, &st->alarm);
      if (rc) return rc;
      rc = prof_begin(c);
      if (rc) return rc;
      packed_verify_kernel<W><<<aux_grid, 256, 0, c->compute>>>(reinterpret_cast<const uint4 *>(c->d_pk), nquads, d_hits + f0, n, st);
      if (W == 4 && c->d_mid) {
        if (c->mid_rounds >= 16) { rc = flush_private(c); if (rc) return rc; }
        packed_apply_mid_kernel<<<apply_grid, 256, 0, c->compute>>>(c->d_pk, nwords, c->d_mid, st, d_hits + f0, n, c->d_diag);
        c->mid_rounds++;
      } else {
        packed_apply_kernel<W><<<apply_grid, 256, 0, c->compute>>>(c->d_pk, nwords, c->d_runcnt, c->ncells, st, d_hits + f0, n, c->d_diag);
      }
      c->launches += 3;
      GCB_CUDA(cudaGetLastError());
      rc = prof_end(c, GCB_PROF_OTHER, 0.0);
      if (rc) return rc;
      DirectSink<MODE_COUNT> redo;
      redo.counts = c->d_runcnt; redo.acc = c->d_acc; redo.w = 1.0f;
      rc = ap ? launch_stream<DirectSink<MODE_COUNT>, PBC, true, 1>(c, xs, ntiles, redo, 0, bs, d_hits + f0, nullptr, &st->alarm)
              : launch_stream<DirectSink<MODE_COUNT>, PBC, false, 1>(c, xs, ntiles, redo, 0, bs, d_hits + f0, nullptr, &st->alarm);
      if (rc) return rc;
      const long long g1 = ntiles * tile_atoms, f1 = g1 / c->natoms;
      e_begin = f1 * c->gnx + rank_host(c, (int)(g1 - f1 * c->natoms));
    }
    int rc = launch_gather<MODE_COUNT, PBC>(c, xs, e_begin, total_e, c->d_runcnt, 1.0f, bs, d_hits + f0);
    if (rc) return rc;
  }
  // the host learns how the rounds went without waiting for them: the state is copied to a pinned mirror behind the
  // call's kernels and read (possibly one call late) by packed_paused() before the next call is queued
  GCB_CUDA(cudaMemcpyAsync(c->h_pstate, c->d_pstate, sizeof(PackedState), cudaMemcpyDeviceToHost, c->compute));
  return GCB_OK;
}

// Input that keeps overflowing the packed lanes (atoms piled onto a few cells) is binned faster by the direct
// kernel alone -- its hot cells stay L2-resident by themselves -- than by a packed pass that is thrown away every
// other round.  Once the device-side back-off has reached 8 rounds the host stops queueing packed passes for 32
// calls, then probes again.
bool packed_paused(gcb_counter *c)
{
  if (c->packed_pause > 0) { c->packed_pause--; return true; }
  if (c->h_pstate && ((volatile PackedState *)c->h_pstate)->backoff >= 8) {
    c->h_pstate->backoff = 0;                    // (the mirror only: the next copy brings the device's value back)
    c->packed_pause = 32;
    return true;
  }
  return false;
}

// The private grids hold hits that the uint32 grid does not have yet: fold them in.  Called by everything that
// reads or rescales d_runcnt (density, counts, fold, clamp, all-reduce) and before a lane could fill up.
int flush_private(gcb_counter *c)
{
  if (c->mid_rounds > 0) {                       // the byte grid of the 4-bit packed rounds
    const size_t nwords = (c->ncells + 7) / 8;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((nwords + 255) / 256, (size_t)c->sm_count * 16));
    mid_flush_kernel<<<grid, 256, 0, c->compute>>>(c->d_mid, nwords, c->d_runcnt, c->ncells);
    c->launches++;
    GCB_CUDA(cudaGetLastError());
    c->mid_rounds = 0;
  }
  if (!c->priv_dirty) return GCB_OK;
  const int grid = (int)std::max<unsigned>(1u, std::min<unsigned>((c->priv_words + 255u) / 256u, (unsigned)c->sm_count * 8u));
  if (c->priv_bits == 8)
    private_flush_kernel<8><<<grid, 256, 0, c->compute>>>(c->d_priv, c->d_priv_hits, c->priv_words, c->priv_ctas, c->d_runcnt, c->ncells);
  else
    private_flush_kernel<4><<<grid, 256, 0, c->compute>>>(c->d_priv, c->d_priv_hits, c->priv_words, c->priv_ctas, c->d_runcnt, c->ncells);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  c->priv_dirty = false;
  c->priv_bound = 0;
  return GCB_OK;
}

// Private path: the streaming kernel with one CTA per SM, each binning into its own shared-memory copy of the grid
// (PrivateSink); the sub-tile tail goes to the gather kernel and straight into the uint32 grid.
template <int W, bool PBC>
int launch_private(gcb_counter *c, const float *d_x, long nframes, const float *d_box3, unsigned long long *d_hits)
{
  const int tile_atoms = c->ap_priv.tile_atoms;
  if (!c->d_priv) {
    GCB_CUDA(cudaMalloc(&c->d_priv, (size_t)c->priv_ctas * c->priv_words * 4));
    GCB_CUDA(cudaMalloc(&c->d_priv_hits, sizeof(unsigned long long) * (size_t)c->priv_ctas));
    GCB_CUDA(cudaMalloc(&c->d_privstate, sizeof(PrivState)));
    GCB_CUDA(cudaMemsetAsync(c->d_priv, 0, (size_t)c->priv_ctas * c->priv_words * 4, c->compute));
    GCB_CUDA(cudaMemsetAsync(c->d_priv_hits, 0, sizeof(unsigned long long) * (size_t)c->priv_ctas, c->compute));
    GCB_CUDA(cudaMemsetAsync(c->d_privstate, 0, sizeof(PrivState), c->compute));
  }
  // what one private grid may receive from a tile: its selected atoms
  const unsigned long long per_tile = (unsigned long long)(tile_atoms / std::max(c->ap.step, 1)) + 2ull;
  // frames per launch: no private grid may be handed more than the flush threshold allows (a multiple of 4 frames:
  // every launch starts 16-byte aligned)
  const unsigned long long tiles_cap = std::max<unsigned long long>(1, c->priv_limit / per_tile) * (unsigned long long)c->priv_ctas;
  const long fpl = (long)std::max<unsigned long long>(4, (tiles_cap * (unsigned long long)tile_atoms / (unsigned long long)c->natoms) & ~3ull);
  for (long f0 = 0; f0 < nframes; f0 += fpl) {
    const long n = std::min(fpl, nframes - f0);
    const float *xs = d_x + (size_t)f0 * c->natoms * 3;
    const float *bs = d_box3 ? d_box3 + 3 * f0 : nullptr;
    const long long total_e = (long long)n * c->gnx, total_atoms = (long long)n * c->natoms;
    const long long ntiles = total_atoms / tile_atoms;
    long long e_begin = 0;
    if (ntiles > 0) {
      const unsigned long long share = ((unsigned long long)ntiles + (unsigned long long)c->priv_ctas - 1ull) / (unsigned long long)c->priv_ctas * per_tile;
      if (c->priv_bound + share > c->priv_limit) { int rc = flush_private(c); if (rc) return rc; }
      private_begin_kernel<<<1, 1, 0, c->compute>>>(c->d_privstate);
      c->launches++;
      GCB_CUDA(cudaGetLastError());
      PrivateSink<W> sink;
      sink.counts = c->d_runcnt; sink.stage = c->d_priv; sink.stage_hits = c->d_priv_hits; sink.st = c->d_privstate;
      sink.nwords = c->priv_words;
      int rc = launch_stream<PrivateSink<W>, PBC, true>(c, xs, ntiles, sink, c->priv_ctas, bs, d_hits + f0, nullptr, nullptr, &c->ap_priv);
      if (rc) return rc;
      c->priv_dirty = true;
      c->priv_bound += share;
      const long long g1 = ntiles * tile_atoms, f1 = g1 / c->natoms;
      e_begin = f1 * c->gnx + rank_host(c, (int)(g1 - f1 * c->natoms));
    }
    int rc = launch_gather<MODE_COUNT, PBC>(c, xs, e_begin, total_e, c->d_runcnt, 1.0f, bs, d_hits + f0);
    if (rc) return rc;
  }
  return GCB_OK;
}

template <int MODE, bool PBC>
int launch_bin(gcb_counter *c, const float *d_x, long nframes, float w, const float *d_box3,
               unsigned long long *d_hits)
{
  const long long total_e = (long long)nframes * c->gnx;
  if (total_e == 0) return GCB_OK;
  uint32_t *cnt = MODE == MODE_COUNT ? c->d_runcnt : c->d_total;
  long long e_begin = 0;
  const bool aligned = ((uintptr_t)d_x & 15u) == 0;
  const int k = c->opts.kernel;
  const bool use_stream = aligned && k != GCB_KERNEL_GATHER &&
                          (c->stream_ok || k == GCB_KERNEL_STREAM || k == GCB_KERNEL_STREAM_INDEXED || k == GCB_KERNEL_PACKED || k == GCB_KERNEL_PRIVATE);
  // AUTO: grids that fit one CTA's shared memory as 6- or 8-bit counters are binned into per-CTA private copies
  // (PrivateSink) when the launch is large enough to pay for loading and storing them
  if (use_stream && MODE == MODE_COUNT && c->priv_bits != 0 && c->ap_ok &&
      (k == GCB_KERNEL_PRIVATE || (k == GCB_KERNEL_AUTO && total_e >= c->priv_min_units)))
    return c->priv_bits == 8 ? launch_private<8, PBC>(c, d_x, nframes, d_box3, d_hits)
                             : launch_private<4, PBC>(c, d_x, nframes, d_box3, d_hits);
  // AUTO: grids beyond L2 whose packed counters fit it take the packed path (PackedSink)
  if (use_stream && MODE == MODE_COUNT && c->packed_bits != 0 &&
      (k == GCB_KERNEL_PACKED || (k == GCB_KERNEL_AUTO && c->packed_auto && total_e >= (1LL << 20) && !packed_paused(c))))
    return c->packed_bits == 8 ? launch_packed<8, PBC>(c, d_x, nframes, d_box3, d_hits)
                               : launch_packed<4, PBC>(c, d_x, nframes, d_box3, d_hits);
  if (use_stream) {
    const bool ap = c->ap_ok && k != GCB_KERNEL_STREAM_INDEXED;
    const int tile_atoms = ap ? c->ap.tile_atoms : TILE_ATOMS;
    const long long total_atoms = (long long)nframes * c->natoms;
    const long long ntiles = total_atoms / tile_atoms;      // full tiles; the remainder goes to the gather kernel
    if (ntiles > 0) {
      DirectSink<MODE> sink;
      sink.counts = cnt; sink.acc = c->d_acc; sink.w = w;
      int rc = ap ? launch_stream<DirectSink<MODE>, PBC, true>(c, d_x, ntiles, sink, 0, d_box3, d_hits, nullptr)
                  : launch_stream<DirectSink<MODE>, PBC, false>(c, d_x, ntiles, sink, 0, d_box3, d_hits, nullptr);
      if (rc) return rc;
      const long long g1 = ntiles * tile_atoms, f1 = g1 / c->natoms;
      e_begin = f1 * c->gnx + rank_host(c, (int)(g1 - f1 * c->natoms));
    }
  }
  return launch_gather<MODE, PBC>(c, d_x, e_begin, total_e, cnt, w, d_box3, d_hits);
}

template <int MODE>
int launch_bin_pbc(gcb_counter *c, const float *d_x, long nframes, float w, const float *d_box3,
                   unsigned long long *d_hits)
{
  if (c->opts.pbc == GCB_PBC_RECT) return launch_bin<MODE, true>(c, d_x, nframes, w, d_box3, d_hits);
  return launch_bin<MODE, false>(c, d_x, nframes, w, d_box3, d_hits);
}

// Bin `nframes` device-resident frames whose per-frame weights are `w[0..nframes)`
// (host array, NULL = 1).  Frames are cut into runs of bit-identical weight.
// Long runs (and the common uniform case) count integers and are folded into
// the f32 accumulator when the weight changes; short runs add floats directly
// (all additions inside one run carry the same value, so they commute, and
// runs are ordered on the compute stream): both reproduce the reference's
// sequential `grid += weight` exactly.
int bin_device_batch(gcb_counter *c, const float *d_x, long nframes, const float *w, const float *d_box3,
                     unsigned long long *d_hits)
{
  long f = 0;
  while (f < nframes) {
    const float wr = w ? w[f] : 1.0f;
    long e = f + 1;
    // one launch never adds more than a quarter of the counter range to a cell (saturation guard)
    const long max_frames = (long)std::max<unsigned long long>(1, (c->clamp_ceiling / 4) / (unsigned long long)std::max(c->gnx, 1));
    while (e < nframes && e - f < max_frames && same_bits(w ? w[e] : 1.0f, wr)) e++;
    const long n = e - f;
    const unsigned long long units = (unsigned long long)n * (unsigned long long)c->gnx;
    if (!c->any) { c->any = true; c->w_first = wr; }
    else if (!same_bits(wr, c->w_first)) c->uniform = false;

    const float *xs = d_x + (size_t)f * c->natoms * 3;
    const float *bs = d_box3 ? d_box3 + 3 * f : nullptr;
    const bool continues = c->have_run && same_bits(wr, c->w_cur);
    const bool fresh = !c->have_run && !c->acc_used;
    const bool long_run = (double)n * c->gnx >= (double)c->ncells / 8.0;
    int rc;
    if (continues || fresh || long_run) {
      if (c->have_run && !continues) { rc = fold_pending(c); if (rc) return rc; }
      c->have_run = true;
      c->w_cur = wr;
      rc = ensure_headroom(c, c->d_runcnt, c->bound_run, units);
      if (rc) return rc;
      c->bound_run += units;
      rc = launch_bin_pbc<MODE_COUNT>(c, xs, n, wr, bs, d_hits + f);
    } else {
      rc = fold_pending(c);
      if (rc) return rc;
      rc = ensure_acc(c);
      if (rc) return rc;
      rc = ensure_headroom(c, c->d_total, c->bound_total, units);
      if (rc) return rc;
      c->bound_total += units;
      rc = launch_bin_pbc<MODE_FLOAT>(c, xs, n, wr, bs, d_hits + f);
    }
    if (rc) return rc;
    f = e;
  }
  return GCB_OK;
}

// collect the hit counters of a slot whose kernels have finished
int collect_slot(gcb_counter *c, Slot &s)
{
  if (s.pending == 0) return GCB_OK;
  GCB_CUDA(cudaEventSynchronize(s.consumed));
  for (long f = 0; f < s.pending; f++) c->hits_per_frame[s.frame0 + f] = s.h_hits[f];
  s.pending = 0;
  return GCB_OK;
}

int alloc_slots(gcb_counter *c)
{
  if (!c->slots.empty()) return GCB_OK;
  const size_t frame_bytes = (size_t)c->natoms * 12;
  long fpb = c->opts.frames_per_batch;
  if (fpb <= 0) fpb = (long)std::max<size_t>(1, ((size_t)32 << 20) / std::max<size_t>(frame_bytes, 1));
  const int ns = c->opts.nslots > 0 ? c->opts.nslots : 3;
  c->slots.resize(ns);
  for (auto &s : c->slots) {
    s.cap = fpb;
    GCB_CUDA(cudaStreamCreateWithFlags(&s.copy, cudaStreamNonBlocking));
    GCB_CUDA(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
    GCB_CUDA(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
    GCB_CUDA(cudaHostAlloc(&s.h_x, frame_bytes * fpb, cudaHostAllocDefault));
    GCB_CUDA(cudaMalloc(&s.d_x, frame_bytes * fpb + 16));
    GCB_CUDA(cudaMalloc(&s.d_hits, sizeof(unsigned long long) * (fpb + 2)));
    GCB_CUDA(cudaHostAlloc(&s.h_hits, sizeof(unsigned long long) * (fpb + 2), cudaHostAllocDefault));
    GCB_CUDA(cudaMalloc(&s.d_box3, sizeof(float) * 3 * fpb));
    GCB_CUDA(cudaHostAlloc(&s.h_box3, sizeof(float) * 3 * fpb, cudaHostAllocDefault));
  }
  return GCB_OK;
}

// stage `n` frames that already sit in slot.h_x (or are copied there from `src`)
int submit_slot(gcb_counter *c, Slot &s, const float *src, bool src_pinned, long n, const float *w,
                const float *box)
{
  const size_t bytes = (size_t)n * c->natoms * 12;
  const float *from = s.h_x;
  if (src) {
    if (src_pinned) from = src;                 // DMA straight out of the caller's pinned memory
    else memcpy(s.h_x, src, bytes);
  }
  GCB_CUDA(cudaMemcpyAsync(s.d_x, from, bytes, cudaMemcpyHostToDevice, s.copy));
  if (c->opts.pbc == GCB_PBC_RECT) {
    if (!box) return gcb::fail(GCB_ERR_ARG, "GCB_PBC_RECT needs the per-frame box");
    for (long f = 0; f < n; f++) { s.h_box3[3 * f] = box[9 * f]; s.h_box3[3 * f + 1] = box[9 * f + 4]; s.h_box3[3 * f + 2] = box[9 * f + 8]; }
    GCB_CUDA(cudaMemcpyAsync(s.d_box3, s.h_box3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, s.copy));
  }
  GCB_CUDA(cudaEventRecord(s.copied, s.copy));
  GCB_CUDA(cudaStreamWaitEvent(c->compute, s.copied, 0));
  GCB_CUDA(cudaMemsetAsync(s.d_hits, 0, sizeof(unsigned long long) * (n + 2), c->compute));
  int rc = bin_device_batch(c, s.d_x, n, w, c->opts.pbc == GCB_PBC_RECT ? s.d_box3 : nullptr, s.d_hits);
  if (rc) return rc;
  GCB_CUDA(cudaMemcpyAsync(s.h_hits, s.d_hits, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, c->compute));
  GCB_CUDA(cudaEventRecord(s.consumed, c->compute));
  // the next copy into this slot must wait for the kernels that read it
  GCB_CUDA(cudaStreamWaitEvent(s.copy, s.consumed, 0));
  s.pending = n;
  s.frame0 = c->hits_per_frame.size();        // frames keep their submission order whatever slot they used
  c->hits_per_frame.resize(s.frame0 + (size_t)n, 0);
  for (long f = 0; f < n; f++) c->weight_per_frame.push_back(w ? w[f] : 1.0f);
  return GCB_OK;
}

void free_counter(gcb_counter *c)
{
  if (!c) return;
  cudaSetDevice(c->opts.device);
  if (c->compute) cudaStreamSynchronize(c->compute);
  if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
  if (c->nccl && c->ops && c->ops->destroy) c->ops->destroy(c->nccl);
  if (c->ctl) gcb_ctl_close(c->ctl);
  if (c->l2_window) {                  // give the set-aside lines back
    cudaStreamAttrValue av = {};
    cudaStreamSetAttribute(c->compute, cudaStreamAttributeAccessPolicyWindow, &av);
    cudaCtxResetPersistingL2Cache();
  }
  for (auto &s : c->slots) {
    if (s.copy) { cudaStreamSynchronize(s.copy); cudaStreamDestroy(s.copy); }
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.consumed) cudaEventDestroy(s.consumed);
    cudaFreeHost(s.h_x); cudaFree(s.d_x); cudaFree(s.d_hits); cudaFreeHost(s.h_hits);
    cudaFree(s.d_box3); cudaFreeHost(s.h_box3);
  }
  for (auto &d : c->devhits) cudaFree(d.d_box3);           // (d.d lives in a hit arena)
  for (auto &a : c->hit_arenas) cudaFree(a.d);
  cudaFree(c->d_gindex); cudaFree(c->d_rank); cudaFree(c->d_runcnt); cudaFree(c->d_total); cudaFree(c->d_acc);
  cudaFree(c->d_diag); cudaFree(c->d_out); cudaFree(c->d_out2);
  cudaFree(c->d_pk); cudaFree(c->d_pstate); cudaFreeHost(c->h_pstate); cudaFree(c->d_mid);
  cudaFree(c->d_priv); cudaFree(c->d_priv_hits); cudaFree(c->d_privstate);
  for (auto &x : c->xtc) {
    cudaFree(x.d_buf); cudaFree(x.d_meta); cudaFree(x.d_status); cudaFree(x.d_x); cudaFree(x.d_work); cudaFreeHost(x.h_meta); cudaFreeHost(x.h_status);
    if (x.skimmed) cudaEventDestroy(x.skimmed);
    if (x.consumed) cudaEventDestroy(x.consumed);
    if (x.stream) { cudaStreamSynchronize(x.stream); cudaStreamDestroy(x.stream); }
  }
  cudaFree(c->d_scr); cudaFreeHost(c->h_scr);
  for (auto e : c->events) if (e) cudaEventDestroy(e);
  for (auto e : c->prof_ev) if (e) cudaEventDestroy(e);
  if (c->compute) cudaStreamDestroy(c->compute);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->reduce_done) cudaEventDestroy(c->reduce_done);
  if (c->scr_up) cudaEventDestroy(c->scr_up);
  if (c->grp_done) cudaEventDestroy(c->grp_done);
  cudaGetLastError();                  // nothing above may leave an error behind for the next counter's launch checks
  delete c;
}

}  // namespace

extern "C" {

int gcb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int gcb_counter_create(gcb_counter **out, const gcb_tgrid *tg, const int *gindex, int gnx, int natoms,
                       const gcb_counter_opts *opts)
{
  if (!out || !tg || natoms <= 0 || gnx < 0 || (gnx > 0 && !gindex))
    return gcb::fail(GCB_ERR_ARG, "gcb_counter_create: bad argument");
  if (tg->mx[0] <= 0 || tg->mx[1] <= 0 || tg->mx[2] <= 0 || gcb::cells(*tg) > (size_t)INT32_MAX)
    return gcb::fail(GCB_ERR_ARG, "gcb_counter_create: grid %d x %d x %d not supported", tg->mx[0], tg->mx[1], tg->mx[2]);
  for (int i = 0; i < gnx; i++)
    if (gindex[i] < 0 || gindex[i] >= natoms)
      return gcb::fail(GCB_ERR_ARG, "index group entry %d = %d is outside [0, %d)", i, gindex[i], natoms);
  if (gcb_device_count() <= 0) return gcb::fail(GCB_ERR_CUDA, "no CUDA device: libgridcount_b200 has no CPU fallback");

  gcb_counter *c = new (std::nothrow) gcb_counter;
  if (!c) return gcb::fail(GCB_ERR_NOMEM, "out of memory");
  if (opts) c->opts = *opts;
  c->tg = *tg;
  c->natoms = natoms;
  c->gnx = gnx;
  c->ncells = gcb::cells(*tg);
  for (int d = 0; d < 3; d++) {
    c->P.a[d] = tg->a[d]; c->P.b[d] = tg->b[d]; c->P.delta[d] = tg->delta[d]; c->P.fmx[d] = (float)tg->mx[d];
  }
  c->P.my = tg->mx[1]; c->P.mz = tg->mx[2];
  {
    const bool fast = !gcb_edge_vulnerable(tg) && tg->mx[0] < (1 << 21) && tg->mx[1] < (1 << 21) && tg->mx[2] < (1 << 21);
    for (int d = 0; d < 3; d++) {
      c->PF.a[d] = tg->a[d]; c->PF.b[d] = tg->b[d]; c->PF.delta[d] = tg->delta[d]; c->PF.fmx[d] = (float)tg->mx[d];
      const double inv = (double)(1.0f / tg->delta[d]);
      c->PF.inv_lo[d] = fast ? (float)(inv * (1.0 - 4.76837158203125e-07)) : NAN;       // 2^-21
      c->PF.inv_hi[d] = fast ? (float)(inv * (1.0 + 4.76837158203125e-07)) : NAN;
    }
    c->PF.my = tg->mx[1]; c->PF.mz = tg->mx[2];
    c->PF.bias = (int)(0x4B400000u * ((uint32_t)tg->mx[1] * (uint32_t)tg->mx[2] + (uint32_t)tg->mx[2] + 1u));
  }

#define GCB_CREATE_CUDA(call)                                                                        \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { gcb::set_error("%s failed: %s", #call, cudaGetErrorString(e_)); \
       free_counter(c); return GCB_ERR_CUDA; } } while (0)

  GCB_CREATE_CUDA(cudaSetDevice(c->opts.device));
  cudaDeviceProp prop;
  GCB_CREATE_CUDA(cudaGetDeviceProperties(&prop, c->opts.device));
  c->sm_count = prop.multiProcessorCount;
  GCB_CREATE_CUDA(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));

  // integer counts are order free, so the group may be sorted (duplicates kept)
  c->h_gindex.assign(gindex, gindex + gnx);
  std::sort(c->h_gindex.begin(), c->h_gindex.end());
  std::vector<unsigned> rank((size_t)natoms + 1);
  {
    size_t p = 0;
    for (int r = 0; r <= natoms; r++) {
      while (p < c->h_gindex.size() && c->h_gindex[p] < r) p++;
      rank[r] = (unsigned)p;
    }
  }
  // streaming reads every byte of every frame: worth it while >= 1/8 of the 32-byte sectors are wanted anyway
  c->stream_ok = (double)gnx * 8.0 >= (double)natoms && (long long)natoms * 3 >= 1;
  {
    // arithmetic progression?  (duplicates or irregular groups take the indexed path)
    const int step = gnx >= 2 ? c->h_gindex[1] - c->h_gindex[0] : 1;
    bool ap = gnx >= 1 && step >= 1 && natoms >= 256;
    for (int i = 2; ap && i < gnx; i++) ap = c->h_gindex[i] - c->h_gindex[i - 1] == step;
    c->ap_ok = ap;
    if (ap) {
      c->ap.g0 = c->h_gindex[0];
      c->ap.step = step;
      c->ap.tile_atoms = std::min(TILE_ATOMS, natoms & ~3);     // <= natoms: a tile spans at most two frames
    }
  }

  {
    if (const char *e = getenv("GCB_CLAMP_AT")) c->clamp_at = strtoull(e, nullptr, 10);
    if (const char *e = getenv("GCB_CLAMP_CEILING")) c->clamp_ceiling = strtoull(e, nullptr, 10);
    if (c->clamp_at < 1 || c->clamp_ceiling > 0xffffffffull || c->clamp_at * 2 > c->clamp_ceiling) {
      c->clamp_at = 1ull << 27; c->clamp_ceiling = 0xffffffffull;
    }
    // Packed path: bytes while they fit ~80 MB of the 126 MB L2 next to the coordinate stream, nibbles up to twice
    // as many cells (500^3, and every grid the XDR format can hold: 512^3 cells = 64 MB of nibbles).  A round is
    // sized so that a uniform density stays far below the lane range: mean 8 hits per byte cell (P[>255] = 0),
    // 1.5 per nibble cell (P[>15] ~ 7e-12 per cell, about one failed round in a thousand on 500^3, which costs that
    // round and the next a pass of the direct kernel); denser spots fail verification and are binned directly.
    {
      const char *benv = getenv("GCB_PACKED_BITS");
      const int want = benv ? atoi(benv) : 0;
      c->packed_bits = want == 4 ? 4 : want == 8 ? 8 : (c->ncells <= ((size_t)80 << 20) ? 8 : 4);
      c->packed_auto = c->ncells * sizeof(uint32_t) > ((size_t)48 << 20) && c->ncells / (c->packed_bits == 8 ? 1 : 2) <= ((size_t)80 << 20) &&
                       !getenv("GCB_NO_PACKED");
      const char *uenv = getenv("GCB_PACKED_UNITS");
      c->packed_units = uenv ? atoll(uenv) : (long long)((double)c->ncells * (c->packed_bits == 8 ? 8.0 : 1.5));
      if (c->packed_units < 4096) c->packed_units = 4096;
    }
    // Private path: the grid as 8-bit or 4-bit counters next to the 3 x 36 KB tile ring in one CTA's shared memory.
    // A private grid is flushed before its mean load could pass 32 hits per cell (8-bit lanes, range 255: densities
    // that peak at several times their mean stay in range) or 1.5 (4-bit lanes, range 15: P[> 15] ~ 1e-11 per cell
    // of a uniform density); anything denser fails the digit-sum check, is binned again directly, and backs off.
    {
      const size_t optin = (size_t)prop.sharedMemPerBlockOptin, ring = stream_smem_bytes<PrivateSink<8>>();
      const size_t budget = optin > ring + 4096 ? optin - ring - 1024 : 0;
      const size_t w8 = (c->ncells + 3) / 4, w4 = (c->ncells + 7) / 8;
      const char *benv = getenv("GCB_PRIVATE_BITS");
      const int want = benv ? atoi(benv) : 0;
      if (w8 * 4 <= budget && want != 4) { c->priv_bits = 8; c->priv_words = (unsigned)w8; }
      else if (w4 * 4 <= budget) { c->priv_bits = 4; c->priv_words = (unsigned)w4; }
      if (getenv("GCB_NO_PRIVATE") || natoms < 256) c->priv_bits = 0;
      c->priv_ctas = c->sm_count;
      c->priv_limit = c->priv_bits == 8 ? (unsigned long long)c->ncells * 32ull : (unsigned long long)c->ncells * 3ull / 2ull;
      if (const char *e = getenv("GCB_PRIVATE_LIMIT")) c->priv_limit = strtoull(e, nullptr, 10);
      // loading and storing the private grids costs 2 x 4 B x words x CTAs of L2 traffic per launch: worth it from
      // about that many atom-frames on (12 B each would stream in the same time)
      c->priv_min_units = (long long)c->priv_words * c->priv_ctas * 2;
      if (const char *e = getenv("GCB_PRIVATE_MIN_UNITS")) c->priv_min_units = atoll(e);
      c->ap_priv = c->ap;
      c->ap_priv.tile_atoms = std::min(PrivateSink<8>::TILE_CAP, natoms & ~3);
    }
  }
  GCB_CREATE_CUDA(cudaMalloc(&c->d_gindex, sizeof(int) * std::max(gnx, 1)));
  GCB_CREATE_CUDA(cudaMalloc(&c->d_rank, sizeof(unsigned) * rank.size()));
  GCB_CREATE_CUDA(cudaMalloc(&c->d_runcnt, sizeof(uint32_t) * c->ncells));
  GCB_CREATE_CUDA(cudaMalloc(&c->d_diag, sizeof(unsigned long long) * 4));   // [0] edge overflow, [3] saturation flag
  if (gnx) GCB_CREATE_CUDA(cudaMemcpy(c->d_gindex, c->h_gindex.data(), sizeof(int) * gnx, cudaMemcpyHostToDevice));
  GCB_CREATE_CUDA(cudaMemcpy(c->d_rank, rank.data(), sizeof(unsigned) * rank.size(), cudaMemcpyHostToDevice));
  // on the compute stream: it is non-blocking, a legacy-stream cudaMemset could still be
  // running when the first binning kernel starts and would wipe its deposits
  GCB_CREATE_CUDA(cudaMemsetAsync(c->d_runcnt, 0, sizeof(uint32_t) * c->ncells, c->compute));
  GCB_CREATE_CUDA(cudaMemsetAsync(c->d_diag, 0, sizeof(unsigned long long) * 4, c->compute));
  GCB_CREATE_CUDA(cudaStreamSynchronize(c->compute));
#undef GCB_CREATE_CUDA
  *out = c;
  return GCB_OK;
}

void gcb_counter_destroy(gcb_counter *c) { free_counter(c); }

int gcb_counter_reset(gcb_counter *c)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  // No host synchronisation: everything below is ordered on the compute stream behind the work already
  // queued, and results that were never collected are simply dropped (their counter buffers go back to the
  // pool; later launches that reuse them are ordered on the same stream).
  for (auto &s : c->slots) s.pending = 0;
  c->reduce_pending = false;          // statistics of a reduce nobody collected are dropped with the rest
  for (auto &h : c->devhits)
    if (h.d_box3) { GCB_CUDA(cudaStreamSynchronize(c->compute)); cudaFree(h.d_box3); }
  c->devhits.clear();
  for (auto &a : c->hit_arenas) a.used = 0;
  GCB_CUDA(cudaMemsetAsync(c->d_runcnt, 0, sizeof(uint32_t) * c->ncells, c->compute));
  if (c->acc_used) {
    GCB_CUDA(cudaMemsetAsync(c->d_acc, 0, sizeof(float) * c->ncells, c->compute));
    GCB_CUDA(cudaMemsetAsync(c->d_total, 0, sizeof(uint32_t) * c->ncells, c->compute));
    c->acc_used = false;
  }
  GCB_CUDA(cudaMemsetAsync(c->d_diag, 0, sizeof(unsigned long long) * 4, c->compute));
  if (c->d_pstate) GCB_CUDA(cudaMemsetAsync(&c->d_pstate->rounds, 0, 2 * sizeof(unsigned long long), c->compute));
  c->packed_rounds = 0; c->packed_redone = 0;
  if (c->d_priv) {
    if (c->priv_dirty) {
      GCB_CUDA(cudaMemsetAsync(c->d_priv, 0, (size_t)c->priv_ctas * c->priv_words * 4, c->compute));
      GCB_CUDA(cudaMemsetAsync(c->d_priv_hits, 0, sizeof(unsigned long long) * (size_t)c->priv_ctas, c->compute));
    }
    GCB_CUDA(cudaMemsetAsync(&c->d_privstate->launches, 0, 2 * sizeof(unsigned long long), c->compute));
  }
  c->priv_dirty = false; c->priv_bound = 0; c->priv_launches = 0; c->priv_redone = 0;
  if (c->mid_rounds > 0) {
    GCB_CUDA(cudaMemsetAsync(c->d_mid, 0, ((c->ncells + 7) / 8) * sizeof(uint2), c->compute));
    c->mid_rounds = 0;
  }
  c->bound_run = 0; c->bound_total = 0;
  c->reduced = false; c->global_frames = 0;
  c->have_run = false; c->any = false; c->uniform = true; c->w_cur = 0; c->w_first = 0;
  c->hits_per_frame.clear(); c->weight_per_frame.clear();
  c->sumP = 0; c->sumP_frames = 0; c->hits_total = 0; c->edge_overflow = 0;
  return GCB_OK;
}

int gcb_counter_slot(gcb_counter *c, int slot, float **x_pinned, long *capacity_frames)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  int rc = alloc_slots(c);
  if (rc) return rc;
  if (slot < 0 || slot >= (int)c->slots.size()) return gcb::fail(GCB_ERR_ARG, "slot %d out of range", slot);
  Slot &s = c->slots[slot];
  rc = collect_slot(c, s);              // also guarantees the previous H2D out of this slot is done
  if (rc) return rc;
  GCB_CUDA(cudaStreamSynchronize(s.copy));
  if (x_pinned) *x_pinned = s.h_x;
  if (capacity_frames) *capacity_frames = s.cap;
  return GCB_OK;
}

int gcb_counter_submit_slot(gcb_counter *c, int slot, long nframes, const float *weight, const float *box)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  int rc = alloc_slots(c);
  if (rc) return rc;
  if (slot < 0 || slot >= (int)c->slots.size() || nframes < 0 || nframes > c->slots[slot].cap)
    return gcb::fail(GCB_ERR_ARG, "gcb_counter_submit_slot: slot %d / %ld frames out of range", slot, nframes);
  rc = collect_slot(c, c->slots[slot]);
  if (rc) return rc;
  return submit_slot(c, c->slots[slot], nullptr, false, nframes, weight, box);
}

int gcb_counter_add_frames(gcb_counter *c, const float *x, long nframes, const float *weight, const float *box)
{
  if (!c || (!x && nframes > 0) || nframes < 0) return gcb::fail(GCB_ERR_ARG, "gcb_counter_add_frames: bad argument");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  int rc = alloc_slots(c);
  if (rc) return rc;
  cudaPointerAttributes attr;
  bool pinned = false;
  if (cudaPointerGetAttributes(&attr, x) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
  else cudaGetLastError();
  const size_t frame_floats = (size_t)c->natoms * 3;
  long done = 0;
  while (done < nframes) {
    Slot &s = c->slots[c->next_slot];
    c->next_slot = (c->next_slot + 1) % (int)c->slots.size();
    rc = collect_slot(c, s);
    if (rc) return rc;
    if (!pinned) GCB_CUDA(cudaStreamSynchronize(s.copy));   // staging buffer free to overwrite
    const long n = std::min(s.cap, nframes - done);
    rc = submit_slot(c, s, x + (size_t)done * frame_floats, pinned, n, weight ? weight + done : nullptr,
                     box ? box + 9 * (size_t)done : nullptr);
    if (rc) return rc;
    done += n;
  }
  // Pinned x is DMA-ed straight out of the caller's memory: the call returns when the last of those copies has
  // left it (the kernels behind them keep running), so the caller may refill x at once, as with pageable memory.
  if (pinned)
    for (auto &s : c->slots) GCB_CUDA(cudaEventSynchronize(s.copied));
  return GCB_OK;
}

// per-frame hit counters (and boxes) of a batch of device-resident frames: carved out of the hit arenas, zeroed on
// the compute stream, registered for collection
static int reserve_devhits(gcb_counter *c, long nframes, const float *box, DevHits *out)
{
  DevHits h;
  h.n = nframes;
  {
    const size_t need = ((size_t)nframes + 2 + 15) & ~(size_t)15;
    for (auto &a : c->hit_arenas)
      if (a.cap - a.used >= need) { h.d = a.d + a.used; a.used += need; break; }
    if (!h.d) {
      gcb_counter::HitArena a{nullptr, std::max<size_t>(need * 4, c->hit_arenas.empty() ? (size_t)1 << 16 : c->hit_arenas.back().cap * 2), need};
      GCB_CUDA(cudaMalloc(&a.d, sizeof(unsigned long long) * a.cap));
      c->hit_arenas.push_back(a);
      h.d = a.d;
    }
  }
  GCB_CUDA(cudaMemsetAsync(h.d, 0, sizeof(unsigned long long) * (nframes + 2), c->compute));
  if (c->opts.pbc == GCB_PBC_RECT) {
    if (!box) return gcb::fail(GCB_ERR_ARG, "GCB_PBC_RECT needs the per-frame box");
    std::vector<float> b3((size_t)nframes * 3);
    for (long f = 0; f < nframes; f++) { b3[3 * f] = box[9 * f]; b3[3 * f + 1] = box[9 * f + 4]; b3[3 * f + 2] = box[9 * f + 8]; }
    GCB_CUDA(cudaMalloc(&h.d_box3, sizeof(float) * 3 * nframes));
    GCB_CUDA(cudaMemcpy(h.d_box3, b3.data(), sizeof(float) * 3 * nframes, cudaMemcpyHostToDevice));
  }
  h.frame0 = c->hits_per_frame.size();
  c->hits_per_frame.resize(h.frame0 + (size_t)nframes, 0);
  c->devhits.push_back(h);
  *out = h;
  return GCB_OK;
}

int gcb_counter_add_device_frames(gcb_counter *c, const float *x_dev, long nframes, const float *weight,
                                  const float *box)
{
  if (!c || (!x_dev && nframes > 0) || nframes < 0)
    return gcb::fail(GCB_ERR_ARG, "gcb_counter_add_device_frames: bad argument");
  if (nframes == 0) return GCB_OK;
  GCB_CUDA(cudaSetDevice(c->opts.device));
  DevHits h;
  int rc = reserve_devhits(c, nframes, box, &h);
  if (rc) return rc;
  rc = bin_device_batch(c, x_dev, nframes, weight, h.d_box3, h.d);
  if (rc) return rc;
  for (long f = 0; f < nframes; f++) c->weight_per_frame.push_back(weight ? weight[f] : 1.0f);
  return GCB_OK;
}

extern "C" int gcb_xtc_skim_on_stream_(const uint8_t *buf_dev, const gcb_xtc_frame *frames, long nframes, int natoms, void *meta_dev,
                                       int *status_dev, void *stream, void *meta_host_pinned, void *work_dev);
extern "C" int gcb_xtc_segments_on_stream_(const uint8_t *buf_dev, long nframes, int natoms, float *x_dev, const void *meta_dev,
                                           int *status_dev, void *stream, const void *work_dev);
extern "C" size_t gcb_xtc_work_bytes_(long nframes, int natoms);
extern "C" size_t gcb_xtc_meta_bytes_(long nframes);

// A batch of xtc frames can be binned straight out of the decoder (xtc_bin_kernel) when it is one run of integer
// counts: one weight for the whole batch that continues the current run (or starts one under the rules of
// bin_device_batch), the uint32 grid as the target (grids on the packed-counter path take the two-step route),
// and not more frames than one launch may add to a cell.  Prepares the run exactly as bin_device_batch would.
static bool xtc_batch_fusable(gcb_counter *c, long n, const float *w)
{
  const bool off = getenv("GCB_XTC_UNFUSED") != nullptr;
  if (off || c->opts.kernel != GCB_KERNEL_AUTO) return false;
  if (c->packed_bits != 0 && c->packed_auto) return false;
  const float wr = w ? w[0] : 1.0f;
  for (long f = 1; f < n; f++) if (!same_bits(w ? w[f] : 1.0f, wr)) return false;
  const long max_frames = (long)std::max<unsigned long long>(1, (c->clamp_ceiling / 4) / (unsigned long long)std::max(c->gnx, 1));
  if (n > max_frames) return false;
  const bool continues = c->have_run && same_bits(wr, c->w_cur);
  const bool fresh = !c->have_run && !c->acc_used;
  const bool long_run = (double)n * c->gnx >= (double)c->ncells / 8.0;
  return continues || fresh || long_run;
}

static int xtc_bin_batch(gcb_counter *c, const uint8_t *d_buf, const void *d_meta, const void *d_work, const int *d_status,
                         long n, const float *w, const float *d_box3, unsigned long long *d_hits)
{
  const float wr = w ? w[0] : 1.0f;
  const unsigned long long units = (unsigned long long)n * (unsigned long long)c->gnx;
  if (!c->any) { c->any = true; c->w_first = wr; }
  else if (!same_bits(wr, c->w_first)) c->uniform = false;
  int rc;
  if (c->have_run && !same_bits(wr, c->w_cur)) { rc = fold_pending(c); if (rc) return rc; }
  c->have_run = true;
  c->w_cur = wr;
  rc = ensure_headroom(c, c->d_runcnt, c->bound_run, units);
  if (rc) return rc;
  c->bound_run += units;
  const int *nseg = static_cast<const int *>(d_work);
  const xtcdev::Checkpoint *ckpt = reinterpret_cast<const xtcdev::Checkpoint *>(static_cast<const char *>(d_work) + xtcdev::work_ckpt_offset(n));
  const long long nthreads = (long long)n * xtcdev::max_segments(c->natoms);
  const unsigned grid = (unsigned)((nthreads + 127) / 128);
  const xtcdev::FrameDev *meta = static_cast<const xtcdev::FrameDev *>(d_meta);
  const bool pbc = c->opts.pbc == GCB_PBC_RECT;
  rc = prof_begin(c);
  if (rc) return rc;
#define GCB_XTC_BIN(PBC_, AP_)                                                                                              \
  xtc_bin_kernel<PBC_, AP_><<<grid, 128, 0, c->compute>>>(d_buf, meta, n, c->natoms, ckpt, nseg, d_status, c->P, c->d_runcnt, \
                                                          c->ap.g0, c->ap.step, c->gnx, c->d_rank, d_box3, d_hits, c->d_diag)
  if (c->ap_ok) { if (pbc) GCB_XTC_BIN(true, true); else GCB_XTC_BIN(false, true); }
  else          { if (pbc) GCB_XTC_BIN(true, false); else GCB_XTC_BIN(false, false); }
#undef GCB_XTC_BIN
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  return prof_end(c, GCB_PROF_OTHER, (double)units);
}

int gcb_counter_add_xtc(gcb_counter *c, const uint8_t *buf, size_t len, const gcb_xtc_frame *frames, long nframes,
                        const float *weight)
{
  if (!c || !buf || !frames || nframes < 0) return gcb::fail(GCB_ERR_ARG, "gcb_counter_add_xtc: bad argument");
  if (nframes == 0) return GCB_OK;
  GCB_CUDA(cudaSetDevice(c->opts.device));
  if (!c->xtc[0].stream) {
    for (auto &x : c->xtc) {
      GCB_CUDA(cudaStreamCreateWithFlags(&x.stream, cudaStreamNonBlocking));
      GCB_CUDA(cudaEventCreateWithFlags(&x.skimmed, cudaEventDisableTiming));
      GCB_CUDA(cudaEventCreateWithFlags(&x.consumed, cudaEventDisableTiming));
    }
  }
  static const bool trace = getenv("GCB_TRACE") != nullptr;
  auto now = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; };
  const double t_begin = now();
  const size_t frame_bytes = (size_t)c->natoms * 12;
  // The pipeline.  A batch goes through: H2D of its compressed bytes -> skim -> segment kernel(s).  The skim walks
  // each frame's control bits as one chain of dependent steps: ~2 ms for a 60 000-atom frame whether the batch has 64
  // frames or 1024.  So every batch has its own set of buffers and its own stream for copy + skim (GCB_XTC_SETS of them
  // in flight: the skims of several batches overlap each other and the copies), and only the segment kernels, which
  // are short and wide, queue on the compute stream.  Batches of ~64 MB of compressed bytes (1.2 ms of PCIe), at
  // least 64 frames.
  const char *env = getenv("GCB_XTC_BATCH");
  long per;
  if (env) per = atol(env);
  else {
    const double bytes_per_frame = (double)(frames[nframes - 1].payload_offset + frames[nframes - 1].payload_bytes - frames[0].payload_offset) / (double)nframes;
    per = (long)(((size_t)64 << 20) / std::max(bytes_per_frame, 1.0));
    // at least 64 frames: a frame's skim takes ~30 ns per atom and its compressed bytes ~0.1 ns per atom of PCIe
    // time, so with GCB_XTC_SETS batches in flight the copies hide the skims from ~55 frames per batch on,
    // whatever the frame size (large systems simply get large batches: 1 M atoms -> 320 MB of compressed bytes)
    per = std::max<long>(64, std::min<long>(per, 16384));
  }
  per = std::max<long>(4, per & ~3L);                         // batches start 16-byte aligned in the decoded array
  std::vector<float> box3;
  int set = 0;
  auto check_set = [&](gcb_counter::XtcSet &x) -> int {       // wait for a set's batch, look at the decoder's verdicts
    if (x.nframes == 0) return GCB_OK;
    GCB_CUDA(cudaEventSynchronize(x.consumed));
    for (long f = 0; f < x.nframes; f++)
      if (x.h_status[f]) {
        x.nframes = 0;
        return gcb::fail(GCB_ERR_FORMAT, "xtc frame %ld: damaged compressed stream (frames of its batch that were decoded before the "
                                         "damage may have been binned: reset the counter)", x.first + f);
      }
    x.nframes = 0;
    return GCB_OK;
  };
  long nfused = 0;
  // the first two batches are a quarter and a half of the others: the pipeline starts working after a quarter of a
  // batch's copy time (nothing overlaps the first copy and the first skim)
  long f0 = 0;
  for (long batch = 0; f0 < nframes; batch++, set = (set + 1) % GCB_XTC_SETS) {
    const long want = batch == 0 ? std::max<long>(16, (per / 4) & ~3L) : batch == 1 ? std::max<long>(16, (per / 2) & ~3L) : per;
    const long n = std::min(want, nframes - f0);
    gcb_counter::XtcSet &x = c->xtc[set];
    int rc = check_set(x);
    if (rc) return rc;
    size_t lo = frames[f0].payload_offset, hi = 0;
    for (long f = f0; f < f0 + n; f++) {
      if (frames[f].natoms != c->natoms) return gcb::fail(GCB_ERR_FORMAT, "xtc frame %ld has %d atoms, the counter %d", f, frames[f].natoms, c->natoms);
      const size_t b = frames[f].payload_offset;
      const size_t nb = frames[f].natoms <= 9 ? (size_t)frames[f].natoms * 12 : ((size_t)frames[f].payload_bytes + 3) & ~(size_t)3;
      if (b + nb > len) return gcb::fail(GCB_ERR_ARG, "xtc frame %ld lies outside the buffer", f);
      lo = std::min(lo, b); hi = std::max(hi, b + nb);
    }
    lo &= ~(size_t)3;
    const float *wb = weight ? weight + f0 : nullptr;
    const bool fused = xtc_batch_fusable(c, n, wb);
    if (hi - lo + 16 > x.buf_cap) { cudaFree(x.d_buf); x.d_buf = nullptr; x.buf_cap = (hi - lo) + (hi - lo) / 8 + 16; GCB_CUDA(cudaMalloc(&x.d_buf, x.buf_cap)); }
    if (n > x.frames_cap) {
      cudaFree(x.d_meta); cudaFree(x.d_status); cudaFree(x.d_work); cudaFreeHost(x.h_meta); cudaFreeHost(x.h_status);
      x.d_meta = x.h_meta = x.d_work = nullptr; x.d_status = x.h_status = nullptr;
      x.frames_cap = n;
      GCB_CUDA(cudaMalloc(&x.d_meta, gcb_xtc_meta_bytes_(n)));
      GCB_CUDA(cudaMalloc(&x.d_status, sizeof(int) * (size_t)n));
      GCB_CUDA(cudaMalloc(&x.d_work, gcb_xtc_work_bytes_(n, c->natoms)));
      GCB_CUDA(cudaHostAlloc(&x.h_meta, gcb_xtc_meta_bytes_(n), cudaHostAllocDefault));
      GCB_CUDA(cudaHostAlloc((void **)&x.h_status, sizeof(int) * (size_t)n, cudaHostAllocDefault));
    }
    if (!fused && (size_t)n * frame_bytes + 16 > x.x_cap) { cudaFree(x.d_x); x.d_x = nullptr; x.x_cap = (size_t)n * frame_bytes + 16; GCB_CUDA(cudaMalloc(&x.d_x, x.x_cap)); }
    // compressed bytes, frame table and skim on the set's stream; the segment kernels on the compute stream
    GCB_CUDA(cudaMemcpyAsync(x.d_buf, buf + lo, hi - lo, cudaMemcpyHostToDevice, x.stream));
    std::vector<gcb_xtc_frame> rel(frames + f0, frames + f0 + n);
    for (auto &r : rel) r.payload_offset -= lo;
    const float *boxarg = nullptr;
    if (c->opts.pbc == GCB_PBC_RECT) {
      box3.resize((size_t)n * 9);
      for (long f = 0; f < n; f++) memcpy(&box3[(size_t)f * 9], frames[f0 + f].box, sizeof(float) * 9);
      boxarg = box3.data();
    }
    rc = gcb_xtc_skim_on_stream_(x.d_buf, rel.data(), n, c->natoms, x.d_meta, x.d_status, x.stream, x.h_meta, x.d_work);
    if (rc) return rc;
    GCB_CUDA(cudaEventRecord(x.skimmed, x.stream));
    GCB_CUDA(cudaStreamWaitEvent(c->compute, x.skimmed, 0));
    if (!fused) {
      rc = gcb_xtc_segments_on_stream_(x.d_buf, n, c->natoms, x.d_x, x.d_meta, x.d_status, c->compute, x.d_work);
      if (rc) return rc;
    }
    GCB_CUDA(cudaMemcpyAsync(x.h_status, x.d_status, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->compute));
    if (fused) {
      DevHits h;
      rc = reserve_devhits(c, n, boxarg, &h);
      if (rc) return rc;
      rc = xtc_bin_batch(c, x.d_buf, x.d_meta, x.d_work, x.d_status, n, wb, h.d_box3, h.d);
      if (rc) return rc;
      for (long f = 0; f < n; f++) c->weight_per_frame.push_back(wb ? wb[f] : 1.0f);
      nfused += n;
    } else {
      rc = gcb_counter_add_device_frames(c, x.d_x, n, wb, boxarg);
      if (rc) return rc;
    }
    // (the next copy into this set is queued only after the host has waited for this event in check_set: no wait on
    // the copy stream here, it would hold back the copy of the NEXT batch, which goes into another set)
    GCB_CUDA(cudaEventRecord(x.consumed, c->compute));
    x.nframes = n; x.first = f0;
    f0 += n;
  }
  const double t_queued = now();
  for (auto &x : c->xtc) { int rc = check_set(x); if (rc) return rc; }
  if (trace)
    fprintf(stderr, "[gcb add_xtc] %ld frames in batches of %ld (%ld binned out of the decoder): queued after %.3f ms, done after %.3f ms\n",
            nframes, per, nfused, t_queued - t_begin, now() - t_begin);
  return GCB_OK;
}

static int finish_reduce(gcb_counter *c);

int gcb_counter_sync(gcb_counter *c)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  { int rc = finish_reduce(c); if (rc) return rc; }
  for (auto &s : c->slots) { int rc = collect_slot(c, s); if (rc) return rc; }
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  for (auto &h : c->devhits) {
    std::vector<unsigned long long> tmp((size_t)h.n);
    GCB_CUDA(cudaMemcpy(tmp.data(), h.d, sizeof(unsigned long long) * h.n, cudaMemcpyDeviceToHost));
    for (long f = 0; f < h.n; f++) c->hits_per_frame[h.frame0 + f] = tmp[f];
    cudaFree(h.d_box3);
  }
  c->devhits.clear();
  for (auto &a : c->hit_arenas) a.used = 0;
  {
    unsigned long long diag[4];
    GCB_CUDA(cudaMemcpy(diag, c->d_diag, sizeof(diag), cudaMemcpyDeviceToHost));
    if (!c->reduced) c->edge_overflow = diag[0];      // after an all-reduce the statistic is the global one
    c->saturated = diag[3] != 0;
    if (c->d_pstate) {
      PackedState ps;
      GCB_CUDA(cudaMemcpy(&ps, c->d_pstate, sizeof(ps), cudaMemcpyDeviceToHost));
      c->packed_rounds = ps.rounds; c->packed_redone = ps.redone;
    }
    if (c->d_privstate) {
      PrivState ps;
      GCB_CUDA(cudaMemcpy(&ps, c->d_privstate, sizeof(ps), cudaMemcpyDeviceToHost));
      c->priv_launches = ps.launches; c->priv_redone = ps.redone + (ps.failed ? 1 : 0);
    }
  }
  // sumP: `sumP += (double)weight` once per hit, frame by frame (g_ri3Dc.c:151,426)
  for (; c->sumP_frames < c->hits_per_frame.size(); c->sumP_frames++) {
    const uint64_t h = c->hits_per_frame[c->sumP_frames];
    c->sumP = gcb::seq_add_f64(c->sumP, (double)c->weight_per_frame[c->sumP_frames], h);
    c->hits_total += h;
  }
  return GCB_OK;
}

int gcb_counter_stats(gcb_counter *c, gcb_stats *out)
{
  if (!c || !out) return gcb::fail(GCB_ERR_ARG, "null argument");
  int rc = gcb_counter_sync(c);
  if (rc) return rc;
  out->frames = c->reduced ? c->global_frames : c->hits_per_frame.size();
  out->atom_frames = out->frames * (uint64_t)c->gnx;
  out->hits = c->hits_total;
  out->edge_overflow = c->edge_overflow;
  out->sumP = c->sumP;
  out->uniform_weight = c->uniform ? 1 : 0;
  out->weight = c->w_first;
  out->kernel_launches = c->launches;
  out->packed_rounds = c->packed_rounds;
  out->packed_redone = c->packed_redone;
  out->private_launches = c->priv_launches;
  out->private_redone = c->priv_redone;
  out->saturated = c->saturated ? 1 : 0;
  return GCB_OK;
}

int gcb_counter_hits_per_frame(gcb_counter *c, uint64_t *hits, long capacity, long *nframes_out)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  int rc = gcb_counter_sync(c);
  if (rc) return rc;
  const long n = (long)c->hits_per_frame.size();
  if (nframes_out) *nframes_out = n;
  if (hits) memcpy(hits, c->hits_per_frame.data(), sizeof(uint64_t) * (size_t)std::min(n, capacity));
  return GCB_OK;
}

__global__ void add_u32_kernel(const uint32_t *__restrict__ p, const uint32_t *__restrict__ q, uint32_t *__restrict__ o, size_t n)
{
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x)
    o[c] = p[c] + (q ? q[c] : 0u);
}

static int ensure_out(gcb_counter *c)
{
  if (!c->d_out) GCB_CUDA(cudaMalloc(&c->d_out, c->ncells * sizeof(float)));
  return GCB_OK;
}

int gcb_counter_device_counts(gcb_counter *c, uint32_t **counts_dev)
{
  if (!c || !counts_dev) return gcb::fail(GCB_ERR_ARG, "null argument");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  int rc = flush_private(c);
  if (rc) return rc;
  if (!c->acc_used) { *counts_dev = c->d_runcnt; return GCB_OK; }
  rc = ensure_out(c);
  if (rc) return rc;
  add_u32_kernel<<<elementwise_grid(c, c->ncells), 256, 0, c->compute>>>(c->d_total, c->have_run ? c->d_runcnt : nullptr,
                                                                          (uint32_t *)c->d_out, c->ncells);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  *counts_dev = (uint32_t *)c->d_out;
  return GCB_OK;
}

int gcb_counter_counts(gcb_counter *c, uint32_t *counts_zfast)
{
  if (!c || !counts_zfast) return gcb::fail(GCB_ERR_ARG, "null argument");
  uint32_t *d = nullptr;
  int rc = gcb_counter_device_counts(c, &d);
  if (rc) return rc;
  GCB_CUDA(cudaMemcpyAsync(counts_zfast, d, c->ncells * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->compute));
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  return GCB_OK;
}

static int density_to_device(gcb_counter *c, double denom)
{
  int rc = ensure_out(c);
  if (rc) return rc;
  rc = flush_private(c);
  if (rc) return rc;
  density_kernel<<<elementwise_grid(c, c->ncells), 256, 0, c->compute>>>(c->d_runcnt, c->acc_used ? c->d_acc : nullptr, c->w_cur,
                                                                          c->have_run ? 1 : 0, denom, c->d_out, c->ncells);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  return GCB_OK;
}

int gcb_counter_device_density(gcb_counter *c, double T_tot, float **grid_zfast_dev)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  const float dV = (float)gcb_delta_v(&c->tg);
  int rc = density_to_device(c, (double)dV * T_tot);
  if (rc) return rc;
  if (grid_zfast_dev) *grid_zfast_dev = c->d_out;
  return GCB_OK;
}

int gcb_counter_occupancy(gcb_counter *c, float *occ_zfast)
{
  if (!c || !occ_zfast) return gcb::fail(GCB_ERR_ARG, "null argument");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  int rc = density_to_device(c, 0.0);
  if (rc) return rc;
  GCB_CUDA(cudaMemcpyAsync(occ_zfast, c->d_out, c->ncells * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  return GCB_OK;
}

int gcb_counter_density(gcb_counter *c, double T_tot, float *grid_zfast, float *grid_xfast, gcb_tgrid *tg_out)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  const float dV = (float)gcb_delta_v(&c->tg);            // `real dV = DeltaV(&tgrid)` (g_ri3Dc.c:453)
  const double denom = (double)dV * T_tot;                // (dV*T_tot) with T_tot double (g_ri3Dc.c:458)
  int rc = density_to_device(c, denom);
  if (rc) return rc;
  if (grid_zfast)
    GCB_CUDA(cudaMemcpyAsync(grid_zfast, c->d_out, c->ncells * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
  if (grid_xfast) {
    if (!c->d_out2) GCB_CUDA(cudaMalloc(&c->d_out2, c->ncells * sizeof(float)));
    dim3 block(32, 8), grid((c->tg.mx[2] + 31) / 32, (c->tg.mx[0] + 31) / 32, c->tg.mx[1]);
    zfast_to_xfast_kernel<<<grid, block, 0, c->compute>>>(c->d_out, c->d_out2, c->tg.mx[0], c->tg.mx[1], c->tg.mx[2]);
    c->launches++;
    GCB_CUDA(cudaGetLastError());
    GCB_CUDA(cudaMemcpyAsync(grid_xfast, c->d_out2, c->ncells * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
  }
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  if (tg_out) { *tg_out = c->tg; tg_out->tweight = (float)T_tot; }   // g_ri3Dc.c:460
  return GCB_OK;
}

// The finished density as file images produced on the device.  The prologues are a few dozen bytes and
// come from grid_io.cpp; the cell payload is reordered and encoded by zfast_to_file_kernel and copied
// straight behind the prologue in the caller's buffer.
extern "C" size_t gcb_xdr_prologue_(const gcb_tgrid *tg, const char *header, uint8_t *out);
extern "C" size_t gcb_plt_prologue_(const gcb_tgrid *tg, uint8_t *out);

static int emit_file_payload(gcb_counter *c, double T_tot, int emit, float scale, uint8_t *payload_out)
{
  const float dV = (float)gcb_delta_v(&c->tg);            // `real dV = DeltaV(&tgrid)` (g_ri3Dc.c:453)
  int rc = density_to_device(c, (double)dV * T_tot);
  if (rc) return rc;
  if (!c->d_out2) GCB_CUDA(cudaMalloc(&c->d_out2, c->ncells * sizeof(float)));
  dim3 block(32, 8), grid((c->tg.mx[2] + 31) / 32, (c->tg.mx[0] + 31) / 32, c->tg.mx[1]);
  if (emit == 1)
    zfast_to_file_kernel<1><<<grid, block, 0, c->compute>>>(c->d_out, (uint32_t *)c->d_out2, c->tg.mx[0], c->tg.mx[1], c->tg.mx[2], 1.0f);
  else
    zfast_to_file_kernel<2><<<grid, block, 0, c->compute>>>(c->d_out, (uint32_t *)c->d_out2, c->tg.mx[0], c->tg.mx[1], c->tg.mx[2], scale);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  GCB_CUDA(cudaMemcpyAsync(payload_out, c->d_out2, c->ncells * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  return GCB_OK;
}

int gcb_counter_xdr(gcb_counter *c, double T_tot, const char *header, uint8_t *out, size_t cap, size_t *written,
                    gcb_tgrid *tg_out)
{
  if (!c || !out) return gcb::fail(GCB_ERR_ARG, "gcb_counter_xdr: null argument");
  if (!header) header = "";
  if (strlen(header) > GCB_HEADER_MAX) return gcb::fail(GCB_ERR_ARG, "header longer than %d bytes", GCB_HEADER_MAX);
  if (c->ncells > GCB_NGRID_MAX) return gcb::fail(GCB_ERR_RANGE, "grid has %zu cells; the format allows %u", c->ncells, GCB_NGRID_MAX);
  GCB_CUDA(cudaSetDevice(c->opts.device));
  gcb_tgrid tg = c->tg;
  tg.tweight = (float)T_tot;                               // g_ri3Dc.c:460
  const size_t need = gcb_xdr_size(&tg, header);
  if (cap < need) return gcb::fail(GCB_ERR_ARG, "gcb_counter_xdr: buffer of %zu bytes, need %zu", cap, need);
  const size_t pro = gcb_xdr_prologue_(&tg, header, out);
  int rc = emit_file_payload(c, T_tot, 1, 1.0f, out + pro);
  if (rc) return rc;
  if (written) *written = need;
  if (tg_out) *tg_out = tg;
  return GCB_OK;
}

int gcb_counter_grid_write(gcb_counter *c, double T_tot, const char *path, const char *header, gcb_tgrid *tg_out)
{
  if (!c || !path) return gcb::fail(GCB_ERR_ARG, "gcb_counter_grid_write: null argument");
  gcb_tgrid tg = c->tg;
  const size_t need = gcb_xdr_size(&tg, header ? header : "");
  uint8_t *buf = nullptr;
  GCB_CUDA(cudaSetDevice(c->opts.device));
  GCB_CUDA(cudaHostAlloc((void **)&buf, need, cudaHostAllocDefault));     // pinned: the payload arrives by DMA
  size_t n = 0;
  int rc = gcb_counter_xdr(c, T_tot, header, buf, need, &n, tg_out);
  if (!rc) {
    FILE *fp = fopen(path, "wb");
    if (!fp) rc = gcb::fail(GCB_ERR_IO, "cannot open %s for writing", path);
    else {
      const bool ok = fwrite(buf, 1, n, fp) == n;
      if (fclose(fp) != 0 || !ok) rc = gcb::fail(GCB_ERR_IO, "short write to %s", path);
    }
  }
  cudaFreeHost(buf);
  return rc;
}

int gcb_counter_plt(gcb_counter *c, double T_tot, float unit, uint8_t *out, size_t cap, size_t *written)
{
  if (!c || !out) return gcb::fail(GCB_ERR_ARG, "gcb_counter_plt: null argument");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  const size_t need = 44 + 4 * c->ncells;
  if (cap < need) return gcb::fail(GCB_ERR_ARG, "gcb_counter_plt: buffer of %zu bytes, need %zu", cap, need);
  const size_t pro = gcb_plt_prologue_(&c->tg, out);
  const float scale = (float)(1. / (double)unit);          // `real norm = 1./unit` (plt.c:48)
  int rc = emit_file_payload(c, T_tot, 2, scale, out + pro);
  if (rc) return rc;
  if (written) *written = need;
  return GCB_OK;
}

int gcb_counter_event_record(gcb_counter *c, int id)
{
  if (!c || id < 0 || id >= GCB_MAX_EVENTS) return gcb::fail(GCB_ERR_ARG, "gcb_counter_event_record: bad argument");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  if ((int)c->events.size() <= id) c->events.resize((size_t)id + 1, nullptr);
  if (!c->events[id]) GCB_CUDA(cudaEventCreate(&c->events[id]));
  GCB_CUDA(cudaEventRecord(c->events[id], c->compute));
  return GCB_OK;
}

int gcb_counter_event_elapsed(gcb_counter *c, int id0, int id1, float *ms)
{
  if (!c || !ms || id0 < 0 || id1 < 0 || id0 >= (int)c->events.size() || id1 >= (int)c->events.size() ||
      !c->events[id0] || !c->events[id1])
    return gcb::fail(GCB_ERR_STATE, "gcb_counter_event_elapsed: events %d/%d were not recorded", id0, id1);
  GCB_CUDA(cudaSetDevice(c->opts.device));
  GCB_CUDA(cudaEventSynchronize(c->events[id1]));
  GCB_CUDA(cudaEventElapsedTime(ms, c->events[id0], c->events[id1]));
  return GCB_OK;
}

int gcb_counter_profile(gcb_counter *c, int enable)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  c->prof = enable != 0;
  c->prof_used = 0;
  c->prof_kind.clear();
  c->prof_units.clear();
  return GCB_OK;
}

int gcb_counter_profile_kinds(gcb_counter *c, double ms[GCB_PROF_KINDS], long launches[GCB_PROF_KINDS], double atom_frames[GCB_PROF_KINDS])
{
  if (!c || !ms || !launches || !atom_frames) return gcb::fail(GCB_ERR_ARG, "gcb_counter_profile_kinds: null argument");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  for (int k = 0; k < GCB_PROF_KINDS; k++) { ms[k] = 0; launches[k] = 0; atom_frames[k] = 0; }
  for (size_t i = 0; i + 1 < c->prof_used; i += 2) {
    float t = 0;
    GCB_CUDA(cudaEventElapsedTime(&t, c->prof_ev[i], c->prof_ev[i + 1]));
    const int k = c->prof_kind[i / 2];
    ms[k] += t; launches[k]++; atom_frames[k] += c->prof_units[i / 2];
  }
  return GCB_OK;
}

// the kind of binning kernel that took the most time (direct, packed pass or its redo pass)
int gcb_counter_profile_read(gcb_counter *c, double *k1_ms, long *k1_launches, double *k1_atom_frames)
{
  double ms[GCB_PROF_KINDS], units[GCB_PROF_KINDS];
  long n[GCB_PROF_KINDS];
  int rc = gcb_counter_profile_kinds(c, ms, n, units);
  if (rc) return rc;
  int best = GCB_PROF_DIRECT;
  if (ms[GCB_PROF_PACKED] > ms[best]) best = GCB_PROF_PACKED;
  if (ms[GCB_PROF_REDO] > ms[best]) best = GCB_PROF_REDO;
  if (ms[GCB_PROF_PRIVATE] > ms[best]) best = GCB_PROF_PRIVATE;
  if (k1_ms) *k1_ms = ms[best];
  if (k1_launches) *k1_launches = n[best];
  if (k1_atom_frames) *k1_atom_frames = units[best];
  return GCB_OK;
}

int gcb_counter_timer_start(gcb_counter *c) { return gcb_counter_event_record(c, 0); }

int gcb_counter_timer_stop(gcb_counter *c, float *ms)
{
  int rc = gcb_counter_event_record(c, 1);
  if (rc) return rc;
  return gcb_counter_event_elapsed(c, 0, 1, ms);
}

int gcb_gen_device_frames(const gcb_gen *p, long frame0, long nframes, int natoms, float *x_dev, int device)
{
  if (!p || !x_dev || nframes < 0 || natoms <= 0) return gcb::fail(GCB_ERR_ARG, "gcb_gen_device_frames: bad argument");
  if (p->mode == GCB_GEN_CLUSTER && p->nsites <= 0) return gcb::fail(GCB_ERR_ARG, "cluster mode needs nsites > 0");
  GCB_CUDA(cudaSetDevice(device));
  if (nframes == 0) return GCB_OK;
  const long long total = (long long)nframes * natoms;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  gen_frames_kernel<<<grid, 256>>>(*p, frame0, nframes, natoms, x_dev);
  GCB_CUDA(cudaGetLastError());
  GCB_CUDA(cudaDeviceSynchronize());
  return GCB_OK;
}

int gcb_device_malloc(void **p, size_t bytes, int device)
{
  if (!p) return gcb::fail(GCB_ERR_ARG, "null argument");
  GCB_CUDA(cudaSetDevice(device));
  GCB_CUDA(cudaMalloc(p, bytes));
  return GCB_OK;
}
int gcb_device_free(void *p, int device)
{
  GCB_CUDA(cudaSetDevice(device));
  GCB_CUDA(cudaFree(p));
  return GCB_OK;
}
int gcb_host_alloc_pinned(void **p, size_t bytes)
{
  if (!p) return gcb::fail(GCB_ERR_ARG, "null argument");
  GCB_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
  return GCB_OK;
}
int gcb_host_free_pinned(void *p)
{
  GCB_CUDA(cudaFreeHost(p));
  return GCB_OK;
}
int gcb_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, int device)
{
  GCB_CUDA(cudaSetDevice(device));
  GCB_CUDA(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
  return GCB_OK;
}
int gcb_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, int device)
{
  GCB_CUDA(cudaSetDevice(device));
  GCB_CUDA(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
  return GCB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// multi-GPU reduce (the NCCL binding lives in nccl_reduce.cpp)
// ---------------------------------------------------------------------------
extern "C" int gcb_counter_comm_attach_(gcb_counter *c, void *comm, const gcb_reduce_ops *ops, int rank, int nranks, gcb_ctl *ctl)
{
  GCB_CUDA(cudaSetDevice(c->opts.device));
  if (comm) {
    if (c->nccl) return gcb::fail(GCB_ERR_STATE, "the counter already has a communicator");
    c->nccl = comm; c->ops = ops; c->rank = rank; c->nranks = nranks; c->ctl = ctl;
  }
  return GCB_OK;
}

// After this call every rank holds the global integer counts and statistics.
// Counts are an integer sum, so the result is independent of the rank count.
// Exact f32 occupancy needs one weight for the whole job (checked); with
// differing weights the ranks' f32 accumulators are summed instead, which is
// not the reference's frame-ordered chain (see gridcount_b200.h).
//
// What the call costs the host:
//   1. the ranks' {weight, uniformity, frame count, saturation bounds} table is host-side knowledge.  On one node it
//      is exchanged through shared memory (shm_ctl.cpp): no kernel, no stream, nothing that has to find an SM while
//      the persistent binning CTAs hold them all.  Ranks on several nodes fall back to a small NCCL all-reduce on a
//      side stream.
//   2. the per-frame hit counters never visit the host on their way: they are copied device to device into this
//      rank's block of the gather buffer, and the grid all-reduce and the (in-place) all-gather of the blocks go
//      out as ONE NCCL group behind the binning kernels.  The call returns after queueing.
extern "C" int gcb_counter_allreduce(gcb_counter *c)
{
  enum { U32 = 3, U64 = 5, F32 = 7, F64 = 8 };
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  if (c->nranks <= 1 || !c->nccl) return gcb_counter_sync(c);
  const gcb_reduce_ops *ops = c->ops;
  GCB_CUDA(cudaSetDevice(c->opts.device));
  static const bool trace = getenv("GCB_TRACE") != nullptr;
  auto now = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; };
  const double t_begin = now();
  int rc = finish_reduce(c);                  // a previous reduce whose statistics nobody asked for
  if (rc) return rc;
  rc = flush_private(c);                      // the grid that is reduced must hold everything
  if (rc) return rc;
  if (c->reduce_done) GCB_CUDA(cudaEventSynchronize(c->reduce_done));          // (or that a reset dropped): its D2H owns the pinned scratch
  for (auto &s : c->slots) { rc = collect_slot(c, s); if (rc) return rc; }     // host-staged batches report through pinned memory
  if (!c->comm_stream) GCB_CUDA(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
  // Persistent scratch (device + pinned mirror): no allocation per call.
  auto scratch = [&](size_t words) -> int {
    if (words <= c->scr_words) return GCB_OK;
    cudaFree(c->d_scr); cudaFreeHost(c->h_scr);
    c->d_scr = nullptr; c->h_scr = nullptr; c->scr_words = 0;
    const size_t cap = words + words / 2 + 64;
    GCB_CUDA(cudaMalloc(&c->d_scr, sizeof(unsigned long long) * cap));
    GCB_CUDA(cudaHostAlloc((void **)&c->h_scr, sizeof(unsigned long long) * cap, cudaHostAllocDefault));
    c->scr_words = cap;
    return GCB_OK;
  };
  // 1. each rank's {has frames, weight bits, uniform, f32 accumulator in use, frame count, saturation bounds}
  const size_t R = (size_t)c->nranks;
  constexpr size_t TW = GCB_CTL_WORDS;
  std::vector<unsigned long long> table(R * TW, 0ull);
  {
    uint32_t wb; memcpy(&wb, &c->w_first, 4);
    unsigned long long mine[TW] = {c->any ? 1ull : 0ull, wb, c->uniform ? 1ull : 0ull, c->acc_used ? 1ull : 0ull,
                                   c->hits_per_frame.size(), c->bound_run, c->bound_total, 0ull};
    if (c->ctl) {
      rc = gcb_ctl_allgather(c->ctl, (const uint64_t *)mine, (uint64_t *)table.data(), 0);    // no timeout: like a collective
      if (rc) return rc;
    } else {
      rc = scratch(R * TW);
      if (rc) return rc;
      memset(c->h_scr, 0, sizeof(unsigned long long) * R * TW);
      memcpy(c->h_scr + (size_t)c->rank * TW, mine, sizeof(mine));
      GCB_CUDA(cudaMemcpyAsync(c->d_scr, c->h_scr, sizeof(unsigned long long) * R * TW, cudaMemcpyHostToDevice, c->comm_stream));
      rc = ops->allreduce(c->d_scr, c->d_scr, R * TW, U64, c->nccl, c->comm_stream);     // disjoint slots: a sum is a gather
      if (rc) return rc;
      GCB_CUDA(cudaMemcpyAsync(c->h_scr, c->d_scr, sizeof(unsigned long long) * R * TW, cudaMemcpyDeviceToHost, c->comm_stream));
      GCB_CUDA(cudaStreamSynchronize(c->comm_stream));
      memcpy(table.data(), c->h_scr, sizeof(unsigned long long) * R * TW);
    }
  }
  bool global_uniform = true, any_acc = false, have_w = false;
  uint32_t w0 = 0;
  size_t total_frames = 0, my_off = 0, max_frames = 0;
  unsigned long long sum_bound_run = 0, sum_bound_total = 0;
  c->pend_frames.assign(R, 0);
  for (int r = 0; r < c->nranks; r++) {
    const unsigned long long *o = table.data() + (size_t)r * TW;
    if (r == c->rank) my_off = total_frames;
    total_frames += (size_t)o[4];
    max_frames = std::max(max_frames, (size_t)o[4]);
    c->pend_frames[(size_t)r] = (size_t)o[4];
    sum_bound_run += o[5]; sum_bound_total += o[6];
    if (!o[0]) continue;
    if (!o[2]) global_uniform = false;
    if (o[3]) any_acc = true;
    if (!have_w) { w0 = (uint32_t)o[1]; have_w = true; }
    else if (w0 != (uint32_t)o[1]) global_uniform = false;
  }
  const double t_gather = now();
  // 2. Statistics: one block [hits(max_frames) | weight bits(max_frames) | edge overflow] per rank, all-gathered in
  //    place.  Ranks hold contiguous frame shards in rank order, so with every rank's per-frame hit counts and
  //    weights each rank redoes the reference's frame-ordered `sumP += (double)weight` chain (g_ri3Dc.c:151,426):
  //    bit-identical to a single-GPU run.  Host-known parts of this rank's block go up from the pinned mirror;
  //    counters of device-resident batches and the overflow counter are copied in on the device, behind the kernels
  //    that produce them.
  const size_t blk = 2 * max_frames + 1;
  rc = scratch(R * blk);
  if (rc) return rc;
  unsigned long long *mine_h = c->h_scr + (size_t)c->rank * blk, *mine_d = c->d_scr + (size_t)c->rank * blk;
  memset(mine_h, 0, sizeof(unsigned long long) * blk);
  for (size_t f = 0; f < c->hits_per_frame.size(); f++) {
    mine_h[f] = c->hits_per_frame[f];
    uint32_t wbits; memcpy(&wbits, &c->weight_per_frame[f], 4);
    mine_h[max_frames + f] = wbits;
  }
  for (auto &h : c->devhits) for (long f = 0; f < h.n; f++) mine_h[h.frame0 + f] = 0;
  // the host-known part goes up on the side stream while the binning kernels are still running
  if (!c->scr_up) GCB_CUDA(cudaEventCreateWithFlags(&c->scr_up, cudaEventDisableTiming));
  if (!c->grp_done) GCB_CUDA(cudaEventCreateWithFlags(&c->grp_done, cudaEventDisableTiming));
  GCB_CUDA(cudaMemcpyAsync(mine_d, mine_h, sizeof(unsigned long long) * blk, cudaMemcpyHostToDevice, c->comm_stream));
  GCB_CUDA(cudaEventRecord(c->scr_up, c->comm_stream));
  GCB_CUDA(cudaStreamWaitEvent(c->compute, c->scr_up, 0));
  for (auto &h : c->devhits)
    GCB_CUDA(cudaMemcpyAsync(mine_d + h.frame0, h.d, sizeof(unsigned long long) * h.n, cudaMemcpyDeviceToDevice, c->compute));
  GCB_CUDA(cudaMemcpyAsync(mine_d + 2 * max_frames, c->d_diag, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->compute));
  // 3. the grids and the statistics, one NCCL group
  if (!(global_uniform && !any_acc)) {
    rc = fold_pending(c);
    if (rc) return rc;
    rc = ensure_acc(c);
    if (rc) return rc;
    sum_bound_total += sum_bound_run;            // every rank has folded its run into the totals by now
    sum_bound_run = 0;
  }
  {
    // saturation guard across ranks: if the ranks' bounds could add up beyond the counter range, every rank
    // clamps its cells to an equal share first (the same decision everywhere: it rests on the exchanged table)
    const unsigned long long share = std::min<unsigned long long>(c->clamp_at, c->clamp_ceiling / R);
    uint32_t *grid_cnt = (global_uniform && !any_acc) ? c->d_runcnt : c->d_total;
    unsigned long long &sum_bound = (global_uniform && !any_acc) ? sum_bound_run : sum_bound_total;
    if (sum_bound > c->clamp_ceiling) {
      clamp_counts_kernel<<<elementwise_grid(c, c->ncells), 256, 0, c->compute>>>(grid_cnt, c->ncells, (uint32_t)share, c->d_diag + 3);
      c->launches++;
      GCB_CUDA(cudaGetLastError());
      sum_bound = share * R;
    }
    c->bound_run = sum_bound_run;                 // bounds of the reduced grids
    c->bound_total = sum_bound_total;
  }
  rc = ops->group_start();
  if (rc) return rc;
  if (global_uniform && !any_acc) {
    // the common case: one integer grid per rank, one weight: sum the counts
    rc = ops->allreduce(c->d_runcnt, c->d_runcnt, c->ncells, U32, c->nccl, c->compute);
  } else {
    rc = ops->allreduce(c->d_acc, c->d_acc, c->ncells, F32, c->nccl, c->compute);
    if (!rc) rc = ops->allreduce(c->d_total, c->d_total, c->ncells, U32, c->nccl, c->compute);
  }
  if (!rc) rc = ops->allgather(mine_d, c->d_scr, blk, U64, c->nccl, c->compute);
  const int rc_end = ops->group_end();
  if (rc) return rc;
  if (rc_end) return rc_end;
  if (global_uniform && !any_acc) {
    if (have_w) { memcpy(&c->w_cur, &w0, 4); c->w_first = c->w_cur; c->have_run = true; c->any = true; }
  } else {
    c->uniform = false;
  }
  // the gathered statistics come back on the side stream: the next job's kernels do not queue behind the copy
  GCB_CUDA(cudaEventRecord(c->grp_done, c->compute));
  GCB_CUDA(cudaStreamWaitEvent(c->comm_stream, c->grp_done, 0));
  GCB_CUDA(cudaMemcpyAsync(c->h_scr, c->d_scr, sizeof(unsigned long long) * R * blk, cudaMemcpyDeviceToHost, c->comm_stream));
  // The grids are reduced in stream order; the host half (per-frame counters -> sumP chain) is finished when
  // somebody asks for it (finish_reduce), so the caller can queue the density and the next job meanwhile.
  if (!c->reduce_done) GCB_CUDA(cudaEventCreateWithFlags(&c->reduce_done, cudaEventDisableTiming));
  GCB_CUDA(cudaEventRecord(c->reduce_done, c->comm_stream));
  c->reduce_pending = true; c->pend_total = total_frames; c->pend_off = my_off; c->pend_ndev = c->devhits.size();
  c->pend_nlocal = c->hits_per_frame.size(); c->pend_blk = blk;
  c->global_frames = total_frames;
  c->reduced = true;
  if (trace && c->rank == 0)
    fprintf(stderr, "[gcb allreduce] rank table (%s) %.3f ms, queueing the grid/statistics group %.3f ms\n",
            c->ctl ? "shared memory" : "NCCL, beside the binning", t_gather - t_begin, now() - t_gather);
  return GCB_OK;
}

// host half of gcb_counter_allreduce: waits for the gathered statistics and redoes the frame-ordered chain
static int finish_reduce(gcb_counter *c)
{
  if (!c->reduce_pending) return GCB_OK;
  GCB_CUDA(cudaEventSynchronize(c->reduce_done));
  c->reduce_pending = false;
  const size_t blk = c->pend_blk, max_frames = (blk - 1) / 2;
  const unsigned long long *mine = c->h_scr + (size_t)c->rank * blk;
  // local bookkeeping that gcb_counter_sync would have done, for the batches the reduce covered
  const size_t ndev = std::min(c->pend_ndev, c->devhits.size());
  for (size_t i = 0; i < ndev; i++) {
    DevHits &h = c->devhits[i];
    for (long f = 0; f < h.n; f++) c->hits_per_frame[h.frame0 + f] = mine[h.frame0 + f];
    cudaFree(h.d_box3);
  }
  c->devhits.erase(c->devhits.begin(), c->devhits.begin() + (long)ndev);
  if (c->devhits.empty()) for (auto &a : c->hit_arenas) a.used = 0;
  c->sumP_frames = c->pend_nlocal;            // frames added after the reduce continue the chain from the global sum
  // consecutive frames of equal weight are one run of identical additions: one closed-form call per run
  double sp = 0;
  unsigned long long hits = 0, run_hits = 0, edge = 0;
  uint32_t run_w = 0;
  for (size_t r = 0; r < c->pend_frames.size(); r++) {
    const unsigned long long *fr = c->h_scr + r * blk;
    for (size_t f = 0; f < c->pend_frames[r]; f++) {
      const uint32_t wbits = (uint32_t)fr[max_frames + f];
      if (wbits != run_w && run_hits) {
        float wf; memcpy(&wf, &run_w, 4);
        sp = gcb::seq_add_f64(sp, (double)wf, run_hits);
        run_hits = 0;
      }
      run_w = wbits;
      run_hits += fr[f];
      hits += fr[f];
    }
    edge += fr[2 * max_frames];
  }
  if (run_hits) { float wf; memcpy(&wf, &run_w, 4); sp = gcb::seq_add_f64(sp, (double)wf, run_hits); }
  c->hits_total = hits; c->edge_overflow = edge; c->sumP = sp;
  return GCB_OK;
}
