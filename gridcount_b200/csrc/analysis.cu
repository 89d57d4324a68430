// K3: a_ri3Dc's reductions over a finished density grid (a_ri3Dc.c:133-201 crop, :488-762
// analysis) and a_gridcalc's weighted average (a_gridcalc.c:176-189), on the GPU.
//
// The reference sums in f32 in one fixed order (k outer, j, i inner; a_ri3Dc.c:567-703), so a
// tree or double-precision reduction would differ from it by the REFERENCE's rounding error
// (~1e-6..1e-5 relative on these sizes; its f32 `average` even saturates on large grids).  Every
// output element is therefore evaluated as the reference's sequential chain, bit for bit:
//
//   short chains: one thread per output, parallel ACROSS outputs
//     xyp/hxyp [mx][my]   chain over k            thread per (i,j), reads the x-fastest copy
//     xzp      [mx][mz]   chain over j            thread per (i,k), k fastest
//     yzp      [my][mz]   chain over i            thread per (j,k), k fastest
//     rzp/hrzp [r][mz]    chain over bin r's cells in (j,i) order, thread per (r,k)
//   long chains: parallel INSIDE the chain, by the exact scan of the rounding recurrence (seqsum.cuh)
//     zdf/lzdf/profile[k] chain over the slice's (j,i) cells      one block per (k, chain)
//     rdf/hrdf[r]         chain over k, then bin r's cells        one block per r
//     sumP (f64), average (f32): whole-grid chains                one cooperative grid, all SMs
//     rweight[r]          constant increments: closed form per exponent, on the host
//   The scans need finite cells >= 0 (any density); otherwise a flag is raised and the serial
//   one-thread-per-chain kernels (an_chain_kernel, an_slice_kernel, an_rdf_kernel) take over.
//
// Compiled with -fmad=false: ci*ci + cj*cj and the divide-then-add steps are separate f32
// operations as in the reference binary.
#include "gcb_internal.h"
#include "seqsum.cuh"
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

namespace {

const double kPi = 3.14159265358979323846;        // M_PI (utilgmx.h:23-28)

struct Dims { int mx, my, mz; };

// z-fastest [mx][my][mz] -> x-fastest [mz][my][mx]
__global__ void an_transpose_kernel(const float *__restrict__ in, float *__restrict__ out, Dims d)
{
  __shared__ float tile[32][33];
  const int j = blockIdx.z, i0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r, k = k0 + threadIdx.x;
    if (i < d.mx && k < d.mz) tile[r][threadIdx.x] = in[((size_t)i * d.my + j) * d.mz + k];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, i = i0 + threadIdx.x;
    if (i < d.mx && k < d.mz) out[((size_t)k * d.my + j) * d.mx + i] = tile[threadIdx.x][r];
  }
}

// readjust_tgrid's copy loop (a_ri3Dc.c:194-197)
__global__ void an_crop_kernel(const float *__restrict__ in, Dims o, float *__restrict__ out, Dims n, int bx, int by, int bz)
{
  const size_t total = (size_t)n.mx * n.my * n.mz;
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(c % n.mz);
    const size_t r = c / n.mz;
    const int j = (int)(r % n.my), i = (int)(r / n.my);
    out[c] = in[((size_t)(i + bx) * o.my + (j + by)) * o.mz + (k + bz)];
  }
}

// xyp[i][j] += g/((real)mz*U) over k; hxyp[i][j] += 1.0/((real)mz*1) when g <= min_occ  (a_ri3Dc.c:573-580)
__global__ void an_xy_kernel(const float *__restrict__ gx, Dims d, float d_xy, float min_occ, double h_inc,
                             float *__restrict__ xyp, float *__restrict__ hxyp)
{
  const int n = d.mx * d.my;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;        // t = j*mx + i
  if (t >= n) return;
  const int j = t / d.mx, i = t - j * d.mx;
  float s = 0.0f, h = 0.0f;
  for (int k = 0; k < d.mz; k++) {
    const float v = gx[(size_t)k * n + t];
    s = __fadd_rn(s, __fdiv_rn(v, d_xy));
    if (v <= min_occ) h = (float)((double)h + h_inc);
  }
  xyp[(size_t)i * d.my + j] = s;
  hxyp[(size_t)i * d.my + j] = h;
}

// xzp[i][k] += g/((real)my*U) over j   (a_ri3Dc.c:575-576)
__global__ void an_xz_kernel(const float *__restrict__ g, Dims d, float d_xz, float *__restrict__ xzp)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;        // t = i*mz + k
  if (t >= d.mx * d.mz) return;
  const int i = t / d.mz, k = t - i * d.mz;
  float s = 0.0f;
  for (int j = 0; j < d.my; j++) s = __fadd_rn(s, __fdiv_rn(g[((size_t)i * d.my + j) * d.mz + k], d_xz));
  xzp[t] = s;
}

// yzp[j][k] += g/((real)mx*U) over i   (a_ri3Dc.c:577-578)
__global__ void an_yz_kernel(const float *__restrict__ g, Dims d, float d_yz, float *__restrict__ yzp)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;        // t = j*mz + k
  if (t >= d.my * d.mz) return;
  const int j = t / d.mz, k = t - j * d.mz;
  float s = 0.0f;
  for (int i = 0; i < d.mx; i++) s = __fadd_rn(s, __fdiv_rn(g[((size_t)i * d.my + j) * d.mz + k], d_yz));
  yzp[t] = s;
}

struct SliceParams {
  float min_occ, d_zdf, DeltaR, U, a_z, delta_z;
  double dV;
};

// per-slice chains (a_ri3Dc.c:589-613, 661-695): zdf, lzdf, Rgyr2 -> profile; nocc[k] = occupied
// in-circle cells of the slice.  Tables are [j][i]: c2 = ci*ci+cj*cj, flags bit0 = bInCircle(i,j).
__global__ void an_slice_kernel(const float *__restrict__ g, Dims d, const float *__restrict__ c2tab,
                                const unsigned char *__restrict__ flags, SliceParams p, float *__restrict__ zdf,
                                float *__restrict__ lzdf, float *__restrict__ profile, unsigned *__restrict__ nocc)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= d.mz) return;
  float zs = 0.0f, lz = 0.0f, rg = 0.0f;
  unsigned n = 0;
  for (int j = 0; j < d.my; j++) {
    const float *row = c2tab + (size_t)j * d.mx;
    const unsigned char *frow = flags + (size_t)j * d.mx;
#pragma unroll 4
    for (int i = 0; i < d.mx; i++) {
      const float v = g[((size_t)i * d.my + j) * d.mz + k];
      if (v > p.min_occ && (frow[i] & 1)) {
        n++;
        lz = __fadd_rn(lz, v);
        rg = __fadd_rn(rg, __fmul_rn(v, row[i]));
      }
      zs = __fadd_rn(zs, v);
    }
  }
  const float z = (float)((double)p.a_z + ((double)k + 0.5) * (double)p.delta_z);        // a_ri3Dc.c:665
  zdf[2 * k] = z; lzdf[2 * k] = z; profile[2 * k] = z;
  zdf[2 * k + 1] = __fdiv_rn(zs, p.d_zdf);                                               // :666
  rg = (float)((double)rg / (0.5 * (double)lz));                                         // :672
  profile[2 * k + 1] = (float)(sqrt((double)rg) * (double)p.DeltaR);                     // :675
  const float Agyr = (float)(kPi * (double)rg);                                          // :677
  const float den = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(p.delta_z, Agyr), p.DeltaR), p.DeltaR), p.U);
  lzdf[2 * k + 1] = (float)((double)lz * p.dV / (double)den);                            // :695
  nocc[k] = n;
}

// rzp[r][k] / hrzp[r][k] raw sums (a_ri3Dc.c:640-649): thread per (r,k), chain over bin r's cells
__global__ void an_rz_kernel(const float *__restrict__ g, Dims d, const int *__restrict__ bin_start,
                             const int *__restrict__ bin_cells, int nr, float min_occ, float *__restrict__ rzp,
                             float *__restrict__ hrzp)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;        // t = r*mz + k
  if (t >= nr * d.mz) return;
  const int r = t / d.mz, k = t - r * d.mz;
  float s = 0.0f, h = 0.0f;
  for (int c = bin_start[r]; c < bin_start[r + 1]; c++) {
    const float v = g[(size_t)bin_cells[c] * d.mz + k];
    if (v > min_occ) s = __fadd_rn(s, v); else h = __fadd_rn(h, 1.0f);
  }
  rzp[t] = s; hrzp[t] = h;
}

// rdf[r][1], hrdf[r][1], rweight[r] raw sums (a_ri3Dc.c:627-645): thread per r, chain over k then cells
__global__ void an_rdf_kernel(const float *__restrict__ g, Dims d, const int *__restrict__ bin_start,
                              const int *__restrict__ bin_cells, int nr, float min_occ, double rw_inc,
                              float *__restrict__ rdf, float *__restrict__ hrdf, float *__restrict__ rweight)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nr) return;
  float s = 0.0f, h = 0.0f, rw = 0.0f;
  const int c0 = bin_start[r], c1 = bin_start[r + 1];
  for (int k = 0; k < d.mz; k++) {
#pragma unroll 4
    for (int c = c0; c < c1; c++) {
      const float v = g[(size_t)bin_cells[c] * d.mz + k];
      rw = (float)((double)rw + rw_inc);
      if (v > min_occ) s = __fadd_rn(s, v); else h = __fadd_rn(h, 1.0f);
    }
  }
  rdf[2 * r + 1] = s; hrdf[2 * r + 1] = h; rweight[r] = rw;
}

// ---------------------------------------------------------------------------------------------------
// Scan forms of the long chains (seqsum.cuh): same bits as the serial kernels above for grids whose
// cells are finite and >= 0; otherwise `bad_out` is raised and the host runs the serial kernels.
// They read the x-fastest copy gx[k][j][i], where a slice's (j,i) traversal is linear memory.
// ---------------------------------------------------------------------------------------------------
// One block per (slice k, chain): chain 0 = zdf, 1 = lzdf (and the count nocc), 2 = Rgyr2; raw sums go to
// raw[chain][k], an_slice_finish_kernel turns them into the three profiles.
#ifndef GCB_SLICE_THREADS
#define GCB_SLICE_THREADS 256
#endif
constexpr int SLICE_THREADS = GCB_SLICE_THREADS;
__global__ void __launch_bounds__(SLICE_THREADS)
an_slice_scan_kernel(const float *__restrict__ gx, Dims d, const float *__restrict__ c2tab,
                     const unsigned char *__restrict__ flags, float min_occ, float *__restrict__ raw,
                     unsigned *__restrict__ nocc, int *__restrict__ bad_out)
{
  __shared__ seqsum::BlockScratch scratch;
  __shared__ unsigned nocc_sh;
  const int k = blockIdx.x, chain = blockIdx.y;
  const size_t nxy = (size_t)d.mx * d.my;
  const float *sl = gx + (size_t)k * nxy;
  bool bad = false;
  float s;
  if (chain == 0) {
    s = seqsum::block_chain<float, 8, SLICE_THREADS>(nxy, [&](size_t t) { return t < nxy ? __ldg(sl + t) : 0.0f; }, scratch, &bad);
  } else if (chain == 1) {
    if (threadIdx.x == 0) nocc_sh = 0;
    s = seqsum::block_chain<float, 8, SLICE_THREADS>(nxy, [&](size_t t) {
      if (t >= nxy) return 0.0f;
      const float v = __ldg(sl + t);
      return (v > min_occ && (__ldg(flags + t) & 1)) ? v : 0.0f; }, scratch, &bad);
    unsigned n = 0;
    for (size_t t = threadIdx.x; t < nxy; t += SLICE_THREADS) n += (__ldg(sl + t) > min_occ && (__ldg(flags + t) & 1)) ? 1u : 0u;
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&nocc_sh, n);
    __syncthreads();
    if (threadIdx.x == 0) nocc[k] = nocc_sh;
  } else {
    s = seqsum::block_chain<float, 8, SLICE_THREADS>(nxy, [&](size_t t) {
      if (t >= nxy) return 0.0f;
      const float v = __ldg(sl + t);
      return (v > min_occ && (__ldg(flags + t) & 1)) ? __fmul_rn(v, __ldg(c2tab + t)) : 0.0f; }, scratch, &bad);
  }
  if (threadIdx.x == 0) { raw[(size_t)chain * d.mz + k] = s; if (bad) atomicExch(bad_out, 1); }
}

__global__ void an_slice_finish_kernel(const float *__restrict__ raw, int mz, SliceParams p, float *__restrict__ zdf,
                                       float *__restrict__ lzdf, float *__restrict__ profile)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= mz) return;
  const float zs = raw[k], lz = raw[mz + k];
  float rg = raw[2 * mz + k];
  const float z = (float)((double)p.a_z + ((double)k + 0.5) * (double)p.delta_z);        // a_ri3Dc.c:665
  zdf[2 * k] = z; lzdf[2 * k] = z; profile[2 * k] = z;
  zdf[2 * k + 1] = __fdiv_rn(zs, p.d_zdf);                                               // :666
  rg = (float)((double)rg / (0.5 * (double)lz));                                         // :672
  profile[2 * k + 1] = (float)(sqrt((double)rg) * (double)p.DeltaR);                     // :675
  const float Agyr = (float)(kPi * (double)rg);                                          // :677
  const float den = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(p.delta_z, Agyr), p.DeltaR), p.DeltaR), p.U);
  lzdf[2 * k + 1] = (float)((double)lz * p.dV / (double)den);                            // :695
}

// One block per radial bin r: rdf[r][1] chains the occupied cells over k, then bin r's cells in (j,i)
// order (a_ri3Dc.c:627-645); hrdf[r][1] is a chain of +1.0f over the other cells, i.e. min(count, 2^24).
// bin_xy[] lists each bin's cells as j*mx + i.
#ifndef GCB_RDF_THREADS
#define GCB_RDF_THREADS 1024
#endif
constexpr int RDF_THREADS = GCB_RDF_THREADS;
__global__ void __launch_bounds__(RDF_THREADS)
an_rdf_scan_kernel(const float *__restrict__ gx, Dims d, const int *__restrict__ bin_start,
                   const int *__restrict__ bin_xy, float min_occ, float *__restrict__ rdf, float *__restrict__ hrdf,
                   int *__restrict__ bad_out)
{
  __shared__ seqsum::BlockScratch scratch;
  __shared__ unsigned long long holes_sh;
  const int r = blockIdx.x;
  const int c0 = bin_start[r];
  const unsigned nc = (unsigned)(bin_start[r + 1] - c0);
  const size_t nxy = (size_t)d.mx * d.my, n = (size_t)nc * d.mz;
  if (threadIdx.x == 0) holes_sh = 0;
  bool bad = false;
  const bool small = n <= 0xffffffffull;
  auto cell = [&](size_t t) {
    size_t k, c;
    if (small) { const uint32_t kk = (uint32_t)t / nc; k = kk; c = (uint32_t)t - kk * nc; }
    else { k = t / nc; c = t - k * nc; }
    return __ldg(gx + k * nxy + (size_t)__ldg(bin_xy + c0 + c));
  };
  const float s = seqsum::block_chain<float, 8, RDF_THREADS>(
      n, [&](size_t t) { if (t >= n) return 0.0f; const float v = cell(t); return v > min_occ ? v : 0.0f; }, scratch, &bad);
  unsigned long long holes = 0;
  for (size_t t = threadIdx.x; t < n; t += RDF_THREADS) holes += cell(t) > min_occ ? 0u : 1u;
  holes = __reduce_add_sync(0xffffffffu, (unsigned)holes);
  if ((threadIdx.x & 31) == 0 && holes) atomicAdd(&holes_sh, holes);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (bad) atomicExch(bad_out, 1);
    rdf[2 * r + 1] = s;
    hrdf[2 * r + 1] = holes_sh > (1ull << 24) ? 16777216.0f : (float)holes_sh;
  }
}

// Whole-grid chains in (k,j,i) order over the x-fastest copy, which is linear memory order.
// block 0: sumP += (double)g (a_ri3Dc.c:499-508) and the count behind sumH
// block 1: average += g for cells with irad < iradNR (a_ri3Dc.c:633); masked cells add +0.0f,
//          which leaves any partial sum that started at +0 unchanged bit for bit
#ifndef GCB_SCAN_THREADS
#define GCB_SCAN_THREADS 1024
#endif
constexpr int SCAN_THREADS = GCB_SCAN_THREADS;
constexpr int SCAN_T = 8;
__global__ void __launch_bounds__(SCAN_THREADS)
an_chain_scan_kernel(const float *__restrict__ gx, size_t ncells, int nxy, const unsigned char *__restrict__ flags,
                     float min_occ_raw, double *__restrict__ sumP_out, unsigned long long *__restrict__ nholes_out,
                     float *__restrict__ average_out, int *__restrict__ bad_out)
{
  __shared__ seqsum::BlockScratch scratch;
  bool bad = false;
  if (blockIdx.x == 0) {
    const double s = seqsum::block_chain<double, SCAN_T, SCAN_THREADS>(
        ncells, [&](size_t i) { return i < ncells ? __ldg(gx + i) : 0.0f; }, scratch, &bad);
    if (threadIdx.x == 0) {
      if (bad) atomicExch(bad_out, 1);
      *sumP_out = s;
    }
  } else if (blockIdx.x == 1) {
    const bool small = ncells <= 0xffffffffull;
    const float s = seqsum::block_chain<float, SCAN_T, SCAN_THREADS>(
        ncells,
        [&](size_t i) {
          if (i >= ncells) return 0.0f;
          const size_t t = small ? (size_t)((uint32_t)i % (uint32_t)nxy) : i % (size_t)nxy;
          return (__ldg(flags + t) & 2) ? __ldg(gx + i) : 0.0f;
        },
        scratch, &bad);
    if (threadIdx.x == 0) {
      if (bad) atomicExch(bad_out, 1);
      *average_out = s;
    }
  } else {                                  // the count behind sumH (a_ri3Dc.c:504): order-free
    unsigned long long holes = 0;
    for (size_t i = (size_t)(blockIdx.x - 2) * SCAN_THREADS + threadIdx.x; i < ncells; i += (size_t)(gridDim.x - 2) * SCAN_THREADS)
      holes += __ldg(gx + i) <= min_occ_raw ? 1u : 0u;
    holes = __reduce_add_sync(0xffffffffu, (unsigned)holes);          // per-lane counts stay far below 2^32/32
    if ((threadIdx.x & 31) == 0 && holes) atomicAdd(nholes_out, holes);
  }
}

// The same two chains spread over a cooperative grid (seqsum::GridChain): one block per SM, rounds
// separated by grid-wide barriers.  fbuf holds 2 (round parity) x 2 (chains) x maxseg segment maps.
__global__ void __launch_bounds__(SCAN_THREADS)
an_chain_grid_kernel(const float *__restrict__ gx, size_t ncells, int nxy, const unsigned char *__restrict__ flags,
                     float min_occ_raw, seqsum::Fn *__restrict__ fbuf, size_t maxseg, double *__restrict__ sumP_out,
                     unsigned long long *__restrict__ nholes_out, float *__restrict__ average_out, int *__restrict__ bad_out)
{
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ seqsum::BlockScratch sh;
  if (threadIdx.x == 0) { sh.first_cross = SCAN_THREADS; sh.bad = 0; }
  __syncthreads();
  {                                         // the count behind sumH (a_ri3Dc.c:504): order-free
    unsigned holes = 0;
    for (size_t i = (size_t)blockIdx.x * SCAN_THREADS + threadIdx.x; i < ncells; i += (size_t)gridDim.x * SCAN_THREADS)
      holes += __ldg(gx + i) <= min_occ_raw ? 1u : 0u;
    holes = __reduce_add_sync(0xffffffffu, holes);
    if ((threadIdx.x & 31) == 0 && holes) atomicAdd(nholes_out, (unsigned long long)holes);
  }
  auto load_all = [&](size_t i, float (&x)[SCAN_T]) {
#pragma unroll
    for (int j = 0; j < SCAN_T; j++) x[j] = i + j < ncells ? __ldg(gx + i + j) : 0.0f;
  };
  auto load_cyl = [&](size_t i, float (&x)[SCAN_T]) {              // cells with irad < iradNR (flag bit 1), else +0
    unsigned t = i < ncells ? (unsigned)(i % (size_t)nxy) : 0u;
#pragma unroll
    for (int j = 0; j < SCAN_T; j++) {
      x[j] = (i + j < ncells && (__ldg(flags + t) & 2)) ? __ldg(gx + i + j) : 0.0f;
      if (++t == (unsigned)nxy) t = 0;
    }
  };
  seqsum::GridChain<double, SCAN_T, SCAN_THREADS> A;
  seqsum::GridChain<float, SCAN_T, SCAN_THREADS> B;
  A.init(ncells); B.init(ncells);
  bool bad = false;
  A.prologue(ncells, load_all, sh, &bad, (size_t)4 * SCAN_THREADS * SCAN_T);
  B.prologue(ncells, load_cyl, sh, &bad, (size_t)4 * SCAN_THREADS * SCAN_T);
  if (bad) {                                                         // every block sees the same terms: uniform
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(bad_out + 1, 1);
    return;
  }
  for (int round = 0; !(A.done && B.done); round++) {
    seqsum::Fn *fa = fbuf + (size_t)(round & 1) * 2 * maxseg, *fb = fa + maxseg;
    A.phase1(ncells, load_all, fa, sh, bad_out);
    B.phase1(ncells, load_cyl, fb, sh, bad_out);
    grid.sync();
    if (*(volatile int *)bad_out) return;                            // same value in every block: set before the barrier
    A.phase2(ncells, load_all, fa, sh, &bad);
    B.phase2(ncells, load_cyl, fb, sh, &bad);
    if (bad) break;                                                  // every block computes the same `bad`
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (bad) atomicExch(bad_out + 1, 1);                             // a second word: bad_out[0] is read between barriers
    *sumP_out = A.value(); *average_out = B.value();
  }
}

// the serial form of the same two chains: one consumer thread each, fed through shared memory by the
// rest of its block.  Used when the grid holds negative or non-finite cells (never for a density).
constexpr int CHAIN_THREADS = 256;
constexpr int CHAIN_CHUNK = 2048;
__global__ void __launch_bounds__(CHAIN_THREADS)
an_chain_kernel(const float *__restrict__ gx, size_t ncells, int nxy, const unsigned char *__restrict__ flags,
                float min_occ_raw, double *__restrict__ sumP_out, unsigned long long *__restrict__ nholes_out,
                float *__restrict__ average_out)
{
  __shared__ __align__(16) float buf[2][CHAIN_CHUNK];
  __shared__ unsigned long long holes_sh;
  const bool is_avg = blockIdx.x == 1;
  const int tid = threadIdx.x;
  constexpr int PER = CHAIN_CHUNK / CHAIN_THREADS;
  if (tid == 0) holes_sh = 0;
  unsigned long long holes = 0;
  double sd = 0.0;
  float sf = 0.0f;
  const size_t nchunks = (ncells + CHAIN_CHUNK - 1) / CHAIN_CHUNK;
  float regs[PER];
  auto load = [&](size_t c) {
#pragma unroll
    for (int r = 0; r < PER; r++) {
      const size_t idx = c * CHAIN_CHUNK + (size_t)r * CHAIN_THREADS + tid;
      float v = 0.0f;
      if (idx < ncells) {
        v = gx[idx];
        if (is_avg) { if (!(flags[idx % (size_t)nxy] & 2)) v = 0.0f; }
        else if (v <= min_occ_raw) holes++;
      }
      regs[r] = v;
    }
  };
  auto stash = [&](int b) {
#pragma unroll
    for (int r = 0; r < PER; r++) buf[b][r * CHAIN_THREADS + tid] = regs[r];
  };
  load(0);
  stash(0);
  __syncthreads();
  for (size_t c = 0; c < nchunks; c++) {
    const int b = (int)(c & 1);
    if (c + 1 < nchunks) load(c + 1);
    if (tid == 0) {
      const size_t left = ncells - c * CHAIN_CHUNK;
      const int n = left < (size_t)CHAIN_CHUNK ? (int)left : CHAIN_CHUNK;
      const float4 *q = reinterpret_cast<const float4 *>(buf[b]);
      int e = 0;
      if (is_avg) {
        for (; e + 4 <= n; e += 4) {
          const float4 v = q[e >> 2];
          sf = __fadd_rn(sf, v.x); sf = __fadd_rn(sf, v.y); sf = __fadd_rn(sf, v.z); sf = __fadd_rn(sf, v.w);
        }
        for (; e < n; e++) sf = __fadd_rn(sf, buf[b][e]);
      } else {
        for (; e + 4 <= n; e += 4) {
          const float4 v = q[e >> 2];
          sd = __dadd_rn(sd, (double)v.x); sd = __dadd_rn(sd, (double)v.y);
          sd = __dadd_rn(sd, (double)v.z); sd = __dadd_rn(sd, (double)v.w);
        }
        for (; e < n; e++) sd = __dadd_rn(sd, (double)buf[b][e]);
      }
    }
    if (c + 1 < nchunks) stash(b ^ 1);
    __syncthreads();
  }
  if (!is_avg) {
    atomicAdd(&holes_sh, holes);
    __syncthreads();
    if (tid == 0) { *sumP_out = sd; *nholes_out = holes_sh; }
  } else if (tid == 0) {
    *average_out = sf;
  }
}

// a_gridcalc.c:180-189: a += g * w with w double, f32 store
__global__ void gridcalc_axpy_kernel(float *__restrict__ avg, const float *__restrict__ g, double w, size_t n)
{
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x)
    avg[c] = (float)__dadd_rn((double)avg[c], __dmul_rn((double)g[c], w));
}

// a_ri3Dc.c:128-130
float discrete_adjust(float x_old, float x_new, float delta)
{
  const float q = (x_old - x_new) / delta;
  const int n = -(int)((double)q + (x_old > x_new ? 0.5 : -0.5));
  return (float)n * delta;
}

struct DevBuf {
  void *p = nullptr;
  bool owned = true;                   // false: a view into an Arena
  ~DevBuf() { if (p && owned) cudaFree(p); }
  template <class T> T *as() const { return static_cast<T *>(p); }
};

// all device buffers of one gcb_analyse call in a single allocation
// Per device: one allocation (up to CACHE_MAX: a 500^3 grid with its x-fastest copy), four streams, pinned staging
// buffers and the per-(i,j) tables of the last geometry are kept between calls: a C2-sized analysis is 1 ms of
// kernels and a C3-sized one a few ms, while cudaMalloc/cudaFree of the grid copies, stream creation, rebuilding the
// tables and pageable copies cost several times that.  gcb_analyse_release_cache() gives everything back.  A second
// thread analysing on the same device at the same time simply allocates its own.
struct TableKey {
  int mx = 0, my = 0, mz = 0, iradNR = -1;
  uint32_t radius_bits = 0, deltaR_bits = 0;
  bool operator==(const TableKey &o) const
  {
    return mx == o.mx && my == o.my && mz == o.mz && iradNR == o.iradNR && radius_bits == o.radius_bits && deltaR_bits == o.deltaR_bits;
  }
};
struct DeviceCache {
  std::mutex m;
  char *base = nullptr;
  size_t cap = 0;
  cudaStream_t st[4] = {nullptr, nullptr, nullptr, nullptr};
  // tables of the last geometry: valid while `base` has not been reallocated (they sit at the front of the arena)
  bool tables_valid = false;
  TableKey key;
  std::vector<int> bin_start;
  std::vector<float> rweight;
  // pinned staging: results on their way back, tables on their way up
  char *h_stage = nullptr;
  size_t h_stage_cap = 0;
  // pinned chunks for the upload of a pageable host grid
  static constexpr int NCHUNK = 16;
  static constexpr size_t CHUNK = (size_t)16 << 20;
  char *h_chunk[NCHUNK] = {nullptr};
  cudaEvent_t chunk_done[NCHUNK] = {nullptr};
  cudaStream_t chunk_st[NCHUNK] = {nullptr};
};
DeviceCache g_cache[64];
constexpr size_t CACHE_MAX = (size_t)1536 << 20;

struct Arena {
  char *base = nullptr;
  size_t total = 0;
  DeviceCache *cache = nullptr;        // locked while this arena borrows its allocation
  std::vector<std::pair<DevBuf *, size_t>> req;
  void want(DevBuf &b, size_t bytes)
  {
    req.emplace_back(&b, total);
    total += (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
  }
  cudaError_t commit(int device)
  {
    cudaError_t e = cudaSuccess;
    DeviceCache &dc = g_cache[device & 63];
    if (total <= CACHE_MAX && dc.m.try_lock()) {
      cache = &dc;
      if (dc.cap < total) {
        if (dc.base) cudaFree(dc.base);
        dc.base = nullptr; dc.cap = 0; dc.tables_valid = false;
        e = cudaMalloc(&dc.base, total);
        if (e == cudaSuccess) dc.cap = total;
      }
      base = dc.base;
    } else {
      e = cudaMalloc(&base, total);
    }
    if (e == cudaSuccess) for (auto &r : req) { r.first->p = base + r.second; r.first->owned = false; }
    return e;
  }
  ~Arena()
  {
    if (cache) cache->m.unlock();
    else if (base) cudaFree(base);
  }
};

// the four analysis streams of a device: cached when the cache is free, else created for the call
struct Streams {
  cudaStream_t st[4] = {nullptr, nullptr, nullptr, nullptr};
  bool owned = false;
  cudaError_t get(DeviceCache *cache)
  {
    cudaStream_t *src = cache ? cache->st : st;
    owned = cache == nullptr;
    for (int i = 0; i < 4; i++) {
      if (!src[i]) { const cudaError_t e = cudaStreamCreateWithFlags(&src[i], cudaStreamNonBlocking); if (e != cudaSuccess) return e; }
      st[i] = src[i];
    }
    return cudaSuccess;
  }
  ~Streams() { if (owned) for (auto &s : st) if (s) cudaStreamDestroy(s); }
};

// pinned staging of at least `bytes` (cached), or NULL: the caller then uses pageable memory
char *stage_pinned(DeviceCache *cache, size_t bytes)
{
  if (!cache) return nullptr;
  if (cache->h_stage_cap < bytes) {
    if (cache->h_stage) cudaFreeHost(cache->h_stage);
    cache->h_stage = nullptr; cache->h_stage_cap = 0;
    const size_t cap = bytes + bytes / 4 + 4096;
    if (cudaHostAlloc((void **)&cache->h_stage, cap, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); cache->h_stage = nullptr; return nullptr; }
    cache->h_stage_cap = cap;
  }
  return cache->h_stage;
}

// Host grid -> device.  Pinned memory goes in one DMA.  Pageable memory would be staged by the driver at ~8 GB/s
// (32 ms for a 400^3 grid, six times the kernels): instead up to 8 threads copy 16 MB chunks into cached pinned
// buffers and queue their DMA themselves, so host copies and DMA of different chunks overlap.
int upload_grid(DeviceCache *cache, int device, void *dst, const float *src, size_t bytes)
{
  cudaPointerAttributes attr;
  bool pinned = false;
  if (cudaPointerGetAttributes(&attr, src) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
  else cudaGetLastError();
  if (pinned || !cache || bytes < 4 * DeviceCache::CHUNK) {
    if (cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess)
      return gcb::fail(GCB_ERR_CUDA, "copying the grid to the device failed: %s", cudaGetErrorString(cudaGetLastError()));
    return GCB_OK;
  }
  for (int i = 0; i < DeviceCache::NCHUNK; i++) {
    if (!cache->h_chunk[i]) {
      if (cudaHostAlloc((void **)&cache->h_chunk[i], DeviceCache::CHUNK, cudaHostAllocDefault) != cudaSuccess ||
          cudaEventCreateWithFlags(&cache->chunk_done[i], cudaEventDisableTiming) != cudaSuccess ||
          cudaStreamCreateWithFlags(&cache->chunk_st[i], cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        if (cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess)
          return gcb::fail(GCB_ERR_CUDA, "copying the grid to the device failed: %s", cudaGetErrorString(cudaGetLastError()));
        return GCB_OK;
      }
    }
  }
  const size_t nchunks = (bytes + DeviceCache::CHUNK - 1) / DeviceCache::CHUNK;
  const int nthreads = (int)std::min<size_t>(DeviceCache::NCHUNK / 2, std::min<size_t>(nchunks, std::max(1u, std::thread::hardware_concurrency() / 2)));
  std::vector<int> rcs((size_t)nthreads, 0);
  auto work = [&](int t) {
    if (cudaSetDevice(device) != cudaSuccess) { rcs[(size_t)t] = 1; return; }
    const int slots[2] = {t, t + DeviceCache::NCHUNK / 2};   // thread t owns two pinned buffers (nthreads <= NCHUNK / 2)
    int turn = 0;
    for (size_t c = (size_t)t; c < nchunks; c += (size_t)nthreads, turn++) {
      const int slot = slots[turn & 1];
      if (turn >= 2 && cudaEventSynchronize(cache->chunk_done[slot]) != cudaSuccess) { rcs[(size_t)t] = 1; return; }
      const size_t off = c * DeviceCache::CHUNK, n = std::min(DeviceCache::CHUNK, bytes - off);
      memcpy(cache->h_chunk[slot], reinterpret_cast<const char *>(src) + off, n);
      if (cudaMemcpyAsync(static_cast<char *>(dst) + off, cache->h_chunk[slot], n, cudaMemcpyHostToDevice, cache->chunk_st[slot]) != cudaSuccess ||
          cudaEventRecord(cache->chunk_done[slot], cache->chunk_st[slot]) != cudaSuccess) { rcs[(size_t)t] = 1; return; }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
  work(0);
  for (auto &x : th) x.join();
  bool bad = false;
  for (int r : rcs) bad |= r != 0;
  for (int i = 0; i < DeviceCache::NCHUNK; i++) if (cudaStreamSynchronize(cache->chunk_st[i]) != cudaSuccess) bad = true;
  if (bad) return gcb::fail(GCB_ERR_CUDA, "copying the grid to the device failed: %s", cudaGetErrorString(cudaGetLastError()));
  return GCB_OK;
}

#define AN_CUDA(call)                                                                            \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) {                                          \
      gcb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return GCB_ERR_CUDA; } } while (0)

template <class T> float *host_copy(const DevBuf &b, size_t n, int *rc)
{
  T *h = (T *)malloc(std::max<size_t>(n, 1) * sizeof(T));
  if (!h) { *rc = gcb::fail(GCB_ERR_NOMEM, "out of memory"); return nullptr; }
  if (n && cudaMemcpy(h, b.p, n * sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess) {
    free(h);
    *rc = gcb::fail(GCB_ERR_CUDA, "copying an analysis result back failed: %s", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return (float *)h;
}

int blocks_for(size_t n, int threads) { return (int)std::max<size_t>(1, (n + threads - 1) / threads); }

}  // namespace

// `on_device`: grid_zfast is device memory on opts->device (a counter's density, gcb_counter_device_density),
// otherwise host memory.  out->grid is filled only when the analysis crops.
static int analyse_impl(const gcb_tgrid *tg_in, const float *grid_zfast, bool on_device, const gcb_analysis_opts *opts,
                        gcb_analysis *out)
{
  if (!tg_in || !grid_zfast || !opts || !out) return gcb::fail(GCB_ERR_ARG, "gcb_analyse: null argument");
  if (opts->unit < 0 || opts->unit >= GCB_UNIT_NR) return gcb::fail(GCB_ERR_ARG, "gcb_analyse: unknown unit %d", opts->unit);
  if (tg_in->mx[0] <= 0 || tg_in->mx[1] <= 0 || tg_in->mx[2] <= 0) return gcb::fail(GCB_ERR_ARG, "gcb_analyse: empty grid");
  if (gcb_device_count() <= 0) return gcb::fail(GCB_ERR_CUDA, "no CUDA device: libgridcount_b200 has no CPU fallback");
  memset(out, 0, sizeof(*out));
  AN_CUDA(cudaSetDevice(opts->device));

  // ---- geometry: update_tgrid + readjust_tgrid (a_ri3Dc.c:473-486) on the host ----------------
  gcb_tgrid old = *tg_in;
  gcb_update_tgrid(&old);                                     // b = a + mx*Delta (a_ri3Dc.c:474)
  gcb_tgrid tg = old;
  gcb_cavity user{};
  user.radius = opts->radius; user.z1 = opts->z1; user.z2 = opts->z2;
  bool cropped = false;
  int base[3] = {0, 0, 0};
  {
    const float dsum = tg.delta[0] + tg.delta[1];
    const float DeltaR = (float)(0.5 * (double)dsum);
    const float R_new = user.radius;
    if (opts->setmask & GCB_SET_Z1) { tg.a[2] += discrete_adjust(tg.a[2], user.z1, tg.delta[2]); cropped = true; }
    if (opts->setmask & GCB_SET_Z2) { tg.b[2] += discrete_adjust(tg.b[2], user.z2, tg.delta[2]); cropped = true; }
    if (tg.a[2] >= tg.b[2])
      return gcb::fail(GCB_ERR_RANGE, "FAILURE: Apparently z2 =< z1, or either of them was too large for the given grid.");
    gcb_tgrid2cavity(&tg, &user);
    if (opts->setmask & GCB_SET_R) { user.radius += discrete_adjust(user.radius, R_new, DeltaR); cropped = true; }
    if (cropped) {
      int rc = gcb_setup_tgrid(&tg, &user, old.delta);
      if (rc) return gcb::fail(GCB_ERR_RANGE, "FAILURE while readjusting the grid.");
      tg.tweight = old.tweight;
      for (int d = 0; d < 3; d++) {
        base[d] = (int)((tg.a[d] - old.a[d]) / tg.delta[d]);
        if (base[d] < 0 || base[d] + tg.mx[d] > old.mx[d])
          return gcb::fail(GCB_ERR_RANGE, "readjusted grid does not lie inside the stored one (dimension %d)", d);
      }
    }
  }
  const Dims od{old.mx[0], old.mx[1], old.mx[2]}, d{tg.mx[0], tg.mx[1], tg.mx[2]};
  const size_t ocells = (size_t)od.mx * od.my * od.mz, ncells = (size_t)d.mx * d.my * d.mz;
  const int nxy = d.mx * d.my;
  const float fmx = (float)d.mx, fmy = (float)d.my, fmz = (float)d.mz;

  gcb_cavity geometry{};
  gcb_tgrid2cavity(&tg, &geometry);                           // a_ri3Dc.c:486
  const float DeltaR = (float)(0.5 * (double)(tg.delta[0] + tg.delta[1]));      // :488
  const double dV = gcb_delta_v(&tg);                         // :495
  const float U = gcb_dens_unit(opts->unit, &tg);             // DensUnit[nunit], :516
  const float Uunity = gcb_dens_unit(GCB_UNIT_UNITY, nullptr);
  const float min_occ_raw = opts->min_occ;                    // sumH uses the unscaled value (:504)
  const float min_occ = min_occ_raw * U;                      // :517
  int iradNR = (int)floor((double)(geometry.radius / DeltaR));                  // :544
  if (iradNR < 0) iradNR = 0;

  static const bool trace = getenv("GCB_TRACE") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t_start = now();
  const double rw_inc = 1.0 / (double)fmz;                                 // a_ri3Dc.c:637
  // ---- device buffers: the tables first, so that they stay where they are from call to call --------
  DevBuf d_in, d_g, d_gx, d_c2, d_flags, d_bs, d_bc, d_bxy, d_fbuf, d_sraw, d_xyp, d_hxyp, d_xzp, d_yzp, d_zdf, d_lzdf, d_prof, d_nocc,
      d_rzp, d_hrzp, d_rdf, d_hrdf, d_rw, d_scal;
  Arena arena;
  const size_t nrz = (size_t)iradNR * d.mz;
  const size_t maxseg = (ncells + (size_t)SCAN_THREADS * SCAN_T - 1) / ((size_t)SCAN_THREADS * SCAN_T) + 1;
  arena.want(d_c2, (size_t)nxy * 4);
  arena.want(d_flags, (size_t)nxy);
  arena.want(d_bs, ((size_t)iradNR + 1) * 4);
  arena.want(d_bc, (size_t)nxy * 4);              // (at most every (i,j) column lies in a radial bin)
  arena.want(d_bxy, (size_t)nxy * 4);
  const size_t table_span = arena.total;
  arena.want(d_xyp, (size_t)nxy * 4);
  arena.want(d_hxyp, (size_t)nxy * 4);
  arena.want(d_xzp, (size_t)d.mx * d.mz * 4);
  arena.want(d_yzp, (size_t)d.my * d.mz * 4);
  arena.want(d_zdf, (size_t)d.mz * 8);
  arena.want(d_lzdf, (size_t)d.mz * 8);
  arena.want(d_prof, (size_t)d.mz * 8);
  arena.want(d_nocc, (size_t)d.mz * 4);
  arena.want(d_sraw, (size_t)d.mz * 12);
  arena.want(d_rzp, nrz * 4);
  arena.want(d_hrzp, nrz * 4);
  arena.want(d_rw, ((size_t)iradNR + 1) * 4);
  arena.want(d_rdf, ((size_t)iradNR + 1) * 8);    // d_rdf .. the head of d_fbuf are zeroed in one go
  arena.want(d_hrdf, ((size_t)iradNR + 1) * 8);
  arena.want(d_scal, 32);
  arena.want(d_fbuf, 16 + 4 * maxseg * sizeof(seqsum::Fn));
  arena.want(d_gx, ncells * 4);
  if (cropped) arena.want(d_g, ncells * 4);
  if (!on_device) arena.want(d_in, ocells * 4);
  AN_CUDA(arena.commit(opts->device));
  DeviceCache *cache = arena.cache;

  // ---- per-(i,j) tables (a_ri3Dc.c:586-592, 625), [j][i] order = the traversal order; kept per device for the
  //      last geometry (host: bin_start, rweight; device: all five tables) ------------------------------
  TableKey key;
  key.mx = d.mx; key.my = d.my; key.mz = d.mz; key.iradNR = iradNR;
  memcpy(&key.radius_bits, &geometry.radius, 4); memcpy(&key.deltaR_bits, &DeltaR, 4);
  std::vector<int> bin_start_local;
  std::vector<float> rweight_local;
  const bool tables_hit = cache && cache->tables_valid && cache->key == key;
  if (!tables_hit) {
    std::vector<float> c2tab((size_t)nxy);
    std::vector<unsigned char> flags((size_t)nxy);
    std::vector<int> irad((size_t)nxy);
    {
      const float rcirc = geometry.radius / DeltaR;
      const float cx = (float)((double)fmx / 2.0), cy = (float)((double)fmy / 2.0);
      for (int j = 0; j < d.my; j++) {
        const float cj = (float)(((double)(float)j + 0.5) - (double)fmy / 2.0);
        for (int i = 0; i < d.mx; i++) {
          const float ci = (float)(((double)(float)i + 0.5) - (double)fmx / 2.0);
          const float c2 = ci * ci + cj * cj;
          const float ex = (float)i - cx, ey = (float)j - cy;
          const int ir = (int)floor(sqrt((double)c2));
          const size_t t = (size_t)j * d.mx + i;
          c2tab[t] = c2;
          irad[t] = ir;
          flags[t] = (unsigned char)((ex * ex + ey * ey <= rcirc * rcirc ? 1 : 0) | (ir < iradNR ? 2 : 0));
        }
      }
    }
    // cells of each radial bin in (j,i) order, as z-fastest row numbers i*my + j and as x-fastest j*mx + i
    std::vector<int> bin_start((size_t)iradNR + 1, 0), bin_cells, bin_xy;
    {
      for (size_t t = 0; t < (size_t)nxy; t++) if (irad[t] < iradNR) bin_start[(size_t)irad[t] + 1]++;
      for (int r = 0; r < iradNR; r++) bin_start[(size_t)r + 1] += bin_start[r];
      bin_cells.resize((size_t)bin_start[iradNR]);
      bin_xy.resize((size_t)bin_start[iradNR]);
      std::vector<int> fill(bin_start.begin(), bin_start.end() - 1);
      for (int j = 0; j < d.my; j++)
        for (int i = 0; i < d.mx; i++) {
          const int r = irad[(size_t)j * d.mx + i];
          if (r < iradNR) { bin_xy[(size_t)fill[r]] = j * d.mx + i; bin_cells[(size_t)fill[r]++] = i * d.my + j; }
        }
    }
    // rweight[r] is a chain of cells(r)*mz constant increments: data-independent, closed form per exponent
    std::vector<float> rweight_host((size_t)iradNR);
    for (int r = 0; r < iradNR; r++)
      rweight_host[(size_t)r] = seqsum::inc_chain_f32(0.0f, rw_inc, (long long)(bin_start[(size_t)r + 1] - bin_start[r]) * d.mz);
    // one upload of the table span (through pinned staging when there is a cache)
    std::vector<char> pageable;
    char *stg = stage_pinned(cache, table_span);
    if (!stg) { pageable.resize(table_span); stg = pageable.data(); }
    const char *t0 = d_c2.as<char>();
    memcpy(stg + (d_c2.as<char>() - t0), c2tab.data(), (size_t)nxy * 4);
    memcpy(stg + (d_flags.as<char>() - t0), flags.data(), (size_t)nxy);
    memcpy(stg + (d_bs.as<char>() - t0), bin_start.data(), bin_start.size() * 4);
    if (!bin_cells.empty()) memcpy(stg + (d_bc.as<char>() - t0), bin_cells.data(), bin_cells.size() * 4);
    if (!bin_xy.empty()) memcpy(stg + (d_bxy.as<char>() - t0), bin_xy.data(), bin_xy.size() * 4);
    AN_CUDA(cudaMemcpy(d_c2.p, stg, table_span, cudaMemcpyHostToDevice));
    if (cache) {
      cache->key = key; cache->tables_valid = true;
      cache->bin_start.swap(bin_start); cache->rweight.swap(rweight_host);
    } else {
      bin_start_local.swap(bin_start); rweight_local.swap(rweight_host);
    }
  }
  const std::vector<int> &bin_start = cache ? cache->bin_start : bin_start_local;
  const std::vector<float> &rweight_host = cache ? cache->rweight : rweight_local;
  const long long ncyl = iradNR > 0 ? bin_start[(size_t)iradNR] : 0;

  if (!on_device) {
    const int rcu = upload_grid(cache, opts->device, d_in.p, grid_zfast, ocells * 4);
    if (rcu) return rcu;
  }
  int64_t launches = 0;
  const float *g_in = on_device ? grid_zfast : d_in.as<float>();
  const float *g = g_in;
  // d_rdf, d_hrdf, d_scal and the two flag words in front of the segment maps: one memset
  AN_CUDA(cudaMemset(d_rdf.p, 0, (size_t)(d_fbuf.as<char>() + 16 - d_rdf.as<char>())));
  if (cropped) {
    an_crop_kernel<<<std::min(blocks_for(ncells, 256), 148 * 16), 256>>>(g_in, od, d_g.as<float>(), d, base[0], base[1], base[2]);
    launches++;
    g = d_g.as<float>();
  }

  // ---- kernels: independent chains on independent streams --------------------------------------
  Streams streams;
  AN_CUDA(streams.get(arena.cache));
  cudaStream_t (&st)[4] = streams.st;
  AN_CUDA(cudaDeviceSynchronize());                           // uploads / crop (legacy stream) done
  const auto t_up = now();
  {
    dim3 block(32, 8), grid((d.mz + 31) / 32, (d.mx + 31) / 32, d.my);
    an_transpose_kernel<<<grid, block, 0, st[0]>>>(g, d_gx.as<float>(), d);
    launches++;
  }
  cudaEvent_t gx_ready;                                       // the x-fastest copy exists: the other streams may read it
  AN_CUDA(cudaEventCreateWithFlags(&gx_ready, cudaEventDisableTiming));
  AN_CUDA(cudaEventRecord(gx_ready, st[0]));
  const float d_xy = fmz * U, d_xz = fmy * U, d_yz = fmx * U;               // a_ri3Dc.c:573-578 divisors
  const double h_inc = 1.0 / (double)(fmz * Uunity);                       // :580
  an_xy_kernel<<<blocks_for(nxy, 128), 128, 0, st[0]>>>(d_gx.as<float>(), d, d_xy, min_occ, h_inc, d_xyp.as<float>(), d_hxyp.as<float>());
  double *d_sumP = d_scal.as<double>();
  unsigned long long *d_holes = reinterpret_cast<unsigned long long *>(d_scal.as<char>() + 8);
  float *d_avg = reinterpret_cast<float *>(d_scal.as<char>() + 16);
  int *d_bad = reinterpret_cast<int *>(d_scal.as<char>() + 24);
  // whole-grid chains: a cooperative grid when the grid is large enough to feed one, else two blocks
  int *d_bad_grid = reinterpret_cast<int *>(d_fbuf.as<char>());       // two flag words in front of the segment maps
  bool coop = false;
  if (ncells >= (size_t)SCAN_THREADS * SCAN_T * 16) {
    int dev = 0, sms = 0, can = 0, per_sm = 0;
    AN_CUDA(cudaGetDevice(&dev));
    AN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AN_CUDA(cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, dev));
    AN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, an_chain_grid_kernel, SCAN_THREADS, 0));
    if (can && per_sm > 0) {
      const size_t segs = (ncells + (size_t)SCAN_THREADS * SCAN_T - 1) / ((size_t)SCAN_THREADS * SCAN_T);
      const int nblk = (int)std::min<size_t>((size_t)sms * per_sm, std::max<size_t>(1, segs / 4));
      const float *a_gx = d_gx.as<float>();
      size_t a_n = ncells; int a_nxy = nxy;
      const unsigned char *a_flags = d_flags.as<unsigned char>();
      float a_min = min_occ_raw;
      seqsum::Fn *a_fbuf = reinterpret_cast<seqsum::Fn *>(d_fbuf.as<char>() + 16);
      size_t a_maxseg = maxseg;
      void *args[] = {&a_gx, &a_n, &a_nxy, &a_flags, &a_min, &a_fbuf, &a_maxseg, &d_sumP, &d_holes, &d_avg, &d_bad_grid};
      AN_CUDA(cudaLaunchCooperativeKernel((const void *)an_chain_grid_kernel, dim3(nblk), dim3(SCAN_THREADS), args, 0, st[0]));
      coop = true;
    }
  }
  if (!coop) {
    const int hole_blocks = (int)std::min<size_t>(148, (ncells + SCAN_THREADS * 16 - 1) / (SCAN_THREADS * 16));
    an_chain_scan_kernel<<<2 + hole_blocks, SCAN_THREADS, 0, st[0]>>>(d_gx.as<float>(), ncells, nxy, d_flags.as<unsigned char>(),
                                                                       min_occ_raw, d_sumP, d_holes, d_avg, d_bad);
  }
  an_xz_kernel<<<blocks_for((size_t)d.mx * d.mz, 128), 128, 0, st[1]>>>(g, d, d_xz, d_xzp.as<float>());
  an_yz_kernel<<<blocks_for((size_t)d.my * d.mz, 128), 128, 0, st[1]>>>(g, d, d_yz, d_yzp.as<float>());
  SliceParams sp;
  sp.min_occ = min_occ; sp.d_zdf = fmx * fmy * U; sp.DeltaR = DeltaR; sp.U = U; sp.a_z = tg.a[2]; sp.delta_z = tg.delta[2];
  sp.dV = dV;
  AN_CUDA(cudaStreamWaitEvent(st[2], gx_ready, 0));
  // (limiting how many blocks of the slice / rdf kernels an SM takes, so that they and the cooperative whole-grid kernel
  // are resident together, was measured: 4.84 ms of kernels at C3 size without, 4.83-5.42 with -- the three are work-bound)
  an_slice_scan_kernel<<<dim3(d.mz, 3), SLICE_THREADS, 0, st[2]>>>(d_gx.as<float>(), d, d_c2.as<float>(), d_flags.as<unsigned char>(), min_occ,
                                                                   d_sraw.as<float>(), d_nocc.as<unsigned>(), d_bad);
  an_slice_finish_kernel<<<blocks_for(d.mz, 128), 128, 0, st[2]>>>(d_sraw.as<float>(), d.mz, sp, d_zdf.as<float>(), d_lzdf.as<float>(),
                                                                   d_prof.as<float>());
  launches += 6;
  if (iradNR > 0) {
    an_rz_kernel<<<blocks_for(nrz, 128), 128, 0, st[3]>>>(g, d, d_bs.as<int>(), d_bc.as<int>(), iradNR, min_occ, d_rzp.as<float>(), d_hrzp.as<float>());
    AN_CUDA(cudaStreamWaitEvent(st[3], gx_ready, 0));
    an_rdf_scan_kernel<<<iradNR, RDF_THREADS, 0, st[3]>>>(d_gx.as<float>(), d, d_bs.as<int>(), d_bxy.as<int>(), min_occ, d_rdf.as<float>(),
                                                          d_hrdf.as<float>(), d_bad);
    launches += 2;
  }
  cudaError_t kerr = cudaGetLastError();
  for (auto &s : st) { cudaError_t e = cudaStreamSynchronize(s); if (kerr == cudaSuccess) kerr = e; }
  cudaEventDestroy(gx_ready);
  if (kerr != cudaSuccess) return gcb::fail(GCB_ERR_CUDA, "analysis kernels failed: %s", cudaGetErrorString(kerr));
  bool serial = false;
  {
    int bad = 0, bad_grid[2] = {0, 0};
    cudaError_t e = cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(bad_grid, d_bad_grid, sizeof(bad_grid), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && (bad || bad_grid[0] || bad_grid[1])) {  // negative or non-finite cells: the scans' precondition fails, chain serially
      serial = true;
      e = cudaMemset(d_scal.p, 0, 32);
      an_chain_kernel<<<2, CHAIN_THREADS, 0, st[0]>>>(d_gx.as<float>(), ncells, nxy, d_flags.as<unsigned char>(), min_occ_raw, d_sumP, d_holes, d_avg);
      an_slice_kernel<<<blocks_for(d.mz, 32), 32, 0, st[2]>>>(g, d, d_c2.as<float>(), d_flags.as<unsigned char>(), sp, d_zdf.as<float>(),
                                                               d_lzdf.as<float>(), d_prof.as<float>(), d_nocc.as<unsigned>());
      launches += 2;
      if (iradNR > 0) {
        an_rdf_kernel<<<blocks_for(iradNR, 32), 32, 0, st[3]>>>(g, d, d_bs.as<int>(), d_bc.as<int>(), iradNR, min_occ, rw_inc, d_rdf.as<float>(),
                                                                 d_hrdf.as<float>(), d_rw.as<float>());
        launches++;
      }
      if (e == cudaSuccess) e = cudaGetLastError();
      for (auto &s : st) { cudaError_t e2 = cudaStreamSynchronize(s); if (e == cudaSuccess) e = e2; }
    }
    if (e != cudaSuccess) return gcb::fail(GCB_ERR_CUDA, "analysis kernels failed: %s", cudaGetErrorString(e));
  }
  if (!serial && iradNR > 0)
    AN_CUDA(cudaMemcpy(d_rw.p, rweight_host.data(), (size_t)iradNR * 4, cudaMemcpyHostToDevice));

  const auto t_kern = now();
  // ---- results back; the O(iradNR*mz) normalisations (a_ri3Dc.c:717-762) on the host ----------
  int rc = GCB_OK;
  out->grid = cropped ? host_copy<float>(d_g, ncells, &rc) : nullptr;    // uncropped: the caller's grid IS the analysed grid
  // everything else was laid out contiguously in the arena (d_xyp ... d_scal): one copy, then cut up on the host
  const char *span0 = d_xyp.as<char>();
  const size_t span = (size_t)(d_scal.as<char>() + 32 - span0);
  std::vector<char> stage_pageable;
  char *stage = stage_pinned(cache, span);
  if (!stage) { stage_pageable.resize(span); stage = stage_pageable.data(); }
  if (cudaMemcpy(stage, span0, span, cudaMemcpyDeviceToHost) != cudaSuccess)
    rc = gcb::fail(GCB_ERR_CUDA, "copying the analysis results back failed: %s", cudaGetErrorString(cudaGetLastError()));
  auto cut = [&](const DevBuf &b, size_t n) -> float * {
    float *h = (float *)malloc(std::max<size_t>(n, 1) * sizeof(float));
    if (!h) { rc = gcb::fail(GCB_ERR_NOMEM, "out of memory"); return nullptr; }
    if (n) memcpy(h, stage + (b.as<char>() - span0), n * sizeof(float));
    return h;
  };
  out->xyp = cut(d_xyp, nxy); out->hxyp = cut(d_hxyp, nxy);
  out->xzp = cut(d_xzp, (size_t)d.mx * d.mz); out->yzp = cut(d_yzp, (size_t)d.my * d.mz);
  out->zdf = cut(d_zdf, (size_t)d.mz * 2); out->lzdf = cut(d_lzdf, (size_t)d.mz * 2);
  out->profile = cut(d_prof, (size_t)d.mz * 2);
  out->rweight = cut(d_rw, iradNR);
  out->rdf = cut(d_rdf, (size_t)iradNR * 2); out->hrdf = cut(d_hrdf, (size_t)iradNR * 2);
  out->rzp = cut(d_rzp, nrz); out->hrzp = cut(d_hrzp, nrz);
  std::vector<unsigned> nocc((size_t)d.mz);
  char scal[32];
  memcpy(nocc.data(), stage + (d_nocc.as<char>() - span0), (size_t)d.mz * 4);
  memcpy(scal, stage + (d_scal.as<char>() - span0), 32);
  if (rc) { gcb_analysis_free(out); return rc; }

  const size_t mz = (size_t)d.mz;
  for (int r = 0; r < iradNR; r++) {                          // a_ri3Dc.c:717-726
    const float rw = out->rweight[r];
    out->rdf[2 * r] = out->hrdf[2 * r] = (float)r * DeltaR;
    out->rdf[2 * r + 1] /= rw * fmz * U;
    out->hrdf[2 * r + 1] /= rw * fmz * Uunity;
    for (size_t k = 0; k < mz; k++) {
      out->rzp[(size_t)r * mz + k] /= rw * U;
      out->hrzp[(size_t)r * mz + k] /= rw * Uunity;
    }
  }
  if (opts->rad_ba_step > 0) {                                // a_ri3Dc.c:731-757, with its `irad+1` indexing
    if (opts->rad_ba_nr > 0 && opts->rad_ba_nr + opts->rad_ba_step > iradNR + 1) {
      gcb_analysis_free(out);
      return gcb::fail(GCB_ERR_RANGE, "-radnr %d / -radstep %d reach beyond the %d radial bins", opts->rad_ba_nr, opts->rad_ba_step, iradNR);
    }
    const float stp = (float)opts->rad_ba_step;
    for (int r = 0; r < opts->rad_ba_nr; r += opts->rad_ba_step) {
      for (int i = 1; i < opts->rad_ba_step; i++) {
        out->rdf[2 * r + 1] += out->rdf[2 * (r + i) + 1];
        out->hrdf[2 * r + 1] += out->hrdf[2 * (r + i) + 1];
        for (size_t k = 0; k < mz; k++) {
          out->rzp[(size_t)r * mz + k] += out->rzp[(size_t)(r + 1) * mz + k];
          out->hrzp[(size_t)r * mz + k] += out->hrzp[(size_t)(r + 1) * mz + k];
        }
      }
      out->rdf[2 * r + 1] /= stp; out->hrdf[2 * r + 1] /= stp;
      for (size_t k = 0; k < mz; k++) { out->rzp[(size_t)r * mz + k] /= stp; out->hrzp[(size_t)r * mz + k] /= stp; }
      for (int i = 1; i < opts->rad_ba_step; i++) {
        out->rdf[2 * (r + 1) + 1] = out->rdf[2 * r + 1];
        out->hrdf[2 * (r + 1) + 1] = out->hrdf[2 * r + 1];
        for (size_t k = 0; k < mz; k++) {
          out->rzp[(size_t)(r + 1) * mz + k] = out->rzp[(size_t)r * mz + k];
          out->hrzp[(size_t)(r + 1) * mz + k] = out->hrzp[(size_t)r * mz + k];
        }
      }
    }
  }

  double sumP; unsigned long long nholes; float average;
  memcpy(&sumP, scal, 8); memcpy(&nholes, scal + 8, 8); memcpy(&average, scal + 16, 4);
  unsigned long long nocc_total = 0;
  for (unsigned v : nocc) nocc_total += v;
  // `volume++` on a float (a_ri3Dc.c:594) stops counting at 2^24
  float volume = nocc_total > (1ull << 24) ? 16777216.0f : (float)nocc_total;
  volume = (float)((double)volume * dV);                      // :759
  const int ivol = (int)(ncyl * d.mz);                        // `ivol++` per cell with irad < iradNR (:634)
  out->tg = tg;
  out->geometry = geometry;
  out->DeltaR = DeltaR; out->dV = dV; out->iradNR = iradNR; out->unit_value = U;
  out->sumP = sumP * dV;                                      // :510-511
  out->sumH = (double)nholes * dV;
  out->height = tg.b[2] - tg.a[2];                            // :760
  out->volume = volume;
  out->reff = (float)sqrt((double)volume / (kPi * (double)out->height));   // :761
  out->average = average / ((float)ivol * U);                 // :762
  out->ivol = ivol;
  out->kernel_launches = launches;
  if (trace)
    fprintf(stderr, "[gcb analyse] %dx%dx%d: tables+alloc+H2D %.3f ms, kernels %.3f ms%s, D2H+host finish %.3f ms\n", d.mx, d.my, d.mz,
            ms(t_start, t_up), ms(t_up, t_kern), serial ? " (serial chains)" : "", ms(t_kern, now()));
  return GCB_OK;
}

extern "C" {

int gcb_analyse(const gcb_tgrid *tg, const float *grid_zfast, const gcb_analysis_opts *opts, gcb_analysis *out)
{
  return analyse_impl(tg, grid_zfast, false, opts, out);
}

int gcb_analyse_device(const gcb_tgrid *tg, const float *grid_zfast_dev, const gcb_analysis_opts *opts, gcb_analysis *out)
{
  return analyse_impl(tg, grid_zfast_dev, true, opts, out);
}

// gives back what gcb_analyse keeps per device between calls (device arena, streams, pinned staging, tables)
void gcb_analyse_release_cache(int device)
{
  if (device < 0 || device >= 64) return;
  DeviceCache &dc = g_cache[device];
  std::lock_guard<std::mutex> lk(dc.m);
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return; }
  cudaDeviceSynchronize();
  if (dc.base) cudaFree(dc.base);
  dc.base = nullptr; dc.cap = 0; dc.tables_valid = false;
  for (auto &s : dc.st) { if (s) cudaStreamDestroy(s); s = nullptr; }
  if (dc.h_stage) cudaFreeHost(dc.h_stage);
  dc.h_stage = nullptr; dc.h_stage_cap = 0;
  for (int i = 0; i < DeviceCache::NCHUNK; i++) {
    if (dc.h_chunk[i]) cudaFreeHost(dc.h_chunk[i]);
    if (dc.chunk_done[i]) cudaEventDestroy(dc.chunk_done[i]);
    if (dc.chunk_st[i]) cudaStreamDestroy(dc.chunk_st[i]);
    dc.h_chunk[i] = nullptr; dc.chunk_done[i] = nullptr; dc.chunk_st[i] = nullptr;
  }
  cudaGetLastError();
}

void gcb_analysis_free(gcb_analysis *a)
{
  if (!a) return;
  free(a->grid); free(a->xyp); free(a->xzp); free(a->yzp); free(a->hxyp); free(a->zdf); free(a->lzdf);
  free(a->profile); free(a->rweight); free(a->rdf); free(a->hrdf); free(a->rzp); free(a->hrzp);
  a->grid = a->xyp = a->xzp = a->yzp = a->hxyp = a->zdf = a->lzdf = a->profile = a->rweight = nullptr;
  a->rdf = a->hrdf = a->rzp = a->hrzp = nullptr;
}

int gcb_gridcalc(int ngrids, const float *const *grids_zfast, const float *tweights, const gcb_tgrid *tg,
                 float *avg_zfast, float *tweight_out, int device)
{
  if (ngrids <= 0 || !grids_zfast || !tweights || !tg || !avg_zfast)
    return gcb::fail(GCB_ERR_ARG, "gcb_gridcalc: bad argument");
  if (gcb_device_count() <= 0) return gcb::fail(GCB_ERR_CUDA, "no CUDA device: libgridcount_b200 has no CPU fallback");
  AN_CUDA(cudaSetDevice(device));
  const size_t n = gcb::cells(*tg);
  double sumT = 0;
  for (int f = 0; f < ngrids; f++) sumT += (double)tweights[f];               // a_gridcalc.c:143
  DevBuf d_avg, d_g;
  AN_CUDA(cudaMalloc(&d_avg.p, n * 4));
  AN_CUDA(cudaMalloc(&d_g.p, n * 4));
  AN_CUDA(cudaMemset(d_avg.p, 0, n * 4));
  for (int f = 0; f < ngrids; f++) {
    const double w = (double)tweights[f] / sumT;                               // :182
    AN_CUDA(cudaMemcpy(d_g.p, grids_zfast[f], n * 4, cudaMemcpyHostToDevice));
    gridcalc_axpy_kernel<<<std::min(blocks_for(n, 256), 148 * 16), 256>>>(d_avg.as<float>(), d_g.as<float>(), w, n);
    AN_CUDA(cudaGetLastError());
  }
  AN_CUDA(cudaMemcpy(avg_zfast, d_avg.p, n * 4, cudaMemcpyDeviceToHost));
  if (tweight_out) *tweight_out = (float)sumT;                                 // :170
  return GCB_OK;
}

}  // extern "C"
