// The passes around K1 and K2: closed-form f32 accumulation (g_ri3Dc.c:149), verification / application of the packed
// counters, folding of the private grids, saturation guard, weight-run folding, density (g_ri3Dc.c:451-460), the
// file-order transposes (xdr_grid.c:113-126) and the synthetic frame generator.  Included by binning.cu only.
#pragma once
#include "k1_kernels.cuh"

namespace k1 {

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

}  // namespace k1
