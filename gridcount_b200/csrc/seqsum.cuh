// Exact parallel evaluation of a SEQUENTIAL floating-point accumulation  s = fl(s + x_i), x_i >= 0.
//
// a_ri3Dc sums the grid in one fixed order into f32 (and one f64) accumulators (a_ri3Dc.c:499-508,
// 567-703); its results carry that order's rounding error (an f32 `average` over 64 M cells even
// saturates), so matching it to 1e-6 means reproducing the chain bit for bit.  A chain looks serial,
// but between two changes of the accumulator's exponent it is an INTEGER recurrence:
//
//   s = M * u  with  u = 2^(E - bias) the accumulator's ulp and  2^(p-1) <= M < 2^p  (p = 24 or 53);
//   x = k * u + r, 0 <= r < u.   Round-to-nearest-even gives   M' = M + k + c,
//       c = [r > u/2]  or, on a tie r == u/2,  c = (M + k) & 1.
//
// So every term is a map  M -> M + (M odd ? a1 : a0)  (a0 == a1 unless the term is a tie), and such
// maps are closed under composition:  (g o f)[q] = f[q] + g[(q + f[q]) & 1].  Composition is
// associative, hence a chunk of terms is reduced with an ordinary parallel scan.  Terms are
// non-negative, so M only grows and the recurrence stops being valid exactly once per exponent: at the
// first term that takes M to 2^p or beyond.  That single term is added with a real floating-point
// addition, the exponent is re-read, and the scan resumes behind it.  Zero and subnormal accumulators
// are the E = 1 case with 0 <= M < 2^p, so the chain starts from s = +0 without a special case.
//
// Inputs must be finite and >= 0 (-0 counts as 0); callers test that and fall back to the serial
// chain otherwise.  Terms are f32 in both instantiations (the f64 chain adds (double)x).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SEQSUM_HD __host__ __device__ __forceinline__
#else
#define SEQSUM_HD inline
#endif

namespace seqsum {

// M -> M + (M & 1 ? a1 : a0); {0,0} is the identity.  The f32 chains work on 32-bit maps (every entry is capped at
// 2^25), the f64 chain on 64-bit ones; buffers in shared and global memory hold the 64-bit form (Fn).
template <class U> struct alignas(2 * sizeof(U)) FnT { U a0, a1; };
using Fn = FnT<unsigned long long>;

SEQSUM_HD int high_bit(uint32_t v)                      // v != 0
{
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)v);
#else
  return 31 - __builtin_clz(v);
#endif
}

template <class R> struct Tr;
template <> struct Tr<float> {
  using U = uint32_t;                                   // map entries: <= 2^25
  static constexpr int MBITS = 24;                      // significand bits incl. the implicit one
  static constexpr int EOFF = 0;                        // a term's shift against the ulp: sh = E - ex - EOFF
  SEQSUM_HD static void unpack(float s, int &E, unsigned long long &M)
  {
    uint32_t b;
#if defined(__CUDA_ARCH__)
    b = __float_as_uint(s);
#else
    __builtin_memcpy(&b, &s, 4);
#endif
    E = (int)((b >> 23) & 0xffu);
    M = b & 0x7fffffu;
    if (E) M |= 0x800000u; else E = 1;
  }
  SEQSUM_HD static float pack(int E, unsigned long long M)
  {
    const uint32_t b = M < 0x800000ull ? (uint32_t)M : (((uint32_t)E << 23) | ((uint32_t)M & 0x7fffffu));
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float s; __builtin_memcpy(&s, &b, 4); return s;
#endif
  }
  SEQSUM_HD static bool finite(float s) { int E; unsigned long long M; unpack(s, E, M); return E != 255; }
  SEQSUM_HD static float add(float s, float x)
  {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(s, x);
#else
    volatile float r = s + x; return r;
#endif
  }
};
template <> struct Tr<double> {
  using U = unsigned long long;                         // map entries: <= 2^54
  static constexpr int MBITS = 53;
  static constexpr int EOFF = 925;                      // 1075 - 150
  SEQSUM_HD static void unpack(double s, int &E, unsigned long long &M)
  {
    unsigned long long b;
#if defined(__CUDA_ARCH__)
    b = (unsigned long long)__double_as_longlong(s);
#else
    __builtin_memcpy(&b, &s, 8);
#endif
    E = (int)((b >> 52) & 0x7ffu);
    M = b & 0xfffffffffffffull;
    if (E) M |= 1ull << 52; else E = 1;
  }
  SEQSUM_HD static double pack(int E, unsigned long long M)
  {
    const unsigned long long b = M < (1ull << 52) ? M : (((unsigned long long)E << 52) | (M & 0xfffffffffffffull));
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double s; __builtin_memcpy(&s, &b, 8); return s;
#endif
  }
  SEQSUM_HD static bool finite(double s) { int E; unsigned long long M; unpack(s, E, M); return E != 2047; }
  SEQSUM_HD static double add(double s, float x)
  {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(s, (double)x);
#else
    volatile double r = s + (double)x; return r;
#endif
  }
};

// a finite, non-negative f32 term?  (-0 is accepted: s + -0 == s for every s >= +0)
SEQSUM_HD bool term_ok(uint32_t xb) { return ((xb >> 23) & 0xffu) != 255u && (!(xb >> 31) || (xb << 1) == 0u); }

template <class R> using Map = FnT<typename Tr<R>::U>;

template <class U> SEQSUM_HD Fn widen(const FnT<U> &f) { Fn o; o.a0 = f.a0; o.a1 = f.a1; return o; }
template <class R> SEQSUM_HD Map<R> narrow(const Fn &f)          // (entries of an R-chain's map fit Tr<R>::U by construction)
{
  Map<R> o; o.a0 = (typename Tr<R>::U)f.a0; o.a1 = (typename Tr<R>::U)f.a1; return o;
}

// the map of one f32 term (bit pattern xb, sign ignored) while the accumulator's exponent field is E
template <class R>
SEQSUM_HD Map<R> term(uint32_t xb, int E)
{
  using U = typename Tr<R>::U;
  constexpr U CAP = (U)1 << (Tr<R>::MBITS + 1);                        // "reaches the next exponent for sure"
  int ex = (int)((xb >> 23) & 0xffu);
  uint32_t mx = xb & 0x7fffffu;
  if (ex) mx |= 0x800000u; else ex = 1;
  Map<R> f; f.a0 = 0; f.a1 = 0;
  if (mx == 0u) return f;
  const int sh = E - ex - Tr<R>::EOFF;                                 // x / u = mx / 2^sh
  if (sh <= 0) {
    const U k = (high_bit(mx) - sh >= Tr<R>::MBITS + 1) ? CAP : ((U)mx << (-sh));
    f.a0 = k; f.a1 = k;
    return f;
  }
  if (sh >= 25) return f;                                              // x < u/2: the accumulator does not move
  const uint32_t k = mx >> sh, r = mx & ((1u << sh) - 1u), half = 1u << (sh - 1);
  if (r != half) { f.a0 = f.a1 = k + (r > half ? 1u : 0u); return f; }
  f.a0 = k + (k & 1u);                                                 // tie: to even
  f.a1 = k + 1u - (k & 1u);
  return f;
}

// g after f, saturating at CAP (a saturated map only ever says "crossed"; its parity is not used)
template <class R>
SEQSUM_HD Map<R> then(const Map<R> &f, const Map<R> &g)
{
  using U = typename Tr<R>::U;
  constexpr U CAP = (U)1 << (Tr<R>::MBITS + 1);
  Map<R> h;
  h.a0 = f.a0 + ((f.a0 & 1u) ? g.a1 : g.a0);
  h.a1 = f.a1 + ((f.a1 & 1u) ? g.a0 : g.a1);
  if (h.a0 > CAP) h.a0 = CAP;
  if (h.a1 > CAP) h.a1 = CAP;
  return h;
}

template <class U> SEQSUM_HD unsigned long long apply(const FnT<U> &f, unsigned long long M) { return M + ((M & 1ull) ? f.a1 : f.a0); }

// n repetitions of  s = (float)((double)s + c)  with a constant c > 0 (a_ri3Dc.c:637, `rweight`): per exponent
// of s the step is one fixed map, so the chain advances a whole exponent at a time.  Host code.
inline float inc_chain_f32(float s, double c, long long n)
{
  constexpr unsigned long long LIMIT = 1ull << 24;
  while (n > 0) {
    int E; unsigned long long M;
    Tr<float>::unpack(s, E, M);
    const double B = __builtin_ldexp(1.0, E - 127);                    // lower end of s's binade
    if (M < 0x800000ull || !(c < B) || E >= 254) {                     // zero, subnormal, or c not small: plain step
      volatile double t = (double)s + c; volatile float r = (float)t; s = r; n--;
      continue;
    }
    const double u = __builtin_ldexp(1.0, E - 150);
    volatile double bc = B + c;                                        // the f64 rounding of s + c, on this binade's grid
    const double cp = bc - B;
    const double q = cp / u, kf = __builtin_floor(q), r = q - kf;      // all exact
    const unsigned long long k = (unsigned long long)kf;
    Fn f;
    if (r != 0.5) f.a0 = f.a1 = k + (r > 0.5 ? 1ull : 0ull);
    else { f.a0 = k + (k & 1ull); f.a1 = k + 1ull - (k & 1ull); }
    const unsigned long long M1 = apply(f, M);
    if (M1 >= LIMIT) {                                                 // this step changes the exponent
      volatile double t = (double)s + c; volatile float rr = (float)t; s = rr; n--;
      continue;
    }
    const unsigned long long A = f.a0;                                 // after a tie the state is even and stays even
    if (A == 0) return Tr<float>::pack(E, M1);                         // c < ulp/2: the accumulator is stuck
    const unsigned long long jmax = (LIMIT - 1ull - M1) / A;
    const unsigned long long take = (unsigned long long)n < 1ull + jmax ? (unsigned long long)n : 1ull + jmax;
    s = Tr<float>::pack(E, M1 + (take - 1ull) * A);
    n -= (long long)take;
  }
  return s;
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------------
// Cooperative chains.  `load(i)` returns term i (f32, already masked; must return +0 for i >= n).
// Every thread owns T consecutive terms of a chunk, so thread order == chain order.
// ---------------------------------------------------------------------------------------------------
template <class U>
__device__ __forceinline__ FnT<U> shfl_up_fn(const FnT<U> &f, int d)
{
  FnT<U> o;
  o.a0 = __shfl_up_sync(0xffffffffu, f.a0, d);
  o.a1 = __shfl_up_sync(0xffffffffu, f.a1, d);
  return o;
}

template <class R>
__device__ __forceinline__ Map<R> warp_scan(Map<R> f, int lane)        // inclusive
{
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const Map<R> o = shfl_up_fn(f, d);
    if (lane >= d) f = then<R>(o, f);
  }
  return f;
}

// One thread's walk over its T terms from state (E, M): returns the number of terms consumed before the
// exponent changes (T if it does not); M is advanced over the consumed terms.
template <class R, int T>
__device__ __forceinline__ int walk(const float (&x)[T], int E, unsigned long long &M)
{
  constexpr unsigned long long LIMIT = 1ull << Tr<R>::MBITS;
#pragma unroll
  for (int j = 0; j < T; j++) {
    const unsigned long long m2 = apply(term<R>(__float_as_uint(x[j]), E), M);
    if (m2 >= LIMIT) return j;
    M = m2;
  }
  return T;
}

struct BlockScratch {
  Fn warp_tot[32];
  unsigned long long cross_M;       // state just before the crossing term
  float cross_x;
  unsigned long long cross_seg;     // GridChain: the segment in which the exponent changes
  int first_cross;                  // lowest thread id whose terms reach the next exponent
  int cross_j;
  int bad;
};

// A chain over n terms by one thread block of NT threads (NT a multiple of 32, <= 1024).  Returns the
// accumulator in every thread; *bad is set (and the result meaningless) when a term is negative or
// not finite, or the sum overflows.
template <class R, int T, int NT, class Load>
__device__ R block_chain(size_t n, Load load, BlockScratch &sh, bool *bad)
{
  constexpr unsigned long long LIMIT = 1ull << Tr<R>::MBITS;
  constexpr int CH = T * NT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int E; unsigned long long M;
  Tr<R>::unpack((R)0, E, M);
  if (tid == 0) { sh.first_cross = NT; sh.bad = 0; }
  __syncthreads();
  size_t pos = 0;
  float x[T];
  bool have = false;                                                   // x[] holds the chunk at `pos`
  float nx[T];
  while (pos < n) {
    if (!have) {
#pragma unroll
      for (int j = 0; j < T; j++) x[j] = load(pos + (size_t)tid * T + j);
    }
    const bool more = pos + CH < n;
    if (more) {
#pragma unroll
      for (int j = 0; j < T; j++) nx[j] = load(pos + CH + (size_t)tid * T + j);
    }
    bool ok = true;
    Map<R> f; f.a0 = 0; f.a1 = 0;
#pragma unroll
    for (int j = 0; j < T; j++) {
      const uint32_t xb = __float_as_uint(x[j]);
      ok &= term_ok(xb);
      f = then<R>(f, term<R>(xb, E));
    }
    if (!ok) sh.bad = 1;
    const Map<R> incl_w = warp_scan<R>(f, lane);
    if (lane == 31) sh.warp_tot[warp] = widen(incl_w);
    __syncthreads();
    if (warp == 0) {
      Map<R> t; t.a0 = 0; t.a1 = 0;
      if (lane < NT / 32) t = narrow<R>(sh.warp_tot[lane]);
      t = warp_scan<R>(t, lane);
      sh.warp_tot[lane] = widen(t);
    }
    __syncthreads();
    Map<R> excl = shfl_up_fn(incl_w, 1);
    if (lane == 0) { excl.a0 = 0; excl.a1 = 0; }
    if (warp > 0) excl = then<R>(narrow<R>(sh.warp_tot[warp - 1]), excl);
    const Map<R> incl = then<R>(excl, f);
    const Map<R> total = narrow<R>(sh.warp_tot[NT / 32 - 1]);
    const bool crosses = apply(incl, M) >= LIMIT;
    if (crosses) atomicMin(&sh.first_cross, tid);
    __syncthreads();
    if (sh.bad) { *bad = true; return (R)0; }
    const int fc = sh.first_cross;
    if (fc == NT) {                                                    // the common case: same exponent throughout
      M = apply(total, M);
      pos += CH;
      if (more) {
#pragma unroll
        for (int j = 0; j < T; j++) x[j] = nx[j];
      }
      have = more;
      continue;                                                        // first_cross is still NT; no barrier needed
    }
    if (tid == fc) {
      unsigned long long m = apply(excl, M);
      const int j = walk<R, T>(x, E, m);                               // j < T by construction
      sh.cross_M = m; sh.cross_j = j;
      float xv = 0.0f;
#pragma unroll
      for (int q = 0; q < T; q++) if (q == j) xv = x[q];
      sh.cross_x = xv;
    }
    __syncthreads();
    const R s = Tr<R>::add(Tr<R>::pack(E, sh.cross_M), sh.cross_x);   // the one real addition of this exponent
    pos += (size_t)fc * T + sh.cross_j + 1;
    have = false;
    __syncthreads();
    if (tid == 0) sh.first_cross = NT;
    if (!Tr<R>::finite(s)) { *bad = true; return (R)0; }
    Tr<R>::unpack(s, E, M);
    __syncthreads();
  }
  return Tr<R>::pack(E, M);
}

// ---------------------------------------------------------------------------------------------------
// A chain spread over a whole cooperative grid.  The chain is cut into segments of T*NT terms.  Each
// round, (1) the blocks reduce the segments of a window behind the current position to one map each,
// for the accumulator's CURRENT exponent; (2) after a grid-wide barrier every block -- redundantly, so
// no second barrier or broadcast is needed -- scans those maps, advances over the window if the
// exponent holds, or else replays the one segment in which it changes, up to the crossing term.  The
// window doubles while nothing happens.  Rounds ~ exponent changes + log2(segments).
//   `load(i, x)` fills x[0..T) with terms i .. i+T-1 (+0 beyond the end).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fn ldcg_fn(const Fn *p)                     // written by other blocks: not through L1
{
  const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(p));
  Fn f; f.a0 = v.x; f.a1 = v.y;
  return f;
}

// inclusive/exclusive block scan of one map per thread; two barriers inside, and the caller must pass
// another barrier before the next block_scan (sh.warp_tot is reused)
template <class R, int NT>
__device__ __forceinline__ void block_scan(const Map<R> &f, BlockScratch &sh, Map<R> &excl, Map<R> &incl, Map<R> &total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const Map<R> incl_w = warp_scan<R>(f, lane);
  if (lane == 31) sh.warp_tot[warp] = widen(incl_w);
  __syncthreads();
  if (warp == 0) {
    Map<R> t; t.a0 = 0; t.a1 = 0;
    if (lane < NT / 32) t = narrow<R>(sh.warp_tot[lane]);
    t = warp_scan<R>(t, lane);
    sh.warp_tot[lane] = widen(t);
  }
  __syncthreads();
  excl = shfl_up_fn(incl_w, 1);
  if (lane == 0) { excl.a0 = 0; excl.a1 = 0; }
  if (warp > 0) excl = then<R>(narrow<R>(sh.warp_tot[warp - 1]), excl);
  incl = then<R>(excl, f);
  total = narrow<R>(sh.warp_tot[NT / 32 - 1]);
}

template <class R, int T, int NT>
struct GridChain {
  static constexpr int CH = T * NT;
  static constexpr unsigned long long LIMIT = 1ull << Tr<R>::MBITS;
  int E; unsigned long long M;
  size_t pos; unsigned long long W;                                    // position; window in segments
  bool done;

  __device__ void init(size_t n) { Tr<R>::unpack((R)0, E, M); pos = 0; W = gridDim.x; done = n == 0; }
  __device__ R value() const { return Tr<R>::pack(E, M); }
  __device__ unsigned long long window(size_t n) const
  {
    const unsigned long long left = ((unsigned long long)(n - pos) + CH - 1) / CH;
    return left < W ? left : W;
  }

  // Most exponent changes of a sum that starts at zero happen within its first few thousand terms (the k-th one
  // near term 2^k / mean).  Every block therefore walks the first `limit` terms by itself, redundantly, one
  // block-level scan per chunk or per exponent change, before the grid-wide rounds (one barrier each) begin.
  template <class Load>
  __device__ void prologue(size_t n, Load load, BlockScratch &sh, bool *bad, size_t limit)
  {
    const int tid = threadIdx.x;
    const size_t end = n < limit ? n : limit;
    while (pos < end && !done) {
      float x[T];
      load(pos + (size_t)tid * T, x);
      bool ok = true;
      Map<R> f; f.a0 = 0; f.a1 = 0;
#pragma unroll
      for (int j = 0; j < T; j++) {
        if (pos + (size_t)tid * T + j >= end) x[j] = 0.0f;             // the rounds take over at `end`
        const uint32_t xb = __float_as_uint(x[j]);
        ok &= term_ok(xb);
        f = then<R>(f, term<R>(xb, E));
      }
      if (!ok) sh.bad = 1;
      Map<R> excl, incl, total;
      block_scan<R, NT>(f, sh, excl, incl, total);
      if (apply(incl, M) >= LIMIT) atomicMin(&sh.first_cross, tid);
      __syncthreads();
      if (sh.bad) { *bad = true; done = true; return; }
      const int fc = sh.first_cross;
      if (fc == NT) {
        M = apply(total, M);
        pos = end - pos > (size_t)CH ? pos + CH : end;
        continue;
      }
      if (tid == fc) {
        unsigned long long m = apply(excl, M);
        const int j = walk<R, T>(x, E, m);
        float xv = 0.0f;
#pragma unroll
        for (int q = 0; q < T; q++) if (q == j) xv = x[q];
        sh.cross_M = m; sh.cross_j = j; sh.cross_x = xv;
      }
      __syncthreads();
      const R snew = Tr<R>::add(Tr<R>::pack(E, sh.cross_M), sh.cross_x);
      pos += (size_t)fc * T + sh.cross_j + 1;
      __syncthreads();
      if (tid == 0) sh.first_cross = NT;
      __syncthreads();
      if (!Tr<R>::finite(snew)) { *bad = true; done = true; return; }
      Tr<R>::unpack(snew, E, M);
    }
    done = pos >= n;
  }

  template <class Load>
  __device__ void phase1(size_t n, Load load, Fn *fbuf, BlockScratch &sh, int *bad_flag) const
  {
    if (done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long nseg = window(n);
    for (unsigned long long s = blockIdx.x; s < nseg; s += gridDim.x) {
      float x[T];
      load(pos + s * CH + (size_t)tid * T, x);
      bool ok = true;
      Map<R> f; f.a0 = 0; f.a1 = 0;
#pragma unroll
      for (int j = 0; j < T; j++) {
        const uint32_t xb = __float_as_uint(x[j]);
        ok &= term_ok(xb);
        f = then<R>(f, term<R>(xb, E));
      }
      if (!ok) atomicExch(bad_flag, 1);
      const Map<R> w = warp_scan<R>(f, lane);
      if (lane == 31) sh.warp_tot[warp] = widen(w);
      __syncthreads();
      if (warp == 0) {
        Map<R> t; t.a0 = 0; t.a1 = 0;
        if (lane < NT / 32) t = narrow<R>(sh.warp_tot[lane]);
        t = warp_scan<R>(t, lane);
        if (lane == 31) fbuf[s] = widen(t);                            // lanes beyond NT/32 hold the identity
      }
      __syncthreads();
    }
  }

  template <class Load>
  __device__ void phase2(size_t n, Load load, const Fn *fbuf, BlockScratch &sh, bool *bad)
  {
    if (done) return;
    const int tid = threadIdx.x;
    const unsigned long long nseg = window(n);
    const unsigned long long per = (nseg + NT - 1) / NT;
    const unsigned long long s0 = (unsigned long long)tid * per;
    const unsigned long long s1 = s0 + per < nseg ? s0 + per : nseg;
    Map<R> f; f.a0 = 0; f.a1 = 0;
    for (unsigned long long s = s0; s < s1; s++) f = then<R>(f, narrow<R>(ldcg_fn(fbuf + s)));
    Map<R> excl, incl, total;
    block_scan<R, NT>(f, sh, excl, incl, total);
    if (apply(incl, M) >= LIMIT) atomicMin(&sh.first_cross, tid);
    __syncthreads();
    const int fc = sh.first_cross;
    if (fc == NT) {                                                    // the window keeps the exponent
      M = apply(total, M);
      const size_t adv = (size_t)nseg * CH;
      pos = adv < n - pos ? pos + adv : n;
      W = W < (1ull << 40) ? W * 2 : W;
      done = pos >= n;
      return;
    }
    if (tid == fc) {
      unsigned long long m = apply(excl, M), s = s0;
      for (; s < s1; s++) {
        const unsigned long long m2 = apply(ldcg_fn(fbuf + s), m);
        if (m2 >= LIMIT) break;
        m = m2;
      }
      sh.cross_M = m; sh.cross_seg = s;
    }
    __syncthreads();
    const unsigned long long Mc = sh.cross_M;
    const size_t base = pos + (size_t)sh.cross_seg * CH;
    if (tid == 0) sh.first_cross = NT;
    // replay that segment from (E, Mc); the scan must find the crossing term
    float x[T];
    load(base + (size_t)tid * T, x);
    Map<R> g; g.a0 = 0; g.a1 = 0;
#pragma unroll
    for (int j = 0; j < T; j++) g = then<R>(g, term<R>(__float_as_uint(x[j]), E));
    block_scan<R, NT>(g, sh, excl, incl, total);                       // its barriers also publish first_cross = NT
    if (apply(incl, Mc) >= LIMIT) atomicMin(&sh.first_cross, tid);
    __syncthreads();
    const int fc2 = sh.first_cross;
    if (fc2 == NT) { *bad = true; done = true; return; }               // cannot happen; never spin on it
    if (tid == fc2) {
      unsigned long long m = apply(excl, Mc);
      const int j = walk<R, T>(x, E, m);
      float xv = 0.0f;
#pragma unroll
      for (int q = 0; q < T; q++) if (q == j) xv = x[q];
      sh.cross_M = m; sh.cross_j = j; sh.cross_x = xv;
    }
    __syncthreads();
    const R snew = Tr<R>::add(Tr<R>::pack(E, sh.cross_M), sh.cross_x);
    pos = base + (size_t)fc2 * T + sh.cross_j + 1;
    __syncthreads();
    if (tid == 0) sh.first_cross = NT;
    __syncthreads();
    if (!Tr<R>::finite(snew)) { *bad = true; done = true; return; }
    Tr<R>::unpack(snew, E, M);
    done = pos >= n;
  }
};

// The same by one warp (no block barriers): for many medium-sized chains, one warp each.
template <class R, int T, class Load>
__device__ R warp_chain(size_t n, Load load, bool *bad)
{
  constexpr unsigned long long LIMIT = 1ull << Tr<R>::MBITS;
  constexpr int CH = T * 32;
  const int lane = threadIdx.x & 31;
  int E; unsigned long long M;
  Tr<R>::unpack((R)0, E, M);
  size_t pos = 0;
  float x[T];
  while (pos < n) {
#pragma unroll
    for (int j = 0; j < T; j++) x[j] = load(pos + (size_t)lane * T + j);
    bool ok = true;
    Map<R> f; f.a0 = 0; f.a1 = 0;
#pragma unroll
    for (int j = 0; j < T; j++) {
      const uint32_t xb = __float_as_uint(x[j]);
      ok &= term_ok(xb);
      f = then<R>(f, term<R>(xb, E));
    }
    if (!__all_sync(0xffffffffu, ok)) { *bad = true; return (R)0; }
    const Map<R> incl = warp_scan<R>(f, lane);
    Map<R> excl = shfl_up_fn(incl, 1);
    if (lane == 0) { excl.a0 = 0; excl.a1 = 0; }
    const unsigned cross = __ballot_sync(0xffffffffu, apply(incl, M) >= LIMIT);
    if (!cross) {
      Map<R> total;
      total.a0 = __shfl_sync(0xffffffffu, incl.a0, 31);
      total.a1 = __shfl_sync(0xffffffffu, incl.a1, 31);
      M = apply(total, M);
      pos += CH;
      continue;
    }
    const int fc = __ffs((int)cross) - 1;
    unsigned long long m = apply(excl, M);
    int j = 0;
    float xv = 0.0f;
    if (lane == fc) {
      j = walk<R, T>(x, E, m);
#pragma unroll
      for (int q = 0; q < T; q++) if (q == j) xv = x[q];
    }
    m = __shfl_sync(0xffffffffu, m, fc);
    j = __shfl_sync(0xffffffffu, j, fc);
    xv = __shfl_sync(0xffffffffu, xv, fc);
    const R s = Tr<R>::add(Tr<R>::pack(E, m), xv);
    if (!Tr<R>::finite(s)) { *bad = true; return (R)0; }
    Tr<R>::unpack(s, E, M);
    pos += (size_t)fc * T + j + 1;
  }
  return Tr<R>::pack(E, M);
}
#endif  // __CUDACC__

}  // namespace seqsum
