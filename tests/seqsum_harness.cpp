// CPU harness for gridcount_b200/csrc/seqsum.cuh: runs the term / then / apply primitives the GPU chains are
// built from through the same chunked procedure (compose a chunk, test for an exponent change, redo the
// crossing term with a real addition), so the arithmetic can be checked without a GPU.
// Built by tests/test_seqsum_cpu.py:  g++ -O2 -shared -fPIC -ffp-contract=off
#include "seqsum.cuh"
#include <cstddef>
#include <vector>
using namespace seqsum;

template <class R>
static R chain(const float *x, size_t n, int chunk, long *crossings, int *bad)
{
  constexpr unsigned long long LIMIT = 1ull << Tr<R>::MBITS;
  int E; unsigned long long M;
  Tr<R>::unpack((R)0, E, M);
  size_t pos = 0;
  *crossings = 0; *bad = 0;
  std::vector<Map<R>> pre((size_t)chunk);
  while (pos < n) {
    const size_t m = n - pos < (size_t)chunk ? n - pos : (size_t)chunk;
    Map<R> acc; acc.a0 = acc.a1 = 0;
    for (size_t i = 0; i < m; i++) {
      uint32_t xb; __builtin_memcpy(&xb, x + pos + i, 4);
      if (!term_ok(xb)) { *bad = 1; return (R)0; }
      acc = then<R>(acc, term<R>(xb, E));
      pre[i] = acc;
    }
    if (apply(acc, M) < LIMIT) { M = apply(acc, M); pos += m; continue; }
    size_t c = 0;
    while (apply(pre[c], M) < LIMIT) c++;
    const unsigned long long mprev = c ? apply(pre[c - 1], M) : M;
    const R s = Tr<R>::add(Tr<R>::pack(E, mprev), x[pos + c]);
    if (!Tr<R>::finite(s)) { *bad = 1; return (R)0; }
    Tr<R>::unpack(s, E, M);
    pos += c + 1;
    ++*crossings;
  }
  return Tr<R>::pack(E, M);
}

extern "C" float ss_chain_f32(const float *x, size_t n, int chunk, long *crossings, int *bad) { return chain<float>(x, n, chunk, crossings, bad); }
extern "C" double ss_chain_f64(const float *x, size_t n, int chunk, long *crossings, int *bad) { return chain<double>(x, n, chunk, crossings, bad); }
extern "C" float ss_inc_chain_f32(float s, double c, long long n) { return inc_chain_f32(s, c, n); }
