// Memory-system self-tests: the measured denominators of K1's roofline (SURVEY 8d asks for the L2-atomic peak
// next to the HBM number; neither is in MEASURED_PEAKS.json).  K1 sends the L2 a coalesced evict-first stream of
// coordinate bytes plus one scattered RED per binned atom; these kernels send it the same traffic with no
// conversion work and no shared memory around it, so what they reach is what the memory system sustains for
// that mix, independently of how K1 is written.  bench.py runs them on the box it measures on.
#include "gcb_internal.h"
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// random-address non-returning adds over a window of `ncells` uint32 cells
__global__ void __launch_bounds__(256) st_red_random(uint32_t *__restrict__ grid, uint32_t ncells, int iters)
{
  uint32_t s = mix32(blockIdx.x * blockDim.x + threadIdx.x + 1);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      s = s * 1664525u + 1013904223u;
      atomicAdd(grid + (uint32_t)(((uint64_t)mix32(s) * ncells) >> 32), 1u);
    }
  }
}

// NUM REDs per DEN streamed float4 (16 B) loads: 4/9 = one RED per 36 B (one atom in three counted),
// 4/3 = one RED per 12 B (every atom counted), 0/9 = the stream alone
template <int NUM, int DEN>
__global__ void __launch_bounds__(256) st_stream_red(const float4 *__restrict__ x, size_t n, uint32_t *__restrict__ grid,
                                                     uint32_t ncells, float *__restrict__ sink)
{
  float acc = 0;
  uint32_t s = mix32(blockIdx.x * blockDim.x + threadIdx.x + 1);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + (DEN - 1) * stride < n; i += DEN * stride) {
    float4 v[DEN];
#pragma unroll
    for (int k = 0; k < DEN; k++) v[k] = __ldcs(x + i + k * stride);
#pragma unroll
    for (int k = 0; k < DEN; k++) acc += v[k].x + v[k].y + v[k].z + v[k].w;
#pragma unroll
    for (int k = 0; k < NUM; k++) {
      s = s * 1664525u + 1013904223u;
      atomicAdd(grid + (uint32_t)(((uint64_t)mix32(s ^ __float_as_uint(acc)) * ncells) >> 32), 1u);
    }
  }
  if (acc == 1234.5f) *sink = acc;
}

template <class F>
int best_of(F launch, int reps, float *best_ms, cudaStream_t st)
{
  cudaEvent_t a, b;
  GCB_CUDA(cudaEventCreate(&a));
  GCB_CUDA(cudaEventCreate(&b));
  float best = 1e30f;
  for (int r = 0; r < reps + 1; r++) {
    cudaEventRecord(a, st);
    launch();
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (r > 0) best = std::min(best, ms);                 // the first run warms up
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  GCB_CUDA(cudaGetLastError());
  *best_ms = best;
  return GCB_OK;
}

}  // namespace

extern "C" int gcb_selftest_l2(int device, size_t window_bytes, int mix, size_t stream_bytes, double *red_gops, double *read_gbs)
{
  if (window_bytes < 4096 || mix < 0 || mix > 3) return gcb::fail(GCB_ERR_ARG, "gcb_selftest_l2: bad argument");
  if (gcb_device_count() <= 0) return gcb::fail(GCB_ERR_CUDA, "no CUDA device");
  GCB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  GCB_CUDA(cudaGetDeviceProperties(&prop, device));
  const int sms = prop.multiProcessorCount;
  uint32_t *g = nullptr;
  float4 *x = nullptr;
  float *sink = nullptr;
  GCB_CUDA(cudaMalloc(&g, window_bytes));
  GCB_CUDA(cudaMemset(g, 0, window_bytes));
  const uint32_t ncells = (uint32_t)std::min<size_t>(window_bytes / 4, 0xffffffffu);
  // windows beyond 32 MB get what K1's packed path gives its counters: a persisting-L2 set-aside and an
  // access-policy window on the launching stream (without it the stream evicts the window: tools/l2_window.cu)
  cudaStream_t st = nullptr;
  GCB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  bool windowed = false;
  if (window_bytes > ((size_t)32 << 20) && !getenv("GCB_NO_L2_PERSIST")) {
    const size_t want = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, window_bytes + ((size_t)2 << 20));
    if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
      cudaStreamAttrValue av = {};
      av.accessPolicyWindow.base_ptr = g;
      av.accessPolicyWindow.num_bytes = std::min<size_t>(window_bytes, (size_t)prop.accessPolicyMaxWindowSize);
      av.accessPolicyWindow.hitRatio = std::min(1.0f, (float)((double)want / (double)window_bytes));
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      windowed = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
    }
    cudaGetLastError();
  }
  float ms = 0;
  int rc = GCB_OK;
  double reds = 0, bytes = 0;
  if (mix == 0) {
    const int blocks = sms * 8, iters = 128;
    rc = best_of([&] { st_red_random<<<blocks, 256, 0, st>>>(g, ncells, iters); }, 3, &ms, st);
    reds = (double)blocks * 256 * iters * 8;
  } else {
    if (stream_bytes < ((size_t)256 << 20)) stream_bytes = (size_t)2 << 30;     // well beyond the L2
    if (cudaMalloc(&x, stream_bytes) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) {
      cudaGetLastError();
      cudaFree(g); cudaFree(x); cudaStreamDestroy(st);
      return gcb::fail(GCB_ERR_NOMEM, "gcb_selftest_l2: cannot allocate the %zu-byte stream", stream_bytes);
    }
    cudaMemset(x, 0, stream_bytes);
    const size_t n = stream_bytes / 16;
    const int blocks = sms * 16;
    if (mix == 1) { rc = best_of([&] { st_stream_red<4, 9><<<blocks, 256, 0, st>>>(x, n, g, ncells, sink); }, 3, &ms, st); reds = (double)n * 4 / 9; }
    if (mix == 2) { rc = best_of([&] { st_stream_red<4, 3><<<blocks, 256, 0, st>>>(x, n, g, ncells, sink); }, 3, &ms, st); reds = (double)n * 4 / 3; }
    if (mix == 3) { rc = best_of([&] { st_stream_red<0, 9><<<blocks, 256, 0, st>>>(x, n, g, ncells, sink); }, 3, &ms, st); reds = 0; }
    bytes = (double)stream_bytes;
  }
  if (windowed) {
    cudaStreamAttrValue av = {};
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
    cudaCtxResetPersistingL2Cache();
  }
  cudaStreamDestroy(st);
  cudaFree(g); cudaFree(x); cudaFree(sink);
  if (rc) return rc;
  if (red_gops) *red_gops = reds / (ms * 1e-3) * 1e-9;
  if (read_gbs) *read_gbs = bytes / (ms * 1e-3) * 1e-9;
  return GCB_OK;
}
