// How to keep a RED window L2-resident next to a streaming read (the packed-counter path of K1 on grids beyond L2):
// one RED per 36 streamed bytes into windows of 16..96 MB, with
//   red hint   : none | createpolicy evict_last on the RED
//   stream hint: ld.global.cs | createpolicy evict_first | createpolicy evict_first + L1::no_allocate
//   carve-out  : persisting-L2 set-aside + access-policy window on the RED array, or none
// Prints one JSON line per combination.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_window tools/l2_window.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

template <int SH>
__device__ __forceinline__ float4 ld_stream(const float4 *p, unsigned long long pol)
{
  float4 v;
  if (SH == 0) return __ldcs(p);
  if (SH == 1)
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  if (SH == 2)
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}

template <int RH>
__device__ __forceinline__ void red_add(uint32_t *p, uint32_t v, unsigned long long pol)
{
  if (RH == 0) atomicAdd(p, v);
  else asm volatile("red.global.add.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

// 4 REDs per 9 float4 loads = one RED per 36 B
template <int RH, int SH, int NUM>
__global__ void __launch_bounds__(256) k(const float4 *x, size_t n, uint32_t *grid, uint32_t ncells, float *sink)
{
  unsigned long long pol_first, pol_last;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  float acc = 0;
  uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + 8 * stride < n; i += 9 * stride) {
    float4 v[9];
#pragma unroll
    for (int q = 0; q < 9; q++) v[q] = ld_stream<SH>(x + i + q * stride, pol_first);
#pragma unroll
    for (int q = 0; q < 9; q++) acc += v[q].x + v[q].y + v[q].z + v[q].w;
#pragma unroll
    for (int q = 0; q < NUM; q++) {
      s = s * 1664525u + 1013904223u;
      red_add<RH>(grid + (uint32_t)(((uint64_t)mix(s ^ __float_as_uint(acc)) * ncells) >> 32), 1u, pol_last);
    }
  }
  if (acc == 1234.5f) *sink = acc;
}

template <int RH, int SH, int NUM>
static void run(const char *tag, cudaStream_t st, const float4 *x, size_t bytes, uint32_t *g, size_t wmb, float *sink, int sms, int carve)
{
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(a, st));
    k<RH, SH, NUM><<<sms * 16, 256, 0, st>>>(x, bytes / 16, g, (uint32_t)((wmb << 20) / 4), sink);
    CK(cudaEventRecord(b, st));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0) best = std::min(best, ms);
  }
  printf("{\"window_mb\": %zu, \"carveout\": %d, \"reds_per_36B\": %.2f, \"variant\": \"%s\", \"read_gbs\": %.0f, \"red_gops\": %.1f}\n", wmb, carve,
         NUM / 4.0, tag, bytes / best * 1e-6, (double)(bytes / 16) * NUM / 9 / best * 1e-6);
  fflush(stdout);
  cudaEventDestroy(a); cudaEventDestroy(b);
}

int main()
{
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"l2_bytes\": %d, \"persisting_l2_max\": %d, \"access_policy_max_window\": %d}\n", prop.name, prop.l2CacheSize,
         prop.persistingL2CacheMaxSize, prop.accessPolicyMaxWindowSize);
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  const size_t bytes = (size_t)3 << 30;
  float4 *x; float *sink;
  CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(x, 0, bytes));
  for (int carve = 0; carve < 2; carve++) {
    for (size_t wmb : {16, 32, 48, 64, 80, 96}) {
      uint32_t *g;
      CK(cudaMalloc(&g, wmb << 20));
      CK(cudaMemset(g, 0, wmb << 20));
      if (carve) {
        const size_t want = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, (wmb << 20) + ((size_t)4 << 20));
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
        cudaStreamAttrValue av = {};
        av.accessPolicyWindow.base_ptr = g;
        av.accessPolicyWindow.num_bytes = std::min<size_t>(wmb << 20, (size_t)prop.accessPolicyMaxWindowSize);
        av.accessPolicyWindow.hitRatio = std::min(1.0f, (float)((double)want / (double)(wmb << 20)));
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
      }
      run<0, 0, 4>("red plain, stream ldcs", st, x, bytes, g, wmb, sink, sms, carve);
      run<0, 1, 4>("red plain, stream evict_first", st, x, bytes, g, wmb, sink, sms, carve);
      run<1, 1, 4>("red evict_last, stream evict_first", st, x, bytes, g, wmb, sink, sms, carve);
      run<1, 2, 4>("red evict_last, stream evict_first+L1 no_allocate", st, x, bytes, g, wmb, sink, sms, carve);
      run<1, 0, 4>("red evict_last, stream ldcs", st, x, bytes, g, wmb, sink, sms, carve);
      run<1, 1, 12>("ALL ATOMS red evict_last, stream evict_first", st, x, bytes, g, wmb, sink, sms, carve);
      run<0, 0, 12>("ALL ATOMS red plain, stream ldcs", st, x, bytes, g, wmb, sink, sms, carve);
      if (carve) {
        cudaStreamAttrValue av = {};
        av.accessPolicyWindow.num_bytes = 0;
        CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
        CK(cudaCtxResetPersistingL2Cache());
      }
      CK(cudaFree(g));
    }
  }
  return 0;
}
