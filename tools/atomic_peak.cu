// Microbenchmarks that give the roofline its second denominator:
//   * random-address red.global.add.u32 throughput as a function of the window size
//     (L2-resident windows vs windows larger than the 126 MB L2),
//   * shared-memory atomic throughput,
//   * HBM streaming read bandwidth (float4 LDG),
// all SMs busy, timed with CUDA events.  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/atomic_peak tools/atomic_peak.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

template <int ILP>
__global__ void __launch_bounds__(256) red_random(uint32_t *grid, uint32_t ncells, int iters)
{
  uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      s = s * 1664525u + 1013904223u;
      const uint32_t cell = (uint32_t)(((uint64_t)mix(s) * ncells) >> 32);
      atomicAdd(grid + cell, 1u);
    }
  }
}

// returning atomics on packed 8-bit counters: the narrow-counter path of K1 adds 1 << 8*(cell & 3) to the word
// and looks at the old byte (spill at 127, carry alarm at 255)
template <int ILP>
__global__ void __launch_bounds__(256) atom_u8_random(uint32_t *grid, uint32_t ncells, int iters, uint32_t *sink)
{
  uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  uint32_t seen = 0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      s = s * 1664525u + 1013904223u;
      const uint32_t cell = (uint32_t)(((uint64_t)mix(s) * ncells) >> 32);
      const uint32_t sh = 8u * (cell & 3u);
      const uint32_t old = atomicAdd(grid + (cell >> 2), 1u << sh);
      const uint32_t ob = (old >> sh) & 255u;
      if (ob == 127u) atomicSub(grid + (cell >> 2), 128u << sh);
      seen |= ob == 255u ? 1u : 0u;
    }
  }
  if (seen) *sink = 1;
}

__global__ void __launch_bounds__(1024) smem_atomics(uint32_t *out, int iters, int ncells)
{
  extern __shared__ uint32_t h[];
  for (int i = threadIdx.x; i < ncells; i += blockDim.x) h[i] = 0;
  __syncthreads();
  uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      s = s * 1664525u + 1013904223u;
      atomicAdd(&h[(uint32_t)(((uint64_t)mix(s) * (uint32_t)ncells) >> 32)], 1u);
    }
  }
  __syncthreads();
  uint32_t acc = 0;
  for (int i = threadIdx.x; i < ncells; i += blockDim.x) acc += h[i];
  if (acc == 0xffffffffu) out[0] = acc;
}

__global__ void __launch_bounds__(256) stream_read(const float4 *x, size_t n, float *sink)
{
  float acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(x + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 1234.5f) *sink = acc;
}

// The mix K1 really sends to the L2: a coalesced evict-first stream of coordinate bytes plus random REDs, NUM REDs
// per DEN float4 loads (1-of-3 water: 1 RED per 36 B = 4 per 9 loads; all atoms: 1 per 12 B = 4 per 3 loads).
// No conversion work, no shared memory: what the memory system sustains for this mix, whatever the kernel around it.
template <int NUM, int DEN>
__global__ void __launch_bounds__(256) stream_plus_red(const float4 *x, size_t n, uint32_t *grid, uint32_t ncells, float *sink)
{
  float acc = 0;
  uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
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
      atomicAdd(grid + (uint32_t)(((uint64_t)mix(s ^ __float_as_uint(acc)) * ncells) >> 32), 1u);
    }
  }
  if (acc == 1234.5f) *sink = acc;
}

template <int NUM, int DEN>
static void run_mix(const float4 *x, size_t bytes, uint32_t *g, uint32_t ncells, float *sink, int sms, cudaEvent_t a, cudaEvent_t b,
                    size_t window_mb, bool &first)
{
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(a));
    stream_plus_red<NUM, DEN><<<sms * 16, 256>>>(x, bytes / 16, g, ncells, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  const double loads = (double)(bytes / 16);
  printf("%s{\"window_mb\": %zu, \"reds_per_36B\": %.3f, \"read_gbs\": %.1f, \"red_gops\": %.2f}", first ? "" : ",\n  ", window_mb,
         (double)NUM / DEN * 2.25, bytes / best * 1e-6, loads * NUM / DEN / best * 1e-6);
  first = false;
}

int main()
{
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d,\n \"red_u32_random\": [", prop.name, sms, prop.l2CacheSize);
  const size_t windows_mb[] = {1, 4, 16, 64, 96, 256, 512};
  bool first = true;
  for (size_t mb : windows_mb) {
    const size_t bytes = mb << 20;
    uint32_t *g;
    CK(cudaMalloc(&g, bytes));
    CK(cudaMemset(g, 0, bytes));
    const uint32_t ncells = (uint32_t)(bytes / 4);
    const int blocks = sms * 8, iters = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
      CK(cudaEventRecord(a));
      red_random<8><<<blocks, 256>>>(g, ncells, iters);
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (rep > 0 && ms < best) best = ms;
    }
    const double ops = (double)blocks * 256 * iters * 8;
    printf("%s{\"window_mb\": %zu, \"gops\": %.2f}", first ? "" : ", ", mb, ops / best * 1e-6);
    first = false;
    CK(cudaFree(g));
  }
  printf("],\n \"atom_u8_random\": [");
  first = true;
  for (size_t cells_m : {16, 64, 96, 125}) {                 // M cells of one byte each
    const size_t bytes = cells_m << 20;
    uint32_t *g, *sink;
    CK(cudaMalloc(&g, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(g, 0, bytes)); CK(cudaMemset(sink, 0, 4));
    const uint32_t ncells = (uint32_t)bytes;
    const int blocks = sms * 8, iters = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
      CK(cudaMemset(g, 0, bytes));
      CK(cudaEventRecord(a));
      atom_u8_random<8><<<blocks, 256>>>(g, ncells, iters, sink);
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (rep > 0 && ms < best) best = ms;
    }
    const double ops = (double)blocks * 256 * iters * 8;
    printf("%s{\"cells_m\": %zu, \"gops\": %.2f}", first ? "" : ", ", cells_m, ops / best * 1e-6);
    first = false;
    CK(cudaFree(g)); CK(cudaFree(sink));
  }
  printf("],\n");
  {
    uint32_t *out; CK(cudaMalloc(&out, 4));
    const int ncells = 32768;   // 128 KiB of u32 bins per CTA
    CK(cudaFuncSetAttribute(smem_atomics, cudaFuncAttributeMaxDynamicSharedMemorySize, ncells * 4));
    float best = 1e30f;
    const int blocks = sms, iters = 512;
    for (int rep = 0; rep < 5; rep++) {
      CK(cudaEventRecord(a));
      smem_atomics<<<blocks, 1024, ncells * 4>>>(out, iters, ncells);
      CK(cudaGetLastError());
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (rep > 0 && ms < best) best = ms;
    }
    printf(" \"smem_atomic_random_gops\": %.2f,\n", (double)blocks * 1024 * iters * 8 / best * 1e-6);
    CK(cudaFree(out));
  }
  {
    const size_t bytes = (size_t)4 << 30;
    float4 *x; float *sink;
    CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(x, 0, bytes));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
      CK(cudaEventRecord(a));
      stream_read<<<sms * 16, 256>>>(x, bytes / 16, sink);
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (rep > 0 && ms < best) best = ms;
    }
    printf(" \"hbm_read_gbs\": %.1f,\n \"stream_plus_red\": [", bytes / best * 1e-6);
    bool f2 = true;
    for (size_t mb : {4, 64}) {
      uint32_t *g;
      CK(cudaMalloc(&g, mb << 20));
      CK(cudaMemset(g, 0, mb << 20));
      const uint32_t nc = (uint32_t)((mb << 20) / 4);
      run_mix<0, 9>(x, bytes, g, nc, sink, sms, a, b, mb, f2);
      run_mix<1, 9>(x, bytes, g, nc, sink, sms, a, b, mb, f2);
      run_mix<2, 9>(x, bytes, g, nc, sink, sms, a, b, mb, f2);
      run_mix<4, 9>(x, bytes, g, nc, sink, sms, a, b, mb, f2);
      run_mix<8, 9>(x, bytes, g, nc, sink, sms, a, b, mb, f2);
      run_mix<4, 3>(x, bytes, g, nc, sink, sms, a, b, mb, f2);
      CK(cudaFree(g));
    }
    printf("]}\n");
  }
  return 0;
}
