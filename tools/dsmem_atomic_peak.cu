// Microbenchmark: random-address red.shared::cluster.add.u32 into the shared memory of the CTAs of a
// thread-block cluster (distributed shared memory), all SMs busy, as a function of the cluster size.
// It answers one design question of DESIGN.md section 3: can a cluster-wide privatised copy of a grid
// that is too large for one SM (C2: 3.4 MB u32, 1.7 MB u16) take hits faster than the L2 (191 G RED/s)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dsmem_atomic_peak tools/dsmem_atomic_peak.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t cluster_nctarank()
{
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_sync()
{
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// LOCAL = 1: every hit goes to this CTA's own shared memory (baseline); 0: uniformly over the cluster's CTAs
template <int LOCAL>
__global__ void __launch_bounds__(1024) dsmem_red(uint32_t *out, int iters, int ncells)
{
  extern __shared__ uint32_t h[];
  for (int i = threadIdx.x; i < ncells; i += blockDim.x) h[i] = 0;
  const uint32_t ncta = cluster_nctarank();
  cluster_sync();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(h);
  uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      s = s * 1664525u + 1013904223u;
      const uint32_t r = mix(s);
      const uint32_t cell = (uint32_t)(((uint64_t)r * (uint32_t)ncells) >> 32);
      if (LOCAL) {
        atomicAdd(&h[cell], 1u);
      } else {
        const uint32_t rank = (r & 0xffffu) * ncta >> 16;
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(base + 4u * cell), "r"(rank));
        asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(remote), "r"(1u) : "memory");
      }
    }
  }
  cluster_sync();
  uint32_t acc = 0;
  for (int i = threadIdx.x; i < ncells; i += blockDim.x) acc += h[i];
  acc = __reduce_add_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

template <int LOCAL>
static void run(int cl, int threads, int ncells, int sms, uint32_t *out, bool &first)
{
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  auto kern = dsmem_red<LOCAL>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ncells * 4));
  if (cl > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = ncells * 4;
  cfg.gridDim = dim3(cl);
  int nclusters = 0;
  CK(cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg));
  if (nclusters < 1) { fprintf(stderr, "cluster %d: not launchable\n", cl); return; }
  const int blocks = nclusters * cl;
  cfg.gridDim = dim3(blocks);
  const int iters = 256;
  float best = 1e30f;
  uint32_t total = 0;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaMemset(out, 0, 4));
    CK(cudaEventRecord(a));
    CK(cudaLaunchKernelEx(&cfg, kern, out, iters, ncells));
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
    CK(cudaMemcpy(&total, out, 4, cudaMemcpyDeviceToHost));
  }
  const double ops = (double)blocks * threads * iters * 8;
  printf("%s\n  {\"cluster\": %d, \"local\": %d, \"threads\": %d, \"clusters_resident\": %d, \"ctas\": %d, \"sms\": %d, "
         "\"gops\": %.2f, \"sum_ok\": %s}",
         first ? "" : ",", cl, LOCAL, threads, nclusters, blocks, sms, ops / best * 1e-6,
         total == (uint32_t)ops ? "true" : "false");
  first = false;
}

int main()
{
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  uint32_t *out; CK(cudaMalloc(&out, 4));
  printf("{\"gpu\": \"%s\", \"what\": \"random red.shared::cluster.add.u32, 26800 u32 cells (107 KB) per CTA\", \"runs\": [", prop.name);
  bool first = true;
  const int ncells = 26800;
  run<1>(1, 1024, ncells, sms, out, first);
  for (int cl : {2, 4, 8, 16}) {
    run<0>(cl, 1024, ncells, sms, out, first);
    run<0>(cl, 256, ncells, sms, out, first);
  }
  printf("]}\n");
  return 0;
}
