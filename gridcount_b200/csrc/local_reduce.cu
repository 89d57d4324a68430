// In-process transport for gcb_counter_allreduce: the ranks are counters driven by threads of ONE process (on one
// GPU or several), and the collectives are stream-ordered copies through a staging buffer on the first rank's device
// plus a rank-ordered sum kernel.  It binds the same gcb_reduce_ops as the NCCL transport (nccl_reduce.cpp), so the
// whole reduce protocol of binning.cu -- rank table over the shared-memory control plane, saturation guard, grid
// sum, in-place gather of the per-frame statistics, the frame-ordered sumP chain -- runs unchanged.  NCCL refuses
// two ranks on one device; this transport is what lets the frame-sharded path (the reference's manual equivalent is
// the grid merge of a_gridcalc.c:180-189) be exercised on a single-GPU box, and it serves a host program that
// drives several GPUs from one process without NCCL.  Throughput is not its purpose.
#include "gcb_internal.h"
#include <cuda_runtime.h>
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <new>
#include <random>
#include <vector>

namespace {

struct LocalArg {
  const void *send = nullptr;
  void *recv = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ready = nullptr;
  int device = 0;
  size_t count = 0;
  int dtype = 0, kind = 0;
};

}  // namespace

struct gcb_local_comm {
  int nranks = 0;
  char id[128];
  std::mutex mu;
  std::condition_variable cv;
  uint64_t generation = 0;
  int arrived = 0, attached = 0, status = GCB_OK;
  std::vector<LocalArg> args;
  int root_device = -1;
  cudaStream_t work = nullptr;
  cudaEvent_t done[2] = {nullptr, nullptr};
  void *stage = nullptr;
  size_t stage_bytes = 0;
};

namespace {

struct LocalRank {
  gcb_local_comm *lc;
  int rank, device;
  cudaEvent_t ready;
};

size_t dtype_bytes(int dtype) { return dtype == 5 ? 8 : 4; }

// stage[0][i] = stage[0][i] + stage[1][i] + ... in rank order (an f32 sum is then the same on every run)
template <class T>
__global__ void local_sum_kernel(T *__restrict__ stage, size_t count, int nranks)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    T v = stage[i];
    for (int r = 1; r < nranks; r++) v += stage[(size_t)r * count + i];
    stage[i] = v;
  }
}

// the last rank to arrive queues the whole collective on the work stream
int perform(gcb_local_comm *lc)
{
  const LocalArg &a0 = lc->args[0];
  for (int r = 1; r < lc->nranks; r++)
    if (lc->args[r].count != a0.count || lc->args[r].dtype != a0.dtype || lc->args[r].kind != a0.kind)
      return gcb::fail(GCB_ERR_STATE, "in-process collective: the ranks disagree (rank %d: kind %d, %zu x dtype %d; rank 0: kind %d, %zu x dtype %d)",
                       r, lc->args[r].kind, lc->args[r].count, lc->args[r].dtype, a0.kind, a0.count, a0.dtype);
  const size_t bytes = a0.count * dtype_bytes(a0.dtype), R = (size_t)lc->nranks;
  GCB_CUDA(cudaSetDevice(lc->root_device));
  if (R * bytes > lc->stage_bytes) {
    GCB_CUDA(cudaStreamSynchronize(lc->work));
    cudaFree(lc->stage);
    lc->stage = nullptr; lc->stage_bytes = 0;
    GCB_CUDA(cudaMalloc(&lc->stage, R * bytes + 256));
    lc->stage_bytes = R * bytes;
  }
  if (bytes == 0) return GCB_OK;
  uint8_t *st = static_cast<uint8_t *>(lc->stage);
  for (int r = 0; r < lc->nranks; r++) {
    GCB_CUDA(cudaStreamWaitEvent(lc->work, lc->args[r].ready, 0));
    GCB_CUDA(cudaMemcpyPeerAsync(st + (size_t)r * bytes, lc->root_device, lc->args[r].send, lc->args[r].device, bytes, lc->work));
  }
  size_t out_bytes = R * bytes;                  // all-gather: every rank receives all blocks
  if (a0.kind == 0) {
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((a0.count + 255) / 256, 148 * 8));
    if (a0.dtype == 3) local_sum_kernel<uint32_t><<<grid, 256, 0, lc->work>>>((uint32_t *)st, a0.count, lc->nranks);
    else if (a0.dtype == 5) local_sum_kernel<unsigned long long><<<grid, 256, 0, lc->work>>>((unsigned long long *)st, a0.count, lc->nranks);
    else if (a0.dtype == 7) local_sum_kernel<float><<<grid, 256, 0, lc->work>>>((float *)st, a0.count, lc->nranks);
    else return gcb::fail(GCB_ERR_ARG, "in-process all-reduce: dtype %d not supported", a0.dtype);
    GCB_CUDA(cudaGetLastError());
    out_bytes = bytes;
  }
  for (int r = 0; r < lc->nranks; r++)
    GCB_CUDA(cudaMemcpyPeerAsync(lc->args[r].recv, lc->args[r].device, st, lc->root_device, out_bytes, lc->work));
  GCB_CUDA(cudaEventRecord(lc->done[lc->generation & 1], lc->work));
  return GCB_OK;
}

int collective(int kind, const void *send, void *recv, size_t count, int dtype, void *comm, void *stream)
{
  LocalRank *lr = static_cast<LocalRank *>(comm);
  gcb_local_comm *lc = lr->lc;
  GCB_CUDA(cudaSetDevice(lr->device));
  GCB_CUDA(cudaEventRecord(lr->ready, (cudaStream_t)stream));
  std::unique_lock<std::mutex> lk(lc->mu);
  LocalArg &a = lc->args[(size_t)lr->rank];
  a.send = send; a.recv = recv; a.stream = (cudaStream_t)stream; a.ready = lr->ready; a.device = lr->device;
  a.count = count; a.dtype = dtype; a.kind = kind;
  const uint64_t gen = lc->generation;
  if (++lc->arrived == lc->nranks) {
    lc->status = perform(lc);
    cudaSetDevice(lr->device);
    lc->arrived = 0;
    lc->generation++;
    lc->cv.notify_all();
  } else {
    lc->cv.wait(lk, [&] { return lc->generation != gen; });
  }
  const int rc = lc->status;
  if (rc) return rc;
  if (count) GCB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, lc->done[gen & 1], 0));
  return GCB_OK;
}

int local_allreduce(const void *send, void *recv, size_t count, int dtype, void *comm, void *stream)
{
  return collective(0, send, recv, count, dtype, comm, stream);
}
int local_allgather(const void *send, void *recv, size_t count_per_rank, int dtype, void *comm, void *stream)
{
  return collective(1, send, recv, count_per_rank, dtype, comm, stream);
}
int local_group(void) { return GCB_OK; }
void local_destroy(void *comm)
{
  LocalRank *lr = static_cast<LocalRank *>(comm);
  if (!lr) return;
  if (lr->ready) cudaEventDestroy(lr->ready);
  delete lr;
}
const gcb_reduce_ops local_ops = {local_allreduce, local_allgather, local_group, local_group, local_destroy};

}  // namespace

extern "C" {

int gcb_local_comm_create(gcb_local_comm **out, int nranks)
{
  if (!out || nranks < 1 || nranks > 64) return gcb::fail(GCB_ERR_ARG, "gcb_local_comm_create: 1..64 ranks");
  gcb_local_comm *lc = new (std::nothrow) gcb_local_comm;
  if (!lc) return gcb::fail(GCB_ERR_NOMEM, "out of memory");
  lc->nranks = nranks;
  lc->args.resize((size_t)nranks);
  std::random_device rd;                                    // names the shared-memory control block of this job
  for (int i = 0; i < 128; i += 4) { const unsigned v = rd(); memcpy(lc->id + i, &v, 4); }
  *out = lc;
  return GCB_OK;
}

// Called once per rank, from `nranks` threads at the same time (the control plane waits for all of them).
int gcb_counter_comm_init_local(gcb_counter *c, gcb_local_comm *lc, int rank)
{
  if (!c || !lc || rank < 0 || rank >= lc->nranks) return gcb::fail(GCB_ERR_ARG, "gcb_counter_comm_init_local: bad argument");
  int rc = gcb_counter_comm_attach_(c, nullptr, nullptr, rank, lc->nranks, nullptr);     // selects the counter's device
  if (rc) return rc;
  LocalRank *lr = new (std::nothrow) LocalRank;
  if (!lr) return gcb::fail(GCB_ERR_NOMEM, "out of memory");
  lr->lc = lc; lr->rank = rank; lr->ready = nullptr;
  GCB_CUDA(cudaGetDevice(&lr->device));
  GCB_CUDA(cudaEventCreateWithFlags(&lr->ready, cudaEventDisableTiming));
  {
    std::lock_guard<std::mutex> lk(lc->mu);
    if (rank == 0) {
      lc->root_device = lr->device;
      GCB_CUDA(cudaStreamCreateWithFlags(&lc->work, cudaStreamNonBlocking));
      for (auto &e : lc->done) GCB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    lc->attached++;
  }
  gcb_ctl *ctl = nullptr;
  if (lc->nranks > 1) {
    rc = gcb_ctl_open(&ctl, lc->id, rank, lc->nranks, GCB_CTL_WORDS, 60000);
    if (rc) { local_destroy(lr); return rc; }
  }
  return gcb_counter_comm_attach_(c, lr, &local_ops, rank, lc->nranks, ctl);
}

// after the counters that used it are destroyed
void gcb_local_comm_destroy(gcb_local_comm *lc)
{
  if (!lc) return;
  if (lc->root_device >= 0) {
    cudaSetDevice(lc->root_device);
    if (lc->work) { cudaStreamSynchronize(lc->work); cudaStreamDestroy(lc->work); }
    for (auto e : lc->done) if (e) cudaEventDestroy(e);
    cudaFree(lc->stage);
  }
  cudaGetLastError();
  delete lc;
}

}  // extern "C"
