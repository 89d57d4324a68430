// Internal helpers shared by the translation units of libgridcount_b200.so.
#pragma once
#include "gridcount_b200.h"
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace gcb {

// thread-local message behind gcb_last_error()
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);

inline size_t cells(const gcb_tgrid &tg) { return (size_t)tg.mx[0] * tg.mx[1] * tg.mx[2]; }

// value of n sequential additions of w to v (g_ri3Dc.c:149 in f32, :151/:426 in f64)
float seq_add_f32(float v, float w, uint64_t n);
double seq_add_f64(double v, double w, uint64_t n);

// host_pack.cpp: x[frame][natoms][3] -> out[frame][gnx][3] on a pool of host threads (step > 0: the group is
// g0 + i*step and gindex is not read; threads <= 0: pack_threads_default())
int pack_threads_default();
int pack_cpus();
void pack_frames(const float *x, long nframes, int natoms, const int *gindex, int gnx, int g0, int step, float *out,
                 int threads);
// the same in two halves: pack_begin returns while the pool's threads work (NULL: a small job, already done);
// pack_end lets the caller join in, waits, and releases the pool.  One job at a time, both calls from one thread.
void *pack_begin(const float *x, long nframes, int natoms, const int *gindex, int gnx, int g0, int step, float *out,
                 int threads);
void pack_end(void *handle);

}  // namespace gcb

// The collective operations gcb_counter_allreduce needs, stream-ordered like NCCL's (dtype: ncclUint32 = 3,
// ncclUint64 = 5, ncclFloat32 = 7).  Bound to NCCL in nccl_reduce.cpp (one process per GPU) and to the in-process
// transport of local_reduce.cu (one host thread per counter, any devices).
struct gcb_reduce_ops {
  int (*allreduce)(const void *send, void *recv, size_t count, int nccl_dtype, void *comm, void *stream);
  int (*allgather)(const void *send, void *recv, size_t count_per_rank, int nccl_dtype, void *comm, void *stream);
  int (*group_start)(void);
  int (*group_end)(void);
  void (*destroy)(void *comm);
};
// binning.cu: selects the counter's device (comm == NULL) / attaches a communicator, its operations and control plane
extern "C" int gcb_counter_comm_attach_(gcb_counter *c, void *comm, const gcb_reduce_ops *ops, int rank, int nranks, gcb_ctl *ctl);

// host control plane between the ranks of one node (shm_ctl.cpp); declared in gridcount_b200.h
#define GCB_CTL_WORDS 8

#define GCB_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return gcb::fail(GCB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),    \
                       __FILE__, __LINE__);                                                     \
  } while (0)
