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

}  // namespace gcb

// the collective operations gcb_counter_allreduce needs; bound to NCCL in nccl_reduce.cpp
struct gcb_reduce_ops {
  int (*allreduce)(const void *send, void *recv, size_t count, int nccl_dtype, void *comm, void *stream);
  int (*group_start)(void);
  int (*group_end)(void);
};

#define GCB_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return gcb::fail(GCB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),    \
                       __FILE__, __LINE__);                                                     \
  } while (0)
