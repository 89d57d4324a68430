// Frame-sharded multi-GPU: integer sum of the per-rank private grids over NCCL
// (NVLink 5 / NVSwitch).  NCCL is bound lazily with dlopen so that the library
// loads on hosts without it and shares torch's copy when one is already mapped.
#include "gcb_internal.h"
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdlib>

// minimal NCCL declarations (stable C ABI of libnccl.so.2)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5,
               ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;

namespace {
struct Nccl {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
Nccl g_nccl;

int load_nccl()
{
  if (g_nccl.lib) return GCB_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return gcb::fail(GCB_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(g_nccl.lib, "ncclAllGather");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(g_nccl.lib, "ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(g_nccl.lib, "ncclGroupEnd");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.GroupStart || !g_nccl.GroupEnd)
    return gcb::fail(GCB_ERR_NCCL, "libnccl.so.2 lacks the expected symbols");
  return GCB_OK;
}
}  // namespace

#define GCB_NCCL(call)                                                                                    \
  do {                                                                                                    \
    ncclResult_t r_ = (call);                                                                             \
    if (r_ != ncclSuccess)                                                                                \
      return gcb::fail(GCB_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
  } while (0)

static int do_allreduce(const void *send, void *recv, size_t count, int dtype, void *comm, void *stream)
{
  GCB_NCCL(g_nccl.AllReduce(send, recv, count, (ncclDataType_t)dtype, ncclSum, (ncclComm_t)comm, (cudaStream_t)stream));
  return GCB_OK;
}

static int do_allgather(const void *send, void *recv, size_t count_per_rank, int dtype, void *comm, void *stream)
{
  GCB_NCCL(g_nccl.AllGather(send, recv, count_per_rank, (ncclDataType_t)dtype, (ncclComm_t)comm, (cudaStream_t)stream));
  return GCB_OK;
}

// called by the counter's destructor
static void do_destroy(void *comm)
{
  if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)comm);
}

static int do_group_start() { GCB_NCCL(g_nccl.GroupStart()); return GCB_OK; }
static int do_group_end() { GCB_NCCL(g_nccl.GroupEnd()); return GCB_OK; }
static const gcb_reduce_ops nccl_ops = {do_allreduce, do_allgather, do_group_start, do_group_end, do_destroy};

extern "C" {

int gcb_nccl_unique_id(char id_out[128])
{
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  GCB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_out, id.internal, 128);
  return GCB_OK;
}

int gcb_counter_comm_init(gcb_counter *c, const char id[128], int rank, int nranks)
{
  if (!c || !id || rank < 0 || rank >= nranks) return gcb::fail(GCB_ERR_ARG, "gcb_counter_comm_init: bad argument");
  int rc = load_nccl();
  if (rc) return rc;
  rc = gcb_counter_comm_attach_(c, nullptr, nullptr, rank, nranks, nullptr);   // selects the device
  if (rc) return rc;
  ncclUniqueId uid;
  memcpy(uid.internal, id, 128);
  ncclComm_t comm;
  GCB_NCCL(g_nccl.CommInitRank(&comm, nranks, uid, rank));
  // control plane: a shared-memory all-gather between the ranks of this node (shm_ctl.cpp).  The communicator is up,
  // so every rank is alive; ranks that cannot attach within the timeout (other nodes) leave every rank without it,
  // and the rank table then travels over NCCL as a small all-reduce.
  gcb_ctl *ctl = nullptr;
  if (!getenv("GCB_NO_SHM_CTL") && nranks > 1) {
    if (gcb_ctl_open(&ctl, id, rank, nranks, GCB_CTL_WORDS, 20000) == GCB_OK) {
      uint64_t mine[GCB_CTL_WORDS] = {1}, all[GCB_CTL_WORDS * 64];
      if (nranks > 64 || gcb_ctl_allgather(ctl, mine, all, 20000) != GCB_OK) { gcb_ctl_close(ctl); ctl = nullptr; }
    } else {
      ctl = nullptr;
    }
  }
  return gcb_counter_comm_attach_(c, comm, &nccl_ops, rank, nranks, ctl);
}

}  // extern "C"
