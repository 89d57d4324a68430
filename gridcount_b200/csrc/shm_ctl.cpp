// Control plane of the frame-sharded multi-GPU run: a tiny all-gather between the ranks of ONE node over POSIX
// shared memory, so that what the ranks must agree on before the grid all-reduce (which buffers to reduce, frame
// counts, saturation bounds: all host-side knowledge) does not travel as a second NCCL collective.  An NCCL kernel
// for a 56-byte-per-rank table has to find an SM while the persistent binning CTAs hold every SM's shared memory,
// and the host has to wait for it: in round 1 that exchange, not the 3.4 MB grid reduce, limited 2-GPU scaling.
// Data plane (grids, per-frame counters) stays on NCCL / NVLink.
//
// Region layout: header | seq[nranks] (one cache line each) | table[2][nranks][words].
// Exchange k (k = 1, 2, ...): rank r writes table[k & 1][r], then publishes seq[r] = k (release); it waits for
// seq[j] >= k for every j (acquire) and reads table[k & 1][*].  A rank starts exchange k + 2 (same parity) only
// after it saw every seq[j] >= k + 1, i.e. after every rank finished reading exchange k: two buffers suffice.
#include "gcb_internal.h"
#include <atomic>
#include <cerrno>
#include <cstdlib>
#include <ctime>
#include <fcntl.h>
#include <new>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

constexpr uint32_t CTL_MAGIC = 0x67636231u;          // "gcb1"

struct CtlHeader {
  std::atomic<uint32_t> magic;
  uint32_t nranks, words;
  std::atomic<uint32_t> attached;
  uint8_t pad[48];
};
struct alignas(64) CtlSeq { std::atomic<uint64_t> v; uint8_t pad[56]; };
static_assert(sizeof(CtlHeader) == 64 && sizeof(CtlSeq) == 64, "control block layout");

double now_s()
{
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

size_t region_bytes(int nranks, int words)
{
  return sizeof(CtlHeader) + sizeof(CtlSeq) * (size_t)nranks + sizeof(uint64_t) * 2 * (size_t)nranks * (size_t)words;
}

}  // namespace

struct gcb_ctl {
  void *base = nullptr;
  size_t bytes = 0;
  int rank = 0, nranks = 0, words = 0;
  uint64_t round = 0;
  CtlHeader *hdr() const { return static_cast<CtlHeader *>(base); }
  CtlSeq *seq() const { return reinterpret_cast<CtlSeq *>(static_cast<uint8_t *>(base) + sizeof(CtlHeader)); }
  uint64_t *table(int parity) const
  {
    return reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(base) + sizeof(CtlHeader) + sizeof(CtlSeq) * (size_t)nranks) +
           (size_t)parity * (size_t)nranks * (size_t)words;
  }
};

extern "C" {

void gcb_ctl_close(gcb_ctl *h)
{
  if (!h) return;
  if (h->base) munmap(h->base, h->bytes);
  delete h;
}

int gcb_ctl_open(gcb_ctl **out, const char id[128], int rank, int nranks, int words, int timeout_ms)
{
  if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks || words < 1 || words > 4096)
    return gcb::fail(GCB_ERR_ARG, "gcb_ctl_open: bad argument");
  uint64_t hash = 1469598103934665603ull;            // FNV-1a over the job's unique id
  for (int i = 0; i < 128; i++) { hash ^= (uint8_t)id[i]; hash *= 1099511628211ull; }
  char name[64];
  snprintf(name, sizeof(name), "/gcb_ctl_%016llx", (unsigned long long)hash);
  gcb_ctl *h = new (std::nothrow) gcb_ctl;
  if (!h) return gcb::fail(GCB_ERR_NOMEM, "out of memory");
  h->rank = rank; h->nranks = nranks; h->words = words; h->bytes = region_bytes(nranks, words);
  const double deadline = now_s() + 1e-3 * (double)timeout_ms;
  int fd = -1;
  if (rank == 0) {
    shm_unlink(name);                                  // a stale segment of a crashed job with the same id
    fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)h->bytes) != 0) {
      const int e = errno;
      if (fd >= 0) { close(fd); shm_unlink(name); }
      delete h;
      return gcb::fail(GCB_ERR_IO, "cannot create shared-memory control block %s: %s", name, strerror(e));
    }
  } else {
    for (;;) {
      fd = shm_open(name, O_RDWR, 0600);
      struct stat sb;
      if (fd >= 0 && fstat(fd, &sb) == 0 && (size_t)sb.st_size >= h->bytes) break;
      if (fd >= 0) { close(fd); fd = -1; }
      if (now_s() > deadline) { delete h; return gcb::fail(GCB_ERR_IO, "control block %s did not appear (ranks on different nodes?)", name); }
      usleep(200);
    }
  }
  h->base = mmap(nullptr, h->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (h->base == MAP_FAILED) {
    h->base = nullptr;
    if (rank == 0) shm_unlink(name);
    delete h;
    return gcb::fail(GCB_ERR_IO, "cannot map the control block %s", name);
  }
  CtlHeader *hd = h->hdr();
  if (rank == 0) {                                     // (ftruncate zero-filled the region)
    hd->nranks = (uint32_t)nranks; hd->words = (uint32_t)words;
    hd->magic.store(CTL_MAGIC, std::memory_order_release);
  } else {
    while (hd->magic.load(std::memory_order_acquire) != CTL_MAGIC) {
      if (now_s() > deadline) { gcb_ctl_close(h); return gcb::fail(GCB_ERR_IO, "control block %s was never initialised", name); }
      sched_yield();
    }
    if (hd->nranks != (uint32_t)nranks || hd->words != (uint32_t)words) {
      const unsigned nr = hd->nranks, nw = hd->words;
      gcb_ctl_close(h);
      return gcb::fail(GCB_ERR_ARG, "control block %s was created for %u ranks x %u words", name, nr, nw);
    }
  }
  hd->attached.fetch_add(1, std::memory_order_acq_rel);
  // everybody is mapped before the name disappears; after that the segment lives as long as a mapping does
  while (hd->attached.load(std::memory_order_acquire) < (uint32_t)nranks) {
    if (now_s() > deadline) {
      const unsigned got = hd->attached.load();
      if (rank == 0) shm_unlink(name);
      gcb_ctl_close(h);
      return gcb::fail(GCB_ERR_IO, "only %u of %d ranks attached to the control block %s", got, nranks, name);
    }
    sched_yield();
  }
  if (rank == 0) shm_unlink(name);
  *out = h;
  return GCB_OK;
}

int gcb_ctl_allgather(gcb_ctl *h, const uint64_t *mine, uint64_t *all, int timeout_ms)
{
  if (!h || !mine || !all) return gcb::fail(GCB_ERR_ARG, "gcb_ctl_allgather: null argument");
  const uint64_t k = ++h->round;
  uint64_t *tab = h->table((int)(k & 1));
  memcpy(tab + (size_t)h->rank * h->words, mine, sizeof(uint64_t) * (size_t)h->words);
  h->seq()[h->rank].v.store(k, std::memory_order_release);
  // timeout_ms <= 0: wait as long as it takes (a rank may still be binning its shard for minutes: the same patience
  // a collective has).  Spin briefly, then yield, then sleep: the other ranks need the cores.
  const double t_begin = now_s(), deadline = t_begin + 1e-3 * (double)timeout_ms;
  for (int j = 0; j < h->nranks; j++) {
    unsigned spins = 0;
    while (h->seq()[j].v.load(std::memory_order_acquire) < k) {
      if ((++spins & 1023u) == 0) {
        const double t = now_s();
        if (timeout_ms > 0 && t > deadline)
          return gcb::fail(GCB_ERR_NCCL, "control all-gather %llu: rank %d never arrived", (unsigned long long)k, j);
        if (t - t_begin > 0.01) usleep(50); else sched_yield();
      }
    }
  }
  memcpy(all, tab, sizeof(uint64_t) * (size_t)h->nranks * (size_t)h->words);
  return GCB_OK;
}

}  // extern "C"
