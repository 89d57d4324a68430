// Host side of the staging of host frames (gcb_counter_add_frames): the index group's atoms of a batch of frames
// x[frame][natoms][3] are copied into a pinned staging buffer as x'[frame][gnx][3] by a pool of host threads, so that
// only the atoms the reference's loop reads (`x[index[i]]`, g_ri3Dc.c:425-426) cross PCIe: 12 B per counted atom
// instead of 12 B per atom of the frame (36 B per counted atom for water oxygens, where every 32-byte sector of a
// frame holds part of a selected atom and neither the copy engine nor zero-copy reads can skip the hydrogens).
// Pure host code: no CUDA in this file.
#include "gcb_internal.h"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <sched.h>
#include <unistd.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace {

inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#elif defined(__aarch64__)
  asm volatile("yield");
#endif
}

// One job at a time.  A job is a number of chunks handed out through one atomic counter; the caller works alongside
// the workers 0 .. helpers-1 and returns when each of them has reported.  Workers poll the generation word for a
// short while after a job (the next batch normally follows at once), then sleep on the condition variable.
struct Job {
  void (*fn)(void *ctx, long chunk);
  void *ctx;
  long nchunks;
  std::atomic<long> next{0};
  std::atomic<int> done{0};
  void work()
  {
    for (;;) {
      const long k = next.fetch_add(1, std::memory_order_relaxed);
      if (k >= nchunks) break;
      fn(ctx, k);
    }
  }
};

class Pool {
 public:
  static Pool &get()
  {
    static Pool *p = new Pool;          // never destroyed: the detached workers may outlive static destructors
    return *p;
  }

  // begin() publishes the job to `threads - 1` workers and returns at once (the number of helpers); end() lets the
  // caller join the work, waits for the helpers and releases the pool.  Same thread for both; `job` lives in between.
  int begin(Job &job, int threads)
  {
    run_mutex_.lock();
    const int helpers = std::max(0, std::min(threads, MAX_THREADS) - 1);
    {
      std::lock_guard<std::mutex> lk(m_);
      if (pid_ != getpid()) { pid_ = getpid(); nworkers_ = 0; }      // a forked child starts without workers
      while (nworkers_ < helpers) {
        const int id = nworkers_++;
        const uint64_t seen = gen_.load(std::memory_order_relaxed);
        std::thread([this, id, seen] { worker(id, seen); }).detach();
      }
      job_ = &job;
      helpers_ = helpers;
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    return helpers;
  }

  void end(Job &job, int helpers)
  {
    job.work();
    for (long spin = 0; job.done.load(std::memory_order_acquire) < helpers; spin++) {
      if (spin < 20000) cpu_relax();
      else std::this_thread::yield();
    }
    run_mutex_.unlock();
  }

  static constexpr int MAX_THREADS = 64;

 private:
  Pool() {}

  void worker(int id, uint64_t seen)
  {
    for (;;) {
      bool fresh = false;
      for (int spin = 0; spin < 4000; spin++) {
        if (gen_.load(std::memory_order_acquire) != seen) { fresh = true; break; }
        cpu_relax();
      }
      Job *job;
      int helpers;
      {
        std::unique_lock<std::mutex> lk(m_);
        if (!fresh) cv_.wait(lk, [&] { return gen_.load(std::memory_order_relaxed) != seen; });
        // a consistent snapshot: run() publishes the job, the helper count and the generation under this lock
        seen = gen_.load(std::memory_order_relaxed);
        job = job_;
        helpers = helpers_;
      }
      // run() waits for exactly the workers below `helpers`, so `job` is alive for them; the others must not touch it
      if (id < helpers) {
        job->work();
        job->done.fetch_add(1, std::memory_order_release);
      }
    }
  }

  std::mutex run_mutex_, m_;
  std::condition_variable cv_;
  std::atomic<uint64_t> gen_{0};
  Job *job_ = nullptr;
  int helpers_ = 0;
  int nworkers_ = 0;
  pid_t pid_ = 0;
};

// `n` atoms at a fixed stride: dst[i] = s[i * stride] (3 floats each).  On x86-64 four atoms at a time, assembled in
// registers (shuffles only: the bits are not touched) and written with non-temporal stores -- the packed batch is read
// next by the DMA engine, not by this core, and a regular store would first fetch the destination line (measured:
// +15-25 % packing rate with 8 threads).  The 16-byte loads reach one float past an atom, so the last atoms of a call
// take the scalar loop (the last atom of the caller's array may end a page).
void copy_strided(const float *s, float *dst, long n, size_t stride)
{
  long i = 0;
#if defined(__x86_64__)
  for (; i < n && ((uintptr_t)dst & 15u); i++, s += stride, dst += 3) { dst[0] = s[0]; dst[1] = s[1]; dst[2] = s[2]; }
  if (stride >= 3) {
    for (; i + 4 < n; i += 4, s += 4 * stride, dst += 12) {
      const __m128 A = _mm_loadu_ps(s), B = _mm_loadu_ps(s + stride), Cc = _mm_loadu_ps(s + 2 * stride),
                   D = _mm_loadu_ps(s + 3 * stride - 1);                                  // (d[-1], d0, d1, d2)
      const __m128 t0 = _mm_shuffle_ps(A, B, _MM_SHUFFLE(0, 0, 2, 2));                    // (a2, a2, b0, b0)
      const __m128 t2 = _mm_shuffle_ps(Cc, D, _MM_SHUFFLE(1, 1, 2, 2));                   // (c2, c2, d0, d0)
      _mm_stream_ps(dst, _mm_shuffle_ps(A, t0, _MM_SHUFFLE(2, 0, 1, 0)));                 // a0 a1 a2 b0
      _mm_stream_ps(dst + 4, _mm_shuffle_ps(B, Cc, _MM_SHUFFLE(1, 0, 2, 1)));             // b1 b2 c0 c1
      _mm_stream_ps(dst + 8, _mm_shuffle_ps(t2, D, _MM_SHUFFLE(3, 2, 2, 0)));             // c2 d0 d1 d2
    }
    _mm_sfence();
  }
#endif
  for (; i < n; i++, s += stride, dst += 3) {
    uint64_t xy; uint32_t z;
    memcpy(&xy, s, 8); memcpy(&z, s + 2, 4);
    memcpy(dst, &xy, 8); memcpy(dst + 2, &z, 4);
  }
}

struct PackJob {
  const float *x;
  float *out;
  long long nsel;            // nframes * gnx
  int natoms, gnx, g0, step; // step > 0: the group is g0 + i*step
  const int *gindex;
  long long per_chunk;
};

void pack_chunk(void *ctx, long chunk)
{
  const PackJob &p = *static_cast<const PackJob *>(ctx);
  long long e = (long long)chunk * p.per_chunk;
  const long long e_end = std::min(p.nsel, e + p.per_chunk);
  while (e < e_end) {
    const long long f = e / p.gnx;
    const int i0 = (int)(e - f * p.gnx);
    const int i1 = (int)std::min<long long>(p.gnx, i0 + (e_end - e));
    const float *src = p.x + (size_t)f * p.natoms * 3;
    float *dst = p.out + (size_t)e * 3;
    if (p.step == 1) {
      memcpy(dst, src + (size_t)(p.g0 + i0) * 3, (size_t)(i1 - i0) * 12);
    } else if (p.step > 1) {
      copy_strided(src + ((size_t)p.g0 + (size_t)i0 * p.step) * 3, dst, i1 - i0, (size_t)p.step * 3);
    } else {
      for (int i = i0; i < i1; i++, dst += 3) {
        const float *s = src + (size_t)p.gindex[i] * 3;
        uint64_t xy; uint32_t z;
        memcpy(&xy, s, 8); memcpy(&z, s + 2, 4);
        memcpy(dst, &xy, 8); memcpy(dst + 2, &z, 4);
      }
    }
    e += i1 - i0;
  }
}

}  // namespace

namespace gcb {

// CPUs this process may run on
int pack_cpus()
{
  static const int n = [] {
    int cpus = 0;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) cpus = CPU_COUNT(&set);
    if (cpus <= 0) cpus = (int)sysconf(_SC_NPROCESSORS_ONLN);
    return std::max(1, cpus);
  }();
  return n;
}

// host threads a packing job may use: GCB_PACK_THREADS, else the CPUs this process may run on, at most 16
int pack_threads_default()
{
  static const int n = [] {
    if (const char *e = getenv("GCB_PACK_THREADS")) { const int v = atoi(e); if (v >= 1) return std::min(v, Pool::MAX_THREADS); }
    return std::min(pack_cpus(), 16);
  }();
  return n;
}

struct PackHandle {
  PackJob p;
  Job job;
  int helpers = 0;
};

// Starts packing on the pool's threads and returns a handle for pack_end (NULL: the job was small and is done).
void *pack_begin(const float *x, long nframes, int natoms, const int *gindex, int gnx, int g0, int step, float *out,
                 int threads)
{
  if (nframes <= 0 || gnx <= 0) return nullptr;
  PackHandle *h = new PackHandle;
  PackJob &p = h->p;
  p.x = x; p.out = out; p.nsel = (long long)nframes * gnx;
  p.natoms = natoms; p.gnx = gnx; p.g0 = g0; p.step = step; p.gindex = gindex;
  p.per_chunk = 8192;                                  // 96 KB of output per chunk
  h->job.fn = pack_chunk; h->job.ctx = &h->p;
  h->job.nchunks = (long)((p.nsel + p.per_chunk - 1) / p.per_chunk);
  if (threads <= 0) threads = pack_threads_default();
  threads = (int)std::min<long>(threads, h->job.nchunks);
  if (threads <= 1 || p.nsel < 32768) {                // not worth waking anybody
    h->job.work();
    delete h;
    return nullptr;
  }
  h->helpers = Pool::get().begin(h->job, threads);
  return h;
}

// The caller joins the packing, waits for the other threads and releases the pool.
void pack_end(void *handle)
{
  if (!handle) return;
  PackHandle *h = static_cast<PackHandle *>(handle);
  Pool::get().end(h->job, h->helpers);
  delete h;
}

void pack_frames(const float *x, long nframes, int natoms, const int *gindex, int gnx, int g0, int step, float *out,
                 int threads)
{
  pack_end(pack_begin(x, nframes, natoms, gindex, gnx, g0, step, out, threads));
}

}  // namespace gcb

extern "C" int gcb_pack_frames(const float *x, long nframes, int natoms, const int *gindex, int gnx, float *out, int threads)
{
  if (nframes < 0 || natoms <= 0 || gnx < 0 || (nframes > 0 && gnx > 0 && (!x || !out || !gindex)))
    return gcb::fail(GCB_ERR_ARG, "gcb_pack_frames: bad argument");
  for (int i = 0; i < gnx; i++)
    if (gindex[i] < 0 || gindex[i] >= natoms)
      return gcb::fail(GCB_ERR_ARG, "index group entry %d = %d is outside [0, %d)", i, gindex[i], natoms);
  // groups of the form g0 + i*step (every water oxygen, all atoms) are copied with a fixed stride
  int step = gnx >= 2 ? gindex[1] - gindex[0] : 1;
  for (int i = 2; step >= 1 && i < gnx; i++)
    if (gindex[i] - gindex[i - 1] != step) step = 0;
  if (step < 1) step = 0;
  gcb::pack_frames(x, nframes, natoms, gindex, gnx, gnx ? gindex[0] : 0, step, out, threads);
  return GCB_OK;
}
