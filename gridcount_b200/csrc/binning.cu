// The counter object behind the C ABI: staging of frames through pinned buffers on CUDA streams, the launch logic of
// K1 (which sink a batch takes: private shared-memory grids, the uint32 grid, packed L2-resident counters; rounds,
// verification, lazy folding), K2 (density normalisation, g_ri3Dc.c:451-460, and the file-order transposes,
// xdr_grid.c:113-126), the XTC pipeline and the frame-sharded reduce.  Replaces the frame loop g_ri3Dc.c:410-429 with
// gridcount() (g_ri3Dc.c:139-152) inside it.  The device code lives in k1_kernels.cuh and k1_passes.cuh.
//
// Data layout in HBM
//   frames   x[frame][atom][3] f32, exactly the host rvec layout (12 B/atom); host frames may arrive PACKED,
//            x'[frame][gnx][3] (the index group only, copied by host threads: host_pack.cpp), and are then binned
//            as frames of gnx atoms that all count
//   counts   u32 [mx][my][mz], z fastest (the reference's contiguous block)
//   acc      f32 same shape, only materialised when frame weights change
//   gindex   i32 sorted copy of the index group, rank[] = prefix count per atom
#include "gcb_internal.h"
#include "k1_kernels.cuh"
#include "k1_passes.cuh"
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <new>
#include <vector>

using namespace k1;

namespace {

// ---------------------------------------------------------------------------
// host object
// ---------------------------------------------------------------------------
struct Slot {
  cudaStream_t copy = nullptr;
  cudaEvent_t copied = nullptr, consumed = nullptr;
  float *h_x = nullptr;               // pinned staging [cap][natoms][3]
  float *d_x = nullptr;               // device staging
  unsigned long long *d_hits = nullptr, *h_hits = nullptr;
  float *d_box3 = nullptr, *h_box3 = nullptr;
  long cap = 0;                       // frames
  long cap_pack = 0;                  // frames when the batch is staged packed (index group only; 0: does not fit)
  bool dma_in_place = false;          // the last copy queued on this slot reads the caller's pinned memory
  bool copy_queued = false;           // ... and has not been seen complete yet (two-lane staging)
  size_t copy_bytes = 0;
  long pending = 0;                   // frames whose hit counts have not been collected yet
  size_t frame0 = 0;                  // global index of the first pending frame
};

struct DevHits {                      // hit counters of an add_device_frames call awaiting collection
  unsigned long long *d = nullptr;    // inside one of the counter's hit arenas
  float *d_box3 = nullptr;
  long n = 0;
  size_t frame0 = 0;
};

}  // namespace

constexpr int GCB_XTC_SETS = 6;

struct gcb_counter {
  gcb_tgrid tg{};
  GridParams P{};
  GridParamsFast PF{};
  gcb_counter_opts opts{};
  int natoms = 0, gnx = 0;
  size_t ncells = 0;
  int sm_count = 148;
  bool stream_ok = false;             // index group dense enough for the streaming kernel
  bool ap_ok = false;                 // index group is g0 + i*step: arithmetic selection inside the tile
  ApGroup ap{0, 1, TILE_ATOMS};
  // packed (W-bit lanes, L2-resident) path
  int packed_bits = 0;                // 0: not available for this grid; 8 or 4
  bool packed_auto = false;           // AUTO takes it: uint32 grid beyond L2, packed counters inside it
  long long packed_units = 0;         // atom-frames per round
  uint32_t *d_pk = nullptr;           // packed counters, always all-zero between rounds
  PackedState *d_pstate = nullptr;
  PackedState *h_pstate = nullptr;    // pinned mirror, refreshed (asynchronously) behind every call's rounds
  int packed_pause = 0;               // calls still to be binned without the packed pass (see packed_paused)
  int packed_bits_home = 0;           // the lane width chosen at create; packed_bits may fall back from 4 to 8 (packed_paused)
  long long packed_units_home = 0;
  bool packed_fallback_ok = false;    // a nibble grid may retry with byte lanes and short rounds before giving up
  int packed_mode_calls = 0, packed_cooldown = 0;
  uint64_t packed_rounds = 0, packed_redone = 0;
  // privatised path (PrivateSink): per-CTA W-bit copies of the grid, resting in d_priv between launches
  int priv_bits = 0;                  // 0: grid too large for one CTA's shared memory; 8 or 4
  unsigned priv_words = 0;            // words per private grid
  int priv_ctas = 0;                  // CTAs (private grids) of every private launch
  uint32_t *d_priv = nullptr;         // [priv_ctas][priv_words]
  unsigned long long *d_priv_hits = nullptr;   // [priv_ctas]
  PrivState *d_privstate = nullptr;
  bool priv_dirty = false;            // d_priv holds hits that are not in d_runcnt yet
  unsigned long long priv_bound = 0;  // upper bound of the hits any one private grid holds
  unsigned long long priv_limit = 0;  // flush before priv_bound could exceed this
  long long priv_min_units = 0;       // AUTO takes the private path for launches of at least this many atom-frames
  uint64_t priv_launches = 0, priv_redone = 0;
  ApGroup ap_priv{0, 1, TILE_ATOMS};
  uint2 *d_mid = nullptr;             // 4-bit packed path: byte grid between the packed rounds and the uint32 grid
  int mid_rounds = 0;                 // rounds folded into d_mid since it was last flushed (<= 16)
  bool l2_window = false;             // persisting-L2 window over d_pk set on the compute stream
  size_t l2_limit_set = 0;            // the persisting-L2 set-aside this counter asked for
  bool saturated = false;             // some cell was clamped (diag[3])
  // xtc path: GCB_XTC_SETS sets of {compressed bytes, frame table, status, decoded frames}, copy stream + events
  struct XtcSet {
    uint8_t *d_buf = nullptr; size_t buf_cap = 0;
    void *d_meta = nullptr, *h_meta = nullptr, *d_work = nullptr; int *d_status = nullptr, *h_status = nullptr; long frames_cap = 0;
    float *d_x = nullptr; size_t x_cap = 0;
    cudaStream_t stream = nullptr;                      // H2D of the compressed bytes + frame table + skim
    cudaEvent_t skimmed = nullptr, consumed = nullptr;  // skim done (set stream) / segment kernels done (compute stream)
    long nframes = 0; long first = 0;
  } xtc[GCB_XTC_SETS];
  // multi-GPU statistics scratch
  unsigned long long *d_scr = nullptr, *h_scr = nullptr;
  size_t scr_words = 0;
  cudaStream_t comm_stream = nullptr; // rank-table exchange, beside the compute stream
  cudaEvent_t reduce_done = nullptr;  // behind the D2H of the gathered statistics
  cudaEvent_t scr_up = nullptr, grp_done = nullptr;   // scratch uploaded / NCCL group finished
  bool reduce_pending = false;        // host half of the last all-reduce still to do (finish_reduce)
  size_t pend_total = 0, pend_off = 0, pend_ndev = 0, pend_nlocal = 0, pend_blk = 0;
  std::vector<size_t> pend_frames;    // frames each rank contributed to the pending reduce
  gcb_ctl *ctl = nullptr;             // shared-memory control plane between the ranks of this node (may be absent)

  std::vector<int> h_gindex;          // sorted copy of the index group
  int *d_gindex = nullptr;
  unsigned *d_rank = nullptr;         // rank[r] = number of group entries < r, r in [0, natoms]
  uint32_t *d_runcnt = nullptr, *d_total = nullptr;
  float *d_acc = nullptr;
  bool acc_used = false;              // d_acc / d_total hold something since the last reset (they stay allocated)
  unsigned long long *d_diag = nullptr;
  float *d_out = nullptr, *d_out2 = nullptr;   // density / transpose outputs

  // saturation guard (clamp_counts_kernel): upper bounds on any cell of d_runcnt / d_total
  unsigned long long bound_run = 0, bound_total = 0;
  unsigned long long clamp_at = 1ull << 27, clamp_ceiling = 0xffffffffull;
  bool have_run = false;              // d_runcnt holds hits of weight w_cur not yet folded into d_acc
  float w_cur = 0;
  bool uniform = true, any = false;
  float w_first = 0;

  // Host frames may be staged PACKED: host threads copy the index group's atoms of a batch into the pinned slot
  // (host_pack.cpp), 12 B per counted atom cross PCIe instead of 12 B per atom, and the kernels see frames of gnx
  // atoms that are all counted (`pk`: the layout fields for that view, swapped in around the launches).
  struct Layout {
    int natoms = 0;
    bool stream_ok = false, ap_ok = false;
    ApGroup ap{0, 1, TILE_ATOMS}, ap_priv{0, 1, TILE_ATOMS};
    int *d_gindex = nullptr;
    unsigned *d_rank = nullptr;
    std::vector<int> h_gindex;
    int priv_bits = 0;
  } pk;
  bool pk_ready = false;
  int host_pack = -1;                 // GCB_HOST_PACK_*: -1 auto, 0 never, 1 always
  double pack_rate = 0;               // measured: source bytes per second the host threads pack (moving average)
  // AUTO, pinned frames: do the two lanes beat plain DMA out of the caller's memory ON THIS HOST?  (They may not
  // where many ranks share the host's memory bandwidth: a packed atom costs the host 36 B read + 12 B written +
  // 12 B read by the DMA engine, an atom DMA-ed in place 36 B.)  The first calls time both, ~40 ms of calls each.
  int lanes_choice = -1;              // -1: still measuring, 0: DMA in place only, 1: two lanes
  int trial_phase = 0;                // 0: timing the two lanes, 1: timing DMA in place
  int trial_calls[2] = {0, 0};
  double trial_s[2] = {0, 0}, trial_frames[2] = {0, 0};
  uint64_t h2d_bytes = 0;             // bytes of frames queued host -> device by add_frames / submit_slot
  uint64_t frames_packed = 0;         // frames that were staged packed
  uint64_t frames_in_place = 0;       // frames DMA-ed straight out of the caller's pinned memory

  cudaStream_t compute = nullptr;
  std::vector<Slot> slots;
  int next_slot = 0;
  std::vector<DevHits> devhits;
  // per-frame hit counters of device-resident batches are carved out of a few large device buffers (no allocation
  // per call in steady state, however many calls lie between two collections); `used` goes back to zero whenever
  // no batch is waiting for collection (later launches that reuse the memory are ordered on the compute stream)
  struct HitArena { unsigned long long *d; size_t cap, used; };
  std::vector<HitArena> hit_arenas;

  std::vector<uint64_t> hits_per_frame;
  std::vector<float> weight_per_frame;
  double sumP = 0;
  size_t sumP_frames = 0;
  uint64_t hits_total = 0, edge_overflow = 0;
  int64_t launches = 0;
  void *nccl = nullptr;               // the communicator: an ncclComm_t (nccl_reduce.cpp) or an in-process rank (local_reduce.cu)
  const gcb_reduce_ops *ops = nullptr;
  int nranks = 1, rank = 0;
  std::vector<cudaEvent_t> events;
  // gcb_counter_profile: an event pair around every launch of the streaming kernel (K1's dominant kernel)
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev;
  size_t prof_used = 0;
  std::vector<int> prof_kind;         // per event pair: GCB_PROF_*
  std::vector<double> prof_units;     // per event pair: atom-frames the launch covered
  bool reduced = false;               // statistics below are global after gcb_counter_allreduce
  uint64_t global_frames = 0;
};

namespace {

inline bool same_bits(float p, float q) { uint32_t x, y; memcpy(&x, &p, 4); memcpy(&y, &q, 4); return x == y; }

unsigned rank_host(const gcb_counter *c, int r)
{
  return (unsigned)(std::lower_bound(c->h_gindex.begin(), c->h_gindex.end(), r) - c->h_gindex.begin());
}

int flush_private(gcb_counter *c);

int ensure_acc(gcb_counter *c)
{
  c->acc_used = true;
  if (c->d_acc) return GCB_OK;
  GCB_CUDA(cudaMalloc(&c->d_acc, c->ncells * sizeof(float)));
  GCB_CUDA(cudaMalloc(&c->d_total, c->ncells * sizeof(uint32_t)));
  GCB_CUDA(cudaMemsetAsync(c->d_acc, 0, c->ncells * sizeof(float), c->compute));
  GCB_CUDA(cudaMemsetAsync(c->d_total, 0, c->ncells * sizeof(uint32_t), c->compute));
  return GCB_OK;
}

int elementwise_grid(const gcb_counter *c, size_t n)
{
  const size_t b = (n + 255) / 256, cap = (size_t)c->sm_count * 8;
  return (int)std::max<size_t>(1, std::min(b, cap));
}

// make sure `units` more hits fit into every cell of `cnt` (bound = what the fullest cell may hold now)
int ensure_headroom(gcb_counter *c, uint32_t *cnt, unsigned long long &bound, unsigned long long units)
{
  if (bound + units <= c->clamp_ceiling) return GCB_OK;
  if (bound > c->clamp_at) {
    if (cnt == c->d_runcnt) { int rc = flush_private(c); if (rc) return rc; }
    clamp_counts_kernel<<<elementwise_grid(c, c->ncells), 256, 0, c->compute>>>(cnt, c->ncells, (uint32_t)c->clamp_at, c->d_diag + 3);
    c->launches++;
    GCB_CUDA(cudaGetLastError());
    bound = c->clamp_at;
  }
  if (bound + units > c->clamp_ceiling)
    return gcb::fail(GCB_ERR_RANGE, "a single launch of %llu atom-frames cannot be guarded against uint32 overflow", units);
  return GCB_OK;
}

int fold_pending(gcb_counter *c)
{
  if (!c->have_run) return GCB_OK;
  int rc = flush_private(c);
  if (rc) return rc;
  rc = ensure_acc(c);
  if (rc) return rc;
  rc = ensure_headroom(c, c->d_total, c->bound_total, c->bound_run);
  if (rc) return rc;
  c->bound_total += c->bound_run;
  c->bound_run = 0;
  fold_run_kernel<<<elementwise_grid(c, c->ncells), 256, 0, c->compute>>>(c->d_runcnt, c->d_total, c->d_acc,
                                                                           c->w_cur, c->ncells);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  c->have_run = false;
  return GCB_OK;
}

// gcb_counter_profile: an event pair on the compute stream around one launch (or a short run of launches)
int prof_begin(gcb_counter *c)
{
  if (!c->prof) return GCB_OK;
  if (c->prof_used + 2 > c->prof_ev.size()) {
    c->prof_ev.resize(c->prof_used + 2, nullptr);
    GCB_CUDA(cudaEventCreate(&c->prof_ev[c->prof_used]));
    GCB_CUDA(cudaEventCreate(&c->prof_ev[c->prof_used + 1]));
  }
  GCB_CUDA(cudaEventRecord(c->prof_ev[c->prof_used], c->compute));
  return GCB_OK;
}
int prof_end(gcb_counter *c, int kind, double units)
{
  if (!c->prof) return GCB_OK;
  GCB_CUDA(cudaEventRecord(c->prof_ev[c->prof_used + 1], c->compute));
  c->prof_used += 2;
  c->prof_kind.push_back(kind);
  c->prof_units.push_back(units);
  return GCB_OK;
}

template <class SINK>
size_t stream_smem_bytes() { return (sizeof(StreamSmem<SINK::NSTAGES, SINK::TILE_CAP>) + 15) & ~(size_t)15; }

template <class SINK, bool PBC, bool AP, int GUARD = 0>
int launch_stream(gcb_counter *c, const float *d_x, long long ntiles, const SINK &sink, int grid_cap,
                  const float *d_box3, unsigned long long *d_hits, int *grid_out, const int *run_if = nullptr,
                  const ApGroup *tiles = nullptr)
{
  static size_t attr_done[64] = {0};             // per instantiation and per device (the attribute is per context)
  const size_t smem = stream_smem_bytes<SINK>() + sink.extra_smem();
  const int dv = c->opts.device & 63;
  if (attr_done[dv] < smem) {
    GCB_CUDA(cudaFuncSetAttribute(bin_stream_kernel<SINK, PBC, AP, GUARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dv] = smem;
  }
  const ApGroup &G = tiles ? *tiles : c->ap;
  const int per_sm = std::max(1, (int)((227 * 1024) / (smem + 1024)));
  long long grid = std::min<long long>(ntiles, (long long)c->sm_count * per_sm);
  if (grid_cap > 0) grid = std::min<long long>(grid, grid_cap);
  const int kind = SINK::PRIVATE ? GCB_PROF_PRIVATE : GUARD == 2 ? GCB_PROF_PACKED : GUARD == 1 ? GCB_PROF_REDO : GCB_PROF_DIRECT;
  int rc = prof_begin(c);
  if (rc) return rc;
  bin_stream_kernel<SINK, PBC, AP, GUARD><<<(unsigned)grid, (SINK::CWARPS + 1) * 32, smem, c->compute>>>(
      d_x, c->natoms, c->d_gindex, (unsigned)c->gnx, c->d_rank, ntiles, G, c->PF, sink, d_box3, d_hits, c->d_diag, run_if);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  rc = prof_end(c, kind, (double)ntiles * (AP ? G.tile_atoms : TILE_ATOMS) / (double)c->natoms * (double)c->gnx);
  if (rc) return rc;
  if (grid_out) *grid_out = (int)grid;
  return GCB_OK;
}

// the part of a batch the streaming kernel leaves: fewer than one tile of atoms at the end
template <int MODE, bool PBC>
int launch_gather(gcb_counter *c, const float *d_x, long long e_begin, long long e_end, uint32_t *cnt, float w,
                  const float *d_box3, unsigned long long *d_hits)
{
  if (e_begin >= e_end) return GCB_OK;
  const long long chunk = (long long)GATHER_THREADS * GATHER_ITEMS;
  const long long nchunks = (e_end - e_begin + chunk - 1) / chunk;
  const long long grid = std::min<long long>(nchunks, (long long)c->sm_count * 8);
  int rc = prof_begin(c);
  if (rc) return rc;
  bin_gather_kernel<MODE, PBC><<<(unsigned)grid, GATHER_THREADS, 0, c->compute>>>(
      d_x, c->natoms, c->d_gindex, (unsigned)c->gnx, e_begin, e_end, c->P, cnt, c->d_acc, w, d_box3, d_hits, c->d_diag);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  return prof_end(c, GCB_PROF_OTHER, (double)(e_end - e_begin));
}

// Keep the packed counters L2-resident next to the coordinate stream: a persisting-L2 set-aside with an access-policy
// window over the `bytes` of them the current lane width uses, on the compute stream.  Measured (tools/l2_window.cu,
// profiles/l2_window_r02.txt): one RED per 36 streamed bytes into a 64 MB window runs at 83 G/s without it (the stream
// evicts the window, whatever evict-first hints the loads carry) and at 123 G/s with it -- the rate of a 4 MB window.
void set_packed_window(gcb_counter *c, size_t bytes)
{
  if (getenv("GCB_NO_L2_PERSIST") || bytes <= ((size_t)32 << 20)) return;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, c->opts.device) != cudaSuccess) { cudaGetLastError(); return; }
  const size_t win = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
  // The set-aside only ever grows: changing cudaLimitPersistingL2CacheSize with work in flight stalls the device for
  // tens of milliseconds (measured: ~90 ms), so a counter that once fell back to byte lanes keeps the larger set-aside
  // when it probes the nibbles again; the window itself is cheap to move.
  const size_t want = std::max(c->l2_limit_set, std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, win + ((size_t)2 << 20)));
  if (want > 0 && (c->l2_limit_set == want || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess)) {
    c->l2_limit_set = want;
    cudaStreamAttrValue av = {};
    av.accessPolicyWindow.base_ptr = c->d_pk;
    av.accessPolicyWindow.num_bytes = win;
    av.accessPolicyWindow.hitRatio = std::min(1.0f, (float)((double)want / (double)win));
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(c->compute, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess) c->l2_window = true;
  }
  cudaGetLastError();
}

// Packed path: rounds of ~packed_units atom-frames; per round the streaming kernel into the packed counters
// (returns at once while backing off), verify + apply, the guarded uint32 kernel (returns at once unless the round
// failed or is skipped), and the gather kernel for the sub-tile tail (straight into the uint32 grid).
template <int W, bool PBC>
int launch_packed(gcb_counter *c, const float *d_x, long nframes, const float *d_box3, unsigned long long *d_hits)
{
  const bool ap = c->ap_ok && c->opts.kernel != GCB_KERNEL_STREAM_INDEXED;
  const int tile_atoms = ap ? c->ap.tile_atoms : TILE_ATOMS;
  // frames per round: as many as packed_units allows, the call cut into equal rounds, a multiple of 4 frames
  // (every round starts 16-byte aligned)
  long fpr = (long)std::max<long long>(4, (c->packed_units / std::max(c->gnx, 1)) & ~3LL);
  {
    const long nrounds = (nframes + fpr - 1) / fpr;
    fpr = std::max<long>(4, (((nframes + nrounds - 1) / nrounds) + 3) & ~3L);
  }
  const size_t nwords = (c->ncells + (32 / W) - 1) / (32 / W), nquads = (nwords + 3) / 4;
  if (!c->d_pk) {
    // (a nibble grid that may fall back to byte lanes gets room for the bytes; the window below covers what fits)
    const size_t pk_bytes = c->packed_fallback_ok ? ((c->ncells + 15) / 16) * 16 : nquads * 16;
    GCB_CUDA(cudaMalloc(&c->d_pk, pk_bytes));
    GCB_CUDA(cudaMemsetAsync(c->d_pk, 0, pk_bytes, c->compute));
    GCB_CUDA(cudaMalloc(&c->d_pstate, sizeof(PackedState)));
    GCB_CUDA(cudaMemsetAsync(c->d_pstate, 0, sizeof(PackedState), c->compute));
    if (W == 4 && !getenv("GCB_NO_MID")) {
      GCB_CUDA(cudaMalloc(&c->d_mid, nwords * sizeof(uint2)));
      GCB_CUDA(cudaMemsetAsync(c->d_mid, 0, nwords * sizeof(uint2), c->compute));
    }
    GCB_CUDA(cudaHostAlloc((void **)&c->h_pstate, sizeof(PackedState), cudaHostAllocDefault));
    memset(c->h_pstate, 0, sizeof(PackedState));
    set_packed_window(c, nquads * 16);
  }
  PackedState *st = c->d_pstate;
  const int aux_grid = (int)std::max<size_t>(1, std::min<size_t>((nquads + 255) / 256, (size_t)c->sm_count * 8));
  const int apply_grid = (int)std::max<size_t>(1, std::min<size_t>((nwords + 255) / 256, (size_t)c->sm_count * 16));
  for (long f0 = 0; f0 < nframes; f0 += fpr) {
    const long n = std::min(fpr, nframes - f0);
    const float *xs = d_x + (size_t)f0 * c->natoms * 3;
    const float *bs = d_box3 ? d_box3 + 3 * f0 : nullptr;
    const long long total_e = (long long)n * c->gnx, total_atoms = (long long)n * c->natoms;
    const long long ntiles = total_atoms / tile_atoms;
    long long e_begin = 0;
    if (ntiles > 0) {
      packed_begin_kernel<<<1, 1, 0, c->compute>>>(st, c->d_diag);
      PackedSink<W> sink;
      sink.pk = c->d_pk;
      int rc = ap ? launch_stream<PackedSink<W>, PBC, true, 2>(c, xs, ntiles, sink, 0, bs, d_hits + f0, nullptr, &st->alarm)
                  : launch_stream<PackedSink<W>, PBC, false, 2>(c, xs, ntiles, sink, 0, bs, d_hits + f0, nullptr, &st->alarm);
      if (rc) return rc;
      rc = prof_begin(c);
      if (rc) return rc;
      packed_verify_kernel<W><<<aux_grid, 256, 0, c->compute>>>(reinterpret_cast<const uint4 *>(c->d_pk), nquads, d_hits + f0, n, st);
      if (W == 4 && c->d_mid) {
        if (c->mid_rounds >= 16) { rc = flush_private(c); if (rc) return rc; }
        packed_apply_mid_kernel<<<apply_grid, 256, 0, c->compute>>>(c->d_pk, nwords, c->d_mid, st, d_hits + f0, n, c->d_diag);
        c->mid_rounds++;
      } else {
        packed_apply_kernel<W><<<apply_grid, 256, 0, c->compute>>>(c->d_pk, nwords, c->d_runcnt, c->ncells, st, d_hits + f0, n, c->d_diag);
      }
      c->launches += 3;
      GCB_CUDA(cudaGetLastError());
      rc = prof_end(c, GCB_PROF_OTHER, 0.0);
      if (rc) return rc;
      DirectSink<MODE_COUNT> redo;
      redo.counts = c->d_runcnt; redo.acc = c->d_acc; redo.w = 1.0f;
      rc = ap ? launch_stream<DirectSink<MODE_COUNT>, PBC, true, 1>(c, xs, ntiles, redo, 0, bs, d_hits + f0, nullptr, &st->alarm)
              : launch_stream<DirectSink<MODE_COUNT>, PBC, false, 1>(c, xs, ntiles, redo, 0, bs, d_hits + f0, nullptr, &st->alarm);
      if (rc) return rc;
      const long long g1 = ntiles * tile_atoms, f1 = g1 / c->natoms;
      e_begin = f1 * c->gnx + rank_host(c, (int)(g1 - f1 * c->natoms));
    }
    int rc = launch_gather<MODE_COUNT, PBC>(c, xs, e_begin, total_e, c->d_runcnt, 1.0f, bs, d_hits + f0);
    if (rc) return rc;
  }
  // the host learns how the rounds went without waiting for them: the state is copied to a pinned mirror behind the
  // call's kernels and read (possibly one call late) by packed_paused() before the next call is queued
  GCB_CUDA(cudaMemcpyAsync(c->h_pstate, c->d_pstate, sizeof(PackedState), cudaMemcpyDeviceToHost, c->compute));
  return GCB_OK;
}

// What to do with input that keeps overflowing the packed lanes (the device-side back-off has reached 8 rounds; the
// host learns of it from the pinned mirror, possibly a call late):
//  * a grid on NIBBLE lanes (its bytes would not fit the L2: 500^3) first falls back to BYTE lanes with short rounds
//    (0.64 hits per cell): if atoms pile up on a small part of the grid, the bytes that are actually touched are
//    L2-resident by themselves -- measured on the clustered 500^3 shape: 83.5 G/s against 57 for the direct kernel
//    and 45 when every nibble round is thrown away; every 256 calls the nibbles are probed again;
//  * byte lanes that overflow as well (everything on a handful of cells) leave the packed passes alone for 32 calls:
//    the direct kernel's hot cells stay L2-resident by themselves.
// -> true: bin this call with the direct kernel.
bool packed_paused(gcb_counter *c)
{
  if (c->packed_pause > 0) {
    if (--c->packed_pause == 0 && c->packed_bits != c->packed_bits_home) {      // after the pause: the home lanes again
      c->packed_bits = c->packed_bits_home; c->packed_units = c->packed_units_home; c->packed_mode_calls = 0;
      set_packed_window(c, (c->ncells + 1) / 2);
    }
    return true;
  }
  if (c->packed_cooldown > 0) { c->packed_cooldown--; return false; }          // (a mirror copy from before the switch may still land)
  const bool failing = c->h_pstate && ((volatile PackedState *)c->h_pstate)->backoff >= 8;
  auto restart = [&] {                                                         // a new mode starts with a clean back-off
    c->h_pstate->backoff = 0;
    cudaMemsetAsync(c->d_pstate, 0, offsetof(PackedState, rounds), c->compute);
    c->packed_cooldown = 2;
    c->packed_mode_calls = 0;
  };
  if (failing) {
    if (c->packed_fallback_ok && c->packed_bits == 4) {
      c->packed_bits = 8;
      c->packed_units = std::max<long long>(4096, (long long)((double)c->ncells * 0.64));
      set_packed_window(c, c->ncells);
      restart();
      return false;
    }
    restart();
    c->packed_pause = 32;
    return true;
  }
  if (c->packed_bits != c->packed_bits_home && ++c->packed_mode_calls >= 256) {  // probe the home lanes again
    c->packed_bits = c->packed_bits_home; c->packed_units = c->packed_units_home;
    set_packed_window(c, (c->ncells + 1) / 2);
    restart();
  }
  return false;
}

// The private grids hold hits that the uint32 grid does not have yet: fold them in.  Called by everything that
// reads or rescales d_runcnt (density, counts, fold, clamp, all-reduce) and before a lane could fill up.
int flush_private(gcb_counter *c)
{
  if (c->mid_rounds > 0) {                       // the byte grid of the 4-bit packed rounds
    const size_t nwords = (c->ncells + 7) / 8;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((nwords + 255) / 256, (size_t)c->sm_count * 16));
    mid_flush_kernel<<<grid, 256, 0, c->compute>>>(c->d_mid, nwords, c->d_runcnt, c->ncells);
    c->launches++;
    GCB_CUDA(cudaGetLastError());
    c->mid_rounds = 0;
  }
  if (!c->priv_dirty) return GCB_OK;
  const int grid = (int)std::max<unsigned>(1u, std::min<unsigned>((c->priv_words + 255u) / 256u, (unsigned)c->sm_count * 8u));
  if (c->priv_bits == 8)
    private_flush_kernel<8><<<grid, 256, 0, c->compute>>>(c->d_priv, c->d_priv_hits, c->priv_words, c->priv_ctas, c->d_runcnt, c->ncells);
  else
    private_flush_kernel<4><<<grid, 256, 0, c->compute>>>(c->d_priv, c->d_priv_hits, c->priv_words, c->priv_ctas, c->d_runcnt, c->ncells);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  c->priv_dirty = false;
  c->priv_bound = 0;
  return GCB_OK;
}

// Private path: the streaming kernel with one CTA per SM, each binning into its own shared-memory copy of the grid
// (PrivateSink); the sub-tile tail goes to the gather kernel and straight into the uint32 grid.
template <int W, bool PBC>
int launch_private(gcb_counter *c, const float *d_x, long nframes, const float *d_box3, unsigned long long *d_hits)
{
  const int tile_atoms = c->ap_priv.tile_atoms;
  if (!c->d_priv) {
    GCB_CUDA(cudaMalloc(&c->d_priv, (size_t)c->priv_ctas * c->priv_words * 4));
    GCB_CUDA(cudaMalloc(&c->d_priv_hits, sizeof(unsigned long long) * (size_t)c->priv_ctas));
    GCB_CUDA(cudaMalloc(&c->d_privstate, sizeof(PrivState)));
    GCB_CUDA(cudaMemsetAsync(c->d_priv, 0, (size_t)c->priv_ctas * c->priv_words * 4, c->compute));
    GCB_CUDA(cudaMemsetAsync(c->d_priv_hits, 0, sizeof(unsigned long long) * (size_t)c->priv_ctas, c->compute));
    GCB_CUDA(cudaMemsetAsync(c->d_privstate, 0, sizeof(PrivState), c->compute));
  }
  // what one private grid may receive from a tile: its selected atoms
  const unsigned long long per_tile = (unsigned long long)(tile_atoms / std::max(c->ap.step, 1)) + 2ull;
  // frames per launch: no private grid may be handed more than the flush threshold allows (a multiple of 4 frames:
  // every launch starts 16-byte aligned)
  const unsigned long long tiles_cap = std::max<unsigned long long>(1, c->priv_limit / per_tile) * (unsigned long long)c->priv_ctas;
  const long fpl = (long)std::max<unsigned long long>(4, (tiles_cap * (unsigned long long)tile_atoms / (unsigned long long)c->natoms) & ~3ull);
  for (long f0 = 0; f0 < nframes; f0 += fpl) {
    const long n = std::min(fpl, nframes - f0);
    const float *xs = d_x + (size_t)f0 * c->natoms * 3;
    const float *bs = d_box3 ? d_box3 + 3 * f0 : nullptr;
    const long long total_e = (long long)n * c->gnx, total_atoms = (long long)n * c->natoms;
    const long long ntiles = total_atoms / tile_atoms;
    long long e_begin = 0;
    if (ntiles > 0) {
      const unsigned long long share = ((unsigned long long)ntiles + (unsigned long long)c->priv_ctas - 1ull) / (unsigned long long)c->priv_ctas * per_tile;
      if (c->priv_bound + share > c->priv_limit) { int rc = flush_private(c); if (rc) return rc; }
      private_begin_kernel<<<1, 1, 0, c->compute>>>(c->d_privstate);
      c->launches++;
      GCB_CUDA(cudaGetLastError());
      PrivateSink<W> sink;
      sink.counts = c->d_runcnt; sink.stage = c->d_priv; sink.stage_hits = c->d_priv_hits; sink.st = c->d_privstate;
      sink.nwords = c->priv_words;
      int rc = launch_stream<PrivateSink<W>, PBC, true>(c, xs, ntiles, sink, c->priv_ctas, bs, d_hits + f0, nullptr, nullptr, &c->ap_priv);
      if (rc) return rc;
      c->priv_dirty = true;
      c->priv_bound += share;
      const long long g1 = ntiles * tile_atoms, f1 = g1 / c->natoms;
      e_begin = f1 * c->gnx + rank_host(c, (int)(g1 - f1 * c->natoms));
    }
    int rc = launch_gather<MODE_COUNT, PBC>(c, xs, e_begin, total_e, c->d_runcnt, 1.0f, bs, d_hits + f0);
    if (rc) return rc;
  }
  return GCB_OK;
}

template <int MODE, bool PBC>
int launch_bin(gcb_counter *c, const float *d_x, long nframes, float w, const float *d_box3,
               unsigned long long *d_hits)
{
  const long long total_e = (long long)nframes * c->gnx;
  if (total_e == 0) return GCB_OK;
  uint32_t *cnt = MODE == MODE_COUNT ? c->d_runcnt : c->d_total;
  long long e_begin = 0;
  const bool aligned = ((uintptr_t)d_x & 15u) == 0;
  const int k = c->opts.kernel;
  const bool use_stream = aligned && k != GCB_KERNEL_GATHER &&
                          (c->stream_ok || k == GCB_KERNEL_STREAM || k == GCB_KERNEL_STREAM_INDEXED || k == GCB_KERNEL_PACKED || k == GCB_KERNEL_PRIVATE);
  // AUTO: grids that fit one CTA's shared memory as 6- or 8-bit counters are binned into per-CTA private copies
  // (PrivateSink) when the launch is large enough to pay for loading and storing them
  if (use_stream && MODE == MODE_COUNT && c->priv_bits != 0 && c->ap_ok &&
      (k == GCB_KERNEL_PRIVATE || (k == GCB_KERNEL_AUTO && total_e >= c->priv_min_units)))
    return c->priv_bits == 8 ? launch_private<8, PBC>(c, d_x, nframes, d_box3, d_hits)
                             : launch_private<4, PBC>(c, d_x, nframes, d_box3, d_hits);
  // AUTO: grids beyond L2 whose packed counters fit it take the packed path (PackedSink)
  if (use_stream && MODE == MODE_COUNT && c->packed_bits != 0 &&
      (k == GCB_KERNEL_PACKED || (k == GCB_KERNEL_AUTO && c->packed_auto && total_e >= (1LL << 20) && !packed_paused(c))))
    return c->packed_bits == 8 ? launch_packed<8, PBC>(c, d_x, nframes, d_box3, d_hits)
                               : launch_packed<4, PBC>(c, d_x, nframes, d_box3, d_hits);
  if (use_stream) {
    const bool ap = c->ap_ok && k != GCB_KERNEL_STREAM_INDEXED;
    const int tile_atoms = ap ? c->ap.tile_atoms : TILE_ATOMS;
    const long long total_atoms = (long long)nframes * c->natoms;
    const long long ntiles = total_atoms / tile_atoms;      // full tiles; the remainder goes to the gather kernel
    if (ntiles > 0) {
      DirectSink<MODE> sink;
      sink.counts = cnt; sink.acc = c->d_acc; sink.w = w;
      int rc = ap ? launch_stream<DirectSink<MODE>, PBC, true>(c, d_x, ntiles, sink, 0, d_box3, d_hits, nullptr)
                  : launch_stream<DirectSink<MODE>, PBC, false>(c, d_x, ntiles, sink, 0, d_box3, d_hits, nullptr);
      if (rc) return rc;
      const long long g1 = ntiles * tile_atoms, f1 = g1 / c->natoms;
      e_begin = f1 * c->gnx + rank_host(c, (int)(g1 - f1 * c->natoms));
    }
  }
  return launch_gather<MODE, PBC>(c, d_x, e_begin, total_e, cnt, w, d_box3, d_hits);
}

template <int MODE>
int launch_bin_pbc(gcb_counter *c, const float *d_x, long nframes, float w, const float *d_box3,
                   unsigned long long *d_hits)
{
  if (c->opts.pbc == GCB_PBC_RECT) return launch_bin<MODE, true>(c, d_x, nframes, w, d_box3, d_hits);
  return launch_bin<MODE, false>(c, d_x, nframes, w, d_box3, d_hits);
}

// Bin `nframes` device-resident frames whose per-frame weights are `w[0..nframes)`
// (host array, NULL = 1).  Frames are cut into runs of bit-identical weight.
// Long runs (and the common uniform case) count integers and are folded into
// the f32 accumulator when the weight changes; short runs add floats directly
// (all additions inside one run carry the same value, so they commute, and
// runs are ordered on the compute stream): both reproduce the reference's
// sequential `grid += weight` exactly.
int bin_device_batch(gcb_counter *c, const float *d_x, long nframes, const float *w, const float *d_box3,
                     unsigned long long *d_hits)
{
  long f = 0;
  while (f < nframes) {
    const float wr = w ? w[f] : 1.0f;
    long e = f + 1;
    // one launch never adds more than a quarter of the counter range to a cell (saturation guard)
    const long max_frames = (long)std::max<unsigned long long>(1, (c->clamp_ceiling / 4) / (unsigned long long)std::max(c->gnx, 1));
    while (e < nframes && e - f < max_frames && same_bits(w ? w[e] : 1.0f, wr)) e++;
    const long n = e - f;
    const unsigned long long units = (unsigned long long)n * (unsigned long long)c->gnx;
    if (!c->any) { c->any = true; c->w_first = wr; }
    else if (!same_bits(wr, c->w_first)) c->uniform = false;

    const float *xs = d_x + (size_t)f * c->natoms * 3;
    const float *bs = d_box3 ? d_box3 + 3 * f : nullptr;
    const bool continues = c->have_run && same_bits(wr, c->w_cur);
    const bool fresh = !c->have_run && !c->acc_used;
    const bool long_run = (double)n * c->gnx >= (double)c->ncells / 8.0;
    int rc;
    if (continues || fresh || long_run) {
      if (c->have_run && !continues) { rc = fold_pending(c); if (rc) return rc; }
      c->have_run = true;
      c->w_cur = wr;
      rc = ensure_headroom(c, c->d_runcnt, c->bound_run, units);
      if (rc) return rc;
      c->bound_run += units;
      rc = launch_bin_pbc<MODE_COUNT>(c, xs, n, wr, bs, d_hits + f);
    } else {
      rc = fold_pending(c);
      if (rc) return rc;
      rc = ensure_acc(c);
      if (rc) return rc;
      rc = ensure_headroom(c, c->d_total, c->bound_total, units);
      if (rc) return rc;
      c->bound_total += units;
      rc = launch_bin_pbc<MODE_FLOAT>(c, xs, n, wr, bs, d_hits + f);
    }
    if (rc) return rc;
    f = e;
  }
  return GCB_OK;
}

// collect the hit counters of a slot whose kernels have finished
int collect_slot(gcb_counter *c, Slot &s)
{
  if (s.pending == 0) return GCB_OK;
  GCB_CUDA(cudaEventSynchronize(s.consumed));
  for (long f = 0; f < s.pending; f++) c->hits_per_frame[s.frame0 + f] = s.h_hits[f];
  s.pending = 0;
  return GCB_OK;
}

// `want` > 0: at least that many slots (the two-lane staging of pinned frames grows the default 3 to 6)
int alloc_slots(gcb_counter *c, int want = 0)
{
  const size_t have = c->slots.size();
  const int ns = std::max<int>(want, have ? (int)have : (c->opts.nslots > 0 ? c->opts.nslots : 3));
  if ((size_t)ns <= have) return GCB_OK;
  const size_t frame_bytes = (size_t)c->natoms * 12;
  long fpb = c->opts.frames_per_batch;
  if (fpb <= 0) fpb = (long)std::max<size_t>(1, ((size_t)32 << 20) / std::max<size_t>(frame_bytes, 1));
  c->slots.resize((size_t)ns);
  // a packed batch holds as many frames as fit the same bytes (at most 16 x: per-frame counters are sized by it;
  // fewer than `fpb`, possibly none, for an index group with more entries than the frame has atoms)
  const long fpb_pack = c->gnx > 0 ? (long)std::min<long long>((long long)fpb * 16, (long long)fpb * c->natoms / c->gnx) : 0;
  const long fpb_max = std::max(fpb, fpb_pack);
  for (size_t k = have; k < c->slots.size(); k++) {
    Slot &s = c->slots[k];
    s.cap = fpb;
    s.cap_pack = fpb_pack;
    GCB_CUDA(cudaStreamCreateWithFlags(&s.copy, cudaStreamNonBlocking));
    GCB_CUDA(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
    GCB_CUDA(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
    GCB_CUDA(cudaHostAlloc(&s.h_x, frame_bytes * fpb, cudaHostAllocDefault));
    GCB_CUDA(cudaMalloc(&s.d_x, frame_bytes * fpb + 16));
    GCB_CUDA(cudaMalloc(&s.d_hits, sizeof(unsigned long long) * (fpb_max + 2)));
    GCB_CUDA(cudaHostAlloc(&s.h_hits, sizeof(unsigned long long) * (fpb_max + 2), cudaHostAllocDefault));
    GCB_CUDA(cudaMalloc(&s.d_box3, sizeof(float) * 3 * fpb_max));
    GCB_CUDA(cudaHostAlloc(&s.h_box3, sizeof(float) * 3 * fpb_max, cudaHostAllocDefault));
  }
  return GCB_OK;
}

// The layout the kernels see while a packed batch is binned: frames of gnx atoms, every one of them counted.
int ensure_pack_layout(gcb_counter *c)
{
  if (c->pk_ready) return GCB_OK;
  auto &L = c->pk;
  const int n = c->gnx;
  L.natoms = n;
  L.stream_ok = true;
  L.ap_ok = n >= 256;                           // (as in gcb_counter_create; smaller frames take the indexed kernels)
  L.ap = ApGroup{0, 1, std::min(TILE_ATOMS, n & ~3)};
  L.ap_priv = ApGroup{0, 1, std::min(PrivateSink<8>::TILE_CAP, n & ~3)};
  L.priv_bits = n >= 256 ? c->priv_bits : 0;
  L.h_gindex.resize((size_t)n);
  for (int i = 0; i < n; i++) L.h_gindex[(size_t)i] = i;
  std::vector<unsigned> rank((size_t)n + 1);
  for (int r = 0; r <= n; r++) rank[(size_t)r] = (unsigned)r;
  GCB_CUDA(cudaMalloc(&L.d_gindex, sizeof(int) * std::max(n, 1)));
  GCB_CUDA(cudaMalloc(&L.d_rank, sizeof(unsigned) * rank.size()));
  if (n) GCB_CUDA(cudaMemcpy(L.d_gindex, L.h_gindex.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
  GCB_CUDA(cudaMemcpy(L.d_rank, rank.data(), sizeof(unsigned) * rank.size(), cudaMemcpyHostToDevice));
  c->pk_ready = true;
  return GCB_OK;
}

void swap_layout(gcb_counter *c)
{
  auto &L = c->pk;
  std::swap(c->natoms, L.natoms);
  std::swap(c->stream_ok, L.stream_ok);
  std::swap(c->ap_ok, L.ap_ok);
  std::swap(c->ap, L.ap);
  std::swap(c->ap_priv, L.ap_priv);
  std::swap(c->d_gindex, L.d_gindex);
  std::swap(c->d_rank, L.d_rank);
  std::swap(c->priv_bits, L.priv_bits);
  c->h_gindex.swap(L.h_gindex);
}

// host threads one packing job of this counter uses: ranks of a multi-GPU job share the node's cores
int pack_threads(const gcb_counter *c)
{
  const int t = gcb::pack_threads_default();
  if (c->nranks <= 1 || getenv("GCB_PACK_THREADS")) return t;
  return std::min(t, std::max(2, gcb::pack_cpus() / c->nranks));
}

// stage `n` frames that already sit in slot.h_x (or are copied there from `src`; `pack`: slot.h_x holds only the
// index group's atoms, x'[n][gnx][3], and the kernels run on the packed layout)
int submit_slot(gcb_counter *c, Slot &s, const float *src, bool src_pinned, long n, const float *w,
                const float *box, bool pack = false)
{
  size_t bytes = (size_t)n * c->natoms * 12;
  const float *from = s.h_x;
  if (pack) {
    int rc = ensure_pack_layout(c);
    if (rc) return rc;
    bytes = (size_t)n * c->gnx * 12;
    c->frames_packed += (uint64_t)n;
  } else if (src) {
    if (src_pinned) { from = src; c->frames_in_place += (uint64_t)n; }   // DMA straight out of the caller's pinned memory
    else memcpy(s.h_x, src, bytes);
  }
  s.dma_in_place = from != s.h_x;
  s.copy_bytes = bytes;
  s.copy_queued = true;
  c->h2d_bytes += bytes;
  GCB_CUDA(cudaMemcpyAsync(s.d_x, from, bytes, cudaMemcpyHostToDevice, s.copy));
  if (c->opts.pbc == GCB_PBC_RECT) {
    if (!box) return gcb::fail(GCB_ERR_ARG, "GCB_PBC_RECT needs the per-frame box");
    for (long f = 0; f < n; f++) { s.h_box3[3 * f] = box[9 * f]; s.h_box3[3 * f + 1] = box[9 * f + 4]; s.h_box3[3 * f + 2] = box[9 * f + 8]; }
    GCB_CUDA(cudaMemcpyAsync(s.d_box3, s.h_box3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, s.copy));
  }
  GCB_CUDA(cudaEventRecord(s.copied, s.copy));
  GCB_CUDA(cudaStreamWaitEvent(c->compute, s.copied, 0));
  GCB_CUDA(cudaMemsetAsync(s.d_hits, 0, sizeof(unsigned long long) * (n + 2), c->compute));
  if (pack) swap_layout(c);
  int rc = bin_device_batch(c, s.d_x, n, w, c->opts.pbc == GCB_PBC_RECT ? s.d_box3 : nullptr, s.d_hits);
  if (pack) swap_layout(c);
  if (rc) return rc;
  GCB_CUDA(cudaMemcpyAsync(s.h_hits, s.d_hits, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, c->compute));
  GCB_CUDA(cudaEventRecord(s.consumed, c->compute));
  // the next copy into this slot must wait for the kernels that read it
  GCB_CUDA(cudaStreamWaitEvent(s.copy, s.consumed, 0));
  s.pending = n;
  s.frame0 = c->hits_per_frame.size();        // frames keep their submission order whatever slot they used
  c->hits_per_frame.resize(s.frame0 + (size_t)n, 0);
  for (long f = 0; f < n; f++) c->weight_per_frame.push_back(w ? w[f] : 1.0f);
  return GCB_OK;
}

void free_counter(gcb_counter *c)
{
  if (!c) return;
  cudaSetDevice(c->opts.device);
  if (c->compute) cudaStreamSynchronize(c->compute);
  if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
  if (c->nccl && c->ops && c->ops->destroy) c->ops->destroy(c->nccl);
  if (c->ctl) gcb_ctl_close(c->ctl);
  if (c->l2_window) {                  // give the set-aside lines back
    cudaStreamAttrValue av = {};
    cudaStreamSetAttribute(c->compute, cudaStreamAttributeAccessPolicyWindow, &av);
    cudaCtxResetPersistingL2Cache();
  }
  for (auto &s : c->slots) {
    if (s.copy) { cudaStreamSynchronize(s.copy); cudaStreamDestroy(s.copy); }
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.consumed) cudaEventDestroy(s.consumed);
    cudaFreeHost(s.h_x); cudaFree(s.d_x); cudaFree(s.d_hits); cudaFreeHost(s.h_hits);
    cudaFree(s.d_box3); cudaFreeHost(s.h_box3);
  }
  for (auto &d : c->devhits) cudaFree(d.d_box3);           // (d.d lives in a hit arena)
  for (auto &a : c->hit_arenas) cudaFree(a.d);
  cudaFree(c->pk.d_gindex); cudaFree(c->pk.d_rank);
  cudaFree(c->d_gindex); cudaFree(c->d_rank); cudaFree(c->d_runcnt); cudaFree(c->d_total); cudaFree(c->d_acc);
  cudaFree(c->d_diag); cudaFree(c->d_out); cudaFree(c->d_out2);
  cudaFree(c->d_pk); cudaFree(c->d_pstate); cudaFreeHost(c->h_pstate); cudaFree(c->d_mid);
  cudaFree(c->d_priv); cudaFree(c->d_priv_hits); cudaFree(c->d_privstate);
  for (auto &x : c->xtc) {
    cudaFree(x.d_buf); cudaFree(x.d_meta); cudaFree(x.d_status); cudaFree(x.d_x); cudaFree(x.d_work); cudaFreeHost(x.h_meta); cudaFreeHost(x.h_status);
    if (x.skimmed) cudaEventDestroy(x.skimmed);
    if (x.consumed) cudaEventDestroy(x.consumed);
    if (x.stream) { cudaStreamSynchronize(x.stream); cudaStreamDestroy(x.stream); }
  }
  cudaFree(c->d_scr); cudaFreeHost(c->h_scr);
  for (auto e : c->events) if (e) cudaEventDestroy(e);
  for (auto e : c->prof_ev) if (e) cudaEventDestroy(e);
  if (c->compute) cudaStreamDestroy(c->compute);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->reduce_done) cudaEventDestroy(c->reduce_done);
  if (c->scr_up) cudaEventDestroy(c->scr_up);
  if (c->grp_done) cudaEventDestroy(c->grp_done);
  cudaGetLastError();                  // nothing above may leave an error behind for the next counter's launch checks
  delete c;
}

}  // namespace

extern "C" {

int gcb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int gcb_counter_create(gcb_counter **out, const gcb_tgrid *tg, const int *gindex, int gnx, int natoms,
                       const gcb_counter_opts *opts)
{
  if (!out || !tg || natoms <= 0 || gnx < 0 || (gnx > 0 && !gindex))
    return gcb::fail(GCB_ERR_ARG, "gcb_counter_create: bad argument");
  if (tg->mx[0] <= 0 || tg->mx[1] <= 0 || tg->mx[2] <= 0 || gcb::cells(*tg) > (size_t)INT32_MAX)
    return gcb::fail(GCB_ERR_ARG, "gcb_counter_create: grid %d x %d x %d not supported", tg->mx[0], tg->mx[1], tg->mx[2]);
  for (int i = 0; i < gnx; i++)
    if (gindex[i] < 0 || gindex[i] >= natoms)
      return gcb::fail(GCB_ERR_ARG, "index group entry %d = %d is outside [0, %d)", i, gindex[i], natoms);
  if (gcb_device_count() <= 0) return gcb::fail(GCB_ERR_CUDA, "no CUDA device: libgridcount_b200 has no CPU fallback");

  gcb_counter *c = new (std::nothrow) gcb_counter;
  if (!c) return gcb::fail(GCB_ERR_NOMEM, "out of memory");
  if (opts) c->opts = *opts;
  c->tg = *tg;
  c->natoms = natoms;
  c->gnx = gnx;
  c->ncells = gcb::cells(*tg);
  for (int d = 0; d < 3; d++) {
    c->P.a[d] = tg->a[d]; c->P.b[d] = tg->b[d]; c->P.delta[d] = tg->delta[d]; c->P.fmx[d] = (float)tg->mx[d];
  }
  c->P.my = tg->mx[1]; c->P.mz = tg->mx[2];
  {
    const bool fast = !gcb_edge_vulnerable(tg) && tg->mx[0] < (1 << 21) && tg->mx[1] < (1 << 21) && tg->mx[2] < (1 << 21);
    for (int d = 0; d < 3; d++) {
      c->PF.a[d] = tg->a[d]; c->PF.b[d] = tg->b[d]; c->PF.delta[d] = tg->delta[d]; c->PF.fmx[d] = (float)tg->mx[d];
      const double inv = (double)(1.0f / tg->delta[d]);
      c->PF.inv_lo[d] = fast ? (float)(inv * (1.0 - 4.76837158203125e-07)) : NAN;       // 2^-21
      c->PF.inv_hi[d] = fast ? (float)(inv * (1.0 + 4.76837158203125e-07)) : NAN;
    }
    c->PF.my = tg->mx[1]; c->PF.mz = tg->mx[2];
    c->PF.bias = (int)(0x4B400000u * ((uint32_t)tg->mx[1] * (uint32_t)tg->mx[2] + (uint32_t)tg->mx[2] + 1u));
  }

#define GCB_CREATE_CUDA(call)                                                                        \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { gcb::set_error("%s failed: %s", #call, cudaGetErrorString(e_)); \
       free_counter(c); return GCB_ERR_CUDA; } } while (0)

  GCB_CREATE_CUDA(cudaSetDevice(c->opts.device));
  cudaDeviceProp prop;
  GCB_CREATE_CUDA(cudaGetDeviceProperties(&prop, c->opts.device));
  c->sm_count = prop.multiProcessorCount;
  GCB_CREATE_CUDA(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));

  // integer counts are order free, so the group may be sorted (duplicates kept)
  c->h_gindex.assign(gindex, gindex + gnx);
  std::sort(c->h_gindex.begin(), c->h_gindex.end());
  std::vector<unsigned> rank((size_t)natoms + 1);
  {
    size_t p = 0;
    for (int r = 0; r <= natoms; r++) {
      while (p < c->h_gindex.size() && c->h_gindex[p] < r) p++;
      rank[r] = (unsigned)p;
    }
  }
  // streaming reads every byte of every frame: worth it while >= 1/8 of the 32-byte sectors are wanted anyway
  c->stream_ok = (double)gnx * 8.0 >= (double)natoms && (long long)natoms * 3 >= 1;
  {
    // arithmetic progression?  (duplicates or irregular groups take the indexed path)
    const int step = gnx >= 2 ? c->h_gindex[1] - c->h_gindex[0] : 1;
    bool ap = gnx >= 1 && step >= 1 && natoms >= 256;
    for (int i = 2; ap && i < gnx; i++) ap = c->h_gindex[i] - c->h_gindex[i - 1] == step;
    c->ap_ok = ap;
    if (ap) {
      c->ap.g0 = c->h_gindex[0];
      c->ap.step = step;
      c->ap.tile_atoms = std::min(TILE_ATOMS, natoms & ~3);     // <= natoms: a tile spans at most two frames
    }
  }

  {
    if (const char *e = getenv("GCB_HOST_PACK")) c->host_pack = e[0] == '0' ? 0 : e[0] == '1' ? 1 : -1;
    if (const char *e = getenv("GCB_CLAMP_AT")) c->clamp_at = strtoull(e, nullptr, 10);
    if (const char *e = getenv("GCB_CLAMP_CEILING")) c->clamp_ceiling = strtoull(e, nullptr, 10);
    if (c->clamp_at < 1 || c->clamp_ceiling > 0xffffffffull || c->clamp_at * 2 > c->clamp_ceiling) {
      c->clamp_at = 1ull << 27; c->clamp_ceiling = 0xffffffffull;
    }
    // Packed path: bytes while they fit ~80 MB of the 126 MB L2 next to the coordinate stream, nibbles up to twice
    // as many cells (500^3, and every grid the XDR format can hold: 512^3 cells = 64 MB of nibbles).  A round is
    // sized so that a uniform density stays far below the lane range: mean 8 hits per byte cell (P[>255] = 0),
    // 1.5 per nibble cell (P[>15] ~ 7e-12 per cell, about one failed round in a thousand on 500^3, which costs that
    // round and the next a pass of the direct kernel); denser spots fail verification and are binned directly.
    {
      const char *benv = getenv("GCB_PACKED_BITS");
      const int want = benv ? atoi(benv) : 0;
      c->packed_bits = want == 4 ? 4 : want == 8 ? 8 : (c->ncells <= ((size_t)80 << 20) ? 8 : 4);
      c->packed_auto = c->ncells * sizeof(uint32_t) > ((size_t)48 << 20) && c->ncells / (c->packed_bits == 8 ? 1 : 2) <= ((size_t)80 << 20) &&
                       !getenv("GCB_NO_PACKED");
      const char *uenv = getenv("GCB_PACKED_UNITS");
      c->packed_units = uenv ? atoll(uenv) : (long long)((double)c->ncells * (c->packed_bits == 8 ? 8.0 : 1.5));
      if (c->packed_units < 4096) c->packed_units = 4096;
      c->packed_bits_home = c->packed_bits; c->packed_units_home = c->packed_units;
      c->packed_fallback_ok = c->packed_bits == 4 && !benv && !uenv && c->packed_auto && !getenv("GCB_NO_BYTE_FALLBACK");
    }
    // Private path: the grid as 8-bit or 4-bit counters next to the 3 x 36 KB tile ring in one CTA's shared memory.
    // A private grid is flushed before its mean load could pass 32 hits per cell (8-bit lanes, range 255: densities
    // that peak at several times their mean stay in range) or 1.5 (4-bit lanes, range 15: P[> 15] ~ 1e-11 per cell
    // of a uniform density); anything denser fails the digit-sum check, is binned again directly, and backs off.
    {
      const size_t optin = (size_t)prop.sharedMemPerBlockOptin, ring = stream_smem_bytes<PrivateSink<8>>();
      const size_t budget = optin > ring + 4096 ? optin - ring - 1024 : 0;
      const size_t w8 = (c->ncells + 3) / 4, w4 = (c->ncells + 7) / 8;
      const char *benv = getenv("GCB_PRIVATE_BITS");
      const int want = benv ? atoi(benv) : 0;
      if (w8 * 4 <= budget && want != 4) { c->priv_bits = 8; c->priv_words = (unsigned)w8; }
      else if (w4 * 4 <= budget) { c->priv_bits = 4; c->priv_words = (unsigned)w4; }
      if (getenv("GCB_NO_PRIVATE") || natoms < 256) c->priv_bits = 0;
      c->priv_ctas = c->sm_count;
      c->priv_limit = c->priv_bits == 8 ? (unsigned long long)c->ncells * 32ull : (unsigned long long)c->ncells * 3ull / 2ull;
      if (const char *e = getenv("GCB_PRIVATE_LIMIT")) c->priv_limit = strtoull(e, nullptr, 10);
      // loading and storing the private grids costs 2 x 4 B x words x CTAs of L2 traffic per launch: worth it from
      // about that many atom-frames on (12 B each would stream in the same time)
      c->priv_min_units = (long long)c->priv_words * c->priv_ctas * 2;
      if (const char *e = getenv("GCB_PRIVATE_MIN_UNITS")) c->priv_min_units = atoll(e);
      c->ap_priv = c->ap;
      c->ap_priv.tile_atoms = std::min(PrivateSink<8>::TILE_CAP, natoms & ~3);
    }
  }
  GCB_CREATE_CUDA(cudaMalloc(&c->d_gindex, sizeof(int) * std::max(gnx, 1)));
  GCB_CREATE_CUDA(cudaMalloc(&c->d_rank, sizeof(unsigned) * rank.size()));
  GCB_CREATE_CUDA(cudaMalloc(&c->d_runcnt, sizeof(uint32_t) * c->ncells));
  GCB_CREATE_CUDA(cudaMalloc(&c->d_diag, sizeof(unsigned long long) * 4));   // [0] edge overflow, [3] saturation flag
  if (gnx) GCB_CREATE_CUDA(cudaMemcpy(c->d_gindex, c->h_gindex.data(), sizeof(int) * gnx, cudaMemcpyHostToDevice));
  GCB_CREATE_CUDA(cudaMemcpy(c->d_rank, rank.data(), sizeof(unsigned) * rank.size(), cudaMemcpyHostToDevice));
  // on the compute stream: it is non-blocking, a legacy-stream cudaMemset could still be
  // running when the first binning kernel starts and would wipe its deposits
  GCB_CREATE_CUDA(cudaMemsetAsync(c->d_runcnt, 0, sizeof(uint32_t) * c->ncells, c->compute));
  GCB_CREATE_CUDA(cudaMemsetAsync(c->d_diag, 0, sizeof(unsigned long long) * 4, c->compute));
  GCB_CREATE_CUDA(cudaStreamSynchronize(c->compute));
#undef GCB_CREATE_CUDA
  *out = c;
  return GCB_OK;
}

void gcb_counter_destroy(gcb_counter *c) { free_counter(c); }

int gcb_counter_reset(gcb_counter *c)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  // No host synchronisation: everything below is ordered on the compute stream behind the work already
  // queued, and results that were never collected are simply dropped (their counter buffers go back to the
  // pool; later launches that reuse them are ordered on the same stream).
  for (auto &s : c->slots) s.pending = 0;
  c->reduce_pending = false;          // statistics of a reduce nobody collected are dropped with the rest
  for (auto &h : c->devhits)
    if (h.d_box3) { GCB_CUDA(cudaStreamSynchronize(c->compute)); cudaFree(h.d_box3); }
  c->devhits.clear();
  for (auto &a : c->hit_arenas) a.used = 0;
  GCB_CUDA(cudaMemsetAsync(c->d_runcnt, 0, sizeof(uint32_t) * c->ncells, c->compute));
  if (c->acc_used) {
    GCB_CUDA(cudaMemsetAsync(c->d_acc, 0, sizeof(float) * c->ncells, c->compute));
    GCB_CUDA(cudaMemsetAsync(c->d_total, 0, sizeof(uint32_t) * c->ncells, c->compute));
    c->acc_used = false;
  }
  GCB_CUDA(cudaMemsetAsync(c->d_diag, 0, sizeof(unsigned long long) * 4, c->compute));
  if (c->d_pstate) GCB_CUDA(cudaMemsetAsync(&c->d_pstate->rounds, 0, 2 * sizeof(unsigned long long), c->compute));
  c->packed_rounds = 0; c->packed_redone = 0;
  if (c->d_priv) {
    if (c->priv_dirty) {
      GCB_CUDA(cudaMemsetAsync(c->d_priv, 0, (size_t)c->priv_ctas * c->priv_words * 4, c->compute));
      GCB_CUDA(cudaMemsetAsync(c->d_priv_hits, 0, sizeof(unsigned long long) * (size_t)c->priv_ctas, c->compute));
    }
    GCB_CUDA(cudaMemsetAsync(&c->d_privstate->launches, 0, 2 * sizeof(unsigned long long), c->compute));
  }
  c->priv_dirty = false; c->priv_bound = 0; c->priv_launches = 0; c->priv_redone = 0;
  if (c->mid_rounds > 0) {
    GCB_CUDA(cudaMemsetAsync(c->d_mid, 0, ((c->ncells + 7) / 8) * sizeof(uint2), c->compute));
    c->mid_rounds = 0;
  }
  c->bound_run = 0; c->bound_total = 0;
  c->reduced = false; c->global_frames = 0;
  c->have_run = false; c->any = false; c->uniform = true; c->w_cur = 0; c->w_first = 0;
  c->hits_per_frame.clear(); c->weight_per_frame.clear();
  c->sumP = 0; c->sumP_frames = 0; c->hits_total = 0; c->edge_overflow = 0;
  return GCB_OK;
}

int gcb_counter_slot(gcb_counter *c, int slot, float **x_pinned, long *capacity_frames)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  int rc = alloc_slots(c);
  if (rc) return rc;
  if (slot < 0 || slot >= (int)c->slots.size()) return gcb::fail(GCB_ERR_ARG, "slot %d out of range", slot);
  Slot &s = c->slots[slot];
  rc = collect_slot(c, s);              // also guarantees the previous H2D out of this slot is done
  if (rc) return rc;
  GCB_CUDA(cudaStreamSynchronize(s.copy));
  if (x_pinned) *x_pinned = s.h_x;
  if (capacity_frames) *capacity_frames = s.cap;
  return GCB_OK;
}

int gcb_counter_submit_slot(gcb_counter *c, int slot, long nframes, const float *weight, const float *box)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  int rc = alloc_slots(c);
  if (rc) return rc;
  if (slot < 0 || slot >= (int)c->slots.size() || nframes < 0 || nframes > c->slots[slot].cap)
    return gcb::fail(GCB_ERR_ARG, "gcb_counter_submit_slot: slot %d / %ld frames out of range", slot, nframes);
  rc = collect_slot(c, c->slots[slot]);
  if (rc) return rc;
  return submit_slot(c, c->slots[slot], nullptr, false, nframes, weight, box);
}

int gcb_counter_add_frames(gcb_counter *c, const float *x, long nframes, const float *weight, const float *box)
{
  if (!c || (!x && nframes > 0) || nframes < 0) return gcb::fail(GCB_ERR_ARG, "gcb_counter_add_frames: bad argument");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  int rc = alloc_slots(c);
  if (rc) return rc;
  cudaPointerAttributes attr;
  bool pinned = false;
  if (cudaPointerGetAttributes(&attr, x) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
  else cudaGetLastError();
  const size_t frame_floats = (size_t)c->natoms * 3;
  // How the frames reach the device (header: gcb_host_staging).  Pageable x has to pass through a pinned slot anyway:
  // host threads copy the index group into it (packed staging).  Pinned x is staged on TWO LANES: the host threads
  // pack batch after batch -- those cross PCIe at 12 B per counted atom -- and whenever the copies already queued
  // would not keep the link busy until the next packed batch is ready, a batch is DMA-ed straight out of the caller's
  // memory to fill the gap.  The host threads and the link are then both busy all the time, whatever their speeds
  // (throughput ~ (link + 2/3 packing rate) / 36 B for water oxygens); a host that packs slowly mostly DMAs in place.
  // A packed batch is queued (copy + kernels) while the pool's threads are already packing the next one.
  const bool can_pack = c->gnx > 0 && nframes > 0 && c->host_pack != GCB_HOST_PACK_NEVER && c->slots[0].cap_pack > 0;
  const bool all_pack = can_pack && (c->host_pack == GCB_HOST_PACK_ALWAYS || !pinned);
  bool two_lanes = can_pack && !all_pack && c->gnx < c->natoms && c->opts.nslots <= 0;
  int measuring = -1;                 // the phase of the trial this call belongs to
  if (two_lanes && nframes > c->slots[0].cap) {
    rc = alloc_slots(c, 6);
    if (rc) return rc;
    if (c->host_pack == GCB_HOST_PACK_AUTO) {
      if (c->lanes_choice < 0) { measuring = c->trial_phase; two_lanes = measuring == 0; }
      else two_lanes = c->lanes_choice == 1;
    }
  } else {
    two_lanes = false;
  }
  timespec call_t0{};
  if (measuring >= 0) clock_gettime(CLOCK_MONOTONIC, &call_t0);
  const double frame_bytes = (double)frame_floats * 4.0, link_rate = 55e9;     // (PCIe 5 x16; a slower link only adds slack)
  const int threads = pack_threads(c);
  const int *gi = c->h_gindex.data();
  const int g0 = c->ap.g0, gstep = c->ap_ok ? c->ap.step : 0;

  struct InFlight {                   // the packed batch whose packing was started last and which is not queued yet
    Slot *s = nullptr; long n = 0, first = 0; void *job = nullptr; timespec t0{};
  } fly;
  auto pack_done = [&]() {            // join the packing of `fly`, note its rate
    gcb::pack_end(fly.job);
    fly.job = nullptr;
    timespec t1{};
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double dt = (double)(t1.tv_sec - fly.t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - fly.t0.tv_nsec);
    const double bytes = (double)fly.n * frame_bytes;
    if (dt > 0 && bytes >= (double)(1 << 20)) {
      const double r = bytes / dt;
      c->pack_rate = c->pack_rate > 0 ? 0.7 * c->pack_rate + 0.3 * r : r;
    }
  };
  auto queue_fly = [&]() -> int {     // copy + kernels of `fly` (its packing is complete)
    if (!fly.s) return GCB_OK;
    Slot *s = fly.s;
    fly.s = nullptr;
    return submit_slot(c, *s, nullptr, false, fly.n, weight ? weight + fly.first : nullptr,
                       box ? box + 9 * (size_t)fly.first : nullptr, true);
  };
  auto bail = [&](int code) { if (fly.job) { gcb::pack_end(fly.job); fly.job = nullptr; } return code; };

  long done = 0;
  while (done < nframes) {
    bool pack = all_pack;
    long n_pack = c->slots[0].cap_pack;
    if (two_lanes) {
      // bytes the link still has to move, and the batch the host threads pack in about the time one in-place copy takes
      double queued = 0;
      for (auto &q : c->slots)
        if (q.copy_queued) {
          if (cudaEventQuery(q.copied) == cudaErrorNotReady) queued += (double)q.copy_bytes;
          else q.copy_queued = false;
        }
      cudaGetLastError();
      if (fly.s) queued += (double)fly.n * (double)c->gnx * 12.0;          // (about to be queued)
      const long cap = c->slots[0].cap;
      const double rate = c->pack_rate > 0 ? c->pack_rate : 4e9 * (double)threads;
      n_pack = std::max<long>(1, std::min<long>(c->slots[0].cap_pack, (long)((double)cap * rate / link_rate)));
      const double t_pack = (double)n_pack * frame_bytes / rate + 100e-6;
      // pack while the link has enough to do meanwhile; the last frames of a call are packed in any case, so that
      // the call does not end with a long copy out of the caller's memory to wait for
      pack = queued >= t_pack * link_rate || nframes - done <= 2 * n_pack;
    }
    Slot &s = c->slots[c->next_slot];
    c->next_slot = (c->next_slot + 1) % (int)c->slots.size();
    if (fly.s == &s) {                // (a single slot) nothing to overlap with
      pack_done();
      rc = queue_fly();
      if (rc) return bail(rc);
    }
    rc = collect_slot(c, s);
    if (rc) return bail(rc);
    if (!pinned || pack) {            // staging buffer free to overwrite
      const cudaError_t err = cudaStreamSynchronize(s.copy);
      if (err != cudaSuccess) return bail(gcb::fail(GCB_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(err)));
    }
    const long n = std::min(pack ? n_pack : s.cap, nframes - done);
    if (pack) {
      if (fly.s) pack_done();         // (one packing job at a time; it has normally finished by now)
      InFlight next;
      next.s = &s; next.n = n; next.first = done;
      clock_gettime(CLOCK_MONOTONIC, &next.t0);
      next.job = gcb::pack_begin(x + (size_t)done * frame_floats, n, c->natoms, gi, c->gnx, g0, gstep, s.h_x, threads);
      rc = queue_fly();               // the previous packed batch: copy + kernels, while the pool packs this one
      fly = next;
      if (rc) return bail(rc);
    } else {
      if (fly.s) {                    // frames are queued in the order of x
        pack_done();
        rc = queue_fly();
        if (rc) return bail(rc);
      }
      rc = submit_slot(c, s, x + (size_t)done * frame_floats, pinned, n, weight ? weight + done : nullptr,
                       box ? box + 9 * (size_t)done : nullptr, false);
      if (rc) return bail(rc);
    }
    done += n;
  }
  if (fly.s) {
    pack_done();
    rc = queue_fly();
    if (rc) return rc;
  }
  // Pinned x that was DMA-ed straight out of the caller's memory: the call returns when the last of those copies has
  // left it (the kernels behind them keep running), so the caller may refill x at once, as with pageable memory.
  if (pinned && !all_pack)
    for (auto &s : c->slots)
      if (s.dma_in_place || !two_lanes) GCB_CUDA(cudaEventSynchronize(s.copied));
  if (measuring >= 0 && c->trial_calls[measuring]++ > 0) {        // (a phase's first call carries the other phase's tail)
    timespec t1{};
    clock_gettime(CLOCK_MONOTONIC, &t1);
    c->trial_s[measuring] += (double)(t1.tv_sec - call_t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - call_t0.tv_nsec);
    c->trial_frames[measuring] += (double)nframes;
    if (c->trial_s[measuring] >= 0.04) {
      if (measuring == 0) c->trial_phase = 1;
      else c->lanes_choice = c->trial_frames[0] / c->trial_s[0] > 1.05 * c->trial_frames[1] / c->trial_s[1] ? 1 : 0;
    }
  }
  return GCB_OK;
}

int gcb_counter_set_host_pack(gcb_counter *c, int mode)
{
  if (!c || mode < GCB_HOST_PACK_AUTO || mode > GCB_HOST_PACK_ALWAYS) return gcb::fail(GCB_ERR_ARG, "gcb_counter_set_host_pack: bad argument");
  c->host_pack = mode;
  return GCB_OK;
}

int gcb_counter_host_staging(gcb_counter *c, gcb_host_staging *out)
{
  if (!c || !out) return gcb::fail(GCB_ERR_ARG, "gcb_counter_host_staging: null argument");
  memset(out, 0, sizeof *out);
  out->mode = c->host_pack;
  out->threads = pack_threads(c);
  out->h2d_bytes = c->h2d_bytes;
  out->frames_packed = c->frames_packed;
  out->frames_in_place = c->frames_in_place;
  out->pack_bytes_per_s = c->pack_rate;
  out->lanes = c->lanes_choice;
  out->two_lane_frames_per_s = c->trial_s[0] > 0 ? c->trial_frames[0] / c->trial_s[0] : 0.0;
  out->in_place_frames_per_s = c->trial_s[1] > 0 ? c->trial_frames[1] / c->trial_s[1] : 0.0;
  return GCB_OK;
}

// per-frame hit counters (and boxes) of a batch of device-resident frames: carved out of the hit arenas, zeroed on
// the compute stream, registered for collection
static int reserve_devhits(gcb_counter *c, long nframes, const float *box, DevHits *out)
{
  DevHits h;
  h.n = nframes;
  {
    const size_t need = ((size_t)nframes + 2 + 15) & ~(size_t)15;
    for (auto &a : c->hit_arenas)
      if (a.cap - a.used >= need) { h.d = a.d + a.used; a.used += need; break; }
    if (!h.d) {
      gcb_counter::HitArena a{nullptr, std::max<size_t>(need * 4, c->hit_arenas.empty() ? (size_t)1 << 16 : c->hit_arenas.back().cap * 2), need};
      GCB_CUDA(cudaMalloc(&a.d, sizeof(unsigned long long) * a.cap));
      c->hit_arenas.push_back(a);
      h.d = a.d;
    }
  }
  GCB_CUDA(cudaMemsetAsync(h.d, 0, sizeof(unsigned long long) * (nframes + 2), c->compute));
  if (c->opts.pbc == GCB_PBC_RECT) {
    if (!box) return gcb::fail(GCB_ERR_ARG, "GCB_PBC_RECT needs the per-frame box");
    std::vector<float> b3((size_t)nframes * 3);
    for (long f = 0; f < nframes; f++) { b3[3 * f] = box[9 * f]; b3[3 * f + 1] = box[9 * f + 4]; b3[3 * f + 2] = box[9 * f + 8]; }
    GCB_CUDA(cudaMalloc(&h.d_box3, sizeof(float) * 3 * nframes));
    GCB_CUDA(cudaMemcpy(h.d_box3, b3.data(), sizeof(float) * 3 * nframes, cudaMemcpyHostToDevice));
  }
  h.frame0 = c->hits_per_frame.size();
  c->hits_per_frame.resize(h.frame0 + (size_t)nframes, 0);
  c->devhits.push_back(h);
  *out = h;
  return GCB_OK;
}

int gcb_counter_add_device_frames(gcb_counter *c, const float *x_dev, long nframes, const float *weight,
                                  const float *box)
{
  if (!c || (!x_dev && nframes > 0) || nframes < 0)
    return gcb::fail(GCB_ERR_ARG, "gcb_counter_add_device_frames: bad argument");
  if (nframes == 0) return GCB_OK;
  GCB_CUDA(cudaSetDevice(c->opts.device));
  DevHits h;
  int rc = reserve_devhits(c, nframes, box, &h);
  if (rc) return rc;
  rc = bin_device_batch(c, x_dev, nframes, weight, h.d_box3, h.d);
  if (rc) return rc;
  for (long f = 0; f < nframes; f++) c->weight_per_frame.push_back(weight ? weight[f] : 1.0f);
  return GCB_OK;
}

extern "C" int gcb_xtc_skim_on_stream_(const uint8_t *buf_dev, const gcb_xtc_frame *frames, long nframes, int natoms, void *meta_dev,
                                       int *status_dev, void *stream, void *meta_host_pinned, void *work_dev);
extern "C" int gcb_xtc_segments_on_stream_(const uint8_t *buf_dev, long nframes, int natoms, float *x_dev, const void *meta_dev,
                                           int *status_dev, void *stream, const void *work_dev);
extern "C" size_t gcb_xtc_work_bytes_(long nframes, int natoms);
extern "C" size_t gcb_xtc_meta_bytes_(long nframes);

// A batch of xtc frames can be binned straight out of the decoder (xtc_bin_kernel) when it is one run of integer
// counts: one weight for the whole batch that continues the current run (or starts one under the rules of
// bin_device_batch), the uint32 grid as the target (grids on the packed-counter path take the two-step route),
// and not more frames than one launch may add to a cell.  Prepares the run exactly as bin_device_batch would.
static bool xtc_batch_fusable(gcb_counter *c, long n, const float *w)
{
  const bool off = getenv("GCB_XTC_UNFUSED") != nullptr;
  if (off || c->opts.kernel != GCB_KERNEL_AUTO) return false;
  if (c->packed_bits != 0 && c->packed_auto) return false;
  const float wr = w ? w[0] : 1.0f;
  for (long f = 1; f < n; f++) if (!same_bits(w ? w[f] : 1.0f, wr)) return false;
  const long max_frames = (long)std::max<unsigned long long>(1, (c->clamp_ceiling / 4) / (unsigned long long)std::max(c->gnx, 1));
  if (n > max_frames) return false;
  const bool continues = c->have_run && same_bits(wr, c->w_cur);
  const bool fresh = !c->have_run && !c->acc_used;
  const bool long_run = (double)n * c->gnx >= (double)c->ncells / 8.0;
  return continues || fresh || long_run;
}

static int xtc_bin_batch(gcb_counter *c, const uint8_t *d_buf, const void *d_meta, const void *d_work, const int *d_status,
                         long n, const float *w, const float *d_box3, unsigned long long *d_hits)
{
  const float wr = w ? w[0] : 1.0f;
  const unsigned long long units = (unsigned long long)n * (unsigned long long)c->gnx;
  if (!c->any) { c->any = true; c->w_first = wr; }
  else if (!same_bits(wr, c->w_first)) c->uniform = false;
  int rc;
  if (c->have_run && !same_bits(wr, c->w_cur)) { rc = fold_pending(c); if (rc) return rc; }
  c->have_run = true;
  c->w_cur = wr;
  rc = ensure_headroom(c, c->d_runcnt, c->bound_run, units);
  if (rc) return rc;
  c->bound_run += units;
  const int *nseg = static_cast<const int *>(d_work);
  const xtcdev::Checkpoint *ckpt = reinterpret_cast<const xtcdev::Checkpoint *>(static_cast<const char *>(d_work) + xtcdev::work_ckpt_offset(n));
  const long long nthreads = (long long)n * xtcdev::max_segments(c->natoms);
  const unsigned grid = (unsigned)((nthreads + 127) / 128);
  const xtcdev::FrameDev *meta = static_cast<const xtcdev::FrameDev *>(d_meta);
  const bool pbc = c->opts.pbc == GCB_PBC_RECT;
  rc = prof_begin(c);
  if (rc) return rc;
#define GCB_XTC_BIN(PBC_, AP_)                                                                                              \
  xtc_bin_kernel<PBC_, AP_><<<grid, 128, 0, c->compute>>>(d_buf, meta, n, c->natoms, ckpt, nseg, d_status, c->P, c->d_runcnt, \
                                                          c->ap.g0, c->ap.step, c->gnx, c->d_rank, d_box3, d_hits, c->d_diag)
  if (c->ap_ok) { if (pbc) GCB_XTC_BIN(true, true); else GCB_XTC_BIN(false, true); }
  else          { if (pbc) GCB_XTC_BIN(true, false); else GCB_XTC_BIN(false, false); }
#undef GCB_XTC_BIN
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  return prof_end(c, GCB_PROF_OTHER, (double)units);
}

int gcb_counter_add_xtc(gcb_counter *c, const uint8_t *buf, size_t len, const gcb_xtc_frame *frames, long nframes,
                        const float *weight)
{
  if (!c || !buf || !frames || nframes < 0) return gcb::fail(GCB_ERR_ARG, "gcb_counter_add_xtc: bad argument");
  if (nframes == 0) return GCB_OK;
  GCB_CUDA(cudaSetDevice(c->opts.device));
  if (!c->xtc[0].stream) {
    for (auto &x : c->xtc) {
      GCB_CUDA(cudaStreamCreateWithFlags(&x.stream, cudaStreamNonBlocking));
      GCB_CUDA(cudaEventCreateWithFlags(&x.skimmed, cudaEventDisableTiming));
      GCB_CUDA(cudaEventCreateWithFlags(&x.consumed, cudaEventDisableTiming));
    }
  }
  static const bool trace = getenv("GCB_TRACE") != nullptr;
  auto now = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; };
  const double t_begin = now();
  const size_t frame_bytes = (size_t)c->natoms * 12;
  // The pipeline.  A batch goes through: H2D of its compressed bytes -> skim -> segment kernel(s).  The skim walks
  // each frame's control bits as one chain of dependent steps: ~2 ms for a 60 000-atom frame whether the batch has 64
  // frames or 1024.  So every batch has its own set of buffers and its own stream for copy + skim (GCB_XTC_SETS of them
  // in flight: the skims of several batches overlap each other and the copies), and only the segment kernels, which
  // are short and wide, queue on the compute stream.  Batches of ~64 MB of compressed bytes (1.2 ms of PCIe), at
  // least 64 frames.
  const char *env = getenv("GCB_XTC_BATCH");
  long per;
  if (env) per = atol(env);
  else {
    const double bytes_per_frame = (double)(frames[nframes - 1].payload_offset + frames[nframes - 1].payload_bytes - frames[0].payload_offset) / (double)nframes;
    per = (long)(((size_t)64 << 20) / std::max(bytes_per_frame, 1.0));
    // at least 64 frames: a frame's skim takes ~30 ns per atom and its compressed bytes ~0.1 ns per atom of PCIe
    // time, so with GCB_XTC_SETS batches in flight the copies hide the skims from ~55 frames per batch on,
    // whatever the frame size (large systems simply get large batches: 1 M atoms -> 320 MB of compressed bytes)
    per = std::max<long>(64, std::min<long>(per, 16384));
  }
  per = std::max<long>(4, per & ~3L);                         // batches start 16-byte aligned in the decoded array
  std::vector<float> box3;
  int set = 0;
  auto check_set = [&](gcb_counter::XtcSet &x) -> int {       // wait for a set's batch, look at the decoder's verdicts
    if (x.nframes == 0) return GCB_OK;
    GCB_CUDA(cudaEventSynchronize(x.consumed));
    for (long f = 0; f < x.nframes; f++)
      if (x.h_status[f]) {
        x.nframes = 0;
        return gcb::fail(GCB_ERR_FORMAT, "xtc frame %ld: damaged compressed stream (frames of its batch that were decoded before the "
                                         "damage may have been binned: reset the counter)", x.first + f);
      }
    x.nframes = 0;
    return GCB_OK;
  };
  long nfused = 0;
  // the first two batches are a quarter and a half of the others: the pipeline starts working after a quarter of a
  // batch's copy time (nothing overlaps the first copy and the first skim)
  long f0 = 0;
  for (long batch = 0; f0 < nframes; batch++, set = (set + 1) % GCB_XTC_SETS) {
    const long want = batch == 0 ? std::max<long>(16, (per / 4) & ~3L) : batch == 1 ? std::max<long>(16, (per / 2) & ~3L) : per;
    const long n = std::min(want, nframes - f0);
    gcb_counter::XtcSet &x = c->xtc[set];
    int rc = check_set(x);
    if (rc) return rc;
    size_t lo = frames[f0].payload_offset, hi = 0;
    for (long f = f0; f < f0 + n; f++) {
      if (frames[f].natoms != c->natoms) return gcb::fail(GCB_ERR_FORMAT, "xtc frame %ld has %d atoms, the counter %d", f, frames[f].natoms, c->natoms);
      const size_t b = frames[f].payload_offset;
      const size_t nb = frames[f].natoms <= 9 ? (size_t)frames[f].natoms * 12 : ((size_t)frames[f].payload_bytes + 3) & ~(size_t)3;
      if (b + nb > len) return gcb::fail(GCB_ERR_ARG, "xtc frame %ld lies outside the buffer", f);
      lo = std::min(lo, b); hi = std::max(hi, b + nb);
    }
    lo &= ~(size_t)3;
    const float *wb = weight ? weight + f0 : nullptr;
    const bool fused = xtc_batch_fusable(c, n, wb);
    if (hi - lo + 16 > x.buf_cap) { cudaFree(x.d_buf); x.d_buf = nullptr; x.buf_cap = (hi - lo) + (hi - lo) / 8 + 16; GCB_CUDA(cudaMalloc(&x.d_buf, x.buf_cap)); }
    if (n > x.frames_cap) {
      cudaFree(x.d_meta); cudaFree(x.d_status); cudaFree(x.d_work); cudaFreeHost(x.h_meta); cudaFreeHost(x.h_status);
      x.d_meta = x.h_meta = x.d_work = nullptr; x.d_status = x.h_status = nullptr;
      x.frames_cap = n;
      GCB_CUDA(cudaMalloc(&x.d_meta, gcb_xtc_meta_bytes_(n)));
      GCB_CUDA(cudaMalloc(&x.d_status, sizeof(int) * (size_t)n));
      GCB_CUDA(cudaMalloc(&x.d_work, gcb_xtc_work_bytes_(n, c->natoms)));
      GCB_CUDA(cudaHostAlloc(&x.h_meta, gcb_xtc_meta_bytes_(n), cudaHostAllocDefault));
      GCB_CUDA(cudaHostAlloc((void **)&x.h_status, sizeof(int) * (size_t)n, cudaHostAllocDefault));
    }
    if (!fused && (size_t)n * frame_bytes + 16 > x.x_cap) { cudaFree(x.d_x); x.d_x = nullptr; x.x_cap = (size_t)n * frame_bytes + 16; GCB_CUDA(cudaMalloc(&x.d_x, x.x_cap)); }
    // compressed bytes, frame table and skim on the set's stream; the segment kernels on the compute stream
    GCB_CUDA(cudaMemcpyAsync(x.d_buf, buf + lo, hi - lo, cudaMemcpyHostToDevice, x.stream));
    std::vector<gcb_xtc_frame> rel(frames + f0, frames + f0 + n);
    for (auto &r : rel) r.payload_offset -= lo;
    const float *boxarg = nullptr;
    if (c->opts.pbc == GCB_PBC_RECT) {
      box3.resize((size_t)n * 9);
      for (long f = 0; f < n; f++) memcpy(&box3[(size_t)f * 9], frames[f0 + f].box, sizeof(float) * 9);
      boxarg = box3.data();
    }
    rc = gcb_xtc_skim_on_stream_(x.d_buf, rel.data(), n, c->natoms, x.d_meta, x.d_status, x.stream, x.h_meta, x.d_work);
    if (rc) return rc;
    GCB_CUDA(cudaEventRecord(x.skimmed, x.stream));
    GCB_CUDA(cudaStreamWaitEvent(c->compute, x.skimmed, 0));
    if (!fused) {
      rc = gcb_xtc_segments_on_stream_(x.d_buf, n, c->natoms, x.d_x, x.d_meta, x.d_status, c->compute, x.d_work);
      if (rc) return rc;
    }
    GCB_CUDA(cudaMemcpyAsync(x.h_status, x.d_status, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->compute));
    if (fused) {
      DevHits h;
      rc = reserve_devhits(c, n, boxarg, &h);
      if (rc) return rc;
      rc = xtc_bin_batch(c, x.d_buf, x.d_meta, x.d_work, x.d_status, n, wb, h.d_box3, h.d);
      if (rc) return rc;
      for (long f = 0; f < n; f++) c->weight_per_frame.push_back(wb ? wb[f] : 1.0f);
      nfused += n;
    } else {
      rc = gcb_counter_add_device_frames(c, x.d_x, n, wb, boxarg);
      if (rc) return rc;
    }
    // (the next copy into this set is queued only after the host has waited for this event in check_set: no wait on
    // the copy stream here, it would hold back the copy of the NEXT batch, which goes into another set)
    GCB_CUDA(cudaEventRecord(x.consumed, c->compute));
    x.nframes = n; x.first = f0;
    f0 += n;
  }
  const double t_queued = now();
  for (auto &x : c->xtc) { int rc = check_set(x); if (rc) return rc; }
  if (trace)
    fprintf(stderr, "[gcb add_xtc] %ld frames in batches of %ld (%ld binned out of the decoder): queued after %.3f ms, done after %.3f ms\n",
            nframes, per, nfused, t_queued - t_begin, now() - t_begin);
  return GCB_OK;
}

static int finish_reduce(gcb_counter *c);

int gcb_counter_sync(gcb_counter *c)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  { int rc = finish_reduce(c); if (rc) return rc; }
  for (auto &s : c->slots) { int rc = collect_slot(c, s); if (rc) return rc; }
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  for (auto &h : c->devhits) {
    std::vector<unsigned long long> tmp((size_t)h.n);
    GCB_CUDA(cudaMemcpy(tmp.data(), h.d, sizeof(unsigned long long) * h.n, cudaMemcpyDeviceToHost));
    for (long f = 0; f < h.n; f++) c->hits_per_frame[h.frame0 + f] = tmp[f];
    cudaFree(h.d_box3);
  }
  c->devhits.clear();
  for (auto &a : c->hit_arenas) a.used = 0;
  {
    unsigned long long diag[4];
    GCB_CUDA(cudaMemcpy(diag, c->d_diag, sizeof(diag), cudaMemcpyDeviceToHost));
    if (!c->reduced) c->edge_overflow = diag[0];      // after an all-reduce the statistic is the global one
    c->saturated = diag[3] != 0;
    if (c->d_pstate) {
      PackedState ps;
      GCB_CUDA(cudaMemcpy(&ps, c->d_pstate, sizeof(ps), cudaMemcpyDeviceToHost));
      c->packed_rounds = ps.rounds; c->packed_redone = ps.redone;
    }
    if (c->d_privstate) {
      PrivState ps;
      GCB_CUDA(cudaMemcpy(&ps, c->d_privstate, sizeof(ps), cudaMemcpyDeviceToHost));
      c->priv_launches = ps.launches; c->priv_redone = ps.redone + (ps.failed ? 1 : 0);
    }
  }
  // sumP: `sumP += (double)weight` once per hit, frame by frame (g_ri3Dc.c:151,426)
  for (; c->sumP_frames < c->hits_per_frame.size(); c->sumP_frames++) {
    const uint64_t h = c->hits_per_frame[c->sumP_frames];
    c->sumP = gcb::seq_add_f64(c->sumP, (double)c->weight_per_frame[c->sumP_frames], h);
    c->hits_total += h;
  }
  return GCB_OK;
}

int gcb_counter_stats(gcb_counter *c, gcb_stats *out)
{
  if (!c || !out) return gcb::fail(GCB_ERR_ARG, "null argument");
  int rc = gcb_counter_sync(c);
  if (rc) return rc;
  out->frames = c->reduced ? c->global_frames : c->hits_per_frame.size();
  out->atom_frames = out->frames * (uint64_t)c->gnx;
  out->hits = c->hits_total;
  out->edge_overflow = c->edge_overflow;
  out->sumP = c->sumP;
  out->uniform_weight = c->uniform ? 1 : 0;
  out->weight = c->w_first;
  out->kernel_launches = c->launches;
  out->packed_rounds = c->packed_rounds;
  out->packed_redone = c->packed_redone;
  out->private_launches = c->priv_launches;
  out->private_redone = c->priv_redone;
  out->packed_bits = c->packed_bits;
  out->saturated = c->saturated ? 1 : 0;
  return GCB_OK;
}

int gcb_counter_hits_per_frame(gcb_counter *c, uint64_t *hits, long capacity, long *nframes_out)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  int rc = gcb_counter_sync(c);
  if (rc) return rc;
  const long n = (long)c->hits_per_frame.size();
  if (nframes_out) *nframes_out = n;
  if (hits) memcpy(hits, c->hits_per_frame.data(), sizeof(uint64_t) * (size_t)std::min(n, capacity));
  return GCB_OK;
}

__global__ void add_u32_kernel(const uint32_t *__restrict__ p, const uint32_t *__restrict__ q, uint32_t *__restrict__ o, size_t n)
{
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x)
    o[c] = p[c] + (q ? q[c] : 0u);
}

static int ensure_out(gcb_counter *c)
{
  if (!c->d_out) GCB_CUDA(cudaMalloc(&c->d_out, c->ncells * sizeof(float)));
  return GCB_OK;
}

int gcb_counter_device_counts(gcb_counter *c, uint32_t **counts_dev)
{
  if (!c || !counts_dev) return gcb::fail(GCB_ERR_ARG, "null argument");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  int rc = flush_private(c);
  if (rc) return rc;
  if (!c->acc_used) { *counts_dev = c->d_runcnt; return GCB_OK; }
  rc = ensure_out(c);
  if (rc) return rc;
  add_u32_kernel<<<elementwise_grid(c, c->ncells), 256, 0, c->compute>>>(c->d_total, c->have_run ? c->d_runcnt : nullptr,
                                                                          (uint32_t *)c->d_out, c->ncells);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  *counts_dev = (uint32_t *)c->d_out;
  return GCB_OK;
}

int gcb_counter_counts(gcb_counter *c, uint32_t *counts_zfast)
{
  if (!c || !counts_zfast) return gcb::fail(GCB_ERR_ARG, "null argument");
  uint32_t *d = nullptr;
  int rc = gcb_counter_device_counts(c, &d);
  if (rc) return rc;
  GCB_CUDA(cudaMemcpyAsync(counts_zfast, d, c->ncells * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->compute));
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  return GCB_OK;
}

static int density_to_device(gcb_counter *c, double denom)
{
  int rc = ensure_out(c);
  if (rc) return rc;
  rc = flush_private(c);
  if (rc) return rc;
  density_kernel<<<elementwise_grid(c, c->ncells), 256, 0, c->compute>>>(c->d_runcnt, c->acc_used ? c->d_acc : nullptr, c->w_cur,
                                                                          c->have_run ? 1 : 0, denom, c->d_out, c->ncells);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  return GCB_OK;
}

int gcb_counter_device_density(gcb_counter *c, double T_tot, float **grid_zfast_dev)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  const float dV = (float)gcb_delta_v(&c->tg);
  int rc = density_to_device(c, (double)dV * T_tot);
  if (rc) return rc;
  if (grid_zfast_dev) *grid_zfast_dev = c->d_out;
  return GCB_OK;
}

int gcb_counter_occupancy(gcb_counter *c, float *occ_zfast)
{
  if (!c || !occ_zfast) return gcb::fail(GCB_ERR_ARG, "null argument");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  int rc = density_to_device(c, 0.0);
  if (rc) return rc;
  GCB_CUDA(cudaMemcpyAsync(occ_zfast, c->d_out, c->ncells * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  return GCB_OK;
}

int gcb_counter_density(gcb_counter *c, double T_tot, float *grid_zfast, float *grid_xfast, gcb_tgrid *tg_out)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  const float dV = (float)gcb_delta_v(&c->tg);            // `real dV = DeltaV(&tgrid)` (g_ri3Dc.c:453)
  const double denom = (double)dV * T_tot;                // (dV*T_tot) with T_tot double (g_ri3Dc.c:458)
  int rc = density_to_device(c, denom);
  if (rc) return rc;
  if (grid_zfast)
    GCB_CUDA(cudaMemcpyAsync(grid_zfast, c->d_out, c->ncells * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
  if (grid_xfast) {
    if (!c->d_out2) GCB_CUDA(cudaMalloc(&c->d_out2, c->ncells * sizeof(float)));
    dim3 block(32, 8), grid((c->tg.mx[2] + 31) / 32, (c->tg.mx[0] + 31) / 32, c->tg.mx[1]);
    zfast_to_xfast_kernel<<<grid, block, 0, c->compute>>>(c->d_out, c->d_out2, c->tg.mx[0], c->tg.mx[1], c->tg.mx[2]);
    c->launches++;
    GCB_CUDA(cudaGetLastError());
    GCB_CUDA(cudaMemcpyAsync(grid_xfast, c->d_out2, c->ncells * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
  }
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  if (tg_out) { *tg_out = c->tg; tg_out->tweight = (float)T_tot; }   // g_ri3Dc.c:460
  return GCB_OK;
}

// The finished density as file images produced on the device.  The prologues are a few dozen bytes and
// come from grid_io.cpp; the cell payload is reordered and encoded by zfast_to_file_kernel and copied
// straight behind the prologue in the caller's buffer.
extern "C" size_t gcb_xdr_prologue_(const gcb_tgrid *tg, const char *header, uint8_t *out);
extern "C" size_t gcb_plt_prologue_(const gcb_tgrid *tg, uint8_t *out);

static int emit_file_payload(gcb_counter *c, double T_tot, int emit, float scale, uint8_t *payload_out)
{
  const float dV = (float)gcb_delta_v(&c->tg);            // `real dV = DeltaV(&tgrid)` (g_ri3Dc.c:453)
  int rc = density_to_device(c, (double)dV * T_tot);
  if (rc) return rc;
  if (!c->d_out2) GCB_CUDA(cudaMalloc(&c->d_out2, c->ncells * sizeof(float)));
  dim3 block(32, 8), grid((c->tg.mx[2] + 31) / 32, (c->tg.mx[0] + 31) / 32, c->tg.mx[1]);
  if (emit == 1)
    zfast_to_file_kernel<1><<<grid, block, 0, c->compute>>>(c->d_out, (uint32_t *)c->d_out2, c->tg.mx[0], c->tg.mx[1], c->tg.mx[2], 1.0f);
  else
    zfast_to_file_kernel<2><<<grid, block, 0, c->compute>>>(c->d_out, (uint32_t *)c->d_out2, c->tg.mx[0], c->tg.mx[1], c->tg.mx[2], scale);
  c->launches++;
  GCB_CUDA(cudaGetLastError());
  GCB_CUDA(cudaMemcpyAsync(payload_out, c->d_out2, c->ncells * sizeof(float), cudaMemcpyDeviceToHost, c->compute));
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  return GCB_OK;
}

int gcb_counter_xdr(gcb_counter *c, double T_tot, const char *header, uint8_t *out, size_t cap, size_t *written,
                    gcb_tgrid *tg_out)
{
  if (!c || !out) return gcb::fail(GCB_ERR_ARG, "gcb_counter_xdr: null argument");
  if (!header) header = "";
  if (strlen(header) > GCB_HEADER_MAX) return gcb::fail(GCB_ERR_ARG, "header longer than %d bytes", GCB_HEADER_MAX);
  if (c->ncells > GCB_NGRID_MAX) return gcb::fail(GCB_ERR_RANGE, "grid has %zu cells; the format allows %u", c->ncells, GCB_NGRID_MAX);
  GCB_CUDA(cudaSetDevice(c->opts.device));
  gcb_tgrid tg = c->tg;
  tg.tweight = (float)T_tot;                               // g_ri3Dc.c:460
  const size_t need = gcb_xdr_size(&tg, header);
  if (cap < need) return gcb::fail(GCB_ERR_ARG, "gcb_counter_xdr: buffer of %zu bytes, need %zu", cap, need);
  const size_t pro = gcb_xdr_prologue_(&tg, header, out);
  int rc = emit_file_payload(c, T_tot, 1, 1.0f, out + pro);
  if (rc) return rc;
  if (written) *written = need;
  if (tg_out) *tg_out = tg;
  return GCB_OK;
}

int gcb_counter_grid_write(gcb_counter *c, double T_tot, const char *path, const char *header, gcb_tgrid *tg_out)
{
  if (!c || !path) return gcb::fail(GCB_ERR_ARG, "gcb_counter_grid_write: null argument");
  gcb_tgrid tg = c->tg;
  const size_t need = gcb_xdr_size(&tg, header ? header : "");
  uint8_t *buf = nullptr;
  GCB_CUDA(cudaSetDevice(c->opts.device));
  GCB_CUDA(cudaHostAlloc((void **)&buf, need, cudaHostAllocDefault));     // pinned: the payload arrives by DMA
  size_t n = 0;
  int rc = gcb_counter_xdr(c, T_tot, header, buf, need, &n, tg_out);
  if (!rc) {
    FILE *fp = fopen(path, "wb");
    if (!fp) rc = gcb::fail(GCB_ERR_IO, "cannot open %s for writing", path);
    else {
      const bool ok = fwrite(buf, 1, n, fp) == n;
      if (fclose(fp) != 0 || !ok) rc = gcb::fail(GCB_ERR_IO, "short write to %s", path);
    }
  }
  cudaFreeHost(buf);
  return rc;
}

int gcb_counter_plt(gcb_counter *c, double T_tot, float unit, uint8_t *out, size_t cap, size_t *written)
{
  if (!c || !out) return gcb::fail(GCB_ERR_ARG, "gcb_counter_plt: null argument");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  const size_t need = 44 + 4 * c->ncells;
  if (cap < need) return gcb::fail(GCB_ERR_ARG, "gcb_counter_plt: buffer of %zu bytes, need %zu", cap, need);
  const size_t pro = gcb_plt_prologue_(&c->tg, out);
  const float scale = (float)(1. / (double)unit);          // `real norm = 1./unit` (plt.c:48)
  int rc = emit_file_payload(c, T_tot, 2, scale, out + pro);
  if (rc) return rc;
  if (written) *written = need;
  return GCB_OK;
}

int gcb_counter_event_record(gcb_counter *c, int id)
{
  if (!c || id < 0 || id >= GCB_MAX_EVENTS) return gcb::fail(GCB_ERR_ARG, "gcb_counter_event_record: bad argument");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  if ((int)c->events.size() <= id) c->events.resize((size_t)id + 1, nullptr);
  if (!c->events[id]) GCB_CUDA(cudaEventCreate(&c->events[id]));
  GCB_CUDA(cudaEventRecord(c->events[id], c->compute));
  return GCB_OK;
}

int gcb_counter_event_elapsed(gcb_counter *c, int id0, int id1, float *ms)
{
  if (!c || !ms || id0 < 0 || id1 < 0 || id0 >= (int)c->events.size() || id1 >= (int)c->events.size() ||
      !c->events[id0] || !c->events[id1])
    return gcb::fail(GCB_ERR_STATE, "gcb_counter_event_elapsed: events %d/%d were not recorded", id0, id1);
  GCB_CUDA(cudaSetDevice(c->opts.device));
  GCB_CUDA(cudaEventSynchronize(c->events[id1]));
  GCB_CUDA(cudaEventElapsedTime(ms, c->events[id0], c->events[id1]));
  return GCB_OK;
}

int gcb_counter_profile(gcb_counter *c, int enable)
{
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  c->prof = enable != 0;
  c->prof_used = 0;
  c->prof_kind.clear();
  c->prof_units.clear();
  return GCB_OK;
}

int gcb_counter_profile_kinds(gcb_counter *c, double ms[GCB_PROF_KINDS], long launches[GCB_PROF_KINDS], double atom_frames[GCB_PROF_KINDS])
{
  if (!c || !ms || !launches || !atom_frames) return gcb::fail(GCB_ERR_ARG, "gcb_counter_profile_kinds: null argument");
  GCB_CUDA(cudaSetDevice(c->opts.device));
  GCB_CUDA(cudaStreamSynchronize(c->compute));
  for (int k = 0; k < GCB_PROF_KINDS; k++) { ms[k] = 0; launches[k] = 0; atom_frames[k] = 0; }
  for (size_t i = 0; i + 1 < c->prof_used; i += 2) {
    float t = 0;
    GCB_CUDA(cudaEventElapsedTime(&t, c->prof_ev[i], c->prof_ev[i + 1]));
    const int k = c->prof_kind[i / 2];
    ms[k] += t; launches[k]++; atom_frames[k] += c->prof_units[i / 2];
  }
  return GCB_OK;
}

// the kind of binning kernel that took the most time (direct, packed pass or its redo pass)
int gcb_counter_profile_read(gcb_counter *c, double *k1_ms, long *k1_launches, double *k1_atom_frames)
{
  double ms[GCB_PROF_KINDS], units[GCB_PROF_KINDS];
  long n[GCB_PROF_KINDS];
  int rc = gcb_counter_profile_kinds(c, ms, n, units);
  if (rc) return rc;
  int best = GCB_PROF_DIRECT;
  if (ms[GCB_PROF_PACKED] > ms[best]) best = GCB_PROF_PACKED;
  if (ms[GCB_PROF_REDO] > ms[best]) best = GCB_PROF_REDO;
  if (ms[GCB_PROF_PRIVATE] > ms[best]) best = GCB_PROF_PRIVATE;
  if (k1_ms) *k1_ms = ms[best];
  if (k1_launches) *k1_launches = n[best];
  if (k1_atom_frames) *k1_atom_frames = units[best];
  return GCB_OK;
}

int gcb_counter_timer_start(gcb_counter *c) { return gcb_counter_event_record(c, 0); }

int gcb_counter_timer_stop(gcb_counter *c, float *ms)
{
  int rc = gcb_counter_event_record(c, 1);
  if (rc) return rc;
  return gcb_counter_event_elapsed(c, 0, 1, ms);
}

int gcb_gen_device_frames(const gcb_gen *p, long frame0, long nframes, int natoms, float *x_dev, int device)
{
  if (!p || !x_dev || nframes < 0 || natoms <= 0) return gcb::fail(GCB_ERR_ARG, "gcb_gen_device_frames: bad argument");
  if (p->mode == GCB_GEN_CLUSTER && p->nsites <= 0) return gcb::fail(GCB_ERR_ARG, "cluster mode needs nsites > 0");
  GCB_CUDA(cudaSetDevice(device));
  if (nframes == 0) return GCB_OK;
  const long long total = (long long)nframes * natoms;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  gen_frames_kernel<<<grid, 256>>>(*p, frame0, nframes, natoms, x_dev);
  GCB_CUDA(cudaGetLastError());
  GCB_CUDA(cudaDeviceSynchronize());
  return GCB_OK;
}

int gcb_device_malloc(void **p, size_t bytes, int device)
{
  if (!p) return gcb::fail(GCB_ERR_ARG, "null argument");
  GCB_CUDA(cudaSetDevice(device));
  GCB_CUDA(cudaMalloc(p, bytes));
  return GCB_OK;
}
int gcb_device_free(void *p, int device)
{
  GCB_CUDA(cudaSetDevice(device));
  GCB_CUDA(cudaFree(p));
  return GCB_OK;
}
int gcb_host_alloc_pinned(void **p, size_t bytes)
{
  if (!p) return gcb::fail(GCB_ERR_ARG, "null argument");
  GCB_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
  return GCB_OK;
}
int gcb_host_free_pinned(void *p)
{
  GCB_CUDA(cudaFreeHost(p));
  return GCB_OK;
}
int gcb_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, int device)
{
  GCB_CUDA(cudaSetDevice(device));
  GCB_CUDA(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
  return GCB_OK;
}
int gcb_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, int device)
{
  GCB_CUDA(cudaSetDevice(device));
  GCB_CUDA(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
  return GCB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// multi-GPU reduce (the NCCL binding lives in nccl_reduce.cpp)
// ---------------------------------------------------------------------------
extern "C" int gcb_counter_comm_attach_(gcb_counter *c, void *comm, const gcb_reduce_ops *ops, int rank, int nranks, gcb_ctl *ctl)
{
  GCB_CUDA(cudaSetDevice(c->opts.device));
  if (comm) {
    if (c->nccl) return gcb::fail(GCB_ERR_STATE, "the counter already has a communicator");
    c->nccl = comm; c->ops = ops; c->rank = rank; c->nranks = nranks; c->ctl = ctl;
  }
  return GCB_OK;
}

// After this call every rank holds the global integer counts and statistics.
// Counts are an integer sum, so the result is independent of the rank count.
// Exact f32 occupancy needs one weight for the whole job (checked); with
// differing weights the ranks' f32 accumulators are summed instead, which is
// not the reference's frame-ordered chain (see gridcount_b200.h).
//
// What the call costs the host:
//   1. the ranks' {weight, uniformity, frame count, saturation bounds} table is host-side knowledge.  On one node it
//      is exchanged through shared memory (shm_ctl.cpp): no kernel, no stream, nothing that has to find an SM while
//      the persistent binning CTAs hold them all.  Ranks on several nodes fall back to a small NCCL all-reduce on a
//      side stream.
//   2. the per-frame hit counters never visit the host on their way: they are copied device to device into this
//      rank's block of the gather buffer, and the grid all-reduce and the (in-place) all-gather of the blocks go
//      out as ONE NCCL group behind the binning kernels.  The call returns after queueing.
extern "C" int gcb_counter_allreduce(gcb_counter *c)
{
  enum { U32 = 3, U64 = 5, F32 = 7, F64 = 8 };
  if (!c) return gcb::fail(GCB_ERR_ARG, "null counter");
  if (c->nranks <= 1 || !c->nccl) return gcb_counter_sync(c);
  const gcb_reduce_ops *ops = c->ops;
  GCB_CUDA(cudaSetDevice(c->opts.device));
  static const bool trace = getenv("GCB_TRACE") != nullptr;
  auto now = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; };
  const double t_begin = now();
  int rc = finish_reduce(c);                  // a previous reduce whose statistics nobody asked for
  if (rc) return rc;
  rc = flush_private(c);                      // the grid that is reduced must hold everything
  if (rc) return rc;
  if (c->reduce_done) GCB_CUDA(cudaEventSynchronize(c->reduce_done));          // (or that a reset dropped): its D2H owns the pinned scratch
  for (auto &s : c->slots) { rc = collect_slot(c, s); if (rc) return rc; }     // host-staged batches report through pinned memory
  if (!c->comm_stream) GCB_CUDA(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
  // Persistent scratch (device + pinned mirror): no allocation per call.
  auto scratch = [&](size_t words) -> int {
    if (words <= c->scr_words) return GCB_OK;
    cudaFree(c->d_scr); cudaFreeHost(c->h_scr);
    c->d_scr = nullptr; c->h_scr = nullptr; c->scr_words = 0;
    const size_t cap = words + words / 2 + 64;
    GCB_CUDA(cudaMalloc(&c->d_scr, sizeof(unsigned long long) * cap));
    GCB_CUDA(cudaHostAlloc((void **)&c->h_scr, sizeof(unsigned long long) * cap, cudaHostAllocDefault));
    c->scr_words = cap;
    return GCB_OK;
  };
  // 1. each rank's {has frames, weight bits, uniform, f32 accumulator in use, frame count, saturation bounds}
  const size_t R = (size_t)c->nranks;
  constexpr size_t TW = GCB_CTL_WORDS;
  std::vector<unsigned long long> table(R * TW, 0ull);
  {
    uint32_t wb; memcpy(&wb, &c->w_first, 4);
    unsigned long long mine[TW] = {c->any ? 1ull : 0ull, wb, c->uniform ? 1ull : 0ull, c->acc_used ? 1ull : 0ull,
                                   c->hits_per_frame.size(), c->bound_run, c->bound_total, 0ull};
    if (c->ctl) {
      rc = gcb_ctl_allgather(c->ctl, (const uint64_t *)mine, (uint64_t *)table.data(), 0);    // no timeout: like a collective
      if (rc) return rc;
    } else {
      rc = scratch(R * TW);
      if (rc) return rc;
      memset(c->h_scr, 0, sizeof(unsigned long long) * R * TW);
      memcpy(c->h_scr + (size_t)c->rank * TW, mine, sizeof(mine));
      GCB_CUDA(cudaMemcpyAsync(c->d_scr, c->h_scr, sizeof(unsigned long long) * R * TW, cudaMemcpyHostToDevice, c->comm_stream));
      rc = ops->allreduce(c->d_scr, c->d_scr, R * TW, U64, c->nccl, c->comm_stream);     // disjoint slots: a sum is a gather
      if (rc) return rc;
      GCB_CUDA(cudaMemcpyAsync(c->h_scr, c->d_scr, sizeof(unsigned long long) * R * TW, cudaMemcpyDeviceToHost, c->comm_stream));
      GCB_CUDA(cudaStreamSynchronize(c->comm_stream));
      memcpy(table.data(), c->h_scr, sizeof(unsigned long long) * R * TW);
    }
  }
  bool global_uniform = true, any_acc = false, have_w = false;
  uint32_t w0 = 0;
  size_t total_frames = 0, my_off = 0, max_frames = 0;
  unsigned long long sum_bound_run = 0, sum_bound_total = 0;
  c->pend_frames.assign(R, 0);
  for (int r = 0; r < c->nranks; r++) {
    const unsigned long long *o = table.data() + (size_t)r * TW;
    if (r == c->rank) my_off = total_frames;
    total_frames += (size_t)o[4];
    max_frames = std::max(max_frames, (size_t)o[4]);
    c->pend_frames[(size_t)r] = (size_t)o[4];
    sum_bound_run += o[5]; sum_bound_total += o[6];
    if (!o[0]) continue;
    if (!o[2]) global_uniform = false;
    if (o[3]) any_acc = true;
    if (!have_w) { w0 = (uint32_t)o[1]; have_w = true; }
    else if (w0 != (uint32_t)o[1]) global_uniform = false;
  }
  const double t_gather = now();
  // 2. Statistics: one block [hits(max_frames) | weight bits(max_frames) | edge overflow] per rank, all-gathered in
  //    place.  Ranks hold contiguous frame shards in rank order, so with every rank's per-frame hit counts and
  //    weights each rank redoes the reference's frame-ordered `sumP += (double)weight` chain (g_ri3Dc.c:151,426):
  //    bit-identical to a single-GPU run.  Host-known parts of this rank's block go up from the pinned mirror;
  //    counters of device-resident batches and the overflow counter are copied in on the device, behind the kernels
  //    that produce them.
  const size_t blk = 2 * max_frames + 1;
  rc = scratch(R * blk);
  if (rc) return rc;
  unsigned long long *mine_h = c->h_scr + (size_t)c->rank * blk, *mine_d = c->d_scr + (size_t)c->rank * blk;
  memset(mine_h, 0, sizeof(unsigned long long) * blk);
  for (size_t f = 0; f < c->hits_per_frame.size(); f++) {
    mine_h[f] = c->hits_per_frame[f];
    uint32_t wbits; memcpy(&wbits, &c->weight_per_frame[f], 4);
    mine_h[max_frames + f] = wbits;
  }
  for (auto &h : c->devhits) for (long f = 0; f < h.n; f++) mine_h[h.frame0 + f] = 0;
  // the host-known part goes up on the side stream while the binning kernels are still running
  if (!c->scr_up) GCB_CUDA(cudaEventCreateWithFlags(&c->scr_up, cudaEventDisableTiming));
  if (!c->grp_done) GCB_CUDA(cudaEventCreateWithFlags(&c->grp_done, cudaEventDisableTiming));
  GCB_CUDA(cudaMemcpyAsync(mine_d, mine_h, sizeof(unsigned long long) * blk, cudaMemcpyHostToDevice, c->comm_stream));
  GCB_CUDA(cudaEventRecord(c->scr_up, c->comm_stream));
  GCB_CUDA(cudaStreamWaitEvent(c->compute, c->scr_up, 0));
  for (auto &h : c->devhits)
    GCB_CUDA(cudaMemcpyAsync(mine_d + h.frame0, h.d, sizeof(unsigned long long) * h.n, cudaMemcpyDeviceToDevice, c->compute));
  GCB_CUDA(cudaMemcpyAsync(mine_d + 2 * max_frames, c->d_diag, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->compute));
  // 3. the grids and the statistics, one NCCL group
  if (!(global_uniform && !any_acc)) {
    rc = fold_pending(c);
    if (rc) return rc;
    rc = ensure_acc(c);
    if (rc) return rc;
    sum_bound_total += sum_bound_run;            // every rank has folded its run into the totals by now
    sum_bound_run = 0;
  }
  {
    // saturation guard across ranks: if the ranks' bounds could add up beyond the counter range, every rank
    // clamps its cells to an equal share first (the same decision everywhere: it rests on the exchanged table)
    const unsigned long long share = std::min<unsigned long long>(c->clamp_at, c->clamp_ceiling / R);
    uint32_t *grid_cnt = (global_uniform && !any_acc) ? c->d_runcnt : c->d_total;
    unsigned long long &sum_bound = (global_uniform && !any_acc) ? sum_bound_run : sum_bound_total;
    if (sum_bound > c->clamp_ceiling) {
      clamp_counts_kernel<<<elementwise_grid(c, c->ncells), 256, 0, c->compute>>>(grid_cnt, c->ncells, (uint32_t)share, c->d_diag + 3);
      c->launches++;
      GCB_CUDA(cudaGetLastError());
      sum_bound = share * R;
    }
    c->bound_run = sum_bound_run;                 // bounds of the reduced grids
    c->bound_total = sum_bound_total;
  }
  rc = ops->group_start();
  if (rc) return rc;
  if (global_uniform && !any_acc) {
    // the common case: one integer grid per rank, one weight: sum the counts
    rc = ops->allreduce(c->d_runcnt, c->d_runcnt, c->ncells, U32, c->nccl, c->compute);
  } else {
    rc = ops->allreduce(c->d_acc, c->d_acc, c->ncells, F32, c->nccl, c->compute);
    if (!rc) rc = ops->allreduce(c->d_total, c->d_total, c->ncells, U32, c->nccl, c->compute);
  }
  if (!rc) rc = ops->allgather(mine_d, c->d_scr, blk, U64, c->nccl, c->compute);
  const int rc_end = ops->group_end();
  if (rc) return rc;
  if (rc_end) return rc_end;
  if (global_uniform && !any_acc) {
    if (have_w) { memcpy(&c->w_cur, &w0, 4); c->w_first = c->w_cur; c->have_run = true; c->any = true; }
  } else {
    c->uniform = false;
  }
  // the gathered statistics come back on the side stream: the next job's kernels do not queue behind the copy
  GCB_CUDA(cudaEventRecord(c->grp_done, c->compute));
  GCB_CUDA(cudaStreamWaitEvent(c->comm_stream, c->grp_done, 0));
  GCB_CUDA(cudaMemcpyAsync(c->h_scr, c->d_scr, sizeof(unsigned long long) * R * blk, cudaMemcpyDeviceToHost, c->comm_stream));
  // The grids are reduced in stream order; the host half (per-frame counters -> sumP chain) is finished when
  // somebody asks for it (finish_reduce), so the caller can queue the density and the next job meanwhile.
  if (!c->reduce_done) GCB_CUDA(cudaEventCreateWithFlags(&c->reduce_done, cudaEventDisableTiming));
  GCB_CUDA(cudaEventRecord(c->reduce_done, c->comm_stream));
  c->reduce_pending = true; c->pend_total = total_frames; c->pend_off = my_off; c->pend_ndev = c->devhits.size();
  c->pend_nlocal = c->hits_per_frame.size(); c->pend_blk = blk;
  c->global_frames = total_frames;
  c->reduced = true;
  if (trace && c->rank == 0)
    fprintf(stderr, "[gcb allreduce] rank table (%s) %.3f ms, queueing the grid/statistics group %.3f ms\n",
            c->ctl ? "shared memory" : "NCCL, beside the binning", t_gather - t_begin, now() - t_gather);
  return GCB_OK;
}

// host half of gcb_counter_allreduce: waits for the gathered statistics and redoes the frame-ordered chain
static int finish_reduce(gcb_counter *c)
{
  if (!c->reduce_pending) return GCB_OK;
  GCB_CUDA(cudaEventSynchronize(c->reduce_done));
  c->reduce_pending = false;
  const size_t blk = c->pend_blk, max_frames = (blk - 1) / 2;
  const unsigned long long *mine = c->h_scr + (size_t)c->rank * blk;
  // local bookkeeping that gcb_counter_sync would have done, for the batches the reduce covered
  const size_t ndev = std::min(c->pend_ndev, c->devhits.size());
  for (size_t i = 0; i < ndev; i++) {
    DevHits &h = c->devhits[i];
    for (long f = 0; f < h.n; f++) c->hits_per_frame[h.frame0 + f] = mine[h.frame0 + f];
    cudaFree(h.d_box3);
  }
  c->devhits.erase(c->devhits.begin(), c->devhits.begin() + (long)ndev);
  if (c->devhits.empty()) for (auto &a : c->hit_arenas) a.used = 0;
  c->sumP_frames = c->pend_nlocal;            // frames added after the reduce continue the chain from the global sum
  // consecutive frames of equal weight are one run of identical additions: one closed-form call per run
  double sp = 0;
  unsigned long long hits = 0, run_hits = 0, edge = 0;
  uint32_t run_w = 0;
  for (size_t r = 0; r < c->pend_frames.size(); r++) {
    const unsigned long long *fr = c->h_scr + r * blk;
    for (size_t f = 0; f < c->pend_frames[r]; f++) {
      const uint32_t wbits = (uint32_t)fr[max_frames + f];
      if (wbits != run_w && run_hits) {
        float wf; memcpy(&wf, &run_w, 4);
        sp = gcb::seq_add_f64(sp, (double)wf, run_hits);
        run_hits = 0;
      }
      run_w = wbits;
      run_hits += fr[f];
      hits += fr[f];
    }
    edge += fr[2 * max_frames];
  }
  if (run_hits) { float wf; memcpy(&wf, &run_w, 4); sp = gcb::seq_add_f64(sp, (double)wf, run_hits); }
  c->hits_total = hits; c->edge_overflow = edge; c->sumP = sp;
  return GCB_OK;
}
