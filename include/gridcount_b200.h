/*
 * gridcount_b200.h -- C ABI of libgridcount_b200.so
 *
 * B200-native (sm_100a) replacement for the one data-parallel hot path of
 * orbeckst/gridcount: g_ri3Dc's binning of index-group atom positions into a
 * 3D occupancy grid over a trajectory, the density normalisation, the XDR grid
 * file, and a_ri3Dc's reductions over the finished grid.
 *
 * The reference has no FFI: its seam is function level inside two main()s.
 * Each entry point below names the reference code it replaces (file:line under
 * the reference tree).  Plain C types only; every function returns 0 on
 * success or a negative gcb_status, and gcb_last_error() gives the message.
 * The library never calls exit(); the host tool turns errors into
 * gmx_fatal-style exits.  There is no CPU fallback for the GPU entry points:
 * without a CUDA device they fail with GCB_ERR_CUDA.
 *
 * Memory layouts (all little-endian host floats):
 *   frames      x[frame][atom][3]             rvec arrays as read_next_x fills them
 *                                             (g_ri3Dc.c:392,428)
 *   grid        z fastest, cell (i,j,k) at (i*my + j)*mz + k, the contiguous block
 *               behind t_tgrid.grid[0][0]     (utilgmx.c:57-69)
 *   file order  x fastest, then y, then z     (xdr_grid.c:113-126, plt.c:69-75)
 *   2D outputs  row major [n0][n1] as grid2_alloc (utilgmx.c:76-110)
 */
#ifndef GRIDCOUNT_B200_H
#define GRIDCOUNT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCB_ABI_VERSION 1
#define GCB_HEADER_MAX 256            /* HEADER_MAX, xdr_grid.h:35 */
#define GCB_GRID_FF_VERSION 2         /* GRID_FF_VERSION, xdr_grid.h:34 */
#define GCB_NGRID_MAX (512u * 512u * 512u)  /* NGRID_MAX, xdr_grid.h:36 */
#define GCB_GMX_REAL_EPS 5.96046448E-08     /* GMX_REAL_EPS of single-precision Gromacs 4.5 (grid3D.c:143,146) */

typedef enum {
  GCB_OK = 0,
  GCB_ERR_ARG = -1,        /* bad argument */
  GCB_ERR_CUDA = -2,       /* CUDA runtime/driver failure, or no device */
  GCB_ERR_NOMEM = -3,
  GCB_ERR_IO = -4,         /* file open/read/write */
  GCB_ERR_FORMAT = -5,     /* not a version-2, 3D grid file (xdr_grid.c:28-34, grid3D.c:81-83) */
  GCB_ERR_STATE = -6,      /* call order (e.g. results requested before sync) */
  GCB_ERR_NCCL = -7,
  GCB_ERR_RANGE = -8       /* crop outside the grid (asserts at a_ri3Dc.c:191-192), z2 <= z1 (:163-165) */
} gcb_status;

const char *gcb_last_error(void);
int gcb_abi_version(void);

/* ------------------------------------------------------------------------
 * Geometry (host).  Replaces t_cavity (count.h:71-77), t_tgrid without its
 * data pointer (grid3D.h:63-71), autoset_cavity (count.c:25-64), setup_corners
 * + setup_tgrid (grid3D.c:125-203), DeltaV (grid3D.c:108-110), update_tgrid /
 * tgrid2cavity (a_ri3Dc.c:86-107).  Same f32/f64 operation order.
 * ---------------------------------------------------------------------- */
typedef struct { float axis[3], cpoint[3], radius, z1, z2, vol; } gcb_cavity;
typedef struct { int mx[3]; float a[3], b[3], delta[3]; float tweight; } gcb_tgrid;

enum { GCB_SET_CPOINT = 1, GCB_SET_R = 2, GCB_SET_Z1 = 4, GCB_SET_Z2 = 8 };   /* which options the user gave */

int gcb_autoset_cavity(gcb_cavity *cav, const float box[9], int setmask);
int gcb_setup_tgrid(gcb_tgrid *tg, gcb_cavity *cav, const float delta[3]);
double gcb_delta_v(const gcb_tgrid *tg);
void gcb_update_tgrid(gcb_tgrid *tg);
void gcb_tgrid2cavity(const gcb_tgrid *tg, gcb_cavity *cav);
/* 1 if a coordinate one ulp below b[d] would bin to mx[d] (the reference then
 * writes out of bounds, g_ri3Dc.c:146-149); such hits are dropped and counted
 * in gcb_stats.edge_overflow. */
int gcb_edge_vulnerable(const gcb_tgrid *tg);

/* Frame weights of the frame loop (g_ri3Dc.c:390-394, 413-416): dt = t - tlast
 * in f32 with tlast0 = t[0] - (t[1]-t[0]); T_tot += dt in f64; weight = dt or 1.
 * `tlast` in/out lets a caller stream frames in pieces (NULL: bootstrap from t). */
int gcb_frame_weights(const float *t, long nframes, int dtweight, float *tlast,
                      float *weight_out, double *T_tot_inout);

/* ------------------------------------------------------------------------
 * Binning on the GPU.  Replaces gridcount() (g_ri3Dc.c:139-152) and the atom
 * loop g_ri3Dc.c:425-426 over many frames.
 * ---------------------------------------------------------------------- */
typedef struct gcb_counter gcb_counter;

enum {
  GCB_PBC_NONE = 0,        /* reference behaviour: atoms outside [a,b) are dropped (g_ri3Dc.c:145) */
  GCB_PBC_RECT = 1         /* opt-in single-image rectangular wrap as count.c:88-92 before binning */
};
enum {                     /* kernel selection (GCB_KERNEL_AUTO picks by index-group density and regularity) */
  GCB_KERNEL_AUTO = 0, GCB_KERNEL_GATHER = 1, GCB_KERNEL_STREAM = 2,
  GCB_KERNEL_STREAM_INDEXED = 3,  /* streaming kernel without the arithmetic-progression shortcut */
  /* 4 was GCB_KERNEL_ROUTED (keys routed to buckets, counted by a second kernel): removed in round 2, it lost to
     both kernels below on every shape */
  GCB_KERNEL_PACKED = 5           /* streaming kernel with non-returning REDs into packed 8- or 4-bit L2-resident
                                     counters, verified (digit sum == hits) and folded into the uint32 grid once per
                                     round; AUTO picks it for grids whose uint32 cells exceed the L2 */
  ,
  GCB_KERNEL_PRIVATE = 6          /* streaming kernel, one CTA per SM, every CTA binning with shared-memory atomics
                                     into its own 8- or 4-bit copy of the whole grid (grids up to ~118 000 / ~236 000
                                     cells: 60^3); the copies rest in global memory between launches and are folded into
                                     the uint32 grid when it is read or a lane could fill up; verified by digit sum
                                     per CTA and launch.  AUTO picks it for such grids when a launch is large
                                     enough to pay for loading and storing the copies; on larger grids this value
                                     means GCB_KERNEL_STREAM */
};

typedef struct {
  int device;              /* CUDA device ordinal */
  int pbc;                 /* GCB_PBC_* */
  int kernel;              /* GCB_KERNEL_* */
  int nslots;              /* staging slots / copy streams for host frames (0: default 3) */
  long frames_per_batch;   /* frames per staged batch (0: sized to ~32 MiB) */
} gcb_counter_opts;

typedef struct {
  uint64_t frames;         /* frames binned so far */
  uint64_t atom_frames;    /* frames * gnx */
  uint64_t hits;           /* atom-frames that landed in the grid */
  uint64_t edge_overflow;  /* dropped: quotient >= mx or NaN coordinate (see gcb_edge_vulnerable) */
  double sumP;             /* Sum over hits of (double)weight in frame order (g_ri3Dc.c:151,426) */
  int uniform_weight;      /* 1 while every frame so far had the same weight */
  float weight;            /* that weight */
  int64_t kernel_launches; /* CUDA kernels launched by this counter so far */
  uint64_t packed_rounds;  /* rounds queued on the packed-counter path (GCB_KERNEL_PACKED) ... */
  uint64_t packed_redone;  /* ... of which the direct uint32 kernel binned (lane overflow, or backing off after one; this rank) */
  int saturated;           /* 1 if some cell reached the saturation guard (2^27 hits and the counter range was about
                              to run out): its count is a lower bound; occupancy and density are still the reference's
                              bits, because the reference's f32 accumulator stops moving long before (this rank) */
  uint64_t private_launches; /* launches on the private shared-memory path (GCB_KERNEL_PRIVATE) ... */
  uint64_t private_redone;   /* ... in which some CTA overflowed a lane and binned its tiles again directly (this rank) */
  int packed_bits;           /* lane width the packed path is using now: 8 or 4 (a nibble grid may have fallen back to
                                bytes with short rounds after its rounds kept overflowing), 0: no packed path */
} gcb_stats;

int gcb_counter_create(gcb_counter **out, const gcb_tgrid *tg, const int *gindex, int gnx,
                       int natoms, const gcb_counter_opts *opts);
void gcb_counter_destroy(gcb_counter *c);
int gcb_counter_reset(gcb_counter *c);

/* Host frames x[nframes][natoms][3]; weight[f] is the reference's per-frame
 * `weight` (NULL: 1); box[f][9] is only read with GCB_PBC_RECT.  Frames are
 * staged through pinned buffers in batches on CUDA streams and binned
 * asynchronously; the call returns when all of x has been consumed: the host threads have copied what they pack into
 * the staging buffers (gcb_host_staging below), and the last DMA that reads pinned x in place has finished, so the
 * caller may refill x at once.
 * Frames whose weight differs from frame to frame in short runs are accumulated with global f32 atomics
 * (RED.ADD.F32), which flush subnormal sums to zero whatever -ftz says; weights are time steps, many orders of
 * magnitude above 1e-38, so this never shows. */
int gcb_counter_add_frames(gcb_counter *c, const float *x, long nframes,
                           const float *weight, const float *box);
/* How host frames reach the device.  PACKED staging: a pool of host threads copies the index group's atoms of a
 * batch -- what the reference's loop reads, x[index[i]] (g_ri3Dc.c:425-426) -- into a pinned slot, so that 12 B per
 * counted atom cross PCIe instead of 12 B per atom of the frame (a third of the bytes for water oxygens), and the
 * kernels bin frames in which every atom counts.  Counts, per-frame hits and everything derived from them are the
 * same bits either way.  AUTO (the default; environment GCB_HOST_PACK=0/1 overrides it at create):
 *   - pageable x is always packed (it has to be copied into a slot anyway);
 *   - pinned x goes down TWO LANES at once: the host threads pack batch after batch, and whenever the copies already
 *     queued would not keep the link busy until the next packed batch is ready, a batch is DMA-ed straight out of the
 *     caller's memory to fill the gap.  Host threads and link are then both busy all the time, whatever their speeds
 *     (throughput ~ (link + 2/3 packing rate) / 36 B for water oxygens); a host that packs slowly mostly DMAs in place.
 *     Only calls of more than one batch, and only counters created with nslots = 0 (the pipeline is then 6 slots deep).
 *     Whether the two lanes beat plain DMA depends on the host as well (a packed atom costs its memory system
 *     36 B read + 12 B written + 12 B read by the DMA engine, an atom DMA-ed in place 36 B; many ranks may share
 *     it), so the first calls time both ways, about 40 ms of calls each, and the counter keeps the two lanes only
 *     if they moved frames at least 5 % faster.
 * NEVER is the staging of round 1 (memcpy of pageable x into the slot, pinned x DMA-ed in place).
 * GCB_PACK_THREADS sets the pool's size (default: the CPUs the process may run on, at most 16; the ranks of a
 * multi-GPU job share them). */
enum { GCB_HOST_PACK_AUTO = -1, GCB_HOST_PACK_NEVER = 0, GCB_HOST_PACK_ALWAYS = 1 };
typedef struct {
  int mode;                 /* GCB_HOST_PACK_* */
  int threads;              /* host threads a packing job uses */
  uint64_t h2d_bytes;       /* bytes of frames queued host -> device so far (not cleared by gcb_counter_reset) */
  uint64_t frames_packed;   /* frames that were staged packed so far ... */
  uint64_t frames_in_place; /* ... and frames DMA-ed straight out of the caller's pinned memory */
  double pack_bytes_per_s;  /* measured packing rate in bytes of source frames (moving average; 0: nothing packed yet) */
  int lanes;                /* AUTO with pinned x: 1 the two lanes won the trial, 0 DMA in place did, -1 still measuring */
  double two_lane_frames_per_s, in_place_frames_per_s;   /* what the trial measured (0: not measured yet) */
} gcb_host_staging;
int gcb_counter_set_host_pack(gcb_counter *c, int mode);
int gcb_counter_host_staging(gcb_counter *c, gcb_host_staging *out);
/* The packing step by itself (pure host code, no device needed): out[f][i][:] = x[f][gindex[i]][:].
 * threads <= 0: the default above. */
int gcb_pack_frames(const float *x, long nframes, int natoms, const int *gindex, int gnx, float *out, int threads);
/* Same for frames already resident in device memory (16-byte aligned). */
int gcb_counter_add_device_frames(gcb_counter *c, const float *x_dev, long nframes,
                                  const float *weight, const float *box);
/* Zero-copy producer interface: fill a pinned slot, then submit it. */
int gcb_counter_slot(gcb_counter *c, int slot, float **x_pinned, long *capacity_frames);
int gcb_counter_submit_slot(gcb_counter *c, int slot, long nframes, const float *weight, const float *box);
int gcb_counter_sync(gcb_counter *c);

int gcb_counter_stats(gcb_counter *c, gcb_stats *out);
/* per-frame in-grid hit counts, in submission order */
int gcb_counter_hits_per_frame(gcb_counter *c, uint64_t *hits, long capacity, long *nframes_out);
/* raw integer hit counts per cell (z fastest) */
int gcb_counter_counts(gcb_counter *c, uint32_t *counts_zfast);
/* the reference's f32 accumulator grid[i][j][k] before normalisation
 * (the value of its `+= weight` chain, g_ri3Dc.c:149), z fastest */
int gcb_counter_occupancy(gcb_counter *c, float *occ_zfast);
/* Normalisation g_ri3Dc.c:451-460: grid / ((double)(float)dV * T_tot); either
 * output may be NULL.  Sets tg_out->tweight = (float)T_tot when tg_out != NULL. */
int gcb_counter_density(gcb_counter *c, double T_tot, float *grid_zfast, float *grid_xfast,
                        gcb_tgrid *tg_out);
/* device pointers for callers that keep results on the GPU (valid until the next
 * call that produces the same kind of result, or destroy); the work is enqueued on
 * the counter's stream, gcb_counter_sync waits for it */
int gcb_counter_device_counts(gcb_counter *c, uint32_t **counts_dev);
int gcb_counter_device_density(gcb_counter *c, double T_tot, float **grid_zfast_dev);
/* CUDA events on the stream the binning kernels are launched on.  _event_record
 * enqueues event number `id` (0..GCB_MAX_EVENTS-1) without blocking;
 * _event_elapsed waits for event `id1` and returns the device time between the
 * two in milliseconds.  timer_start/stop are the pair (0, 1) with the wait. */
#define GCB_MAX_EVENTS 4096
int gcb_counter_event_record(gcb_counter *c, int id);
int gcb_counter_event_elapsed(gcb_counter *c, int id0, int id1, float *ms);
int gcb_counter_timer_start(gcb_counter *c);
int gcb_counter_timer_stop(gcb_counter *c, float *ms);
/* Per-launch timing of K1's dominant kernel (the streaming binning kernel that replaces the atom loop
 * g_ri3Dc.c:425-426): while enabled, every launch is bracketed by a CUDA event pair on the launching stream.
 * _profile_read waits for the stream and returns the summed launch durations, the number of launches and the
 * atom-frames those launches covered (everything but the sub-tile tails).  Measurement aid: bench.py's roofline. */
int gcb_counter_profile(gcb_counter *c, int enable);
int gcb_counter_profile_read(gcb_counter *c, double *k1_ms, long *k1_launches, double *k1_atom_frames);
/* The same split by what the launch was: [0] the direct uint32 kernel, [1] the packed-counter pass of a round,
 * [2] the guarded uint32 pass behind it (returns at once unless the round failed verification or was skipped),
 * [3] everything else K1 queued for those frames (verify/apply passes, gather tails, the fused xtc kernel),
 * [4] the streaming kernel with per-CTA private shared-memory grids. */
enum { GCB_PROF_DIRECT = 0, GCB_PROF_PACKED = 1, GCB_PROF_REDO = 2, GCB_PROF_OTHER = 3, GCB_PROF_PRIVATE = 4, GCB_PROF_KINDS = 5 };
int gcb_counter_profile_kinds(gcb_counter *c, double ms[GCB_PROF_KINDS], long launches[GCB_PROF_KINDS], double atom_frames[GCB_PROF_KINDS]);

/* Frame-sharded multi-GPU: one counter per rank; integer sum of the private
 * grids and of the statistics over NCCL (the reference's manual equivalent is
 * a_gridcalc.c:180-189).  After the call every rank holds the global result.
 * gcb_counter_allreduce returns once the collectives are queued behind the binning
 * kernels (device-side consumers such as gcb_counter_density are ordered on the
 * stream); the gathered per-frame statistics are waited for by the first getter that
 * needs them (gcb_counter_stats, _hits_per_frame, _sync).  Before the sum the cells are clamped
 * if the ranks' bounds could add up to 2^32 (see gcb_stats.saturated). */
int gcb_nccl_unique_id(char id_out[128]);
int gcb_counter_comm_init(gcb_counter *c, const char id[128], int rank, int nranks);
int gcb_counter_allreduce(gcb_counter *c);
/* With frame weights that differ between frames the reduce adds the ranks' f32 accumulators (ncclSum), which is
 * not the reference's frame-ordered `grid += weight` chain (g_ri3Dc.c:149,413-416): occupancy and density then
 * agree with a single-process run to f32 accumulation error (the order of an f32 sum), not bit for bit; counts,
 * per-frame hits and sumP stay exact.  One weight for the whole job (constant dt) is exact for any rank count.
 *
 * The control plane under gcb_counter_comm_init: what the ranks must agree on before the grid reduce (one weight
 * or not, frame counts, saturation bounds) is host-side knowledge and is exchanged through a shared-memory
 * all-gather between the ranks of one node, not through a second NCCL collective.  Exposed for tests: */
typedef struct gcb_ctl gcb_ctl;
int gcb_ctl_open(gcb_ctl **out, const char id[128], int rank, int nranks, int words, int timeout_ms);
int gcb_ctl_allgather(gcb_ctl *h, const uint64_t *mine /* [words] */, uint64_t *all /* [nranks][words] */,
                      int timeout_ms /* <= 0: wait for the other ranks as long as it takes */);
void gcb_ctl_close(gcb_ctl *h);
/* The same reduce between counters of ONE process (one host thread per counter, on one GPU or several): an
 * in-process transport behind the same protocol, for host programs that drive their GPUs from threads and for
 * exercising the frame-sharded path on a single-GPU box (NCCL refuses two ranks on one device).
 * gcb_counter_comm_init_local is called once per rank from `nranks` threads at the same time, gcb_counter_allreduce
 * likewise; gcb_local_comm_destroy after the counters are destroyed. */
typedef struct gcb_local_comm gcb_local_comm;
int gcb_local_comm_create(gcb_local_comm **out, int nranks);
int gcb_counter_comm_init_local(gcb_counter *c, gcb_local_comm *lc, int rank);
void gcb_local_comm_destroy(gcb_local_comm *lc);

/* ------------------------------------------------------------------------
 * XTC trajectories decoded on the GPU.  Replaces, for .xtc input, the frames that
 * read_first_x / read_next_x deliver to the frame loop (g_ri3Dc.c:392,428; the decoder
 * itself is Gromacs' xdr3dfcoord, not part of the reference tree).  The host parses the
 * frame headers only; compressed payloads cross PCIe as they are (about 5 bytes per atom instead of 12).
 * On the GPU one warp per frame walks the stream's control bits and drops a checkpoint every 64 atoms; one
 * thread per checkpoint then decodes its atoms -- and, in gcb_counter_add_xtc, bins them on the spot (no
 * coordinates are stored) whenever a batch is one run of integer counts.  PARITY WITH GROMACS UNPINNED (the
 * library is not in the reference tree): pinned bit for bit to the host codec of host/trajio.c.
 * ---------------------------------------------------------------------- */
typedef struct {
  uint64_t payload_offset;  /* byte offset of the compressed coordinates in the scanned buffer */
  uint32_t payload_bytes;   /* 0 for frames of <= 9 atoms, which are stored as plain floats */
  int natoms, step;
  float time, box[9], precision;
  int minint[3], maxint[3], smallidx;
} gcb_xtc_frame;
/* frames == NULL counts; stops quietly at an incomplete trailing frame; *consumed = bytes of whole frames */
int gcb_xtc_scan(const uint8_t *buf, size_t len, gcb_xtc_frame *frames, long cap, long *nframes_out,
                 size_t *consumed);
/* host bytes -> x_dev[nframes][natoms][3] on the device (synchronous) */
int gcb_xtc_decode_device(const uint8_t *buf, size_t len, const gcb_xtc_frame *frames, long nframes,
                          int natoms, float *x_dev, int device);
/* decode + bin: the xtc counterpart of gcb_counter_add_frames.  `buf` is host memory (pinned memory is
 * copied asynchronously); weight[f] as in gcb_counter_add_frames; with GCB_PBC_RECT the boxes of the
 * frame headers are used.  The call returns when all batches have been consumed.  A frame whose compressed
 * stream is damaged fails the call with GCB_ERR_FORMAT and bins nothing itself, but frames of the same call
 * before it have been binned: reset the counter. */
int gcb_counter_add_xtc(gcb_counter *c, const uint8_t *buf, size_t len, const gcb_xtc_frame *frames,
                        long nframes, const float *weight);

/* ------------------------------------------------------------------------
 * Grid files and header.  Replaces the header string (g_ri3Dc.c:462-473),
 * grid_write / grid_read (grid3D.c:22-106), xdr_grid_v2 and the
 * (un)serialisers (xdr_grid.c:89-158), density_write_plt / _ascii
 * (plt.c:42-99).  Byte-exact with the reference's files.
 * ---------------------------------------------------------------------- */
int gcb_header(char out[GCB_HEADER_MAX], const char *subtitle, double T_tot, const char *grpname,
               const gcb_cavity *cav, const gcb_tgrid *tg, double sumP, const char *trxname,
               int dtweight);
size_t gcb_xdr_size(const gcb_tgrid *tg, const char *header);
/* grid_xfast: cells in file order (what gcb_counter_density emits), or NULL with grid_zfast given */
int gcb_xdr_encode(const gcb_tgrid *tg, const float *grid_zfast, const float *grid_xfast,
                   const char *header, uint8_t *out, size_t cap, size_t *written);
int gcb_grid_write(const char *path, const gcb_tgrid *tg, const float *grid_zfast,
                   const float *grid_xfast, const char *header);
/* *grid_zfast is malloc'ed by the library (free with gcb_free); b is rebuilt as a + mx*delta */
int gcb_grid_read(const char *path, gcb_tgrid *tg, char header[GCB_HEADER_MAX + 1], float **grid_zfast);
int gcb_plt_write(const char *path, const gcb_tgrid *tg, const float *grid_zfast, float unit);
/* The same files for a counter's result, encoded ON THE DEVICE: K2, the z-fastest -> x-fastest reordering of
 * grid3_serialise (xdr_grid.c:113-126) and the XDR big-endian byte order (or plt's `norm * grid`, plt.c:69-75)
 * are one pass over the grid in HBM; the host writes only the prologue in front of the cells.  `out` receives
 * the complete file image (gcb_xdr_size bytes; 44 + 4*cells for plt); pinned memory is filled by DMA.
 * tweight is (float)T_tot as at g_ri3Dc.c:460. */
int gcb_counter_xdr(gcb_counter *c, double T_tot, const char *header, uint8_t *out, size_t cap,
                    size_t *written, gcb_tgrid *tg_out);
int gcb_counter_grid_write(gcb_counter *c, double T_tot, const char *path, const char *header,
                           gcb_tgrid *tg_out);
int gcb_counter_plt(gcb_counter *c, double T_tot, float unit, uint8_t *out, size_t cap, size_t *written);
int gcb_ascii_write(const char *path, const gcb_tgrid *tg, const float *grid_zfast, float unit);
void gcb_free(void *p);

/* ------------------------------------------------------------------------
 * Analysis on the GPU.  Replaces readjust_tgrid (a_ri3Dc.c:133-201) and the
 * reductions of a_ri3Dc's main (a_ri3Dc.c:488-762).  Every output is the
 * reference's sequential f32 (sumP: f64) chain in its k,j,i order, bit for bit;
 * long chains are evaluated in parallel by an exact scan of the rounding
 * recurrence (csrc/seqsum.cuh), which needs finite cells >= 0 -- grids with
 * negative or non-finite cells take one-thread-per-chain kernels instead.
 * ---------------------------------------------------------------------- */
enum { GCB_UNIT_UNITY, GCB_UNIT_SPC, GCB_UNIT_MOLAR, GCB_UNIT_ANG, GCB_UNIT_VOXEL, GCB_UNIT_NR };  /* count.h:28 */

typedef struct {
  int unit;                /* GCB_UNIT_*  (-unit) */
  float min_occ;           /* -minocc, in the chosen unit */
  int rad_ba_nr, rad_ba_step;   /* hidden -radnr / -radstep (0, 1 by default) */
  int setmask;             /* GCB_SET_R|Z1|Z2: crop as readjust_tgrid with the values below */
  float radius, z1, z2;
  int device;
} gcb_analysis_opts;

typedef struct {
  gcb_tgrid tg;            /* the (possibly cropped) grid that was analysed */
  gcb_cavity geometry;     /* tgrid2cavity of it */
  float DeltaR; double dV; int iradNR; float unit_value;
  double sumP, sumH;       /* a_ri3Dc.c:499-512 */
  float volume, reff, average, height; int ivol;    /* a_ri3Dc.c:759-762 */
  float *grid;             /* the cropped grid, z fastest [mx][my][mz]; NULL when the analysis did not crop
                              (no -R/-z1/-z2): the grid that was passed in is the grid that was analysed */
  float *xyp, *xzp, *yzp, *hxyp;        /* [mx][my] [mx][mz] [my][mz] [mx][my] */
  float *zdf, *lzdf, *profile;          /* [mz][2] */
  float *rweight;                       /* [iradNR] */
  float *rdf, *hrdf;                    /* [iradNR][2] */
  float *rzp, *hrzp;                    /* [iradNR][mz] */
  int64_t kernel_launches;
} gcb_analysis;

float gcb_dens_unit(int unit, const gcb_tgrid *tg);      /* DensUnit[] (count.c:21-22, a_ri3Dc.c:516) */
int gcb_analyse(const gcb_tgrid *tg, const float *grid_zfast, const gcb_analysis_opts *opts,
                gcb_analysis *out);
/* the same over a grid that already sits in device memory on opts->device (gcb_counter_device_density):
 * K1 -> K2 -> K3 without the grid crossing PCIe. */
int gcb_analyse_device(const gcb_tgrid *tg, const float *grid_zfast_dev, const gcb_analysis_opts *opts,
                       gcb_analysis *out);
void gcb_analysis_free(gcb_analysis *a);
/* gcb_analyse keeps, per device, what a call needs again at once: its device buffers (up to 1.5 GB: a 500^3 grid and
 * its x-fastest copy), four streams, pinned staging and the per-(i,j) tables of the last geometry.  This gives
 * them back. */
void gcb_analyse_release_cache(int device);

/* a_gridcalc.c:132-189: time-weighted average of n grids of identical shape */
int gcb_gridcalc(int ngrids, const float *const *grids_zfast, const float *tweights,
                 const gcb_tgrid *tg, float *avg_zfast, float *tweight_out, int device);

/* ------------------------------------------------------------------------
 * Synthetic trajectories (benchmarks/tests): counter-based generator whose
 * bits are reproducible on the CPU.  Writes device memory.
 * ---------------------------------------------------------------------- */
enum { GCB_GEN_UNIFORM = 0, GCB_GEN_QUANTISED = 1, GCB_GEN_PORE = 2, GCB_GEN_CLUSTER = 3 };
typedef struct {
  uint64_t seed; int mode; float L[3];
  float slab_lo, slab_hi, pore_r; int nsites; float sigma; int box_period; float box_amp;
} gcb_gen;
int gcb_gen_device_frames(const gcb_gen *p, long frame0, long nframes, int natoms, float *x_dev, int device);

/* device memory helpers so that non-CUDA hosts (C tools, ctypes) can hold frames on the GPU */
int gcb_device_malloc(void **p, size_t bytes, int device);
int gcb_device_free(void *p, int device);
int gcb_host_alloc_pinned(void **p, size_t bytes);
int gcb_host_free_pinned(void *p);
int gcb_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, int device);
int gcb_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, int device);
int gcb_device_count(void);

/* ------------------------------------------------------------------------
 * Memory-system self-test: the measured denominators of K1's roofline next to the HBM copy bandwidth
 * ("the L2-atomic peak ... the builder must measure it").  mix 0: random-address
 * red.global.add.u32 over a window of window_bytes (the L2-atomic peak for that window); mix 1 / 2: a coalesced
 * evict-first read stream with one such RED per 36 / 12 streamed bytes (one atom in three counted / every atom
 * counted: the traffic K1 sends the L2, without K1); mix 3: the stream alone.  stream_bytes 0: 2 GiB.
 * ---------------------------------------------------------------------- */
int gcb_selftest_l2(int device, size_t window_bytes, int mix, size_t stream_bytes, double *red_gops, double *read_gbs);

#ifdef __cplusplus
}
#endif
#endif
