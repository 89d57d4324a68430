/*
 * g_ri3Dc -- drop-in for the reference tool of the same name (src/g_ri3Dc.c): places a 3D grid
 * into the simulation box and counts the index-group atoms per cell over a trajectory; writes
 * the time-averaged number density as an XDR grid file for a_ri3Dc.
 *
 * Same options, defaults, messages to stderr, header string and output file as the reference
 * (option table g_ri3Dc.c:265-291).  The frame loop (g_ri3Dc.c:410-429) hands batches of frames
 * to the GPU through the C ABI; the time arithmetic (dt, T_tot, weight) stays here, in the
 * reference's precision.  Differences, because Gromacs is not linked: -s takes a .gro (or is
 * omitted: the atom count then comes from the trajectory) instead of a .tpr; trajectories are
 * .xtc, multi-frame .gro or the raw GCTRJ1 test format.  Extra options: -gpu, -gpus (shard an .xtc trajectory's frames
 * over several GPUs), -pbc, -hostdecode.
 */
#include "gcb_host.h"
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const char *synopsis =
    "g_ri3Dc places a 3D grid into a simulation box and counts the number of molecules (typically water)\n"
    "in each cell over a trajectory. The rectangular grid is large enough to encompass a cylinder of given\n"
    "radius and height (-R, -z1, -z2, and a point on the pore axis -cpoint; the axis is always parallel to z).\n"
    "Without these, a cylinder is fit into the simulation box. At the end the number density, averaged over\n"
    "the trajectory, is written to the grid file; read it into a_ri3Dc for radial distribution functions,\n"
    "density plots and pore profiles.";

#define CHECK(call) do { if ((call) != GCB_OK) gh_fatal("%s", gcb_last_error()); } while (0)

/* ---- .xtc input: the file goes to the GPU(s) as it is ------------------------------------------------------------
 * Pass 1 reads only the frame headers (seeking over the payloads): offsets, sizes and times of all frames; -b/-e/-dt
 * and the weights (g_ri3Dc.c:413-416) follow from the times.  Pass 2: the selected frames are cut into one contiguous
 * shard per GPU; every GPU has its own thread, which reads its byte range in bounded chunks through a pinned buffer
 * and hands the bytes to gcb_counter_add_xtc.  Several GPUs are combined by gcb_counter_allreduce over the
 * in-process transport (the reference's manual equivalent: one g_ri3Dc run per trajectory piece + a_gridcalc). */
typedef struct { uint64_t off; uint64_t size; float time; } xtc_entry;

static uint32_t rd_be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static long xtc_index(const char *path, int natoms, xtc_entry **out)
{
  FILE *f = fopen(path, "rb");
  if (!f) gh_fatal("Cannot read trajectory %s", path);
  long n = 0, cap = 0;
  xtc_entry *e = NULL;
  uint64_t off = 0;
  uint8_t h[92];
  for (;;) {
    if (fseeko(f, (off_t)off, SEEK_SET) != 0) break;
    const size_t got = fread(h, 1, sizeof h, f);
    if (got < 56) break;
    if (rd_be32(h) != 1995u || (int)rd_be32(h + 4) != natoms) gh_fatal("%s: not an xtc frame of %d atoms at byte %llu", path, natoms, (unsigned long long)off);
    uint64_t size;
    if (natoms <= 9) size = 56 + (uint64_t)natoms * 12;
    else { if (got < 92) break; size = 92 + (((uint64_t)rd_be32(h + 88) + 3) & ~(uint64_t)3); }
    if (n == cap) { cap = cap ? 2 * cap : 1024; e = (xtc_entry *)gh_realloc(e, sizeof(xtc_entry) * (size_t)cap); }
    uint32_t tb = rd_be32(h + 12);
    e[n].off = off; e[n].size = size; memcpy(&e[n].time, &tb, 4);
    n++;
    off += size;
  }
  /* a frame cut short by the end of the file does not count */
  if (n > 0) {
    fseeko(f, 0, SEEK_END);
    const uint64_t len = (uint64_t)ftello(f);
    if (e[n - 1].off + e[n - 1].size > len) n--;
  }
  fclose(f);
  *out = e;
  return n;
}

typedef struct {
  /* in */
  const char *path;
  const gcb_tgrid *tg;
  const gh_group *grp;
  int natoms, device, pbc, rank, nranks;
  gcb_local_comm *lc;
  const xtc_entry *sel;          /* this rank's selected frames, in file order */
  const float *w;                /* their weights */
  long nsel;
  size_t chunk;
  /* out */
  gcb_counter *ctr;
  int rc;
  char err[512];
} xtc_job;

static void *xtc_worker(void *arg)
{
  xtc_job *j = (xtc_job *)arg;
  j->rc = 1;
#define WCHECK(call) do { if ((call) != GCB_OK) { snprintf(j->err, sizeof j->err, "%s", gcb_last_error()); goto done; } } while (0)
  gcb_counter_opts copt;
  memset(&copt, 0, sizeof copt);
  copt.device = j->device;
  copt.pbc = j->pbc ? GCB_PBC_RECT : GCB_PBC_NONE;
  uint8_t *pin = NULL;
  gcb_xtc_frame *frames = NULL;
  float *wsel = NULL;
  FILE *f = NULL;
  int attached = 0, joined = 0;
  WCHECK(gcb_counter_create(&j->ctr, j->tg, j->grp->atoms, j->grp->n, j->natoms, &copt));
  if (j->nranks > 1) { WCHECK(gcb_counter_comm_init_local(j->ctr, j->lc, j->rank)); attached = 1; }
  if (j->nsel > 0) {
    f = fopen(j->path, "rb");
    if (!f) { snprintf(j->err, sizeof j->err, "Cannot read trajectory %s", j->path); goto done; }
    WCHECK(gcb_host_alloc_pinned((void **)&pin, j->chunk + 16));
    long cap = 0, i = 0;
    while (i < j->nsel) {
      /* as many of the next selected frames as fit the buffer (frames skipped by -dt in between are read along) */
      long k = i;
      const uint64_t lo = j->sel[i].off;
      while (k < j->nsel && j->sel[k].off + j->sel[k].size - lo <= j->chunk) k++;
      if (k == i) { snprintf(j->err, sizeof j->err, "A frame of %s does not fit the %zu-byte read buffer (GCB_XTC_CHUNK_BYTES)", j->path, j->chunk); goto done; }
      const size_t len = (size_t)(j->sel[k - 1].off + j->sel[k - 1].size - lo);
      if (fseeko(f, (off_t)lo, SEEK_SET) != 0 || fread(pin, 1, len, f) != len) { snprintf(j->err, sizeof j->err, "Cannot read trajectory %s", j->path); goto done; }
      long nall = 0;
      WCHECK(gcb_xtc_scan(pin, len, NULL, 0, &nall, NULL));
      if (nall > cap) {
        cap = nall;
        frames = (gcb_xtc_frame *)gh_realloc(frames, sizeof(gcb_xtc_frame) * (size_t)cap);
        wsel = (float *)gh_realloc(wsel, sizeof(float) * (size_t)cap);
      }
      WCHECK(gcb_xtc_scan(pin, len, frames, nall, &nall, NULL));
      /* keep the selected ones: a frame's payload starts 92 (or 56) bytes behind its header */
      const uint64_t hdr = j->natoms <= 9 ? 56 : 92;
      long q = i, m = 0, a;
      for (a = 0; a < nall && q < k; a++)
        if (frames[a].payload_offset == j->sel[q].off - lo + hdr) { frames[m] = frames[a]; wsel[m++] = j->w[q++]; }
      if (q != k) { snprintf(j->err, sizeof j->err, "%s changed while it was read", j->path); goto done; }
      WCHECK(gcb_counter_add_xtc(j->ctr, pin, len, frames, m, wsel));      /* returns when the bytes have been consumed */
      i = k;
    }
  }
  if (j->nranks > 1) { joined = 1; WCHECK(gcb_counter_allreduce(j->ctr)); }
  j->rc = 0;
done:
  /* a rank that failed on its way still takes part in the reduce, or the other ranks would wait for it forever */
  if (j->rc && attached && !joined) gcb_counter_allreduce(j->ctr);
  if (f) fclose(f);
  free(frames); free(wsel);
  if (pin) gcb_host_free_pinned(pin);
  return NULL;
#undef WCHECK
}

int main(int argc, char **argv)
{
  int bdtWeight = 1, bMolecular = 0, ngroups = 1, device = 0, bPbc = 0, bHostDecode = 0;
  const char *gpulist = "";
  gcb_cavity cavity;
  float Delta[3] = {0.05f, 0.05f, 0.05f};
  const char *subtitle = "";
  memset(&cavity, 0, sizeof cavity);
  cavity.axis[2] = 1;

  gh_option pa[] = {
      {"-m", GH_BOOL, &bMolecular, NULL, 0},
      {"-cpoint", GH_RVEC, cavity.cpoint, "Point on the central symmetry axis of the grid", 0},
      {"-R", GH_REAL, &cavity.radius, "Maximum radius that should be contained in the grid", 0},
      {"-z1", GH_REAL, &cavity.z1, "Center grid between z1, and ...", 0},
      {"-z2", GH_REAL, &cavity.z2, "z2 (these boundaries are kept fixed!)", 0},
      {"-delta", GH_RVEC, Delta, "Spatial resolution in X, Y, and Z (in nm)", 0},
      {"-dtweight", GH_BOOL, &bdtWeight, NULL, 0},
      {"-subtitle", GH_STR, &subtitle, "Some text to add to the output graphs", 0},
      {"-ng", GH_INT, &ngroups, NULL, 0},
      {"-gpu", GH_INT, &device, "CUDA device to use", 0},
      {"-gpus", GH_STR, &gpulist, "CUDA devices to shard the frames of an .xtc trajectory over, e.g. 0,1,2,3", 0},
      {"-pbc", GH_BOOL, &bPbc, "wrap atoms once into the rectangular box before counting (the reference does not)", 0},
      {"-hostdecode", GH_BOOL, &bHostDecode, "decode .xtc frames on the host instead of on the GPU", 0},
  };
  gh_file fnm[] = {
      {"-f", "traj", ".xtc", GH_F_READ, "Trajectory: xtc gro raw", 0, NULL, NULL, 0},
      {"-s", "topol", ".gro", GH_F_OPT_READ, "Structure (atom count): gro", 0, NULL, NULL, 0},
      {"-n", "index", ".ndx", GH_F_OPT_READ, "Index file", 0, NULL, NULL, 0},
      {"-grid", "gridxdr", ".dat", GH_F_WRITE, "3D density grid", 0, NULL, NULL, 0},
  };
  const int NPA = (int)(sizeof pa / sizeof pa[0]), NFILE = (int)(sizeof fnm / sizeof fnm[0]);
  gh_timesel ts;

  gh_parse_args(argc, argv, synopsis, fnm, NFILE, pa, NPA, &ts);
  if (bMolecular) gh_fatal("Sorry, -m option not working at the moment");          /* g_ri3Dc.c:357-358 */
  if (ngroups != 1) gh_fatal("Sorry, only a single group currently allowed.");     /* :365-367 */

  /* ---- inputs: atom count, index group, time step ---- */
  const char *trxname = gh_fn(fnm, NFILE, "-f");
  gh_traj *trj = gh_traj_open(trxname);
  if (!trj) gh_fatal("Cannot read trajectory %s", trxname);
  int natoms = gh_traj_natoms(trj);
  if (gh_fn_set(fnm, NFILE, "-s")) {
    int ntop = 0;
    if (gh_top_natoms(gh_fn(fnm, NFILE, "-s"), &ntop, NULL)) gh_fatal("Cannot read structure file %s", gh_fn(fnm, NFILE, "-s"));
    if (ntop != natoms) gh_fatal("Structure file has %d atoms, trajectory has %d", ntop, natoms);
  }
  gh_group *groups = NULL, whole;
  int ng = 0;
  const gh_group *grp;
  if (gh_fn_set(fnm, NFILE, "-n")) {
    if (gh_ndx_read(gh_fn(fnm, NFILE, "-n"), &groups, &ng) || ng == 0) gh_fatal("Cannot read index file %s", gh_fn(fnm, NFILE, "-n"));
    grp = gh_ndx_select(groups, ng);
  } else {                                  /* get_index without a file: the default group of all atoms */
    int i;
    whole.name = (char *)"System"; whole.n = natoms; whole.atoms = (int *)gh_alloc(sizeof(int) * (size_t)natoms);
    for (i = 0; i < natoms; i++) whole.atoms[i] = i;
    grp = &whole;
    gh_msg("Selected %d: '%s'\n", 0, whole.name);
  }

  /* get_timestep (g_ri3Dc.c:155-177): the time between the first two frames */
  float *x = (float *)gh_alloc(sizeof(float) * 3 * (size_t)natoms);
  gh_frame_info fr, fr2;
  float dt0;
  {
    gh_traj *peek = gh_traj_open(trxname);
    if (!peek || gh_traj_next_selected(peek, &ts, &fr, x) != 1) gh_fatal("No frames in %s", trxname);
    if (gh_traj_next_selected(peek, &ts, &fr2, x) != 1) gh_fatal("Need at least two frames in %s to get the time step", trxname);
    dt0 = fr2.time - fr.time;
    gh_traj_close(peek);
  }
  if (gh_traj_next_selected(trj, &ts, &fr, x) != 1) gh_fatal("No frames in %s", trxname);

  /* ---- geometry from the FIRST frame's box (g_ri3Dc.c:392-403) ---- */
  gh_msg("\n");
  const int mask = (gh_opt_set(pa, NPA, "-cpoint") ? GCB_SET_CPOINT : 0) | (gh_opt_set(pa, NPA, "-R") ? GCB_SET_R : 0) |
                   (gh_opt_set(pa, NPA, "-z1") ? GCB_SET_Z1 : 0) | (gh_opt_set(pa, NPA, "-z2") ? GCB_SET_Z2 : 0);
  gcb_tgrid tg;
  CHECK(gcb_autoset_cavity(&cavity, fr.box, mask));
  if (gcb_setup_tgrid(&tg, &cavity, Delta) != GCB_OK) gh_fatal("FAILED: setup_tgrid()\n%s", gcb_last_error());
  gh_msg("Grid: %d x %d x %d cells, lower corner (%g, %g, %g) nm, upper corner (%g, %g, %g) nm\n", tg.mx[0], tg.mx[1], tg.mx[2],
         tg.a[0], tg.a[1], tg.a[2], tg.b[0], tg.b[1], tg.b[2]);
  if (gcb_edge_vulnerable(&tg))
    gh_msg("Note: an atom within one ulp of the upper grid edge would index past the grid; such atoms are dropped and reported.\n");

  gcb_counter_opts copt;
  memset(&copt, 0, sizeof copt);
  copt.device = device;
  copt.pbc = bPbc ? GCB_PBC_RECT : GCB_PBC_NONE;
  gcb_counter *ctr = NULL;

  long nframes = 0;
  double T_tot = 0;
  float tlast = fr.time - dt0;                 /* g_ri3Dc.c:394 */
  const size_t tl = strlen(trxname);
  xtc_job *jobs = NULL;
  int njobs = 0;
  gcb_local_comm *lc = NULL;
  if (!bHostDecode && tl > 4 && !strcmp(trxname + tl - 4, ".xtc")) {
    gh_traj_close(trj);
    /* the devices: -gpus a,b,c or the single -gpu */
    int devs[64], ndev = 0;
    if (gpulist[0]) {
      const char *q = gpulist;
      while (*q && ndev < 64) {
        char *end;
        const long d = strtol(q, &end, 10);
        if (end == q || d < 0) gh_fatal("-gpus expects a comma-separated list of device numbers, got '%s'", gpulist);
        devs[ndev++] = (int)d;
        q = *end == ',' ? end + 1 : end;
        if (*end && *end != ',') gh_fatal("-gpus expects a comma-separated list of device numbers, got '%s'", gpulist);
      }
    }
    if (ndev == 0) devs[ndev++] = device;
    /* pass 1: headers only; -b / -e / -dt as read_next_x applies them; weights from the selected times */
    xtc_entry *all = NULL;
    const long nall = xtc_index(trxname, natoms, &all);
    long i;
    double t0sel = 0;
    int have0 = 0;
    for (i = 0; i < nall; i++) {
      const double t = all[i].time;
      if (ts.have_b && t < ts.b) continue;
      if (ts.have_e && t > ts.e) break;
      if (ts.have_dt && ts.dt > 0) {
        if (!have0) { have0 = 1; t0sel = t; }
        const double r = fmod(t - t0sel, ts.dt);
        if (fabs(r) > 1e-4 * ts.dt && fabs(r - ts.dt) > 1e-4 * ts.dt) continue;
      }
      all[nframes++] = all[i];
    }
    float *tt = (float *)gh_alloc(sizeof(float) * (size_t)(nframes > 0 ? nframes : 1));
    float *w = (float *)gh_alloc(sizeof(float) * (size_t)(nframes > 0 ? nframes : 1));
    for (i = 0; i < nframes; i++) tt[i] = all[i].time;
    CHECK(gcb_frame_weights(tt, nframes, bdtWeight, &tlast, w, &T_tot));        /* g_ri3Dc.c:413-416 */
    /* pass 2: one contiguous shard of the selected frames per device, one thread per device */
    size_t chunk = (size_t)512 << 20;
    if (getenv("GCB_XTC_CHUNK_BYTES")) chunk = (size_t)strtoull(getenv("GCB_XTC_CHUNK_BYTES"), NULL, 10);
    if (chunk < ((size_t)natoms * 16 + 4096) * 2) chunk = ((size_t)natoms * 16 + 4096) * 2;     /* at least two frames */
    njobs = ndev;
    jobs = (xtc_job *)gh_zalloc((size_t)njobs, sizeof(xtc_job));
    if (njobs > 1) CHECK(gcb_local_comm_create(&lc, njobs));
    pthread_t *th = (pthread_t *)gh_zalloc((size_t)njobs, sizeof(pthread_t));
    int r;
    for (r = 0; r < njobs; r++) {
      const long f0 = nframes * r / njobs, f1 = nframes * (r + 1) / njobs;
      jobs[r].path = trxname; jobs[r].tg = &tg; jobs[r].grp = grp; jobs[r].natoms = natoms; jobs[r].device = devs[r];
      jobs[r].pbc = bPbc; jobs[r].rank = r; jobs[r].nranks = njobs; jobs[r].lc = lc;
      jobs[r].sel = all + f0; jobs[r].w = w + f0; jobs[r].nsel = f1 - f0; jobs[r].chunk = chunk;
    }
    if (njobs > 1) gh_msg("Sharding %ld frames over %d GPUs\n", nframes, njobs);
    for (r = 1; r < njobs; r++) if (pthread_create(&th[r], NULL, xtc_worker, &jobs[r]) != 0) gh_fatal("cannot start a worker thread");
    xtc_worker(&jobs[0]);
    for (r = 1; r < njobs; r++) pthread_join(th[r], NULL);
    for (r = 0; r < njobs; r++) if (jobs[r].rc) gh_fatal("%s", jobs[r].err);
    ctr = jobs[0].ctr;                          /* after the reduce every rank holds the whole result */
    free(th); free(tt); free(w); free(all);
  } else {
  /* ---- main loop over frames: batches through the pinned slots ---- */
  if (gpulist[0]) gh_msg("Note: -gpus shards .xtc trajectories decoded on the GPU; this run uses device %d only\n", device);
  CHECK(gcb_counter_create(&ctr, &tg, grp->atoms, grp->n, natoms, &copt));
  const int NSLOTS = 3;
  int slot = 0;
  long cap = 0, n = 0;
  float *xs = NULL, *w, *boxes, *tbuf;
  CHECK(gcb_counter_slot(ctr, slot, &xs, &cap));
  w = (float *)gh_alloc(sizeof(float) * (size_t)cap);
  tbuf = (float *)gh_alloc(sizeof(float) * (size_t)cap);
  boxes = (float *)gh_alloc(sizeof(float) * 9 * (size_t)cap);
  int more = 1;
  while (more) {
    memcpy(xs + (size_t)n * natoms * 3, x, sizeof(float) * 3 * (size_t)natoms);
    memcpy(boxes + 9 * n, fr.box, sizeof(float) * 9);
    tbuf[n++] = fr.time;
    nframes++;
    more = gh_traj_next_selected(trj, &ts, &fr, x) == 1;
    if (n == cap || !more) {
      /* dt = t - tlast; T_tot += dt; weight = dtweight ? dt : 1   (g_ri3Dc.c:413-416) */
      CHECK(gcb_frame_weights(tbuf, n, bdtWeight, &tlast, w, &T_tot));
      CHECK(gcb_counter_submit_slot(ctr, slot, n, w, bPbc ? boxes : NULL));
      slot = (slot + 1) % NSLOTS;
      n = 0;
      if (more) CHECK(gcb_counter_slot(ctr, slot, &xs, &cap));
    }
  }
  gh_traj_close(trj);
  free(w); free(tbuf); free(boxes);
  }

  gcb_stats st;
  CHECK(gcb_counter_stats(ctr, &st));
  gh_msg("Read %ld frames, %lu of %lu atom-frames inside the grid\n", nframes, (unsigned long)st.hits, (unsigned long)st.atom_frames);
  if (st.edge_overflow) gh_msg("WARNING: %lu atoms on the upper grid edge were dropped\n", (unsigned long)st.edge_overflow);

  /* ---- normalisation (g_ri3Dc.c:451-460), header (:462-473), file (:476-478) ---- */
  /* the file image is produced on the device (K2 + grid3_serialise's reordering + XDR byte order) */
  gcb_tgrid tgout = tg;
  tgout.tweight = (float)T_tot;
  char header[GCB_HEADER_MAX];
  CHECK(gcb_header(header, subtitle, T_tot, grp->name, &cavity, &tgout, st.sumP, trxname, bdtWeight));
  gh_msg("Writing 3D density grid with header\n%s\n", header);
  if (gcb_counter_grid_write(ctr, T_tot, gh_fn(fnm, NFILE, "-grid"), header, &tgout) != GCB_OK)
    gh_msg("Writing the 3D grid failed.\n%s\n", gcb_last_error());
  if (jobs) {
    int r;
    for (r = 0; r < njobs; r++) gcb_counter_destroy(jobs[r].ctr);
    if (lc) gcb_local_comm_destroy(lc);
    free(jobs);
  } else {
    gcb_counter_destroy(ctr);
  }
  free(x);
  if (groups) gh_ndx_free(groups, ng);
  gh_msg("\n");
  return 0;
}
