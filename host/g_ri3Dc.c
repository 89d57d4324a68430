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
 * .xtc, multi-frame .gro or the raw GCTRJ1 test format.  Extra options: -gpu, -pbc.
 */
#include "gcb_host.h"
#include <math.h>
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

int main(int argc, char **argv)
{
  int bdtWeight = 1, bMolecular = 0, ngroups = 1, device = 0, bPbc = 0, bHostDecode = 0;
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
    whole.name = (char *)"System"; whole.n = natoms; whole.atoms = (int *)malloc(sizeof(int) * (size_t)natoms);
    for (i = 0; i < natoms; i++) whole.atoms[i] = i;
    grp = &whole;
    gh_msg("Selected %d: '%s'\n", 0, whole.name);
  }

  /* get_timestep (g_ri3Dc.c:155-177): the time between the first two frames */
  float *x = (float *)malloc(sizeof(float) * 3 * (size_t)natoms);
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
  gcb_counter *ctr;
  CHECK(gcb_counter_create(&ctr, &tg, grp->atoms, grp->n, natoms, &copt));

  long nframes = 0;
  double T_tot = 0;
  float tlast = fr.time - dt0;                 /* g_ri3Dc.c:394 */
  const size_t tl = strlen(trxname);
  if (!bHostDecode && tl > 4 && !strcmp(trxname + tl - 4, ".xtc")) {
    /* ---- .xtc: the file goes to the GPU as it is; only the frame headers are read here.  It is read in bounded
     * chunks through one pinned buffer (a trajectory need not fit into pinnable memory); a frame cut by the end of
     * a chunk is carried over to the next one. ---- */
    gh_traj_close(trj);
    FILE *f = fopen(trxname, "rb");
    if (!f) gh_fatal("Cannot read trajectory %s", trxname);
    size_t chunk = (size_t)512 << 20;
    if (getenv("GCB_XTC_CHUNK_BYTES")) chunk = (size_t)strtoull(getenv("GCB_XTC_CHUNK_BYTES"), NULL, 10);
    if (chunk < ((size_t)natoms * 16 + 4096) * 2) chunk = ((size_t)natoms * 16 + 4096) * 2;     /* at least two frames */
    uint8_t *pin = NULL;
    CHECK(gcb_host_alloc_pinned((void **)&pin, chunk + 16));
    gcb_xtc_frame *frames = NULL;
    float *tt = NULL, *w = NULL;
    long cap = 0, i;
    size_t have = 0;                           /* bytes in the buffer: a carried partial frame + what was read */
    double t0sel = 0;
    int have0 = 0, done = 0, eof = 0;
    while (!done) {
      const size_t got = eof ? 0 : fread(pin + have, 1, chunk - have, f);
      if (got < chunk - have) eof = 1;
      have += got;
      long nall = 0;
      size_t used = 0;
      CHECK(gcb_xtc_scan(pin, have, NULL, 0, &nall, &used));
      if (nall == 0) {
        if (have > 0 && eof) gh_msg("WARNING: %zu trailing bytes of %s are not a whole frame\n", have, trxname);
        if (have >= chunk) gh_fatal("A frame of %s does not fit the %zu-byte read buffer (GCB_XTC_CHUNK_BYTES)", trxname, chunk);
        break;
      }
      if (nall > cap) {
        cap = nall;
        frames = (gcb_xtc_frame *)realloc(frames, sizeof(gcb_xtc_frame) * (size_t)cap);
        tt = (float *)realloc(tt, sizeof(float) * (size_t)cap);
        w = (float *)realloc(w, sizeof(float) * (size_t)cap);
      }
      CHECK(gcb_xtc_scan(pin, have, frames, nall, &nall, &used));
      /* -b / -e / -dt as read_next_x applies them */
      long nsel = 0;
      for (i = 0; i < nall; i++) {
        const double t = frames[i].time;
        if (ts.have_b && t < ts.b) continue;
        if (ts.have_e && t > ts.e) { done = 1; break; }
        if (ts.have_dt && ts.dt > 0) {
          if (!have0) { have0 = 1; t0sel = t; }
          const double r = fmod(t - t0sel, ts.dt);
          if (fabs(r) > 1e-4 * ts.dt && fabs(r - ts.dt) > 1e-4 * ts.dt) continue;
        }
        frames[nsel++] = frames[i];
      }
      for (i = 0; i < nsel; i++) tt[i] = frames[i].time;
      CHECK(gcb_frame_weights(tt, nsel, bdtWeight, &tlast, w, &T_tot));          /* g_ri3Dc.c:413-416; tlast, T_tot carry over */
      CHECK(gcb_counter_add_xtc(ctr, pin, have, frames, nsel, w));               /* returns when the bytes have been consumed */
      nframes += nsel;
      memmove(pin, pin + used, have - used);                                     /* the partial frame behind the last whole one */
      have -= used;
      if (eof && have == 0) break;
    }
    fclose(f);
    free(tt); free(w); free(frames);
    gcb_host_free_pinned(pin);
  } else {
  /* ---- main loop over frames: batches through the pinned slots ---- */
  const int NSLOTS = 3;
  int slot = 0;
  long cap = 0, n = 0;
  float *xs = NULL, *w, *boxes, *tbuf;
  CHECK(gcb_counter_slot(ctr, slot, &xs, &cap));
  w = (float *)malloc(sizeof(float) * (size_t)cap);
  tbuf = (float *)malloc(sizeof(float) * (size_t)cap);
  boxes = (float *)malloc(sizeof(float) * 9 * (size_t)cap);
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
  gcb_counter_destroy(ctr);
  free(x);
  if (groups) gh_ndx_free(groups, ng);
  gh_msg("\n");
  return 0;
}
