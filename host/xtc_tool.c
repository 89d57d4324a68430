/*
 * gcb_xtc -- small companion for trajectories (test and benchmark helper, no reference counterpart):
 *   gcb_xtc convert IN OUT.xtc [precision]   any readable trajectory (xtc, gro, GCTRJ1 raw) -> xtc
 *   gcb_xtc info FILE                         frames, atoms, time range
 *   gcb_xtc bench FILE [repeats]              host decode throughput (atoms/s), the separately timed XTC decode
 */
#include "gcb_host.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char **argv)
{
  gh_set_program(argv[0]);
  if (argc < 3) { fprintf(stderr, "usage: %s convert IN OUT.xtc [precision] | info FILE | bench FILE [repeats]\n", argv[0]); return 2; }
  gh_traj *t = gh_traj_open(argv[2]);
  if (!t) gh_fatal("Cannot read trajectory %s", argv[2]);
  const int natoms = gh_traj_natoms(t);
  float *x = (float *)gh_alloc(sizeof(float) * 3 * (size_t)natoms);
  gh_frame_info fr;
  if (!strcmp(argv[1], "convert")) {
    if (argc < 4) gh_fatal("convert needs an output name");
    const float prec = argc > 4 ? (float)atof(argv[4]) : 1000.0f;
    gh_xtc_writer *w = gh_xtc_create(argv[3]);
    long n = 0;
    int rc;
    if (!w) gh_fatal("Cannot write %s", argv[3]);
    while ((rc = gh_traj_next(t, &fr, x)) == 1) { if (gh_xtc_write(w, natoms, fr.step, fr.time, fr.box, x, prec)) gh_fatal("encode failed"); n++; }
    if (rc < 0) gh_fatal("damaged input after %ld frames", n);
    gh_xtc_close(w);
    printf("%ld frames, %d atoms\n", n, natoms);
  } else if (!strcmp(argv[1], "info")) {
    long n = 0;
    float t0 = 0, t1 = 0;
    int rc;
    while ((rc = gh_traj_next(t, &fr, x)) == 1) { if (!n) t0 = fr.time; t1 = fr.time; n++; }
    printf("%ld frames, %d atoms, t = %g .. %g ps%s\n", n, natoms, t0, t1, rc < 0 ? " (damaged tail)" : "");
  } else if (!strcmp(argv[1], "bench")) {
    const int reps = argc > 3 ? atoi(argv[3]) : 1;
    long n = 0;
    int r;
    double t0 = now();
    for (r = 0; r < reps; r++) {
      gh_traj *u = gh_traj_open(argv[2]);
      while (gh_traj_next(u, &fr, x) == 1) n++;
      gh_traj_close(u);
    }
    const double dt = now() - t0;
    printf("{\"frames\": %ld, \"natoms\": %d, \"seconds\": %.6f, \"atom_frames_per_s\": %.6g}\n", n, natoms, dt, (double)n * natoms / dt);
  } else {
    gh_fatal("unknown command %s", argv[1]);
  }
  gh_traj_close(t);
  free(x);
  return 0;
}
