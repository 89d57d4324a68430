/*
 * a_gridcalc -- drop-in for the reference tool (src/a_gridcalc.c): time-weighted average of
 * several 3D grids, avg += g * (T_g / Sum T) (a_gridcalc.c:176-189), written as a new grid file.
 * Options as the reference: -f grid files, -o output (a_gridcalc.c:92-100).
 */
#include "gcb_host.h"
#include <stdlib.h>
#include <string.h>

static const char *synopsis =
    "a_gridcalc computes the time-weighted average of 3D grids written by g_ri3Dc. All grids must have the same\n"
    "dimensions; no sanity checks are performed.";

#define CHECK(call) do { if ((call) != GCB_OK) gh_fatal("%s", gcb_last_error()); } while (0)

int main(int argc, char **argv)
{
  int device = 0;
  gh_option pa[] = {{"-gpu", GH_INT, &device, "CUDA device to use", 0}};
  gh_file fnm[] = {
      {"-f", "gridxdr", ".dat", GH_F_READ_MULT, "3D density grids", 0, NULL, NULL, 0},
      {"-o", "gridout", ".dat", GH_F_WRITE, "averaged grid", 0, NULL, NULL, 0},
  };
  gh_parse_args(argc, argv, synopsis, fnm, 2, pa, 1, NULL);
  const int nfile = fnm[0].nnames;
  if (!nfile) gh_fatal("No input files!");
  float **grids = (float **)gh_zalloc((size_t)nfile, sizeof(float *));
  float *tw = (float *)gh_zalloc((size_t)nfile, sizeof(float));
  gcb_tgrid tg0, tg;
  char header[GCB_HEADER_MAX + 1];
  double sumT = 0;
  int i;
  for (i = 0; i < nfile; i++) {
    gh_msg("Reading grid 3D file %s\n", fnm[0].names[i]);
    if (gcb_grid_read(fnm[0].names[i], &tg, header, &grids[i]) != GCB_OK)
      gh_fatal("Error reading the 3D grid---no point in continuing!\n%s", gcb_last_error());
    gh_msg("Tweight = %g ps\nmx = %d %d %d\n.. done!\n", tg.tweight, tg.mx[0], tg.mx[1], tg.mx[2]);
    if (i == 0) tg0 = tg;
    else if (tg.mx[0] != tg0.mx[0] || tg.mx[1] != tg0.mx[1] || tg.mx[2] != tg0.mx[2])
      gh_fatal("Grid %s is %d x %d x %d, the first one %d x %d x %d", fnm[0].names[i], tg.mx[0], tg.mx[1], tg.mx[2], tg0.mx[0],
               tg0.mx[1], tg0.mx[2]);
    tw[i] = tg.tweight;
    sumT += (double)tg.tweight;
  }
  gh_msg("No sanity checks... proceeding with the above %d grids.\n", nfile);
  const size_t cells = (size_t)tg0.mx[0] * tg0.mx[1] * tg0.mx[2];
  float *avg = (float *)gh_alloc(sizeof(float) * cells);
  float tweight = 0;
  CHECK(gcb_gridcalc(nfile, (const float *const *)grids, tw, &tg0, avg, &tweight, device));
  tg0.tweight = tweight;
  /* header as a_gridcalc.c:192-198 */
  snprintf(header, GCB_HEADER_MAX, "Average density: {Ttot}%gps{Ngrid}%d{<Delta>}%g{mx,my,mz}%d,%d,%d", sumT, nfile,
           (tg0.delta[0] + tg0.delta[1] + tg0.delta[2]) / 3, tg0.mx[0], tg0.mx[1], tg0.mx[2]);
  header[GCB_HEADER_MAX - 1] = 0;
  if (gcb_grid_write(gh_fn(fnm, 2, "-o"), &tg0, avg, NULL, header) != GCB_OK) gh_msg("Writing the 3D grid failed.\n%s\n", gcb_last_error());
  for (i = 0; i < nfile; i++) gcb_free(grids[i]);
  free(grids); free(tw); free(avg);
  return 0;
}
