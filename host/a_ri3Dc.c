/*
 * a_ri3Dc -- drop-in for the reference's grid analysis tool (src/a_ri3Dc.c): reads the XDR grid
 * file g_ri3Dc wrote and produces the projections, radial / axial distribution functions, the
 * pore profile R*(z), and the DATA_RESULT lines.
 *
 * Option table and defaults: a_ri3Dc.c:324-378.  All reductions (a_ri3Dc.c:488-762) run on the
 * GPU through gcb_analyse(); this file only parses options and writes the reference's text
 * formats (a_ri3Dc.c:764-962, xf.c).
 */
#include "gcb_host.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const char *synopsis =
    "a_ri3Dc analyses a 3D grid produced by g_ri3Dc. The 2D projections are written in a format suitable for the\n"
    "fast 2D density plotter xfarbe, together with the parameter file `XFarbe'. All geometry options (radius, z1, z2)\n"
    "that are not set are taken from the grid dimensions. -dump converts the binary grid into ascii text, -plt\n"
    "produces a binary density file for gOpenMol / VMD.\n"
    "  xy, z-averaged 1/Lz Int_z n;  xz, y-averaged;  yz, x-averaged;  radially averaged 1/2pi r Int_phi n(r,phi,z)\n"
    "  pore profile (z, R*(z), R*(z)/R);  rdf P(r);  zdf P(z);  lzdf: local density over an effective area pi R*(z)^2";

static const char *unit_names[] = {"nm^-3", "n(SPC,T=300K,P=1bar)", "mol/l", "Angstrom^-3", "Voxel"};   /* count.c:16-17 */

#define CHECK(call) do { if ((call) != GCB_OK) gh_fatal("%s", gcb_last_error()); } while (0)

static void write_matrix(const char *path, const char *header, const float *a, int n0, int n1, int s0, int s1,
                         float maxlevel, float x1, float x2, float y1, float y2)
{
  /* rows run over the second index; element (p, q) of the [n0][n1] block sits at a[p*s0 + q*s1] */
  FILE *f = gh_xf_open(path, header, n0, n1);
  int p, q;
  for (q = 0; q < n1; q++) {
    for (p = 0; p < n0; p++) fprintf(f, "%f ", a[(size_t)p * s0 + (size_t)q * s1]);
    fprintf(f, "\n");
  }
  gh_xf_levels(f, gh_xf_ncols(), maxlevel);
  gh_xf_axes(f, x1, x2, y1, y2);
  fclose(f);
}

static void write_rz(const char *path, const char *header, const float *rz, int nr, int mz, int mirror, float maxlevel,
                     float radius, float DeltaR, float z1, float z2)
{
  FILE *f = gh_xf_open(path, header, mirror ? 2 * nr - 1 : nr, mz);
  int k, ir;
  for (k = 0; k < mz; k++) {
    for (ir = mirror ? -nr + 1 : 0; ir < nr; ir++) fprintf(f, "%f ", rz[(size_t)abs(ir) * mz + k]);
    fprintf(f, "\n");
  }
  gh_xf_levels(f, gh_xf_ncols(), maxlevel);
  gh_xf_axes(f, mirror ? -radius + DeltaR : radius, radius, z1, z2);      /* a_ri3Dc.c:860-862 */
  fclose(f);
}

int main(int argc, char **argv)
{
  gcb_cavity geometry;
  float min_occ = 0, rsolvent = 0.14f, maxDensity = 48.48405f, Delta[3] = {0.02f, 0.02f, 0.02f};
  int bMirror = 1, bDoHoles = 0, rad_ba_nr = 0, rad_ba_step = 1, device = 0;
  const char *unitstr[] = {NULL, "unity", "SPC", "molar", "Angstrom", "Voxel", NULL};
  const char *subtitle = NULL;
  memset(&geometry, 0, sizeof geometry);

  gh_option pa[] = {
      {"-R", GH_REAL, &geometry.radius, "Maximum radius that should be contained in the grid", 0},
      {"-z1", GH_REAL, &geometry.z1, "Center grid between z1, and ...", 0},
      {"-z2", GH_REAL, &geometry.z2, "z2 (these boundaries are kept fixed!)", 0},
      {"-delta", GH_RVEC, Delta, NULL, 0},
      {"-minocc", GH_REAL, &min_occ, NULL, 0},
      {"-rsolvent", GH_REAL, &rsolvent, NULL, 0},
      {"-subtitle", GH_STR, &subtitle, "Some text to add to the output graphs", 0},
      {"-mirror", GH_BOOL, &bMirror, "mirror the radial projection P(r,z) to create the impression of a full view of the pore", 0},
      {"-unit", GH_ENUM, unitstr, "output density unit: unity (nm^-3), SPC, molar, Angstrom, Voxel", 0},
      {"-xfarbe-maxlevel", GH_REAL, &maxDensity, "xfarbe will plot 15 equally spaced levels up to this density (default 1.5 SPC bulk)", 0},
      {"-holes", GH_BOOL, &bDoHoles, NULL, 0},
      {"-radnr", GH_INT, &rad_ba_nr, NULL, 0},
      {"-radstep", GH_INT, &rad_ba_step, NULL, 0},
      {"-gpu", GH_INT, &device, "CUDA device to use", 0},
  };
  gh_file fnm[] = {
      {"-grid", "gridxdr", ".dat", GH_F_READ, "3D density grid", 0, NULL, NULL, 0},
      {"-profile", "profile", ".xvg", GH_F_WRITE, "pore profile", 0, NULL, NULL, 0},
      {"-xyp", "xyp", ".dat", GH_F_WRITE, "xy projection", 0, NULL, NULL, 0},
      {"-xzp", "xzp", ".dat", GH_F_OPT_WRITE, "xz projection", 0, NULL, NULL, 0},
      {"-yzp", "yzp", ".dat", GH_F_OPT_WRITE, "yz projection", 0, NULL, NULL, 0},
      {"-rzp", "rzp", ".dat", GH_F_WRITE, "radial projection", 0, NULL, NULL, 0},
      {"-rdf", "rdf", ".xvg", GH_F_WRITE, "radial distribution function", 0, NULL, NULL, 0},
      {"-zdf", "zdf", ".xvg", GH_F_WRITE, "axial distribution function", 0, NULL, NULL, 0},
      {"-lzdf", "lzdf", ".xvg", GH_F_WRITE, "local density axial distribution function", 0, NULL, NULL, 0},
      {"-hxyp", "hxyp", ".dat", GH_F_OPT_WRITE, "xy projection of unoccupied cells", 0, NULL, NULL, 0},
      {"-hrzp", "hrzp", ".dat", GH_F_OPT_WRITE, "radial projection of unoccupied cells", 0, NULL, NULL, 0},
      {"-hrdf", "hrdf", ".xvg", GH_F_OPT_WRITE, "rdf of unoccupied cells", 0, NULL, NULL, 0},
      {"-dump", "gridasc", ".dat", GH_F_OPT_WRITE, "ascii dump of the grid", 0, NULL, NULL, 0},
      {"-plt", "plt", ".dat", GH_F_OPT_WRITE, "gOpenMol density", 0, NULL, NULL, 0},
  };
  const int NPA = (int)(sizeof pa / sizeof pa[0]), NFILE = (int)(sizeof fnm / sizeof fnm[0]);

  gh_parse_args(argc, argv, synopsis, fnm, NFILE, pa, NPA, NULL);
  if (rad_ba_step <= 0 || rad_ba_nr % rad_ba_step != 0) gh_fatal("-radnr must be a multiple of -radstep");   /* assert at a_ri3Dc.c:438 */
  if (bDoHoles) gh_msg("Sorry, this option does not work at the moment.\n");

  int nunit = GCB_UNIT_UNITY;                                   /* a_ri3Dc.c:451-470: first letter decides */
  switch (unitstr[0][0]) {
    case 'S': nunit = GCB_UNIT_SPC; break;
    case 'm': nunit = GCB_UNIT_MOLAR; break;
    case 'A': nunit = GCB_UNIT_ANG; break;
    case 'V': nunit = GCB_UNIT_VOXEL; break;
    default: nunit = GCB_UNIT_UNITY;
  }

  gh_msg("Reading grid 3D file...\n");
  gcb_tgrid tg;
  char header[GCB_HEADER_MAX + 1];
  float *grid = NULL;
  if (gcb_grid_read(gh_fn(fnm, NFILE, "-grid"), &tg, header, &grid) != GCB_OK)
    gh_fatal("Error reading the 3D grid---no point in continuing!\n%s", gcb_last_error());
  gh_msg(".. done!\n");
  /* -subtitle is accepted but has no effect, as in the reference: grid_read decodes the stored header into the
   * buffer the option pointed to (grid3D.c:65, a_ri3Dc.c:473), so the file's header always wins */
  (void)subtitle;

  gcb_analysis_opts ao;
  memset(&ao, 0, sizeof ao);
  ao.unit = nunit; ao.min_occ = min_occ; ao.rad_ba_nr = rad_ba_nr; ao.rad_ba_step = rad_ba_step; ao.device = device;
  ao.setmask = (gh_opt_set(pa, NPA, "-R") ? GCB_SET_R : 0) | (gh_opt_set(pa, NPA, "-z1") ? GCB_SET_Z1 : 0) |
               (gh_opt_set(pa, NPA, "-z2") ? GCB_SET_Z2 : 0);
  ao.radius = geometry.radius; ao.z1 = geometry.z1; ao.z2 = geometry.z2;
  gcb_analysis an;
  CHECK(gcb_analyse(&tg, grid, &ao, &an));     /* an.grid is NULL unless the analysis cropped: `grid` is needed until the end */
  if (ao.setmask & GCB_SET_Z1) gh_msg("Readjusted: z1 = %g nm (supplied: %g)\n", an.tg.a[2], geometry.z1);
  if (ao.setmask & GCB_SET_Z2) gh_msg("Readjusted: z2 = %g nm (supplied: %g)\n", an.tg.b[2], geometry.z2);
  if (ao.setmask & GCB_SET_R) gh_msg("Readjusted: R = %g nm (supplied: %g)\n", an.geometry.radius, geometry.radius);

  const float U = an.unit_value;
  if (!gh_opt_set(pa, NPA, "-xfarbe-maxlevel")) maxDensity /= U;            /* a_ri3Dc.c:520-524 */
  const int mx = an.tg.mx[0], my = an.tg.mx[1], mz = an.tg.mx[2], nr = an.iradNR;
  const float *a = an.tg.a, *b = an.tg.b;
  const float R = an.geometry.radius;

  gh_msg("\n# ---> Result <-----\n# Densities have the unit %s\n  V=%g nm^3  R*=%g nm sumP=%g sumH=%g dV=%f\n", unit_names[nunit],
         an.volume, an.reff, an.sumP, an.sumH, an.dV);
  printf("\nDATA_RESULT:  V %f   R* %f   sumP %f sumH %f  dV %f # %s\n", an.volume, an.reff, an.sumP, an.sumH, an.dV, header);
  printf("DATA_RESULT:  average %f  iVolume %f\n", an.average, (float)an.ivol * an.dV);

  char ylabel[256];
  FILE *f;
  int k, ir;
  f = gh_xvg_open(gh_fn(fnm, NFILE, "-profile"), "Effective radius R*(z) pore profile", header, "z [nm]", "R* [nm]");
  for (k = 0; k < mz; k++) fprintf(f, "%.6f   %.6f  %.6f\n", an.profile[2 * k], an.profile[2 * k + 1], an.profile[2 * k + 1] / R);
  fclose(f);

  /* projections: xyp[i][j] at i*my + j, xzp[i][k] at i*mz + k, yzp[j][k] at j*mz + k */
  write_matrix(gh_fn(fnm, NFILE, "-xyp"), header, an.xyp, mx, my, my, 1, maxDensity, a[0], b[0], a[1], b[1]);
  if (gh_fn_set(fnm, NFILE, "-hxyp")) write_matrix(gh_fn(fnm, NFILE, "-hxyp"), header, an.hxyp, mx, my, my, 1, 1, a[0], b[0], a[1], b[1]);
  if (gh_fn_set(fnm, NFILE, "-xzp")) write_matrix(gh_fn(fnm, NFILE, "-xzp"), header, an.xzp, mx, mz, mz, 1, maxDensity, a[0], b[0], a[2], b[2]);
  if (gh_fn_set(fnm, NFILE, "-yzp")) {
    /* the reference announces mx columns here but writes my per row (a_ri3Dc.c:834-838): kept */
    FILE *g = gh_xf_open(gh_fn(fnm, NFILE, "-yzp"), header, mx, mz);
    int j;
    for (k = 0; k < mz; k++) {
      for (j = 0; j < my; j++) fprintf(g, "%f ", an.yzp[(size_t)j * mz + k]);
      fprintf(g, "\n");
    }
    gh_xf_levels(g, gh_xf_ncols(), maxDensity);
    gh_xf_axes(g, a[1], b[1], a[2], b[2]);
    fclose(g);
  }
  write_rz(gh_fn(fnm, NFILE, "-rzp"), header, an.rzp, nr, mz, bMirror, maxDensity, R, an.DeltaR, a[2], b[2]);
  if (gh_fn_set(fnm, NFILE, "-hrzp")) write_rz(gh_fn(fnm, NFILE, "-hrzp"), header, an.hrzp, nr, mz, bMirror, 1, R, an.DeltaR, a[2], b[2]);

  snprintf(ylabel, sizeof ylabel, "n(z) [%s], PMF G(z)/kT", unit_names[nunit]);
  f = gh_xvg_open(gh_fn(fnm, NFILE, "-zdf"), "Axial distribution function n(z)", header, "z [nm]", ylabel);
  for (k = 0; k < mz; k++) fprintf(f, "%.6f   %.6f   %.6f\n", an.zdf[2 * k], an.zdf[2 * k + 1], -log(an.zdf[2 * k + 1]));
  fclose(f);

  snprintf(ylabel, sizeof ylabel, "n\\slocal\\N(z) [%s]", unit_names[nunit]);
  f = gh_xvg_open(gh_fn(fnm, NFILE, "-lzdf"), "Local density axial distribution function n\\slocal\\N(z)", header, "z [nm]", ylabel);
  for (k = 0; k < mz; k++) {
    const float lz = an.lzdf[2 * k + 1], q = rsolvent / an.profile[2 * k + 1];
    fprintf(f, "%.6f   %.6f  %.6f  %.6f\n", an.lzdf[2 * k], lz, lz / (1 + q * q), -log(lz));      /* a_ri3Dc.c:896-901 */
  }
  fclose(f);

  snprintf(ylabel, sizeof ylabel, "n(r) [%s]", unit_names[nunit]);
  f = gh_xvg_open(gh_fn(fnm, NFILE, "-rdf"), "Radial distribution function P(r)", header, "r [nm]", ylabel);
  for (ir = 0; ir < nr; ir++) fprintf(f, "%.6f   %.6f\n", an.rdf[2 * ir], an.rdf[2 * ir + 1]);
  fclose(f);
  if (gh_fn_set(fnm, NFILE, "-hrdf")) {
    f = gh_xvg_open(gh_fn(fnm, NFILE, "-hrdf"), "Radial distribution function of unoccupied cells", header, "r [nm]", "h(r)");
    for (ir = 0; ir < nr; ir++) fprintf(f, "%.6f   %.6f\n", an.hrdf[2 * ir], an.hrdf[2 * ir + 1]);
    fclose(f);
  }

  if (gh_fn_set(fnm, NFILE, "-dump")) {
    gh_msg("Dumping the 3D grid into ascii file %s...[Unit is %s]", gh_fn(fnm, NFILE, "-dump"), unit_names[nunit]);
    CHECK(gcb_ascii_write(gh_fn(fnm, NFILE, "-dump"), &an.tg, an.grid ? an.grid : grid, U));
    gh_msg(".. done!\n");
  }
  if (gh_fn_set(fnm, NFILE, "-plt")) {
    gh_msg("Dumping the 3D grid into gOpenMol density file %s... (rename to xxx.plt)\n[Unit is %s]", gh_fn(fnm, NFILE, "-plt"),
           unit_names[nunit]);
    CHECK(gcb_plt_write(gh_fn(fnm, NFILE, "-plt"), &an.tg, an.grid ? an.grid : grid, U));
    gh_msg("... done!\n");
  }
  gh_xf_write_resource();
  gcb_analysis_free(&an);
  gcb_free(grid);
  gh_msg("\n");
  return 0;
}
