/*
 * gc_oracle.c -- TEST INFRASTRUCTURE ONLY (see gc_oracle.h).
 *
 * Plain-C restatement of the reference algorithm on flat arrays.  The
 * arithmetic types and the order of every rounding step follow the reference
 * expression by expression (C's usual arithmetic conversions applied to
 * `real` = float operands and double constants); build with
 * -ffp-contract=off and without -march flags that enable FMA.
 *
 * Citations are file:line into /root/reference.
 */
#include "gc_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define PI_D 3.14159265358979323846 /* M_PI, include/utilgmx.h:24-28 */

static float minf_d(double p, double q) { return (float)(p < q ? p : q); }

/* ------------------------------------------------------------------ */
/* geometry                                                            */
/* ------------------------------------------------------------------ */

/* count.c:25-64 */
int gco_autoset_cavity(gco_cavity *cav, const float box9[9], int setmask)
{
  int changed = 0;
  const float Lx = box9[0], Ly = box9[4], Lz = box9[8];
  if (!(setmask & GCO_SET_CPOINT)) {
    cav->cpoint[0] = (float)(0.5 * (double)Lx);         /* count.c:29 */
    cav->cpoint[1] = (float)(0.5 * (double)Ly);
    cav->cpoint[2] = (float)(0.5 * (double)Lz);
    changed = 1;
  }
  if (!(setmask & GCO_SET_R)) {                          /* count.c:35-38 */
    float dx = cav->cpoint[0] - Lx, dy = cav->cpoint[1] - Ly;
    double mx_ = (double)cav->cpoint[0] < fabs((double)dx) ? (double)cav->cpoint[0] : fabs((double)dx);
    double my_ = (double)cav->cpoint[1] < fabs((double)dy) ? (double)cav->cpoint[1] : fabs((double)dy);
    cav->radius = minf_d(mx_, my_);
    changed = 1;
  }
  if (!(setmask & GCO_SET_Z1)) { cav->z1 = 0; changed = 1; }      /* count.c:42-46 */
  if (!(setmask & GCO_SET_Z2)) { cav->z2 = Lz; changed = 1; }     /* count.c:47-51 */
  if (cav->z1 > cav->z2) {                                        /* count.c:53-59 */
    float tmp = cav->z1; cav->z1 = cav->z2; cav->z2 = tmp; changed = 1;
  }
  {                                                               /* count.c:60 */
    float h = cav->z2 - cav->z1;
    cav->vol = (float)(PI_D * (double)cav->radius * (double)cav->radius * (double)h);
  }
  return changed;
}

/* grid3D.c:125-203 (setup_corners folded in: only c[0] and c[7] are used) */
int gco_setup_grid(gco_geom *g, gco_cavity *cav, const float delta[3])
{
  const float R = cav->radius;
  const float twoR = 2 * R;                                       /* grid3D.c:182-183 */
  const double eps100 = 100 * GCO_GMX_REAL_EPS;
  int d;
  /* c[0]  (grid3D.c:190-192) */
  g->a[0] = cav->cpoint[0] - R;
  g->a[1] = cav->cpoint[1] - R;
  g->a[2] = cav->z1;
  /* c[7] via c[4] -> c[5] -> c[6] -> c[7]  (grid3D.c:196-201) */
  {
    float c4x = cav->cpoint[0] - R, c4y = cav->cpoint[1] + R, c4z = cav->z2;
    float c5x = c4x - 0.0f, c5y = c4y - twoR, c5z = c4z - 0.0f;
    float c6x = c5x + twoR, c6y = c5y + 0.0f, c6z = c5z + 0.0f;
    g->b[0] = c6x + 0.0f; g->b[1] = c6y + twoR; g->b[2] = c6z + 0.0f;
  }
  for (d = 0; d < 3; d++) {
    float span, du, fill;
    g->delta[d] = delta[d];
    span = g->b[d] - g->a[d];
    g->mx[d] = (int)floor((double)(span / delta[d]) + eps100);   /* grid3D.c:143 */
    fill = (float)g->mx[d] * delta[d];
    du = (float)(0.5 * (double)(span - fill));                    /* grid3D.c:145 */
    if ((double)(float)fabs((double)du) < eps100) du = 0;         /* grid3D.c:146 */
    g->a[d] += du;
    g->b[d] -= du;
  }
  cav->z1 = g->a[2];                                              /* grid3D.c:153-156 */
  cav->z2 = g->b[2];
  {
    double wx = fabs((double)(g->b[0] - g->a[0])), wy = fabs((double)(g->b[1] - g->a[1]));
    cav->radius = (float)(0.5 * (wx < wy ? wx : wy));
  }
  return g->mx[0] > 0 && g->mx[1] > 0 && g->mx[2] > 0;
}

double gco_delta_v(const gco_geom *g)
{
  return (double)g->delta[0] * (double)g->delta[1] * (double)g->delta[2];
}

int gco_edge_vulnerable(const gco_geom *g)
{
  int d, bad = 0;
  for (d = 0; d < 3; d++) {
    float x = nextafterf(g->b[d], -INFINITY);
    float u = x - g->a[d];
    if (u >= 0 && !(u / g->delta[d] < (float)g->mx[d])) bad = 1;
  }
  return bad;
}

/* ------------------------------------------------------------------ */
/* frame loop bookkeeping                                              */
/* ------------------------------------------------------------------ */

/* g_ri3Dc.c:390-394 (dt from the first two frames, tlast = t0 - dt) and
 * :413-416 (dt = t - tlast in f32, T_tot += dt in f64, weight = dt or 1). */
void gco_frame_weights(const float *t, long nframes, int dtweight, float *weight, double *T_tot)
{
  long f;
  double T = 0;
  float dt0 = nframes > 1 ? t[1] - t[0] : 0;   /* get_timestep, g_ri3Dc.c:155-177 */
  float tlast = nframes > 0 ? t[0] - dt0 : 0;
  for (f = 0; f < nframes; f++) {
    float dt = t[f] - tlast;
    tlast = t[f];
    T += (double)dt;
    weight[f] = dtweight ? dt : 1.0f;
  }
  *T_tot = T;
}

/* ------------------------------------------------------------------ */
/* binning                                                             */
/* ------------------------------------------------------------------ */

/* g_ri3Dc.c:144-147; returns the flat z-fastest cell, -1 if outside, -2 if
 * the atom passes all three bounds tests but a quotient lands on or beyond mx
 * or is NaN (undefined behaviour in the reference, SURVEY 7.3 item 2: dropped
 * and counted here and in the CUDA path).  The reference returns at the first
 * failing bounds test, so an atom that is outside in any dimension is never
 * dangerous: "outside" takes precedence over "overflow". */
static inline long cell_of(const gco_geom *g, const float *x)
{
  int ijk[3], d, ovf = 0;
  for (d = 0; d < 3; d++) {
    float u = x[d] - g->a[d];
    float q;
    if (u < 0 || x[d] - g->b[d] >= 0) return -1;
    q = u / g->delta[d];
    if (!(q < (float)g->mx[d])) { ovf = 1; ijk[d] = 0; }
    else ijk[d] = (int)floor((double)q);
  }
  if (ovf) return -2;
  return ((long)ijk[0] * g->mx[1] + ijk[1]) * g->mx[2] + ijk[2];  /* utilgmx.c:57-69 */
}

void gco_bin_frames(const gco_geom *g, float *grid, const float *x, long nframes,
                    int natoms, const int *gindex, int gnx, const float *weight,
                    double *sumP, uint64_t *hits_per_frame, uint64_t *edge_overflow)
{
  long f;
  int i;
  double s = sumP ? *sumP : 0;
  uint64_t ovf = 0;
  for (f = 0; f < nframes; f++) {
    const float *xf = x + (size_t)f * natoms * 3;
    const float w = weight ? weight[f] : 1.0f;
    uint64_t hits = 0;
    for (i = 0; i < gnx; i++) {
      long c = cell_of(g, xf + 3 * (size_t)gindex[i]);
      if (c >= 0) {
        grid[c] += w;             /* g_ri3Dc.c:149 */
        s += (double)w;           /* g_ri3Dc.c:151,426 */
        hits++;
      } else if (c == -2) ovf++;
    }
    if (hits_per_frame) hits_per_frame[f] = hits;
  }
  if (sumP) *sumP = s;
  if (edge_overflow) *edge_overflow += ovf;
}

/* The same accumulation with `real = double` (a Gromacs -DDOUBLE build of g_ri3Dc.c:149: `grid[i][j][k] += dt`
 * in f64; bins from the same f32 coordinates).  The f32 chain above carries its own summation error (SURVEY.md
 * section 0: 1.5e-4 after 15 625 additions of 0.1); this is the value that error is measured against, and the
 * yardstick for results whose f32 additions happen in another order (several ranks with differing weights). */
void gco_bin_frames_f64(const gco_geom *g, double *grid, const float *x, long nframes,
                        int natoms, const int *gindex, int gnx, const float *weight)
{
  long f;
  int i;
  for (f = 0; f < nframes; f++) {
    const float *xf = x + (size_t)f * natoms * 3;
    const double w = weight ? (double)weight[f] : 1.0;
    for (i = 0; i < gnx; i++) {
      long c = cell_of(g, xf + 3 * (size_t)gindex[i]);
      if (c >= 0) grid[c] += w;
    }
  }
}

void gco_count_frames(const gco_geom *g, uint32_t *counts, const float *x, long nframes,
                      int natoms, const int *gindex, int gnx,
                      uint64_t *hits_per_frame, uint64_t *edge_overflow)
{
  long f;
  int i;
  uint64_t ovf = 0;
  for (f = 0; f < nframes; f++) {
    const float *xf = x + (size_t)f * natoms * 3;
    uint64_t hits = 0;
    for (i = 0; i < gnx; i++) {
      long c = cell_of(g, xf + 3 * (size_t)gindex[i]);
      if (c >= 0) { counts[c]++; hits++; }
      else if (c == -2) ovf++;
    }
    if (hits_per_frame) hits_per_frame[f] = hits;
  }
  if (edge_overflow) *edge_overflow += ovf;
}

/* count.c:88-92: one image at most, lower test `< 0`, upper test `> L` */
void gco_wrap_frames(float *x, long nframes, int natoms, const float *box9_per_frame)
{
  long f;
  int i, d;
  for (f = 0; f < nframes; f++) {
    const float *box = box9_per_frame + 9 * f;
    float *xf = x + (size_t)f * natoms * 3;
    for (i = 0; i < natoms; i++)
      for (d = 0; d < 3; d++) {
        float L = box[4 * d];
        float v = xf[3 * i + d];
        if (v < 0) v += L;
        if (v > L) v -= L;
        xf[3 * i + d] = v;
      }
  }
}

/* n sequential `v += w` in f32.  Within one binade the rounded increment is
 * a constant number of ulps once one addition has been made inside it (with
 * round-half-even a tie can only change the first in-binade step, after which
 * the running mantissa stays even), so whole binades are skipped at a time. */
float gco_seq_add_f32(float v, float w, uint64_t n)
{
  while (n > 0) {
    float v1 = v + w, v2, v3;
    uint32_t b1, b2, b3;
    n--;
    if (v1 == v || v1 != v1) return v1;   /* saturated (or NaN): nothing changes any more */
    v = v1;
    if (n < 2) continue;
    v2 = v + w;                            /* first step certainly made inside v's binade ... */
    v3 = v2 + w;                           /* ... and the steady-state step after it */
    memcpy(&b1, &v, 4);
    memcpy(&b2, &v2, 4);
    memcpy(&b3, &v3, 4);
    if (v > 0 && w > 0 && (b1 >> 23) == (b2 >> 23) && (b2 >> 23) == (b3 >> 23) && b3 > b2) {
      uint32_t inc = b3 - b2;
      uint32_t top = b2 | 0x7FFFFFu;              /* last pattern of this binade */
      uint64_t k;
      n--;                                        /* commit v -> v2 */
      k = (uint64_t)(top - b2) / inc;             /* further steps that stay inside */
      if (k > n) k = n;
      b2 += (uint32_t)(k * inc);
      n -= k;
      memcpy(&v, &b2, 4);
    }
  }
  return v;
}

double gco_seq_add_f64(double v, double w, uint64_t n)
{
  while (n > 0) {
    double v1 = v + w, v2, v3;
    uint64_t b1, b2, b3;
    n--;
    if (v1 == v || v1 != v1) return v1;
    v = v1;
    if (n < 2) continue;
    v2 = v + w;
    v3 = v2 + w;
    memcpy(&b1, &v, 8);
    memcpy(&b2, &v2, 8);
    memcpy(&b3, &v3, 8);
    if (v > 0 && w > 0 && (b1 >> 52) == (b2 >> 52) && (b2 >> 52) == (b3 >> 52) && b3 > b2) {
      uint64_t inc = b3 - b2;
      uint64_t top = b2 | 0xFFFFFFFFFFFFFull;
      uint64_t k;
      n--;
      k = (top - b2) / inc;
      if (k > n) k = n;
      b2 += k * inc;
      n -= k;
      memcpy(&v, &b2, 8);
    }
  }
  return v;
}

/* ------------------------------------------------------------------ */
/* normalisation, header                                               */
/* ------------------------------------------------------------------ */

/* g_ri3Dc.c:451-460 */
void gco_normalise(gco_geom *g, float *grid, double T_tot)
{
  const float dV = (float)gco_delta_v(g);                /* `real dV = DeltaV()` */
  const double denom = (double)dV * T_tot;
  size_t n = (size_t)g->mx[0] * g->mx[1] * g->mx[2], c;
  for (c = 0; c < n; c++) grid[c] = (float)((double)grid[c] / denom);
  g->tweight = (float)T_tot;
}

/* g_ri3Dc.c:462-473 */
int gco_header(char out[256], const char *subtitle, double T_tot, const char *grpname,
               const gco_cavity *cav, const gco_geom *g, double sumP,
               const char *trxname, int dtweight)
{
  float dsum = g->delta[0] + g->delta[1] + g->delta[2];
  int n = snprintf(out, 256,
                   "%s{T}%gps{grp}%s{z1}%.2f{z2}%.2f{Rmax}%.2f{<Delta>}%.3f{<N>}%.2f{XTC}%s"
                   "{Delta}(%.3f,%.3f,%.3f){sumP}%g%s",
                   subtitle ? subtitle : "", T_tot, grpname, cav->z1, cav->z2, cav->radius,
                   (double)dsum / 3., sumP / T_tot, trxname,
                   g->delta[0], g->delta[1], g->delta[2], sumP, dtweight ? "ps" : "");
  out[255] = '\0';
  return n;
}

/* ------------------------------------------------------------------ */
/* XDR grid file                                                       */
/* ------------------------------------------------------------------ */
static uint8_t *put32(uint8_t *p, uint32_t w)
{
  p[0] = (uint8_t)(w >> 24); p[1] = (uint8_t)(w >> 16); p[2] = (uint8_t)(w >> 8); p[3] = (uint8_t)w;
  return p + 4;
}
static uint8_t *putf(uint8_t *p, float f) { uint32_t w; memcpy(&w, &f, 4); return put32(p, w); }
static uint32_t get32(const uint8_t *p)
{
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}
static float getf(const uint8_t *p) { uint32_t w = get32(p); float f; memcpy(&f, &w, 4); return f; }

size_t gco_xdr_size(const gco_geom *g, const char *header)
{
  size_t len = strlen(header), cells = (size_t)g->mx[0] * g->mx[1] * g->mx[2];
  return 4 + 4 + ((len + 3) & ~(size_t)3) + 4 + 4 + 4 + 12 + 12 + 12 + 4 + 4 * cells;
}

/* field order of xdr_grid_v2 (xdr_grid.c:89-107); data x-fastest (xdr_grid.c:113-126) */
size_t gco_xdr_encode(const gco_geom *g, const float *grid, const char *header,
                      uint8_t *out, size_t cap)
{
  size_t len = strlen(header), need = gco_xdr_size(g, header);
  uint8_t *p = out;
  int i, j, k, d;
  const int mx = g->mx[0], my = g->mx[1], mz = g->mx[2];
  if (need > cap || len > 256) return 0;
  p = put32(p, 2);                                   /* GRID_FF_VERSION */
  p = put32(p, (uint32_t)len);
  memcpy(p, header, len);
  memset(p + len, 0, ((len + 3) & ~(size_t)3) - len);
  p += (len + 3) & ~(size_t)3;
  p = put32(p, 0);                                   /* egtyREGULAR */
  p = putf(p, g->tweight);
  p = put32(p, 3);
  for (d = 0; d < 3; d++) p = put32(p, (uint32_t)g->mx[d]);
  for (d = 0; d < 3; d++) p = putf(p, g->delta[d]);
  for (d = 0; d < 3; d++) p = putf(p, g->a[d]);
  p = put32(p, (uint32_t)((size_t)mx * my * mz));
  for (k = 0; k < mz; k++)
    for (j = 0; j < my; j++)
      for (i = 0; i < mx; i++)
        p = putf(p, grid[((size_t)i * my + j) * mz + k]);
  return (size_t)(p - out);
}

int gco_xdr_decode(const uint8_t *buf, size_t len, gco_geom *g, char header[257], float **grid_out)
{
  const uint8_t *p = buf, *end = buf + len;
  uint32_t hl, n, type, dim;
  int d, i, j, k;
  float *grid;
#define NEED(nb) do { if ((size_t)(end - p) < (size_t)(nb)) return -1; } while (0)
  NEED(8);
  if (get32(p) != 2) return -2;                      /* version_check, xdr_grid.c:28-34 */
  p += 4;
  hl = get32(p); p += 4;
  if (hl > 256) return -1;
  NEED((hl + 3) & ~3u);
  memcpy(header, p, hl); header[hl] = '\0';
  p += (hl + 3) & ~3u;
  NEED(12);
  type = get32(p); p += 4; (void)type;
  g->tweight = getf(p); p += 4;
  dim = get32(p); p += 4;
  if (dim != 3) return -3;                           /* grid3D.c:81-83 */
  NEED(40);
  for (d = 0; d < 3; d++) { g->mx[d] = (int)get32(p); p += 4; }
  for (d = 0; d < 3; d++) { g->delta[d] = getf(p); p += 4; }
  for (d = 0; d < 3; d++) { g->a[d] = getf(p); p += 4; }
  n = get32(p); p += 4;
  if (n > 512u * 512u * 512u) return -1;             /* NGRID_MAX, xdr_grid.h:36 */
  if ((size_t)n != (size_t)g->mx[0] * g->mx[1] * g->mx[2]) return -4;
  NEED((size_t)n * 4);
  grid = malloc(sizeof(float) * (n ? n : 1));
  for (k = 0; k < g->mx[2]; k++)
    for (j = 0; j < g->mx[1]; j++)
      for (i = 0; i < g->mx[0]; i++) {
        grid[((size_t)i * g->mx[1] + j) * g->mx[2] + k] = getf(p);
        p += 4;
      }
  gco_update_b(g);
  *grid_out = grid;
  return 0;
#undef NEED
}

/* a_gridcalc.c:176-189: avg += g * (double)(tweight/sumT), f32 store per grid */
void gco_gridcalc(int ngrids, const float *const *grids, const float *tweights,
                  size_t cx, size_t cy, size_t cz, float *avg, float *tweight_out)
{
  double sumT = 0;
  size_t n = cx * cy * cz, c;
  int f;
  for (f = 0; f < ngrids; f++) sumT += (double)tweights[f];
  memset(avg, 0, sizeof(float) * n);
  for (f = 0; f < ngrids; f++) {
    double w = (double)tweights[f] / sumT;
    for (c = 0; c < n; c++) avg[c] = (float)((double)avg[c] + (double)grids[f][c] * w);
  }
  *tweight_out = (float)sumT;
}

/* plt.c:42-77, host-native byte order */
size_t gco_plt_encode(const gco_geom *g, const float *grid, float unit, uint8_t *out, size_t cap)
{
  size_t cells = (size_t)g->mx[0] * g->mx[1] * g->mx[2], need = 4 * (2 + 3 + 6 + cells);
  const float nrm = (float)(1. / (double)unit);
  uint8_t *p = out;
  int32_t iv;
  float fv;
  int d, i, j, k;
  if (need > cap) return 0;
  iv = 3; memcpy(p, &iv, 4); p += 4;
  iv = 42; memcpy(p, &iv, 4); p += 4;
  for (d = 2; d >= 0; d--) { iv = g->mx[d]; memcpy(p, &iv, 4); p += 4; }
  for (d = 2; d >= 0; d--) {
    fv = (float)(10. * (double)g->a[d]); memcpy(p, &fv, 4); p += 4;
    fv = (float)(10. * (double)g->b[d]); memcpy(p, &fv, 4); p += 4;
  }
  for (k = 0; k < g->mx[2]; k++)
    for (j = 0; j < g->mx[1]; j++)
      for (i = 0; i < g->mx[0]; i++) {
        fv = nrm * grid[((size_t)i * g->mx[1] + j) * g->mx[2] + k];
        memcpy(p, &fv, 4); p += 4;
      }
  return (size_t)(p - out);
}

/* ------------------------------------------------------------------ */
/* a_ri3Dc                                                             */
/* ------------------------------------------------------------------ */
static const float k_dens_unit[GCO_UNIT_NR] = { 1.000000f, 32.3227f, 0.602214f, 1000.0f, -1.0f }; /* count.c:21-22 */

void gco_update_b(gco_geom *g)                            /* a_ri3Dc.c:86-92 */
{
  int d;
  for (d = 0; d < 3; d++) {
    float ext = (float)g->mx[d] * g->delta[d];
    g->b[d] = g->a[d] + ext;
  }
}

void gco_tgrid2cavity(const gco_geom *g, gco_cavity *c)   /* a_ri3Dc.c:94-107 */
{
  int d;
  float v[3];
  c->axis[0] = 0; c->axis[1] = 0; c->axis[2] = 1;
  for (d = 0; d < 3; d++) {
    float s = g->a[d] + g->b[d];
    float w = g->b[d] - g->a[d];
    c->cpoint[d] = 0.5f * s;
    v[d] = 0.5f * w;
  }
  c->radius = minf_d(fabs((double)v[0]), fabs((double)v[1]));
  c->z1 = g->a[2] < g->b[2] ? g->a[2] : g->b[2];
  c->z2 = g->a[2] > g->b[2] ? g->a[2] : g->b[2];
  {
    float r2 = c->radius * c->radius, h = c->z2 - c->z1;
    c->vol = (float)(PI_D * (double)r2 * (double)h);
  }
}

static float discrete_adjust(float x_old, float x_new, float delta)   /* a_ri3Dc.c:128-130 */
{
  float q = (x_old - x_new) / delta;
  int n = -(int)((double)q + (x_old > x_new ? 0.5 : -0.5));
  return (float)n * delta;
}

int gco_readjust(gco_geom *g, float **grid, gco_cavity *user, int setmask)   /* a_ri3Dc.c:133-201 */
{
  gco_geom old = *g;
  float *og = *grid, *ng;
  float dsum = g->delta[0] + g->delta[1];
  float DeltaR = (float)(0.5 * (double)dsum);
  float R_new = user->radius;
  int changed = 0, base[3], d, i, j, k;
  if (setmask & GCO_SET_Z1) { g->a[2] += discrete_adjust(g->a[2], user->z1, g->delta[2]); changed = 1; }
  if (setmask & GCO_SET_Z2) { g->b[2] += discrete_adjust(g->b[2], user->z2, g->delta[2]); changed = 1; }
  if (g->a[2] >= g->b[2]) return -1;
  gco_tgrid2cavity(g, user);
  if (setmask & GCO_SET_R) { user->radius += discrete_adjust(user->radius, R_new, DeltaR); changed = 1; }
  if (!changed) return 0;
  if (!gco_setup_grid(g, user, old.delta)) return -2;
  for (d = 0; d < 3; d++) {
    base[d] = (int)((g->a[d] - old.a[d]) / g->delta[d]);
    if (base[d] < 0 || base[d] + g->mx[d] > old.mx[d]) return -3;   /* the asserts at :191-192 */
  }
  ng = malloc(sizeof(float) * (size_t)g->mx[0] * g->mx[1] * g->mx[2]);
  for (i = 0; i < g->mx[0]; i++)
    for (j = 0; j < g->mx[1]; j++)
      for (k = 0; k < g->mx[2]; k++)
        ng[((size_t)i * g->mx[1] + j) * g->mx[2] + k] =
          og[((size_t)(i + base[0]) * old.mx[1] + (j + base[1])) * old.mx[2] + (k + base[2])];
  free(og);
  *grid = ng;
  return 1;
}

void gco_analysis_free(gco_analysis *an)
{
  free(an->xyp); free(an->xzp); free(an->yzp); free(an->hxyp);
  free(an->zdf); free(an->lzdf); free(an->profile); free(an->rweight);
  free(an->rdf); free(an->hrdf); free(an->rzp); free(an->hrzp);
}

/* a_ri3Dc.c:486-762.  `g->b` must be current (gco_update_b / gco_readjust). */
int gco_analyse(const gco_geom *g, const float *grid, gco_analysis *an)
{
  const int mx = g->mx[0], my = g->mx[1], mz = g->mx[2];
  const float fmx = (float)mx, fmy = (float)my, fmz = (float)mz;
  const float Uunity = k_dens_unit[GCO_UNIT_UNITY];
  float U, min_occ, DeltaR, rcirc, average = 0, volume = 0;
  double dV, sumP = 0, sumH = 0;
  int i, j, k, irad, iradNR, ivol = 0;
  float d_xy, d_xz, d_yz, d_zdf;
  double h_inc, rw_inc;
  float *ci, *cj;

  gco_tgrid2cavity(g, &an->geometry);                    /* a_ri3Dc.c:486 */
  {
    float dsum = g->delta[0] + g->delta[1];
    DeltaR = (float)(0.5 * (double)dsum);                /* a_ri3Dc.c:488 */
  }
  dV = gco_delta_v(g);                                   /* a_ri3Dc.c:495 */

  min_occ = an->min_occ;
  for (k = 0; k < mz; k++)                               /* a_ri3Dc.c:499-508 */
    for (j = 0; j < my; j++)
      for (i = 0; i < mx; i++) {
        float v = grid[((size_t)i * my + j) * mz + k];
        sumP += (double)v;
        if (v <= min_occ) sumH++;
      }
  sumP *= dV; sumH *= dV;

  U = an->unit == GCO_UNIT_VOXEL ? (float)dV : k_dens_unit[an->unit];  /* a_ri3Dc.c:516 */
  min_occ = min_occ * U;                                 /* a_ri3Dc.c:517 */

  iradNR = (int)floor((double)(an->geometry.radius / DeltaR));   /* a_ri3Dc.c:544 */
  if (iradNR < 0) iradNR = 0;
  an->DeltaR = DeltaR; an->dV = dV; an->iradNR = iradNR; an->unit_value = U;
  an->sumP = sumP; an->sumH = sumH;

  an->xyp = calloc((size_t)mx * my, 4); an->xzp = calloc((size_t)mx * mz, 4);
  an->yzp = calloc((size_t)my * mz, 4); an->hxyp = calloc((size_t)mx * my, 4);
  an->zdf = calloc((size_t)mz * 2, 4); an->lzdf = calloc((size_t)mz * 2, 4);
  an->profile = calloc((size_t)mz * 2, 4);
  an->rweight = calloc((size_t)iradNR + 1, 4);
  an->rdf = calloc((size_t)iradNR * 2 + 2, 4); an->hrdf = calloc((size_t)iradNR * 2 + 2, 4);
  an->rzp = calloc((size_t)iradNR * mz + 1, 4); an->hrzp = calloc((size_t)iradNR * mz + 1, 4);

  d_xy = fmz * U; d_xz = fmy * U; d_yz = fmx * U;        /* a_ri3Dc.c:573-578 divisors */
  d_zdf = fmx * fmy * U;                                  /* a_ri3Dc.c:666 */
  h_inc = 1.0 / (double)(fmz * Uunity);                   /* a_ri3Dc.c:580 */
  rw_inc = 1.0 / (double)fmz;                             /* a_ri3Dc.c:637 */
  rcirc = an->geometry.radius / DeltaR;                   /* a_ri3Dc.c:592 */

  ci = malloc(sizeof(float) * (size_t)mx); cj = malloc(sizeof(float) * (size_t)my);
  for (i = 0; i < mx; i++) ci[i] = (float)(((double)(float)i + 0.5) - (double)fmx / 2.0);   /* a_ri3Dc.c:586 */
  for (j = 0; j < my; j++) cj[j] = (float)(((double)(float)j + 0.5) - (double)fmy / 2.0);

  for (k = 0; k < mz; k++) {
    float Rgyr2 = 0, Agyr;
    float *zd = an->zdf + 2 * (size_t)k, *lz = an->lzdf + 2 * (size_t)k, *pr = an->profile + 2 * (size_t)k;
    for (j = 0; j < my; j++) {
      for (i = 0; i < mx; i++) {
        const float v = grid[((size_t)i * my + j) * mz + k];
        const float c2 = ci[i] * ci[i] + cj[j] * cj[j];
        an->xyp[(size_t)i * my + j] += v / d_xy;
        an->xzp[(size_t)i * mz + k] += v / d_xz;
        an->yzp[(size_t)j * mz + k] += v / d_yz;
        if (v <= min_occ)
          an->hxyp[(size_t)i * my + j] = (float)((double)an->hxyp[(size_t)i * my + j] + h_inc);
        if (v > min_occ) {                                 /* a_ri3Dc.c:589-613 */
          const float cx = (float)((double)fmx / 2.0), cy = (float)((double)fmy / 2.0);
          const float ex = (float)i - cx, ey = (float)j - cy;
          if (ex * ex + ey * ey <= rcirc * rcirc) {
            volume++;
            lz[1] += v;
            Rgyr2 += v * c2;
          }
        }
        irad = (int)floor(sqrt((double)c2));               /* a_ri3Dc.c:625 */
        if (irad < iradNR) {                               /* a_ri3Dc.c:627-649 */
          average += v;
          ivol++;
          an->rweight[irad] = (float)((double)an->rweight[irad] + rw_inc);
          if (v > min_occ) {
            an->rdf[2 * irad + 1] += v;
            an->rzp[(size_t)irad * mz + k] += v;
          } else {
            an->hrdf[2 * irad + 1]++;
            an->hrzp[(size_t)irad * mz + k]++;
          }
        }
        zd[1] += v;                                        /* a_ri3Dc.c:661 */
      }
    }
    zd[0] = (float)((double)g->a[2] + ((double)k + 0.5) * (double)g->delta[2]);   /* a_ri3Dc.c:667 */
    zd[1] /= d_zdf;
    lz[0] = pr[0] = zd[0];
    Rgyr2 = (float)((double)Rgyr2 / (0.5 * (double)lz[1]));                        /* a_ri3Dc.c:674 */
    pr[1] = (float)(sqrt((double)Rgyr2) * (double)DeltaR);                         /* a_ri3Dc.c:677 */
    Agyr = (float)(PI_D * (double)Rgyr2);                                          /* a_ri3Dc.c:679 */
    {
      float den = g->delta[2] * Agyr * DeltaR * DeltaR * U;                        /* a_ri3Dc.c:695 */
      lz[1] = (float)((double)lz[1] * dV / (double)den);
    }
  }
  free(ci); free(cj);

  for (irad = 0; irad < iradNR; irad++) {                  /* a_ri3Dc.c:717-726 */
    const float rw = an->rweight[irad];
    an->rdf[2 * irad] = an->hrdf[2 * irad] = (float)irad * DeltaR;
    an->rdf[2 * irad + 1] /= rw * fmz * U;
    an->hrdf[2 * irad + 1] /= rw * fmz * Uunity;
    for (k = 0; k < mz; k++) {
      an->rzp[(size_t)irad * mz + k] /= rw * U;
      an->hrzp[(size_t)irad * mz + k] /= rw * Uunity;
    }
  }
  /* a_ri3Dc.c:731-757, including its `irad+1` indexing */
  if (an->rad_ba_step > 0)
    for (irad = 0; irad < an->rad_ba_nr; irad += an->rad_ba_step) {
      const float st = (float)an->rad_ba_step;
      for (i = 1; i < an->rad_ba_step; i++) {
        an->rdf[2 * irad + 1] += an->rdf[2 * (irad + i) + 1];
        an->hrdf[2 * irad + 1] += an->hrdf[2 * (irad + i) + 1];
        for (k = 0; k < mz; k++) {
          an->rzp[(size_t)irad * mz + k] += an->rzp[(size_t)(irad + 1) * mz + k];
          an->hrzp[(size_t)irad * mz + k] += an->hrzp[(size_t)(irad + 1) * mz + k];
        }
      }
      an->rdf[2 * irad + 1] /= st; an->hrdf[2 * irad + 1] /= st;
      for (k = 0; k < mz; k++) { an->rzp[(size_t)irad * mz + k] /= st; an->hrzp[(size_t)irad * mz + k] /= st; }
      for (i = 1; i < an->rad_ba_step; i++) {
        an->rdf[2 * (irad + 1) + 1] = an->rdf[2 * irad + 1];
        an->hrdf[2 * (irad + 1) + 1] = an->hrdf[2 * irad + 1];
        for (k = 0; k < mz; k++) {
          an->rzp[(size_t)(irad + 1) * mz + k] = an->rzp[(size_t)irad * mz + k];
          an->hrzp[(size_t)(irad + 1) * mz + k] = an->hrzp[(size_t)irad * mz + k];
        }
      }
    }

  volume = (float)((double)volume * dV);                   /* a_ri3Dc.c:759 */
  an->height = g->b[2] - g->a[2];
  an->reff = (float)sqrt((double)volume / (PI_D * (double)an->height));
  average /= (float)ivol * U;                              /* a_ri3Dc.c:762 */
  an->volume = volume; an->average = average; an->ivol = ivol;
  return 0;
}

/* ------------------------------------------------------------------ */
/* synthetic trajectories                                              */
/* ------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint32_t gen_u24(uint64_t seed, uint64_t frame, uint64_t atom, uint64_t lane)
{
  uint64_t key = seed ^ (frame * 0xD1342543DE82EF95ull) ^ (atom * 0xA24BAED4963EE407ull)
                 ^ (lane * 0x9FB21C651E98DF25ull);
  return (uint32_t)(splitmix64(key) >> 40);
}
static inline float u24_to_unit(uint32_t u) { return (float)u * 5.9604644775390625e-08f; /* 2^-24 */ }

void gco_gen_box(const gco_gen *p, long frame, float L_out[3])
{
  int d;
  float s = 1.0f;
  if (p->box_period > 0) {
    /* triangle wave in [-1, 1], exact in f32 */
    long ph = frame % p->box_period;
    long half = p->box_period / 2;
    float tri = ((float)labs(ph - half) - 0.5f * (float)half) / (0.5f * (float)half);
    s = 1.0f + p->box_amp * tri;
  }
  for (d = 0; d < 3; d++) L_out[d] = p->L[d] * s;
}

void gco_gen_frames(const gco_gen *p, long frame0, long nframes, int natoms, float *x)
{
  long f;
  int i, d;
  for (f = 0; f < nframes; f++) {
    const uint64_t fr = (uint64_t)(frame0 + f);
    float L[3];
    gco_gen_box(p, frame0 + f, L);
    for (i = 0; i < natoms; i++) {
      float *r = x + ((size_t)f * natoms + i) * 3;
      if (p->mode == GCO_GEN_CLUSTER) {
        /* Irwin-Hall(4) offset around one of nsites fixed centres, wrapped once */
        uint32_t site = gen_u24(p->seed, fr, (uint64_t)i, 3) % (uint32_t)p->nsites;
        for (d = 0; d < 3; d++) {
          float c = u24_to_unit(gen_u24(p->seed ^ 0x5151515151515151ull, 0, site, (uint64_t)d)) * L[d];
          uint32_t s4 = gen_u24(p->seed, fr, (uint64_t)i, 4 + (uint64_t)d) + gen_u24(p->seed, fr, (uint64_t)i, 8 + (uint64_t)d)
                        + gen_u24(p->seed, fr, (uint64_t)i, 12 + (uint64_t)d) + gen_u24(p->seed, fr, (uint64_t)i, 16 + (uint64_t)d);
          float off = ((float)s4 * 5.9604644775390625e-08f - 2.0f) * (p->sigma * 1.7320508f);
          float v = c + off;
          if (v < 0) v += L[d];
          if (v >= L[d]) v -= L[d];
          r[d] = v;
        }
        continue;
      }
      for (d = 0; d < 3; d++) r[d] = u24_to_unit(gen_u24(p->seed, fr, (uint64_t)i, (uint64_t)d)) * L[d];
      if (p->mode == GCO_GEN_PORE) {
        /* membrane slab [slab_lo, slab_hi) is empty except for a pore of radius pore_r */
        float dx = r[0] - 0.5f * L[0], dy = r[1] - 0.5f * L[1];
        float r2 = dx * dx + dy * dy;
        if (r[2] >= p->slab_lo && r[2] < p->slab_hi && r2 > p->pore_r * p->pore_r) {
          float thick = p->slab_hi - p->slab_lo;
          float z = u24_to_unit(gen_u24(p->seed, fr, (uint64_t)i, 3)) * (L[2] - thick);
          if (z >= p->slab_lo) z += thick;
          r[2] = z;
        }
      }
      if (p->mode == GCO_GEN_QUANTISED)
        for (d = 0; d < 3; d++) r[d] = rintf(r[d] * 1000.0f) * 0.001f;
    }
  }
}

/* ------------------------------------------------------------------ */
/* CPU port timing (bench.py cpu_baseline, kind "port")                */
/* ------------------------------------------------------------------ */
typedef struct {
  const gco_geom *g; float *grid; const float *x; long f0, f1; int natoms;
  const int *gindex; int gnx; float w; double sumP;
} port_job;

static void *port_worker(void *arg)
{
  port_job *j = arg;
  j->sumP = 0;
  gco_bin_frames(j->g, j->grid, j->x + (size_t)j->f0 * j->natoms * 3, j->f1 - j->f0, j->natoms,
                 j->gindex, j->gnx, NULL, &j->sumP, NULL, NULL);
  (void)j->w;
  return NULL;
}

double gco_bench(const gco_geom *g, const float *x, long nframes, int natoms,
                 const int *gindex, int gnx, float weight, int nthreads, double *sumP_out)
{
  size_t cells = (size_t)g->mx[0] * g->mx[1] * g->mx[2];
  port_job *jobs = calloc((size_t)nthreads, sizeof(*jobs));
  pthread_t *th = calloc((size_t)nthreads, sizeof(*th));
  struct timespec t0, t1;
  double s = 0;
  int t;
  for (t = 0; t < nthreads; t++) {
    jobs[t].g = g; jobs[t].grid = malloc(sizeof(float) * cells);
    memset(jobs[t].grid, 0, sizeof(float) * cells);
    jobs[t].x = x; jobs[t].natoms = natoms; jobs[t].gindex = gindex; jobs[t].gnx = gnx; jobs[t].w = weight;
    jobs[t].f0 = nframes * t / nthreads; jobs[t].f1 = nframes * (t + 1) / nthreads;
  }
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, port_worker, &jobs[t]);
  port_worker(&jobs[0]);
  for (t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  for (t = 0; t < nthreads; t++) { s += jobs[t].sumP; free(jobs[t].grid); }
  if (sumP_out) *sumP_out = s;
  free(jobs); free(th);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
