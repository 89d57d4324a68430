// Host-side geometry, frame weights and the header string of the g_ri3Dc path.
//
// These run once per job and decide where every bin edge lies, so each
// expression keeps the reference's operand types and rounding order
// (real = float, double constants promote): compile with -ffp-contract=off.
// Citations are file:line in the reference tree.
#include "gcb_internal.h"
#include <cmath>
#include <cstdlib>

namespace gcb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int fail(int code, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// Closed form of `for (n times) v += w;`.  Inside one binade every rounded
// increment is the same number of ulps once one addition has happened there
// (round-half-even can only alter the first in-binade step), so the chain is
// advanced a binade at a time; the steps that cross binades are real additions.
float seq_add_f32(float v, float w, uint64_t n)
{
  while (n > 0) {
    const float v1 = v + w;
    --n;
    if (v1 == v || v1 != v1) return v1;
    v = v1;
    if (n < 2) continue;
    const float v2 = v + w, v3 = v2 + w;
    uint32_t b1, b2, b3;
    memcpy(&b1, &v, 4); memcpy(&b2, &v2, 4); memcpy(&b3, &v3, 4);
    if (v > 0 && w > 0 && (b1 >> 23) == (b2 >> 23) && (b2 >> 23) == (b3 >> 23) && b3 > b2) {
      const uint32_t inc = b3 - b2, last = b2 | 0x7FFFFFu;
      --n;
      uint64_t k = (uint64_t)(last - b2) / inc;
      if (k > n) k = n;
      b2 += (uint32_t)(k * inc);
      n -= k;
      memcpy(&v, &b2, 4);
    }
  }
  return v;
}

double seq_add_f64(double v, double w, uint64_t n)
{
  while (n > 0) {
    const double v1 = v + w;
    --n;
    if (v1 == v || v1 != v1) return v1;
    v = v1;
    if (n < 2) continue;
    const double v2 = v + w, v3 = v2 + w;
    uint64_t b1, b2, b3;
    memcpy(&b1, &v, 8); memcpy(&b2, &v2, 8); memcpy(&b3, &v3, 8);
    if (v > 0 && w > 0 && (b1 >> 52) == (b2 >> 52) && (b2 >> 52) == (b3 >> 52) && b3 > b2) {
      const uint64_t inc = b3 - b2, last = b2 | 0xFFFFFFFFFFFFFull;
      --n;
      uint64_t k = (last - b2) / inc;
      if (k > n) k = n;
      b2 += k * inc;
      n -= k;
      memcpy(&v, &b2, 8);
    }
  }
  return v;
}

}  // namespace gcb

static const double kPi = 3.14159265358979323846;   // M_PI (utilgmx.h:24-28)

extern "C" {

const char *gcb_last_error(void) { return gcb::g_err; }
int gcb_abi_version(void) { return GCB_ABI_VERSION; }
void gcb_free(void *p) { free(p); }

int gcb_autoset_cavity(gcb_cavity *cav, const float box[9], int setmask)
{
  if (!cav || !box) return gcb::fail(GCB_ERR_ARG, "gcb_autoset_cavity: null argument");
  const float L[3] = {box[0], box[4], box[8]};
  if (!(setmask & GCB_SET_CPOINT))                       // count.c:28-33
    for (int d = 0; d < 3; d++) cav->cpoint[d] = (float)(0.5 * (double)L[d]);
  if (!(setmask & GCB_SET_R)) {                          // count.c:35-40
    double lim[2];
    for (int d = 0; d < 2; d++) {
      const float off = cav->cpoint[d] - L[d];
      const double far = std::fabs((double)off);
      lim[d] = (double)cav->cpoint[d] < far ? (double)cav->cpoint[d] : far;
    }
    cav->radius = (float)(lim[0] < lim[1] ? lim[0] : lim[1]);
  }
  if (!(setmask & GCB_SET_Z1)) cav->z1 = 0;              // count.c:42-46
  if (!(setmask & GCB_SET_Z2)) cav->z2 = L[2];           // count.c:47-51
  if (cav->z1 > cav->z2) { const float t = cav->z1; cav->z1 = cav->z2; cav->z2 = t; }   // count.c:53-59
  const float height = cav->z2 - cav->z1;
  cav->vol = (float)(kPi * (double)cav->radius * (double)cav->radius * (double)height);  // count.c:60
  cav->axis[0] = 0; cav->axis[1] = 0; cav->axis[2] = 1;
  return GCB_OK;
}

int gcb_setup_tgrid(gcb_tgrid *tg, gcb_cavity *cav, const float delta[3])
{
  if (!tg || !cav || !delta) return gcb::fail(GCB_ERR_ARG, "gcb_setup_tgrid: null argument");
  for (int d = 0; d < 3; d++)
    if (!(delta[d] > 0)) return gcb::fail(GCB_ERR_ARG, "gcb_setup_tgrid: delta[%d] must be > 0", d);
  const float R = cav->radius, D = 2 * R;                // edge vectors rx, ry (grid3D.c:182-183)
  // lower corner c[0] (grid3D.c:190-192)
  float lo[3] = {cav->cpoint[0] - R, cav->cpoint[1] - R, cav->z1};
  // upper corner: c[4] -> -ry -> +rx -> +ry (grid3D.c:196-201); the zero steps are identities in f32
  float hi[3] = {cav->cpoint[0] - R, cav->cpoint[1] + R, cav->z2};
  hi[1] = hi[1] - D;
  hi[0] = hi[0] + D;
  hi[1] = hi[1] + D;
  hi[2] = hi[2] + 0.0f;
  const double slack = 100 * GCB_GMX_REAL_EPS;
  for (int d = 0; d < 3; d++) {
    const float width = hi[d] - lo[d];
    const int m = (int)std::floor((double)(width / delta[d]) + slack);   // grid3D.c:143
    const float covered = (float)m * delta[d];
    float du = (float)(0.5 * (double)(width - covered));                 // grid3D.c:145
    if ((double)(float)std::fabs((double)du) < slack) du = 0;            // grid3D.c:146
    tg->mx[d] = m;
    tg->delta[d] = delta[d];
    tg->a[d] = lo[d] + du;
    tg->b[d] = hi[d] - du;
  }
  cav->z1 = tg->a[2];                                                    // grid3D.c:153-156
  cav->z2 = tg->b[2];
  const double wx = std::fabs((double)(tg->b[0] - tg->a[0])), wy = std::fabs((double)(tg->b[1] - tg->a[1]));
  cav->radius = (float)(0.5 * (wx < wy ? wx : wy));
  if (tg->mx[0] <= 0 || tg->mx[1] <= 0 || tg->mx[2] <= 0)
    return gcb::fail(GCB_ERR_RANGE, "gcb_setup_tgrid: empty grid %d x %d x %d", tg->mx[0], tg->mx[1], tg->mx[2]);
  return GCB_OK;
}

double gcb_delta_v(const gcb_tgrid *tg)
{
  return (double)tg->delta[0] * (double)tg->delta[1] * (double)tg->delta[2];   // grid3D.c:108-110
}

void gcb_update_tgrid(gcb_tgrid *tg)                     // a_ri3Dc.c:86-92
{
  for (int d = 0; d < 3; d++) {
    const float extent = (float)tg->mx[d] * tg->delta[d];
    tg->b[d] = tg->a[d] + extent;
  }
}

void gcb_tgrid2cavity(const gcb_tgrid *tg, gcb_cavity *cav)   // a_ri3Dc.c:94-107
{
  float half[3];
  for (int d = 0; d < 3; d++) {
    const float sum = tg->a[d] + tg->b[d], diff = tg->b[d] - tg->a[d];
    cav->cpoint[d] = 0.5f * sum;
    half[d] = 0.5f * diff;
  }
  cav->axis[0] = 0; cav->axis[1] = 0; cav->axis[2] = 1;
  const double hx = std::fabs((double)half[0]), hy = std::fabs((double)half[1]);
  cav->radius = (float)(hx < hy ? hx : hy);
  cav->z1 = tg->a[2] < tg->b[2] ? tg->a[2] : tg->b[2];
  cav->z2 = tg->a[2] > tg->b[2] ? tg->a[2] : tg->b[2];
  const float r2 = cav->radius * cav->radius, h = cav->z2 - cav->z1;
  cav->vol = (float)(kPi * (double)r2 * (double)h);
}

int gcb_edge_vulnerable(const gcb_tgrid *tg)
{
  for (int d = 0; d < 3; d++) {
    const float x = std::nextafterf(tg->b[d], -INFINITY);
    const float u = x - tg->a[d];
    if (u >= 0 && !(u / tg->delta[d] < (float)tg->mx[d])) return 1;
  }
  return 0;
}

int gcb_frame_weights(const float *t, long nframes, int dtweight, float *tlast, float *weight_out,
                      double *T_tot_inout)
{
  if (!t || nframes < 0 || !weight_out || !T_tot_inout)
    return gcb::fail(GCB_ERR_ARG, "gcb_frame_weights: null argument");
  float last;
  if (tlast) last = *tlast;
  else {
    // get_timestep + `tlast = t - dt` (g_ri3Dc.c:155-177, 390-394)
    const float dt0 = nframes > 1 ? t[1] - t[0] : 0.0f;
    last = nframes > 0 ? t[0] - dt0 : 0.0f;
  }
  double T = *T_tot_inout;
  for (long f = 0; f < nframes; f++) {
    const float dt = t[f] - last;                        // g_ri3Dc.c:413
    last = t[f];
    T += (double)dt;                                     // g_ri3Dc.c:415
    weight_out[f] = dtweight ? dt : 1.0f;                // g_ri3Dc.c:416
  }
  if (tlast) *tlast = last;
  *T_tot_inout = T;
  return GCB_OK;
}

int gcb_header(char out[GCB_HEADER_MAX], const char *subtitle, double T_tot, const char *grpname,
               const gcb_cavity *cav, const gcb_tgrid *tg, double sumP, const char *trxname, int dtweight)
{
  if (!out || !cav || !tg) return gcb::fail(GCB_ERR_ARG, "gcb_header: null argument");
  const float dsum = tg->delta[0] + tg->delta[1] + tg->delta[2];
  // field list and formats of g_ri3Dc.c:463-472 (part of the file format: a_ri3Dc
  // copies it into every plot as the subtitle)
  snprintf(out, GCB_HEADER_MAX,
           "%s{T}%gps{grp}%s{z1}%.2f{z2}%.2f{Rmax}%.2f{<Delta>}%.3f{<N>}%.2f{XTC}%s"
           "{Delta}(%.3f,%.3f,%.3f){sumP}%g%s",
           subtitle ? subtitle : "", T_tot, grpname ? grpname : "", cav->z1, cav->z2, cav->radius,
           (double)dsum / 3., sumP / T_tot, trxname ? trxname : "",
           tg->delta[0], tg->delta[1], tg->delta[2], sumP, dtweight ? "ps" : "");
  out[GCB_HEADER_MAX - 1] = '\0';
  return GCB_OK;
}

float gcb_dens_unit(int unit, const gcb_tgrid *tg)
{
  static const float table[GCB_UNIT_NR] = {1.000000f, 32.3227f, 0.602214f, 1000.0f, -1.0f};   // count.c:21-22
  if (unit == GCB_UNIT_VOXEL && tg) return (float)gcb_delta_v(tg);                            // a_ri3Dc.c:516
  if (unit < 0 || unit >= GCB_UNIT_NR) return 1.0f;
  return table[unit];
}

}  // extern "C"
