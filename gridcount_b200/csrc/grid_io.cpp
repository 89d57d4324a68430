// Grid files of the gridcount path: the XDR density file (version 2), the
// gOpenMol plt dump and the ascii dump.  Byte-exact with what the reference
// writes (grid3D.c:22-106, xdr_grid.c:89-158, plt.c:42-99); the wire format is
// RFC 4506 XDR: big-endian 4-byte units, strings as length + bytes + zero pad.
#include "gcb_internal.h"
#include <cstdlib>
#include <vector>

namespace {

inline uint32_t bswap(uint32_t w) { return __builtin_bswap32(w); }
inline uint32_t f2u(float f) { uint32_t w; memcpy(&w, &f, 4); return w; }
inline float u2f(uint32_t w) { float f; memcpy(&f, &w, 4); return f; }

struct Writer {
  uint8_t *p;
  void u32(uint32_t w) { w = bswap(w); memcpy(p, &w, 4); p += 4; }
  void f32(float f) { u32(f2u(f)); }
  void str(const char *s, size_t len)
  {
    u32((uint32_t)len);
    memcpy(p, s, len);
    const size_t padded = (len + 3) & ~(size_t)3;
    memset(p + len, 0, padded - len);
    p += padded;
  }
};

struct Reader {
  const uint8_t *p, *end;
  bool have(size_t n) const { return (size_t)(end - p) >= n; }
  uint32_t u32() { uint32_t w; memcpy(&w, p, 4); p += 4; return bswap(w); }
  float f32() { return u2f(u32()); }
};

// z-fastest memory -> x-fastest file order (grid3_serialise, xdr_grid.c:113-126), big-endian words
void serialise_be(const gcb_tgrid &tg, const float *zfast, uint8_t *out)
{
  const int mx = tg.mx[0], my = tg.mx[1], mz = tg.mx[2];
  uint32_t *w = reinterpret_cast<uint32_t *>(out);   // out is 4-byte aligned relative to the record start
  // tile over (i,k) to keep both sides cache friendly on 256-500 MB grids
  const int T = 32;
  for (int j = 0; j < my; j++)
    for (int k0 = 0; k0 < mz; k0 += T)
      for (int i0 = 0; i0 < mx; i0 += T) {
        const int k1 = k0 + T < mz ? k0 + T : mz, i1 = i0 + T < mx ? i0 + T : mx;
        for (int k = k0; k < k1; k++)
          for (int i = i0; i < i1; i++) {
            uint32_t v = bswap(f2u(zfast[((size_t)i * my + j) * mz + k]));
            memcpy(&w[((size_t)k * my + j) * mx + i], &v, 4);
          }
      }
}

}  // namespace

extern "C" {

size_t gcb_xdr_size(const gcb_tgrid *tg, const char *header)
{
  const size_t len = header ? strlen(header) : 0;
  return 4 + 4 + ((len + 3) & ~(size_t)3) + 4 + 4 + 4 + 3 * 4 + 3 * 4 + 3 * 4 + 4 + 4 * gcb::cells(*tg);
}

// everything of the xdr_grid_v2 record (xdr_grid.c:93-106) in front of the cells; returns its length
size_t gcb_xdr_prologue_(const gcb_tgrid *tg, const char *header, uint8_t *out)
{
  Writer w{out};
  w.u32(GCB_GRID_FF_VERSION);
  w.str(header, strlen(header));
  w.u32(0);                       // egtyREGULAR
  w.f32(tg->tweight);
  w.u32(3);                       // dim
  for (int d = 0; d < 3; d++) w.u32((uint32_t)tg->mx[d]);
  for (int d = 0; d < 3; d++) w.f32(tg->delta[d]);
  for (int d = 0; d < 3; d++) w.f32(tg->a[d]);
  w.u32((uint32_t)gcb::cells(*tg));
  return (size_t)(w.p - out);
}

// the 44 bytes in front of a plt file's cells (plt.c:52-67), host byte order
size_t gcb_plt_prologue_(const gcb_tgrid *tg, uint8_t *out)
{
  const int32_t head[5] = {3, 42, tg->mx[2], tg->mx[1], tg->mx[0]};   // rank, surface type, Z/Y/X dims
  float corners[6];
  for (int n = 0, d = 2; d >= 0; d--) {                              // Angstrom, Z then Y then X
    corners[n++] = (float)(10. * (double)tg->a[d]);
    corners[n++] = (float)(10. * (double)tg->b[d]);
  }
  memcpy(out, head, 20);
  memcpy(out + 20, corners, 24);
  return 44;
}

int gcb_xdr_encode(const gcb_tgrid *tg, const float *grid_zfast, const float *grid_xfast, const char *header,
                   uint8_t *out, size_t cap, size_t *written)
{
  if (!tg || !out || (!grid_zfast && !grid_xfast)) return gcb::fail(GCB_ERR_ARG, "gcb_xdr_encode: null argument");
  if (!header) header = "";
  const size_t len = strlen(header), n = gcb::cells(*tg);
  if (len > GCB_HEADER_MAX) return gcb::fail(GCB_ERR_ARG, "header longer than %d bytes", GCB_HEADER_MAX);
  if (n > GCB_NGRID_MAX) return gcb::fail(GCB_ERR_RANGE, "grid has %zu cells; the format allows %u", n, GCB_NGRID_MAX);
  const size_t need = gcb_xdr_size(tg, header);
  if (cap < need) return gcb::fail(GCB_ERR_ARG, "gcb_xdr_encode: buffer of %zu bytes, need %zu", cap, need);
  Writer w{out + gcb_xdr_prologue_(tg, header, out)};
  if (grid_xfast) {
    for (size_t c = 0; c < n; c++) { uint32_t v = bswap(f2u(grid_xfast[c])); memcpy(w.p + 4 * c, &v, 4); }
  } else {
    serialise_be(*tg, grid_zfast, w.p);
  }
  w.p += 4 * n;
  if (written) *written = (size_t)(w.p - out);
  return GCB_OK;
}

int gcb_grid_write(const char *path, const gcb_tgrid *tg, const float *grid_zfast, const float *grid_xfast,
                   const char *header)
{
  if (!path || !tg) return gcb::fail(GCB_ERR_ARG, "gcb_grid_write: null argument");
  const size_t need = gcb_xdr_size(tg, header);
  std::vector<uint8_t> buf(need);
  size_t n = 0;
  int rc = gcb_xdr_encode(tg, grid_zfast, grid_xfast, header, buf.data(), need, &n);
  if (rc) return rc;
  FILE *fp = fopen(path, "wb");
  if (!fp) return gcb::fail(GCB_ERR_IO, "cannot open %s for writing", path);
  const bool ok = fwrite(buf.data(), 1, n, fp) == n;
  if (fclose(fp) != 0 || !ok) return gcb::fail(GCB_ERR_IO, "short write to %s", path);
  return GCB_OK;
}

int gcb_grid_read(const char *path, gcb_tgrid *tg, char header[GCB_HEADER_MAX + 1], float **grid_zfast)
{
  if (!path || !tg || !grid_zfast) return gcb::fail(GCB_ERR_ARG, "gcb_grid_read: null argument");
  FILE *fp = fopen(path, "rb");
  if (!fp) return gcb::fail(GCB_ERR_IO, "cannot open %s", path);
  fseek(fp, 0, SEEK_END);
  const long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  std::vector<uint8_t> buf((size_t)(sz > 0 ? sz : 0));
  const bool ok = sz > 0 && fread(buf.data(), 1, (size_t)sz, fp) == (size_t)sz;
  fclose(fp);
  if (!ok) return gcb::fail(GCB_ERR_IO, "cannot read %s", path);

  Reader r{buf.data(), buf.data() + buf.size()};
  if (!r.have(8)) return gcb::fail(GCB_ERR_FORMAT, "%s: truncated", path);
  const uint32_t version = r.u32();
  if (version != GCB_GRID_FF_VERSION)     // version_check rejects both older and newer (xdr_grid.c:28-34)
    return gcb::fail(GCB_ERR_FORMAT, "%s: grid file format version %u, this library reads version %d", path,
                     version, GCB_GRID_FF_VERSION);
  const uint32_t hl = r.u32();
  const size_t hpad = ((size_t)hl + 3) & ~(size_t)3;
  if (hl > GCB_HEADER_MAX || !r.have(hpad)) return gcb::fail(GCB_ERR_FORMAT, "%s: bad header length %u", path, hl);
  if (header) { memcpy(header, r.p, hl); header[hl] = '\0'; }
  r.p += hpad;
  if (!r.have(12)) return gcb::fail(GCB_ERR_FORMAT, "%s: truncated", path);
  (void)r.u32();                          // grid type
  tg->tweight = r.f32();
  const uint32_t dim = r.u32();
  if (dim != 3) return gcb::fail(GCB_ERR_FORMAT, "%s: dim=%u, only 3D grids are supported", path, dim);   // grid3D.c:81-83
  if (!r.have(40)) return gcb::fail(GCB_ERR_FORMAT, "%s: truncated", path);
  for (int d = 0; d < 3; d++) tg->mx[d] = (int)r.u32();
  for (int d = 0; d < 3; d++) tg->delta[d] = r.f32();
  for (int d = 0; d < 3; d++) tg->a[d] = r.f32();
  const uint32_t n = r.u32();
  if (n > GCB_NGRID_MAX || tg->mx[0] <= 0 || tg->mx[1] <= 0 || tg->mx[2] <= 0 || (size_t)n != gcb::cells(*tg) ||
      !r.have((size_t)n * 4))
    return gcb::fail(GCB_ERR_FORMAT, "%s: grid of %u cells does not match %d x %d x %d or the file size", path, n,
                     tg->mx[0], tg->mx[1], tg->mx[2]);
  float *g = (float *)malloc(sizeof(float) * (size_t)n);
  if (!g) return gcb::fail(GCB_ERR_NOMEM, "out of memory for %u cells", n);
  const int mx = tg->mx[0], my = tg->mx[1], mz = tg->mx[2];
  // grid3_unserialise (xdr_grid.c:146-158): file is x fastest
  const int T = 32;
  for (int j = 0; j < my; j++)
    for (int k0 = 0; k0 < mz; k0 += T)
      for (int i0 = 0; i0 < mx; i0 += T) {
        const int k1 = k0 + T < mz ? k0 + T : mz, i1 = i0 + T < mx ? i0 + T : mx;
        for (int i = i0; i < i1; i++)
          for (int k = k0; k < k1; k++) {
            uint32_t v;
            memcpy(&v, r.p + 4 * (((size_t)k * my + j) * mx + i), 4);
            g[((size_t)i * my + j) * mz + k] = u2f(bswap(v));
          }
      }
  gcb_update_tgrid(tg);                   // b is not stored (a_ri3Dc.c:86-92)
  *grid_zfast = g;
  return GCB_OK;
}

int gcb_plt_write(const char *path, const gcb_tgrid *tg, const float *grid, float unit)
{
  if (!path || !tg || !grid) return gcb::fail(GCB_ERR_ARG, "gcb_plt_write: null argument");
  FILE *fp = fopen(path, "wb");
  if (!fp) return gcb::fail(GCB_ERR_IO, "cannot open %s for writing", path);
  const int mx = tg->mx[0], my = tg->mx[1], mz = tg->mx[2];
  const float scale = (float)(1. / (double)unit);                 // `real norm = 1./unit` (plt.c:48)
  uint8_t pro[44];
  gcb_plt_prologue_(tg, pro);                                     // plt.c:52-67
  bool ok = fwrite(pro, 1, sizeof(pro), fp) == sizeof(pro);
  std::vector<float> row((size_t)mx);
  for (int k = 0; k < mz && ok; k++)
    for (int j = 0; j < my && ok; j++) {
      for (int i = 0; i < mx; i++) row[i] = scale * grid[((size_t)i * my + j) * mz + k];
      ok = fwrite(row.data(), 4, (size_t)mx, fp) == (size_t)mx;
    }
  if (fclose(fp) != 0 || !ok) return gcb::fail(GCB_ERR_IO, "short write to %s", path);
  return GCB_OK;
}

int gcb_ascii_write(const char *path, const gcb_tgrid *tg, const float *grid, float unit)
{
  if (!path || !tg || !grid) return gcb::fail(GCB_ERR_ARG, "gcb_ascii_write: null argument");
  FILE *fp = fopen(path, "w");
  if (!fp) return gcb::fail(GCB_ERR_IO, "cannot open %s for writing", path);
  const int mx = tg->mx[0], my = tg->mx[1], mz = tg->mx[2];
  const float scale = (float)(1. / (double)unit);
  for (int k = 0; k < mz; k++) {                                  // plt.c:89-97
    for (int j = 0; j < my; j++) {
      for (int i = 0; i < mx; i++) fprintf(fp, "%.6f ", scale * grid[((size_t)i * my + j) * mz + k]);
      fprintf(fp, "\n");
    }
    fprintf(fp, "\n");
  }
  if (fclose(fp) != 0) return gcb::fail(GCB_ERR_IO, "short write to %s", path);
  return GCB_OK;
}

}  // extern "C"
