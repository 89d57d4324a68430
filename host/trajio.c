/*
 * Trajectory input for the host tools: XTC (Gromacs' xdr3dfcoord compression), the raw GCTRJ1
 * format of the test harness, and multi-frame .gro.
 *
 * The XTC codec is a from-scratch statement of the published xtc/xdr3dfcoord format (Gromacs
 * libxdrf / xdrfile; not under the reference tree and not installed here, so PARITY WITH GROMACS
 * IS UNPINNED -- tests cover write->read round trips and the quantisation invariants only):
 *   frame  := int32 magic(1995) natoms step | f32 time | f32 box[9] | coords
 *   coords := int32 natoms | natoms<=9: 3*natoms f32
 *             else f32 precision | int32 minint[3] maxint[3] smallidx | int32 nbytes | bytes, padded to x4
 * all big-endian 4-byte units.  Coordinates are integers round(x*precision); each atom is either
 * a full triple packed as one mixed-radix number over the bounding-box extents, or part of a run
 * of small differences to its predecessor packed over magicints[smallidx]^3; the first two atoms
 * of a run are swapped (water: O first, then the H closest to it).  Decoding multiplies by
 * fl(1/precision) in f32.
 */
#include "gcb_host.h"
#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ bit packing */
static const int magicints[] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812,
    1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015,
    65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122,
    1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216};
#define FIRSTIDX 9
#define LASTIDX ((int)(sizeof magicints / sizeof magicints[0]) - 1)

typedef struct { uint8_t *p; size_t cap, cnt; unsigned lastbits, lastbyte; int err; } bitbuf;

static void put_bits(bitbuf *b, int nbits, unsigned v)
{
  while (nbits >= 8) {
    b->lastbyte = (b->lastbyte << 8) | ((v >> (nbits - 8)) & 0xffu);
    if (b->cnt >= b->cap) { b->err = 1; return; }
    b->p[b->cnt++] = (uint8_t)(b->lastbyte >> b->lastbits);
    nbits -= 8;
  }
  if (nbits > 0) {
    b->lastbyte = (b->lastbyte << nbits) | (v & ((1u << nbits) - 1u));
    b->lastbits += (unsigned)nbits;
    if (b->lastbits >= 8) {
      b->lastbits -= 8;
      if (b->cnt >= b->cap) { b->err = 1; return; }
      b->p[b->cnt++] = (uint8_t)(b->lastbyte >> b->lastbits);
    }
  }
}

static size_t finish_bits(bitbuf *b)
{
  if (b->lastbits > 0) {
    if (b->cnt >= b->cap) { b->err = 1; return b->cnt; }
    b->p[b->cnt++] = (uint8_t)(b->lastbyte << (8 - b->lastbits));
  }
  return b->cnt;
}

static unsigned get_bits(bitbuf *b, int nbits)
{
  const unsigned mask = nbits >= 32 ? 0xffffffffu : ((1u << nbits) - 1u);
  unsigned v = 0;
  while (nbits >= 8) {
    if (b->cnt >= b->cap) { b->err = 1; return 0; }
    b->lastbyte = (b->lastbyte << 8) | b->p[b->cnt++];
    v |= (b->lastbyte >> b->lastbits) << (nbits - 8);
    nbits -= 8;
  }
  if (nbits > 0) {
    if (b->lastbits < (unsigned)nbits) {
      if (b->cnt >= b->cap) { b->err = 1; return 0; }
      b->lastbits += 8;
      b->lastbyte = (b->lastbyte << 8) | b->p[b->cnt++];
    }
    b->lastbits -= (unsigned)nbits;
    v |= (b->lastbyte >> b->lastbits) & ((1u << nbits) - 1u);
  }
  return v & mask;
}

static int bits_for(unsigned size)            /* bits needed for values 0 .. size-1 ... as the format defines it */
{
  unsigned num = 1;
  int nbits = 0;
  while (size >= num && nbits < 32) { nbits++; num <<= 1; }
  return nbits;
}

static int bits_for_triple(const unsigned sizes[3])     /* bits needed for sizes[0]*sizes[1]*sizes[2] */
{
  unsigned __int128 prod = (unsigned __int128)sizes[0] * sizes[1] * sizes[2];
  /* number of bytes below the top byte, then bits of the top byte, as the byte-wise original does */
  int nbytes = 0, nbits = 0;
  unsigned __int128 t = prod;
  unsigned top, num = 1;
  while (t >> 8) { t >>= 8; nbytes++; }
  top = (unsigned)t;
  while (top >= num) { nbits++; num *= 2; }
  return nbits + nbytes * 8;
}

/* one mixed-radix number v0*s1*s2 + v1*s2 + v2, sent low byte first */
static void put_triple(bitbuf *b, int nbits, const unsigned sizes[3], const unsigned v[3])
{
  unsigned __int128 num = ((unsigned __int128)v[0] * sizes[1] + v[1]) * sizes[2] + v[2];
  uint8_t bytes[16];
  int nbytes = 0, i;
  do { bytes[nbytes++] = (uint8_t)(num & 0xff); num >>= 8; } while (num);
  if (nbits >= nbytes * 8) {
    for (i = 0; i < nbytes; i++) put_bits(b, 8, bytes[i]);
    put_bits(b, nbits - nbytes * 8, 0);
  } else {
    for (i = 0; i < nbytes - 1; i++) put_bits(b, 8, bytes[i]);
    put_bits(b, nbits - (nbytes - 1) * 8, bytes[i]);
  }
}

static void get_triple(bitbuf *b, int nbits, const unsigned sizes[3], int v[3])
{
  unsigned __int128 num = 0;
  int shift = 0;
  while (nbits > 8) { num |= (unsigned __int128)get_bits(b, 8) << shift; shift += 8; nbits -= 8; }
  if (nbits > 0) num |= (unsigned __int128)get_bits(b, nbits) << shift;
  v[2] = (int)(num % sizes[2]); num /= sizes[2];
  v[1] = (int)(num % sizes[1]); num /= sizes[1];
  v[0] = (int)num;
}

/* ------------------------------------------------------------------ big-endian words */
static void be_put32(uint8_t *p, uint32_t w) { p[0] = (uint8_t)(w >> 24); p[1] = (uint8_t)(w >> 16); p[2] = (uint8_t)(w >> 8); p[3] = (uint8_t)w; }
static uint32_t be_get32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static void be_putf(uint8_t *p, float f) { uint32_t w; memcpy(&w, &f, 4); be_put32(p, w); }
static float be_getf(const uint8_t *p) { uint32_t w = be_get32(p); float f; memcpy(&f, &w, 4); return f; }

#define XTC_MAGIC 1995
#define XTC_HEADER_BYTES (4 * (3 + 1 + 9 + 1))

/* ------------------------------------------------------------------ encoder */
static int iabs(int v) { return v < 0 ? -v : v; }

size_t gh_xtc_encode_frame(int natoms, int step, float time, const float box9[9], const float *x, float prec,
                           uint8_t *out, size_t cap)
{
  size_t off = 0;
  int i, d;
  if (natoms < 0 || cap < (size_t)XTC_HEADER_BYTES + 64) return 0;
  be_put32(out + off, XTC_MAGIC); off += 4;
  be_put32(out + off, (uint32_t)natoms); off += 4;
  be_put32(out + off, (uint32_t)step); off += 4;
  be_putf(out + off, time); off += 4;
  for (i = 0; i < 9; i++) { be_putf(out + off, box9[i]); off += 4; }
  be_put32(out + off, (uint32_t)natoms); off += 4;
  if (natoms <= 9) {
    if (cap < off + (size_t)natoms * 12) return 0;
    for (i = 0; i < natoms * 3; i++) { be_putf(out + off, x[i]); off += 4; }
    return off;
  }
  if (!(prec > 0)) prec = 1000.0f;
  int *ip = (int *)malloc(sizeof(int) * 3 * (size_t)natoms);
  int minint[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, maxint[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
  int mindiff = INT32_MAX, oldl[3] = {0, 0, 0};
  if (!ip) return 0;
  for (i = 0; i < natoms; i++) {
    int l[3];
    for (d = 0; d < 3; d++) {
      const float lf = x[3 * i + d] * prec;
      if (fabsf(lf) > 2147483000.0f) { free(ip); return 0; }       /* does not fit the integer range */
      l[d] = lf >= 0 ? (int)(lf + 0.5f) : (int)(lf - 0.5f);
      if (l[d] < minint[d]) minint[d] = l[d];
      if (l[d] > maxint[d]) maxint[d] = l[d];
      ip[3 * i + d] = l[d];
    }
    if (i > 0) {
      const int diff = iabs(oldl[0] - l[0]) + iabs(oldl[1] - l[1]) + iabs(oldl[2] - l[2]);
      if (diff < mindiff) mindiff = diff;
    }
    oldl[0] = l[0]; oldl[1] = l[1]; oldl[2] = l[2];
  }
  unsigned sizeint[3], bitsizeint[3] = {0, 0, 0};
  int bitsize;
  for (d = 0; d < 3; d++) {
    if ((double)maxint[d] - (double)minint[d] >= 2147483645.0) { free(ip); return 0; }
    sizeint[d] = (unsigned)(maxint[d] - minint[d]) + 1u;
  }
  if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) {
    for (d = 0; d < 3; d++) bitsizeint[d] = (unsigned)bits_for(sizeint[d]);
    bitsize = 0;                                   /* the three components are sent separately */
  } else {
    bitsize = bits_for_triple(sizeint);
  }
  int smallidx = FIRSTIDX;
  while (smallidx < LASTIDX && magicints[smallidx] < mindiff) smallidx++;
  const int maxidx = smallidx + 8 < LASTIDX ? smallidx + 8 : LASTIDX, minidx = maxidx - 8;
  int smaller = magicints[smallidx - 1 > FIRSTIDX ? smallidx - 1 : FIRSTIDX] / 2;
  int smallnum = magicints[smallidx] / 2;
  unsigned sizesmall[3];
  sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
  const int larger = magicints[maxidx] / 2;

  be_putf(out + off, prec); off += 4;
  for (d = 0; d < 3; d++) { be_put32(out + off, (uint32_t)minint[d]); off += 4; }
  for (d = 0; d < 3; d++) { be_put32(out + off, (uint32_t)maxint[d]); off += 4; }
  be_put32(out + off, (uint32_t)smallidx); off += 4;
  const size_t len_at = off;
  off += 4;
  bitbuf b;
  memset(&b, 0, sizeof b);
  b.p = out + off; b.cap = cap - off;

  int prevrun = -1, *prevcoord = NULL;
  i = 0;
  while (i < natoms) {
    int is_small = 0, is_smaller, run = 0, k;
    int *thiscoord = ip + 3 * i;
    unsigned tmp[3], small[8 * 3];
    if (smallidx < maxidx && i >= 1 && iabs(thiscoord[0] - prevcoord[0]) < larger &&
        iabs(thiscoord[1] - prevcoord[1]) < larger && iabs(thiscoord[2] - prevcoord[2]) < larger)
      is_smaller = 1;
    else if (smallidx > minidx)
      is_smaller = -1;
    else
      is_smaller = 0;
    if (i + 1 < natoms && iabs(thiscoord[0] - thiscoord[3]) < smallnum && iabs(thiscoord[1] - thiscoord[4]) < smallnum &&
        iabs(thiscoord[2] - thiscoord[5]) < smallnum) {
      /* the second atom of the pair goes out in full, the first as a small difference to it */
      for (d = 0; d < 3; d++) { const int t = thiscoord[d]; thiscoord[d] = thiscoord[d + 3]; thiscoord[d + 3] = t; }
      is_small = 1;
    }
    for (d = 0; d < 3; d++) tmp[d] = (unsigned)(thiscoord[d] - minint[d]);
    if (bitsize == 0) { for (d = 0; d < 3; d++) put_bits(&b, (int)bitsizeint[d], tmp[d]); }
    else put_triple(&b, bitsize, sizeint, tmp);
    prevcoord = thiscoord;
    thiscoord += 3;
    i++;
    if (!is_small && is_smaller == -1) is_smaller = 0;
    while (is_small && run < 8 * 3) {
      if (is_smaller == -1) {
        const long long dx = thiscoord[0] - prevcoord[0], dy = thiscoord[1] - prevcoord[1], dz = thiscoord[2] - prevcoord[2];
        if (dx * dx + dy * dy + dz * dz >= (long long)smaller * smaller) is_smaller = 0;
      }
      for (d = 0; d < 3; d++) small[run++] = (unsigned)(thiscoord[d] - prevcoord[d] + smallnum);
      prevcoord = thiscoord;
      i++;
      thiscoord += 3;
      is_small = 0;
      if (i < natoms && iabs(thiscoord[0] - prevcoord[0]) < smallnum && iabs(thiscoord[1] - prevcoord[1]) < smallnum &&
          iabs(thiscoord[2] - prevcoord[2]) < smallnum)
        is_small = 1;
    }
    if (run != prevrun || is_smaller != 0) {
      prevrun = run;
      put_bits(&b, 1, 1);
      put_bits(&b, 5, (unsigned)(run + is_smaller + 1));
    } else {
      put_bits(&b, 1, 0);
    }
    for (k = 0; k < run; k += 3) put_triple(&b, smallidx, sizesmall, small + k);
    if (is_smaller != 0) {
      smallidx += is_smaller;
      if (is_smaller < 0) { smallnum = smaller; smaller = magicints[smallidx - 1] / 2; }
      else { smaller = smallnum; smallnum = magicints[smallidx] / 2; }
      sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
    }
    if (b.err) { free(ip); return 0; }
  }
  free(ip);
  const size_t nbytes = finish_bits(&b);
  if (b.err) return 0;
  be_put32(out + len_at, (uint32_t)nbytes);
  off += nbytes;
  while (off & 3) { if (off >= cap) return 0; out[off++] = 0; }
  return off;
}

/* ------------------------------------------------------------------ decoder */
size_t gh_xtc_decode_frame(const uint8_t *in, size_t len, gh_frame_info *info, float *x, int max_atoms)
{
  size_t off = 0;
  int i, d;
  if (len < (size_t)XTC_HEADER_BYTES) return 0;
  if (be_get32(in) != XTC_MAGIC) return 0;
  const int natoms = (int)be_get32(in + 4);
  info->natoms = natoms;
  info->step = (int)be_get32(in + 8);
  info->time = be_getf(in + 12);
  for (i = 0; i < 9; i++) info->box[i] = be_getf(in + 16 + 4 * i);
  off = 52;
  const int lsize = (int)be_get32(in + off); off += 4;
  info->prec = 0;
  if (natoms < 0 || lsize != natoms || (x && natoms > max_atoms)) return 0;
  if (natoms <= 9) {
    if (len < off + (size_t)natoms * 12) return 0;
    for (i = 0; i < natoms * 3; i++) { if (x) x[i] = be_getf(in + off); off += 4; }
    return off;
  }
  if (len < off + 36) return 0;
  const float prec = be_getf(in + off); off += 4;
  int minint[3], maxint[3];
  for (d = 0; d < 3; d++) { minint[d] = (int)be_get32(in + off); off += 4; }
  for (d = 0; d < 3; d++) { maxint[d] = (int)be_get32(in + off); off += 4; }
  int smallidx = (int)be_get32(in + off); off += 4;
  const size_t nbytes = be_get32(in + off); off += 4;
  const size_t padded = (nbytes + 3) & ~(size_t)3;
  info->prec = prec;
  if (len < off + padded || smallidx < FIRSTIDX || smallidx > LASTIDX) return 0;
  if (!x) return off + padded;                       /* header only: skip the frame */
  unsigned sizeint[3], bitsizeint[3] = {0, 0, 0};
  int bitsize;
  for (d = 0; d < 3; d++) {
    if (maxint[d] < minint[d] || !(prec > 0)) return 0;     /* damaged header: sizeint would wrap to 0 (division by zero below) */
    sizeint[d] = (unsigned)(maxint[d] - minint[d]) + 1u;
    if (sizeint[d] == 0u) return 0;
  }
  if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) {
    for (d = 0; d < 3; d++) bitsizeint[d] = (unsigned)bits_for(sizeint[d]);
    bitsize = 0;
  } else {
    bitsize = bits_for_triple(sizeint);
  }
  int smaller = magicints[smallidx - 1 > FIRSTIDX ? smallidx - 1 : FIRSTIDX] / 2;
  int smallnum = magicints[smallidx] / 2;
  unsigned sizesmall[3];
  sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
  const float inv = 1.0f / prec;
  bitbuf b;
  memset(&b, 0, sizeof b);
  b.p = (uint8_t *)(uintptr_t)(in + off); b.cap = nbytes;
  int run = 0;
  float *o = x;
  i = 0;
  while (i < natoms) {
    int cur[3], prev[3], k, is_smaller = 0;
    if (bitsize == 0) { for (d = 0; d < 3; d++) cur[d] = (int)get_bits(&b, (int)bitsizeint[d]); }
    else get_triple(&b, bitsize, sizeint, cur);
    i++;
    for (d = 0; d < 3; d++) { cur[d] += minint[d]; prev[d] = cur[d]; }
    if (get_bits(&b, 1) == 1) {
      run = (int)get_bits(&b, 5);
      is_smaller = run % 3;
      run -= is_smaller;
      is_smaller--;
    }
    if (b.err) return 0;
    if (run > 0) {
      if (i + run / 3 > natoms) return 0;
      for (k = 0; k < run; k += 3) {
        get_triple(&b, smallidx, sizesmall, cur);
        i++;
        for (d = 0; d < 3; d++) cur[d] += prev[d] - smallnum;
        if (k == 0) {
          /* undo the swap of the first two atoms of the run */
          for (d = 0; d < 3; d++) { const int t = cur[d]; cur[d] = prev[d]; prev[d] = t; }
          for (d = 0; d < 3; d++) *o++ = (float)prev[d] * inv;
        } else {
          for (d = 0; d < 3; d++) prev[d] = cur[d];
        }
        for (d = 0; d < 3; d++) *o++ = (float)cur[d] * inv;
      }
    } else {
      for (d = 0; d < 3; d++) *o++ = (float)cur[d] * inv;
    }
    smallidx += is_smaller;
    if (smallidx < FIRSTIDX || smallidx > LASTIDX) return 0;
    if (is_smaller < 0) { smallnum = smaller; smaller = smallidx > FIRSTIDX ? magicints[smallidx - 1] / 2 : 0; }
    else if (is_smaller > 0) { smaller = smallnum; smallnum = magicints[smallidx] / 2; }
    sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
    if (b.err) return 0;
  }
  return off + padded;
}

/* ------------------------------------------------------------------ files */
struct gh_xtc_writer { FILE *f; uint8_t *buf; size_t cap; };

gh_xtc_writer *gh_xtc_create(const char *path)
{
  gh_xtc_writer *w = (gh_xtc_writer *)calloc(1, sizeof *w);
  if (!w) return NULL;
  w->f = fopen(path, "wb");
  if (!w->f) { free(w); return NULL; }
  return w;
}

int gh_xtc_write(gh_xtc_writer *w, int natoms, int step, float time, const float box9[9], const float *x, float prec)
{
  const size_t need = (size_t)natoms * 16 + 256;
  if (need > w->cap) {
    uint8_t *nb = (uint8_t *)realloc(w->buf, need);
    if (!nb) return -1;
    w->buf = nb; w->cap = need;
  }
  const size_t n = gh_xtc_encode_frame(natoms, step, time, box9, x, prec, w->buf, w->cap);
  if (!n) return -1;
  return fwrite(w->buf, 1, n, w->f) == n ? 0 : -1;
}

void gh_xtc_close(gh_xtc_writer *w)
{
  if (!w) return;
  if (w->f) fclose(w->f);
  free(w->buf);
  free(w);
}

enum { TRJ_XTC, TRJ_RAW, TRJ_GRO };
struct gh_traj {
  FILE *f;
  int kind, natoms;
  long nframes, iframe;        /* raw */
  uint8_t *buf; size_t cap;    /* xtc frame buffer */
  int have_first;              /* selection state for -dt */
  double t0;
};

static int has_ext(const char *path, const char *ext)
{
  const size_t n = strlen(path), m = strlen(ext);
  return n >= m && !strcmp(path + n - m, ext);
}

static int xtc_peek_natoms(FILE *f)
{
  uint8_t h[12];
  if (fread(h, 1, 12, f) != 12 || be_get32(h) != XTC_MAGIC) return -1;
  rewind(f);
  return (int)be_get32(h + 4);
}

gh_traj *gh_traj_open(const char *path)
{
  gh_traj *t = (gh_traj *)calloc(1, sizeof *t);
  if (!t) return NULL;
  t->f = fopen(path, "rb");
  if (!t->f) { free(t); return NULL; }
  if (has_ext(path, ".xtc")) {
    t->kind = TRJ_XTC;
    t->natoms = xtc_peek_natoms(t->f);
    if (t->natoms < 0) { gh_traj_close(t); return NULL; }
  } else if (has_ext(path, ".gro")) {
    char line[512];
    t->kind = TRJ_GRO;
    if (!fgets(line, sizeof line, t->f) || !fgets(line, sizeof line, t->f)) { gh_traj_close(t); return NULL; }
    t->natoms = atoi(line);
    rewind(t->f);
  } else {
    char magic[8];
    int32_t hdr[2];
    t->kind = TRJ_RAW;
    if (fread(magic, 1, 8, t->f) != 8 || memcmp(magic, "GCTRJ1\n", 7) || fread(hdr, 4, 2, t->f) != 2) { gh_traj_close(t); return NULL; }
    t->natoms = hdr[0];
    t->nframes = hdr[1];
  }
  return t;
}

int gh_traj_natoms(gh_traj *t) { return t->natoms; }

static int next_xtc(gh_traj *t, gh_frame_info *info, float *x)
{
  uint8_t head[XTC_HEADER_BYTES + 36];
  size_t got = fread(head, 1, XTC_HEADER_BYTES, t->f);
  if (got == 0) return 0;
  if (got != XTC_HEADER_BYTES || be_get32(head) != XTC_MAGIC) return -1;
  const int natoms = (int)be_get32(head + 4);
  size_t total;
  if (natoms != t->natoms) return -1;
  if (natoms <= 9) {
    total = XTC_HEADER_BYTES + (size_t)natoms * 12;
  } else {
    if (fread(head + XTC_HEADER_BYTES, 1, 36, t->f) != 36) return -1;
    const size_t nbytes = be_get32(head + XTC_HEADER_BYTES + 32);
    total = XTC_HEADER_BYTES + 36 + ((nbytes + 3) & ~(size_t)3);
  }
  if (total > t->cap) {                     /* (a damaged header may ask for gigabytes: that is a bad file, not a crash) */
    uint8_t *nb = (uint8_t *)realloc(t->buf, total);
    if (!nb) return -1;
    t->buf = nb; t->cap = total;
  }
  const size_t have = natoms <= 9 ? XTC_HEADER_BYTES : XTC_HEADER_BYTES + 36;
  memcpy(t->buf, head, have);
  if (fread(t->buf + have, 1, total - have, t->f) != total - have) return -1;
  return gh_xtc_decode_frame(t->buf, total, info, x, t->natoms) == total ? 1 : -1;
}

static int next_raw(gh_traj *t, gh_frame_info *info, float *x)
{
  if (t->iframe >= t->nframes) return 0;
  if (fread(&info->time, 4, 1, t->f) != 1 || fread(info->box, 4, 9, t->f) != 9) return -1;
  if (fread(x, 12, (size_t)t->natoms, t->f) != (size_t)t->natoms) return -1;
  info->natoms = t->natoms;
  info->step = (int)t->iframe++;
  info->prec = 0;
  return 1;
}

static int next_gro(gh_traj *t, gh_frame_info *info, float *x)
{
  char line[512];
  int i;
  if (!fgets(line, sizeof line, t->f)) return 0;
  const char *tp = strstr(line, "t=");
  info->time = tp ? (float)atof(tp + 2) : (float)t->iframe;
  if (!fgets(line, sizeof line, t->f) || atoi(line) != t->natoms) return -1;
  for (i = 0; i < t->natoms; i++) {
    if (!fgets(line, sizeof line, t->f) || strlen(line) < 44) return -1;
    /* fixed columns: resnr(5) resname(5) atomname(5) atomnr(5) x(8) y(8) z(8) */
    char fld[9];
    int d;
    for (d = 0; d < 3; d++) { memcpy(fld, line + 20 + 8 * d, 8); fld[8] = 0; x[3 * i + d] = (float)atof(fld); }
  }
  if (!fgets(line, sizeof line, t->f)) return -1;
  memset(info->box, 0, sizeof info->box);
  {
    float v[9] = {0};
    const int n = sscanf(line, "%f %f %f %f %f %f %f %f %f", v, v + 1, v + 2, v + 3, v + 4, v + 5, v + 6, v + 7, v + 8);
    if (n < 3) return -1;
    info->box[0] = v[0]; info->box[4] = v[1]; info->box[8] = v[2];
    if (n == 9) { info->box[1] = v[3]; info->box[2] = v[4]; info->box[3] = v[5]; info->box[5] = v[6]; info->box[6] = v[7]; info->box[7] = v[8]; }
  }
  info->natoms = t->natoms;
  info->step = (int)t->iframe++;
  info->prec = 0;
  return 1;
}

int gh_traj_next(gh_traj *t, gh_frame_info *info, float *x)
{
  switch (t->kind) {
    case TRJ_XTC: return next_xtc(t, info, x);
    case TRJ_GRO: return next_gro(t, info, x);
    default: return next_raw(t, info, x);
  }
}

int gh_traj_next_selected(gh_traj *t, const gh_timesel *ts, gh_frame_info *info, float *x)
{
  for (;;) {
    const int rc = gh_traj_next(t, info, x);
    if (rc <= 0) return rc;
    if (!ts) return 1;
    if (ts->have_b && (double)info->time < ts->b) continue;
    if (ts->have_e && (double)info->time > ts->e) return 0;
    if (ts->have_dt && ts->dt > 0) {
      /* frames whose time is a multiple of dt counted from the first selected frame */
      if (!t->have_first) { t->have_first = 1; t->t0 = info->time; }
      const double r = fmod((double)info->time - t->t0, ts->dt);
      if (fabs(r) > 1e-4 * ts->dt && fabs(r - ts->dt) > 1e-4 * ts->dt) continue;
    }
    return 1;
  }
}

void gh_traj_close(gh_traj *t)
{
  if (!t) return;
  if (t->f) fclose(t->f);
  free(t->buf);
  free(t);
}

int gh_top_natoms(const char *path, int *natoms, float box9[9])
{
  FILE *f = fopen(path, "r");
  char line[512];
  int n = -1;
  if (!f) return -1;
  if (box9) memset(box9, 0, sizeof(float) * 9);
  if (!fgets(line, sizeof line, f)) { fclose(f); return -1; }
  if (!strncmp(line, "GCTOP1", 6)) {
    n = atoi(line + 6);
  } else if (fgets(line, sizeof line, f)) {       /* .gro: title, natoms, atoms, box */
    int i;
    n = atoi(line);
    for (i = 0; i < n; i++) if (!fgets(line, sizeof line, f)) { n = -1; break; }
    if (n >= 0 && box9 && fgets(line, sizeof line, f)) {
      float v[3] = {0, 0, 0};
      if (sscanf(line, "%f %f %f", v, v + 1, v + 2) == 3) { box9[0] = v[0]; box9[4] = v[1]; box9[8] = v[2]; }
    }
  }
  fclose(f);
  if (n <= 0) return -1;
  *natoms = n;
  return 0;
}
