/*
 * fakegmx.c -- TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * A small stand-in runtime for the Gromacs 4.5 / SunRPC functions the
 * UNMODIFIED reference sources call, so that /root/reference/src/*.c can be
 * compiled and executed in this image as the parity oracle.
 * It supplies: fatal/log plumbing, a table-driven option parser with the
 * Gromacs conventions the three tools rely on, a raw float32 trajectory
 * reader in place of the XTC/TRR readers, a text topology stub, an .ndx
 * reader, XDR-on-stdio (RFC 4506), and an optional full-precision capture of
 * every floating value the reference prints (see capture.h).
 *
 * Raw trajectory ("GCTRJ1"): 8-byte magic, int32 natoms, int32 nframes, then
 * per frame float32 t, float32 box[9] (row major), float32 x[natoms][3];
 * all little-endian.  Topology stub: text "GCTOP1 <natoms>".
 */
#include "gmxshim.h"
#include "statutil.h"
#include "rpc/xdr.h"
#include <stdarg.h>
#include <stdint.h>
#include <ctype.h>

/* ------------------------------------------------------------------ */
/* logging / fatal                                                     */
/* ------------------------------------------------------------------ */
FILE *debug = NULL;
static char g_program[256] = "gcshim";
static int g_debug_mode = 0;

gmx_bool bDebugMode(void) { return g_debug_mode; }
const char *Program(void) { return g_program; }

void gmx_fatal(int fatal_errno, const char *file, int line, const char *fmt, ...)
{
  va_list ap;
  (void)fatal_errno;
  fprintf(stderr, "\nFatal error (%s, line %d):\n", file, line);
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fprintf(stderr, "\n");
  exit(1);
}

void CopyRight(FILE *out, const char *szProgram)
{
  const char *p = strrchr(szProgram, '/');
  (void)out;
  snprintf(g_program, sizeof(g_program), "%s", p ? p + 1 : szProgram);
}
void thanx(FILE *fp) { (void)fp; }

/* ------------------------------------------------------------------ */
/* file registry (so that captures know the name behind a FILE*)       */
/* ------------------------------------------------------------------ */
#define MAXREG 64
static struct { FILE *fp; char name[1024]; FILE *cap; } g_reg[MAXREG];

static void reg_add(FILE *fp, const char *name)
{
  int i;
  for (i = 0; i < MAXREG; i++)
    if (!g_reg[i].fp) {
      g_reg[i].fp = fp;
      snprintf(g_reg[i].name, sizeof(g_reg[i].name), "%s", name);
      g_reg[i].cap = NULL;
      return;
    }
}
static int reg_find(FILE *fp)
{
  int i;
  for (i = 0; i < MAXREG; i++)
    if (g_reg[i].fp == fp) return i;
  return -1;
}

FILE *ffopen(const char *fn, const char *mode)
{
  FILE *fp = fopen(fn, mode);
  if (!fp) gmx_fatal(FARGS, "Cannot open file %s (mode %s)", fn, mode);
  reg_add(fp, fn);
  return fp;
}
static int reg_close(FILE *fp)
{
  int i = reg_find(fp);
  if (i >= 0) {
    if (g_reg[i].cap) fclose(g_reg[i].cap);
    g_reg[i].fp = NULL;
    g_reg[i].cap = NULL;
  }
  return fclose(fp);
}
void ffclose(FILE *fp) { reg_close(fp); }

/* ---- capture entry points used through capture.h -------------------- */
static FILE *g_stdout_cap = NULL;
static int capture_on(void)
{
  static int on = -1;
  if (on < 0) { const char *e = getenv("GCSHIM_CAPTURE"); on = (e && *e && *e != '0'); }
  return on;
}
static FILE *cap_stream(FILE *fp)
{
  int i;
  char nm[1100];
  if (!capture_on()) return NULL;
  if (fp == stdout) {
    if (!g_stdout_cap) g_stdout_cap = fopen("stdout.f64", "wb");
    return g_stdout_cap;
  }
  i = reg_find(fp);
  if (i < 0) return NULL;
  if (!g_reg[i].cap) {
    snprintf(nm, sizeof(nm), "%s.f64", g_reg[i].name);
    g_reg[i].cap = fopen(nm, "wb");
  }
  return g_reg[i].cap;
}

/* Walk a printf format, pulling the arguments exactly as printf would, and
 * append every floating argument (as the double that was passed) to `cap`. */
static void cap_walk(FILE *cap, const char *fmt, va_list ap)
{
  const char *p = fmt;
  while (*p) {
    int lmod = 0;
    if (*p++ != '%') continue;
    if (*p == '%') { p++; continue; }
    while (*p && strchr("-+ #0", *p)) p++;
    if (*p == '*') { (void)va_arg(ap, int); p++; } else while (isdigit((unsigned char)*p)) p++;
    if (*p == '.') {
      p++;
      if (*p == '*') { (void)va_arg(ap, int); p++; } else while (isdigit((unsigned char)*p)) p++;
    }
    while (*p == 'l' || *p == 'h' || *p == 'z' || *p == 'L') { if (*p == 'l') lmod++; if (*p == 'z') lmod = 1; p++; }
    switch (*p) {
      case 'd': case 'i': case 'u': case 'x': case 'X': case 'o': case 'c':
        if (lmod >= 2) (void)va_arg(ap, long long);
        else if (lmod == 1) (void)va_arg(ap, long);
        else (void)va_arg(ap, int);
        break;
      case 'f': case 'F': case 'e': case 'E': case 'g': case 'G': case 'a': case 'A': {
        double d = va_arg(ap, double);
        if (cap) fwrite(&d, sizeof(d), 1, cap);
        break;
      }
      case 's': (void)va_arg(ap, char *); break;
      case 'p': (void)va_arg(ap, void *); break;
      default: break;
    }
    if (*p) p++;
  }
}

int gcshim_fprintf(FILE *fp, const char *fmt, ...)
{
  va_list ap, aq;
  int n;
  FILE *cap = cap_stream(fp);
  va_start(ap, fmt);
  if (cap) { va_copy(aq, ap); cap_walk(cap, fmt, aq); va_end(aq); }
  n = vfprintf(fp, fmt, ap);
  va_end(ap);
  return n;
}
int gcshim_printf(const char *fmt, ...)
{
  va_list ap, aq;
  int n;
  FILE *cap = cap_stream(stdout);
  va_start(ap, fmt);
  if (cap) { va_copy(aq, ap); cap_walk(cap, fmt, aq); va_end(aq); }
  n = vfprintf(stdout, fmt, ap);
  va_end(ap);
  if (cap) fflush(cap);
  return n;
}
FILE *gcshim_fopen(const char *fn, const char *mode)
{
  FILE *fp = fopen(fn, mode);
  if (fp) reg_add(fp, fn);
  return fp;
}
int gcshim_fclose(FILE *fp) { return reg_close(fp); }

/* ------------------------------------------------------------------ */
/* option parsing                                                      */
/* ------------------------------------------------------------------ */
static int g_nfile = 0;
static double g_tbegin = -1e300, g_tend = 1e300;

static const char *ftp_default_opt(int ftp)
{
  switch (ftp) {
    case efTRX: return "-f";
    case efTPS: return "-s";
    case efNDX: return "-n";
    case efXVG: return "-o";
    default: return "-f";
  }
}
static const char *ftp_default_name(int ftp)
{
  switch (ftp) {
    case efTRX: return "traj";
    case efTPS: return "topol";
    case efNDX: return "index";
    case efXVG: return "out";
    default: return "data";
  }
}
static const char *ftp_ext(int ftp)
{
  switch (ftp) {
    case efTRX: return ".xtc";
    case efTPS: return ".tpr";
    case efNDX: return ".ndx";
    case efXVG: return ".xvg";
    case efDAT: return ".dat";
    default: return "";
  }
}
/* Gromacs appends the type's default extension when the name has none. */
static char *with_ext(const char *name, int ftp)
{
  const char *slash = strrchr(name, '/');
  const char *dot = strrchr(slash ? slash : name, '.');
  size_t n = strlen(name) + 8;
  char *out = malloc(n);
  if (dot) snprintf(out, n, "%s", name);
  else snprintf(out, n, "%s%s", name, ftp_ext(ftp));
  return out;
}

static int is_option(const char *s)
{
  /* "-x" where x is not a digit or '.' (negative numbers are values) */
  return s[0] == '-' && s[1] && !isdigit((unsigned char)s[1]) && s[1] != '.';
}

void parse_common_args(int *argc, char *argv[], unsigned long Flags,
                       int nfile, t_filenm fnm[], int npargs, t_pargs *pa,
                       int ndesc, const char **desc, int nbugs, const char **bugs,
                       output_env_t *oenv)
{
  int i, k, a;
  (void)ndesc; (void)desc; (void)nbugs; (void)bugs;
  if (oenv) *oenv = NULL;
  g_nfile = nfile;

  for (i = 0; i < nfile; i++) {
    if (!fnm[i].opt) fnm[i].opt = ftp_default_opt(fnm[i].ftp);
    fnm[i].nfiles = 0;
    fnm[i].fns = NULL;
    fnm[i].fn = with_ext(fnm[i].fn ? fnm[i].fn : ftp_default_name(fnm[i].ftp), fnm[i].ftp);
  }
  for (k = 0; k < npargs; k++) {
    pa[k].bSet = FALSE;
    if (pa[k].type == etENUM) pa[k].u.c[0] = NULL;
  }

  for (a = 1; a < *argc; a++) {
    const char *arg = argv[a];
    int done = 0;
    if (!strcmp(arg, "-debug")) {
      g_debug_mode = 1;
      if (a + 1 < *argc && !is_option(argv[a + 1])) a++;
      if (!debug) debug = fopen("gcshim.debug", "w");
      continue;
    }
    if (!strcmp(arg, "-nice") || !strcmp(arg, "-tu") || !strcmp(arg, "-xvg")) { a++; continue; }
    if (!strcmp(arg, "-h") || !strcmp(arg, "-quiet") || !strcmp(arg, "-w") || !strcmp(arg, "-now")) continue;
    if ((Flags & PCA_CAN_BEGIN) && !strcmp(arg, "-b") && a + 1 < *argc) { g_tbegin = atof(argv[++a]); continue; }
    if ((Flags & PCA_CAN_END) && !strcmp(arg, "-e") && a + 1 < *argc) { g_tend = atof(argv[++a]); continue; }
    if ((Flags & PCA_CAN_DT) && !strcmp(arg, "-dt") && a + 1 < *argc) { a++; continue; }

    for (i = 0; i < nfile && !done; i++) {
      if (strcmp(arg, fnm[i].opt)) continue;
      fnm[i].flag |= ffSET;
      if (fnm[i].flag & ffMULT) {
        while (a + 1 < *argc && !is_option(argv[a + 1])) {
          fnm[i].fns = realloc(fnm[i].fns, sizeof(char *) * (size_t)(fnm[i].nfiles + 1));
          fnm[i].fns[fnm[i].nfiles++] = with_ext(argv[++a], fnm[i].ftp);
        }
        if (fnm[i].nfiles) fnm[i].fn = fnm[i].fns[0];
      } else if (a + 1 < *argc && !is_option(argv[a + 1])) {
        fnm[i].fn = with_ext(argv[++a], fnm[i].ftp);
      }
      done = 1;
    }
    for (k = 0; k < npargs && !done; k++) {
      const char *o = pa[k].option;
      if (pa[k].type == etBOOL) {
        if (!strcmp(arg, o)) { *pa[k].u.b = TRUE; pa[k].bSet = TRUE; done = 1; }
        else if (!strncmp(arg, "-no", 3) && !strcmp(arg + 3, o + 1)) { *pa[k].u.b = FALSE; pa[k].bSet = TRUE; done = 1; }
        continue;
      }
      if (strcmp(arg, o)) continue;
      if (a + 1 >= *argc) gmx_fatal(FARGS, "Option %s needs a value", o);
      pa[k].bSet = TRUE;
      done = 1;
      switch (pa[k].type) {
        case etINT: *pa[k].u.i = atoi(argv[++a]); break;
        case etREAL: case etTIME: *pa[k].u.r = (real)strtod(argv[++a], NULL); break;
        case etSTR: *pa[k].u.c = argv[++a]; break;
        case etRVEC: {
          int d, got = 0;
          real v[DIM] = {0, 0, 0};
          for (d = 0; d < DIM && a + 1 < *argc && !is_option(argv[a + 1]); d++) { v[d] = (real)strtod(argv[++a], NULL); got++; }
          if (got == 1) v[1] = v[2] = v[0];       /* Gromacs: one number fills all three */
          else if (got != 3) gmx_fatal(FARGS, "Option %s needs one or three numbers", o);
          for (d = 0; d < DIM; d++) (*pa[k].u.rv)[d] = v[d];
          break;
        }
        case etENUM: {
          int e, hit = 0;
          const char *val = argv[++a];
          for (e = 1; pa[k].u.c[e]; e++)
            if (!strncasecmp(val, pa[k].u.c[e], strlen(val))) { pa[k].u.c[0] = pa[k].u.c[e]; hit = 1; break; }
          if (!hit) gmx_fatal(FARGS, "Invalid value %s for option %s", val, o);
          break;
        }
        default: break;
      }
    }
    if (!done) gmx_fatal(FARGS, "Invalid command line argument:\n%s", arg);
  }
  for (k = 0; k < npargs; k++)
    if (pa[k].type == etENUM && !pa[k].u.c[0]) pa[k].u.c[0] = pa[k].u.c[1];
}

gmx_bool opt2parg_bSet(const char *opt, int nparg, t_pargs pa[])
{
  int k;
  for (k = 0; k < nparg; k++)
    if (!strcmp(opt, pa[k].option)) return pa[k].bSet;
  gmx_fatal(FARGS, "No such option %s in pargs", opt);
  return FALSE;
}
const char *ftp2fn(int ftp, int nfile, t_filenm fnm[])
{
  int i;
  for (i = 0; i < nfile; i++)
    if (fnm[i].ftp == ftp) return fnm[i].fn;
  gmx_fatal(FARGS, "ftp2fn: no file of type %d", ftp);
  return NULL;
}
const char *ftp2fn_null(int ftp, int nfile, t_filenm fnm[])
{
  int i;
  for (i = 0; i < nfile; i++)
    if (fnm[i].ftp == ftp) {
      if ((fnm[i].flag & ffOPT) && !(fnm[i].flag & ffSET)) return NULL;
      return fnm[i].fn;
    }
  return NULL;
}
const char *opt2fn(const char *opt, int nfile, t_filenm fnm[])
{
  int i;
  for (i = 0; i < nfile; i++)
    if (!strcmp(opt, fnm[i].opt)) return fnm[i].fn;
  gmx_fatal(FARGS, "opt2fn: no option %s", opt);
  return NULL;
}
int opt2fns(char **fns[], const char *opt, int nfile, t_filenm fnm[])
{
  int i;
  for (i = 0; i < nfile; i++)
    if (!strcmp(opt, fnm[i].opt)) { *fns = fnm[i].fns; return fnm[i].nfiles; }
  return 0;
}
gmx_bool opt2bSet(const char *opt, int nfile, t_filenm fnm[])
{
  int i;
  for (i = 0; i < nfile; i++)
    if (!strcmp(opt, fnm[i].opt)) return (fnm[i].flag & ffSET) != 0;
  return FALSE;
}

/* ------------------------------------------------------------------ */
/* topology / index / trajectory                                       */
/* ------------------------------------------------------------------ */
gmx_bool read_tps_conf(const char *infile, char *title, t_topology *top, int *ePBC,
                       rvec **x, rvec **v, matrix box, gmx_bool bMass)
{
  FILE *fp = fopen(infile, "r");
  char tag[32];
  int natoms = 0, i, j;
  (void)bMass;
  if (!fp) gmx_fatal(FARGS, "Cannot open topology stub %s", infile);
  if (fscanf(fp, "%31s %d", tag, &natoms) != 2 || strcmp(tag, "GCTOP1"))
    gmx_fatal(FARGS, "%s is not a GCTOP1 topology stub", infile);
  fclose(fp);
  if (title) strcpy(title, "gcshim");
  memset(top, 0, sizeof(*top));
  top->atoms.nr = natoms;
  snew(top->atoms.atom, natoms);
  for (i = 0; i < natoms; i++) top->atoms.atom[i].m = 1;
  if (ePBC) *ePBC = 0;
  if (x) snew(*x, natoms);
  if (v) *v = NULL;
  for (i = 0; i < DIM; i++) for (j = 0; j < DIM; j++) box[i][j] = 0;
  return TRUE;
}

void get_index(t_atoms *atoms, const char *fnm, int ngrps, int isize[],
               atom_id *index[], char *grpnames[])
{
  int i;
  if (ngrps != 1) gmx_fatal(FARGS, "gcshim get_index: one group only");
  if (!fnm) {
    isize[0] = atoms->nr;
    snew(index[0], atoms->nr);
    for (i = 0; i < atoms->nr; i++) index[0][i] = i;
    grpnames[0] = strdup("System");
    return;
  }
  {
    /* .ndx: "[ name ]" headers followed by 1-based atom numbers */
    FILE *fp = fopen(fnm, "r");
    const char *sel = getenv("GCSHIM_GROUP");
    int want = sel ? atoi(sel) : 0, cur = -1, n = 0, cap = 0;
    atom_id *list = NULL;
    char name[256] = "", line[8192];
    if (!fp) gmx_fatal(FARGS, "Cannot open index file %s", fnm);
    while (fgets(line, sizeof(line), fp)) {
      char *p = line;
      while (isspace((unsigned char)*p)) p++;
      if (*p == '[') {
        cur++;
        if (cur == want) sscanf(p + 1, " %255[^] \t\n]", name);
        continue;
      }
      if (cur != want) continue;
      while (*p) {
        char *end;
        long v = strtol(p, &end, 10);
        if (end == p) break;
        if (n == cap) { cap = cap ? 2 * cap : 1024; list = realloc(list, sizeof(atom_id) * (size_t)cap); }
        list[n++] = (atom_id)(v - 1);
        p = end;
      }
    }
    fclose(fp);
    if (cur < want) gmx_fatal(FARGS, "Index file %s has no group %d", fnm, want);
    isize[0] = n;
    index[0] = list;
    grpnames[0] = strdup(name);
  }
}

struct gcshim_trx { FILE *fp; int natoms, nframes, iframe; };

static t_trxstatus *trx_open(const char *fn)
{
  char magic[8];
  t_trxstatus *st = calloc(1, sizeof(*st));
  st->fp = fopen(fn, "rb");
  if (!st->fp) gmx_fatal(FARGS, "Cannot open trajectory %s", fn);
  if (fread(magic, 1, 8, st->fp) != 8 || memcmp(magic, "GCTRJ1", 6))
    gmx_fatal(FARGS, "%s is not a GCTRJ1 raw trajectory", fn);
  if (fread(&st->natoms, 4, 1, st->fp) != 1 || fread(&st->nframes, 4, 1, st->fp) != 1)
    gmx_fatal(FARGS, "%s: short header", fn);
  return st;
}
static gmx_bool trx_read(t_trxstatus *st, float *t, rvec x[], matrix box)
{
  float b9[9];
  int i, j;
  for (;;) {
    if (st->nframes >= 0 && st->iframe >= st->nframes) return FALSE;
    if (fread(t, 4, 1, st->fp) != 1) return FALSE;
    if (fread(b9, 4, 9, st->fp) != 9) return FALSE;
    if (fread(x, sizeof(float) * 3, (size_t)st->natoms, st->fp) != (size_t)st->natoms) return FALSE;
    st->iframe++;
    if ((double)*t < g_tbegin) continue;   /* -b: skip early frames */
    if ((double)*t > g_tend) return FALSE; /* -e: stop */
    break;
  }
  for (i = 0; i < DIM; i++) for (j = 0; j < DIM; j++) box[i][j] = b9[3 * i + j];
  return TRUE;
}

#ifdef DOUBLE
#error "the raw trajectory reader stores float32 frames; build the oracle with real=float"
#endif

int read_first_x(const output_env_t oenv, t_trxstatus **status, const char *fn,
                 real *t, rvec **x, matrix box)
{
  t_trxstatus *st = trx_open(fn);
  (void)oenv;
  snew(*x, st->natoms);
  if (!trx_read(st, t, *x, box)) gmx_fatal(FARGS, "No frames in %s", fn);
  *status = st;
  return st->natoms;
}
gmx_bool read_next_x(const output_env_t oenv, t_trxstatus *status, real *t,
                     int natoms, rvec x[], matrix box)
{
  (void)oenv; (void)natoms;
  return trx_read(status, t, x, box);
}
gmx_bool read_first_frame(const output_env_t oenv, t_trxstatus **status,
                          const char *fn, t_trxframe *fr, int flags)
{
  t_trxstatus *st = trx_open(fn);
  (void)oenv;
  memset(fr, 0, sizeof(*fr));
  fr->flags = flags;
  fr->natoms = st->natoms;
  snew(fr->x, st->natoms);
  *status = st;
  return read_next_frame(oenv, st, fr);
}
gmx_bool read_next_frame(const output_env_t oenv, t_trxstatus *status, t_trxframe *fr)
{
  float t;
  (void)oenv;
  if (!trx_read(status, &t, fr->x, fr->box)) return FALSE;
  fr->time = t; fr->bTime = TRUE; fr->bX = TRUE; fr->bBox = TRUE;
  return TRUE;
}
void close_trj(t_trxstatus *status)
{
  if (status) { fclose(status->fp); free(status); }
}

real calc_xcm(rvec x[], int gnx, atom_id *index, t_atom *atom, rvec xcm, gmx_bool bQ)
{
  (void)x; (void)gnx; (void)index; (void)atom; (void)bQ;
  xcm[XX] = xcm[YY] = xcm[ZZ] = 0;
  return 0;
}
void read_tpx(const char *fn, t_inputrec *ir, matrix box, int *natoms,
              rvec *x, rvec *v, rvec *f, void *mtop)
{
  (void)fn; (void)box; (void)x; (void)v; (void)f; (void)mtop;
  if (ir) { ir->delta_t = 0; ir->nstxtcout = 0; }
  if (natoms) *natoms = 0;
}

/* ------------------------------------------------------------------ */
/* XDR on stdio (RFC 4506: big-endian 4-byte units)                    */
/* ------------------------------------------------------------------ */
void xdrstdio_create(XDR *xdrs, FILE *fp, enum xdr_op op) { xdrs->x_op = op; xdrs->x_fp = fp; }
void xdr_destroy(XDR *xdrs) { (void)xdrs; }

static bool_t xdr_word(XDR *xdrs, uint32_t *w)
{
  unsigned char b[4];
  if (xdrs->x_op == XDR_ENCODE) {
    b[0] = (unsigned char)(*w >> 24); b[1] = (unsigned char)(*w >> 16);
    b[2] = (unsigned char)(*w >> 8);  b[3] = (unsigned char)(*w);
    return fwrite(b, 1, 4, xdrs->x_fp) == 4;
  }
  if (xdrs->x_op == XDR_DECODE) {
    if (fread(b, 1, 4, xdrs->x_fp) != 4) return 0;
    *w = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3];
    return 1;
  }
  return 1;
}
bool_t xdr_int(XDR *xdrs, int *ip) { uint32_t w = (uint32_t)*ip; if (!xdr_word(xdrs, &w)) return 0; *ip = (int)w; return 1; }
bool_t xdr_u_int(XDR *xdrs, u_int *up) { uint32_t w = *up; if (!xdr_word(xdrs, &w)) return 0; *up = w; return 1; }
bool_t xdr_enum(XDR *xdrs, enum_t *ep) { return xdr_int(xdrs, (int *)ep); }
bool_t xdr_float(XDR *xdrs, float *fp)
{
  uint32_t w;
  memcpy(&w, fp, 4);
  if (!xdr_word(xdrs, &w)) return 0;
  memcpy(fp, &w, 4);
  return 1;
}
bool_t xdr_double(XDR *xdrs, double *dp)
{
  uint64_t q;
  uint32_t hi, lo;
  memcpy(&q, dp, 8);
  hi = (uint32_t)(q >> 32); lo = (uint32_t)q;
  if (!xdr_word(xdrs, &hi) || !xdr_word(xdrs, &lo)) return 0;
  q = ((uint64_t)hi << 32) | lo;
  memcpy(dp, &q, 8);
  return 1;
}
bool_t xdr_string(XDR *xdrs, char **cpp, u_int maxsize)
{
  static const char zero[4] = {0, 0, 0, 0};
  u_int len = 0, pad;
  char junk[4];
  if (xdrs->x_op == XDR_ENCODE) {
    if (!*cpp) return 0;
    len = (u_int)strlen(*cpp);
    if (len > maxsize) return 0;
  }
  if (xdrs->x_op == XDR_FREE) return 1;
  if (!xdr_u_int(xdrs, &len)) return 0;
  if (len > maxsize) return 0;
  pad = (4 - (len & 3)) & 3;
  if (xdrs->x_op == XDR_ENCODE) {
    if (len && fwrite(*cpp, 1, len, xdrs->x_fp) != len) return 0;
    if (pad && fwrite(zero, 1, pad, xdrs->x_fp) != pad) return 0;
    return 1;
  }
  if (!*cpp) *cpp = malloc((size_t)len + 1);
  if (len && fread(*cpp, 1, len, xdrs->x_fp) != len) return 0;
  (*cpp)[len] = '\0';
  if (pad && fread(junk, 1, pad, xdrs->x_fp) != pad) return 0;
  return 1;
}
typedef bool_t (*elemproc_t)(XDR *, void *);
bool_t xdr_vector(XDR *xdrs, char *basep, u_int nelem, u_int elemsize, xdrproc_t proc)
{
  u_int i;
  elemproc_t f = (elemproc_t)proc;
  for (i = 0; i < nelem; i++)
    if (!f(xdrs, basep + (size_t)i * elemsize)) return 0;
  return 1;
}
bool_t xdr_array(XDR *xdrs, char **addrp, u_int *sizep, u_int maxsize, u_int elsize, xdrproc_t proc)
{
  u_int i, c = *sizep;
  elemproc_t f = (elemproc_t)proc;
  if (xdrs->x_op == XDR_FREE) return 1;
  if (!xdr_u_int(xdrs, &c)) return 0;
  if (c > maxsize) return 0;
  *sizep = c;
  if (xdrs->x_op == XDR_DECODE && !*addrp) {
    *addrp = calloc((size_t)(c ? c : 1), elsize);
    if (!*addrp) return 0;
  }
  for (i = 0; i < c; i++)
    if (!f(xdrs, *addrp + (size_t)i * elsize)) return 0;
  return 1;
}
