/*
 * gmxshim.h -- TEST INFRASTRUCTURE ONLY (oracle/): a stand-in for the slice of
 * the Gromacs 4.5.x C API that the unmodified reference sources under
 * /root/reference/src use.  Gromacs itself is not installable here (SURVEY.md
 * section 8c), so the types below are declared from knowledge of the 4.5 API;
 * only their layout as seen by the reference matters.  Nothing in the product
 * path (gridcount_b200/) includes this file.
 *
 * Every Gromacs header name the reference includes (typedefs.h, vec.h, ...) is
 * a one-line file in this directory that includes this one.
 */
#ifndef GC_ORACLE_GMXSHIM_H
#define GC_ORACLE_GMXSHIM_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <sys/types.h>

/* ---- scalar types (Gromacs single precision build) -------------------- */
#ifdef DOUBLE
typedef double real;
#define GMX_REAL_EPS 1.11022302E-16
#else
typedef float real;
/* GMX_FLOAT_EPS of Gromacs 4.5.x (later releases use 1.19e-7). [gmx-memory] */
#define GMX_REAL_EPS 5.96046448E-08
#endif

typedef int gmx_bool;
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#define gmx_inline inline

#define XX 0
#define YY 1
#define ZZ 2
#define DIM 3
#define STRLEN 4096

typedef real rvec[DIM];
typedef int ivec[DIM];
typedef real matrix[DIM][DIM];
typedef int atom_id;

/* ---- topology fragments the reference touches ------------------------- */
typedef struct { int nr; atom_id *index; int nra; atom_id *a; } t_block;
typedef struct { real m, q; } t_atom;
typedef struct { int nr; t_atom *atom; } t_atoms;
typedef struct { char **name; t_atoms atoms; t_block mols; } t_topology;
typedef struct { real delta_t; int nstxtcout; } t_inputrec;

/* ---- command line tables ---------------------------------------------- */
enum { etINT, etREAL, etTIME, etSTR, etBOOL, etRVEC, etENUM, etNR };
typedef struct {
  const char *option;
  gmx_bool bSet;
  int type;
  union { void *v; int *i; real *r; const char **c; gmx_bool *b; rvec *rv; } u;
  const char *desc;
} t_pargs;

enum { efTRX, efTPS, efNDX, efDAT, efXVG, efNR };
#define ffSET   (1 << 0)
#define ffREAD  (1 << 1)
#define ffWRITE (1 << 2)
#define ffOPT   (1 << 3)
#define ffLIB   (1 << 4)
#define ffMULT  (1 << 5)
#define ffOPTRD (ffREAD | ffOPT)
#define ffOPTWR (ffWRITE | ffOPT)
#define ffRDMULT (ffREAD | ffMULT)
typedef struct {
  int ftp;
  const char *opt;
  const char *fn;
  unsigned long flag;
  int nfiles;
  char **fns;
} t_filenm;

typedef struct gcshim_oenv *output_env_t;
typedef struct gcshim_trx t_trxstatus;

#define asize(a) ((int)(sizeof(a) / sizeof((a)[0])))
#define ENUM_NAME(e, max, names) ((((e) < 0) || ((e) >= (max))) ? "UNDEFINED" : (names)[e])

/* ---- smalloc.h --------------------------------------------------------- */
#define snew(ptr, n) (ptr) = calloc((size_t)((n) > 0 ? (n) : 1), sizeof(*(ptr)))
#define srenew(ptr, n) (ptr) = realloc((ptr), (size_t)(n) * sizeof(*(ptr)))
#define sfree(ptr) free(ptr)

/* ---- gmx_fatal.h / futil.h -------------------------------------------- */
#define FARGS 0, __FILE__, __LINE__
void gmx_fatal(int fatal_errno, const char *file, int line, const char *fmt, ...);
extern FILE *debug;
gmx_bool bDebugMode(void);
const char *Program(void);
FILE *ffopen(const char *fn, const char *mode);
void ffclose(FILE *fp);

/* ---- pieces of other Gromacs libraries -------------------------------- */
gmx_bool opt2parg_bSet(const char *opt, int nparg, t_pargs pa[]);
real calc_xcm(rvec x[], int gnx, atom_id *index, t_atom *atom, rvec xcm, gmx_bool bQ);
void read_tpx(const char *fn, t_inputrec *ir, matrix box, int *natoms,
              rvec *x, rvec *v, rvec *f, void *mtop);

/* ---- vec.h inlines (single-rounding component-wise ops) ---------------- */
static inline void copy_rvec(const rvec a, rvec b) { b[XX] = a[XX]; b[YY] = a[YY]; b[ZZ] = a[ZZ]; }
static inline void copy_ivec(const ivec a, ivec b) { b[XX] = a[XX]; b[YY] = a[YY]; b[ZZ] = a[ZZ]; }
static inline void rvec_add(const rvec a, const rvec b, rvec c)
{ real x = a[XX] + b[XX], y = a[YY] + b[YY], z = a[ZZ] + b[ZZ]; c[XX] = x; c[YY] = y; c[ZZ] = z; }
static inline void rvec_sub(const rvec a, const rvec b, rvec c)
{ real x = a[XX] - b[XX], y = a[YY] - b[YY], z = a[ZZ] - b[ZZ]; c[XX] = x; c[YY] = y; c[ZZ] = z; }
static inline void rvec_dec(rvec a, const rvec b)
{ real x = a[XX] - b[XX], y = a[YY] - b[YY], z = a[ZZ] - b[ZZ]; a[XX] = x; a[YY] = y; a[ZZ] = z; }
static inline void svmul(real s, const rvec v, rvec w)
{ w[XX] = s * v[XX]; w[YY] = s * v[YY]; w[ZZ] = s * v[ZZ]; }
static inline real iprod(const rvec a, const rvec b)
{ return a[XX] * b[XX] + a[YY] * b[YY] + a[ZZ] * b[ZZ]; }
static inline real norm(const rvec a)
{ return (real)sqrt(a[XX] * a[XX] + a[YY] * a[YY] + a[ZZ] * a[ZZ]); }
static inline real sqr(real x) { return x * x; }

#endif /* GC_ORACLE_GMXSHIM_H */
