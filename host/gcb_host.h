/*
 * gcb_host.h -- plain C host layer of the drop-in tools g_ri3Dc / a_ri3Dc / a_gridcalc.
 *
 * The reference tools are thin shells around Gromacs 4.5 (parse_common_args, read_tps_conf,
 * get_index, read_first_x/read_next_x; g_ri3Dc.c:336-392, a_ri3Dc.c:427-474).  Gromacs is not
 * available, so this layer provides the small part of it the tools use: the option syntax
 * (-opt value, -[no]bool, one-or-three values for vectors, default file names and
 * extensions), an .ndx reader, and trajectory readers for XTC (own xdr3dfcoord codec), the raw
 * GCTRJ1 format of the test harness and multi-frame .gro.  All numerical work goes through the
 * C ABI in include/gridcount_b200.h; nothing here touches the GPU directly.
 */
#ifndef GCB_HOST_H
#define GCB_HOST_H
#include <stdio.h>
#include <stddef.h>
#include <stdint.h>
#include "gridcount_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- messages (utilgmx.c:124-176: msg -> stderr, fatal -> message + exit) ---- */
void gh_set_program(const char *argv0);
const char *gh_program(void);
void gh_msg(const char *fmt, ...);
void gh_fatal(const char *fmt, ...);
/* checked allocation: like Gromacs' snew/srenew (smalloc.h) these end the program with a message instead of
 * returning NULL */
void *gh_alloc(size_t bytes);
void *gh_zalloc(size_t count, size_t size);
void *gh_realloc(void *p, size_t bytes);

/* ---- options ---------------------------------------------------------------- */
typedef enum { GH_BOOL, GH_INT, GH_REAL, GH_RVEC, GH_STR, GH_ENUM } gh_opt_type;
typedef struct {
  const char *name;        /* with the leading '-' */
  gh_opt_type type;
  void *value;             /* int* (BOOL, INT), float* (REAL), float[3] (RVEC), const char** (STR), const char** list (ENUM) */
  const char *help;        /* NULL: hidden */
  int set;                 /* out: the user gave it (opt2parg_bSet) */
} gh_option;
/* ENUM: value points to a NULL-terminated list whose element [0] receives the choice, Gromacs style */

typedef enum { GH_F_READ, GH_F_WRITE, GH_F_OPT_READ, GH_F_OPT_WRITE, GH_F_READ_MULT } gh_file_mode;
typedef struct {
  const char *opt;         /* "-f", "-grid", ... */
  const char *defname;     /* default base name, e.g. "gridxdr" */
  const char *ext;         /* default extension incl. dot, e.g. ".dat" */
  gh_file_mode mode;
  const char *help;
  int set;                 /* out */
  char *name;              /* out: resolved file name (first one for READ_MULT) */
  char **names; int nnames;/* out: all names for READ_MULT */
} gh_file;

typedef struct { int have_b, have_e, have_dt; double b, e, dt; } gh_timesel;   /* -b -e -dt (PCA_CAN_TIME) */

/* Parses argv; prints usage and exits on -h; calls gh_fatal on a bad command line. */
void gh_parse_args(int argc, char **argv, const char *synopsis, gh_file *files, int nfiles, gh_option *opts,
                   int nopts, gh_timesel *ts);
const char *gh_fn(const gh_file *files, int nfiles, const char *opt);        /* opt2fn */
int gh_fn_set(const gh_file *files, int nfiles, const char *opt);            /* opt2bSet */
int gh_opt_set(const gh_option *opts, int nopts, const char *name);          /* opt2parg_bSet */

/* ---- index groups ----------------------------------------------------------- */
typedef struct { char *name; int n; int *atoms; } gh_group;                   /* atoms are 0-based */
int gh_ndx_read(const char *path, gh_group **groups, int *ngroups);
/* one group: taken; several: the list goes to stderr and the choice (number or name) is read from stdin */
const gh_group *gh_ndx_select(const gh_group *groups, int ngroups);
void gh_ndx_free(gh_group *groups, int ngroups);

/* ---- trajectories ----------------------------------------------------------- */
typedef struct gh_traj gh_traj;
typedef struct { int natoms; int step; float time; float box[9]; float prec; } gh_frame_info;
gh_traj *gh_traj_open(const char *path);                  /* by extension: .xtc, .gro, anything else GCTRJ1 */
int gh_traj_natoms(gh_traj *t);
/* reads the next frame into x[natoms][3]; returns 1, 0 at end of file, <0 on a damaged file */
int gh_traj_next(gh_traj *t, gh_frame_info *info, float *x);
void gh_traj_close(gh_traj *t);
/* frames selected by -b/-e/-dt, as read_next_x would deliver them */
int gh_traj_next_selected(gh_traj *t, const gh_timesel *ts, gh_frame_info *info, float *x);
/* topology substitute for -s: .gro (title, natoms, atoms, box) or the GCTOP1 stub "GCTOP1 <natoms>" */
int gh_top_natoms(const char *path, int *natoms, float box9[9]);

/* XTC codec (Gromacs xdr3dfcoord; magic 1995) */
typedef struct gh_xtc_writer gh_xtc_writer;
gh_xtc_writer *gh_xtc_create(const char *path);
int gh_xtc_write(gh_xtc_writer *w, int natoms, int step, float time, const float box9[9], const float *x, float prec);
void gh_xtc_close(gh_xtc_writer *w);
/* in-memory codec: returns bytes written / consumed, 0 on error */
size_t gh_xtc_encode_frame(int natoms, int step, float time, const float box9[9], const float *x, float prec,
                           uint8_t *out, size_t cap);
size_t gh_xtc_decode_frame(const uint8_t *in, size_t len, gh_frame_info *info, float *x, int max_atoms);

/* ---- text output (xf.c:28-132) ---------------------------------------------- */
FILE *gh_xvg_open(const char *path, const char *title, const char *subtitle, const char *xaxis, const char *yaxis);
FILE *gh_xf_open(const char *path, const char *header, int nx, int ny);
void gh_xf_levels(FILE *f, int ncols, float max);
void gh_xf_axes(FILE *f, float x1, float x2, float y1, float y2);
int gh_xf_write_resource(void);       /* ./XFarbe */
int gh_xf_ncols(void);

#ifdef __cplusplus
}
#endif
#endif
