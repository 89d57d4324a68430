/*
 * gc_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement, in plain C on flat arrays, of the reference algorithm for
 * the gridcount hot path.  It is the checker for the
 * CUDA path: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may load it.  The product library (gridcount_b200/) never links or calls
 * anything in oracle/.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_*.py,
 * -m "not gpu") against outputs of the UNMODIFIED reference sources compiled
 * from /root/reference by oracle/Makefile (oracle/_ref) -- live when oracle/_ref
 * is present, and through the fixtures committed under tests/golden/ (made by
 * tests/golden/make_golden.py) everywhere.  The reference ships no tests or
 * golden vectors of its own.
 *
 * All citations are file:line into /root/reference.
 */
#ifndef GC_ORACLE_H
#define GC_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* GMX_REAL_EPS of a single-precision Gromacs 4.5.x (grid3D.c:143,146) */
#define GCO_GMX_REAL_EPS 5.96046448E-08

typedef struct {               /* t_cavity, include/count.h:71-77 */
  float axis[3], cpoint[3];
  float radius, z1, z2, vol;
} gco_cavity;

typedef struct {               /* t_tgrid without the data, include/grid3D.h:63-71 */
  int mx[3];
  float a[3], b[3], delta[3];
  float tweight;
} gco_geom;

enum { GCO_SET_CPOINT = 1, GCO_SET_R = 2, GCO_SET_Z1 = 4, GCO_SET_Z2 = 8 };
enum { GCO_UNIT_UNITY, GCO_UNIT_SPC, GCO_UNIT_MOLAR, GCO_UNIT_ANG, GCO_UNIT_VOXEL, GCO_UNIT_NR };

/* ---- geometry (count.c:25-64, grid3D.c:125-203) ---------------------- */
int gco_autoset_cavity(gco_cavity *cav, const float box9[9], int setmask);
int gco_setup_grid(gco_geom *g, gco_cavity *cav, const float delta[3]);
double gco_delta_v(const gco_geom *g);          /* grid3D.c:108-110 */
/* 1 if x = b[d]-1ulp would bin to mx[d] in any dimension */
int gco_edge_vulnerable(const gco_geom *g);

/* ---- frame weights (g_ri3Dc.c:390-394, 413-416) ----------------------- */
void gco_frame_weights(const float *t, long nframes, int dtweight,
                       float *weight, double *T_tot);

/* ---- binning (g_ri3Dc.c:139-152, 425-426) ------------------------------ */
/* float accumulation exactly as the reference; grid is z-fastest */
void gco_bin_frames(const gco_geom *g, float *grid, const float *x, long nframes,
                    int natoms, const int *gindex, int gnx, const float *weight,
                    double *sumP, uint64_t *hits_per_frame, uint64_t *edge_overflow);
/* the same with `real = double` (a -DDOUBLE build): f64 accumulation, the yardstick for f32 summation error */
void gco_bin_frames_f64(const gco_geom *g, double *grid, const float *x, long nframes,
                        int natoms, const int *gindex, int gnx, const float *weight);
/* integer hit counts (what the CUDA path accumulates) */
void gco_count_frames(const gco_geom *g, uint32_t *counts, const float *x, long nframes,
                      int natoms, const int *gindex, int gnx,
                      uint64_t *hits_per_frame, uint64_t *edge_overflow);
/* optional rectangular single-image wrap, definition from count.c:88-92 */
void gco_wrap_frames(float *x, long nframes, int natoms, const float *box9_per_frame);

/* value of n sequential additions of w to v in f32 / f64 (g_ri3Dc.c:149,151) */
float gco_seq_add_f32(float v, float w, uint64_t n);
double gco_seq_add_f64(double v, double w, uint64_t n);

/* ---- normalisation + header (g_ri3Dc.c:451-473) ------------------------ */
void gco_normalise(gco_geom *g, float *grid, double T_tot);
int gco_header(char out[256], const char *subtitle, double T_tot, const char *grpname,
               const gco_cavity *cav, const gco_geom *g, double sumP,
               const char *trxname, int dtweight);

/* ---- XDR grid file (grid3D.c:22-106, xdr_grid.c:89-158) ---------------- */
size_t gco_xdr_size(const gco_geom *g, const char *header);
size_t gco_xdr_encode(const gco_geom *g, const float *grid, const char *header,
                      uint8_t *out, size_t cap);
/* returns 0 on success; *grid_out is malloc'ed (z-fastest) */
int gco_xdr_decode(const uint8_t *buf, size_t len, gco_geom *g, char header[257],
                   float **grid_out);
/* a_gridcalc.c:132-189 */
void gco_gridcalc(int ngrids, const float *const *grids, const float *tweights,
                  size_t cells_x, size_t cells_y, size_t cells_z, float *avg, float *tweight_out);

/* ---- gOpenMol plt / ascii (plt.c:42-99) -------------------------------- */
size_t gco_plt_encode(const gco_geom *g, const float *grid, float unit, uint8_t *out, size_t cap);

/* ---- a_ri3Dc analysis (a_ri3Dc.c:86-201, 488-762) ---------------------- */
typedef struct {
  /* inputs */
  int unit;              /* GCO_UNIT_* */
  float min_occ;         /* -minocc, in the chosen unit (a_ri3Dc.c:333,517) */
  int rad_ba_nr, rad_ba_step;   /* hidden -radnr/-radstep (a_ri3Dc.c:731-757) */
  /* derived geometry */
  gco_cavity geometry;   /* tgrid2cavity, a_ri3Dc.c:94-107 */
  float DeltaR;
  double dV;
  int iradNR;
  float unit_value;      /* DensUnit[nunit] after the voxel fix-up (a_ri3Dc.c:516) */
  /* scalars */
  double sumP, sumH;
  float volume, reff, average, height;
  int ivol;
  /* arrays, malloc'ed by gco_analyse; 2D arrays row-major [n0][n1] as grid2_alloc */
  float *xyp, *xzp, *yzp, *hxyp;         /* [mx][my], [mx][mz], [my][mz], [mx][my] */
  float *zdf, *lzdf, *profile;           /* [mz][2] */
  float *rweight;                        /* [iradNR] */
  float *rdf, *hrdf;                     /* [iradNR][2] */
  float *rzp, *hrzp;                     /* [iradNR][mz] */
} gco_analysis;

/* geometry helpers of a_ri3Dc */
void gco_update_b(gco_geom *g);                                   /* a_ri3Dc.c:86-92 */
void gco_tgrid2cavity(const gco_geom *g, gco_cavity *c);          /* a_ri3Dc.c:94-107 */
/* a_ri3Dc.c:133-201; returns 1 if the grid was re-cropped (then *grid is
 * replaced by a new malloc'ed block), 0 if untouched, <0 on the fatal case */
int gco_readjust(gco_geom *g, float **grid, gco_cavity *user, int setmask);
int gco_analyse(const gco_geom *g, const float *grid, gco_analysis *an);
void gco_analysis_free(gco_analysis *an);

/* ---- counter-based synthetic trajectories ------------------ */
enum { GCO_GEN_UNIFORM = 0, GCO_GEN_QUANTISED = 1, GCO_GEN_PORE = 2, GCO_GEN_CLUSTER = 3 };
typedef struct {
  uint64_t seed;
  int mode;
  float L[3];            /* box edge lengths */
  float slab_lo, slab_hi, pore_r;   /* GCO_GEN_PORE */
  int nsites; float sigma;          /* GCO_GEN_CLUSTER */
  int box_period; float box_amp;    /* >0: breathing box, triangle wave */
} gco_gen;
void gco_gen_frames(const gco_gen *p, long frame0, long nframes, int natoms, float *x);
void gco_gen_box(const gco_gen *p, long frame, float L_out[3]);

/* ---- multi-threaded port timing (bench.py cpu_baseline kind "port") ---- */
double gco_bench(const gco_geom *g, const float *x, long nframes, int natoms,
                 const int *gindex, int gnx, float weight, int nthreads, double *sumP_out);

#ifdef __cplusplus
}
#endif
#endif
