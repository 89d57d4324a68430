/* Stand-in for Gromacs 4.5 statutil.h: trajectory reader + option parser
 * prototypes used by the three reference main()s.  TEST INFRASTRUCTURE ONLY. */
#ifndef GC_ORACLE_STATUTIL_H
#define GC_ORACLE_STATUTIL_H
#include "gmxshim.h"

#define TRX_READ_X (1 << 0)
#define TRX_NEED_X (1 << 1)
#define PCA_CAN_VIEW (1 << 5)
#define PCA_CAN_BEGIN (1 << 6)
#define PCA_CAN_END (1 << 7)
#define PCA_CAN_DT (1 << 14)
#define PCA_CAN_TIME (PCA_CAN_BEGIN | PCA_CAN_END | PCA_CAN_DT)

typedef struct {
  int flags;
  int natoms;
  gmx_bool bTime;
  real time;
  gmx_bool bX;
  rvec *x;
  gmx_bool bBox;
  matrix box;
} t_trxframe;

void CopyRight(FILE *out, const char *szProgram);
void thanx(FILE *fp);
void parse_common_args(int *argc, char *argv[], unsigned long Flags,
                       int nfile, t_filenm fnm[], int npargs, t_pargs *pa,
                       int ndesc, const char **desc, int nbugs, const char **bugs,
                       output_env_t *oenv);
gmx_bool read_tps_conf(const char *infile, char *title, t_topology *top, int *ePBC,
                       rvec **x, rvec **v, matrix box, gmx_bool bMass);
void get_index(t_atoms *atoms, const char *fnm, int ngrps, int isize[],
               atom_id *index[], char *grpnames[]);
int read_first_x(const output_env_t oenv, t_trxstatus **status, const char *fn,
                 real *t, rvec **x, matrix box);
gmx_bool read_next_x(const output_env_t oenv, t_trxstatus *status, real *t,
                     int natoms, rvec x[], matrix box);
gmx_bool read_first_frame(const output_env_t oenv, t_trxstatus **status,
                          const char *fn, t_trxframe *fr, int flags);
gmx_bool read_next_frame(const output_env_t oenv, t_trxstatus *status, t_trxframe *fr);
void close_trj(t_trxstatus *status);

const char *ftp2fn(int ftp, int nfile, t_filenm fnm[]);
const char *ftp2fn_null(int ftp, int nfile, t_filenm fnm[]);
const char *opt2fn(const char *opt, int nfile, t_filenm fnm[]);
int opt2fns(char **fns[], const char *opt, int nfile, t_filenm fnm[]);
gmx_bool opt2bSet(const char *opt, int nfile, t_filenm fnm[]);
#endif
