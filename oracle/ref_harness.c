/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * Function-level access to the UNMODIFIED reference for the parity tests and
 * for bench.py's CPU-reference arm.  gridcount() is `static` in
 * /root/reference/src/g_ri3Dc.c:115,139, so this translation unit textually
 * includes that file (with its main() renamed and never called) and wraps the
 * static function.  Everything else (autoset_cavity, setup_tgrid, grid_write,
 * grid_read, ...) is linked from the reference's own library objects.
 *
 * Built only where /root/reference exists; the resulting oracle/_ref/libgcref.so
 * travels to the GPU box as a binary.
 */
#define main gcref_unused_g_ri3Dc_main
#include REF_G_RI3DC_C
#undef main

#include <pthread.h>
#include <time.h>
#include <unistd.h>
#include <fcntl.h>

/* The reference chats on stderr through msg() (utilgmx.c:124-137); keep the
 * test logs readable unless GCREF_VERBOSE is set. */
static int quiet_begin(void)
{
  int saved, nul;
  if (getenv("GCREF_VERBOSE")) return -1;
  fflush(stderr);
  saved = dup(2);
  nul = open("/dev/null", O_WRONLY);
  if (nul >= 0) { dup2(nul, 2); close(nul); }
  return saved;
}
static void quiet_end(int saved)
{
  if (saved < 0) return;
  fflush(stderr);
  dup2(saved, 2);
  close(saved);
}

typedef struct {
  t_tgrid tg;
  t_cavity cav;
} gcref_grid;

/* setmask bits: 1 = -cpoint given, 2 = -R, 4 = -z1, 8 = -z2 (count.c:28-48) */
void *gcref_grid_create(const float box9[9], const float delta[3], int setmask,
                        const float cpoint[3], float radius, float z1, float z2)
{
  gcref_grid *g = calloc(1, sizeof(*g));
  matrix box;
  rvec Delta;
  int i, j;
  t_pargs pa[4] = {
    { "-cpoint", FALSE, etRVEC, {NULL}, "" }, { "-R", FALSE, etREAL, {NULL}, "" },
    { "-z1", FALSE, etREAL, {NULL}, "" },     { "-z2", FALSE, etREAL, {NULL}, "" } };
  for (i = 0; i < 4; i++) pa[i].bSet = (setmask >> i) & 1;
  for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) box[i][j] = box9[3 * i + j];
  g->cav.axis[ZZ] = 1;
  if (cpoint) for (i = 0; i < 3; i++) g->cav.cpoint[i] = cpoint[i];
  g->cav.radius = radius; g->cav.z1 = z1; g->cav.z2 = z2;
  for (i = 0; i < 3; i++) Delta[i] = delta[i];
  {
    int q = quiet_begin(), ok;
    autoset_cavity(&g->cav, box, 4, pa);
    ok = setup_tgrid(&g->tg, &g->cav, Delta);
    quiet_end(q);
    if (!ok) { free(g); return NULL; }
  }
  return g;
}

void gcref_grid_info(void *h, int mx[3], float a[3], float b[3], float delta[3], float cav7[7])
{
  gcref_grid *g = h;
  int d;
  for (d = 0; d < 3; d++) { mx[d] = g->tg.mx[d]; a[d] = g->tg.a[d]; b[d] = g->tg.b[d]; delta[d] = g->tg.Delta[d]; }
  if (cav7) {
    for (d = 0; d < 3; d++) cav7[d] = g->cav.cpoint[d];
    cav7[3] = g->cav.radius; cav7[4] = g->cav.z1; cav7[5] = g->cav.z2; cav7[6] = g->cav.vol;
  }
}

float *gcref_grid_data(void *h) { return ((gcref_grid *)h)->tg.grid[0][0]; }
void gcref_grid_set_tweight(void *h, float tw) { ((gcref_grid *)h)->tg.tweight = tw; }
float gcref_grid_tweight(void *h) { return ((gcref_grid *)h)->tg.tweight; }
void gcref_grid_free(void *h) { gcref_grid *g = h; if (g) { free_tgrid(&g->tg); free(g); } }
double gcref_deltaV(void *h) { return DeltaV(&((gcref_grid *)h)->tg); }

/* The reference's inner loop (g_ri3Dc.c:425-426) over nframes frames held in
 * memory; weights[f] is the per-frame `weight` of g_ri3Dc.c:416. */
void gcref_bin_frames(void *h, const float *x, long nframes, int natoms,
                      const int *gindex, int gnx, const float *weights, double *sumP)
{
  gcref_grid *g = h;
  long f;
  int i;
  double s = *sumP;
  for (f = 0; f < nframes; f++) {
    rvec *xf = (rvec *)(x + (size_t)f * natoms * 3);
    real w = weights ? weights[f] : 1;
    for (i = 0; i < gnx; i++)
      s += gridcount(&g->tg, xf[gindex[i]], w);
  }
  *sumP = s;
}

int gcref_grid_write(void *h, const char *path, char *header)
{
  gcref_grid *g = h;
  FILE *fp = fopen(path, "w");
  int ok;
  if (!fp) return 0;
  ok = grid_write(fp, &g->tg, header);
  fclose(fp);
  return ok;
}

void *gcref_grid_read(const char *path, char *header256)
{
  gcref_grid *g = calloc(1, sizeof(*g));
  FILE *fp = fopen(path, "r");
  int d;
  if (!fp) { free(g); return NULL; }
  {
    int q = quiet_begin(), ok = grid_read(fp, &g->tg, header256);
    quiet_end(q);
    if (!ok) { fclose(fp); free(g); return NULL; }
  }
  fclose(fp);
  for (d = 0; d < 3; d++) g->tg.b[d] = g->tg.a[d] + g->tg.mx[d] * g->tg.Delta[d]; /* as update_tgrid */
  return g;
}

/* ---- multi-threaded CPU baseline: one private reference grid per thread,
 * disjoint frame shards, i.e. the "run per segment and merge" workflow the
 * reference documents (a_gridcalc.c:180-189, FAQ:300-349). ---------------- */
typedef struct {
  gcref_grid *g;
  const float *x; long f0, f1; int natoms; const int *gindex; int gnx; float w; double sumP;
  long cells; int id; pthread_barrier_t *bar; struct timespec *t0;
} gcref_job;

static void *gcref_worker(void *arg)
{
  gcref_job *j = arg;
  long f;
  int i;
  double s = 0;
  /* every thread zeroes its own grid (first touch and clearing are not timed), then all start together */
  memset(gcref_grid_data(j->g), 0, sizeof(real) * (size_t)j->cells);
  pthread_barrier_wait(j->bar);
  if (j->id == 0) clock_gettime(CLOCK_MONOTONIC, j->t0);
  for (f = j->f0; f < j->f1; f++) {
    rvec *xf = (rvec *)(j->x + (size_t)f * j->natoms * 3);
    for (i = 0; i < j->gnx; i++) s += gridcount(&j->g->tg, xf[j->gindex[i]], j->w);
  }
  j->sumP = s;
  return NULL;
}

/* the per-thread grids are kept between calls with the same geometry and thread count: allocating and
 * page-faulting 16 x 256 MB per call costs seconds, many times the binning that is being timed */
static gcref_grid *pool[1024];
static int pool_n = 0;
static float pool_key[12];

/* Returns elapsed seconds for binning nframes frames with nthreads threads;
 * *sumP_out gets the merged hit weight, merged_out (optional, cells floats)
 * the sum of the per-thread grids. */
double gcref_bench(const float box9[9], const float delta[3], const float *x, long nframes,
                   int natoms, const int *gindex, int gnx, float weight, int nthreads,
                   double *sumP_out, float *merged_out)
{
  gcref_job *jobs = calloc((size_t)nthreads, sizeof(*jobs));
  pthread_t *th = calloc((size_t)nthreads, sizeof(*th));
  pthread_barrier_t bar;
  struct timespec t0, t1;
  float key[12];
  int t;
  long cells = 0, c;
  double s = 0;
  if (nthreads > 1024) nthreads = 1024;
  memcpy(key, box9, sizeof(float) * 9); memcpy(key + 9, delta, sizeof(float) * 3);
  if (pool_n != nthreads || memcmp(key, pool_key, sizeof key) != 0) {
    for (t = 0; t < pool_n; t++) gcref_grid_free(pool[t]);
    for (t = 0; t < nthreads; t++) pool[t] = gcref_grid_create(box9, delta, 0, NULL, 0, 0, 0);
    pool_n = nthreads;
    memcpy(pool_key, key, sizeof key);
  }
  pthread_barrier_init(&bar, NULL, (unsigned)nthreads);
  for (t = 0; t < nthreads; t++) {
    jobs[t].g = pool[t];
    jobs[t].x = x; jobs[t].natoms = natoms; jobs[t].gindex = gindex; jobs[t].gnx = gnx; jobs[t].w = weight;
    jobs[t].f0 = nframes * t / nthreads; jobs[t].f1 = nframes * (t + 1) / nthreads;
    cells = (long)jobs[t].g->tg.mx[0] * jobs[t].g->tg.mx[1] * jobs[t].g->tg.mx[2];
    jobs[t].cells = cells; jobs[t].id = t; jobs[t].bar = &bar; jobs[t].t0 = &t0;
  }
  for (t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, gcref_worker, &jobs[t]);
  gcref_worker(&jobs[0]);
  for (t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  pthread_barrier_destroy(&bar);
  for (t = 0; t < nthreads; t++) {
    s += jobs[t].sumP;
    if (merged_out) {
      float *gd = gcref_grid_data(jobs[t].g);
      for (c = 0; c < cells; c++) merged_out[c] = (t ? merged_out[c] : 0) + gd[c];
    }
  }
  if (sumP_out) *sumP_out = s;
  free(jobs); free(th);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
