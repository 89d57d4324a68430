/* Option parsing, index groups and the text writers of the host tools (see gcb_host.h). */
#include "gcb_host.h"
#include <ctype.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static const char *g_program = "gridcount";

void gh_set_program(const char *argv0)
{
  const char *s = strrchr(argv0, '/');
  g_program = s ? s + 1 : argv0;
}
const char *gh_program(void) { return g_program; }

void gh_msg(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
}

void gh_fatal(const char *fmt, ...)
{
  va_list ap;
  fprintf(stderr, "\n-------------------------------------------------------\nProgram %s\nFatal error:\n", g_program);
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fprintf(stderr, "\n-------------------------------------------------------\n");
  exit(1);
}

void *gh_alloc(size_t bytes)
{
  void *p = malloc(bytes ? bytes : 1);
  if (!p) gh_fatal("Out of memory (%zu bytes)", bytes);
  return p;
}

void *gh_zalloc(size_t count, size_t size)
{
  void *p = calloc(count ? count : 1, size ? size : 1);
  if (!p) gh_fatal("Out of memory (%zu x %zu bytes)", count, size);
  return p;
}

void *gh_realloc(void *old, size_t bytes)
{
  void *p = realloc(old, bytes ? bytes : 1);
  if (!p) gh_fatal("Out of memory (%zu bytes)", bytes);
  return p;
}

/* ------------------------------------------------------------------ options */
static int looks_like_option(const char *s)
{
  /* "-x" where x is neither a digit nor '.', so that negative numbers stay values */
  return s[0] == '-' && s[1] && !isdigit((unsigned char)s[1]) && s[1] != '.';
}

static char *dup_with_ext(const char *name, const char *ext)
{
  const char *slash = strrchr(name, '/');
  const char *dot = strrchr(slash ? slash : name, '.');
  size_t n = strlen(name) + strlen(ext) + 1;
  char *out = (char *)gh_alloc(n);
  snprintf(out, n, "%s%s", name, dot ? "" : ext);
  return out;
}

static double parse_number(const char *opt, const char *s)
{
  char *end;
  double v = strtod(s, &end);
  if (end == s || *end) gh_fatal("Expected a number for option %s, got '%s'", opt, s);
  return v;
}

static int is_number(const char *s)
{
  char *end;
  strtod(s, &end);
  return end != s && !*end;
}

static void usage(const char *synopsis, gh_file *files, int nfiles, gh_option *opts, int nopts)
{
  int i;
  printf("%s\n\n%s\n\nFiles:\n", g_program, synopsis);
  for (i = 0; i < nfiles; i++)
    printf("  %-10s %s%-6s %s %s\n", files[i].opt, files[i].defname, files[i].ext,
           files[i].mode == GH_F_READ ? "Input      " : files[i].mode == GH_F_WRITE ? "Output     "
           : files[i].mode == GH_F_OPT_READ ? "Input, Opt." : files[i].mode == GH_F_READ_MULT ? "Input, Mult." : "Output, Opt.",
           files[i].help ? files[i].help : "");
  printf("\nOptions:\n  -h         print this help\n  -b/-e/-dt  first / last time (ps) and time step of frames to use\n");
  for (i = 0; i < nopts; i++) {
    if (!opts[i].help) continue;
    printf("  %s%-12s %s\n", opts[i].type == GH_BOOL ? "-[no]" : "-", opts[i].name + 1, opts[i].help);
  }
}

void gh_parse_args(int argc, char **argv, const char *synopsis, gh_file *files, int nfiles, gh_option *opts,
                   int nopts, gh_timesel *ts)
{
  int i, j;
  gh_set_program(argv[0]);
  for (j = 0; j < nfiles; j++) { files[j].set = 0; files[j].name = NULL; files[j].names = NULL; files[j].nnames = 0; }
  for (j = 0; j < nopts; j++) opts[j].set = 0;
  if (ts) memset(ts, 0, sizeof(*ts));
  for (i = 1; i < argc; i++) {
    const char *a = argv[i];
    int done = 0;
    if (!strcmp(a, "-h") || !strcmp(a, "-help") || !strcmp(a, "--help")) { usage(synopsis, files, nfiles, opts, nopts); exit(0); }
    if (!strcmp(a, "-nice") || !strcmp(a, "-debug") || !strcmp(a, "-xvg")) {      /* accepted and ignored (Gromacs generic) */
      if (i + 1 < argc && !looks_like_option(argv[i + 1])) i++;
      continue;
    }
    if (!strcmp(a, "-quiet") || !strcmp(a, "-w") || !strcmp(a, "-now")) continue;
    if (ts && (!strcmp(a, "-b") || !strcmp(a, "-e") || !strcmp(a, "-dt"))) {
      if (i + 1 >= argc) gh_fatal("Option %s needs a value", a);
      const double v = parse_number(a, argv[++i]);
      if (a[1] == 'b') { ts->have_b = 1; ts->b = v; }
      else if (a[1] == 'e') { ts->have_e = 1; ts->e = v; }
      else { ts->have_dt = 1; ts->dt = v; }
      continue;
    }
    for (j = 0; j < nfiles && !done; j++) {
      if (strcmp(a, files[j].opt)) continue;
      files[j].set = 1;
      done = 1;
      if (files[j].mode == GH_F_READ_MULT) {
        while (i + 1 < argc && !looks_like_option(argv[i + 1])) {
          files[j].names = (char **)gh_realloc(files[j].names, sizeof(char *) * (size_t)(files[j].nnames + 1));
          files[j].names[files[j].nnames++] = dup_with_ext(argv[++i], files[j].ext);
        }
      } else if (i + 1 < argc && !looks_like_option(argv[i + 1])) {
        files[j].name = dup_with_ext(argv[++i], files[j].ext);
      }
    }
    for (j = 0; j < nopts && !done; j++) {
      gh_option *o = &opts[j];
      if (o->type == GH_BOOL) {
        if (!strcmp(a, o->name)) { *(int *)o->value = 1; o->set = 1; done = 1; }
        else if (!strncmp(a, "-no", 3) && !strcmp(a + 3, o->name + 1)) { *(int *)o->value = 0; o->set = 1; done = 1; }
        continue;
      }
      if (strcmp(a, o->name)) continue;
      if (i + 1 >= argc) gh_fatal("Option %s needs a value", a);
      done = 1;
      o->set = 1;
      switch (o->type) {
        case GH_INT: *(int *)o->value = (int)parse_number(a, argv[++i]); break;
        case GH_REAL: *(float *)o->value = (float)parse_number(a, argv[++i]); break;
        case GH_STR: *(const char **)o->value = argv[++i]; break;
        case GH_RVEC: {
          float *v = (float *)o->value;
          v[0] = (float)parse_number(a, argv[++i]);
          if (i + 1 < argc && is_number(argv[i + 1])) {          /* three values, or one for all three */
            if (i + 2 >= argc || !is_number(argv[i + 2])) gh_fatal("Expected one or three numbers for option %s", a);
            v[1] = (float)parse_number(a, argv[++i]);
            v[2] = (float)parse_number(a, argv[++i]);
          } else {
            v[1] = v[2] = v[0];
          }
          break;
        }
        case GH_ENUM: {
          const char **list = (const char **)o->value;
          const char *want = argv[++i];
          int k, found = 0;
          for (k = 1; list[k]; k++)
            if (!strncmp(want, list[k], strlen(want))) { list[0] = list[k]; found = 1; break; }
          if (!found) gh_fatal("Invalid argument %s for option %s", want, a);
          break;
        }
        default: break;
      }
    }
    if (!done) gh_fatal("Invalid command line argument:\n%s", a);
  }
  for (j = 0; j < nfiles; j++) {
    if (files[j].mode == GH_F_READ_MULT) {
      if (files[j].nnames == 0) {
        files[j].names = (char **)gh_alloc(sizeof(char *));
        files[j].names[0] = dup_with_ext(files[j].defname, files[j].ext);
        files[j].nnames = 1;
      }
      files[j].name = files[j].names[0];
    } else if (!files[j].name) {
      files[j].name = dup_with_ext(files[j].defname, files[j].ext);
    }
  }
  for (j = 0; j < nopts; j++)
    if (opts[j].type == GH_ENUM && !opts[j].set) { const char **list = (const char **)opts[j].value; list[0] = list[1]; }
}

static const gh_file *find_file(const gh_file *files, int nfiles, const char *opt)
{
  int j;
  for (j = 0; j < nfiles; j++) if (!strcmp(files[j].opt, opt)) return &files[j];
  gh_fatal("internal: no file option %s", opt);
  return NULL;
}
const char *gh_fn(const gh_file *files, int nfiles, const char *opt) { return find_file(files, nfiles, opt)->name; }
int gh_fn_set(const gh_file *files, int nfiles, const char *opt) { return find_file(files, nfiles, opt)->set; }
int gh_opt_set(const gh_option *opts, int nopts, const char *name)
{
  int j;
  for (j = 0; j < nopts; j++) if (!strcmp(opts[j].name, name)) return opts[j].set;
  return 0;
}

/* ------------------------------------------------------------------ index files */
int gh_ndx_read(const char *path, gh_group **groups, int *ngroups)
{
  FILE *f = fopen(path, "r");
  char line[4096];
  gh_group *g = NULL;
  int ng = 0, cap = 0;
  if (!f) return -1;
  while (fgets(line, sizeof line, f)) {
    char *p = line;
    while (isspace((unsigned char)*p)) p++;
    if (*p == '[') {
      char *q = strchr(p, ']');
      if (!q) { fclose(f); return -2; }
      *q = 0;
      p++;
      while (isspace((unsigned char)*p)) p++;
      while (q > p && isspace((unsigned char)q[-1])) *--q = 0;
      g = (gh_group *)gh_realloc(g, sizeof(gh_group) * (size_t)(ng + 1));
      g[ng].name = strdup(p); g[ng].n = 0; g[ng].atoms = NULL;
      ng++;
      cap = 0;
      continue;
    }
    if (ng == 0) continue;
    while (*p) {
      char *end;
      long v = strtol(p, &end, 10);
      if (end == p) break;
      if (g[ng - 1].n == cap) { cap = cap ? 2 * cap : 256; g[ng - 1].atoms = (int *)gh_realloc(g[ng - 1].atoms, sizeof(int) * (size_t)cap); }
      g[ng - 1].atoms[g[ng - 1].n++] = (int)v - 1;          /* 1-based on disk (count.c:173-174) */
      p = end;
    }
  }
  fclose(f);
  *groups = g; *ngroups = ng;
  return 0;
}

const gh_group *gh_ndx_select(const gh_group *groups, int ngroups)
{
  char buf[256];
  int i;
  if (ngroups <= 0) return NULL;
  if (ngroups == 1) {
    gh_msg("Selected %d: '%s'\n", 0, groups[0].name);
    return &groups[0];
  }
  for (i = 0; i < ngroups; i++) gh_msg("Group %5d (%15s) has %5d elements\n", i, groups[i].name, groups[i].n);
  for (;;) {
    gh_msg("Select a group: ");
    if (!fgets(buf, sizeof buf, stdin)) gh_fatal("Cannot read the group selection from standard input");
    char *p = buf, *end;
    while (isspace((unsigned char)*p)) p++;
    long v = strtol(p, &end, 10);
    if (end != p && v >= 0 && v < ngroups) { gh_msg("Selected %ld: '%s'\n", v, groups[v].name); return &groups[v]; }
    for (end = p + strlen(p); end > p && isspace((unsigned char)end[-1]);) *--end = 0;
    for (i = 0; i < ngroups; i++)
      if (!strcmp(p, groups[i].name)) { gh_msg("Selected %d: '%s'\n", i, groups[i].name); return &groups[i]; }
  }
}

void gh_ndx_free(gh_group *groups, int ngroups)
{
  int i;
  for (i = 0; i < ngroups; i++) { free(groups[i].name); free(groups[i].atoms); }
  free(groups);
}

/* ------------------------------------------------------------------ xvg / xfarbe text (xf.c:28-132) */
static const unsigned k_xf_colours[] = {0x000099, 0x1111DD, 0x5555FF, 0x6677DD, 0x8899AA, 0xAABB66, 0xCCDD33, 0xEEEE00,
                                        0xFFFF00, 0xFFDD00, 0xFFBB00, 0xFF9900, 0xFF6600, 0xEE1100, 0xBB0000, 0x770000};
#define XF_NCOLS ((int)(sizeof k_xf_colours / sizeof k_xf_colours[0]))
#define XF_MAX_HEADER 79

int gh_xf_ncols(void) { return XF_NCOLS; }

FILE *gh_xvg_open(const char *path, const char *title, const char *subtitle, const char *xaxis, const char *yaxis)
{
  FILE *f = fopen(path, "w");
  time_t now;
  if (!f) gh_fatal("Cannot open %s for writing", path);
  /* the first two lines run together in the reference too (no newline after "project file") */
  fprintf(f, "# Grace project file# This file was created by %s\n", g_program);
  time(&now);
  fprintf(f, "# All this happened at: %s", ctime(&now));
  fprintf(f, "#\n@version 50102\n@with g0\n");
  fprintf(f, "@    title \"%s\"\n@    subtitle \"%s\"\n", title, subtitle);
  fprintf(f, "@    xaxis  label \"%s\"\n@    yaxis  label \"%s\"\n", xaxis, yaxis);
  fprintf(f, "@target G0.S0\n@type nxy\n");
  return f;
}

FILE *gh_xf_open(const char *path, const char *header, int nx, int ny)
{
  FILE *f = fopen(path, "w");
  if (!f) gh_fatal("Cannot open %s for writing", path);
  if (header && header[0]) fprintf(f, "%.*s", XF_MAX_HEADER - 1, header);      /* xfarbe misreads longer headings */
  fprintf(f, "\n%d %d\n", nx, ny);
  return f;
}

void gh_xf_levels(FILE *f, int ncols, float max)
{
  const float delta = max / (float)(ncols - 1);
  int i;
  fprintf(f, "%d\n", ncols - 1);
  fprintf(f, "%d  %f\n", 0, delta / 2);          /* the first level filters out noise */
  for (i = 1; i < ncols - 1; i++) fprintf(f, "%d  %f\n", i, (float)i * delta);
  fprintf(f, "%d\n", ncols - 1);
}

void gh_xf_axes(FILE *f, float x1, float x2, float y1, float y2)
{
  fprintf(f, "%f  %f\n%f  %f\n", x1, x2, y1, y2);
}

int gh_xf_write_resource(void)
{
  static const char *onoff[][2] = {{"autolev", "off"}, {"xyratio", "on"}, {"legend", "on"}, {"legend_format", "%4.2f"},
                                   {"x-annotation", "on"}, {"y-annotation", "on"}};
  FILE *f = fopen("XFarbe", "w");
  size_t i;
  if (!f) { gh_msg("error: Cannot open file XFarbe for writing.\n"); return 0; }
  gh_msg("Writing resource file './XFarbe' for xfarbe with %d colours.\n", XF_NCOLS);
  fprintf(f, "xfarbe.ncol: %d\nxfarbe.mode: %d\n", XF_NCOLS, 0);
  for (i = 0; i < sizeof onoff / sizeof onoff[0]; i++) fprintf(f, "xfarbe.%s: %s\n", onoff[i][0], onoff[i][1]);
  fprintf(f, "xfarbe.x-annotation-incr: %d\nxfarbe.y-annotation-incr: %d\n", 0, 0);
  for (i = 0; i < (size_t)XF_NCOLS; i++) fprintf(f, "xfarbe.col%d:   #%.6X\n", (int)i, k_xf_colours[i]);
  fclose(f);
  return 1;
}
