/* capture.h -- TEST INFRASTRUCTURE ONLY.  Force-included (gcc -include) when
 * compiling the UNMODIFIED /root/reference/src/a_ri3Dc.c so that every
 * floating value its main() prints is also appended, as the raw double that
 * was handed to printf, to "<output file>.f64" (stdout -> "stdout.f64") when
 * GCSHIM_CAPTURE=1.  This is how the analysis arrays, which are locals of the
 * reference main(), are recovered at full precision for the parity tests. */
#ifndef GC_ORACLE_CAPTURE_H
#define GC_ORACLE_CAPTURE_H
#include <stdio.h>
int gcshim_fprintf(FILE *fp, const char *fmt, ...);
int gcshim_printf(const char *fmt, ...);
FILE *gcshim_fopen(const char *fn, const char *mode);
int gcshim_fclose(FILE *fp);
#define fprintf gcshim_fprintf
#define printf gcshim_printf
#define fopen gcshim_fopen
#define fclose gcshim_fclose
#endif
