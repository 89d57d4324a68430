/* Minimal SunRPC XDR-on-stdio declarations (RFC 4506 wire format), because
 * glibc no longer ships <rpc/xdr.h> and libtirpc is not in the image.
 * Implemented in ../fakegmx.c.  TEST INFRASTRUCTURE ONLY. */
#ifndef GC_ORACLE_RPC_XDR_H
#define GC_ORACLE_RPC_XDR_H
#include <stdio.h>
#include <sys/types.h>

typedef int bool_t;
typedef int enum_t;
enum xdr_op { XDR_ENCODE = 0, XDR_DECODE = 1, XDR_FREE = 2 };
typedef struct { enum xdr_op x_op; FILE *x_fp; } XDR;
typedef bool_t (*xdrproc_t)(XDR *, void *, ...);

void xdrstdio_create(XDR *xdrs, FILE *fp, enum xdr_op op);
void xdr_destroy(XDR *xdrs);
bool_t xdr_int(XDR *xdrs, int *ip);
bool_t xdr_u_int(XDR *xdrs, u_int *up);
bool_t xdr_enum(XDR *xdrs, enum_t *ep);
bool_t xdr_float(XDR *xdrs, float *fp);
bool_t xdr_double(XDR *xdrs, double *dp);
bool_t xdr_string(XDR *xdrs, char **cpp, u_int maxsize);
bool_t xdr_vector(XDR *xdrs, char *basep, u_int nelem, u_int elemsize, xdrproc_t proc);
bool_t xdr_array(XDR *xdrs, char **addrp, u_int *sizep, u_int maxsize, u_int elsize,
                 xdrproc_t proc);
#endif
