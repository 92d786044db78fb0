/* oracle/shim/metis.h -- TEST INFRASTRUCTURE ONLY (never part of the product).
 *
 * The reference includes <metis.h> (partitioning.cc:25) and calls
 * METIS_PartGraphRecursive with 32-bit `int*` arrays (partitioning.cc:153-155,
 * 186-188).  The only METIS available offline is the CUDA toolkit's
 * libmetis_static.a, built with idx_t = int64_t / real_t = float and shipped
 * without a header.  This shim declares the real entry point and adds a C++
 * overload taking the reference's `int*` arguments that widens the graph,
 * calls METIS, and narrows the partition vector back.
 */
#ifndef DC_ORACLE_METIS_SHIM_H
#define DC_ORACLE_METIS_SHIM_H

#include <stdint.h>
#include <stdlib.h>

extern "C" int METIS_PartGraphRecursive(int64_t *nvtxs, int64_t *ncon, int64_t *xadj,
                                        int64_t *adjncy, int64_t *vwgt, int64_t *vsize,
                                        int64_t *adjwgt, int64_t *nparts, float *tpwgts,
                                        float *ubvec, int64_t *options, int64_t *objval,
                                        int64_t *part);

static inline int METIS_PartGraphRecursive(int *nvtxs, int *ncon, int *xadj, int *adjncy,
                                           void *, void *, void *, int *nparts, void *,
                                           void *, void *, int *objval, int *part)
{
    int64_t n = *nvtxs, nc = *ncon, np = *nparts, obj = 0;
    int64_t *x = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int64_t *a = (int64_t *)malloc(sizeof(int64_t) * (size_t)(xadj[n] > 0 ? xadj[n] : 1));
    int64_t *p = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i <= n; i++) x[i] = xadj[i];
    for (int64_t i = 0; i < xadj[n]; i++) a[i] = adjncy[i];
    int rc = METIS_PartGraphRecursive(&n, &nc, x, a, (int64_t *)0, (int64_t *)0, (int64_t *)0,
                                      &np, (float *)0, (float *)0, (int64_t *)0, &obj, p);
    for (int64_t i = 0; i < n; i++) part[i] = (int)p[i];
    *objval = (int)obj;
    free(p); free(a); free(x);
    return rc;
}

#endif
