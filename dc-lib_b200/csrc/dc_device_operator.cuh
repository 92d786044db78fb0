/* dc_device_operator.cuh -- the device-operator hook.
 *
 * The reference takes the element kernel as a HOST function pointer
 * (userSeqFct / userVecFct of DC_tree_traversal, src/tree_traversal.cc:106-117).  A host
 * pointer cannot run on the GPU, and a __device__ function pointer called through the
 * library's kernels would defeat inlining, so the hook is a *template*: the user's
 * translation unit includes this header, defines an operator functor and instantiates the
 * library's kernels with it through one macro:
 *
 *     struct MyOp : dcdev::KeOperator<MyOp, 1, true> {                 // D = 1, symmetric
 *         static constexpr int NPARAM = 0;
 *         __device__ static void ke(const double (&x)[4][3], const double *params, double *k);
 *     };                                                               // k[(4*a+b)*D*D + i*D + j]
 *     DC_DEVICE_OPERATOR(my_op, MyOp)       // defines:  extern "C" int my_op(...launcher ABI...)
 *
 * and the application passes the launcher to dc_cuda_assemble_with() (include/dc_cuda.h).
 * KeOperator is the convenience form: ke() fills the whole element matrix exactly like a host
 * callback would, the record kept per element is that matrix, and a contribution is a plain
 * `+=` of one block.  Operators that want the fast path implement the full contract of
 * dc_operators.cuh (setup / fetch / apply ...) like the built-ins do.
 *
 * nvcc -gencode arch=compute_100a,code=sm_100a -I<repo>/include -I<repo>/dc-lib_b200/csrc ...
 */
#ifndef DC_DEVICE_OPERATOR_CUH
#define DC_DEVICE_OPERATOR_CUH

#include <cuda_runtime.h>
#include "dc_cuda.h"
#include "dc_kernels.cuh"

/* What the library hands to a launcher: its device arrays and the geometry of the schedule.
 * Filled by dc_cuda_assemble / dc_cuda_assemble_with; opaque to operator code. */
struct dc_launch_args {
    int mode, threads, smCount, symmetric, colBase, persistent;
    int nbElem, nbNodes;
    int64_t nnz, nbUnits, nbBigEdges, nbListEdges;
    int maxSlots, maxHalo, maxNodes, maxTasks, maxItems, maxTgt;
    const int *d_conn, *d_row, *d_col;
    const double *d_coord;
    const dc_unit_t *d_units;
    const dc_task_t *d_tasks;
    const uint16_t *d_items, *d_slotNodes;
    const int *d_haloNodes, *d_diagRel, *d_tgt;
    const int64_t *d_bigEdge, *d_bigPtr, *d_bigEdgeT;
    const uint32_t *d_bigItem;
    const int64_t *d_listEdge, *d_listPtr;
    const uint32_t *d_listItem;
    long long *d_phaseClock;
    /* which part of the schedule this launch covers (multi-GPU overlap, dc_cuda_assemble_exchange):
     * units [unitFirst, unitFirst + unitCount), the persistent grid leaves reserveCtas CTA slots free,
     * the rows of the fallback lists are assembled unless skipBigRows */
    int64_t unitFirst, unitCount;
    int reserveCtas, skipBigRows;
    int launches;          /* out: kernels launched */
};

/* launcher return codes (dc_cuda_last_error() explains them) */
enum {
    DC_LAUNCH_OK = 0, DC_LAUNCH_NO_PLAN = -6, DC_LAUNCH_BAD_MODE = -7, DC_LAUNCH_PARAMS = -8,
    DC_LAUNCH_CUDA = -10, DC_LAUNCH_SMEM = -11, DC_LAUNCH_NOT_SYMMETRIC = -13
};

namespace dcdev {

template <class Op, int THREADS, int MINB, bool SYM>
inline int launch_units(dc_launch_args *a, double *d_values, const dc_params_t &P, cudaStream_t s)
{
    UnitLayout L = UnitSmem<Op>::layout(a->maxSlots, a->maxHalo, a->maxNodes, a->maxTasks, a->maxItems, a->maxTgt,
                                        THREADS);
    if (L.total > 227 * 1024) return DC_LAUNCH_SMEM;
    auto kern = gather_units<Op, THREADS, MINB, SYM>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return DC_LAUNCH_CUDA;
    if (a->persistent) {
        /* double-buffer the schedule (tasks, targets, items) if that costs no resident CTA */
        const UnitLayout L2 = UnitSmem<Op>::layout(a->maxSlots, a->maxHalo, a->maxNodes, a->maxTasks, a->maxItems,
                                                   a->maxTgt, THREADS, true);
        int one = 0, two = 0;
        if (L2.total <= 227 * 1024 &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&one, kern, THREADS, (size_t)L.total) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&two, kern, THREADS, (size_t)L2.total) == cudaSuccess &&
            two == one && two > 0)
            L = L2;
    }
    /* persistent grid: SMs x resident CTAs (the kernel strides over the units), or one CTA per
     * unit when a->persistent == 0 */
    int64_t grid = a->unitCount;
    if (a->persistent) {
        int perSm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, THREADS, (size_t)L.total) != cudaSuccess || perSm < 1)
            return DC_LAUNCH_CUDA;
        int64_t resident = (int64_t)perSm * a->smCount - a->reserveCtas;
        if (resident < 1) resident = 1;
        if (resident < grid) grid = resident;
    }
#ifdef DC_KO
    {
        const char *ko = getenv("DC_KO");
        const int bits = ko ? atoi(ko) : 0;
        cudaMemcpyToSymbolAsync(c_ko, &bits, sizeof(int), 0, cudaMemcpyHostToDevice, s);
    }
#endif
    kern<<<(unsigned)grid, THREADS, (size_t)L.total, s>>>(a->d_units + a->unitFirst, (int)a->unitCount, a->d_tasks, a->d_items,
                                                           a->d_slotNodes, a->d_haloNodes, a->d_diagRel,
                                                           reinterpret_cast<const int2 *>(a->d_tgt), a->d_coord,
                                                           d_values, L, P, a->d_phaseClock);
    return DC_LAUNCH_OK;
}

/* one assembly with operator Op: the body behind dc_cuda_assemble */
template <class Op>
inline int launch_operator(dc_launch_args *a, const double *h_params, int nbParams, double *d_values,
                           cudaStream_t s)
{
    constexpr int DD = Op::D * Op::D;
    a->launches = 0;
    if (nbParams < Op::NPARAM || nbParams > DC_MAX_PARAMS) return DC_LAUNCH_PARAMS;
    /* the coefficients are a kernel argument of every launch below (no device-global state: two
     * assemblies on different streams of one device do not interfere) */
    dc_params_t P;
    for (int i = 0; i < DC_MAX_PARAMS; i++) P.v[i] = i < nbParams ? h_params[i] : 0.0;
    if (a->mode == DC_MODE_DETERMINISTIC) {
        if (!a->d_units && a->nbBigEdges == 0) return DC_LAUNCH_NO_PLAN;
        if (a->unitCount > 0) {
            if (a->symmetric && !Op::SYMMETRIC) return DC_LAUNCH_NOT_SYMMETRIC;
            int rc;
            if (a->symmetric)
                rc = (a->threads == 128)   ? launch_units<Op, 128, 4, true>(a, d_values, P, s)
                     : (a->threads == 512) ? launch_units<Op, 512, 1, true>(a, d_values, P, s)
                                           : launch_units<Op, 256, 2, true>(a, d_values, P, s);
            else
                rc = (a->threads == 128) ? launch_units<Op, 128, 4, false>(a, d_values, P, s)
                                         : launch_units<Op, 256, 2, false>(a, d_values, P, s);
            if (rc) return rc;
            a->launches++;
        }
        if (a->nbBigEdges > 0 && !a->skipBigRows) {
            gather_lists<Op><<<(unsigned)((a->nbBigEdges + 127) / 128), 128, 0, s>>>(
                a->d_bigEdge, a->d_bigPtr, a->d_bigItem, a->d_bigEdgeT, a->nbBigEdges, a->d_conn, a->d_coord, d_values, P);
            a->launches++;
        }
    } else if (a->mode == DC_MODE_ATOMIC) {
        reset_values<<<a->smCount * 8, 256, 0, s>>>(d_values, a->nnz * DD);
        scatter_atomic<Op><<<(unsigned)((a->nbElem + 127) / 128), 128, 0, s>>>(a->d_conn, a->d_coord, a->d_row, a->d_col,
                                                                                 a->colBase, a->nbElem, d_values, P);
        a->launches += 2;
    } else if (a->mode == DC_MODE_LIST) {
        if (!a->d_listPtr) return DC_LAUNCH_NO_PLAN;
        gather_lists<Op><<<(unsigned)((a->nbListEdges + 127) / 128), 128, 0, s>>>(
            a->d_listEdge, a->d_listPtr, a->d_listItem, (const int64_t *)nullptr, a->nbListEdges, a->d_conn, a->d_coord,
            d_values, P);
        a->launches++;
    } else {
        return DC_LAUNCH_BAD_MODE;
    }
    return cudaGetLastError() == cudaSuccess ? DC_LAUNCH_OK : DC_LAUNCH_CUDA;
}

/* Convenience base: the user writes ke() -- the whole element matrix, as in a host callback --
 * and gets an operator.  Record = the matrix (16*D*D doubles), contribution = `+=` of a block.
 * SYM must only be true if ke() is symmetric BIT FOR BIT (k[b][a] == transpose(k[a][b])). */
template <class Derived, int D_, bool SYM>
struct KeOperator {
    static constexpr int D = D_;
    static constexpr int DD = D_ * D_;
    static constexpr int NS = 16 * DD;
    static constexpr bool SYMMETRIC = SYM;
    static constexpr int NOPS = DD, NOPS_DIAG = DD;
    __device__ __forceinline__ static void setup(const double (&x)[4][3], const double *params, double *rec)
    {
        Derived::ke(x, params, rec);
    }
    template <typename R>
    __device__ __forceinline__ static void fetch(R rec, uint32_t ab, double (&ops)[DD])
    {
#pragma unroll
        for (int q = 0; q < DD; q++) ops[q] = rec[ab * DD + q];
    }
    template <typename R>
    __device__ __forceinline__ static void fetch_diag(R rec, uint32_t a, double (&ops)[DD]) { fetch(rec, 5u * a, ops); }
    __device__ __forceinline__ static void apply(const double (&ops)[DD], const double *, double (&acc)[DD])
    {
#pragma unroll
        for (int q = 0; q < DD; q++) acc[q] = __dadd_rn(acc[q], ops[q]);
    }
    __device__ __forceinline__ static void apply_diag(const double (&ops)[DD], const double *p, double (&acc)[DD])
    {
        apply(ops, p, acc);
    }
    template <typename R>
    __device__ __forceinline__ static void accumulate(R rec, const double *p, uint32_t ab, double (&acc)[DD])
    {
        double ops[DD];
        fetch(rec, ab, ops);
        apply(ops, p, acc);
    }
    template <typename R>
    __device__ __forceinline__ static void accumulate_diag(R rec, const double *p, uint32_t a, double (&acc)[DD])
    {
        accumulate(rec, p, 5u * a, acc);
    }
    __device__ __forceinline__ static void mirror(double (&)[DD]) {}
};

}  // namespace dcdev

#define DC_DEVICE_OPERATOR(symbol, Op)                                                                     \
    extern "C" int symbol(dc_launch_args *a, const double *h_params, int nbParams, double *d_values,        \
                          void *stream)                                                                    \
    {                                                                                                      \
        return dcdev::launch_operator<Op>(a, h_params, nbParams, d_values, (cudaStream_t)stream);          \
    }

#endif
