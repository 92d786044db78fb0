/* dc_kernels.cuh -- sm_100a kernels of the assembly path.
 *
 * gather_units   deterministic assembly over work units (see dc_plan.h): one CTA per
 *                unit; own-element connectivity staged in shared memory with a TMA
 *                bulk copy; one record (gradients + scalars) per touched element in
 *                shared memory; then one lane per CSR edge accumulating its D x D
 *                block in registers in ascending element order; values leave the
 *                SM once, as full-line coalesced stores -- the reset is fused.
 * gather_lists   fallback / generic: one thread per edge over explicit lists.
 * scatter_atomic "fast mode" of the north-star: reset pass + one thread per element
 *                + red.global.add.f64 into the CSR.
 * reset_values, permute_rows, renumber: the memset / gather kernels that replace
 *                the CSR reset and src/permutations.cc.
 */
#ifndef DC_KERNELS_CUH
#define DC_KERNELS_CUH

#include <stdint.h>
#include "dc_plan.h"
#include "dc_operators.cuh"

namespace dcdev {

/* ---- small PTX helpers -------------------------------------------------- */

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
/* TMA bulk copy global -> shared, completion signalled on an mbarrier (UBLKCP) */
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void red_add_f64(double *addr, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void st_stream_f64(double *addr, double v)
{
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

/* operator coefficients, set with cudaMemcpyToSymbolAsync before a launch */
#define DC_MAX_PARAMS 8
__constant__ double c_params[DC_MAX_PARAMS];

template <class Op>
__device__ __forceinline__ void load_element(const int *__restrict__ conn, const double *__restrict__ coord,
                                             int64_t e, double *rec)
{
    const int4 c = __ldg(reinterpret_cast<const int4 *>(conn) + e);
    const int n[4] = {c.x - 1, c.y - 1, c.z - 1, c.w - 1};
    double x[4][3];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int k = 0; k < 3; k++) x[a][k] = __ldg(coord + (int64_t)n[a] * 3 + k);
    Op::setup(x, c_params, rec);
}

/* ---- deterministic assembly over work units ------------------------------ */

template <class Op>
struct UnitSmem {
    static constexpr int DD = Op::D * Op::D;
    /* record stride in doubles: odd multiple of 2 words keeps 64-bit accesses of
     * different slots spread over the banks */
    static constexpr int RS = (Op::NS % 2 == 0) ? Op::NS + 1 : Op::NS;
};

template <class Op, int THREADS>
__global__ void __launch_bounds__(THREADS) gather_units(const dc_unit_t *__restrict__ units,
                                                        const dc_task_t *__restrict__ tasks,
                                                        const uint16_t *__restrict__ items,
                                                        const int *__restrict__ haloElems,
                                                        const int *__restrict__ diagRel,
                                                        const int *__restrict__ conn,
                                                        const double *__restrict__ coord,
                                                        double *__restrict__ values, int maxSlots)
{
    constexpr int DD = UnitSmem<Op>::DD;
    constexpr int RS = UnitSmem<Op>::RS;
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    /* layout: [0,16) mbarrier | region A: own connectivity (TMA target), later the
     * per-warp output staging | region B: element records */
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    int4 *connS = reinterpret_cast<int4 *>(smem_raw + 16);
    double *stage = reinterpret_cast<double *>(smem_raw + 16);
    const size_t regionA = (size_t)((maxSlots * 16 > NW * 32 * DD * 8) ? maxSlots * 16 : NW * 32 * DD * 8);
    double *recs = reinterpret_cast<double *>(smem_raw + 16 + regionA);

    const dc_unit_t u = units[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nSlots = u.ownElemCount + u.haloCount;

    /* P0: own connectivity rows are contiguous in tree order -> one bulk copy */
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    const uint32_t ownBytes = (uint32_t)u.ownElemCount * 16u;
    if (tid == 0 && ownBytes) {
        mbar_expect_tx(bar, ownBytes);
        tma_bulk_g2s(connS, reinterpret_cast<const int4 *>(conn) + u.ownElemFirst, ownBytes, bar);
    }
    if (ownBytes) mbar_wait(bar, 0);

    /* P1: one record per touched element */
    for (int s = tid; s < nSlots; s += THREADS) {
        int4 c;
        if (s < u.ownElemCount) c = connS[s];
        else c = __ldg(reinterpret_cast<const int4 *>(conn) + __ldg(haloElems + u.haloFirst + (s - u.ownElemCount)));
        const int n[4] = {c.x - 1, c.y - 1, c.z - 1, c.w - 1};
        double x[4][3];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int k = 0; k < 3; k++) x[a][k] = __ldg(coord + (int64_t)n[a] * 3 + k);
        double rec[Op::NS];
        Op::setup(x, c_params, rec);
#pragma unroll
        for (int i = 0; i < Op::NS; i++) recs[s * RS + i] = rec[i];
    }
    __syncthreads();   /* records complete; region A is free for staging from here on */

    /* P2: tasks round-robin over warps */
    double *myStage = stage + warp * 32 * DD;
    double *out = values + u.edgeFirst * DD;
    for (int t = warp; t < u.taskCount; t += NW) {
        const dc_task_t task = tasks[u.taskFirst + t];
        const uint16_t *it = items + u.itemFirst + task.itemRel + lane;
        double acc[DD];
#pragma unroll
        for (int c = 0; c < DD; c++) acc[c] = 0.0;
        uint32_t item = task.iters ? __ldg(it) : DC_ITEM_NULL;
        for (int i = 0; i < task.iters; i++) {
            const uint32_t cur = item;
            if (i + 1 < task.iters) item = __ldg(it + (i + 1) * 32);
            if (cur != DC_ITEM_NULL) {
                double blk[DD];
                Op::block(recs + (cur >> 4) * RS, (cur >> 2) & 3, cur & 3, blk);
#pragma unroll
                for (int c = 0; c < DD; c++) acc[c] = __dadd_rn(acc[c], blk[c]);
            }
        }
        if (task.kind == DC_TASK_EDGES) {
            if (DD == 1) {
                const int e = (int)task.first + lane;
                if (e < u.edgeCount && !((task.mask >> lane) & 1u)) st_stream_f64(out + e, acc[0]);
            } else {
#pragma unroll
                for (int c = 0; c < DD; c++) myStage[lane * DD + c] = acc[c];
                __syncwarp();
                const int remaining = u.edgeCount - (int)task.first;   /* edges of this task that exist */
#pragma unroll
                for (int r = 0; r < DD; r++) {
                    const int idx = r * 32 + lane;
                    const int ent = idx / DD;
                    if (ent < remaining && !((task.mask >> ent) & 1u))
                        st_stream_f64(out + (int64_t)task.first * DD + idx, myStage[idx]);
                }
                __syncwarp();
            }
        } else {
            if ((task.mask >> lane) & 1u) {
                const int d = __ldg(diagRel + u.nodeFirst + (int)task.first + lane);
#pragma unroll
                for (int c = 0; c < DD; c++) st_stream_f64(out + (int64_t)d * DD + c, acc[c]);
            }
        }
    }
}

/* ---- fallback: explicit per-edge lists, everything from global memory ----- */

template <class Op>
__global__ void __launch_bounds__(128) gather_lists(const int64_t *__restrict__ edge,
                                                    const int64_t *__restrict__ ptr,
                                                    const uint32_t *__restrict__ item, int64_t nbEdges,
                                                    const int *__restrict__ conn,
                                                    const double *__restrict__ coord,
                                                    double *__restrict__ values)
{
    constexpr int DD = Op::D * Op::D;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbEdges) return;
    double acc[DD];
#pragma unroll
    for (int c = 0; c < DD; c++) acc[c] = 0.0;
    for (int64_t q = ptr[i]; q < ptr[i + 1]; q++) {
        const uint32_t it = __ldg(item + q);
        double rec[Op::NS], blk[DD];
        load_element<Op>(conn, coord, (int64_t)(it >> 4), rec);
        Op::block(rec, (it >> 2) & 3, it & 3, blk);
#pragma unroll
        for (int c = 0; c < DD; c++) acc[c] = __dadd_rn(acc[c], blk[c]);
    }
    double *dst = values + edge[i] * DD;
#pragma unroll
    for (int c = 0; c < DD; c++) dst[c] = acc[c];
}

/* ---- fast mode: one thread per element, red.global.add.f64 --------------- */

template <class Op>
__global__ void __launch_bounds__(128) scatter_atomic(const int *__restrict__ conn,
                                                      const double *__restrict__ coord,
                                                      const int *__restrict__ row,
                                                      const int *__restrict__ col, int colBase,
                                                      int64_t nbElem, double *__restrict__ values)
{
    constexpr int DD = Op::D * Op::D;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nbElem) return;
    const int4 c = __ldg(reinterpret_cast<const int4 *>(conn) + e);
    const int n[4] = {c.x - 1, c.y - 1, c.z - 1, c.w - 1};
    double rec[Op::NS];
    load_element<Op>(conn, coord, e, rec);
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int r0 = __ldg(row + n[a]), r1 = __ldg(row + n[a] + 1);
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int target = n[b] + colBase;
            int k = r0;
            while (k < r1 && __ldg(col + k) != target) k++;
            if (k == r1) continue;
            double blk[DD];
            Op::block(rec, a, b, blk);
            double *dst = values + (int64_t)k * DD;
#pragma unroll
            for (int q = 0; q < DD; q++) red_add_f64(dst + q, blk[q]);
        }
    }
}

/* ---- reset / permutation kernels ------------------------------------------ */

/* zero n doubles starting at p: 16-byte stores in the aligned body */
__global__ void __launch_bounds__(256) reset_values(double *__restrict__ p, int64_t n)
{
    const int64_t head = ((reinterpret_cast<uintptr_t>(p) & 15) && n > 0) ? 1 : 0;
    const int64_t pairs = (n - head) / 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double2 *body = reinterpret_cast<double2 *>(p + head);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride)
        body[i] = make_double2(0.0, 0.0);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (head) p[0] = 0.0;
        if (head + 2 * pairs < n) p[n - 1] = 0.0;
    }
}

/* out[perm[i]][:] = in[i][:]; consecutive threads walk consecutive words of `in` */
template <typename T>
__global__ void __launch_bounds__(256) permute_rows(const T *__restrict__ in, T *__restrict__ out,
                                                    const int *__restrict__ perm, int64_t nbItem, int dimItem)
{
    const int64_t total = nbItem * dimItem;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
        const int64_t i = w / dimItem;
        const int j = (int)(w - i * dimItem);
        out[(int64_t)__ldg(perm + i) * dimItem + j] = in[w];
    }
}

/* 16-byte rows (the tetrahedron connectivity): one int4 per thread */
__global__ void __launch_bounds__(256) permute_rows_int4(const int4 *__restrict__ in, int4 *__restrict__ out,
                                                         const int *__restrict__ perm, int64_t nbItem)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbItem; i += stride)
        out[__ldg(perm + i)] = in[i];
}

__global__ void __launch_bounds__(256) renumber(int *__restrict__ tab, const int *__restrict__ perm, int64_t size,
                                                int base)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += stride)
        tab[i] = __ldg(perm + (tab[i] - base)) + base;
}

}  // namespace dcdev
#endif
