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

/* shared-memory geometry of a unit launch, computed on the host from the plan */
struct UnitLayout {
    int offStage;    /* region A: own connectivity (TMA target), later per-warp output staging */
    int offHalo;     /* halo element ids                                                     */
    int offTasks;    /* dc_task_t records                                                    */
    int offItems;    /* uint16 items                                                         */
    int offRecs;     /* element records                                                      */
    int total;
};

template <class Op>
struct UnitSmem {
    static constexpr int DD = Op::D * Op::D;
    /* record stride in doubles: an odd number of 8-byte words keeps 64-bit accesses of
     * different slots spread over the banks */
    static constexpr int RS = (Op::NS % 2 == 0) ? Op::NS + 1 : Op::NS;
    static UnitLayout layout(int maxSlots, int maxHalo, int maxTasks, int maxItems, int threads)
    {
        auto up = [](int v) { return (v + 15) & ~15; };
        UnitLayout L;
        const int a = maxSlots * 16, b = (threads / 32) * 32 * DD * 8;
        L.offStage = 32;
        L.offHalo = L.offStage + up(a > b ? a : b);
        L.offTasks = L.offHalo + up(maxHalo * 4);
        L.offItems = L.offTasks + up(maxTasks * 16);
        L.offRecs = L.offItems + up(maxItems * 2);
        L.total = L.offRecs + maxSlots * RS * 8;
        return L;
    }
};

template <class Op, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) gather_units(const dc_unit_t *__restrict__ units,
                                                              const dc_task_t *__restrict__ tasks,
                                                              const uint16_t *__restrict__ items,
                                                              const int *__restrict__ haloElems,
                                                              const int *__restrict__ diagRel,
                                                              const int *__restrict__ conn,
                                                              const double *__restrict__ coord,
                                                              double *__restrict__ values, const UnitLayout L)
{
    constexpr int DD = UnitSmem<Op>::DD;
    constexpr int RS = UnitSmem<Op>::RS;
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);          /* bar[0]: mesh data, bar[1]: schedule */
    int4 *connS = reinterpret_cast<int4 *>(smem_raw + L.offStage);
    double *stage = reinterpret_cast<double *>(smem_raw + L.offStage);
    int *haloS = reinterpret_cast<int *>(smem_raw + L.offHalo);
    dc_task_t *taskS = reinterpret_cast<dc_task_t *>(smem_raw + L.offTasks);
    uint16_t *itemS = reinterpret_cast<uint16_t *>(smem_raw + L.offItems);
    double *recs = reinterpret_cast<double *>(smem_raw + L.offRecs);

    const dc_unit_t u = units[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nSlots = u.ownElemCount + u.haloCount;

    /* P0: everything the unit reads that is contiguous in HBM comes in as TMA bulk copies:
     * own connectivity rows (tree order makes them one range), halo ids, tasks, items */
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t ownBytes = (uint32_t)u.ownElemCount * 16u;
        const uint32_t haloBytes = (uint32_t)((u.haloCount + 3) & ~3) * 4u;
        mbar_expect_tx(&bar[0], ownBytes + haloBytes);
        if (ownBytes) tma_bulk_g2s(connS, reinterpret_cast<const int4 *>(conn) + u.ownElemFirst, ownBytes, &bar[0]);
        if (haloBytes) tma_bulk_g2s(haloS, haloElems + u.haloFirst, haloBytes, &bar[0]);
        const uint32_t taskBytes = (uint32_t)u.taskCount * 16u, itemBytes = (uint32_t)u.itemCount * 2u;
        mbar_expect_tx(&bar[1], taskBytes + itemBytes);
        if (taskBytes) tma_bulk_g2s(taskS, tasks + u.taskFirst, taskBytes, &bar[1]);
        if (itemBytes) tma_bulk_g2s(itemS, items + u.itemFirst, itemBytes, &bar[1]);
    }
    mbar_wait(&bar[0], 0);

    /* P1: one record per touched element; two elements per thread in flight so that the
     * (L2-latency) coordinate gathers of one overlap the arithmetic of the other */
    for (int s0 = tid; s0 < nSlots; s0 += 2 * THREADS) {
        const int s1 = s0 + THREADS;
        const bool two = s1 < nSlots;
        int4 c0, c1;
        c0 = (s0 < u.ownElemCount) ? connS[s0]
                                   : __ldg(reinterpret_cast<const int4 *>(conn) + haloS[s0 - u.ownElemCount]);
        c1 = c0;
        if (two)
            c1 = (s1 < u.ownElemCount) ? connS[s1]
                                       : __ldg(reinterpret_cast<const int4 *>(conn) + haloS[s1 - u.ownElemCount]);
        const int n0[4] = {c0.x - 1, c0.y - 1, c0.z - 1, c0.w - 1};
        const int n1[4] = {c1.x - 1, c1.y - 1, c1.z - 1, c1.w - 1};
        double x0[4][3], x1[4][3];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                x0[a][k] = __ldg(coord + (int64_t)n0[a] * 3 + k);
                x1[a][k] = __ldg(coord + (int64_t)n1[a] * 3 + k);
            }
        double rec[Op::NS];
        Op::setup(x0, c_params, rec);
#pragma unroll
        for (int i = 0; i < Op::NS; i++) recs[s0 * RS + i] = rec[i];
        if (two) {
            Op::setup(x1, c_params, rec);
#pragma unroll
            for (int i = 0; i < Op::NS; i++) recs[s1 * RS + i] = rec[i];
        }
    }
    mbar_wait(&bar[1], 0);
    __syncthreads();   /* records complete; region A is free for staging from here on */

    /* P2: tasks round-robin over warps; a lane owns one CSR edge and sums its
     * contributions in ascending element order (the order of the item rows) */
    double *myStage = stage + warp * 32 * DD;
    double *out = values + u.edgeFirst * DD;
    for (int t = warp; t < u.taskCount; t += NW) {
        const dc_task_t task = taskS[t];
        const uint16_t *it = itemS + task.itemRel + lane;
        double acc[DD];
#pragma unroll
        for (int c = 0; c < DD; c++) acc[c] = 0.0;
        /* two contributions in flight per lane: the blocks are independent, only the
         * final adds into acc are ordered (ascending element = item row order) */
        int i = 0;
        for (; i + 1 < task.iters; i += 2) {
            const uint32_t cur0 = it[i * 32], cur1 = it[(i + 1) * 32];
            double blk0[DD], blk1[DD];
            if (cur0 != DC_ITEM_NULL) Op::block(recs + (cur0 >> 4) * RS, c_params, cur0 & 15u, blk0);
            if (cur1 != DC_ITEM_NULL) Op::block(recs + (cur1 >> 4) * RS, c_params, cur1 & 15u, blk1);
            if (cur0 != DC_ITEM_NULL) {
#pragma unroll
                for (int c = 0; c < DD; c++) acc[c] = __dadd_rn(acc[c], blk0[c]);
            }
            if (cur1 != DC_ITEM_NULL) {
#pragma unroll
                for (int c = 0; c < DD; c++) acc[c] = __dadd_rn(acc[c], blk1[c]);
            }
        }
        if (i < task.iters) {
            const uint32_t cur = it[i * 32];
            if (cur != DC_ITEM_NULL) {
                double blk[DD];
                Op::block(recs + (cur >> 4) * RS, c_params, cur & 15u, blk);
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
        Op::block(rec, c_params, it & 15u, blk);
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
            Op::block(rec, c_params, (uint32_t)(4 * a + b), blk);
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

/* ---- interface exchange (multi-GPU) ------------------------------------------ */

/* buf[k][:] = values[edge[k]][:]  (perEdge doubles per edge) */
__global__ void __launch_bounds__(256) pack_edges(const double *__restrict__ values, const int64_t *__restrict__ edge,
                                                  int64_t nbEdges, int perEdge, double *__restrict__ buf)
{
    const int64_t total = nbEdges * perEdge, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
        const int64_t k = w / perEdge;
        buf[w] = values[__ldg(edge + k) * perEdge + (w - k * perEdge)];
    }
}

/* values[edge[k]][:] += buf[k][:]; buf may be a peer GPU's buffer mapped over NVLink */
__global__ void __launch_bounds__(256) unpack_add_edges(double *__restrict__ values, const int64_t *__restrict__ edge,
                                                        int64_t nbEdges, int perEdge, const double *__restrict__ buf)
{
    const int64_t total = nbEdges * perEdge, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
        const int64_t k = w / perEdge;
        double *dst = values + __ldg(edge + k) * perEdge + (w - k * perEdge);
        *dst = __dadd_rn(*dst, buf[w]);
    }
}

}  // namespace dcdev
#endif
