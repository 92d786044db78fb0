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
/* non-blocking: has the phase with this parity completed? */
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
/* TMA bulk copy global -> shared, completion signalled on an mbarrier (UBLKCP) */
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
/* asynchronous 8-byte copy global -> shared (LDGSTS): no register holds the value in flight */
__device__ __forceinline__ void cp_async_8(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}
__device__ __forceinline__ void red_add_f64(double *addr, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void st_stream_f64(double *addr, double v)
{
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

/* operator coefficients travel with every launch as a kernel argument (by value): two
 * assemblies with different coefficients on different streams cannot see each other's */
#define DC_MAX_PARAMS 8
struct dc_params_t { double v[DC_MAX_PARAMS]; };

/* the operator's coefficients in registers for the whole kernel: without the empty asm the
 * shuffle (whose result ptxas cannot rematerialise) the constant bank is re-read inside the
 * accumulation loops, one uniform load per trip */
template <class Op>
__device__ __forceinline__ void load_params(const dc_params_t &P, double (&prm)[DC_MAX_PARAMS])
{
#pragma unroll
    for (int i = 0; i < DC_MAX_PARAMS; i++) {
        prm[i] = 0.0;
        if (i < Op::NPARAM) {
            const long long b = __double_as_longlong(P.v[i]);
            const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)b, 0), hi = __shfl_sync(0xffffffffu, (unsigned)(b >> 32), 0);
            prm[i] = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
        }
    }
}

/* experiment builds only (make variant DEFS=-DDC_KO): knock-out bits from the environment variable
 * DC_KO switch phases of the unit kernel off (wrong results) to measure what each one costs:
 * 1 transposed stores, 2 direct stores, 4 diagonal tasks, 8 record build, 16 accumulation loops,
 * 32 upper-edge tasks */
#ifdef DC_KO
__constant__ int c_ko;
#define DC_KO_ON(bit) (c_ko & (bit))
#else
#define DC_KO_ON(bit) false
#endif

template <class Op>
__device__ __forceinline__ void load_element(const int *__restrict__ conn, const double *__restrict__ coord,
                                             int64_t e, const double *prm, double *rec)
{
    const int4 c = __ldg(reinterpret_cast<const int4 *>(conn) + e);
    const int n[4] = {c.x - 1, c.y - 1, c.z - 1, c.w - 1};
    double x[4][3];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int k = 0; k < 3; k++) x[a][k] = __ldg(coord + (int64_t)n[a] * 3 + k);
    Op::setup(x, prm, rec);
}

/* ---- deterministic assembly over work units ------------------------------ */

/* shared-memory geometry of a unit launch, computed on the host from the plan */
struct UnitLayout {
    int offStage;    /* per-warp output staging                                              */
    int offSlotN;    /* per slot: 4 uint16 indices into the unit's node table (TMA target)    */
    int offHnode;    /* halo node ids (TMA target)                                           */
    int offCoord;    /* coordinates of the node table: own interval (TMA target) + halo nodes */
    int coordStride; /* halo node ids + coordinates are double-buffered, second copy this far away */
    int offTasks;    /* dc_task_t records                                                    */
    int offTgt;      /* symmetric schedule: (edge, transposed edge) pairs                    */
    int offItems;    /* uint16 items                                                         */
    int offRecs;     /* element records                                                      */
    int schedStride; /* > 0: tasks/targets/items are double-buffered, second copy this far away */
    int total;
};

template <class Op>
struct UnitSmem {
    static constexpr int DD = Op::D * Op::D;
    /* record stride in doubles: an odd number of 8-byte words keeps 64-bit accesses of
     * different slots spread over the banks */
    static constexpr int RS = (Op::NS % 2 == 0) ? Op::NS + 1 : Op::NS;
    /* per-warp staging: 32 blocks */
    static constexpr int STAGE = 32 * DD * 8;
    static UnitLayout layout(int maxSlots, int maxHalo, int maxNodes, int maxTasks, int maxItems, int maxTgt,
                             int threads, bool doubleSched = false)
    {
        auto up = [](int v) { return (v + 15) & ~15; };
        UnitLayout L;
        L.offStage = 48 + 4 * (int)sizeof(dc_unit_t);      /* barriers, counters, four unit descriptors */
        L.offSlotN = L.offStage + up((threads / 32) * STAGE);
        L.offHnode = L.offSlotN + up(((maxSlots + 1) & ~1) * 8);
        L.offCoord = L.offHnode + up(maxHalo * 4);
        L.coordStride = up(maxHalo * 4) + up((maxNodes + maxHalo) * 24);
        L.offTasks = L.offHnode + 2 * L.coordStride;
        L.offTgt = L.offTasks + up(maxTasks * 16);
        L.offItems = L.offTgt + up(maxTgt * 8);
        L.offRecs = L.offItems + up(maxItems * 2);
        L.schedStride = 0;
        if (doubleSched) {
            L.schedStride = L.offRecs - L.offTasks;
            L.offRecs += L.schedStride;
        }
        L.total = L.offRecs + (maxSlots + 1) * RS * 8;      /* + the all-zero record padding items point to */
        return L;
    }
};

/* store the 32 staged blocks of a warp: idx runs over the 32*DD doubles so that consecutive
 * lanes write consecutive addresses inside a block (and across blocks of consecutive edges).
 * myTarget = the edge this lane's block goes to (-1: none); the target of the block a lane
 * stores a piece of comes from its owner by shuffle, not through shared memory. */
template <int DD>
__device__ __forceinline__ void flush_stage(double *__restrict__ values, const double *stage, int myTarget,
                                            int lane)
{
#pragma unroll
    for (int r = 0; r < DD; r++) {
        const int idx = r * 32 + lane;
        const int ent = idx / DD;
        const int t = __shfl_sync(0xffffffffu, myTarget, ent);
        if (t >= 0) st_stream_f64(values + (int64_t)t * DD + (idx - ent * DD), stage[idx]);
    }
}

/* depth of the software pipeline of the accumulation loops: operands are fetched this many
 * contributions ahead of their use, items twice as far.  Measured at 160^3 (profiles/README.md,
 * round 2): 1/1 and 2/1 are equal, deeper rings lose (more registers live across the loop, the
 * pipe is throughput-bound, not latency-bound) */
#ifndef DC_PIPE_DIAG
#define DC_PIPE_DIAG 2
#endif
#ifndef DC_PIPE_UPPER
#define DC_PIPE_UPPER 1
#endif

/* The contributions of one lane, in the order of its item column it[0], it[32], ... (ascending
 * element index = the reference's order).  A ring of P operand sets: a contribution is applied,
 * then its registers are refilled with the operands of the contribution P places further on (its
 * item was loaded another P places earlier), so P - 1 applications hide a shared-memory round
 * trip.  iters is the same for the whole warp.  The loop body has no test: T full trips of P
 * contributions, then the last P .. 2P-1 contributions with the remaining fetches behind P - 1
 * uniform tests (per task, not per contribution); item loads past the end are clamped (one
 * wavefront each, at most P per task). */
template <class Op, int P, bool DIAG>
__device__ __forceinline__ void accumulate_lane(const uint16_t *it, const int iters, const double *recs,
                                                const double *prm, double (&acc)[Op::D * Op::D])
{
    constexpr int RS = (Op::NS % 2 == 0) ? Op::NS + 1 : Op::NS;
    constexpr int NO = DIAG ? Op::NOPS_DIAG : Op::NOPS;
    auto fetch = [&](double (&o)[NO], uint32_t item) {
        if constexpr (DIAG) Op::fetch_diag(recs + (item >> 4) * RS, item & 3u, o);
        else Op::fetch(recs + (item >> 4) * RS, item & 15u, o);
    };
    auto apply = [&](const double (&o)[NO]) {
        if constexpr (DIAG) Op::apply_diag(o, prm, acc);
        else Op::apply(o, prm, acc);
    };
    if (iters < P) {                       /* short lists (boundary rows): plain loop */
        for (int i = 0; i < iters; i++) {
            double o[NO];
            fetch(o, it[i * 32]);
            apply(o);
        }
        return;
    }
    double ops[P][NO];
    uint32_t itm[P];
    const int last = iters - 1;
#pragma unroll
    for (int j = 0; j < P; j++) itm[j] = it[j * 32];
#pragma unroll
    for (int j = 0; j < P; j++) {
        fetch(ops[j], itm[j]);
        itm[j] = it[min(j + P, last) * 32];
    }
    const int trips = (iters - P) / P;
    int c = 0;
    for (int t = 0; t < trips; t++, c += P) {
#pragma unroll
        for (int j = 0; j < P; j++) {
            apply(ops[j]);
            fetch(ops[j], itm[j]);
            itm[j] = it[min(c + j + 2 * P, last) * 32];
        }
    }
    const int rest = iters - c;            /* P <= rest < 2P: rest - P operand sets still to fetch */
#pragma unroll
    for (int j = 0; j < P; j++) {
        apply(ops[j]);
        if (j + P < rest) fetch(ops[j], itm[j]);
    }
#pragma unroll
    for (int j = 0; j + 1 < P; j++)
        if (j + P < rest) apply(ops[j]);
}

/* One task (P2): a lane owns one CSR edge and adds its contributions in ascending element order
 * (the order of the item rows); the block -- and, in the symmetric schedule, its transpose --
 * leaves through the warp's staging area as full-line streaming stores. */
template <class Op, bool SYM>
__device__ __forceinline__ void do_task(const dc_unit_t &u, const dc_task_t task, const uint16_t *itemS,
                                        const int2 *tgtS, const double *recs, double *myStage,
                                        const int *__restrict__ diagRel, double *__restrict__ values,
                                        const double *prm, int lane)
{
    constexpr int DD = UnitSmem<Op>::DD;
    const uint16_t *it = itemS + task.itemRel + lane;
    double acc[DD];
#pragma unroll
    for (int c = 0; c < DD; c++) acc[c] = 0.0;
    if (SYM && task.kind == DC_TASK_OWNDIAG) {
        /* diagonal edges only: one gradient per contribution, symmetric block */
        if (DC_KO_ON(4)) return;
        const int iters = DC_KO_ON(16) ? 0 : task.iters;
        /* lanes without an edge (the last task of a class is partial) stay out of the loop, so a
         * half-warp without edges costs no shared-memory wavefronts */
        if (iters > 0 && ((task.mask >> lane) & 1u)) accumulate_lane<Op, DC_PIPE_DIAG, true>(it, iters, recs, prm, acc);
        Op::mirror(acc);
        const bool valid = (task.mask >> lane) & 1u;
        int t0 = -1;
        if (valid) t0 = tgtS[(int)task.first + lane].x;
        if (DC_KO_ON(2)) t0 = -1;
        if (DD == 1) {
            if (valid && t0 >= 0) st_stream_f64(values + t0, acc[0]);
        } else {
#pragma unroll
            for (int c = 0; c < DD; c++) myStage[lane * DD + c] = acc[c];
            __syncwarp();
            flush_stage<DD>(values, myStage, t0, lane);
            __syncwarp();
        }
        return;
    }
    if (SYM && DC_KO_ON(32)) return;
    {
        const int iters = DC_KO_ON(16) ? 0 : task.iters;
        const bool holdsEdge = task.kind != DC_TASK_UPPER || ((task.mask >> lane) & 1u);
        if (iters > 0 && holdsEdge) accumulate_lane<Op, DC_PIPE_UPPER, false>(it, iters, recs, prm, acc);
    }
    if (task.kind == DC_TASK_DIAG) {
        if ((task.mask >> lane) & 1u) {
            const int d = __ldg(diagRel + u.nodeFirst + (int)task.first + lane);
#pragma unroll
            for (int c = 0; c < DD; c++) st_stream_f64(values + (u.edgeFirst + d) * DD + c, acc[c]);
        }
    } else if (SYM) {
        /* owned edge: block to (row,col); upper edges also store the transpose at (col,row) */
        const bool valid = (task.mask >> lane) & 1u;
        int2 tg = make_int2(-1, -1);
        if (valid) tg = tgtS[(int)task.first + lane];
        if (DC_KO_ON(2)) tg.x = -1;
        if (DC_KO_ON(1)) tg.y = -1;
        if (DD == 1) {
            if (valid) {
                if (tg.x >= 0) st_stream_f64(values + tg.x, acc[0]);
                if (tg.y >= 0) st_stream_f64(values + tg.y, acc[0]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < DD; c++) myStage[lane * DD + c] = acc[c];
            __syncwarp();
            {
                /* the block and its transpose leave from ONE pass over the staged words: a lane reads
                 * the word it would store for the block and also stores it at the transposed position
                 * of the transposed edge's block (the 32 stores of an instruction cover the same lines
                 * either way) */
                const int tgx = DC_KO_ON(2) ? -1 : tg.x, tgy = DC_KO_ON(1) ? -1 : tg.y;
#pragma unroll
                for (int r = 0; r < DD; r++) {
                    const int idx = r * 32 + lane;
                    const int ent = idx / DD;
                    const int comp = idx - ent * DD;
                    const int tx = __shfl_sync(0xffffffffu, tgx, ent);
                    const int ty = __shfl_sync(0xffffffffu, tgy, ent);
                    const double v = myStage[idx];
                    if (tx >= 0) st_stream_f64(values + (int64_t)tx * DD + comp, v);
                    if (ty >= 0) st_stream_f64(values + (int64_t)ty * DD + (comp % Op::D) * Op::D + comp / Op::D, v);
                }
            }
            __syncwarp();
        }
    } else {
        /* 32 consecutive edges of the unit's own interval (diagonal lanes skipped) */
        const int e = (int)task.first + lane;
        const bool valid = e < u.edgeCount && !((task.mask >> lane) & 1u);
        if (DD == 1) {
            if (valid) st_stream_f64(values + u.edgeFirst + e, acc[0]);
        } else {
#pragma unroll
            for (int c = 0; c < DD; c++) myStage[lane * DD + c] = acc[c];
            __syncwarp();
            flush_stage<DD>(values, myStage, valid ? (int)(u.edgeFirst + e) : -1, lane);
            __syncwarp();
        }
    }
}

/* The accumulation + output phase of one unit: warps take tasks from a shared counter (tasks are
 * listed longest first). */
template <class Op, int NW, bool SYM>
__device__ __forceinline__ void run_tasks(const dc_unit_t &u, const dc_task_t *taskS, const uint16_t *itemS,
                                          const int2 *tgtS, const double *recs, int *nextTask, double *myStage,
                                          const int *__restrict__ diagRel, double *__restrict__ values,
                                          const double *prm, int warp, int lane)
{
    for (int t = warp; t < u.taskCount;) {
        const dc_task_t task = taskS[t];
        {
            int nxt = 0;
            if (lane == 0) nxt = atomicAdd(nextTask, 1);
            t = __shfl_sync(0xffffffffu, nxt, 0);
        }
        do_task<Op, SYM>(u, task, itemS, tgtS, recs, myStage, diagRel, values, prm, lane);
    }
}

/* One CTA processes units blockIdx.x, blockIdx.x + gridDim.x, ... (a persistent grid of
 * SMs x resident CTAs, or one CTA per unit when the grid equals the unit count).  The fixed
 * latencies of a unit are taken off the path of the unit before it (iteration k = unit k of the CTA):
 *
 *   top of k     request descriptor k+2; bulk copy of halo node ids + own coordinates of unit k+1
 *                into the OTHER coordinate buffer (descriptor k+1 was requested an iteration ago)
 *   records(k)   from the slot -> node table and the coordinate buffer of unit k
 *   sync A       records complete, the slot -> node table is dead: bulk copy of the table of unit
 *                k+1 (and of its schedule when that is double-buffered)
 *   halo(k+1)    every thread starts the asynchronous copy (cp.async) of one halo node's coordinates
 *                of unit k+1 into the other coordinate buffer (the ids landed while the records
 *                were built) ...
 *   tasks(k)     accumulate + store
 *                ... which completes behind the tasks; nodes beyond the CTA size are copied by the
 *                warps that run out of tasks first
 *   sync B       the schedule region is dead
 *
 * so no bulk-copy or gather latency sits between the tasks of one unit and the records of the
 * next.  The warp that starts with the longest task (warp 0: tasks are listed longest first) is
 * the critical path of a unit, so the bulk copies are issued by the last warp. */
template <class Op, int THREADS, int MINB, bool SYM>
__global__ void __launch_bounds__(THREADS, MINB) gather_units(const dc_unit_t *__restrict__ units, int nbUnits,
                                                              const dc_task_t *__restrict__ tasks,
                                                              const uint16_t *__restrict__ items,
                                                              const uint16_t *__restrict__ slotNodes,
                                                              const int *__restrict__ haloNodes,
                                                              const int *__restrict__ diagRel,
                                                              const int2 *__restrict__ tgtPairs,
                                                              const double *__restrict__ coord,
                                                              double *__restrict__ values, const UnitLayout L,
                                                              const dc_params_t P,
                                                              long long *__restrict__ phaseClock)
{
    double prm[DC_MAX_PARAMS];
    load_params<Op>(P, prm);
    /* phaseClock (normally null): 24 clock64() stamps per unit: 8 written by thread 0, then the
     * end of the tasks of every warp (tools/phase_times.py) */
#define DC_STAMP(unit, i) do { if (phaseClock && tid == 0) phaseClock[(int64_t)(unit) * 24 + (i)] = clock64(); } while (0)
    constexpr int DD = UnitSmem<Op>::DD;
    constexpr int RS = UnitSmem<Op>::RS;
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);   /* [0] slot table, [1] schedule, [2] descriptor, [3] coordinates */
    int *nextTask = reinterpret_cast<int *>(smem_raw + 32);           /* dynamic task counter */
    int *nextHalo = reinterpret_cast<int *>(smem_raw + 36);           /* halo nodes of the next unit handed out so far */
    constexpr int ISSUER = (NW - 1) * 32;                             /* thread that issues the bulk copies */
    dc_unit_t *descS = reinterpret_cast<dc_unit_t *>(smem_raw + 48);  /* descriptor of unit k of this CTA at k & 3 */
    uint2 *slotN = reinterpret_cast<uint2 *>(smem_raw + L.offSlotN);
    double *recs = reinterpret_cast<double *>(smem_raw + L.offRecs);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *myStage = reinterpret_cast<double *>(smem_raw + L.offStage + warp * UnitSmem<Op>::STAGE);
    auto hnode_of = [&](int k) { return reinterpret_cast<int *>(smem_raw + L.offHnode + (k & 1) * L.coordStride); };
    auto coord_of = [&](int k) { return reinterpret_cast<double *>(smem_raw + L.offCoord + (k & 1) * L.coordStride); };

    /* bulk copies of a unit's inputs (one thread).  The node table starts at the even node at or
     * below nodeFirst and spans an even number of nodes: 48-byte granules. */
    auto issue_coords = [&](const dc_unit_t &d, int k) {
        const int alignedFirst = d.nodeFirst & ~1;
        const int ownSpan = ((d.nodeFirst + d.nodeCount + 1) & ~1) - alignedFirst;
        const uint32_t hnodeBytes = (uint32_t)d.hnodeCount * 4u, coordBytes = (uint32_t)ownSpan * 24u;
        mbar_expect_tx(&bar[3], hnodeBytes + coordBytes);
        if (hnodeBytes) tma_bulk_g2s(hnode_of(k), haloNodes + d.hnodeFirst, hnodeBytes, &bar[3]);
        if (coordBytes) tma_bulk_g2s(coord_of(k), coord + (int64_t)alignedFirst * 3, coordBytes, &bar[3]);
    };
    auto issue_slots = [&](const dc_unit_t &d) {
        const int nSl = d.ownElemCount + d.haloCount;
        const uint32_t slotBytes = (uint32_t)((nSl + 1) & ~1) * 8u;
        mbar_expect_tx(&bar[0], slotBytes);
        if (slotBytes) tma_bulk_g2s(slotN, slotNodes + d.slotFirst * 4, slotBytes, &bar[0]);
    };
    /* schedule of iteration `iter` (double-buffered when the layout has room for it) */
    auto issue_sched = [&](const dc_unit_t &d, int iter) {
        unsigned char *base = smem_raw + (iter & 1) * L.schedStride;
        const uint32_t taskBytes = (uint32_t)d.taskCount * 16u, itemBytes = (uint32_t)d.itemCount * 2u;
        const uint32_t tgtBytes = SYM ? (uint32_t)((d.tgtCount + 1) & ~1) * 8u : 0u;
        mbar_expect_tx(&bar[1], taskBytes + itemBytes + tgtBytes);
        if (taskBytes) tma_bulk_g2s(base + L.offTasks, tasks + d.taskFirst, taskBytes, &bar[1]);
        if (itemBytes) tma_bulk_g2s(base + L.offItems, items + d.itemFirst, itemBytes, &bar[1]);
        if (tgtBytes) tma_bulk_g2s(base + L.offTgt, tgtPairs + d.tgtFirst, tgtBytes, &bar[1]);
    };

    int unit = blockIdx.x;
    if (unit >= nbUnits) return;
    const int stride = (int)gridDim.x;
    DC_STAMP(unit, 0);
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], 1);
        mbar_init(&bar[3], 1);
        *nextTask = NW;
        *nextHalo = THREADS > 32 ? THREADS - 32 : THREADS;
        descS[0] = units[unit];
        if (unit + stride < nbUnits) descS[1] = units[unit + stride];
    }
    __syncthreads();
    if (tid == 0) {
        issue_slots(descS[0]);
        issue_coords(descS[0], 0);
        issue_sched(descS[0], 0);
    }
    mbar_wait(&bar[3], 0);
    {
        /* coordinates of the first unit's halo nodes: one gather per node (not per element) */
        const dc_unit_t &d = descS[0];
        const int ownSpan = ((d.nodeFirst + d.nodeCount + 1) & ~1) - (d.nodeFirst & ~1);
        const int *hn = hnode_of(0);
        double *cs = coord_of(0);
        for (int i = tid; i < d.hnodeCount; i += THREADS) {
            const double *src = coord + (int64_t)hn[i] * 3;
            double *dst = cs + (ownSpan + i) * 3;
            dst[0] = __ldg(src);
            dst[1] = __ldg(src + 1);
            dst[2] = __ldg(src + 2);
        }
    }
    __syncthreads();

    for (int k = 0;; k++) {
        const dc_unit_t u = descS[k & 3];
        const dc_task_t *taskS = reinterpret_cast<const dc_task_t *>(smem_raw + (k & 1) * L.schedStride + L.offTasks);
        const int2 *tgtS = reinterpret_cast<const int2 *>(smem_raw + (k & 1) * L.schedStride + L.offTgt);
        const uint16_t *itemS = reinterpret_cast<const uint16_t *>(smem_raw + (k & 1) * L.schedStride + L.offItems);
        const double *coordS = coord_of(k);
        const int nSlots = u.ownElemCount + u.haloCount;
        const int next = unit + stride;
        const bool hasNext = next < nbUnits;
        /* descriptor k+1 was requested at the top of iteration k-1 (k = 0: loaded directly); its
         * halo node ids and own coordinates go to the other coordinate buffer now, descriptor k+2
         * is requested for the next iteration.  One thread waits and re-arms bar[2], in this order
         * (phase k of bar[2] = descriptor k+2; the last two iterations arm nothing and wait for
         * nothing that was not armed). */
        if (tid == ISSUER && hasNext) {
            if (k >= 1) mbar_wait(&bar[2], (uint32_t)(k - 1) & 1u);
            if (next + stride < nbUnits) {
                mbar_expect_tx(&bar[2], (uint32_t)sizeof(dc_unit_t));
                tma_bulk_g2s(&descS[(k + 2) & 3], units + next + stride, (uint32_t)sizeof(dc_unit_t), &bar[2]);
            }
            issue_coords(descS[(k + 1) & 3], k + 1);
        }
        DC_STAMP(unit, 1);

        /* P1b: one record per touched element, everything from shared memory; slot nSlots is an
         * all-zero record: padding items point at it and add exact zeros, so the accumulation
         * loops need no test and can be software-pipelined */
        mbar_wait(&bar[0], (uint32_t)k & 1u);          /* this unit's slot -> node table */
        if (tid < RS) recs[nSlots * RS + tid] = 0.0;
        for (int s0 = tid; s0 < (DC_KO_ON(8) ? 0 : nSlots); s0 += 2 * THREADS) {
            /* two elements per thread in flight: the dependent chain of one setup (cross products,
             * reciprocal, square root) overlaps the other's.  A warp in which no thread has a second
             * slot builds one record per thread; in a mixed warp the threads without one repeat
             * their first (a divergent branch there costs more than it saves, measured). */
            const int s1 = s0 + THREADS;
            const bool two = s1 < nSlots;
            const uint2 la = slotN[s0];
            const uint32_t na[4] = {la.x & 0xFFFFu, la.x >> 16, la.y & 0xFFFFu, la.y >> 16};
            double xa[4][3], ra[Op::NS];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int q = 0; q < 3; q++) xa[a][q] = coordS[na[a] * 3 + q];
            if (__any_sync(__activemask(), two)) {
                const uint2 lb = slotN[two ? s1 : s0];
                const uint32_t nb[4] = {lb.x & 0xFFFFu, lb.x >> 16, lb.y & 0xFFFFu, lb.y >> 16};
                double xb[4][3], rb[Op::NS];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int q = 0; q < 3; q++) xb[a][q] = coordS[nb[a] * 3 + q];
                Op::setup(xa, prm, ra);
                Op::setup(xb, prm, rb);
#pragma unroll
                for (int i = 0; i < Op::NS; i++) recs[s0 * RS + i] = ra[i];
                if (two) {
#pragma unroll
                    for (int i = 0; i < Op::NS; i++) recs[s1 * RS + i] = rb[i];
                }
            } else {
                Op::setup(xa, prm, ra);
#pragma unroll
                for (int i = 0; i < Op::NS; i++) recs[s0 * RS + i] = ra[i];
            }
        }
        mbar_wait(&bar[1], k & 1);   /* this unit's schedule (every thread passes before the phase is re-armed) */
        __syncthreads();   /* A: records complete; the slot -> node table is dead */
        DC_STAMP(unit, 2);
        if (tid == ISSUER && hasNext) {
            issue_slots(descS[(k + 1) & 3]);                             /* lands while this unit accumulates */
            if (L.schedStride) issue_sched(descS[(k + 1) & 3], k + 1);   /* into the other buffer */
        }
        /* halo coordinates of the next unit: the node ids have normally landed while the records
         * were built; then every thread starts the asynchronous copy of one node's coordinates
         * into the other coordinate buffer now and it completes behind the tasks.  A thread that
         * finds the bulk copy still in flight does not wait here: it copies its node after the
         * tasks. */
        auto copy_halo_node = [&](int i) {
            const dc_unit_t &nx = descS[(k + 1) & 3];
            const int hSpan = ((nx.nodeFirst + nx.nodeCount + 1) & ~1) - (nx.nodeFirst & ~1);
            const double *src = coord + (int64_t)hnode_of(k + 1)[i] * 3;
            double *dst = coord_of(k + 1) + (hSpan + i) * 3;
            cp_async_8(dst, src);
            cp_async_8(dst + 1, src + 1);
            cp_async_8(dst + 2, src + 2);
        };
        /* warp 0 takes no part in the halo copies: it starts with the longest task (the diagonal
         * chain) and is the last warp at sync B in 90 % of the units (tools/phase_times.py), so
         * every copy it issued or waited for sat on the critical path; the other warps copy
         * nodes tid - 32, tid - 32 + HT, ... */
        constexpr int HT = THREADS > 32 ? THREADS - 32 : THREADS;      /* threads that copy halo nodes */
        const int htid = THREADS > 32 ? tid - 32 : tid;
        const bool copier = htid >= 0;
        const bool early = copier && hasNext && mbar_test(&bar[3], (uint32_t)(k + 1) & 1u);
        if (early && htid < descS[(k + 1) & 3].hnodeCount) copy_halo_node(htid);
        DC_STAMP(unit, 3);

        run_tasks<Op, NW, SYM>(u, taskS, itemS, tgtS, recs, nextTask, myStage, diagRel, values, prm, warp, lane);
        DC_STAMP(unit, 4);
        if (phaseClock && lane == 0 && warp < 16) phaseClock[(int64_t)unit * 24 + 8 + warp] = clock64();   /* end of this warp's tasks */
        if (hasNext && copier) {
            if (!early) mbar_wait(&bar[3], (uint32_t)(k + 1) & 1u);
            const int hCnt = descS[(k + 1) & 3].hnodeCount;
            if (!early && htid < hCnt) copy_halo_node(htid);
            /* the rare unit with more halo nodes than copying threads: chunks of 32, taken by the
             * warps that run out of tasks first */
            if (hCnt > HT) {
                for (;;) {
                    int c = 0;
                    if (lane == 0) c = atomicAdd(nextHalo, 32);
                    c = __shfl_sync(0xffffffffu, c, 0);
                    if (c >= hCnt) break;
                    if (c + lane < hCnt) copy_halo_node(c + lane);
                }
            }
            cp_async_wait_all();
        }
        DC_STAMP(unit, 5);
        __syncthreads();   /* B: accumulation finished everywhere; schedule region is dead */
        DC_STAMP(unit, 6);
        if (!hasNext) break;
        if (tid == 0) {
            *nextTask = NW;
            *nextHalo = HT;
            if (!L.schedStride) issue_sched(descS[(k + 1) & 3], k + 1);
        }
        unit = next;
    }
#undef DC_STAMP
}

/* ---- fallback: explicit per-edge lists, everything from global memory ----- */

template <class Op>
__global__ void __launch_bounds__(128) gather_lists(const int64_t *__restrict__ edge,
                                                    const int64_t *__restrict__ ptr,
                                                    const uint32_t *__restrict__ item,
                                                    const int64_t *__restrict__ edgeT, int64_t nbEdges,
                                                    const int *__restrict__ conn,
                                                    const double *__restrict__ coord,
                                                    double *__restrict__ values, const dc_params_t P)
{
    constexpr int DD = Op::D * Op::D;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbEdges) return;
    double prm[DC_MAX_PARAMS];
    load_params<Op>(P, prm);
    double acc[DD];
#pragma unroll
    for (int c = 0; c < DD; c++) acc[c] = 0.0;
    for (int64_t q = ptr[i]; q < ptr[i + 1]; q++) {
        const uint32_t it = __ldg(item + q);
        double rec[Op::NS];
        load_element<Op>(conn, coord, (int64_t)(it >> 4), prm, rec);
        Op::accumulate(rec, prm, it & 15u, acc);
    }
    double *dst = values + edge[i] * DD;
#pragma unroll
    for (int c = 0; c < DD; c++) dst[c] = acc[c];
    if (edgeT && edgeT[i] >= 0) {          /* symmetric lists: the transposed block goes to (col,row) */
        double *dt = values + edgeT[i] * DD;
#pragma unroll
        for (int c = 0; c < DD; c++) dt[(c % Op::D) * Op::D + c / Op::D] = acc[c];
    }
}

/* ---- fast mode: one thread per element, red.global.add.f64 --------------- */

template <class Op>
__global__ void __launch_bounds__(128) scatter_atomic(const int *__restrict__ conn,
                                                      const double *__restrict__ coord,
                                                      const int *__restrict__ row,
                                                      const int *__restrict__ col, int colBase,
                                                      int64_t nbElem, double *__restrict__ values,
                                                      const dc_params_t P)
{
    constexpr int DD = Op::D * Op::D;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nbElem) return;
    double prm[DC_MAX_PARAMS];
    load_params<Op>(P, prm);
    const int4 c = __ldg(reinterpret_cast<const int4 *>(conn) + e);
    const int n[4] = {c.x - 1, c.y - 1, c.z - 1, c.w - 1};
    double rec[Op::NS];
    load_element<Op>(conn, coord, e, prm, rec);
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
#pragma unroll
            for (int q = 0; q < DD; q++) blk[q] = 0.0;
            Op::accumulate(rec, prm, (uint32_t)(4 * a + b), blk);   /* = K_e[a][b] */
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

/* out[d][:] = in[inv[d]][:] -- the same permutation told by destination (inv = inverse of perm).
 * Measured on B200 (profiles/README.md, round 2): a partial write of a 32-byte sector makes the L2
 * fetch the sector from DRAM first, so the scatter form above moves twice the algorithmic bytes
 * (rows of 16 or 24 bytes never fill a sector on their own).  Told by destination, consecutive
 * threads write consecutive words -- whole sectors, whole lines -- and only the READS are scattered,
 * which costs nothing extra when neighbouring rows are fetched close in time (tree orderings are
 * local).  The library owns both directions of its permutations, so its own callers use this form. */
template <typename T>
__global__ void __launch_bounds__(256) gather_rows(const T *__restrict__ in, T *__restrict__ out,
                                                   const int *__restrict__ inv, int64_t nbItem, int dimItem)
{
    const int64_t total = nbItem * dimItem;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
        const int64_t d = w / dimItem;
        const int j = (int)(w - d * dimItem);
        out[w] = __ldg(in + (int64_t)__ldg(inv + d) * dimItem + j);
    }
}

/* 16-byte rows (the tetrahedron connectivity): two rows in flight per thread */
__global__ void __launch_bounds__(256) gather_rows_int4(const int4 *__restrict__ in, int4 *__restrict__ out,
                                                        const int *__restrict__ inv, int64_t nbItem)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < nbItem; d += 2 * stride) {
        const int64_t d1 = d + stride;
        const int s0 = __ldg(inv + d), s1 = d1 < nbItem ? __ldg(inv + d1) : 0;
        const int4 a = __ldg(in + s0);
        if (d1 < nbItem) {
            const int4 b = __ldg(in + s1);
            out[d1] = b;
        }
        out[d] = a;
    }
}

/* tab[i] = perm[tab[i] - base] + base, four entries per thread (16-byte loads and stores of tab; the
 * four look-ups of perm are independent) with a scalar head / tail */
__global__ void __launch_bounds__(256) renumber(int *__restrict__ tab, const int *__restrict__ perm, int64_t size,
                                                int base)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t head = ((16 - (reinterpret_cast<uintptr_t>(tab) & 15)) & 15) / 4;        /* ints before the first 16-byte boundary */
    const int64_t h = head < size ? head : size;
    const int64_t quads = (size - h) / 4;
    int4 *body = reinterpret_cast<int4 *>(tab + h);
    for (int64_t q = t0; q < quads; q += stride) {
        int4 v = body[q];
        v.x = __ldg(perm + (v.x - base)) + base;
        v.y = __ldg(perm + (v.y - base)) + base;
        v.z = __ldg(perm + (v.z - base)) + base;
        v.w = __ldg(perm + (v.w - base)) + base;
        body[q] = v;
    }
    if (t0 < h) tab[t0] = __ldg(perm + (tab[t0] - base)) + base;
    const int64_t tail = h + 4 * quads + t0;
    if (tail < size) tab[tail] = __ldg(perm + (tab[tail] - base)) + base;
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

/* ---- consumer of the assembled matrix -------------------------- */

/* y = A x for the block-CSR matrix the assembly leaves on the device (D x D row-major block per
 * edge, x and y with D components per node): one thread per (row, component), blocks of the row
 * in ascending edge order, one fma per entry -- a fixed summation order.  The three threads of a
 * row read the 72 contiguous bytes of a block between them; the values are streamed (read once). */
template <int D>
__global__ void __launch_bounds__(256) bsr_spmv(const int *__restrict__ row, const int *__restrict__ col, int colBase,
                                                const double *__restrict__ values, const double *__restrict__ x,
                                                double *__restrict__ y, int64_t nbRows)
{
    const int64_t total = nbRows * D, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t r = t / D;
        const int i = (int)(t - r * D);
        const int k0 = __ldg(row + r), k1 = __ldg(row + r + 1);
        double acc = 0.0;
        for (int k = k0; k < k1; k++) {
            const int64_t c = __ldg(col + k) - colBase;
            const double *v = values + ((int64_t)k * D + i) * D;
#pragma unroll
            for (int j = 0; j < D; j++) acc = __fma_rn(__ldcs(v + j), __ldg(x + c * D + j), acc);
        }
        y[t] = acc;
    }
}

/* Block-Jacobi preconditioner of the assembled matrix (what the reference's user
 * builds next): inv[r] = inverse of the D x D diagonal block of row r, one thread per row -- the diagonal
 * edge is found in the row's columns, the 3 x 3 inverse is the adjugate over the determinant in a fixed
 * operation order (cofactors fma(p, q, -(r * s)), det = fma(a02, c02, fma(a01, c01, a00 * c00))).
 * A row without a diagonal entry, or a singular block, gets zeros and is counted in *bad. */
template <int D>
__global__ void __launch_bounds__(256) block_jacobi(const int *__restrict__ row, const int *__restrict__ col, int colBase,
                                                    const double *__restrict__ values, double *__restrict__ inv,
                                                    int64_t nbRows, int *__restrict__ bad)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nbRows; r += stride) {
        int k = __ldg(row + r);
        const int k1 = __ldg(row + r + 1);
        while (k < k1 && __ldg(col + k) - colBase != r) k++;
        double *o = inv + r * D * D;
        if (k == k1) {
#pragma unroll
            for (int q = 0; q < D * D; q++) o[q] = 0.0;
            atomicAdd(bad, 1);
            continue;
        }
        const double *a = values + (int64_t)k * D * D;
        if (D == 1) {
            const double v = a[0];
            o[0] = v != 0.0 ? __drcp_rn(v) : 0.0;
            if (v == 0.0) atomicAdd(bad, 1);
        } else {
            double m[9], c[9];
#pragma unroll
            for (int q = 0; q < 9; q++) m[q] = __ldcs(a + q);
            c[0] = cross_comp(m[4], m[8], m[5], m[7]);
            c[1] = cross_comp(m[5], m[6], m[3], m[8]);
            c[2] = cross_comp(m[3], m[7], m[4], m[6]);
            c[3] = cross_comp(m[2], m[7], m[1], m[8]);
            c[4] = cross_comp(m[0], m[8], m[2], m[6]);
            c[5] = cross_comp(m[1], m[6], m[0], m[7]);
            c[6] = cross_comp(m[1], m[5], m[2], m[4]);
            c[7] = cross_comp(m[2], m[3], m[0], m[5]);
            c[8] = cross_comp(m[0], m[4], m[1], m[3]);
            const double det = __fma_rn(m[2], c[2], __fma_rn(m[1], c[1], __dmul_rn(m[0], c[0])));
            const double id = det != 0.0 ? __drcp_rn(det) : 0.0;
            if (det == 0.0) atomicAdd(bad, 1);
            /* inverse = adjugate / det; adjugate = transposed cofactor matrix */
            o[0] = __dmul_rn(c[0], id); o[1] = __dmul_rn(c[3], id); o[2] = __dmul_rn(c[6], id);
            o[3] = __dmul_rn(c[1], id); o[4] = __dmul_rn(c[4], id); o[5] = __dmul_rn(c[7], id);
            o[6] = __dmul_rn(c[2], id); o[7] = __dmul_rn(c[5], id); o[8] = __dmul_rn(c[8], id);
        }
    }
}

}  // namespace dcdev
#endif
