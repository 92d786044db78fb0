/* dc_csr.cuh -- CSR pattern (nodeToNodeRow / nodeToNodeColumn) and node -> element lists built on
 * the device from the renumbered connectivity.
 *
 * In the reference this is user-side work done between DC_renumber_int_array and
 * DC_finalize_tree; the library only ships the node -> element
 * helper DC_create_nodeToElem (src/coloring.cc:79-112: (element, node) couples sorted by node).
 * At 10^8 elements the host build dominates the set-up, so the same arrays are produced here:
 *
 *   csr_count_nodes     deg[node]++ for the 4 nodes of every element
 *   scan_*              exclusive prefix sum (block sums, scan of the sums, apply) -> n2ePtr
 *   csr_fill_n2e        n2eVal[n2ePtr[node] + cursor[node]++] = element   (any order)
 *   csr_rows<FILL>      one warp per node: the candidates are the 4 nodes of every incident
 *                       element; a candidate is kept if it is the first occurrence of its value
 *                       and placed at the rank of its value among the kept ones, so a row comes
 *                       out ascending and without duplicates whatever the order of n2eVal --
 *                       the arrays equal the host builder's (dcx_build_csr) bit for bit.
 *                       FILL = false counts (-> row pointers), FILL = true writes the columns.
 */
#ifndef DC_CSR_CUH
#define DC_CSR_CUH
#include <stdint.h>

namespace dcdev {

__global__ void __launch_bounds__(256) csr_count_nodes(const int *__restrict__ conn, int64_t nbEntries, int nbNodes,
                                                       int *__restrict__ deg, int *__restrict__ bad)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbEntries; i += stride) {
        const int n = __ldg(conn + i) - 1;
        if (n < 0 || n >= nbNodes) { *bad = 1; continue; }
        atomicAdd(deg + n, 1);
    }
}

/* exclusive scan of int32 counts into int64 offsets, SCAN_TILE items per block */
constexpr int SCAN_TILE = 2048;

__global__ void __launch_bounds__(256) scan_block_sums(const int *__restrict__ in, int64_t n, int64_t *__restrict__ sums)
{
    __shared__ int64_t part[8];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t s = 0;
    for (int k = threadIdx.x; k < SCAN_TILE; k += 256)
        if (base + k < n) s += in[base + k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int w = 0; w < 8; w++) t += part[w];
        sums[blockIdx.x] = t;
    }
}

/* one block: sums[] -> exclusive offsets in place, total in sums[nbBlocks] */
__global__ void __launch_bounds__(1024) scan_sums(int64_t *__restrict__ sums, int64_t nbBlocks)
{
    __shared__ int64_t warpTot[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nbBlocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < nbBlocks ? sums[i] : 0;
        int64_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warpTot[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int64_t w = warpTot[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            warpTot[threadIdx.x] = w;
        }
        __syncthreads();
        const int64_t before = carry + ((threadIdx.x >> 5) ? warpTot[(threadIdx.x >> 5) - 1] : 0) + x - v;
        if (i < nbBlocks) sums[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[nbBlocks] = carry;
}

/* out[i] = offset of item i (int64 or int32 view), out[n] = total */
template <typename OUT>
__global__ void __launch_bounds__(256) scan_apply(const int *__restrict__ in, int64_t n, const int64_t *__restrict__ sums,
                                                  OUT *__restrict__ out)
{
    __shared__ int64_t warpTot[8];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    constexpr int PER = SCAN_TILE / 256;      /* consecutive items per thread */
    int v[PER];
    int64_t mine = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int64_t i = base + (int64_t)threadIdx.x * PER + k;
        v[k] = i < n ? in[i] : 0;
        mine += v[k];
    }
    int64_t x = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warpTot[threadIdx.x >> 5] = x;
    __syncthreads();
    int64_t before = sums[blockIdx.x] + x - mine;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) before += warpTot[w];
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int64_t i = base + (int64_t)threadIdx.x * PER + k;
        if (i < n) out[i] = (OUT)before;
        before += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 255) out[n] = (OUT)sums[gridDim.x];
}

__global__ void __launch_bounds__(256) csr_fill_n2e(const int *__restrict__ conn, int64_t nbEntries,
                                                    const int64_t *__restrict__ n2ePtr, int *__restrict__ cursor,
                                                    int *__restrict__ n2eVal)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbEntries; i += stride) {
        const int n = __ldg(conn + i) - 1;
        const int pos = atomicAdd(cursor + n, 1);
        n2eVal[n2ePtr[n] + pos] = (int)(i >> 2);
    }
}

/* ascending element lists (the order DC_create_nodeToElem's couples have after its sort):
 * one warp per node, rank = number of smaller entries */
__global__ void __launch_bounds__(256) csr_sort_n2e(const int64_t *__restrict__ n2ePtr, int nbNodes,
                                                    const int *__restrict__ in, int *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warpsTotal = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); n < nbNodes; n += warpsTotal) {
        const int64_t p0 = n2ePtr[n];
        const int m = (int)(n2ePtr[n + 1] - p0);
        for (int i = lane; i < m; i += 32) {
            const int v = in[p0 + i];
            int rank = 0;
            for (int j = 0; j < m; j++) rank += in[p0 + j] < v;     /* an element holds a node once: all distinct */
            out[p0 + rank] = v;
        }
    }
}

constexpr int CSR_STAGE = 512;      /* candidates staged per warp; larger rows read them through L1 */

template <bool FILL>
__global__ void __launch_bounds__(256) csr_rows(const int *__restrict__ conn, const int64_t *__restrict__ n2ePtr,
                                                const int *__restrict__ n2eVal, int nbNodes,
                                                int *__restrict__ count, const int *__restrict__ row,
                                                int *__restrict__ col, int colBase)
{
    __shared__ int stageAll[8][CSR_STAGE];
    __shared__ unsigned keptAll[8][CSR_STAGE / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *stage = stageAll[warp];
    unsigned *keptBits = keptAll[warp];
    const int64_t warpsTotal = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; n < nbNodes; n += warpsTotal) {
        const int64_t p0 = n2ePtr[n];
        const int m = 4 * (int)(n2ePtr[n + 1] - p0);                  /* candidates */
        const bool staged = m <= CSR_STAGE;
        auto cand = [&](int i) -> int {
            return staged ? stage[i] : __ldg(conn + (int64_t)__ldg(n2eVal + p0 + (i >> 2)) * 4 + (i & 3)) - 1;
        };
        __syncwarp();
        if (staged)
            for (int i = lane; i < m; i += 32)
                stage[i] = __ldg(conn + (int64_t)__ldg(n2eVal + p0 + (i >> 2)) * 4 + (i & 3)) - 1;
        __syncwarp();
        /* pass 1: a candidate is kept iff no earlier candidate has its value */
        int kept = 0;
        for (int i0 = 0; i0 < m; i0 += 32) {
            const int i = i0 + lane;
            bool first = i < m;
            const int v = first ? cand(i) : 0;
            if (first)
                for (int j = 0; j < i; j++) first &= cand(j) != v;
            const unsigned bits = __ballot_sync(0xffffffffu, first);
            if (staged && lane == 0) keptBits[i0 >> 5] = bits;
            kept += __popc(bits);
            if (FILL && !staged && first) {
                /* rows too long for the staging area: rank by re-deriving "kept" on the fly */
                int rank = 0;
                for (int j = 0; j < m; j++) {
                    const int w = cand(j);
                    if (w < v) {
                        bool firstJ = true;
                        for (int q = 0; q < j && firstJ; q++) firstJ = cand(q) != w;
                        rank += firstJ;
                    }
                }
                col[row[n] + rank] = v + colBase;
            }
        }
        if (!FILL) {
            if (lane == 0) count[n] = kept;
            continue;
        }
        if (!staged) continue;
        __syncwarp();
        /* pass 2: a kept candidate goes to the rank of its value among the kept ones */
        for (int i0 = 0; i0 < m; i0 += 32) {
            const int i = i0 + lane;
            if (i < m && ((keptBits[i0 >> 5] >> lane) & 1u)) {
                const int v = stage[i];
                int rank = 0;
                for (int j = 0; j < m; j++) rank += (((keptBits[j >> 5] >> (j & 31)) & 1u) != 0u) & (stage[j] < v);
                col[row[n] + rank] = v + colBase;
            }
        }
    }
}

}  // namespace dcdev
#endif
