/* dc_cuda_tree.cu -- the level-wise element split of DC_create_tree on the device (dc_cuda_tree_*,
 * include/dc_cuda.h).
 *
 * What the reference does at every node of the D&C tree (src/tree_creation.cc:379-479):
 * create_elem_part classes the node's elements by the partition of their nodes (:352-375: all on
 * the left of the pivot part / all on the right / straddling = separator), DC_create_permutation
 * turns the classes into a stable 3-way order (src/permutations.cc:169-180) and
 * DC_permute_int_2d_array moves the connectivity rows (:51-80).  Near the root these passes stream
 * the whole connectivity through the host caches once per level.  Here every level of the tree
 * proper (the separator subtrees stay on the host) is one pass over ALL elements on the device:
 *
 *   tree_classify      class of every element from the pivot of its segment (= tree node of the level)
 *   tree_tile_counts   per tile of 1024 elements, how many of each class
 *   tree_scan_tiles    exclusive prefix over the tiles (one block)
 *   tree_prefix        P_k[i] = number of class-k elements before position i, k = 0, 1, 2
 *   tree_seg_totals    per segment (left, right, separator) counts = P_k[end] - P_k[first]
 *   tree_scatter       row, original index and next-level segment of every element to its place:
 *                      first + offset of its class + its rank inside (segment, class)
 *
 * The host keeps the bookkeeping the recursion did (tree nodes, part ranges, node ranges) and
 * only exchanges three integers per segment and level with the device.  Orders are stable, so
 * connectivity, element permutation and tree equal the host builder's byte for byte.
 * No CPU fallback: without a CUDA device every entry point returns an error. */
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "dc_cuda.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char *what, cudaError_t e = cudaSuccess)
{
    g_err = what;
    if (e != cudaSuccess) {
        g_err += ": ";
        g_err += cudaGetErrorString(e);
    }
    return code;
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return fail(-10, #call, e_);        \
    } while (0)

constexpr int TILE = 1024;            /* elements per block of the class scans: 256 threads x 4 */
constexpr int CLS_IDLE = 3;           /* element of no segment of this level (leaf or separator already) */

__global__ void __launch_bounds__(256) tree_fill(int *__restrict__ posToOrig, int *__restrict__ segOf, int64_t nbElem)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbElem; i += stride) {
        posToOrig[i] = (int)i;
        segOf[i] = 0;
    }
}

/* reference: create_elem_part, tree_creation.cc:352-375 */
__global__ void __launch_bounds__(256) tree_classify(const int *__restrict__ conn, int dimElem, int64_t nbElem,
                                                     const int *__restrict__ segOf, const int *__restrict__ pivot,
                                                     const int *__restrict__ nodePart, int nbNodes,
                                                     uint8_t *__restrict__ cls, int *__restrict__ bad)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbElem; i += stride) {
        const int s = segOf[i];
        if (s < 0) { cls[i] = CLS_IDLE; continue; }
        const int p = __ldg(pivot + s);
        int onLeft = 0;
        if (dimElem == 4) {
            const int4 c = __ldg(reinterpret_cast<const int4 *>(conn) + i);
            const int n[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int node = n[j];
                if ((unsigned)node >= (unsigned)nbNodes) { *bad = 1; node = 0; }
                onLeft += __ldg(nodePart + node) <= p;
            }
        } else {
            for (int j = 0; j < dimElem; j++) {
                int node = __ldg(conn + i * dimElem + j);
                if ((unsigned)node >= (unsigned)nbNodes) { *bad = 1; node = 0; }
                onLeft += __ldg(nodePart + node) <= p;
            }
        }
        cls[i] = (uint8_t)(onLeft == dimElem ? 0 : (onLeft == 0 ? 1 : 2));
    }
}

/* tileCnt[k * stride + b] = elements of class k in tile b (cls is padded with CLS_IDLE) */
__global__ void __launch_bounds__(256) tree_tile_counts(const uint8_t *__restrict__ cls, int *__restrict__ tileCnt,
                                                        int64_t tileStride)
{
    __shared__ int part[8][3];
    const uchar4 c = reinterpret_cast<const uchar4 *>(cls)[(int64_t)blockIdx.x * (TILE / 4) + threadIdx.x];
    const int v[4] = {c.x, c.y, c.z, c.w};
    int cnt[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < 3; k++) cnt[k] += v[j] == k;
#pragma unroll
    for (int k = 0; k < 3; k++)
        for (int o = 16; o > 0; o >>= 1) cnt[k] += __shfl_down_sync(0xffffffffu, cnt[k], o);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 3; k++) part[threadIdx.x >> 5][k] = cnt[k];
    __syncthreads();
    if (threadIdx.x < 3) {
        int t = 0;
        for (int w = 0; w < 8; w++) t += part[w][threadIdx.x];
        tileCnt[threadIdx.x * tileStride + blockIdx.x] = t;
    }
}

/* one block: the three rows of tileCnt -> exclusive prefixes in place */
__global__ void __launch_bounds__(1024) tree_scan_tiles(int *__restrict__ tileCnt, int64_t nbTiles, int64_t tileStride)
{
    __shared__ int warpTot[32];
    __shared__ int carry;
    for (int k = 0; k < 3; k++) {
        int *row = tileCnt + k * tileStride;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (int64_t base = 0; base < nbTiles; base += 1024) {
            const int64_t i = base + threadIdx.x;
            const int v = i < nbTiles ? row[i] : 0;
            int x = v;
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if ((threadIdx.x & 31) >= o) x += y;
            }
            if ((threadIdx.x & 31) == 31) warpTot[threadIdx.x >> 5] = x;
            __syncthreads();
            if (threadIdx.x < 32) {
                int w = warpTot[threadIdx.x];
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, w, o);
                    if (threadIdx.x >= o) w += y;
                }
                warpTot[threadIdx.x] = w;
            }
            __syncthreads();
            const int before = carry + ((threadIdx.x >> 5) ? warpTot[(threadIdx.x >> 5) - 1] : 0) + x - v;
            if (i < nbTiles) row[i] = before;
            __syncthreads();
            if (threadIdx.x == 1023) carry = before + v;
            __syncthreads();
        }
    }
}

/* P_k[i] for the four consecutive positions of every thread, k = 0, 1, 2 (16-byte stores) */
__global__ void __launch_bounds__(256) tree_prefix(const uint8_t *__restrict__ cls, const int *__restrict__ tileCnt,
                                                   int64_t tileStride, int *__restrict__ P0, int *__restrict__ P1,
                                                   int *__restrict__ P2)
{
    __shared__ int warpTot[8][3];
    const int64_t quad = (int64_t)blockIdx.x * (TILE / 4) + threadIdx.x;
    const uchar4 c = reinterpret_cast<const uchar4 *>(cls)[quad];
    const int v[4] = {c.x, c.y, c.z, c.w};
    int mine[3] = {0, 0, 0}, x[3];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < 3; k++) mine[k] += v[j] == k;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        x[k] = mine[k];
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x[k], o);
            if ((threadIdx.x & 31) >= o) x[k] += y;
        }
        if ((threadIdx.x & 31) == 31) warpTot[threadIdx.x >> 5][k] = x[k];
    }
    __syncthreads();
    int *const P[3] = {P0, P1, P2};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int before = tileCnt[k * tileStride + blockIdx.x] + x[k] - mine[k];
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) before += warpTot[w][k];
        int4 o;
        o.x = before; before += v[0] == k;
        o.y = before; before += v[1] == k;
        o.z = before; before += v[2] == k;
        o.w = before;
        reinterpret_cast<int4 *>(P[k])[quad] = o;
    }
}

__global__ void __launch_bounds__(256) tree_seg_totals(const int *__restrict__ segFirst, const int *__restrict__ segEnd,
                                                       int nbSeg, const int *__restrict__ P0, const int *__restrict__ P1,
                                                       const int *__restrict__ P2, int *__restrict__ totals)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nbSeg) return;
    const int a = segFirst[s], b = segEnd[s];
    totals[3 * s] = P0[b] - P0[a];
    totals[3 * s + 1] = P1[b] - P1[a];
    totals[3 * s + 2] = P2[b] - P2[a];
}

/* reference: the stable 3-way order of DC_create_permutation (permutations.cc:169-180) applied by
 * DC_permute_int_2d_array (:51-80) and merged into elemPerm (:118-130; here the inverse, position ->
 * original element, travels with the rows) */
__global__ void __launch_bounds__(256) tree_scatter(const int *__restrict__ connIn, int *__restrict__ connOut, int dimElem,
                                                    int64_t nbElem, const int *__restrict__ posIn, int *__restrict__ posOut,
                                                    const int *__restrict__ segIn, int *__restrict__ segOut,
                                                    const uint8_t *__restrict__ cls, const int *__restrict__ segFirst,
                                                    const int *__restrict__ segEnd, const int *__restrict__ child,
                                                    const int *__restrict__ P0, const int *__restrict__ P1,
                                                    const int *__restrict__ P2)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbElem; i += stride) {
        const int s = segIn[i];
        int64_t dest = i;
        int nextSeg = -1;
        if (s >= 0) {
            const int k = cls[i];
            const int a = __ldg(segFirst + s), b = __ldg(segEnd + s);
            const int t0 = P0[b] - P0[a], t1 = P1[b] - P1[a];
            const int *Pk = k == 0 ? P0 : (k == 1 ? P1 : P2);
            const int off = k == 0 ? 0 : (k == 1 ? t0 : t0 + t1);
            dest = (int64_t)a + off + (Pk[i] - Pk[a]);
            nextSeg = __ldg(child + 3 * s + k);
        }
        if (dimElem == 4) {
            reinterpret_cast<int4 *>(connOut)[dest] = __ldg(reinterpret_cast<const int4 *>(connIn) + i);
        } else {
            for (int j = 0; j < dimElem; j++) connOut[dest * dimElem + j] = __ldg(connIn + i * dimElem + j);
        }
        posOut[dest] = posIn[i];
        segOut[dest] = nextSeg;
    }
}

int grid_for(int64_t n, int threads)
{
    const int64_t b = (n + threads - 1) / threads;
    const int64_t cap = 148 * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

struct dc_cuda_tree {
    int device = 0;
    int64_t nbElem = 0, nbTiles = 0;
    int dimElem = 0, nbNodes = 0;
    int *conn[2] = {nullptr, nullptr}, *pos[2] = {nullptr, nullptr}, *seg[2] = {nullptr, nullptr};
    int cur = 0;
    uint8_t *cls = nullptr;
    int *P[3] = {nullptr, nullptr, nullptr};
    int *tileCnt = nullptr, *nodePart = nullptr, *bad = nullptr;
    int *segFirst = nullptr, *segEnd = nullptr, *pivot = nullptr, *child = nullptr, *totals = nullptr;
    int segCap = 0, nbSeg = 0;
    int64_t launches = 0;
    cudaStream_t stream = nullptr;
};

extern "C" const char *dc_cuda_tree_last_error(void) { return g_err.c_str(); }

static void tree_free(dc_cuda_tree *t)
{
    if (!t) return;
    cudaSetDevice(t->device);
    for (int k = 0; k < 2; k++) { cudaFree(t->conn[k]); cudaFree(t->pos[k]); cudaFree(t->seg[k]); }
    for (int k = 0; k < 3; k++) cudaFree(t->P[k]);
    cudaFree(t->cls); cudaFree(t->tileCnt); cudaFree(t->nodePart); cudaFree(t->bad);
    cudaFree(t->segFirst); cudaFree(t->segEnd); cudaFree(t->pivot); cudaFree(t->child); cudaFree(t->totals);
    if (t->stream) cudaStreamDestroy(t->stream);
    delete t;
}

extern "C" int dc_cuda_tree_begin(dc_cuda_tree **out, int device, const int *h_conn, int nbElem, int dimElem,
                                  const int *h_nodePart, int nbNodes)
{
    if (!out || !h_conn || !h_nodePart) return fail(-1, "dc_cuda_tree_begin: null argument");
    if (nbElem < 1 || dimElem < 1 || nbNodes < 1) return fail(-1, "dc_cuda_tree_begin: empty mesh");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(-2, "dc_cuda_tree_begin: no CUDA device (this library has no CPU fallback)", e);
    if (device < 0 || device >= n) return fail(-3, "dc_cuda_tree_begin: bad device index");
    CK(cudaSetDevice(device));
    dc_cuda_tree *t = new dc_cuda_tree;
    t->device = device;
    t->nbElem = nbElem;
    t->dimElem = dimElem;
    t->nbNodes = nbNodes;
    t->nbTiles = ((int64_t)nbElem + 1 + TILE - 1) / TILE;          /* position nbElem has prefixes too */
    const size_t rowBytes = (size_t)nbElem * dimElem * sizeof(int), padded = (size_t)t->nbTiles * TILE;
#define CKT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { tree_free(t); return fail(-10, #call, e_); } } while (0)
    CKT(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
        CKT(cudaMalloc(&t->conn[k], rowBytes));
        CKT(cudaMalloc(&t->pos[k], (size_t)nbElem * sizeof(int)));
        CKT(cudaMalloc(&t->seg[k], (size_t)nbElem * sizeof(int)));
    }
    for (int k = 0; k < 3; k++) CKT(cudaMalloc(&t->P[k], padded * sizeof(int)));
    CKT(cudaMalloc(&t->cls, padded));
    CKT(cudaMalloc(&t->tileCnt, (size_t)t->nbTiles * 3 * sizeof(int)));
    CKT(cudaMalloc(&t->nodePart, (size_t)nbNodes * sizeof(int)));
    CKT(cudaMalloc(&t->bad, sizeof(int)));
    CKT(cudaMemsetAsync(t->bad, 0, sizeof(int), t->stream));
    CKT(cudaMemsetAsync(t->cls, CLS_IDLE, padded, t->stream));
    CKT(cudaMemcpyAsync(t->conn[0], h_conn, rowBytes, cudaMemcpyHostToDevice, t->stream));
    CKT(cudaMemcpyAsync(t->nodePart, h_nodePart, (size_t)nbNodes * sizeof(int), cudaMemcpyHostToDevice, t->stream));
    tree_fill<<<grid_for(nbElem, 256), 256, 0, t->stream>>>(t->pos[0], t->seg[0], nbElem);
    t->launches++;
    CKT(cudaGetLastError());
    CKT(cudaStreamSynchronize(t->stream));
#undef CKT
    *out = t;
    return 0;
}

/* Classes, prefixes and per-segment counts of one level.  Segments: disjoint position ranges
 * [h_segFirst[s], h_segEnd[s]) whose elements carry segment number s (set by the previous scatter; level 0 =
 * one segment holding everything), pivot part h_pivot[s].  h_totals[3 s + {0,1,2}] = elements on the left /
 * on the right / straddling. */
extern "C" int dc_cuda_tree_classify(dc_cuda_tree *t, int nbSeg, const int *h_segFirst, const int *h_segEnd,
                                     const int *h_pivot, int *h_totals)
{
    if (!t || nbSeg < 1 || !h_segFirst || !h_segEnd || !h_pivot || !h_totals)
        return fail(-1, "dc_cuda_tree_classify: bad argument");
    CK(cudaSetDevice(t->device));
    if (nbSeg > t->segCap) {
        cudaFree(t->segFirst); cudaFree(t->segEnd); cudaFree(t->pivot); cudaFree(t->child); cudaFree(t->totals);
        t->segFirst = t->segEnd = t->pivot = t->child = t->totals = nullptr;
        t->segCap = 0;
        const int cap = nbSeg + nbSeg / 2 + 1024;
        CK(cudaMalloc(&t->segFirst, (size_t)cap * sizeof(int)));
        CK(cudaMalloc(&t->segEnd, (size_t)cap * sizeof(int)));
        CK(cudaMalloc(&t->pivot, (size_t)cap * sizeof(int)));
        CK(cudaMalloc(&t->child, (size_t)cap * 3 * sizeof(int)));
        CK(cudaMalloc(&t->totals, (size_t)cap * 3 * sizeof(int)));
        t->segCap = cap;
    }
    t->nbSeg = nbSeg;
    cudaStream_t st = t->stream;
    CK(cudaMemcpyAsync(t->segFirst, h_segFirst, (size_t)nbSeg * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(t->segEnd, h_segEnd, (size_t)nbSeg * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(t->pivot, h_pivot, (size_t)nbSeg * sizeof(int), cudaMemcpyHostToDevice, st));
    tree_classify<<<grid_for(t->nbElem, 256), 256, 0, st>>>(t->conn[t->cur], t->dimElem, t->nbElem, t->seg[t->cur], t->pivot,
                                                            t->nodePart, t->nbNodes, t->cls, t->bad);
    tree_tile_counts<<<(unsigned)t->nbTiles, 256, 0, st>>>(t->cls, t->tileCnt, t->nbTiles);
    tree_scan_tiles<<<1, 1024, 0, st>>>(t->tileCnt, t->nbTiles, t->nbTiles);
    tree_prefix<<<(unsigned)t->nbTiles, 256, 0, st>>>(t->cls, t->tileCnt, t->nbTiles, t->P[0], t->P[1], t->P[2]);
    tree_seg_totals<<<(nbSeg + 255) / 256, 256, 0, st>>>(t->segFirst, t->segEnd, nbSeg, t->P[0], t->P[1], t->P[2], t->totals);
    t->launches += 5;
    CK(cudaGetLastError());
    int bad = 0;
    CK(cudaMemcpyAsync(h_totals, t->totals, (size_t)nbSeg * 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&bad, t->bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (bad) return fail(-4, "dc_cuda_tree_classify: node index outside [0, nbNodes)");
    return 0;
}

/* Moves every element of the level just classified to its place.  h_child[3 s + k] = segment number
 * the elements of class k of segment s carry at the next level, or -1 when that child is a leaf
 * (k = 2, the separator, is always -1: separator subtrees are built on the host). */
extern "C" int dc_cuda_tree_scatter(dc_cuda_tree *t, const int *h_child)
{
    if (!t || !h_child || t->nbSeg < 1) return fail(-1, "dc_cuda_tree_scatter: no classified level");
    CK(cudaSetDevice(t->device));
    cudaStream_t st = t->stream;
    const int in = t->cur, out = in ^ 1;
    CK(cudaMemcpyAsync(t->child, h_child, (size_t)t->nbSeg * 3 * sizeof(int), cudaMemcpyHostToDevice, st));
    tree_scatter<<<grid_for(t->nbElem, 256), 256, 0, st>>>(t->conn[in], t->conn[out], t->dimElem, t->nbElem, t->pos[in],
                                                           t->pos[out], t->seg[in], t->seg[out], t->cls, t->segFirst,
                                                           t->segEnd, t->child, t->P[0], t->P[1], t->P[2]);
    t->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    t->cur = out;
    t->nbSeg = 0;
    return 0;
}

/* Copies connectivity rows and the position -> original element map back (either may be NULL) and
 * frees the work space; *launches (may be NULL) = kernels launched since begin. */
extern "C" int dc_cuda_tree_end(dc_cuda_tree *t, int *h_conn, int *h_posToOrig, int64_t *launches)
{
    if (!t) return fail(-1, "dc_cuda_tree_end: null handle");
    int rc = 0;
    cudaError_t e = cudaSetDevice(t->device);
    if (e == cudaSuccess && h_conn)
        e = cudaMemcpy(h_conn, t->conn[t->cur], (size_t)t->nbElem * t->dimElem * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && h_posToOrig)
        e = cudaMemcpy(h_posToOrig, t->pos[t->cur], (size_t)t->nbElem * sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(-10, "dc_cuda_tree_end: copy back", e);
    if (launches) *launches = t->launches;
    tree_free(t);
    return rc;
}
