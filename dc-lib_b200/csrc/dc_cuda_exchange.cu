/* dc_cuda_exchange.cu -- interface exchange between the GPUs of one node (include/dc_cuda.h).
 *
 * The reference leaves the halo exchange to the user's MPI/GASPI code (userCommFct,
 * src/tree_traversal.cc:65-73) and only prepares, per neighbour domain, the list of interface nodes
 * and where they go in the neighbour's receive buffer (intfIndex / intfNodes / intfDstOffsets,
 * include/DC.h:51-53,141-143; src/tree_creation.cc:48-169).  Here the same triple exists at CSR-edge
 * granularity, and the transport is peer memory: every rank owns one RECEIVE buffer per neighbour
 * (cudaMalloc, exported with a CUDA IPC handle -- or used through its plain pointer when both ranks
 * live in one process); the neighbour's pack kernel gathers the partial sums of the shared edges
 * and stores them straight into that buffer over NVLink (posted writes), then raises a flag in the
 * same buffer; the owner's unpack kernel waits for the flag and adds the buffer to its values.  No
 * host synchronisation, no NCCL kernel that would need room beside the persistent assembly grid:
 * a step is two kernel launches per neighbour and can be captured in a CUDA graph.
 *
 * Buffers alternate with the parity of the step.  Rank A may overwrite parity p two steps later
 * only after its own unpack of the step in between, which waited for B's flag of that step, which
 * B raised after (in stream order) its unpack of the earlier step: two parities are enough.
 *
 * Kernels that wait for a flag another GPU raises must not share one GPU with the kernel that
 * raises it unless that kernel was launched earlier into the same stream; dc_cuda_exchange_run on
 * several exchange objects of ONE process and ONE stream must therefore pack all before unpacking
 * any (dc_cuda_exchange_pack / _unpack are exported separately for that and for overlap).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>
#include <string>
#include <vector>
#include "dc_cuda.h"

namespace {

constexpr int HEADER_BYTES = 256;           /* flag of parity 0 at byte 0, of parity 1 at byte 128 */
constexpr uint32_t MAGIC = 0x44435832u;     /* "DCX2" */

struct Handle {                              /* DC_EXCHANGE_HANDLE_BYTES = 128 */
    uint32_t magic;
    int32_t pid, device, perEdge;
    int64_t nbEdges;
    uint64_t ptr;                            /* device pointer in the owner's process */
    cudaIpcMemHandle_t ipc;                  /* 64 bytes */
    char pad[128 - 32 - sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(Handle) == DC_EXCHANGE_HANDLE_BYTES, "handle layout");

struct Neighbour {
    int rank = -1;
    int64_t nbEdges = 0, words = 0, stride = 0;      /* stride: bytes between the two parities */
    const int64_t *d_edges = nullptr;                /* into Exchange::d_edges */
    unsigned char *recv = nullptr;                   /* my receive buffer for this neighbour (owned) */
    unsigned char *peer = nullptr;                   /* the neighbour's receive buffer for me (mapped) */
    bool peerMapped = false;                         /* opened with cudaIpcOpenMemHandle */
    unsigned int *d_done = nullptr;                  /* blocks of the pack kernel that have finished */
};

thread_local std::string g_xerr;
int xfail(int code, const char *what, cudaError_t e = cudaSuccess)
{
    g_xerr = what;
    if (e != cudaSuccess) { g_xerr += ": "; g_xerr += cudaGetErrorString(e); }
    return code;
}
#define XCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return xfail(-10, #call, e_); } while (0)

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int XCH_ILP = 8;     /* words in flight per thread */

/* buf[k][:] = values[edge[k]][:] into the NEIGHBOUR's receive buffer; the block that finishes last
 * raises the neighbour's flag (every block fences its stores system-wide before it counts itself
 * done, so the flag is raised after all of them) */
__global__ void __launch_bounds__(256) push_edges(const double *__restrict__ values, const int64_t *__restrict__ edge,
                                                  int64_t nbEdges, int perEdge, double *__restrict__ buf,
                                                  unsigned long long *flag, unsigned long long stamp, unsigned int *done)
{
    const int64_t total = nbEdges * perEdge, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w0 < total; w0 += XCH_ILP * stride) {
        double v[XCH_ILP];
#pragma unroll
        for (int q = 0; q < XCH_ILP; q++) {
            const int64_t w = w0 + q * stride;
            if (w < total) {
                const int64_t k = w / perEdge;
                v[q] = values[__ldg(edge + k) * perEdge + (w - k * perEdge)];
            }
        }
#pragma unroll
        for (int q = 0; q < XCH_ILP; q++)
            if (w0 + q * stride < total) buf[w0 + q * stride] = v[q];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(done, 1u) == gridDim.x - 1) {
            *done = 0;
            __threadfence_system();
            st_release_sys(flag, stamp);
        }
    }
}

/* values[edge[k]][:] += buf[k][:] once the flag shows this step's stamp */
__global__ void __launch_bounds__(256) wait_add_edges(double *__restrict__ values, const int64_t *__restrict__ edge,
                                                      int64_t nbEdges, int perEdge, const double *__restrict__ buf,
                                                      const unsigned long long *flag, unsigned long long stamp)
{
    if (threadIdx.x == 0)
        while (ld_acquire_sys(flag) < stamp) __nanosleep(64);
    __syncthreads();
    const int64_t total = nbEdges * perEdge, stride = (int64_t)gridDim.x * blockDim.x;
    /* XCH_ILP independent words per thread and trip: the kernel often runs on the few CTA slots the
     * persistent assembly grid leaves free, where only memory-level parallelism hides the latency */
    for (int64_t w0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w0 < total; w0 += XCH_ILP * stride) {
        double *dst[XCH_ILP];
        double add[XCH_ILP], old[XCH_ILP];
#pragma unroll
        for (int q = 0; q < XCH_ILP; q++) {
            const int64_t w = w0 + q * stride;
            dst[q] = nullptr;
            if (w < total) {
                const int64_t k = w / perEdge;
                dst[q] = values + __ldg(edge + k) * perEdge + (w - k * perEdge);
                add[q] = __ldcv(buf + w);                /* written by another GPU: not through the read-only path */
                old[q] = *dst[q];
            }
        }
#pragma unroll
        for (int q = 0; q < XCH_ILP; q++)
            if (dst[q]) *dst[q] = __dadd_rn(old[q], add[q]);
    }
}

unsigned grid_of(int64_t work, int cap)
{
    int64_t b = (work + 255) / 256;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace

struct dc_cuda_exchange {
    int device = 0, perEdge = 1, maxBlocks = 64;
    uint64_t step = 0;
    int64_t *d_edges = nullptr;
    bool ownsEdges = true;
    std::vector<Neighbour> nb;
};

extern "C" const char *dc_cuda_exchange_last_error(void) { return g_xerr.c_str(); }

static int create_common(dc_cuda_exchange **out, int device, int perEdge, int nbNeighbours, const int *neighbourRank,
                         const int64_t *edgePtr, const int64_t *edges, bool onDevice)
{
    if (!out || perEdge < 1 || nbNeighbours < 0 || (nbNeighbours > 0 && (!neighbourRank || !edgePtr)))
        return xfail(-1, "dc_cuda_exchange_create: bad argument");
    XCK(cudaSetDevice(device));
    dc_cuda_exchange *x = new dc_cuda_exchange;
    x->device = device;
    x->perEdge = perEdge;
    const int64_t total = nbNeighbours ? edgePtr[nbNeighbours] : 0;
    if (onDevice) {
        x->d_edges = const_cast<int64_t *>(edges);
        x->ownsEdges = false;
    } else {
        cudaError_t e = cudaMalloc((void **)&x->d_edges, sizeof(int64_t) * (size_t)(total > 0 ? total : 1));
        if (e == cudaSuccess && total > 0) e = cudaMemcpy(x->d_edges, edges, sizeof(int64_t) * (size_t)total, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { delete x; return xfail(-10, "dc_cuda_exchange_create: edge list upload", e); }
    }
    x->nb.resize((size_t)nbNeighbours);
    for (int i = 0; i < nbNeighbours; i++) {
        Neighbour &n = x->nb[(size_t)i];
        n.rank = neighbourRank[i];
        n.nbEdges = edgePtr[i + 1] - edgePtr[i];
        n.words = n.nbEdges * perEdge;
        n.stride = ((n.words * 8 + 255) / 256) * 256;
        n.d_edges = x->d_edges + edgePtr[i];
        const size_t bytes = (size_t)HEADER_BYTES + 2 * (size_t)n.stride;
        cudaError_t e = cudaMalloc((void **)&n.recv, bytes);
        if (e == cudaSuccess) e = cudaMemset(n.recv, 0, bytes);
        if (e == cudaSuccess) e = cudaMalloc((void **)&n.d_done, sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMemset(n.d_done, 0, sizeof(unsigned int));
        if (e != cudaSuccess) { dc_cuda_exchange_destroy(x); return xfail(-10, "dc_cuda_exchange_create: receive buffer", e); }
    }
    XCK(cudaDeviceSynchronize());
    *out = x;
    return 0;
}

extern "C" int dc_cuda_exchange_create(dc_cuda_exchange **x, int device, int perEdge, int nbNeighbours,
                                       const int *neighbourRank, const int64_t *edgePtr, const int64_t *h_edges)
{
    return create_common(x, device, perEdge, nbNeighbours, neighbourRank, edgePtr, h_edges, false);
}

extern "C" int dc_cuda_exchange_create_device(dc_cuda_exchange **x, int device, int perEdge, int nbNeighbours,
                                              const int *neighbourRank, const int64_t *edgePtr, const int64_t *d_edges)
{
    return create_common(x, device, perEdge, nbNeighbours, neighbourRank, edgePtr, d_edges, true);
}

extern "C" void dc_cuda_exchange_destroy(dc_cuda_exchange *x)
{
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (Neighbour &n : x->nb) {
        if (n.peerMapped && n.peer) cudaIpcCloseMemHandle(n.peer);
        cudaFree(n.recv);
        cudaFree(n.d_done);
    }
    if (x->ownsEdges) cudaFree(x->d_edges);
    delete x;
}

extern "C" int dc_cuda_exchange_neighbours(const dc_cuda_exchange *x) { return x ? (int)x->nb.size() : 0; }

extern "C" int dc_cuda_exchange_export(dc_cuda_exchange *x, int nbIndex, void *handle)
{
    if (!x || !handle || nbIndex < 0 || nbIndex >= (int)x->nb.size()) return xfail(-1, "dc_cuda_exchange_export: bad argument");
    XCK(cudaSetDevice(x->device));
    Neighbour &n = x->nb[(size_t)nbIndex];
    Handle h;
    memset(&h, 0, sizeof h);
    h.magic = MAGIC;
    h.pid = (int32_t)getpid();
    h.device = x->device;
    h.perEdge = x->perEdge;
    h.nbEdges = n.nbEdges;
    h.ptr = (uint64_t)(uintptr_t)n.recv;
    XCK(cudaIpcGetMemHandle(&h.ipc, n.recv));
    memcpy(handle, &h, sizeof h);
    return 0;
}

extern "C" int dc_cuda_exchange_import(dc_cuda_exchange *x, int nbIndex, const void *handle)
{
    if (!x || !handle || nbIndex < 0 || nbIndex >= (int)x->nb.size()) return xfail(-1, "dc_cuda_exchange_import: bad argument");
    XCK(cudaSetDevice(x->device));
    Neighbour &n = x->nb[(size_t)nbIndex];
    Handle h;
    memcpy(&h, handle, sizeof h);
    if (h.magic != MAGIC) return xfail(-2, "dc_cuda_exchange_import: not an exchange handle");
    if (h.nbEdges != n.nbEdges || h.perEdge != x->perEdge)
        return xfail(-3, "dc_cuda_exchange_import: the neighbour lists another number of shared edges");
    if (h.pid == (int32_t)getpid()) {
        /* both ranks in this process: the pointer itself; other device -> peer access */
        if (h.device != x->device) {
            int can = 0;
            XCK(cudaDeviceCanAccessPeer(&can, x->device, h.device));
            if (!can) return xfail(-4, "dc_cuda_exchange_import: no peer access between the two devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return xfail(-10, "cudaDeviceEnablePeerAccess", e);
            cudaGetLastError();
        }
        n.peer = (unsigned char *)(uintptr_t)h.ptr;
        n.peerMapped = false;
    } else {
        void *p = nullptr;
        XCK(cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess));
        n.peer = (unsigned char *)p;
        n.peerMapped = true;
    }
    return 0;
}

extern "C" int dc_cuda_exchange_pack(dc_cuda_exchange *x, const double *d_values, void *stream)
{
    if (!x || !d_values) return xfail(-1, "dc_cuda_exchange_pack: bad argument");
    XCK(cudaSetDevice(x->device));
    const uint64_t stamp = x->step + 1;
    const int par = (int)(x->step & 1);
    for (Neighbour &n : x->nb) {
        if (!n.peer) return xfail(-5, "dc_cuda_exchange_pack: a neighbour's buffer was not imported");
        unsigned long long *flag = reinterpret_cast<unsigned long long *>(n.peer + par * 128);
        double *buf = reinterpret_cast<double *>(n.peer + HEADER_BYTES + (size_t)par * (size_t)n.stride);
        push_edges<<<grid_of(n.words, x->maxBlocks), 256, 0, (cudaStream_t)stream>>>(d_values, n.d_edges, n.nbEdges, x->perEdge,
                                                                                    buf, flag, stamp, n.d_done);
    }
    XCK(cudaGetLastError());
    return 0;
}

extern "C" int dc_cuda_exchange_unpack(dc_cuda_exchange *x, double *d_values, void *stream)
{
    if (!x || !d_values) return xfail(-1, "dc_cuda_exchange_unpack: bad argument");
    XCK(cudaSetDevice(x->device));
    const uint64_t stamp = x->step + 1;
    const int par = (int)(x->step & 1);
    /* neighbours in the order given at creation (ascending rank by convention): a fixed summation
     * order, so a value shared by several ranks is reproducible from run to run */
    for (Neighbour &n : x->nb) {
        const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(n.recv + par * 128);
        const double *buf = reinterpret_cast<const double *>(n.recv + HEADER_BYTES + (size_t)par * (size_t)n.stride);
        wait_add_edges<<<grid_of(n.words, x->maxBlocks), 256, 0, (cudaStream_t)stream>>>(d_values, n.d_edges, n.nbEdges,
                                                                                        x->perEdge, buf, flag, stamp);
    }
    XCK(cudaGetLastError());
    x->step++;
    return 0;
}

extern "C" int dc_cuda_exchange_run(dc_cuda_exchange *x, double *d_values, void *stream)
{
    const int rc = dc_cuda_exchange_pack(x, d_values, stream);
    return rc ? rc : dc_cuda_exchange_unpack(x, d_values, stream);
}

extern "C" int64_t dc_cuda_exchange_bytes(const dc_cuda_exchange *x)
{
    int64_t b = 0;
    if (x) for (const Neighbour &n : x->nb) b += 2 * 8 * n.words;      /* sent + received */
    return b;
}

extern "C" int dc_cuda_exchange_set_blocks(dc_cuda_exchange *x, int maxBlocks)
{
    if (!x || maxBlocks < 1) return xfail(-1, "dc_cuda_exchange_set_blocks: bad argument");
    x->maxBlocks = maxBlocks;
    return 0;
}
