/* dc_cuda_api.cu -- implementation of the C ABI in include/dc_cuda.h.
 * There is no CPU fallback anywhere in this file: every entry point either runs
 * on the selected CUDA device or returns an error. */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "dc_cuda.h"
#include "dc_device_operator.cuh"
#include "dc_csr.cuh"

using namespace dcdev;

static thread_local std::string g_err;

static int fail(int code, const char *what, cudaError_t e = cudaSuccess)
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

struct dc_cuda_ctx {
    int device = 0;
    int smCount = 148;
    int nbElem = 0, nbNodes = 0, colBase = 0;
    int64_t nnz = 0;
    bool ownsMesh = false;
    int *d_conn = nullptr, *d_row = nullptr, *d_col = nullptr;
    double *d_coord = nullptr;
    /* plan */
    int maxSlots = 0, maxHalo = 0, maxNodes = 0, maxTasks = 0, maxItems = 0, maxTgt = 0, symmetric = 0;
    uint16_t *d_slotNodes = nullptr;
    int *d_tgt = nullptr;
    int64_t *d_bigEdgeT = nullptr;
    int threads = 256;       /* CTA size of the unit kernel (128, 256 or 512) */
    int persistent = 1;      /* persistent grid with next-unit prefetch (0: one CTA per unit) */
    int64_t nbUnits = 0, nbBigEdges = 0;
    bool ownsPlan = true;
    dc_unit_t *d_units = nullptr;
    dc_task_t *d_tasks = nullptr;
    uint16_t *d_items = nullptr;
    int *d_halo = nullptr, *d_diagRel = nullptr;
    int64_t *d_bigEdge = nullptr, *d_bigPtr = nullptr;
    uint32_t *d_bigItem = nullptr;
    /* full lists (DC_MODE_LIST) */
    int64_t nbListEdges = 0;
    int64_t *d_listEdge = nullptr, *d_listPtr = nullptr;
    uint32_t *d_listItem = nullptr;
    /* context-owned values */
    double *d_values = nullptr;
    int64_t valuesLen = 0;
    int64_t launches = 0;
    long long *d_phaseClock = nullptr;   /* debug: per-unit phase stamps (dc_cuda_debug_phases) */
    /* multi-GPU step (dc_cuda_assemble_exchange): units that own interface rows come first in d_units */
    int64_t nbFrontUnits = 0;
    int exchangeReserve = -1;            /* CTA slots the interior launch leaves to the pack / unpack kernels; -1: 4 or 16 (below) */
    int exchangePackEarly = 0;           /* 1: pack on the whole GPU before the interior launch, 0: beside it */
    cudaStream_t sideStream = nullptr;
    cudaEvent_t evFront = nullptr, evJoin = nullptr;
};

extern "C" const char *dc_cuda_last_error(void) { return g_err.c_str(); }

extern "C" int dc_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int dc_cuda_create(dc_cuda_ctx **out, int device)
{
    if (!out) return fail(-1, "dc_cuda_create: null output");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(-2, "dc_cuda_create: no CUDA device (this library has no CPU fallback)", e);
    if (device < 0 || device >= n) return fail(-3, "dc_cuda_create: bad device index");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(-4, "dc_cuda_create: kernels are built for sm_100a only");
    dc_cuda_ctx *c = new dc_cuda_ctx;
    c->device = device;
    c->smCount = prop.multiProcessorCount;
    *out = c;
    return 0;
}

static void free_mesh(dc_cuda_ctx *c)
{
    if (c->ownsMesh) {
        cudaFree(c->d_conn); cudaFree(c->d_coord); cudaFree(c->d_row); cudaFree(c->d_col);
    }
    c->d_conn = c->d_row = c->d_col = nullptr;
    c->d_coord = nullptr;
    c->ownsMesh = false;
}
static void free_plan(dc_cuda_ctx *c)
{
    if (!c->ownsPlan) {
        c->d_units = nullptr; c->d_tasks = nullptr; c->d_items = nullptr; c->d_halo = c->d_diagRel = c->d_tgt = nullptr;
        c->d_slotNodes = nullptr;
        c->d_bigEdge = c->d_bigPtr = c->d_bigEdgeT = nullptr; c->d_bigItem = nullptr;
        c->ownsPlan = true;
    }
    cudaFree(c->d_slotNodes); c->d_slotNodes = nullptr;
    cudaFree(c->d_tgt); cudaFree(c->d_bigEdgeT);
    c->d_tgt = nullptr; c->d_bigEdgeT = nullptr;
    cudaFree(c->d_units); cudaFree(c->d_tasks); cudaFree(c->d_items); cudaFree(c->d_halo);
    cudaFree(c->d_diagRel); cudaFree(c->d_bigEdge); cudaFree(c->d_bigPtr); cudaFree(c->d_bigItem);
    c->d_units = nullptr; c->d_tasks = nullptr; c->d_items = nullptr; c->d_halo = c->d_diagRel = nullptr;
    c->d_bigEdge = c->d_bigPtr = nullptr; c->d_bigItem = nullptr;
    c->nbUnits = c->nbBigEdges = 0;
    c->nbFrontUnits = 0;
}
static void free_lists(dc_cuda_ctx *c)
{
    cudaFree(c->d_listEdge); cudaFree(c->d_listPtr); cudaFree(c->d_listItem);
    c->d_listEdge = c->d_listPtr = nullptr; c->d_listItem = nullptr;
    c->nbListEdges = 0;
}

extern "C" void dc_cuda_destroy(dc_cuda_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    free_mesh(c); free_plan(c); free_lists(c);
    cudaFree(c->d_values);
    if (c->sideStream) cudaStreamDestroy(c->sideStream);
    if (c->evFront) cudaEventDestroy(c->evFront);
    if (c->evJoin) cudaEventDestroy(c->evJoin);
    delete c;
}

template <typename T>
static int upload(T **dst, const T *src, int64_t count)
{
    *dst = nullptr;
    CK(cudaMalloc((void **)dst, sizeof(T) * (size_t)(count > 0 ? count : 1)));
    if (count > 0) CK(cudaMemcpy(*dst, src, sizeof(T) * (size_t)count, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int dc_cuda_set_mesh(dc_cuda_ctx *c, const int *h_conn, const double *h_coord, const int *h_row,
                                const int *h_col, int nbElem, int nbNodes, int colBase)
{
    if (!c) return fail(-1, "dc_cuda_set_mesh: null context");
    CK(cudaSetDevice(c->device));
    free_mesh(c);
    c->ownsMesh = true;
    c->nbElem = nbElem; c->nbNodes = nbNodes; c->colBase = colBase;
    c->nnz = h_row[nbNodes];
    int rc;
    if ((rc = upload(&c->d_conn, h_conn, (int64_t)nbElem * 4))) return rc;
    /* two spare nodes behind the array: the bulk copy of a unit's own coordinates is rounded up
     * to a 48-byte granule */
    CK(cudaMalloc((void **)&c->d_coord, sizeof(double) * 3 * ((size_t)nbNodes + 2)));
    CK(cudaMemset(c->d_coord, 0, sizeof(double) * 3 * ((size_t)nbNodes + 2)));
    CK(cudaMemcpy(c->d_coord, h_coord, sizeof(double) * 3 * (size_t)nbNodes, cudaMemcpyHostToDevice));
    if ((rc = upload(&c->d_row, h_row, (int64_t)nbNodes + 1))) return rc;
    if ((rc = upload(&c->d_col, h_col, c->nnz))) return rc;
    return 0;
}

extern "C" int dc_cuda_set_mesh_device(dc_cuda_ctx *c, const int *d_conn, const double *d_coord,
                                       const int *d_row, const int *d_col, int nbElem, int nbNodes,
                                       int64_t nnz, int colBase)
{
    if (!c) return fail(-1, "dc_cuda_set_mesh_device: null context");
    if ((reinterpret_cast<uintptr_t>(d_conn) & 15) != 0)
        return fail(-5, "dc_cuda_set_mesh_device: connectivity must be 16-byte aligned");
    CK(cudaSetDevice(c->device));
    free_mesh(c);
    c->ownsMesh = false;
    c->d_conn = const_cast<int *>(d_conn);
    c->d_coord = const_cast<double *>(d_coord);
    c->d_row = const_cast<int *>(d_row);
    c->d_col = const_cast<int *>(d_col);
    c->nbElem = nbElem; c->nbNodes = nbNodes; c->nnz = nnz; c->colBase = colBase;
    return 0;
}

extern "C" int dc_cuda_update_coords(dc_cuda_ctx *c, const double *h_coord, void *stream)
{
    if (!c || !c->d_coord) return fail(-1, "dc_cuda_update_coords: no mesh");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->d_coord, h_coord, sizeof(double) * 3 * (size_t)c->nbNodes, cudaMemcpyHostToDevice,
                       (cudaStream_t)stream));
    return 0;
}

extern "C" int dc_cuda_upload_plan(dc_cuda_ctx *c, const struct dc_plan_s *p)
{
    if (!c || !p) return fail(-1, "dc_cuda_upload_plan: null argument");
    CK(cudaSetDevice(c->device));
    free_plan(c);
    int rc;
    if ((rc = upload(&c->d_units, p->units, p->nbUnits))) return rc;
    if ((rc = upload(&c->d_tasks, p->tasks, p->nbTasks))) return rc;
    if ((rc = upload(&c->d_items, p->items, p->nbItems))) return rc;
    if ((rc = upload(&c->d_halo, p->haloNodes, p->nbHaloNodes))) return rc;
    if ((rc = upload(&c->d_slotNodes, p->slotNodes, p->nbSlotNodes))) return rc;
    if ((rc = upload(&c->d_diagRel, p->diagRel, (int64_t)c->nbNodes))) return rc;
    if ((rc = upload(&c->d_tgt, p->tgt, 2 * p->nbTgt))) return rc;
    if (p->big.nbEdges > 0 && p->big.edgeT)
        if ((rc = upload(&c->d_bigEdgeT, p->big.edgeT, p->big.nbEdges))) return rc;
    if (p->big.nbEdges > 0) {
        if ((rc = upload(&c->d_bigEdge, p->big.edge, p->big.nbEdges))) return rc;
        if ((rc = upload(&c->d_bigPtr, p->big.ptr, p->big.nbEdges + 1))) return rc;
        if ((rc = upload(&c->d_bigItem, p->big.item, p->big.nbItems))) return rc;
    }
    c->nbUnits = p->nbUnits;
    c->nbBigEdges = p->big.nbEdges;
    c->maxSlots = p->maxSlots;
    c->maxHalo = p->maxHalo;
    c->maxNodes = p->maxNodes;
    c->maxTasks = p->maxTasks;
    c->maxItems = p->maxItems;
    c->maxTgt = p->maxTgt;
    c->symmetric = p->symmetric;
    return 0;
}

extern "C" int dc_cuda_set_plan_device(dc_cuda_ctx *c, const void *d_units, const void *d_tasks,
                                       const void *d_items, const void *d_slotNodes, const int *d_haloNodes,
                                       const int *d_diagRel, const int *d_tgt, int64_t nbUnits, int maxSlots,
                                       int maxHalo, int maxNodes, int maxTasks, int maxItems, int maxTgt,
                                       int symmetric)
{
    if (!c) return fail(-1, "dc_cuda_set_plan_device: null context");
    if ((reinterpret_cast<uintptr_t>(d_units) | reinterpret_cast<uintptr_t>(d_tasks) |
         reinterpret_cast<uintptr_t>(d_items) | reinterpret_cast<uintptr_t>(d_haloNodes) |
         reinterpret_cast<uintptr_t>(d_slotNodes) | reinterpret_cast<uintptr_t>(d_tgt)) & 15)
        return fail(-5, "dc_cuda_set_plan_device: arrays must be 16-byte aligned");
    CK(cudaSetDevice(c->device));
    free_plan(c);
    c->ownsPlan = false;
    c->d_units = (dc_unit_t *)d_units;
    c->d_tasks = (dc_task_t *)d_tasks;
    c->d_items = (uint16_t *)d_items;
    c->d_halo = const_cast<int *>(d_haloNodes);
    c->d_slotNodes = (uint16_t *)d_slotNodes;
    c->d_diagRel = const_cast<int *>(d_diagRel);
    c->d_tgt = const_cast<int *>(d_tgt);
    c->nbUnits = nbUnits;
    c->nbBigEdges = 0;
    c->maxSlots = maxSlots; c->maxHalo = maxHalo; c->maxNodes = maxNodes; c->maxTasks = maxTasks;
    c->maxItems = maxItems; c->maxTgt = maxTgt; c->symmetric = symmetric;
    return 0;
}

/* fallback lists (rows too large for a unit) for a schedule adopted with dc_cuda_set_plan_device */
extern "C" int dc_cuda_set_plan_big_device(dc_cuda_ctx *c, const int64_t *d_bigEdge, const int64_t *d_bigPtr,
                                           const uint32_t *d_bigItem, const int64_t *d_bigEdgeT, int64_t nbBigEdges)
{
    if (!c || c->ownsPlan) return fail(-1, "dc_cuda_set_plan_big_device: call dc_cuda_set_plan_device first");
    if (nbBigEdges < 0 || (nbBigEdges > 0 && (!d_bigEdge || !d_bigPtr || !d_bigItem)))
        return fail(-1, "dc_cuda_set_plan_big_device: bad argument");
    if (c->symmetric && nbBigEdges > 0 && !d_bigEdgeT)
        return fail(-1, "dc_cuda_set_plan_big_device: the symmetric schedule needs the transposed edges");
    c->d_bigEdge = const_cast<int64_t *>(d_bigEdge);
    c->d_bigPtr = const_cast<int64_t *>(d_bigPtr);
    c->d_bigItem = const_cast<uint32_t *>(d_bigItem);
    c->d_bigEdgeT = const_cast<int64_t *>(d_bigEdgeT);
    c->nbBigEdges = nbBigEdges;
    return 0;
}

extern "C" int dc_cuda_upload_lists(dc_cuda_ctx *c, const int64_t *h_edge, const int64_t *h_ptr,
                                    const uint32_t *h_item, int64_t nbEdges, int64_t nbItems)
{
    if (!c) return fail(-1, "dc_cuda_upload_lists: null context");
    CK(cudaSetDevice(c->device));
    free_lists(c);
    int rc;
    if ((rc = upload(&c->d_listEdge, h_edge, nbEdges))) return rc;
    if ((rc = upload(&c->d_listPtr, h_ptr, nbEdges + 1))) return rc;
    if ((rc = upload(&c->d_listItem, h_item, nbItems))) return rc;
    c->nbListEdges = nbEdges;
    return 0;
}

static void fill_args(const dc_cuda_ctx *c, int mode, dc_launch_args *a)
{
    memset(a, 0, sizeof *a);
    a->mode = mode; a->threads = c->threads; a->smCount = c->smCount; a->symmetric = c->symmetric;
    a->persistent = c->persistent;
    a->colBase = c->colBase; a->nbElem = c->nbElem; a->nbNodes = c->nbNodes; a->nnz = c->nnz;
    a->nbUnits = c->nbUnits; a->nbBigEdges = c->nbBigEdges; a->nbListEdges = c->nbListEdges;
    a->maxSlots = c->maxSlots; a->maxHalo = c->maxHalo; a->maxNodes = c->maxNodes; a->maxTasks = c->maxTasks;
    a->maxItems = c->maxItems; a->maxTgt = c->maxTgt;
    a->d_conn = c->d_conn; a->d_row = c->d_row; a->d_col = c->d_col; a->d_coord = c->d_coord;
    a->d_units = c->d_units; a->d_tasks = c->d_tasks; a->d_items = c->d_items; a->d_slotNodes = c->d_slotNodes;
    a->d_haloNodes = c->d_halo; a->d_diagRel = c->d_diagRel; a->d_tgt = c->d_tgt;
    a->d_bigEdge = c->d_bigEdge; a->d_bigPtr = c->d_bigPtr; a->d_bigEdgeT = c->d_bigEdgeT; a->d_bigItem = c->d_bigItem;
    a->d_listEdge = c->d_listEdge; a->d_listPtr = c->d_listPtr; a->d_listItem = c->d_listItem;
    a->d_phaseClock = c->d_phaseClock;
    a->unitFirst = 0; a->unitCount = c->nbUnits; a->reserveCtas = 0; a->skipBigRows = 0;
}

static int explain(int rc)
{
    switch (rc) {
        case DC_LAUNCH_OK: return 0;
        case DC_LAUNCH_NO_PLAN: return fail(rc, "dc_cuda_assemble: no schedule / contribution lists uploaded for this mode");
        case DC_LAUNCH_BAD_MODE: return fail(rc, "dc_cuda_assemble: unknown mode");
        case DC_LAUNCH_PARAMS: return fail(rc, "dc_cuda_assemble: wrong number of operator parameters");
        case DC_LAUNCH_SMEM: return fail(rc, "dc_cuda_assemble: unit does not fit in shared memory (lower unit_slots)");
        case DC_LAUNCH_NOT_SYMMETRIC:
            return fail(rc, "dc_cuda_assemble: symmetric schedule with a non-symmetric operator");
        default: return fail(rc, "dc_cuda_assemble: CUDA launch failed", cudaGetLastError());
    }
}

/* the built-in operators are instantiated here exactly the way a user operator is in its own file */
DC_DEVICE_OPERATOR(dc_builtin_laplacian, dcdev::OpLaplacian)
DC_DEVICE_OPERATOR(dc_builtin_elasticity, dcdev::OpElasticity)

extern "C" int dc_cuda_assemble_with(dc_cuda_ctx *c, dc_operator_launcher launcher, const double *h_params,
                                     int nbParams, int mode, double *d_values, void *stream)
{
    if (!c || !c->d_coord) return fail(-1, "dc_cuda_assemble: no mesh");
    if (!d_values || !launcher) return fail(-1, "dc_cuda_assemble: null argument");
    CK(cudaSetDevice(c->device));
    dc_launch_args a;
    fill_args(c, mode, &a);
    const int rc = launcher(&a, h_params, nbParams, d_values, stream);
    c->launches += a.launches;
    return explain(rc);
}

extern "C" int dc_cuda_assemble(dc_cuda_ctx *c, int op, const double *h_params, int nbParams, int mode,
                                double *d_values, void *stream)
{
    switch (op) {
        case DC_OP_P1_TET_LAPLACIAN:
            return dc_cuda_assemble_with(c, dc_builtin_laplacian, h_params, nbParams, mode, d_values, stream);
        case DC_OP_P1_TET_ELASTICITY:
            return dc_cuda_assemble_with(c, dc_builtin_elasticity, h_params, nbParams, mode, d_values, stream);
        default:
            return fail(-9, "dc_cuda_assemble: unknown operator");
    }
}

/* a range of the units (the pieces dc_cuda_assemble_exchange is made of) */
extern "C" int dc_cuda_assemble_units(dc_cuda_ctx *c, int op, const double *h_params, int nbParams, double *d_values,
                                      void *stream, int64_t unitFirst, int64_t unitCount, int reserveCtas, int skipBigRows)
{
    if (!c || !c->d_coord) return fail(-1, "dc_cuda_assemble_units: no mesh");
    if (!d_values || unitFirst < 0 || unitCount < 0 || unitFirst + unitCount > c->nbUnits)
        return fail(-1, "dc_cuda_assemble_units: bad argument");
    if (op != DC_OP_P1_TET_LAPLACIAN && op != DC_OP_P1_TET_ELASTICITY) return fail(-9, "dc_cuda_assemble_units: unknown operator");
    CK(cudaSetDevice(c->device));
    dc_launch_args a;
    fill_args(c, DC_MODE_DETERMINISTIC, &a);
    a.unitFirst = unitFirst; a.unitCount = unitCount; a.reserveCtas = reserveCtas; a.skipBigRows = skipBigRows;
    const int rc = (op == DC_OP_P1_TET_LAPLACIAN ? dc_builtin_laplacian : dc_builtin_elasticity)(&a, h_params, nbParams, d_values, stream);
    c->launches += a.launches;
    return explain(rc);
}

extern "C" int dc_cuda_set_front_units(dc_cuda_ctx *c, int64_t nbFrontUnits)
{
    if (!c || nbFrontUnits < 0 || nbFrontUnits > c->nbUnits) return fail(-1, "dc_cuda_set_front_units: bad argument");
    c->nbFrontUnits = nbFrontUnits;
    return 0;
}

/* front units (+ fallback rows) -> pack on the side stream || other units with a few CTA slots
 * left free -> unpack on the side stream -> join */
extern "C" int dc_cuda_assemble_exchange_with(dc_cuda_ctx *c, dc_cuda_exchange *x, dc_operator_launcher launcher,
                                              const double *h_params, int nbParams, double *d_values, void *stream)
{
    if (!c || !c->d_coord) return fail(-1, "dc_cuda_assemble_exchange: no mesh");
    if (!d_values || !launcher) return fail(-1, "dc_cuda_assemble_exchange: null argument");
    CK(cudaSetDevice(c->device));
    if (!x || dc_cuda_exchange_neighbours(x) == 0)
        return dc_cuda_assemble_with(c, launcher, h_params, nbParams, DC_MODE_DETERMINISTIC, d_values, stream);
    if (!c->sideStream) {
        /* LOWEST priority: when the front units finish, the interior launch and the pack kernel become
         * ready together; the interior CTAs must be placed first (two per SM, all registers) and the
         * pack / unpack blocks take the slots left free -- a pack block placed first would keep a whole
         * interior CTA off its SM, and with the static unit stride that CTA finishes late by as much */
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&c->sideStream, cudaStreamNonBlocking, lo));
        CK(cudaEventCreateWithFlags(&c->evFront, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t front = c->nbFrontUnits > 0 ? c->nbFrontUnits : c->nbUnits;
    /* The pack and unpack kernels run beside the interior units on the CTA slots the persistent grid
     * leaves free, where a gather is latency-bound (about 10 GB/s per slot), so the right number of slots
     * follows the bytes sent per interior element.  Weak scaling sends 0.2-0.3: four slots hide the
     * exchange completely (ms/step = 1.002 x kernel) and more only slow the interior down.  One mesh cut
     * over several GPUs sends 2-4 bytes per interior element to 3-7 neighbours: measured optimum 16 slots
     * at 2.3 (4 GPUs, 25 M tets each: 2.86 / 2.61 / 2.78 ms with 4 / 16 / 48) and 32 at 4.0 (8 GPUs, 12.6 M
     * tets each: 1.83 / 1.57 / 1.39 ms with 4 / 16 / 32), i.e. about 8 slots per byte (profiles/README.md).
     * Packing on the whole GPU BEFORE the interior launch (exchange_pack_early) measured slower in both. */
    int reserve = c->exchangeReserve;
    if (reserve < 0) {
        const double interiorElems = (double)c->nbElem * (double)(c->nbUnits - front) / (double)(c->nbUnits > 0 ? c->nbUnits : 1);
        const double bytesPerElem = (double)dc_cuda_exchange_bytes(x) * 0.5 / (interiorElems > 1.0 ? interiorElems : 1.0);
        reserve = (int)(8.0 * bytesPerElem + 0.5);
        reserve = (reserve + 3) & ~3;
        if (reserve < 4) reserve = 4;
        if (reserve > 48) reserve = 48;
    }
    const bool early = c->exchangePackEarly > 0;
    dc_launch_args a;
    fill_args(c, DC_MODE_DETERMINISTIC, &a);
    a.unitFirst = 0; a.unitCount = front;
    int rc = launcher(&a, h_params, nbParams, d_values, stream);
    c->launches += a.launches;
    if (rc) return explain(rc);
    if (early && dc_cuda_exchange_pack(x, d_values, s)) return fail(-15, dc_cuda_exchange_last_error());
    CK(cudaEventRecord(c->evFront, s));
    if (front < c->nbUnits) {            /* enqueued BEFORE a pack on the side stream: see above */
        fill_args(c, DC_MODE_DETERMINISTIC, &a);
        a.unitFirst = front; a.unitCount = c->nbUnits - front;
        a.reserveCtas = reserve; a.skipBigRows = 1;
        rc = launcher(&a, h_params, nbParams, d_values, stream);
        c->launches += a.launches;
        if (rc) return explain(rc);
    }
    CK(cudaStreamWaitEvent(c->sideStream, c->evFront, 0));
    if (!early && dc_cuda_exchange_pack(x, d_values, c->sideStream)) return fail(-15, dc_cuda_exchange_last_error());
    if (dc_cuda_exchange_unpack(x, d_values, c->sideStream)) return fail(-15, dc_cuda_exchange_last_error());
    c->launches += 2 * dc_cuda_exchange_neighbours(x);
    CK(cudaEventRecord(c->evJoin, c->sideStream));
    CK(cudaStreamWaitEvent(s, c->evJoin, 0));
    return 0;
}

extern "C" int dc_cuda_assemble_exchange(dc_cuda_ctx *c, dc_cuda_exchange *x, int op, const double *h_params, int nbParams,
                                         double *d_values, void *stream)
{
    switch (op) {
        case DC_OP_P1_TET_LAPLACIAN:
            return dc_cuda_assemble_exchange_with(c, x, dc_builtin_laplacian, h_params, nbParams, d_values, stream);
        case DC_OP_P1_TET_ELASTICITY:
            return dc_cuda_assemble_exchange_with(c, x, dc_builtin_elasticity, h_params, nbParams, d_values, stream);
        default:
            return fail(-9, "dc_cuda_assemble_exchange: unknown operator");
    }
}

extern "C" int dc_cuda_values(dc_cuda_ctx *c, int dim, double **d_values)
{
    if (!c || !d_values) return fail(-1, "dc_cuda_values: null argument");
    CK(cudaSetDevice(c->device));
    const int64_t need = c->nnz * dim * dim;
    if (need > c->valuesLen) {
        cudaFree(c->d_values);
        c->d_values = nullptr;
        CK(cudaMalloc((void **)&c->d_values, sizeof(double) * (size_t)(need > 0 ? need : 1)));
        c->valuesLen = need;
    }
    *d_values = c->d_values;
    return 0;
}

extern "C" int dc_cuda_download_values(dc_cuda_ctx *c, int dim, double *h_values, void *stream)
{
    if (!c || !c->d_values) return fail(-1, "dc_cuda_download_values: no values");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(h_values, c->d_values, sizeof(double) * (size_t)(c->nnz * dim * dim),
                       cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return 0;
}

extern "C" int dc_cuda_download_range(const double *d_values, int64_t first, int64_t count, double *h_buf,
                                      void *stream)
{
    if (count <= 0) return 0;
    CK(cudaMemcpyAsync(h_buf, d_values + first, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost,
                       (cudaStream_t)stream));
    return 0;
}

static unsigned grid_for(int64_t work, int perBlock)
{
    int64_t b = (work + perBlock - 1) / perBlock;
    const int64_t cap = 148 * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

extern "C" int dc_cuda_permute_double_2d(const double *d_in, double *d_out, const int *d_perm,
                                         int64_t nbItem, int dimItem, void *stream)
{
    if (nbItem <= 0) return 0;
    permute_rows<double><<<grid_for(nbItem * dimItem, 256), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, d_perm,
                                                                                           nbItem, dimItem);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int dc_cuda_permute_int_2d(const int *d_in, int *d_out, const int *d_perm, int64_t nbItem,
                                      int dimItem, void *stream)
{
    if (nbItem <= 0) return 0;
    const bool vec = dimItem == 4 && ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0;
    if (vec)
        permute_rows_int4<<<grid_for(nbItem, 256), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const int4 *>(d_in), reinterpret_cast<int4 *>(d_out), d_perm, nbItem);
    else
        permute_rows<int><<<grid_for(nbItem * dimItem, 256), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, d_perm,
                                                                                            nbItem, dimItem);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int dc_cuda_gather_double_2d(const double *d_in, double *d_out, const int *d_inv, int64_t nbItem, int dimItem,
                                        void *stream)
{
    if (nbItem <= 0) return 0;
    gather_rows<double><<<grid_for(nbItem * dimItem, 256), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, d_inv, nbItem, dimItem);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int dc_cuda_gather_int_2d(const int *d_in, int *d_out, const int *d_inv, int64_t nbItem, int dimItem, void *stream)
{
    if (nbItem <= 0) return 0;
    const bool vec = dimItem == 4 && ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0;
    if (vec)
        gather_rows_int4<<<grid_for((nbItem + 1) / 2, 256), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const int4 *>(d_in), reinterpret_cast<int4 *>(d_out), d_inv, nbItem);
    else
        gather_rows<int><<<grid_for(nbItem * dimItem, 256), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, d_inv, nbItem, dimItem);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int dc_cuda_renumber(int *d_tab, const int *d_perm, int64_t size, int isFortran, void *stream)
{
    if (size <= 0) return 0;
    renumber<<<grid_for(size / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(d_tab, d_perm, size, isFortran ? 1 : 0);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int dc_cuda_reset_edges(double *d_values, int64_t firstEdge, int64_t lastEdge, int perEdge,
                                   void *stream)
{
    if (lastEdge < firstEdge) return 0;
    const int64_t n = (lastEdge - firstEdge + 1) * perEdge;
    reset_values<<<grid_for(n / 2 + 1, 256), 256, 0, (cudaStream_t)stream>>>(d_values + firstEdge * perEdge, n);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int dc_cuda_pack_edges(const double *d_values, const int64_t *d_edge, int64_t nbEdges, int perEdge,
                                  double *d_buf, void *stream)
{
    if (nbEdges <= 0) return 0;
    pack_edges<<<grid_for(nbEdges * perEdge, 256), 256, 0, (cudaStream_t)stream>>>(d_values, d_edge, nbEdges,
                                                                                   perEdge, d_buf);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int dc_cuda_unpack_add_edges(double *d_values, const int64_t *d_edge, int64_t nbEdges, int perEdge,
                                        const double *d_buf, void *stream)
{
    if (nbEdges <= 0) return 0;
    unpack_add_edges<<<grid_for(nbEdges * perEdge, 256), 256, 0, (cudaStream_t)stream>>>(d_values, d_edge, nbEdges,
                                                                                         perEdge, d_buf);
    CK(cudaGetLastError());
    return 0;
}

/* y = A x with the assembled block-CSR values (dim = 1 or 3) and the context's CSR pattern */
extern "C" int dc_cuda_spmv(dc_cuda_ctx *c, int dim, const double *d_values, const double *d_x, double *d_y,
                            void *stream)
{
    if (!c || !c->d_row || !c->d_col) return fail(-1, "dc_cuda_spmv: no CSR pattern on the device");
    if (!d_values || !d_x || !d_y) return fail(-1, "dc_cuda_spmv: null argument");
    CK(cudaSetDevice(c->device));
    if (c->nbNodes <= 0) return 0;
    const unsigned grid = grid_for((int64_t)c->nbNodes * dim, 256);
    if (dim == 1)
        bsr_spmv<1><<<grid, 256, 0, (cudaStream_t)stream>>>(c->d_row, c->d_col, c->colBase, d_values, d_x, d_y, c->nbNodes);
    else if (dim == 3)
        bsr_spmv<3><<<grid, 256, 0, (cudaStream_t)stream>>>(c->d_row, c->d_col, c->colBase, d_values, d_x, d_y, c->nbNodes);
    else
        return fail(-9, "dc_cuda_spmv: dim must be 1 or 3");
    CK(cudaGetLastError());
    c->launches++;
    return 0;
}

/* inverse diagonal blocks of the assembled matrix (block-Jacobi preconditioner); *nbBad (host, may be NULL)
 * receives the number of rows without a usable diagonal block -- reading it synchronises the stream */
extern "C" int dc_cuda_block_jacobi(dc_cuda_ctx *c, int dim, const double *d_values, double *d_invDiag, int *nbBad,
                                    void *stream)
{
    if (!c || !c->d_row || !c->d_col) return fail(-1, "dc_cuda_block_jacobi: no CSR pattern on the device");
    if (!d_values || !d_invDiag) return fail(-1, "dc_cuda_block_jacobi: null argument");
    if (dim != 1 && dim != 3) return fail(-9, "dc_cuda_block_jacobi: dim must be 1 or 3");
    CK(cudaSetDevice(c->device));
    if (c->nbNodes <= 0) { if (nbBad) *nbBad = 0; return 0; }
    int *d_bad = nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaMallocAsync((void **)&d_bad, sizeof(int), s));
    CK(cudaMemsetAsync(d_bad, 0, sizeof(int), s));
    const unsigned grid = grid_for((int64_t)c->nbNodes, 256);
    if (dim == 1)
        block_jacobi<1><<<grid, 256, 0, s>>>(c->d_row, c->d_col, c->colBase, d_values, d_invDiag, c->nbNodes, d_bad);
    else
        block_jacobi<3><<<grid, 256, 0, s>>>(c->d_row, c->d_col, c->colBase, d_values, d_invDiag, c->nbNodes, d_bad);
    CK(cudaGetLastError());
    c->launches++;
    if (nbBad) {
        CK(cudaMemcpyAsync(nbBad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    CK(cudaFreeAsync(d_bad, s));
    return 0;
}

/* debug: enable per-unit phase stamps; the next assembly fills h_out[nbUnits][24] on request */
extern "C" int dc_cuda_debug_phases(dc_cuda_ctx *c, long long *h_out)
{
    if (!c) return fail(-1, "dc_cuda_debug_phases: null context");
    CK(cudaSetDevice(c->device));
    if (!c->d_phaseClock) {
        CK(cudaMalloc((void **)&c->d_phaseClock, sizeof(long long) * 24 * (size_t)(c->nbUnits > 0 ? c->nbUnits : 1)));
        return 0;
    }
    CK(cudaDeviceSynchronize());
    if (h_out)
        CK(cudaMemcpy(h_out, c->d_phaseClock, sizeof(long long) * 24 * (size_t)c->nbUnits, cudaMemcpyDeviceToHost));
    cudaFree(c->d_phaseClock);
    c->d_phaseClock = nullptr;
    return 0;
}

/* ---- CSR pattern / node -> element lists on the device (csrc/dc_csr.cuh) -------------------- */

struct dc_cuda_csr {
    const int *d_conn = nullptr;
    int nbElem = 0, nbNodes = 0;
    int64_t nnz = 0;
    int *d_deg = nullptr, *d_cnt = nullptr, *d_n2eVal = nullptr, *d_flag = nullptr;
    const int *d_row = nullptr;
    int64_t *d_n2ePtr = nullptr, *d_sums = nullptr;
};

static void csr_release(dc_cuda_csr *h)
{
    if (!h) return;
    cudaFree(h->d_deg); cudaFree(h->d_cnt); cudaFree(h->d_n2eVal); cudaFree(h->d_n2ePtr); cudaFree(h->d_sums);
    cudaFree(h->d_flag);
    delete h;
}

/* exclusive scan of d_in[n] (int32 counts) into d_out[n+1]; the total stays in d_sums[nbBlocks] */
template <typename OUT>
static int scan_counts(const int *d_in, int64_t n, int64_t *d_sums, OUT *d_out, cudaStream_t s)
{
    using namespace dcdev;
    const int64_t nbBlocks = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nbBlocks == 0) {
        CK(cudaMemsetAsync(d_out, 0, sizeof(OUT), s));
        CK(cudaMemsetAsync(d_sums, 0, sizeof(int64_t), s));
        return 0;
    }
    scan_block_sums<<<(unsigned)nbBlocks, 256, 0, s>>>(d_in, n, d_sums);
    scan_sums<<<1, 1024, 0, s>>>(d_sums, nbBlocks);
    scan_apply<OUT><<<(unsigned)nbBlocks, 256, 0, s>>>(d_in, n, d_sums, d_out);
    CK(cudaGetLastError());
    return 0;
}

static int csr_begin_body(dc_cuda_csr *h, int *d_row, cudaStream_t s)
{
    using namespace dcdev;
    const int nbNodes = h->nbNodes;
    const int64_t entries = 4 * (int64_t)h->nbElem;
    const int64_t nbBlocks = ((int64_t)nbNodes + SCAN_TILE - 1) / SCAN_TILE;
    int rc = 0, bad = 0;
    CK(cudaMalloc(&h->d_deg, sizeof(int) * (size_t)(nbNodes + 1)));
    CK(cudaMalloc(&h->d_cnt, sizeof(int) * (size_t)(nbNodes + 1)));
    CK(cudaMalloc(&h->d_n2ePtr, sizeof(int64_t) * (size_t)(nbNodes + 1)));
    CK(cudaMalloc(&h->d_n2eVal, sizeof(int) * (size_t)(entries ? entries : 1)));
    CK(cudaMalloc(&h->d_sums, sizeof(int64_t) * (size_t)(nbBlocks + 1)));
    CK(cudaMalloc(&h->d_flag, sizeof(int)));
    CK(cudaMemsetAsync(h->d_deg, 0, sizeof(int) * (size_t)(nbNodes + 1), s));
    CK(cudaMemsetAsync(h->d_cnt, 0, sizeof(int) * (size_t)(nbNodes + 1), s));
    CK(cudaMemsetAsync(h->d_flag, 0, sizeof(int), s));
    if (entries)
        csr_count_nodes<<<grid_for(entries, 256), 256, 0, s>>>(h->d_conn, entries, nbNodes, h->d_deg, h->d_flag);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&bad, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (bad) return fail(-2, "dc_cuda_csr_begin: the connectivity names a node outside [1, nbNodes]");
    if ((rc = scan_counts<int64_t>(h->d_deg, nbNodes, h->d_sums, h->d_n2ePtr, s))) return rc;
    /* d_cnt serves as the fill cursor first, then holds the edges per row */
    if (entries)
        csr_fill_n2e<<<grid_for(entries, 256), 256, 0, s>>>(h->d_conn, entries, h->d_n2ePtr, h->d_cnt, h->d_n2eVal);
    if (nbNodes)
        csr_rows<false><<<grid_for((int64_t)nbNodes * 32, 256), 256, 0, s>>>(h->d_conn, h->d_n2ePtr, h->d_n2eVal, nbNodes,
                                                                           h->d_cnt, nullptr, nullptr, 0);
    CK(cudaGetLastError());
    if ((rc = scan_counts<int>(h->d_cnt, nbNodes, h->d_sums, d_row, s))) return rc;
    CK(cudaMemcpyAsync(&h->nnz, h->d_sums + nbBlocks, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (h->nnz > INT32_MAX) return fail(-3, "dc_cuda_csr_begin: the pattern exceeds 2^31-1 edges");
    return 0;
}

extern "C" int dc_cuda_csr_begin(const int *d_conn, int nbElem, int nbNodes, int *d_row, int64_t *nnz,
                                 dc_cuda_csr **handle, void *stream)
{
    if ((!d_conn && nbElem > 0) || !d_row || !nnz || !handle || nbElem < 0 || nbNodes < 0)
        return fail(-1, "dc_cuda_csr_begin: bad argument");
    dc_cuda_csr *h = new dc_cuda_csr;
    h->d_conn = d_conn; h->nbElem = nbElem; h->nbNodes = nbNodes; h->d_row = d_row;
    const int rc = csr_begin_body(h, d_row, (cudaStream_t)stream);
    if (rc) { csr_release(h); return rc; }
    *nnz = h->nnz;
    *handle = h;
    return 0;
}

extern "C" int dc_cuda_csr_columns(dc_cuda_csr *h, int *d_col, int colBase, void *stream)
{
    using namespace dcdev;
    if (!h || (!d_col && h->nnz)) return fail(-1, "dc_cuda_csr_columns: bad argument");
    if (h->nbNodes)
        csr_rows<true><<<grid_for((int64_t)h->nbNodes * 32, 256), 256, 0, (cudaStream_t)stream>>>(
            h->d_conn, h->d_n2ePtr, h->d_n2eVal, h->nbNodes, nullptr, h->d_row, d_col, colBase);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int dc_cuda_csr_node_to_elem(dc_cuda_csr *h, int64_t *d_ptr, int *d_val, void *stream)
{
    using namespace dcdev;
    if (!h) return fail(-1, "dc_cuda_csr_node_to_elem: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (d_ptr)
        CK(cudaMemcpyAsync(d_ptr, h->d_n2ePtr, sizeof(int64_t) * (size_t)(h->nbNodes + 1), cudaMemcpyDeviceToDevice, s));
    if (d_val && h->nbNodes)
        csr_sort_n2e<<<grid_for((int64_t)h->nbNodes * 32, 256), 256, 0, s>>>(h->d_n2ePtr, h->nbNodes, h->d_n2eVal, d_val);
    CK(cudaGetLastError());
    return 0;
}

extern "C" void dc_cuda_csr_end(dc_cuda_csr *h) { csr_release(h); }

extern "C" int dc_cuda_set_option(dc_cuda_ctx *c, const char *name, int value)
{
    if (!c || !name) return fail(-1, "dc_cuda_set_option: null argument");
    if (!strcmp(name, "unit_threads")) {
        if (value != 128 && value != 256 && value != 512)
            return fail(-12, "dc_cuda_set_option: unit_threads must be 128, 256 or 512");
        c->threads = value;
        return 0;
    }
    if (!strcmp(name, "exchange_pack_early")) {
        if (value < 0 || value > 1) return fail(-12, "dc_cuda_set_option: exchange_pack_early must be 0 or 1");
        c->exchangePackEarly = value;
        return 0;
    }
    if (!strcmp(name, "exchange_reserve")) {
        if (value < -1 || value > 64) return fail(-12, "dc_cuda_set_option: exchange_reserve must be -1 (automatic) or 0..64");
        c->exchangeReserve = value;
        return 0;
    }
    if (!strcmp(name, "persistent")) {
        c->persistent = value ? 1 : 0;
        return 0;
    }
    return fail(-12, "dc_cuda_set_option: unknown option");
}

extern "C" int64_t dc_cuda_launch_count(const dc_cuda_ctx *c) { return c ? c->launches : 0; }

extern "C" int dc_cuda_sync(void *stream)
{
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
