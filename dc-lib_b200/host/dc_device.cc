/* dc_device.cc -- C++/Fortran front end of the device assembly (DC_device.h):
 * glue between the library-owned tree (treeHead) and the C ABI in dc_cuda.h.
 * No computation happens here and nothing falls back to the CPU. */
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "dc_internal.h"
#include "DC_device.h"
#include "dc_cuda.h"
#include "../csrc/dc_plan.h"

namespace {

dc_cuda_ctx *g_ctx = nullptr;
int g_mode = DC_DEVICE_DETERMINISTIC;
int g_maxSlots = 512;     /* measured best with 256-thread CTAs, two per SM (profiles/README.md) */
int g_symmetric = 1;     /* both built-in operators have bitwise symmetric element matrices */

[[noreturn]] void die(const char *what)
{
    fprintf(stderr, "Error: %s (%s)\n", what, dc_cuda_last_error());
    exit(EXIT_FAILURE);
}

void flatten(const tree_t *t, std::vector<int> &out)
{
    const bool leaf = (t->left == nullptr && t->right == nullptr);
    const int rec[8] = {t->firstElem, t->lastElem, t->lastSep, t->firstNode, t->lastNode,
                        t->isSep ? 1 : 0, leaf ? 1 : 0, t->sep ? 1 : 0};
    out.insert(out.end(), rec, rec + 8);
    if (leaf) return;
    flatten(t->left, out);
    flatten(t->right, out);
    if (t->sep) flatten(t->sep, out);
}

}  // namespace

extern "C" int64_t dc_flatten_tree (int *out, int64_t maxNodes)
{
    if (!treeHead) return 0;
    std::vector<int> rec;
    flatten(treeHead, rec);
    const int64_t n = (int64_t)rec.size() / 8;
    if (out && n <= maxNodes) std::copy(rec.begin(), rec.end(), out);
    return n;
}

/* copies of the library-owned permutations (valid between create/read and store) */
extern "C" int dc_get_node_perm (int *out, int nbNodes)
{
    if (!nodePerm) return -1;
    std::copy(nodePerm, nodePerm + nbNodes, out);
    return 0;
}
extern "C" int dc_get_elem_perm (int *out, int nbElem)
{
    if (!elemPerm) return -1;
    std::copy(elemPerm, elemPerm + nbElem, out);
    return 0;
}

void DC_device_init (int device)
{
    if (g_ctx) { dc_cuda_destroy(g_ctx); g_ctx = nullptr; }
    if (dc_cuda_create(&g_ctx, device)) die("DC_device_init failed");
}

void DC_device_set_mode (int mode) { g_mode = mode; }
void DC_device_set_unit_slots (int maxSlots) { g_maxSlots = maxSlots; }

void DC_device_set_mesh (const double *coord, const int *elemToNode, const int *nodeToNodeRow,
                         const int *nodeToNodeColumn, int nbElem, int nbNodes, int colBase)
{
    if (!g_ctx) DC_device_init(0);
    if (!treeHead) dc::fatal("Error: DC_device_set_mesh called without a D&C tree.");
    if (dc_cuda_set_mesh(g_ctx, elemToNode, coord, nodeToNodeRow, nodeToNodeColumn, nbElem, nbNodes, colBase))
        die("DC_device_set_mesh: upload failed");
    std::vector<int> rec;
    flatten(treeHead, rec);
    dc_plan_t plan;
    int rc = dc_plan_build(&plan, elemToNode, nbElem, nbNodes, nodeToNodeRow, nodeToNodeColumn, colBase,
                           rec.data(), (int64_t)rec.size() / 8, g_maxSlots, dc::options().nbThreads, g_symmetric);
    if (rc) {
        fprintf(stderr, "Error: DC_device_set_mesh: cannot build the device schedule (code %d).\n", rc);
        exit(EXIT_FAILURE);
    }
    if (dc_cuda_upload_plan(g_ctx, &plan)) die("DC_device_set_mesh: schedule upload failed");
    dc_plan_free(&plan);
}

void DC_device_update_coords (const double *coord)
{
    if (!g_ctx || dc_cuda_update_coords(g_ctx, coord, nullptr)) die("DC_device_update_coords failed");
}

double *DC_assembly_device_resident (int op, const double *opParams, int nbParams, int operatorDim)
{
    if (!g_ctx) dc::fatal("Error: DC_assembly_device called before DC_device_set_mesh.");
    const int wantDim = (op == DC_OP_ELASTICITY) ? 3 : 1;
    if (operatorDim != wantDim) dc::fatal("Error: operatorDim does not match the device operator.");
    double *d_values = nullptr;
    if (dc_cuda_values(g_ctx, operatorDim, &d_values)) die("DC_assembly_device: cannot allocate values");
    const int mode = (g_mode == DC_DEVICE_ATOMIC) ? DC_MODE_ATOMIC : DC_MODE_DETERMINISTIC;
    if (dc_cuda_assemble(g_ctx, op, opParams, nbParams, mode, d_values, nullptr)) die("DC_assembly_device failed");
    return d_values;
}

void DC_assembly_device (int op, const double *opParams, int nbParams, double *nodeToNodeValue,
                         int operatorDim)
{
    DC_assembly_device_resident(op, opParams, nbParams, operatorDim);
    if (dc_cuda_download_values(g_ctx, operatorDim, nodeToNodeValue, nullptr) || dc_cuda_sync(nullptr))
        die("DC_assembly_device: device->host copy failed");
}

void DC_device_finalize ()
{
    if (g_ctx) dc_cuda_destroy(g_ctx);
    g_ctx = nullptr;
}

extern "C" {
void dc_device_init_ (int *device) { DC_device_init(*device); }
void dc_device_set_mesh_ (const double *coord, const int *elemToNode, const int *nodeToNodeRow,
                          const int *nodeToNodeColumn, int *nbElem, int *nbNodes, int *colBase)
{
    DC_device_set_mesh(coord, elemToNode, nodeToNodeRow, nodeToNodeColumn, *nbElem, *nbNodes, *colBase);
}
void dc_device_update_coords_ (const double *coord) { DC_device_update_coords(coord); }
void dc_device_set_mode_ (int *mode) { DC_device_set_mode(*mode); }
void dc_assembly_device_ (int *op, const double *opParams, int *nbParams, double *nodeToNodeValue,
                          int *operatorDim)
{
    DC_assembly_device(*op, opParams, *nbParams, nodeToNodeValue, *operatorDim);
}
void dc_device_finalize_ () { DC_device_finalize(); }
}
