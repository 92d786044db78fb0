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
/* interface of this domain (DC_device_set_interface), consumed by DC_device_set_mesh */
std::vector<int> g_intfRank, g_intfIndex, g_intfNodes;
dc_cuda_exchange *g_exchange = nullptr;
int g_exchangeDim = 0;                       /* operatorDim the exchange buffers were sized for */
std::vector<int64_t> g_intfEdgePtr, g_intfEdges;
int g_device = 0;

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
    if (g_exchange) { dc_cuda_exchange_destroy(g_exchange); g_exchange = nullptr; }
    if (g_ctx) { dc_cuda_destroy(g_ctx); g_ctx = nullptr; }
    if (dc_cuda_create(&g_ctx, device)) die("DC_device_init failed");
    g_device = device;
}

void DC_device_select (int device) { g_device = device; }

/* DC_create_tree with the element passes of the tree proper on the device (dc_cuda_tree_*): same
 * contract and the same bytes (connectivity, elemPerm, nodePerm, tree) as the host entry points.
 * Runs on the device of DC_device_init / DC_device_select (device 0 before the first call of either). */
void DC_create_tree_device (int *elemToNode, int nbElem, int dimElem, int nbNodes)
{
    dc::create_tree_impl(elemToNode, nbElem, dimElem, nbNodes, nullptr, 3, g_device);
}

void DC_create_tree_geometric_device (int *elemToNode, int nbElem, int dimElem, int nbNodes,
                                      const double *coord, int dimNode)
{
    if (!coord) dc::fatal("Error: DC_create_tree_geometric_device needs node coordinates.");
    dc::create_tree_impl(elemToNode, nbElem, dimElem, nbNodes, coord, dimNode, g_device);
}

void DC_device_set_interface (int nbIntf, const int *neighbourRank, const int *intfIndex, const int *intfNodes)
{
    g_intfRank.assign(neighbourRank, neighbourRank + nbIntf);
    g_intfIndex.assign(intfIndex, intfIndex + nbIntf + 1);
    g_intfNodes.assign(intfNodes, intfNodes + (nbIntf > 0 ? intfIndex[nbIntf] : 0));
}

void DC_device_exchange_handle (int intf, void *handle128)
{
    if (!g_exchange || dc_cuda_exchange_export(g_exchange, intf, handle128)) {
        fprintf(stderr, "Error: DC_device_exchange_handle: %s\n", g_exchange ? dc_cuda_exchange_last_error() : "no interface set");
        exit(EXIT_FAILURE);
    }
}

void DC_device_exchange_connect (int intf, const void *handle128)
{
    if (!g_exchange || dc_cuda_exchange_import(g_exchange, intf, handle128)) {
        fprintf(stderr, "Error: DC_device_exchange_connect: %s\n", g_exchange ? dc_cuda_exchange_last_error() : "no interface set");
        exit(EXIT_FAILURE);
    }
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
    /* several domains: interface edges from the node lists, interface units first, exchange buffers */
    int64_t nbFront = 0;
    if (g_exchange) { dc_cuda_exchange_destroy(g_exchange); g_exchange = nullptr; }
    const int nbIntf = (int)g_intfRank.size();
    std::vector<int64_t> edgePtr((size_t)nbIntf + 1, 0), edges;
    if (nbIntf > 0) {
        std::vector<unsigned char> flag((size_t)nbNodes, 0);
        for (int i = 0; i < nbIntf; i++) {
            const int *nodes = g_intfNodes.data() + g_intfIndex[i];
            const int cnt = g_intfIndex[i + 1] - g_intfIndex[i];
            const int64_t ne = dc_interface_edges(nodeToNodeRow, nodeToNodeColumn, colBase, nodes, cnt, nullptr, nullptr);
            edges.resize((size_t)(edgePtr[i] + ne));
            dc_interface_edges(nodeToNodeRow, nodeToNodeColumn, colBase, nodes, cnt, edges.data() + edgePtr[i], nullptr);
            edgePtr[(size_t)i + 1] = edgePtr[i] + ne;
            for (int p = 0; p < cnt; p++) flag[(size_t)nodes[p] - 1] = 1;
        }
        nbFront = dc_plan_order_units(&plan, flag.data());
    }
    if (dc_cuda_upload_plan(g_ctx, &plan)) die("DC_device_set_mesh: schedule upload failed");
    dc_plan_free(&plan);
    if (nbIntf > 0) {
        if (dc_cuda_set_front_units(g_ctx, nbFront)) die("DC_device_set_mesh: front units");
        /* per-edge words follow the operator; both built-ins are created lazily at the first assembly */
    }
    g_intfEdgePtr = edgePtr;
    g_intfEdges = edges;
}

void DC_device_update_coords (const double *coord)
{
    if (!g_ctx || dc_cuda_update_coords(g_ctx, coord, nullptr)) die("DC_device_update_coords failed");
}

/* the exchange is created when the first assembly tells how many doubles an edge carries */
static void ensure_exchange(int operatorDim)
{
    if (g_intfRank.empty() || (g_exchange && g_exchangeDim == operatorDim)) return;
    if (g_exchange) dc_cuda_exchange_destroy(g_exchange);
    g_exchange = nullptr;
    if (dc_cuda_exchange_create(&g_exchange, g_device, operatorDim * operatorDim, (int)g_intfRank.size(), g_intfRank.data(),
                                g_intfEdgePtr.data(), g_intfEdges.data())) {
        fprintf(stderr, "Error: DC_device: cannot create the interface exchange (%s)\n", dc_cuda_exchange_last_error());
        exit(EXIT_FAILURE);
    }
    g_exchangeDim = operatorDim;
}

/* DC_device_set_interface users create the exchange explicitly (the handles must be swapped before
 * the first assembly): operatorDim of the operator that will be assembled */
void DC_device_prepare_exchange (int operatorDim) { ensure_exchange(operatorDim); }

static double *assemble_resident(int op, DC_device_operator launcher, const double *opParams, int nbParams, int operatorDim)
{
    if (!g_ctx) dc::fatal("Error: DC_assembly_device called before DC_device_set_mesh.");
    double *d_values = nullptr;
    if (dc_cuda_values(g_ctx, operatorDim, &d_values)) die("DC_assembly_device: cannot allocate values");
    const int mode = (g_mode == DC_DEVICE_ATOMIC) ? DC_MODE_ATOMIC : DC_MODE_DETERMINISTIC;
    int rc;
    if (!g_intfRank.empty()) {
        if (mode != DC_MODE_DETERMINISTIC) dc::fatal("Error: DC_assembly_device: several domains need the deterministic mode.");
        ensure_exchange(operatorDim);
        rc = launcher ? dc_cuda_assemble_exchange_with(g_ctx, g_exchange, (dc_operator_launcher)launcher, opParams, nbParams, d_values, nullptr)
                      : dc_cuda_assemble_exchange(g_ctx, g_exchange, op, opParams, nbParams, d_values, nullptr);
    } else {
        rc = launcher ? dc_cuda_assemble_with(g_ctx, (dc_operator_launcher)launcher, opParams, nbParams, mode, d_values, nullptr)
                      : dc_cuda_assemble(g_ctx, op, opParams, nbParams, mode, d_values, nullptr);
    }
    if (rc) die("DC_assembly_device failed");
    return d_values;
}

double *DC_assembly_device_resident (int op, const double *opParams, int nbParams, int operatorDim)
{
    const int wantDim = (op == DC_OP_ELASTICITY) ? 3 : 1;
    if (operatorDim != wantDim) dc::fatal("Error: operatorDim does not match the device operator.");
    return assemble_resident(op, nullptr, opParams, nbParams, operatorDim);
}

double *DC_assembly_device_resident_with (DC_device_operator launcher, const double *opParams, int nbParams, int operatorDim)
{
    if (!launcher) dc::fatal("Error: DC_assembly_device_with: null operator.");
    return assemble_resident(-1, launcher, opParams, nbParams, operatorDim);
}

void DC_assembly_device (int op, const double *opParams, int nbParams, double *nodeToNodeValue,
                         int operatorDim)
{
    DC_assembly_device_resident(op, opParams, nbParams, operatorDim);
    if (dc_cuda_download_values(g_ctx, operatorDim, nodeToNodeValue, nullptr) || dc_cuda_sync(nullptr))
        die("DC_assembly_device: device->host copy failed");
}

void DC_assembly_device_with (DC_device_operator launcher, const double *opParams, int nbParams, double *nodeToNodeValue,
                              int operatorDim)
{
    DC_assembly_device_resident_with(launcher, opParams, nbParams, operatorDim);
    if (dc_cuda_download_values(g_ctx, operatorDim, nodeToNodeValue, nullptr) || dc_cuda_sync(nullptr))
        die("DC_assembly_device_with: device->host copy failed");
}

void DC_device_finalize ()
{
    if (g_exchange) dc_cuda_exchange_destroy(g_exchange);
    g_exchange = nullptr;
    if (g_ctx) dc_cuda_destroy(g_ctx);
    g_ctx = nullptr;
    g_intfRank.clear(); g_intfIndex.clear(); g_intfNodes.clear();
}

extern "C" {
void dc_device_init_ (int *device) { DC_device_init(*device); }
void dc_device_select_ (int *device) { DC_device_select(*device); }
void dc_create_tree_device_ (int *elemToNode, int *nbElem, int *dimElem, int *nbNodes)
{
    DC_create_tree_device(elemToNode, *nbElem, *dimElem, *nbNodes);
}
void dc_create_tree_geometric_device_ (int *elemToNode, int *nbElem, int *dimElem, int *nbNodes,
                                       const double *coord, int *dimNode)
{
    DC_create_tree_geometric_device(elemToNode, *nbElem, *dimElem, *nbNodes, coord, *dimNode);
}
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
void dc_device_set_interface_ (int *nbIntf, const int *neighbourRank, const int *intfIndex, const int *intfNodes)
{
    DC_device_set_interface(*nbIntf, neighbourRank, intfIndex, intfNodes);
}
void dc_device_prepare_exchange_ (int *operatorDim) { DC_device_prepare_exchange(*operatorDim); }
void dc_device_exchange_handle_ (int *intf, void *handle128) { DC_device_exchange_handle(*intf, handle128); }
void dc_device_exchange_connect_ (int *intf, const void *handle128) { DC_device_exchange_connect(*intf, handle128); }
}
