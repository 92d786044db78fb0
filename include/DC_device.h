/* DC_device.h -- device-operator entry points of dc-b200 (C++ and Fortran callable).
 *
 * The reference's assembly runs HOST callbacks on every leaf of the D&C tree
 * (DC_tree_traversal, src/tree_traversal.cc:106-117; README alias DC_assembly,
 * README.md:120-140).  Host function pointers cannot run on the GPU, so the
 * accelerated path registers a DEVICE operator instead: one of the built-ins
 * below, or a user functor compiled against dc-lib_b200/csrc/dc_device_operator.cuh.
 * Call order (the same as with the reference, Appendix A of SURVEY.md):
 *   DC_create_tree | DC_read_tree -> DC_permute_* / DC_renumber_int_array ->
 *   user builds the CSR pattern -> DC_finalize_tree -> DC_device_set_mesh ->
 *   per step: [DC_device_update_coords] DC_assembly_device.
 * Errors follow the reference's convention: message on stderr, exit(EXIT_FAILURE).
 */
#ifndef DC_DEVICE_H
#define DC_DEVICE_H

#include "DC.h"

/* built-in device operators */
#define DC_OP_LAPLACIAN   0   /* P1 tetrahedron, scalar Laplacian, operatorDim 1            */
#define DC_OP_ELASTICITY  1   /* P1 tetrahedron, isotropic elasticity, operatorDim 3,       */
                              /* opParams = {lambda, mu}                                     */
/* assembly modes */
#define DC_DEVICE_DETERMINISTIC 0   /* every CSR value summed in ascending element order:   */
                                    /* bit-identical to the reference traversal              */
#define DC_DEVICE_ATOMIC        1   /* reset + red.global.add.f64 scatter (order-free)       */

/* Select the CUDA device and create the library's device context. */
void DC_device_init (int device);
/* DC_create_tree (METIS node partition on the host, include/DC.h) and DC_create_tree_geometric with the
 * level-wise element passes of the tree proper -- create_elem_part, the stable 3-way order and the row
 * permutation of every tree node, src/tree_creation.cc:352-375, 415-430 -- on the device
 * (dc_cuda_tree_*, include/dc_cuda.h); separator subtrees and the node partition stay on the host.
 * Same contract, same bytes: connectivity order, elemPerm, nodePerm and tree equal the host builders'.
 * They run on the device of DC_device_init, or of DC_device_select when the tree is created before the
 * device context exists (device 0 by default). */
void DC_device_select (int device);
void DC_create_tree_device (int *elemToNode, int nbElem, int dimElem, int nbNodes);
void DC_create_tree_geometric_device (int *elemToNode, int nbElem, int dimElem, int nbNodes,
                                      const double *coord, int dimNode);
/* Upload mesh + CSR pattern and flatten the finalized tree into the device schedule.
 * elemToNode: 1-based, tree order, renumbered; coord: [nbNodes][3] permuted;
 * nodeToNodeColumn numbering base colBase (0 or 1). */
void DC_device_set_mesh (const double *coord, const int *elemToNode, const int *nodeToNodeRow,
                         const int *nodeToNodeColumn, int nbElem, int nbNodes, int colBase);
void DC_device_update_coords (const double *coord);
void DC_device_set_mode (int mode);
/* Shared-memory budget of a work unit in touched elements (default 512). */
void DC_device_set_unit_slots (int maxSlots);
/* Reset + assemble into the HOST array nodeToNodeValue[nnz][operatorDim^2]
 * (device assembly followed by one device->host copy). */
void DC_assembly_device (int op, const double *opParams, int nbParams, double *nodeToNodeValue,
                         int operatorDim);
/* Same without the copy; the values stay in device memory (pointer returned). */
double *DC_assembly_device_resident (int op, const double *opParams, int nbParams, int operatorDim);
/* Same with a user-defined device operator: `launcher` is the function a user's .cu file defines with
 * DC_DEVICE_OPERATOR(name, Functor) (dc-lib_b200/csrc/dc_device_operator.cuh) -- what replaces the
 * host function pointers userSeqFct / userVecFct of the reference (src/tree_traversal.cc:106-117). */
struct dc_launch_args;
typedef int (*DC_device_operator) (struct dc_launch_args *args, const double *opParams, int nbParams,
                                   double *deviceValues, void *stream);
void DC_assembly_device_with (DC_device_operator launcher, const double *opParams, int nbParams,
                              double *nodeToNodeValue, int operatorDim);
double *DC_assembly_device_resident_with (DC_device_operator launcher, const double *opParams, int nbParams,
                                          int operatorDim);
void DC_device_finalize ();

/* ---- several domains, one GPU each (one process per domain, as with the reference's MPI/GASPI users) ----
 * The interface of this domain in the reference's layout (DC_finalize_tree, include/DC.h:141-143): for
 * neighbour i the nodes intfNodes[intfIndex[i] .. intfIndex[i+1]) -- 1-based local ids AFTER
 * DC_renumber_int_array, listed in the same order by both domains -- shared with rank neighbourRank[i]
 * (ascending).  Call before DC_device_set_mesh, which then puts the work units that own interface rows
 * first and creates the peer-memory exchange (dc_cuda_exchange_*, include/dc_cuda.h).  Every domain hands
 * the 128-byte handle of its receive buffer for neighbour i to that neighbour (MPI_Sendrecv, a file, ...),
 * which connects it; from then on DC_assembly_device sums the interface entries across the domains, the
 * exchange overlapped with the interior units (the reference's MULTITHREADED_COMM, at unit granularity). */
void DC_device_set_interface (int nbIntf, const int *neighbourRank, const int *intfIndex, const int *intfNodes);
/* after DC_device_set_mesh: size the exchange buffers for the operator that will be assembled
 * (operatorDim^2 doubles per interface edge); then swap the handles */
void DC_device_prepare_exchange (int operatorDim);
void DC_device_exchange_handle (int intf, void *handle128);
void DC_device_exchange_connect (int intf, const void *handle128);

extern "C" {
    void dc_device_init_ (int *device);
    void dc_device_select_ (int *device);
    void dc_create_tree_device_ (int *elemToNode, int *nbElem, int *dimElem, int *nbNodes);
    void dc_create_tree_geometric_device_ (int *elemToNode, int *nbElem, int *dimElem, int *nbNodes,
                                           const double *coord, int *dimNode);
    void dc_device_set_mesh_ (const double *coord, const int *elemToNode, const int *nodeToNodeRow,
                              const int *nodeToNodeColumn, int *nbElem, int *nbNodes, int *colBase);
    void dc_device_update_coords_ (const double *coord);
    void dc_device_set_mode_ (int *mode);
    void dc_assembly_device_ (int *op, const double *opParams, int *nbParams,
                              double *nodeToNodeValue, int *operatorDim);
    void dc_device_finalize_ ();
    void dc_device_set_interface_ (int *nbIntf, const int *neighbourRank, const int *intfIndex, const int *intfNodes);
    void dc_device_prepare_exchange_ (int *operatorDim);
    void dc_device_exchange_handle_ (int *intf, void *handle128);
    void dc_device_exchange_connect_ (int *intf, const void *handle128);
    /* flatten the current tree in pre-order: 8 ints per tree node
     * {firstElem,lastElem,lastSep,firstNode,lastNode,isSep,isLeaf,hasSep};
     * returns the node count (call with out = NULL to size the buffer) */
    int64_t dc_flatten_tree (int *out, int64_t maxNodes);
    /* copies of the library-owned permutations (between create/read and store); 0 on success */
    int dc_get_node_perm (int *out, int nbNodes);
    int dc_get_elem_perm (int *out, int nbElem);
}

#endif
