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
void DC_device_finalize ();

extern "C" {
    void dc_device_init_ (int *device);
    void dc_device_set_mesh_ (const double *coord, const int *elemToNode, const int *nodeToNodeRow,
                              const int *nodeToNodeColumn, int *nbElem, int *nbNodes, int *colBase);
    void dc_device_update_coords_ (const double *coord);
    void dc_device_set_mode_ (int *mode);
    void dc_assembly_device_ (int *op, const double *opParams, int *nbParams,
                              double *nodeToNodeValue, int *operatorDim);
    void dc_device_finalize_ ();
    /* flatten the current tree in pre-order: 8 ints per tree node
     * {firstElem,lastElem,lastSep,firstNode,lastNode,isSep,isLeaf,hasSep};
     * returns the node count (call with out = NULL to size the buffer) */
    int64_t dc_flatten_tree (int *out, int64_t maxNodes);
    /* copies of the library-owned permutations (between create/read and store); 0 on success */
    int dc_get_node_perm (int *out, int nbNodes);
    int dc_get_elem_perm (int *out, int nbElem);
}

#endif
