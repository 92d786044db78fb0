/* DC_f.h -- Fortran-callable ABI of dc-b200 (reference: include/DC_f.h:23-79).
 * Same symbols and by-address calling convention as the reference's wrappers;
 * dc_assembly_, the option setters and the device entry points (DC_device.h) are
 * additions following the same convention. */
#ifndef DC_F_H
#define DC_F_H

#include "DC.h"

extern "C" {
    double   dc_get_time_ ();
    uint64_t dc_get_cycles_ ();

    void dc_tree_traversal_ (void (*userSeqFct)  (void *, DCargs_t *),
                             void (*userVecFct)  (void *, DCargs_t *),
                             void (*userCommFct) (void *, DCcommArgs_t *),
                             void *userArgs, void *userCommArgs);
    void dc_assembly_ (void (*userSeqFct) (void *, int, int),
                       void (*userVecFct) (void *, int, int),
                       void *userArgs, double *nodeToNodeValue, int *operatorDim);

    void dc_permute_double_2d_array_ (double *tab, int *nbItem, int *dimItem);
    void dc_permute_int_2d_array_ (int *tab, int *nbItem, int *dimItem, int *offset);
    void dc_permute_int_1d_array_ (int *tab, int *size);
    void dc_renumber_int_array_ (int *tab, int *size);
    void dc_create_permutation_ (int *perm, int *part, int *size, int *nbPart);

    void dc_read_tree_ (char *treePath, int *nbElem, int *nbNodes, int *nbIntf,
                        int *nbMaxComm);
    void dc_store_tree_ (char *treePath, int *nbElem, int *nbNodes, int *nbIntf,
                         int *nbMaxComm);
    void dc_finalize_tree_ (int *nodeToNodeRow, int *elemToNode, int *intfIndex,
                            int *intfNodes, int *intfDstOffsets, int *nbDCcomm,
                            int *nbElem, int *dimElem, int *nbBlocks, int *nbIntf,
                            int *rank);
    void dc_create_tree_ (int *elemToNode, int *nbElem, int *dimElem, int *nbNodes);

    /* additions */
    void dc_create_tree_geometric_ (int *elemToNode, int *nbElem, int *dimElem, int *nbNodes,
                                    double *coord, int *dimNode);
    void dc_partition_mesh_ (int *elemToNode, int *nbElem, int *dimElem, int *nbNodes, int *nbPart,
                             double *coord, int *dimNode, int *elemPart);
    void dc_set_vec_size_ (int *vecSize);
    void dc_set_max_elem_per_part_ (int *maxElem);
    void dc_set_num_threads_ (int *nbThreads);
    void dc_free_tree_ ();
}

#endif
