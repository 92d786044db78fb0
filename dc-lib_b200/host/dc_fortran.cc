/* dc_fortran.cc -- Fortran ABI (reference: src/fortran_wrapper.cc:64-146,
 * include/DC_f.h): extern "C", lower case, trailing underscore, every scalar by
 * address, NUL-terminated path strings.  dc_assembly_ and the device entry
 * points are additions in the same style. */
#include <string>
#include "DC_f.h"

extern "C" {

double   dc_get_time_ ()   { return DC_get_time(); }
uint64_t dc_get_cycles_ () { return DC_get_cycles(); }

void dc_tree_traversal_ (void (*userSeqFct)(void *, DCargs_t *), void (*userVecFct)(void *, DCargs_t *),
                         void (*userCommFct)(void *, DCcommArgs_t *), void *userArgs,
                         void *userCommArgs)
{
    DC_tree_traversal(userSeqFct, userVecFct, userCommFct, userArgs, userCommArgs);
}

void dc_assembly_ (void (*userSeqFct)(void *, int, int), void (*userVecFct)(void *, int, int),
                   void *userArgs, double *nodeToNodeValue, int *operatorDim)
{
    DC_assembly(userSeqFct, userVecFct, userArgs, nodeToNodeValue, *operatorDim);
}

void dc_permute_double_2d_array_ (double *tab, int *nbItem, int *dimItem)
{
    DC_permute_double_2d_array(tab, *nbItem, *dimItem);
}

/* the Fortran form always applies the element permutation */
void dc_permute_int_2d_array_ (int *tab, int *nbItem, int *dimItem, int *offset)
{
    DC_permute_int_2d_array(tab, nullptr, *nbItem, *dimItem, *offset);
}

void dc_permute_int_1d_array_ (int *tab, int *size) { DC_permute_int_1d_array(tab, *size); }

/* values are Fortran (1-based) node numbers */
void dc_renumber_int_array_ (int *tab, int *size) { DC_renumber_int_array(tab, *size, true); }

void dc_create_permutation_ (int *perm, int *part, int *size, int *nbPart)
{
    DC_create_permutation(perm, part, *size, *nbPart);
}

void dc_read_tree_ (char *treePath, int *nbElem, int *nbNodes, int *nbIntf, int *nbMaxComm)
{
    std::string path(treePath);
    DC_read_tree(path, *nbElem, *nbNodes, *nbIntf, nbMaxComm);
}

void dc_store_tree_ (char *treePath, int *nbElem, int *nbNodes, int *nbIntf, int *nbMaxComm)
{
    std::string path(treePath);
    DC_store_tree(path, *nbElem, *nbNodes, *nbIntf, *nbMaxComm);
}

void dc_finalize_tree_ (int *nodeToNodeRow, int *elemToNode, int *intfIndex, int *intfNodes,
                        int *intfDstOffsets, int *nbDCcomm, int *nbElem, int *dimElem,
                        int *nbBlocks, int *nbIntf, int *rank)
{
    DC_finalize_tree(nodeToNodeRow, elemToNode, intfIndex, intfNodes, intfDstOffsets, nbDCcomm,
                     *nbElem, *dimElem, *nbBlocks, *nbIntf, *rank);
}

void dc_create_tree_ (int *elemToNode, int *nbElem, int *dimElem, int *nbNodes)
{
    DC_create_tree(elemToNode, *nbElem, *dimElem, *nbNodes);
}

void dc_create_tree_geometric_ (int *elemToNode, int *nbElem, int *dimElem, int *nbNodes, double *coord,
                               int *dimNode)
{
    DC_create_tree_geometric(elemToNode, *nbElem, *dimElem, *nbNodes, coord, *dimNode);
}

void dc_partition_mesh_ (int *elemToNode, int *nbElem, int *dimElem, int *nbNodes, int *nbPart, double *coord,
                         int *dimNode, int *elemPart)
{
    DC_partition_mesh(elemToNode, *nbElem, *dimElem, *nbNodes, *nbPart, coord, *dimNode, elemPart);
}

void dc_set_vec_size_ (int *vecSize) { DC_set_vec_size(*vecSize); }
void dc_set_max_elem_per_part_ (int *maxElem) { DC_set_max_elem_per_part(*maxElem); }
void dc_set_num_threads_ (int *nbThreads) { DC_set_num_threads(*nbThreads); }
void dc_free_tree_ () { DC_free_tree(); }

}
