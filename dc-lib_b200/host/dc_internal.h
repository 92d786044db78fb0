/* dc_internal.h -- private declarations shared by the host-side sources. */
#ifndef DC_INTERNAL_H
#define DC_INTERNAL_H

#include <stdint.h>
#include <vector>
#include "DC.h"

/* Library-owned process state, same three globals as the reference
 * (tree_creation.cc:33-34). */
extern tree_t *treeHead;
extern int *elemPerm, *nodePerm;

namespace dc {

struct Options {
    int vecSize = 0;          /* 0 = pure D&C */
    int maxElemPerPart = MAX_ELEM_PER_PART;
    int nbThreads = 0;
};
Options &options();

[[noreturn]] void fatal(const char *msg);

/* METIS 5.x from the CUDA toolkit's libmetis_static.a: idx_t = int64, real_t = float. */
extern "C" int METIS_PartGraphRecursive(int64_t *nvtxs, int64_t *ncon, int64_t *xadj,
                                        int64_t *adjncy, int64_t *vwgt, int64_t *vsize,
                                        int64_t *adjwgt, int64_t *nparts, float *tpwgts,
                                        float *ubvec, int64_t *options, int64_t *objval,
                                        int64_t *part);

/* permutations (dc_perm.cc) */
void stable_partition_perm(int *perm, const int *part, int size, int nbPart);
template <typename T>
void scatter_rows_inplace(T *tab, const int *perm, int64_t nbItem, int dimItem);

/* colouring (dc_coloring.cc) */
void color_leaves(int *elemToNode, int nbElem, int dimElem, int nbNodes, int vecSize,
                  int *posToOrig);

/* DC_create_tree / DC_create_tree_geometric (coord != nullptr); device >= 0: the element passes of
 * the tree proper run on that CUDA device (dc_tree.cc) */
void create_tree_impl(int *elemToNode, int nbElem, int dimElem, int nbNodes, const double *coord,
                      int dimNode, int device);

/* tree helpers */
void free_subtree(tree_t *t);
tree_t *new_tree_node(int firstElem, int lastElem, int nbSepElem, int firstNode,
                      int lastNode, bool isSep, bool isLeaf);

}  // namespace dc
#endif
