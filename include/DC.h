/* DC.h -- public C++ interface of dc-b200, a B200-native drop-in for DC-lib's
 * assembly path.
 *
 * Every declaration in the first half of this header has the same name, argument
 * list and meaning as the corresponding declaration of the reference's
 * include/DC.h (reference: include/DC.h:23-146), so a program written against
 * DC-lib re-links against libdc_b200.so unchanged.  The record layouts of
 * tree_t / DCargs_t / DCcommArgs_t / index_t / list_t / couple_t are kept
 * field-for-field because user callbacks receive pointers to them.
 *
 * The second half adds (a) the README-generation entry points that the reference
 * documents but no longer ships (README.md:48-140) and (b) run-time options that
 * replace the reference's compile-time variants (build/CMakeLists.txt:31-58).
 * The device (CUDA) entry points live in DC_device.h.
 */
#ifndef DC_H
#define DC_H

#include <stdint.h>
#include <string>

/* Tunables the reference fixes at compile time (include/DC.h:23-24). */
#define MAX_ELEM_NEIGHBORS 250
#define MAX_ELEM_PER_PART 200

/* One node of the D&C tree.  Element / node / edge ranges are 0-based and
 * inclusive.  Inner nodes own [firstElem,lastElem] = left+right subtrees and
 * (lastElem,lastSep] = separator subtree; leaves have lastSep == lastElem. */
typedef struct tree_s {
    int *ownedNodes, *intfIndex, *intfNodes, *intfDst;
    int nbOwnedNodes, nbIntfNodes,
        firstElem, lastElem, lastSep,
        firstNode, lastNode,
        firstEdge, lastEdge,
        vecOffset;
    bool isSep;
    struct tree_s *left, *right, *sep;
} tree_t;

/* What a leaf callback receives (reference: include/DC.h:39-48). */
typedef struct DCargs_s {
#ifdef MULTITHREADED_COMM
    int *ownedNodes;
    int nbOwnedNodes;
#endif
    int firstElem, lastElem,
        firstNode, lastNode,
        firstEdge, lastEdge,
        isSep;
} DCargs_t;

/* Per-leaf interface slices for the user's halo exchange (include/DC.h:51-53). */
typedef struct DCcommArgs_s {
    int *intfIndex, *intfNodes, *intfDst, *intfOffset;
} DCcommArgs_t;

typedef struct { int *index, *value; } index_t;
typedef struct { int list[MAX_ELEM_NEIGHBORS]; int size; } list_t;
typedef struct { int elem, node; } couple_t;

/* Running-average wall-clock / TSC timer (reference: include/DC.h:67-93). */
class DC_timer
{
    public:
        DC_timer ();
        double   get_avg_time ();
        void     reset_time ();
        void     stop_time ();
        void     start_time ();
        uint64_t get_avg_cycles ();
        void     reset_cycles ();
        void     stop_cycles ();
        void     start_cycles ();
    private:
        double   startTime, avgTime;
        int      timeCtr;
        uint64_t startCycles, avgCycles;
        int      cyclesCtr;
};

double   DC_get_time ();
uint64_t DC_get_cycles ();

/* ---- shipped-generation API (reference: include/DC.h:102-146) ------------- */

/* Walk the D&C tree and run HOST callbacks on every leaf (left || right, then
 * separator).  This is the compatibility entry point for arbitrary host loops;
 * the accelerated assembly is DC_assembly_device() in DC_device.h. */
void DC_tree_traversal (void (*userSeqFct)  (void *, DCargs_t *),
                        void (*userVecFct)  (void *, DCargs_t *),
                        void (*userCommFct) (void *, DCcommArgs_t *),
                        void *userArgs, void *userCommArgs);

void DC_create_elemToElem (list_t *elemToElem, index_t &nodeToElem, int *elemToNode,
                           int firstElem, int lastElem, int dimElem);
void DC_create_nodeToElem (index_t &nodeToElem, int *elemToNode, int nbElem,
                           int dimElem, int nbNodes);

/* new[perm[i]] = old[i], in place.  The double/1-D forms use the node
 * permutation, the 2-D int form uses `perm` (elemPerm when perm == nullptr). */
void DC_permute_double_2d_array (double *tab, int nbItem, int dimItem);
void DC_permute_int_2d_array (int *tab, int *perm, int nbItem, int dimItem, int offset);
void DC_permute_int_1d_array (int *tab, int size);
void DC_renumber_int_array (int *tab, int size, bool isFortran);
void DC_create_permutation (int *perm, int *part, int size, int nbPart);

void DC_read_tree (std::string &treePath, int nbElem, int nbNodes, int nbIntf,
                   int *nbMaxComm);
void DC_store_tree (std::string &treePath, int nbElem, int nbNodes, int nbIntf,
                    int nbMaxComm);
void DC_finalize_tree (int *nodeToNodeRow, int *elemToNode, int *intfIndex,
                       int *intfNodes, int *intfDstOffsets, int *nbDCcomm, int nbElem,
                       int dimElem, int nbBlocks, int nbIntf, int rank);
void DC_create_tree (int *elemToNode, int nbElem, int dimElem, int nbNodes);

/* ---- README-generation API (reference: README.md:48-140) ------------------ */

/* Host-callback assembly: the library zeroes nodeToNodeValue over each
 * non-separator leaf's edge interval (operatorDim*operatorDim doubles per edge),
 * then calls fn(userArgs, firstElem, lastElem) on the leaf.  With a vector
 * length set, userVecFct gets [firstElem, vecOffset] and userSeqFct the rest. */
void DC_assembly (void (*userSeqFct) (void *, int, int),
                  void (*userVecFct) (void *, int, int),
                  void *userArgs, double *nodeToNodeValue, int operatorDim);
void DC_finalize_tree (int *nodeToNodeRow, int *elemToNode, int nbNodes);
void DC_store_tree (std::string &treePath, int nbElem, int nbNodes);
void DC_read_tree (std::string &treePath, int nbElem, int nbNodes);

/* ---- run-time options (replace the reference's -D variants) --------------- */

/* vecSize 0 = pure D&C (default); 2/4/8 = D&C + bounded colouring with that
 * many elements per colour (reference: -DDC_VEC -DVEC_SIZE=n). */
void DC_set_vec_size (int vecSize);
int  DC_get_vec_size ();
/* Leaf capacity used by DC_create_tree (reference constant MAX_ELEM_PER_PART). */
void DC_set_max_elem_per_part (int maxElem);
int  DC_get_max_elem_per_part ();
/* DC_create_tree with recursive coordinate bisection of the nodes (coord: [nbNodes][dimNode],
 * original numbering) in place of METIS -- same tree invariants, usable at 10^8 elements. */
void DC_create_tree_geometric (int *elemToNode, int nbElem, int dimElem, int nbNodes,
                               const double *coord, int dimNode);
/* Element partition for nbPart devices (BASELINE config 4 "METIS-split"): nodes cut
 * by METIS recursive bisection of the nodal graph as in partitioning.cc:167-228 (coord == nullptr)
 * or by coordinate bisection; an element follows the majority of its nodes.  elemPart: [nbElem]. */
void DC_partition_mesh (const int *elemToNode, int nbElem, int dimElem, int nbNodes, int nbPart,
                        const double *coord, int dimNode, int *elemPart);
/* Worker threads for DC_tree_traversal / DC_assembly host callbacks (0 = all). */
void DC_set_num_threads (int nbThreads);
/* Root of the current tree (nullptr before create/read). */
tree_t *DC_get_tree ();
/* Release the tree and the permutations. */
void DC_free_tree ();

#endif
