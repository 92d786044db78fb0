/* dc_tree.cc -- D&C tree construction and finalisation (host, once per mesh).
 *
 * Produces the same tree, elemPerm and nodePerm as the reference's
 * DC_create_tree (tree_creation.cc:482-501 -> partitioning.cc:167-228 ->
 * tree_creation.cc:379-479 -> partitioning.cc:115-164), byte for byte, but in
 * O(n log n): the reference rescans the whole element permutation at every tree
 * node (permutations.cc:118-130) and builds partition orders with an
 * O(size*nbPart) scan (permutations.cc:169-180); here the inverse permutation
 * (position -> original element) travels with the connectivity rows, so every
 * tree node only touches its own range, and partition orders are counting sorts.
 *
 * What must be reproduced exactly for METIS to return the reference's partition:
 * the nodal graph's adjacency ORDER (partitioning.cc:38-85: neighbours listed by
 * ascending incident element, then element-local node order, first occurrence
 * wins) and the call arguments (ncon = 1, no weights, default options).
 */
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <unordered_map>
#include "dc_internal.h"

tree_t *treeHead = nullptr;
int *elemPerm = nullptr, *nodePerm = nullptr;

namespace dc {

Options &options()
{
    static Options o;
    return o;
}

tree_t *new_tree_node(int firstElem, int lastElem, int nbSepElem, int firstNode,
                      int lastNode, bool isSep, bool isLeaf)
{
    /* field values as set by the reference's init_dc_tree (tree_creation.cc:283-312) */
    tree_t *t = new tree_t;
    t->ownedNodes = t->intfIndex = t->intfNodes = t->intfDst = nullptr;
    t->nbOwnedNodes = t->nbIntfNodes = 0;
    t->firstElem = firstElem;
    t->lastElem  = lastElem - nbSepElem;
    t->lastSep   = lastElem;
    t->firstNode = firstNode;
    t->lastNode  = lastNode;
    t->firstEdge = t->lastEdge = -1;
    t->vecOffset = 0;
    t->isSep = isSep;
    t->left = t->right = t->sep = nullptr;
    (void)isLeaf;
    return t;
}

void free_subtree(tree_t *t)
{
    if (!t) return;
    free_subtree(t->left);
    free_subtree(t->right);
    free_subtree(t->sep);
    delete[] t->ownedNodes;
    delete[] t->intfIndex;
    delete[] t->intfNodes;
    delete[] t->intfDst;
    delete t;
}

/* Nodal graph of a P1 mesh in METIS CSR form, adjacency order as in the
 * reference's mesh_to_nodal (partitioning.cc:38-85). */
static void build_nodal_graph(std::vector<int64_t> &xadj, std::vector<int64_t> &adjncy,
                              const int *conn, int64_t nbElem, int dimElem, int nbNodes)
{
    std::vector<int64_t> ptr((size_t)nbNodes + 1, 0);
    for (int64_t k = 0; k < nbElem * dimElem; k++) ptr[(size_t)conn[k] + 1]++;
    for (int n = 0; n < nbNodes; n++) ptr[(size_t)n + 1] += ptr[n];
    std::vector<int> incident((size_t)ptr[nbNodes]);
    {
        std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
        for (int64_t e = 0; e < nbElem; e++)
            for (int j = 0; j < dimElem; j++)
                incident[(size_t)fill[conn[e * dimElem + j]]++] = (int)e;
    }
    xadj.assign((size_t)nbNodes + 1, 0);
    adjncy.clear();
    adjncy.reserve((size_t)nbNodes * 16);
    std::vector<int> seenBy((size_t)nbNodes, -1);
    for (int n = 0; n < nbNodes; n++) {
        seenBy[n] = n;
        for (int64_t k = ptr[n]; k < ptr[(size_t)n + 1]; k++) {
            const int *row = conn + (int64_t)incident[(size_t)k] * dimElem;
            for (int j = 0; j < dimElem; j++) {
                int m = row[j];
                if (seenBy[m] != n) {
                    seenBy[m] = n;
                    adjncy.push_back(m);
                }
            }
        }
        xadj[(size_t)n + 1] = (int64_t)adjncy.size();
    }
}

static void metis_recursive(std::vector<int> &part, std::vector<int64_t> &xadj,
                            std::vector<int64_t> &adjncy, int nbNodes, int nbPart)
{
    int64_t n = nbNodes, ncon = 1, np = nbPart, objval = 0;
    std::vector<int64_t> p((size_t)(nbNodes > 0 ? nbNodes : 1), 0);
    if (nbPart < 2 || nbNodes < 1) {
        /* one part: METIS leaves the vector untouched (the reference then reads
         * uninitialised memory); the only sensible answer is "everything in part 0" */
        part.assign((size_t)nbNodes, 0);
        return;
    }
    if (adjncy.empty()) adjncy.push_back(0);
    METIS_PartGraphRecursive(&n, &ncon, xadj.data(), adjncy.data(), nullptr, nullptr, nullptr,
                             &np, nullptr, nullptr, nullptr, &objval, p.data());
    part.resize((size_t)nbNodes);
    for (int i = 0; i < nbNodes; i++) part[i] = (int)p[i];
}

struct TreeBuilder {
    int *conn;        /* 0-based connectivity, rows follow the current element order */
    int *posToOrig;   /* original element index of the row at each position */
    int dimElem;
    int maxElem;

    void sep_partitioning(tree_t *&tree, int firstSepElem, int lastSepElem, int firstNode,
                          int lastNode);
    void tree_creation(tree_t *&tree, int *sepToNode, const int *nodePart,
                       const int *nodePartSize, int firstPart, int lastPart, int firstElem,
                       int lastElem, int firstNode, int lastNode, int sepOffset, bool isSep);
};

/* reference: tree_creation.cc:379-479 (recursion) + 352-375 (element classes) */
void TreeBuilder::tree_creation(tree_t *&tree, int *sepToNode, const int *nodePart,
                                const int *nodePartSize, int firstPart, int lastPart,
                                int firstElem, int lastElem, int firstNode, int lastNode,
                                int sepOffset, bool isSep)
{
    const int nbPart = lastPart - firstPart + 1;
    const int n = lastElem - firstElem + 1;
    if (nbPart < 2 || n <= maxElem) {
        tree = new_tree_node(firstElem, lastElem, 0, firstNode, lastNode, isSep, true);
        return;
    }
    const int pivot = firstPart + (lastPart - firstPart) / 2;

    int nbLeftNodes = 0, nbRightNodes = 0;
    if (!isSep) {
        for (int p = firstPart; p <= pivot; p++) nbLeftNodes += nodePartSize[p];
        nbRightNodes = (lastNode - firstNode + 1) - nbLeftNodes;
    }

    /* class 0: every node in a part <= pivot, 1: every node beyond it, 2: straddles */
    std::vector<int> cls((size_t)n), order((size_t)n);
    const int *rows = isSep ? sepToNode + (int64_t)sepOffset * dimElem
                            : conn + (int64_t)firstElem * dimElem;
    int nbLeft = 0, nbSep = 0;
    for (int e = 0; e < n; e++) {
        int onLeft = 0;
        for (int j = 0; j < dimElem; j++) onLeft += (nodePart[rows[(int64_t)e * dimElem + j]] <= pivot);
        int c = (onLeft == dimElem) ? 0 : (onLeft == 0 ? 1 : 2);
        cls[e] = c;
        nbLeft += (c == 0);
        nbSep  += (c == 2);
    }
    stable_partition_perm(order.data(), cls.data(), n, 3);

    scatter_rows_inplace<int>(conn + (int64_t)firstElem * dimElem, order.data(), n, dimElem);
    scatter_rows_inplace<int>(posToOrig + firstElem, order.data(), n, 1);
    if (isSep) scatter_rows_inplace<int>(sepToNode + (int64_t)sepOffset * dimElem, order.data(), n, dimElem);
    std::vector<int>().swap(cls);
    std::vector<int>().swap(order);

    tree = new_tree_node(firstElem, lastElem, nbSep, firstNode, lastNode, isSep, false);
    tree_creation(tree->right, sepToNode, nodePart, nodePartSize, pivot + 1, lastPart,
                  firstElem + nbLeft, lastElem - nbSep, firstNode + nbLeftNodes, lastNode,
                  sepOffset + nbLeft, isSep);
    tree_creation(tree->left, sepToNode, nodePart, nodePartSize, firstPart, pivot, firstElem,
                  firstElem + nbLeft - 1, firstNode, lastNode - nbRightNodes, sepOffset, isSep);
    if (nbSep > 0)
        sep_partitioning(tree->sep, lastElem - nbSep + 1, lastElem, firstNode, lastNode);
}

/* reference: partitioning.cc:115-164 (+ create_sepToNode 89-112) */
void TreeBuilder::sep_partitioning(tree_t *&tree, int firstSepElem, int lastSepElem,
                                   int firstNode, int lastNode)
{
    const int n = lastSepElem - firstSepElem + 1;
    int nbSepPart = (int)ceil(n / (double)maxElem);
    if (nbSepPart < 2 || n <= maxElem) {
        tree = new_tree_node(firstSepElem, lastSepElem, 0, firstNode, lastNode, true, true);
        return;
    }
    /* separator-local node numbers in order of first appearance */
    std::vector<int> sepToNode((size_t)n * dimElem);
    std::unordered_map<int, int> localId;
    localId.reserve((size_t)n * 2);
    const int *rows = conn + (int64_t)firstSepElem * dimElem;
    for (int64_t k = 0; k < (int64_t)n * dimElem; k++) {
        auto it = localId.find(rows[k]);
        if (it == localId.end()) it = localId.emplace(rows[k], (int)localId.size()).first;
        sepToNode[(size_t)k] = it->second;
    }
    const int nbSepNodes = (int)localId.size();

    std::vector<int64_t> xadj, adjncy;
    std::vector<int> nodePart;
    build_nodal_graph(xadj, adjncy, sepToNode.data(), n, dimElem, nbSepNodes);
    metis_recursive(nodePart, xadj, adjncy, nbSepNodes, nbSepPart);
    std::vector<int64_t>().swap(xadj);
    std::vector<int64_t>().swap(adjncy);

    tree_creation(tree, sepToNode.data(), nodePart.data(), nullptr, 0, nbSepPart - 1,
                  firstSepElem, lastSepElem, firstNode, lastNode, 0, true);
}

}  // namespace dc

/* reference: tree_creation.cc:482-501 + partitioning.cc:167-228 */
void DC_create_tree (int *elemToNode, int nbElem, int dimElem, int nbNodes)
{
    using namespace dc;
    DC_free_tree();
    elemPerm = new int[nbElem > 0 ? nbElem : 1];
    nodePerm = new int[nbNodes > 0 ? nbNodes : 1];
    const int maxElem = options().maxElemPerPart;
    const int64_t nbConn = (int64_t)nbElem * dimElem;

    #pragma omp parallel for schedule(static) if (nbConn > (1 << 16))
    for (int64_t k = 0; k < nbConn; k++) elemToNode[k]--;

    int nbPart = (int)ceil(nbElem / (double)maxElem);
    std::vector<int> nodePart;
    {
        std::vector<int64_t> xadj, adjncy;
        build_nodal_graph(xadj, adjncy, elemToNode, nbElem, dimElem, nbNodes);
        metis_recursive(nodePart, xadj, adjncy, nbNodes, nbPart);
    }
    stable_partition_perm(nodePerm, nodePart.data(), nbNodes, nbPart);
    std::vector<int> nodePartSize((size_t)(nbPart > 0 ? nbPart : 1), 0);
    for (int i = 0; i < nbNodes; i++) nodePartSize[nodePart[i]]++;

    std::vector<int> posToOrig((size_t)(nbElem > 0 ? nbElem : 1));
    for (int i = 0; i < nbElem; i++) posToOrig[i] = i;

    TreeBuilder b{elemToNode, posToOrig.data(), dimElem, maxElem};
    b.tree_creation(treeHead, nullptr, nodePart.data(), nodePartSize.data(), 0, nbPart - 1, 0,
                    nbElem - 1, 0, nbNodes - 1, 0, false);

    #pragma omp parallel for schedule(static) if (nbConn > (1 << 16))
    for (int64_t k = 0; k < nbConn; k++) elemToNode[k]++;

    if (options().vecSize > 0)
        color_leaves(elemToNode, nbElem, dimElem, nbNodes, options().vecSize, posToOrig.data());

    for (int pos = 0; pos < nbElem; pos++) elemPerm[posToOrig[pos]] = pos;
}

/* reference: tree_creation.cc:176-239 (BULK_SYNCHRONOUS, no STATS): every leaf,
 * separator leaves included, gets the CSR edge interval of its node interval. */
static void finalize_rec(tree_t *t, const int *nodeToNodeRow)
{
    if (!t) return;
    if (t->left == nullptr && t->right == nullptr) {
        t->firstEdge = nodeToNodeRow[t->firstNode];
        t->lastEdge  = nodeToNodeRow[t->lastNode + 1] - 1;
        return;
    }
    finalize_rec(t->right, nodeToNodeRow);
    finalize_rec(t->left, nodeToNodeRow);
    finalize_rec(t->sep, nodeToNodeRow);
}

void DC_finalize_tree (int *nodeToNodeRow, int *elemToNode, int *intfIndex, int *intfNodes,
                       int *intfDstOffsets, int *nbDCcomm, int nbElem, int dimElem,
                       int nbBlocks, int nbIntf, int rank)
{
    (void)elemToNode; (void)intfIndex; (void)intfNodes; (void)intfDstOffsets; (void)nbDCcomm;
    (void)nbElem; (void)dimElem; (void)nbBlocks; (void)nbIntf; (void)rank;
    if (!treeHead) dc::fatal("Error: DC_finalize_tree called without a D&C tree.");
    finalize_rec(treeHead, nodeToNodeRow);
}

void DC_finalize_tree (int *nodeToNodeRow, int *elemToNode, int nbNodes)
{
    (void)nbNodes;
    DC_finalize_tree(nodeToNodeRow, elemToNode, nullptr, nullptr, nullptr, nullptr, 0, 0, 1, 0, 0);
}

void DC_set_vec_size (int vecSize) { dc::options().vecSize = vecSize; }
int  DC_get_vec_size () { return dc::options().vecSize; }
void DC_set_max_elem_per_part (int maxElem) { if (maxElem > 0) dc::options().maxElemPerPart = maxElem; }
int  DC_get_max_elem_per_part () { return dc::options().maxElemPerPart; }
void DC_set_num_threads (int nbThreads) { dc::options().nbThreads = nbThreads; }
tree_t *DC_get_tree () { return treeHead; }

void DC_free_tree ()
{
    dc::free_subtree(treeHead);
    treeHead = nullptr;
    delete[] elemPerm; elemPerm = nullptr;
    delete[] nodePerm; nodePerm = nullptr;
}
