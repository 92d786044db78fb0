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
#include <memory>
#include <algorithm>
#include <omp.h>
#include "dc_internal.h"
#include "dc_cuda.h"

tree_t *treeHead = nullptr;
int *elemPerm = nullptr, *nodePerm = nullptr;

namespace dc {

/* The reference fixes the vector length at compile time (-DDC_VEC -DVEC_SIZE=n, one .so per
 * variant, build/CMakeLists.txt:87-109); here it is a run-time option.  A program that is only
 * re-linked (no DC_set_vec_size call) selects the variant with the environment instead:
 * DC_VEC_SIZE=n, DC_MAX_ELEM_PER_PART=m. */
Options &options()
{
    static Options o = [] {
        Options init;
        if (const char *v = getenv("DC_VEC_SIZE")) init.vecSize = atoi(v) > 0 ? atoi(v) : 0;
        if (const char *v = getenv("DC_MAX_ELEM_PER_PART")) init.maxElemPerPart = atoi(v) > 0 ? atoi(v) : init.maxElemPerPart;
        return init;
    }();
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

/* Recursive coordinate bisection into nbPart parts, numbered so that the parts of the
 * left half of every bisection are [first, first + (last-first)/2] -- the same midpoint
 * rule tree_creation uses to split a part range, hence the same shape of tree that a
 * recursive-bisection METIS partition gives.  Used where METIS is too slow (10^8 elements);
 * not part of the reference. */
struct Rcb {
    const double *coord;
    int dimNode;
    const int *ids;      /* local node -> row of coord (nullptr: identity) */
    int *part;

    inline double key(int local, int axis) const
    {
        return coord[(int64_t)(ids ? ids[local] : local) * dimNode + axis];
    }
    void split(int *idx, int64_t n, int firstPart, int lastPart)
    {
        if (firstPart == lastPart || n <= 0) {
            for (int64_t i = 0; i < n; i++) part[idx[i]] = firstPart;
            return;
        }
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int64_t i = 0; i < n; i++)
            for (int a = 0; a < dimNode && a < 3; a++) {
                double v = key(idx[i], a);
                lo[a] = v < lo[a] ? v : lo[a];
                hi[a] = v > hi[a] ? v : hi[a];
            }
        int axis = 0;
        for (int a = 1; a < dimNode && a < 3; a++)
            if (hi[a] - lo[a] > hi[axis] - lo[axis]) axis = a;
        const int pivot = firstPart + (lastPart - firstPart) / 2;
        const int64_t nbParts = lastPart - firstPart + 1, leftParts = pivot - firstPart + 1;
        int64_t nLeft = (n * leftParts + nbParts / 2) / nbParts;
        if (nLeft < 0) nLeft = 0;
        if (nLeft > n) nLeft = n;
        if (nLeft > 0 && nLeft < n)
            std::nth_element(idx, idx + nLeft, idx + n, [&](int p, int q) {
                double kp = key(p, axis), kq = key(q, axis);
                return kp < kq || (kp == kq && p < q);
            });
        #pragma omp task default(shared) if (n > 50000)
        split(idx, nLeft, firstPart, pivot);
        #pragma omp task default(shared) if (n > 50000)
        split(idx + nLeft, n - nLeft, pivot + 1, lastPart);
        #pragma omp taskwait
    }
};

/* tab[perm[i]] <- tab[i] (rows of dim ints) from inside a task region; src = scratch for n*dim
 * ints.  The scratch comes from buffers allocated once per tree (TreeBuilder): fresh gigabyte
 * allocations at every level of the recursion cost more in page faults than the copy itself. */
static void task_scatter_rows(int *tab, const int *perm, int n, int dim, int *src)
{
    const int chunk = 1 << 16, nbChunks = (n + chunk - 1) / chunk;
    #pragma omp taskloop default(shared) grainsize(1) if (nbChunks > 4)
    for (int c = 0; c < nbChunks; c++) {
        const int64_t e0 = (int64_t)c * chunk, e1 = std::min<int64_t>(n, e0 + chunk);
        memcpy(src + e0 * dim, tab + e0 * dim, sizeof(int) * (size_t)((e1 - e0) * dim));
    }
    #pragma omp taskloop default(shared) grainsize(1) if (nbChunks > 4)
    for (int c = 0; c < nbChunks; c++) {
        const int e0 = c * chunk, e1 = std::min(n, e0 + chunk);
        for (int e = e0; e < e1; e++)
            for (int j = 0; j < dim; j++) tab[(int64_t)perm[e] * dim + j] = src[(size_t)((int64_t)e * dim + j)];
    }
}

struct TreeBuilder {
    int *conn;        /* 0-based connectivity, rows follow the current element order */
    int *posToOrig;   /* original element index of the row at each position */
    int dimElem;
    int maxElem;
    const double *coord = nullptr;   /* node coordinates (original numbering): geometric mode */
    int dimNode = 3;
    /* scratch indexed by element POSITION (subtrees work on disjoint position ranges) */
    int *scrRows = nullptr;          /* [nbElem][dimElem] copy of the rows being scattered */
    int *scrOrder = nullptr;         /* [nbElem] destination of every element of a range */
    unsigned char *scrCls = nullptr; /* [nbElem] class of every element of a range */

    /* node partition of a (sub)mesh: METIS on the nodal graph (reference behaviour) or
     * coordinate bisection when coordinates were supplied */
    void partition_nodes(std::vector<int> &part, const int *localConn, int64_t nbElem, int nbLocalNodes,
                         int nbPart, const int *localToGlobal)
    {
        if (coord) {
            part.assign((size_t)nbLocalNodes, 0);
            std::vector<int> idx((size_t)nbLocalNodes);
            for (int i = 0; i < nbLocalNodes; i++) idx[i] = i;
            Rcb r{coord, dimNode, localToGlobal, part.data()};
            r.split(idx.data(), nbLocalNodes, 0, nbPart - 1);
            return;
        }
        std::vector<int64_t> xadj, adjncy;
        build_nodal_graph(xadj, adjncy, localConn, nbElem, dimElem, nbLocalNodes);
        #pragma omp critical(dc_metis)
        metis_recursive(part, xadj, adjncy, nbLocalNodes, nbPart);
    }

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
    int nbLeftOut = 0, nbSepOut = 0;

    int nbLeftNodes = 0, nbRightNodes = 0;
    if (!isSep) {
        for (int p = firstPart; p <= pivot; p++) nbLeftNodes += nodePartSize[p];
        nbRightNodes = (lastNode - firstNode + 1) - nbLeftNodes;
    }

    /* class 0: every node in a part <= pivot, 1: every node beyond it, 2: straddles.
     * The stable 3-way order (reference: DC_create_permutation with 3 parts) is computed
     * chunk-wise so that large ranges near the root are processed by all threads. */
    int *order = scrOrder + firstElem;                      /* every entry is written below */
    const int *rows = isSep ? sepToNode + (int64_t)sepOffset * dimElem
                            : conn + (int64_t)firstElem * dimElem;
    const int chunk = 1 << 16;
    const int nbChunks = (n + chunk - 1) / chunk;
    std::vector<int> hist((size_t)nbChunks * 3, 0);
    {
        unsigned char *cls = scrCls + firstElem;
        #pragma omp taskloop default(shared) grainsize(1) if (nbChunks > 4)
        for (int c = 0; c < nbChunks; c++) {
            const int e0 = c * chunk, e1 = std::min(n, e0 + chunk);
            int h[3] = {0, 0, 0};
            for (int e = e0; e < e1; e++) {
                int onLeft = 0;
                for (int j = 0; j < dimElem; j++) onLeft += (nodePart[rows[(int64_t)e * dimElem + j]] <= pivot);
                const int k = (onLeft == dimElem) ? 0 : (onLeft == 0 ? 1 : 2);
                cls[(size_t)e] = (unsigned char)k;
                h[k]++;
            }
            hist[(size_t)c * 3] = h[0]; hist[(size_t)c * 3 + 1] = h[1]; hist[(size_t)c * 3 + 2] = h[2];
        }
        int total[3] = {0, 0, 0};
        for (int c = 0; c < nbChunks; c++)
            for (int k = 0; k < 3; k++) total[k] += hist[(size_t)c * 3 + k];
        int base[3] = {0, total[0], total[0] + total[1]};
        for (int c = 0; c < nbChunks; c++)
            for (int k = 0; k < 3; k++) {
                const int cnt = hist[(size_t)c * 3 + k];
                hist[(size_t)c * 3 + k] = base[k];
                base[k] += cnt;
            }
        #pragma omp taskloop default(shared) grainsize(1) if (nbChunks > 4)
        for (int c = 0; c < nbChunks; c++) {
            const int e0 = c * chunk, e1 = std::min(n, e0 + chunk);
            int next[3] = {hist[(size_t)c * 3], hist[(size_t)c * 3 + 1], hist[(size_t)c * 3 + 2]};
            for (int e = e0; e < e1; e++) order[(size_t)e] = next[cls[(size_t)e]]++;
        }
        nbLeftOut = total[0];
        nbSepOut = total[2];
    }
    const int nbLeft = nbLeftOut, nbSep = nbSepOut;
    if (nbSep == n) {
        /* every element straddles the cut (a clump around a few nodes, seen with coordinate bisection of
         * separators of irregular parts): cutting again would recurse on the same set for ever -- the
         * reference has no guard either (tree_creation.cc:415-478) -- so the set stays one leaf; the
         * schedule builder cuts over-sized leaves by rows */
        tree = new_tree_node(firstElem, lastElem, 0, firstNode, lastNode, isSep, true);
        return;
    }

    int *scratch = scrRows + (int64_t)firstElem * dimElem;
    task_scatter_rows(conn + (int64_t)firstElem * dimElem, order, n, dimElem, scratch);
    task_scatter_rows(posToOrig + firstElem, order, n, 1, scratch);
    if (isSep) task_scatter_rows(sepToNode + (int64_t)sepOffset * dimElem, order, n, dimElem, scratch);

    tree = new_tree_node(firstElem, lastElem, nbSep, firstNode, lastNode, isSep, false);
    /* the three subtrees work on disjoint element ranges: build them concurrently (the
     * reference builds the separator after its siblings, but nothing it reads is written
     * by them) */
    #pragma omp task default(shared) if (n > 4096)
    tree_creation(tree->right, sepToNode, nodePart, nodePartSize, pivot + 1, lastPart,
                  firstElem + nbLeft, lastElem - nbSep, firstNode + nbLeftNodes, lastNode,
                  sepOffset + nbLeft, isSep);
    #pragma omp task default(shared) if (n > 4096)
    tree_creation(tree->left, sepToNode, nodePart, nodePartSize, firstPart, pivot, firstElem,
                  firstElem + nbLeft - 1, firstNode, lastNode - nbRightNodes, sepOffset, isSep);
    if (nbSep > 0)
        sep_partitioning(tree->sep, lastElem - nbSep + 1, lastElem, firstNode, lastNode);
    #pragma omp taskwait
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
    std::vector<int> sepToNode((size_t)n * dimElem), localToGlobal;
    /* open-addressing table node -> local id (at most 4n keys, load <= 1/2) */
    size_t cap = 16;
    while (cap < (size_t)n * dimElem * 2) cap <<= 1;
    std::vector<int> key(cap, -1), val(cap);
    const size_t hmask = cap - 1;
    const int *rows = conn + (int64_t)firstSepElem * dimElem;
    for (int64_t k = 0; k < (int64_t)n * dimElem; k++) {
        const int node = rows[k];
        size_t h = ((uint32_t)node * 2654435761u >> 5) & hmask;
        while (key[h] != node && key[h] != -1) h = (h + 1) & hmask;
        if (key[h] == -1) {
            key[h] = node;
            val[h] = (int)localToGlobal.size();
            localToGlobal.push_back(node);
        }
        sepToNode[(size_t)k] = val[h];
    }
    const int nbSepNodes = (int)localToGlobal.size();

    std::vector<int> nodePart;
    partition_nodes(nodePart, sepToNode.data(), n, nbSepNodes, nbSepPart, localToGlobal.data());

    tree_creation(tree, sepToNode.data(), nodePart.data(), nullptr, 0, nbSepPart - 1,
                  firstSepElem, lastSepElem, firstNode, lastNode, 0, true);
}

/* The tree proper (every node that is not inside a separator subtree) level by level, the element
 * passes of a level on the device (dc_cuda_tree_*, csrc/dc_cuda_tree.cu): what tree_creation() does
 * for isSep = false, with the recursion turned into a breadth-first loop so that one device pass
 * serves every tree node of a level.  The bookkeeping (tree nodes, part ranges, node ranges, the leaf
 * rule, the all-straddling guard) is the recursion's, line for line; the separator ranges are
 * collected and built afterwards on the host, concurrently (they are disjoint element ranges).
 * reference: tree_creation.cc:379-479 */
struct SplitSeg { tree_t **slot; int firstPart, lastPart, firstElem, lastElem, firstNode, lastNode; };
struct SepJob { tree_t **slot; int firstElem, lastElem, firstNode, lastNode; };

static void main_tree_on_device(TreeBuilder &b, int device, int nbElem, int nbNodes, const int *nodePart,
                                const int *nodePartSize, int nbPart, std::vector<SepJob> &jobs, int64_t *launches)
{
    auto is_leaf = [&](int parts, int n) { return parts < 2 || n <= b.maxElem; };
    if (is_leaf(nbPart, nbElem)) {
        treeHead = new_tree_node(0, nbElem - 1, 0, 0, nbNodes - 1, false, true);
        return;
    }
    std::vector<int64_t> partPrefix((size_t)nbPart + 1, 0);        /* nodes in parts [0, p) */
    for (int p = 0; p < nbPart; p++) partPrefix[(size_t)p + 1] = partPrefix[(size_t)p] + nodePartSize[p];
    auto check = [&](int rc, const char *what) {
        if (rc == 0) return;
        fprintf(stderr, "Error: %s (%s)\n", what, dc_cuda_tree_last_error());
        exit(EXIT_FAILURE);
    };
    dc_cuda_tree *t = nullptr;
    check(dc_cuda_tree_begin(&t, device, b.conn, nbElem, b.dimElem, nodePart, nbNodes), "DC_create_tree on the device: begin");
    std::vector<SplitSeg> cur{SplitSeg{&treeHead, 0, nbPart - 1, 0, nbElem - 1, 0, nbNodes - 1}}, next;
    std::vector<int> segFirst, segEnd, pivot, totals, child;
    while (!cur.empty()) {
        const size_t ns = cur.size();
        segFirst.resize(ns); segEnd.resize(ns); pivot.resize(ns); totals.resize(3 * ns); child.assign(3 * ns, -1);
        for (size_t s = 0; s < ns; s++) {
            segFirst[s] = cur[s].firstElem;
            segEnd[s] = cur[s].lastElem + 1;
            pivot[s] = cur[s].firstPart + (cur[s].lastPart - cur[s].firstPart) / 2;
        }
        check(dc_cuda_tree_classify(t, (int)ns, segFirst.data(), segEnd.data(), pivot.data(), totals.data()),
              "DC_create_tree on the device: classify");
        next.clear();
        for (size_t s = 0; s < ns; s++) {
            const SplitSeg g = cur[s];
            const int n = g.lastElem - g.firstElem + 1;
            const int nbLeft = totals[3 * s], nbSep = totals[3 * s + 2];
            if (nbSep == n) {                         /* every element straddles the cut: stays one leaf */
                *g.slot = new_tree_node(g.firstElem, g.lastElem, 0, g.firstNode, g.lastNode, false, true);
                continue;
            }
            const int nbLeftNodes = (int)(partPrefix[(size_t)pivot[s] + 1] - partPrefix[(size_t)g.firstPart]);
            const int nbRightNodes = (g.lastNode - g.firstNode + 1) - nbLeftNodes;
            tree_t *node = new_tree_node(g.firstElem, g.lastElem, nbSep, g.firstNode, g.lastNode, false, false);
            *g.slot = node;
            const SplitSeg kids[2] = {
                SplitSeg{&node->left, g.firstPart, pivot[s], g.firstElem, g.firstElem + nbLeft - 1, g.firstNode,
                         g.lastNode - nbRightNodes},
                SplitSeg{&node->right, pivot[s] + 1, g.lastPart, g.firstElem + nbLeft, g.lastElem - nbSep,
                         g.firstNode + nbLeftNodes, g.lastNode}};
            for (int k = 0; k < 2; k++) {
                const SplitSeg &c = kids[k];
                if (is_leaf(c.lastPart - c.firstPart + 1, c.lastElem - c.firstElem + 1)) {
                    *c.slot = new_tree_node(c.firstElem, c.lastElem, 0, c.firstNode, c.lastNode, false, true);
                } else {
                    child[3 * s + (size_t)k] = (int)next.size();
                    next.push_back(c);
                }
            }
            if (nbSep > 0) jobs.push_back(SepJob{&node->sep, g.lastElem - nbSep + 1, g.lastElem, g.firstNode, g.lastNode});
        }
        check(dc_cuda_tree_scatter(t, child.data()), "DC_create_tree on the device: scatter");
        cur.swap(next);
    }
    check(dc_cuda_tree_end(t, b.conn, b.posToOrig, launches), "DC_create_tree on the device: copy back");
}

/* reference: tree_creation.cc:482-501 + partitioning.cc:167-228.  device >= 0: the element passes of
 * the tree proper run on that CUDA device (DC_create_tree_device, DC_device.h). */
void create_tree_impl(int *elemToNode, int nbElem, int dimElem, int nbNodes, const double *coord,
                      int dimNode, int device)
{
    DC_free_tree();
    elemPerm = new int[nbElem > 0 ? nbElem : 1];
    nodePerm = new int[nbNodes > 0 ? nbNodes : 1];
    const int maxElem = options().maxElemPerPart;
    const int64_t nbConn = (int64_t)nbElem * dimElem;
    const int nt = options().nbThreads > 0 ? options().nbThreads : omp_get_max_threads();

    #pragma omp parallel for schedule(static) if (nbConn > (1 << 16))
    for (int64_t k = 0; k < nbConn; k++) elemToNode[k]--;

    int nbPart = (int)ceil(nbElem / (double)maxElem);
    std::vector<int> posToOrig((size_t)(nbElem > 0 ? nbElem : 1));
    for (int i = 0; i < nbElem; i++) posToOrig[i] = i;
    TreeBuilder b{elemToNode, posToOrig.data(), dimElem, maxElem, coord, dimNode};
    std::unique_ptr<int[]> scrRows(new int[(size_t)(nbConn > 0 ? nbConn : 1)]), scrOrder(new int[(size_t)(nbElem > 0 ? nbElem : 1)]);
    std::unique_ptr<unsigned char[]> scrCls(new unsigned char[(size_t)(nbElem > 0 ? nbElem : 1)]);
    b.scrRows = scrRows.get();
    b.scrOrder = scrOrder.get();
    b.scrCls = scrCls.get();
    /* first touch with one contiguous piece per thread: the chunk-wise tasks of the recursion would
     * fault the same huge pages from several threads at once, which serialises in the kernel */
    #pragma omp parallel for schedule(static) num_threads(nt) if (nbConn > (1 << 20))
    for (int64_t k = 0; k < nbConn; k += 1024) {
        scrRows[(size_t)k] = 0;
        if (k < nbElem) { scrOrder[(size_t)k] = 0; scrCls[(size_t)k] = 0; }
    }

    const bool verbose = getenv("DC_PLAN_VERBOSE") != nullptr;
    double t0 = omp_get_wtime();
    auto lap = [&](const char *what) {
        if (verbose) { double t1 = omp_get_wtime(); fprintf(stderr, "[dc_tree] %-12s %.2fs\n", what, t1 - t0); t0 = t1; }
    };
    std::vector<int> nodePart;
    #pragma omp parallel num_threads(nt)
    #pragma omp single
    b.partition_nodes(nodePart, elemToNode, nbElem, nbNodes, nbPart, nullptr);
    lap("partition");
    stable_partition_perm(nodePerm, nodePart.data(), nbNodes, nbPart);
    std::vector<int> nodePartSize((size_t)(nbPart > 0 ? nbPart : 1), 0);
    for (int i = 0; i < nbNodes; i++) nodePartSize[nodePart[i]]++;

    if (device >= 0) {
        std::vector<SepJob> jobs;
        int64_t launches = 0;
        main_tree_on_device(b, device, nbElem, nbNodes, nodePart.data(), nodePartSize.data(), nbPart, jobs, &launches);
        lap("device split");
        if (verbose) fprintf(stderr, "[dc_tree] %zu separator subtrees, %lld kernel launches\n", jobs.size(), (long long)launches);
        /* separator subtrees: disjoint element ranges, built concurrently */
        #pragma omp parallel num_threads(nt)
        #pragma omp single
        {
            for (size_t j = 0; j < jobs.size(); j++) {
                const SepJob job = jobs[j];
                if (job.lastElem - job.firstElem + 1 <= maxElem) {
                    b.sep_partitioning(*job.slot, job.firstElem, job.lastElem, job.firstNode, job.lastNode);
                } else {
                    #pragma omp task default(shared) firstprivate(job)
                    b.sep_partitioning(*job.slot, job.firstElem, job.lastElem, job.firstNode, job.lastNode);
                }
            }
            #pragma omp taskwait
        }
        lap("separators");
    } else {
        #pragma omp parallel num_threads(nt)
        #pragma omp single
        b.tree_creation(treeHead, nullptr, nodePart.data(), nodePartSize.data(), 0, nbPart - 1, 0,
                        nbElem - 1, 0, nbNodes - 1, 0, false);
        lap("recursion");
    }
    #pragma omp parallel for schedule(static) if (nbConn > (1 << 16))
    for (int64_t k = 0; k < nbConn; k++) elemToNode[k]++;

    if (options().vecSize > 0)
        color_leaves(elemToNode, nbElem, dimElem, nbNodes, options().vecSize, posToOrig.data());

    #pragma omp parallel for schedule(static) if (nbElem > (1 << 16))
    for (int pos = 0; pos < nbElem; pos++) elemPerm[posToOrig[pos]] = pos;
}

}  // namespace dc

void DC_create_tree (int *elemToNode, int nbElem, int dimElem, int nbNodes)
{
    dc::create_tree_impl(elemToNode, nbElem, dimElem, nbNodes, nullptr, 3, -1);
}

/* Same contract as DC_create_tree, with the node partitions cut by recursive coordinate
 * bisection of `coord` ([nbNodes][dimNode], ORIGINAL node numbering) instead of METIS.
 * For meshes where METIS is too slow; the resulting tree obeys the same invariants. */
void DC_create_tree_geometric (int *elemToNode, int nbElem, int dimElem, int nbNodes,
                               const double *coord, int dimNode)
{
    if (!coord) dc::fatal("Error: DC_create_tree_geometric needs node coordinates.");
    dc::create_tree_impl(elemToNode, nbElem, dimElem, nbNodes, coord, dimNode, -1);
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
    /* The interface arguments only matter to the reference's MULTITHREADED_COMM variant
     * (tree_creation.cc:48-169: per-leaf slices of the MPI interface); this library behaves like the
     * BULK_SYNCHRONOUS variant, which ignores them too (tree_creation.cc:176-202) -- the interface
     * exchange of the device path is dc_cuda_exchange_* (include/dc_cuda.h).  A caller re-linked from a
     * MULTITHREADED_COMM build is told so once instead of silently getting no per-leaf communication. */
    (void)elemToNode; (void)intfIndex; (void)intfNodes; (void)intfDstOffsets;
    (void)nbElem; (void)dimElem; (void)nbBlocks; (void)rank;
    if (nbDCcomm) *nbDCcomm = 0;
    if (nbIntf > 0 && intfIndex) {
        static bool told = false;
        if (!told) {
            told = true;
            fprintf(stderr, "DC_finalize_tree: %d interfaces given; this library is bulk-synchronous (no per-leaf interface "
                            "lists, nbDCcomm = 0): exchange after the assembly, or use dc_cuda_exchange_* on the device path.\n", nbIntf);
        }
    }
    if (!treeHead) dc::fatal("Error: DC_finalize_tree called without a D&C tree.");
    finalize_rec(treeHead, nodeToNodeRow);
}

void DC_finalize_tree (int *nodeToNodeRow, int *elemToNode, int nbNodes)
{
    (void)nbNodes;
    DC_finalize_tree(nodeToNodeRow, elemToNode, nullptr, nullptr, nullptr, nullptr, 0, 0, 1, 0, 0);
}

/* Element partition of a mesh for nbPart devices: the NODES are cut into nbPart parts -- METIS
 * recursive bisection of the nodal graph, exactly as DC_create_tree cuts them (reference:
 * partitioning.cc:167-228), or recursive coordinate bisection when coord is given -- and an element
 * goes to the part that holds most of its nodes (ties: the smallest part).  elemToNode is 1-based
 * and is not modified. */
void DC_partition_mesh (const int *elemToNode, int nbElem, int dimElem, int nbNodes, int nbPart,
                        const double *coord, int dimNode, int *elemPart)
{
    if (nbPart < 1) dc::fatal("Error: DC_partition_mesh needs at least one part.");
    std::vector<int> nodePart((size_t)(nbNodes > 0 ? nbNodes : 1), 0);
    if (nbPart > 1 && coord) {
        std::vector<int> idx((size_t)nbNodes);
        for (int i = 0; i < nbNodes; i++) idx[(size_t)i] = i;
        dc::Rcb rcb{coord, dimNode, nullptr, nodePart.data()};
        #pragma omp parallel
        #pragma omp single
        rcb.split(idx.data(), nbNodes, 0, nbPart - 1);
    } else if (nbPart > 1) {
        std::vector<int> conn0((size_t)nbElem * dimElem);
        for (int64_t k = 0; k < (int64_t)nbElem * dimElem; k++) conn0[(size_t)k] = elemToNode[k] - 1;
        std::vector<int64_t> xadj, adjncy;
        dc::build_nodal_graph(xadj, adjncy, conn0.data(), nbElem, dimElem, nbNodes);
        dc::metis_recursive(nodePart, xadj, adjncy, nbNodes, nbPart);
    }
    #pragma omp parallel for schedule(static) if (nbElem > (1 << 16))
    for (int e = 0; e < nbElem; e++) {
        int best = -1, bestCount = 0;
        for (int j = 0; j < dimElem; j++) {
            const int p = nodePart[(size_t)(elemToNode[(int64_t)e * dimElem + j] - 1)];
            int count = 0;
            for (int k = 0; k < dimElem; k++)
                if (nodePart[(size_t)(elemToNode[(int64_t)e * dimElem + k] - 1)] == p) count++;
            if (count > bestCount || (count == bestCount && p < best)) { best = p; bestCount = count; }
        }
        elemPart[e] = best;
    }
}

void DC_set_vec_size (int vecSize) { dc::options().vecSize = vecSize; }
int  DC_get_vec_size () { return dc::options().vecSize; }
void DC_set_max_elem_per_part (int maxElem) { if (maxElem > 0) dc::options().maxElemPerPart = maxElem; }
int  DC_get_max_elem_per_part () { return dc::options().maxElemPerPart; }
void DC_set_num_threads (int nbThreads)
{
    /* 0 = every processor, whatever OMP_NUM_THREADS said at start-up (torchrun exports 1) */
    dc::options().nbThreads = nbThreads > 0 ? nbThreads : omp_get_num_procs();
    omp_set_num_threads(dc::options().nbThreads);
}
tree_t *DC_get_tree () { return treeHead; }

void DC_free_tree ()
{
    dc::free_subtree(treeHead);
    treeHead = nullptr;
    delete[] elemPerm; elemPerm = nullptr;
    delete[] nodePerm; nodePerm = nullptr;
}
