/* dc_coloring.cc -- D&C + colouring variant (reference: src/coloring.cc).
 *
 * Each leaf of the tree is coloured greedily with BOUNDED colours: elements are
 * visited in ascending index, an element takes the lowest colour that no
 * already-coloured neighbour (= element of the same leaf sharing a node) holds
 * and that has fewer than vecSize members (coloring.cc:118-156).  The leaf is
 * then reordered so that complete colours (exactly vecSize members) come first,
 * in colour order, followed by the incomplete ones (permutations.cc:135-165),
 * and tree.vecOffset records the last element of the complete part
 * (coloring.cc:192-194).  The outcome does not depend on the order of the
 * node->element lists, so these are built with a counting sort instead of the
 * reference's task-parallel quicksort (tools.cc:118-148).
 */
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "dc_internal.h"

#define DC_MAX_COLOR 1024   /* reference: coloring.h:22-24 (32 blocks x 32 bits) */

void DC_create_nodeToElem (index_t &nodeToElem, int *elemToNode, int nbElem, int dimElem,
                           int nbNodes)
{
    int *idx = nodeToElem.index, *val = nodeToElem.value;
    for (int n = 0; n <= nbNodes; n++) idx[n] = 0;
    const int64_t nbConn = (int64_t)nbElem * dimElem;
    for (int64_t k = 0; k < nbConn; k++) idx[elemToNode[k]]++;          /* 1-based node -> slot n+1 */
    for (int n = 0; n < nbNodes; n++) idx[n + 1] += idx[n];
    std::vector<int> fill(idx, idx + nbNodes);
    for (int e = 0; e < nbElem; e++)
        for (int j = 0; j < dimElem; j++)
            val[fill[elemToNode[(int64_t)e * dimElem + j] - 1]++] = e;
}

void DC_create_elemToElem (list_t *elemToElem, index_t &nodeToElem, int *elemToNode,
                           int firstElem, int lastElem, int dimElem)
{
    #pragma omp parallel for schedule(dynamic, 64) if (lastElem - firstElem > 4096)
    for (int e = firstElem; e <= lastElem; e++) {
        list_t &out = elemToElem[e - firstElem];
        int count = 0;
        for (int j = 0; j < dimElem; j++) {
            int node = elemToNode[(int64_t)e * dimElem + j] - 1;
            for (int k = nodeToElem.index[node]; k < nodeToElem.index[node + 1]; k++) {
                int other = nodeToElem.value[k];
                if (other < firstElem || other > lastElem || other == e) continue;
                int rel = other - firstElem, l = 0;
                while (l < count && out.list[l] != rel) l++;
                if (l < count) continue;
                if (count >= MAX_ELEM_NEIGHBORS)
                    dc::fatal("Error: more than MAX_ELEM_NEIGHBORS neighbors for one element.");
                out.list[count++] = rel;
            }
        }
        out.size = count;
    }
}

namespace dc {

struct LeafColorer {
    int *conn;             /* 1-based connectivity in tree order */
    int *posToOrig;
    const int *n2eIndex, *n2eValue;
    int dimElem, vecSize;

    void color_leaf(tree_t *leaf)
    {
        const int first = leaf->firstElem, n = leaf->lastSep - leaf->firstElem + 1;
        std::vector<int> color((size_t)(n > 0 ? n : 1), -1), card, order((size_t)(n > 0 ? n : 1));
        std::vector<int> bannedAt;            /* bannedAt[c] == e+1  <=>  colour c is taken around e */
        int nbColors = 0;
        for (int e = 0; e < n; e++) {
            const int *row = conn + (int64_t)(first + e) * dimElem;
            for (int j = 0; j < dimElem; j++) {
                int node = row[j] - 1;
                for (int k = n2eIndex[node]; k < n2eIndex[node + 1]; k++) {
                    int rel = n2eValue[k] - first;
                    if (rel < 0 || rel >= e) continue;       /* outside the leaf, itself, or not coloured yet */
                    bannedAt[color[rel]] = e + 1;
                }
            }
            int c = 0;
            while (c < (int)card.size() && (bannedAt[c] == e + 1 || card[c] >= vecSize)) c++;
            if (c >= DC_MAX_COLOR) fatal("Error: Not enough colors.");
            if (c == (int)card.size()) { card.push_back(0); bannedAt.push_back(0); }
            color[e] = c;
            card[c]++;
            if (c + 1 > nbColors) nbColors = c + 1;
        }
        /* complete colours first, then the rest; stable inside a colour */
        std::vector<int> start((size_t)nbColors + 1, 0);
        int ptr = 0;
        for (int pass = 0; pass < 2; pass++)
            for (int c = 0; c < nbColors; c++)
                if ((card[c] == vecSize) == (pass == 0)) { start[c] = ptr; ptr += card[c]; }
        int nbInFull = 0;
        for (int c = 0; c < nbColors; c++) if (card[c] == vecSize) nbInFull += card[c];
        for (int e = 0; e < n; e++) order[e] = start[color[e]]++;
        leaf->vecOffset = (nbInFull - 1) + first;
        if (n > 0) {
            scatter_rows_inplace<int>(conn + (int64_t)first * dimElem, order.data(), n, dimElem);
            scatter_rows_inplace<int>(posToOrig + first, order.data(), n, 1);
        }
    }

    void walk(tree_t *t)
    {
        if (!t) return;
        if (t->left == nullptr && t->right == nullptr) { color_leaf(t); return; }
        walk(t->right);
        walk(t->left);
        walk(t->sep);
    }
};

/* reference: coloring.cc:251-283 */
void color_leaves(int *elemToNode, int nbElem, int dimElem, int nbNodes, int vecSize,
                  int *posToOrig)
{
    std::vector<int> index((size_t)nbNodes + 1), value((size_t)nbElem * dimElem + 1);
    index_t n2e{index.data(), value.data()};
    DC_create_nodeToElem(n2e, elemToNode, nbElem, dimElem, nbNodes);
    LeafColorer lc{elemToNode, posToOrig, index.data(), value.data(), dimElem, vecSize};
    lc.walk(treeHead);
}

}  // namespace dc
