/* dc_io.cc -- binary tree file, byte-compatible with the reference (src/IO.cc).
 *
 * Layout (IO.cc:143-158, 102-140): int nbMaxComm; int elemPerm[nbElem];
 * int nodePerm[nbNodes]; then the tree in pre-order (node, left, right, sep).
 * A node record is 8 ints (firstElem lastElem lastSep firstNode lastNode
 * firstEdge lastEdge vecOffset) + bool isSep + bool isLeaf; a leaf record adds
 * int nbOwnedNodes, int nbIntfNodes and, when non-zero, ownedNodes[] and
 * intfIndex[nbIntf+1] intfNodes[] intfDst[].  On reading, an inner node has a
 * separator child iff lastSep > lastElem (IO.cc:66-68).
 */
#include <cstdio>
#include <cstdlib>
#include <string>
#include "dc_internal.h"

namespace {

void put(FILE *f, const void *p, size_t bytes)
{
    if (bytes && fwrite(p, 1, bytes, f) != bytes) dc::fatal("Error: cannot store DC tree & permutations.");
}
void get(FILE *f, void *p, size_t bytes)
{
    if (bytes && fread(p, 1, bytes, f) != bytes) dc::fatal("Error: cannot read DC tree & permutations.");
}

void store_rec(const tree_t *t, FILE *f, int nbIntf)
{
    const int rec[8] = {t->firstElem, t->lastElem, t->lastSep, t->firstNode, t->lastNode,
                        t->firstEdge, t->lastEdge, t->vecOffset};
    const bool leaf = (t->left == nullptr && t->right == nullptr);
    put(f, rec, sizeof rec);
    put(f, &t->isSep, sizeof(bool));
    put(f, &leaf, sizeof(bool));
    if (leaf) {
        put(f, &t->nbOwnedNodes, sizeof(int));
        put(f, &t->nbIntfNodes, sizeof(int));
        if (t->nbOwnedNodes > 0) put(f, t->ownedNodes, sizeof(int) * (size_t)t->nbOwnedNodes);
        if (t->nbIntfNodes > 0) {
            put(f, t->intfIndex, sizeof(int) * (size_t)(nbIntf + 1));
            put(f, t->intfNodes, sizeof(int) * (size_t)t->nbIntfNodes);
            put(f, t->intfDst, sizeof(int) * (size_t)t->nbIntfNodes);
        }
        return;
    }
    store_rec(t->left, f, nbIntf);
    store_rec(t->right, f, nbIntf);
    if (t->sep) store_rec(t->sep, f, nbIntf);
}

tree_t *read_rec(FILE *f, int nbIntf)
{
    int rec[8];
    bool isSep, leaf;
    get(f, rec, sizeof rec);
    get(f, &isSep, sizeof(bool));
    get(f, &leaf, sizeof(bool));
    tree_t *t = dc::new_tree_node(rec[0], rec[2], rec[2] - rec[1], rec[3], rec[4], isSep, leaf);
    t->firstEdge = rec[5];
    t->lastEdge = rec[6];
    t->vecOffset = rec[7];
    if (leaf) {
        get(f, &t->nbOwnedNodes, sizeof(int));
        get(f, &t->nbIntfNodes, sizeof(int));
        if (t->nbOwnedNodes > 0) {
            t->ownedNodes = new int[t->nbOwnedNodes];
            get(f, t->ownedNodes, sizeof(int) * (size_t)t->nbOwnedNodes);
        }
        if (t->nbIntfNodes > 0) {
            t->intfIndex = new int[nbIntf + 1];
            t->intfNodes = new int[t->nbIntfNodes];
            t->intfDst = new int[t->nbIntfNodes];
            get(f, t->intfIndex, sizeof(int) * (size_t)(nbIntf + 1));
            get(f, t->intfNodes, sizeof(int) * (size_t)t->nbIntfNodes);
            get(f, t->intfDst, sizeof(int) * (size_t)t->nbIntfNodes);
        }
        return t;
    }
    t->left = read_rec(f, nbIntf);
    t->right = read_rec(f, nbIntf);
    if (t->lastSep > t->lastElem) t->sep = read_rec(f, nbIntf);
    return t;
}

}  // namespace

void DC_read_tree (std::string &treePath, int nbElem, int nbNodes, int nbIntf, int *nbMaxComm)
{
    FILE *f = fopen(treePath.c_str(), "rb");
    if (!f) dc::fatal("Error: cannot read DC tree & permutations.");
    DC_free_tree();
    elemPerm = new int[nbElem > 0 ? nbElem : 1];
    nodePerm = new int[nbNodes > 0 ? nbNodes : 1];
    int comm = 0;
    get(f, &comm, sizeof(int));
    if (nbMaxComm) *nbMaxComm = comm;
    get(f, elemPerm, sizeof(int) * (size_t)nbElem);
    get(f, nodePerm, sizeof(int) * (size_t)nbNodes);
    treeHead = read_rec(f, nbIntf);
    fclose(f);
}

void DC_store_tree (std::string &treePath, int nbElem, int nbNodes, int nbIntf, int nbMaxComm)
{
    if (!treeHead || !elemPerm || !nodePerm) dc::fatal("Error: cannot store DC tree & permutations.");
    FILE *f = fopen(treePath.c_str(), "wb");
    if (!f) dc::fatal("Error: cannot store DC tree & permutations.");
    put(f, &nbMaxComm, sizeof(int));
    put(f, elemPerm, sizeof(int) * (size_t)nbElem);
    put(f, nodePerm, sizeof(int) * (size_t)nbNodes);
    store_rec(treeHead, f, nbIntf);
    fclose(f);
    /* as in the reference (IO.cc:157) the permutations are released once stored */
    delete[] nodePerm; nodePerm = nullptr;
    delete[] elemPerm; elemPerm = nullptr;
}

void DC_read_tree (std::string &treePath, int nbElem, int nbNodes)
{
    int ignored = 0;
    DC_read_tree(treePath, nbElem, nbNodes, 0, &ignored);
}

void DC_store_tree (std::string &treePath, int nbElem, int nbNodes)
{
    DC_store_tree(treePath, nbElem, nbNodes, 0, 0);
}
