/* tests/mock/tree_split_mock.cc -- TEST INFRASTRUCTURE, not product code.
 *
 * A plain-loop stand-in for the dc_cuda_tree_* entry points (include/dc_cuda.h), written from the
 * contract in that header: classes by pivot part, stable left / right / separator order inside every
 * segment, next-level segment numbers.  tests/test_tree_device_driver.py preloads it (LD_PRELOAD) into
 * a child process so that the HOST side of DC_create_tree_device -- the breadth-first bookkeeping in
 * dc-lib_b200/host/dc_tree.cc -- can be compared with the recursive builder on a machine without a
 * GPU.  The CUDA kernels themselves are checked on the GPU (tests/test_gpu_tree_device.py).  Nothing
 * under dc-lib_b200/ links, loads or falls back to this file. */
#include <cstdint>
#include <cstring>
#include <vector>

struct dc_cuda_tree {
    int64_t nbElem; int dimElem, nbNodes;
    std::vector<int> conn, pos, seg, nodePart, segFirst, segEnd;
    std::vector<unsigned char> cls;
    int nbSeg = 0;
};

extern "C" {

const char *dc_cuda_tree_last_error(void) { return "mock"; }

int dc_cuda_tree_begin(dc_cuda_tree **out, int, const int *h_conn, int nbElem, int dimElem, const int *h_nodePart, int nbNodes)
{
    dc_cuda_tree *t = new dc_cuda_tree;
    t->nbElem = nbElem; t->dimElem = dimElem; t->nbNodes = nbNodes;
    t->conn.assign(h_conn, h_conn + (int64_t)nbElem * dimElem);
    t->nodePart.assign(h_nodePart, h_nodePart + nbNodes);
    t->pos.resize(nbElem); t->seg.assign(nbElem, 0); t->cls.assign(nbElem, 3);
    for (int i = 0; i < nbElem; i++) t->pos[i] = i;
    *out = t;
    return 0;
}

int dc_cuda_tree_classify(dc_cuda_tree *t, int nbSeg, const int *segFirst, const int *segEnd, const int *pivot, int *totals)
{
    t->nbSeg = nbSeg;
    t->segFirst.assign(segFirst, segFirst + nbSeg);
    t->segEnd.assign(segEnd, segEnd + nbSeg);
    memset(totals, 0, sizeof(int) * 3 * (size_t)nbSeg);
    for (int64_t i = 0; i < t->nbElem; i++) {
        const int s = t->seg[i];
        if (s < 0) { t->cls[i] = 3; continue; }
        if (i < segFirst[s] || i >= segEnd[s]) return -1;          /* segment numbers and ranges disagree */
        int onLeft = 0;
        for (int j = 0; j < t->dimElem; j++) onLeft += t->nodePart[t->conn[i * t->dimElem + j]] <= pivot[s];
        const int k = onLeft == t->dimElem ? 0 : (onLeft == 0 ? 1 : 2);
        t->cls[i] = (unsigned char)k;
        totals[3 * s + k]++;
    }
    return 0;
}

int dc_cuda_tree_scatter(dc_cuda_tree *t, const int *child)
{
    std::vector<int> conn(t->conn.size()), pos(t->pos.size()), seg(t->seg.size());
    std::vector<int> next((size_t)t->nbSeg * 3, 0);
    for (int s = 0; s < t->nbSeg; s++) {
        int cnt[3] = {0, 0, 0};
        for (int i = t->segFirst[s]; i < t->segEnd[s]; i++) cnt[t->cls[i]]++;
        next[3 * s] = t->segFirst[s];
        next[3 * s + 1] = next[3 * s] + cnt[0];
        next[3 * s + 2] = next[3 * s + 1] + cnt[1];
    }
    for (int64_t i = 0; i < t->nbElem; i++) {
        const int s = t->seg[i];
        int64_t d = i;
        int ns = -1;
        if (s >= 0) { d = next[3 * s + t->cls[i]]++; ns = child[3 * s + t->cls[i]]; }
        memcpy(&conn[d * t->dimElem], &t->conn[i * t->dimElem], sizeof(int) * t->dimElem);
        pos[d] = t->pos[i];
        seg[d] = ns;
    }
    t->conn.swap(conn); t->pos.swap(pos); t->seg.swap(seg);
    return 0;
}

int dc_cuda_tree_end(dc_cuda_tree *t, int *h_conn, int *h_pos, int64_t *launches)
{
    if (h_conn) memcpy(h_conn, t->conn.data(), sizeof(int) * t->conn.size());
    if (h_pos) memcpy(h_pos, t->pos.data(), sizeof(int) * t->pos.size());
    if (launches) *launches = 0;
    delete t;
    return 0;
}

}
