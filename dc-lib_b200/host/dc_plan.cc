/* dc_plan.cc -- host-side builder of the device schedule (see csrc/dc_plan.h).
 *
 * Input is what the reference's user holds after DC_create_tree /
 * DC_renumber_int_array / DC_finalize_tree: tree-ordered, renumbered connectivity,
 * the CSR pattern, and the tree (flattened in the pre-order of src/IO.cc:102-140).
 * Everything here is O(nbElem) integer work, parallel over units.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <memory>
#include <algorithm>
#include <omp.h>
#include "../csrc/dc_plan.h"

namespace {

struct Mesh {
    const int *conn;     /* 1-based */
    int nbElem, nbNodes;
    const int *row, *col;
    int colBase;
    std::vector<int64_t> n2ePtr;
    int *n2eVal = nullptr;     /* incident elements of each node, ascending */
    std::unique_ptr<int[]> n2eHold;
};

/* int -> int map with open addressing for the per-unit lookups (element -> slot, column -> edge,
 * node -> table entity): a few hundred keys, cleared per use */
struct IntMap {
    std::vector<int> key, val;
    uint32_t mask = 0;
    void reset(size_t n)
    {
        size_t cap = 16;
        while (cap < 2 * n + 2) cap <<= 1;
        if (key.size() != cap) { key.assign(cap, -1); val.resize(cap); }
        else std::fill(key.begin(), key.end(), -1);
        mask = (uint32_t)cap - 1;
    }
    static inline uint32_t hash(int k) { return (uint32_t)k * 2654435761u >> 7; }
    /* slot of k: either holds k or is free (key -1) */
    inline size_t probe(int k) const
    {
        size_t h = hash(k) & mask;
        while (key[h] != k && key[h] != -1) h = (h + 1) & mask;
        return h;
    }
    inline void put(int k, int v) { const size_t h = probe(k); key[h] = k; val[h] = v; }
    inline int get(int k) const { const size_t h = probe(k); return key[h] == k ? val[h] : -1; }
};

void build_node_to_elem(Mesh &m)
{
    const int64_t nc = (int64_t)m.nbElem * 4;
    m.n2ePtr.assign((size_t)m.nbNodes + 1, 0);
    #pragma omp parallel for schedule(static) if (nc > (1 << 18))
    for (int64_t k = 0; k < nc; k++) {
        #pragma omp atomic
        m.n2ePtr[(size_t)m.conn[k]]++;                 /* node n (1-based) -> slot n */
    }
    for (int n = 0; n < m.nbNodes; n++) m.n2ePtr[(size_t)n + 1] += m.n2ePtr[n];
    /* 4 ints per element: not zero-filled by one thread; every thread first touches one contiguous piece */
    m.n2eHold.reset(new int[(size_t)(nc > 0 ? nc : 1)]);
    m.n2eVal = m.n2eHold.get();
    #pragma omp parallel for schedule(static) if (nc > (1 << 18))
    for (int64_t k = 0; k < nc; k += 1024) m.n2eVal[k] = 0;
    std::vector<int64_t> fill(m.n2ePtr.begin(), m.n2ePtr.end() - 1);
    #pragma omp parallel for schedule(static) if (nc > (1 << 18))
    for (int e = 0; e < m.nbElem; e++)
        for (int j = 0; j < 4; j++) {
            int64_t slot;
            #pragma omp atomic capture
            slot = fill[(size_t)m.conn[(int64_t)e * 4 + j] - 1]++;
            m.n2eVal[(size_t)slot] = e;
        }
    #pragma omp parallel for schedule(dynamic, 4096) if (nc > (1 << 18))
    for (int n = 0; n < m.nbNodes; n++)
        std::sort(m.n2eVal + m.n2ePtr[n], m.n2eVal + m.n2ePtr[(size_t)n + 1]);
}

struct UnitSeed {
    int nodeFirst, nodeCount, ownFirst, ownCount;
    std::vector<int> halo;     /* sorted */
};

/* halo = incident elements of the node range that are not in [ownFirst, ownFirst+ownCount) */
void collect_halo(const Mesh &m, int nodeFirst, int nodeCount, int ownFirst, int ownCount,
                  std::vector<int> &halo)
{
    halo.clear();
    for (int64_t k = m.n2ePtr[nodeFirst]; k < m.n2ePtr[(size_t)nodeFirst + nodeCount]; k++) {
        int e = m.n2eVal[(size_t)k];
        if (e < ownFirst || e >= ownFirst + ownCount) halo.push_back(e);
    }
    std::sort(halo.begin(), halo.end());
    halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
}

struct TreeIdx {
    int firstElem, lastSep, firstNode, lastNode;
    bool isSep, isLeaf;
    int64_t left, right;
};

/* pre-order records (node, left, right, sep) -> indexed nodes; separator subtrees are skipped */
int64_t index_tree(const int *rec, int64_t nbRec, int64_t &cursor, std::vector<TreeIdx> &out, bool keep)
{
    if (cursor >= nbRec) return -1;
    const int *r = rec + 8 * cursor++;
    int64_t me = -1;
    if (keep) {
        me = (int64_t)out.size();
        out.push_back(TreeIdx{r[0], r[2], r[3], r[4], r[5] != 0, r[6] != 0, -1, -1});
    }
    if (r[6]) return me;
    const bool hasSep = r[7] != 0, descend = keep && !r[5];
    int64_t l = index_tree(rec, nbRec, cursor, out, descend);
    int64_t rr = index_tree(rec, nbRec, cursor, out, descend);
    if (hasSep) index_tree(rec, nbRec, cursor, out, false);
    if (keep) { out[(size_t)me].left = l; out[(size_t)me].right = rr; }
    return me;
}

struct Selector {
    const Mesh &m;
    int maxSlots;
    std::vector<TreeIdx> tree;
    std::vector<std::vector<UnitSeed>> seedsT;
    std::vector<std::vector<int>> bigT;
    std::vector<IntMap> seenT;
    bool evenSplit = true;
    int maxRows = 0;                 /* > 0: no unit has more rows (DC_PLAN_MAX_ROWS; experiment switch) */

    /* rows [nodeFirst, nodeFirst+nodeCount) without a usable subtree: cut greedily -- a piece takes
     * rows while the distinct elements they touch fit the slot budget */
    void split_rows(int nodeFirst, int nodeCount)
    {
        const int tid = omp_get_thread_num();
        IntMap &seen = seenT[(size_t)tid];
        int start = nodeFirst, end = nodeFirst + nodeCount;
        std::vector<int> acc;
        while (start < end) {
            int stop = start;
            acc.clear();
            seen.reset((size_t)maxSlots);
            while (stop < end) {
                const size_t before = acc.size();
                for (int64_t k = m.n2ePtr[stop]; k < m.n2ePtr[(size_t)stop + 1] && (int)acc.size() <= maxSlots; k++) {
                    const int e = m.n2eVal[(size_t)k];
                    const size_t h = seen.probe(e);
                    if (seen.key[h] == e) continue;
                    seen.key[h] = e;
                    acc.push_back(e);
                }
                if ((int)acc.size() > maxSlots) {              /* the row does not fit: the piece ends before it */
                    acc.resize(before);
                    break;
                }
                stop++;
            }
            if (stop == start) {           /* a single row exceeds the slot budget */
                bigT[tid].push_back(start);
                start++;
                continue;
            }
            std::sort(acc.begin(), acc.end());
            UnitSeed s{start, stop - start, 0, 0, {}};
            s.halo = acc;
            seedsT[tid].push_back(std::move(s));
            start = stop;
        }
    }

    /* A leaf that is too big for one unit.  The greedy cut fills the first pieces and leaves a short
     * last one; a piece of more than 32 rows needs a second diagonal task that is mostly padding
     * (24 iterations for a few lanes).  When the same number of pieces can be had with at most 32
     * rows each, cut evenly instead: every piece then has exactly one diagonal task. */
    /* k even pieces of rows [nodeFirst, +nodeCount), or nothing if one of them exceeds the budget */
    bool even_pieces(int nodeFirst, int nodeCount, int k, std::vector<UnitSeed> &pieces) const
    {
        pieces.clear();
        int start = nodeFirst;
        for (int p = 0; p < k; p++) {
            const int cnt = nodeCount / k + (p < nodeCount % k ? 1 : 0);
            UnitSeed s{start, cnt, 0, 0, {}};
            s.halo.assign(m.n2eVal + m.n2ePtr[start], m.n2eVal + m.n2ePtr[(size_t)start + cnt]);
            std::sort(s.halo.begin(), s.halo.end());
            s.halo.erase(std::unique(s.halo.begin(), s.halo.end()), s.halo.end());
            if ((int)s.halo.size() > maxSlots) { pieces.clear(); return false; }
            pieces.push_back(std::move(s));
            start += cnt;
        }
        return true;
    }

    void split_leaf(int nodeFirst, int nodeCount)
    {
        const int tid = omp_get_thread_num();
        auto even = [&](int k, std::vector<UnitSeed> &pieces) { return even_pieces(nodeFirst, nodeCount, k, pieces); };
        std::vector<UnitSeed> pieces;
        /* the common case first: both callers have found that the leaf does not fit one unit, so if two even pieces fit,
         * the greedy cut needs two pieces as well (its first piece is at least as long as the even
         * one and its second is a subset of the second even piece) -- no need to run it */
        if (evenSplit && nodeCount >= 2 && (nodeCount + 1) / 2 <= 32 && even(2, pieces)) {
            for (UnitSeed &s : pieces) seedsT[tid].push_back(std::move(s));
            return;
        }
        const size_t seeds0 = seedsT[tid].size(), big0 = bigT[tid].size();
        split_rows(nodeFirst, nodeCount);
        int k = (int)(seedsT[tid].size() - seeds0);
        if (bigT[tid].size() != big0) return;          /* a single row exceeds the budget: fallback lists */
        if (maxRows > 0 && (nodeCount + k - 1) / k > maxRows) {
            /* row cap: more, even pieces (they get smaller, so they fit sooner or later) */
            for (k = (nodeCount + maxRows - 1) / maxRows; k <= nodeCount && !even(k, pieces); k++) {}
            if (pieces.empty()) return;
        } else {
            if (!evenSplit || k < 2 || (nodeCount + k - 1) / k > 32) return;
            if (!even(k, pieces)) return;              /* an even piece does not fit: keep the greedy cut */
        }
        seedsT[tid].resize(seeds0);
        for (UnitSeed &s : pieces) seedsT[tid].push_back(std::move(s));
    }

    /* Pack consecutive non-separator leaves (their node intervals are adjacent, in ascending
     * node order) into units until the slot budget is full.  Unlike whole subtrees, whose sizes
     * jump by factors of two, this fills the budget, which lowers the halo ratio. */
    void pack_leaves(const std::vector<std::pair<int, int>> &leaves, size_t first, size_t last)
    {
        const int tid = omp_get_thread_num();
        std::vector<int> cur, leafSet, merged;
        int curFirst = -1, curCount = 0;
        auto flush = [&]() {
            if (curCount > 0) {
                UnitSeed s{curFirst, curCount, 0, 0, {}};
                s.halo.swap(cur);
                seedsT[tid].push_back(std::move(s));
            }
            cur.clear();
            curFirst = -1;
            curCount = 0;
        };
        for (size_t l = first; l < last; l++) {
            const int nf = leaves[l].first, nc = leaves[l].second;
            if (nc <= 0) continue;
            leafSet.assign(m.n2eVal + m.n2ePtr[nf], m.n2eVal + m.n2ePtr[(size_t)nf + nc]);
            std::sort(leafSet.begin(), leafSet.end());
            leafSet.erase(std::unique(leafSet.begin(), leafSet.end()), leafSet.end());
            if ((int)leafSet.size() > maxSlots) {          /* a leaf too big on its own: cut by rows */
                flush();
                split_leaf(nf, nc);
                continue;
            }
            if (maxRows > 0 && nc > maxRows) {             /* row cap (DC_PLAN_MAX_ROWS): even pieces of a leaf that fits */
                std::vector<UnitSeed> pieces;
                if (even_pieces(nf, nc, (nc + maxRows - 1) / maxRows, pieces)) {
                    flush();
                    for (UnitSeed &s : pieces) seedsT[tid].push_back(std::move(s));
                    continue;
                }
            }
            if (curCount > 0 && curFirst + curCount == nf && (maxRows <= 0 || curCount + nc <= maxRows)) {
                merged.resize(cur.size() + leafSet.size());
                merged.resize((size_t)(std::set_union(cur.begin(), cur.end(), leafSet.begin(), leafSet.end(), merged.begin()) -
                                       merged.begin()));
                if ((int)merged.size() <= maxSlots) {
                    cur.swap(merged);
                    curCount += nc;
                    continue;
                }
            }
            flush();
            cur.swap(leafSet);
            curFirst = nf;
            curCount = nc;
        }
        flush();
    }

    /* Units of a fixed number of consecutive rows (cut short where the slot budget would be
     * exceeded).  With 32 rows a unit has exactly one full diagonal task and about seven
     * upper-edge tasks: one task per warp of a 256-thread CTA. */
    void window_rows(int first, int last, int width)
    {
        const int tid = omp_get_thread_num();
        std::vector<int> acc, trial;
        int start = first;
        while (start < last) {
            int stop = start;
            acc.clear();
            while (stop < last && stop - start < width) {
                trial.assign(acc.begin(), acc.end());
                for (int64_t k = m.n2ePtr[stop]; k < m.n2ePtr[(size_t)stop + 1]; k++)
                    trial.push_back(m.n2eVal[(size_t)k]);
                std::sort(trial.begin(), trial.end());
                trial.erase(std::unique(trial.begin(), trial.end()), trial.end());
                if ((int)trial.size() > maxSlots) break;
                acc.swap(trial);
                stop++;
            }
            if (stop == start) {
                bigT[tid].push_back(start);
                start++;
                continue;
            }
            UnitSeed s{start, stop - start, 0, 0, {}};
            s.halo = acc;
            seedsT[tid].push_back(std::move(s));
            start = stop;
        }
    }

    /* take the subtree as one unit if it fits the slot budget, else descend */
    void select(int64_t idx)
    {
        if (idx < 0) return;
        const TreeIdx t = tree[(size_t)idx];
        if (t.isSep) return;               /* separator subtrees own no rows: they appear as halo */
        const int own = t.lastSep - t.firstElem + 1, nodes = t.lastNode - t.firstNode + 1;
        if (own <= maxSlots && nodes > 0) {
            std::vector<int> halo;
            collect_halo(m, t.firstNode, nodes, t.firstElem, own, halo);
            if (own + (int)halo.size() <= maxSlots) {
                UnitSeed s{t.firstNode, nodes, t.firstElem, own, {}};
                s.halo.swap(halo);
                seedsT[omp_get_thread_num()].push_back(std::move(s));
                return;
            }
        }
        if (t.isLeaf) {
            if (nodes > 0) split_leaf(t.firstNode, nodes);
            return;
        }
        #pragma omp task default(shared) if (own > 20000)
        select(t.left);
        #pragma omp task default(shared) if (own > 20000)
        select(t.right);
        #pragma omp taskwait
    }
};

struct UnitOut {
    std::vector<dc_task_t> tasks;
    std::vector<uint16_t> items;
    std::vector<uint16_t> slotNodes;   /* 4 per slot */
    std::vector<int32_t> haloNodes;
    std::vector<int32_t> tgt;      /* symmetric schedule: (edge, transposed edge) per upper edge */
    std::vector<int32_t> slotElems; /* element of every slot */
    int64_t useful = 0;
};

/* Bank-aware slot numbering.  In the accumulation phase the 16 lanes of a half-warp read, in one
 * instruction, one component of the node gradients (slot, node a) their items name; the record
 * stride is 13 doubles and a gradient has 3, so the 8-byte bank is 3*(a - slot) + k (mod 16):
 * two different gradients collide iff (a - slot) mod 16 is equal.  The slot numbers inside a
 * unit are free, so they are chosen here: slots are taken in order of decreasing use and put in
 * the residue class (slot mod 16) that adds the fewest collisions to the half-warp accesses
 * seen so far (one greedy pass; class sizes are fixed by the slot count, so no slot is wasted).
 * Measured on the plan (tools/smem_model.py): 2.5 -> 1.65 wavefronts per access. */
struct BankOpt {
    /* which record layout the slot classes are chosen for: 0 = 13-word records of four 3-word node
     * gradients (bank key (a - slot) mod 16, two reads per off-diagonal contribution: elasticity and
     * the default); 1 = 11-word records of the 10 entries of a symmetric 4x4 element matrix, one read
     * of entry q(a,b) per contribution (P1 Laplacian): bank = 11*slot + q = q - 5*slot (mod 16) */
    int model = 0;
    int nOff = 4;                   /* offsets an entity can be read at (4 nodes / 10 entries) */
    std::vector<int> incPtr, incGroup, order, fill;
    std::vector<uint8_t> incNode, cnt, gmax;

    std::vector<int> stamp;          /* last group that touched a (slot, node) address */
    std::vector<int> tmpG, tmpX;     /* incidences in visiting order */
    std::vector<uint8_t> tmpO;

    /* every half-warp access of the accumulation phase: f(group, slot, node) once per distinct
     * address (equal addresses are a broadcast, not a conflict) */
    template <class F>
    int visit(const UnitOut &out, int nSlots, F &&f)
    {
        static const unsigned char entry[16] = {0, 1, 2, 3, 1, 4, 5, 6, 2, 5, 7, 8, 3, 6, 8, 9};
        stamp.assign((size_t)nSlots * 16, -1);
        int g = 0;
        for (const dc_task_t &t : out.tasks) {
            const uint16_t *it = out.items.data() + t.itemRel;
            const int sides = (model == 1 || t.kind == DC_TASK_OWNDIAG) ? 1 : 2;
            for (int half = 0; half < 2; half++)
                for (int side = 0; side < sides; side++)
                    for (int r = 0; r < (int)t.iters; r++, g++) {
                        const uint16_t *row = it + r * 32 + half * 16;
                        for (int l = 0; l < 16; l++) {
                            const uint32_t item = row[l];
                            const int slot = (int)(item >> 4);
                            if (slot >= nSlots) continue;
                            const int node = model == 1 ? (int)entry[item & 15u]
                                                        : side ? (int)(item & 3u) : (int)((item >> 2) & 3u);
                            int &st = stamp[(size_t)slot * 16 + node];
                            if (st == g) continue;
                            st = g;
                            f(g, slot, node);
                        }
                    }
        }
        return g;
    }

    /* entities with incidences (group, offset) in incPtr/incGroup/incNode; an entity in class c
     * occupies key (offset - c) mod 16 of each of its groups.  fixedCls[x] >= 0 pins an entity;
     * the others are placed in order of decreasing use into the class (with room left, cap[c])
     * that adds the fewest collisions.  Result in cls. */
    void greedy(int nEnt, int nbGroups, const int *fixedCls, const int *capIn, std::vector<int> &cls)
    {
        cls.assign((size_t)nEnt, -1);
        cnt.assign((size_t)nbGroups * 16, 0);
        gmax.assign((size_t)nbGroups, 0);
        int used[16], cap[16];
        for (int c = 0; c < 16; c++) { used[c] = 0; cap[c] = capIn[c]; }
        auto place = [&](int x, int c) {
            cls[(size_t)x] = c;
            for (int q = incPtr[(size_t)x]; q < incPtr[(size_t)x + 1]; q++) {
                const int g = incGroup[(size_t)q], k = ((int)incNode[(size_t)q] - c) & 15;
                const uint8_t now = ++cnt[(size_t)g * 16 + k];
                if (now > gmax[(size_t)g]) gmax[(size_t)g] = now;
            }
        };
        order.clear();
        for (int x = 0; x < nEnt; x++) {
            if (fixedCls && fixedCls[x] >= 0) place(x, fixedCls[x]);
            else order.push_back(x);
        }
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
            return incPtr[(size_t)x + 1] - incPtr[(size_t)x] > incPtr[(size_t)y + 1] - incPtr[(size_t)y];
        });
        for (int x : order) {
            /* cost of every class at once: per offset, the sum over the entity's groups of
             * (entries already on that key) + 64 if one more raises the group's maximum */
            uint16_t w[16][16];
            for (int o = 0; o < nOff; o++)
                for (int k = 0; k < 16; k++) w[o][k] = 0;
            for (int q = incPtr[(size_t)x]; q < incPtr[(size_t)x + 1]; q++) {
                const uint8_t *row = cnt.data() + (size_t)incGroup[(size_t)q] * 16;
                const uint8_t mx = gmax[(size_t)incGroup[(size_t)q]];
                uint16_t *dst = w[incNode[(size_t)q] & 15];
                #pragma omp simd
                for (int k = 0; k < 16; k++) dst[k] = (uint16_t)(dst[k] + row[k] + (row[k] >= mx ? 64 : 0));
            }
            int bestC = -1, bestV = 0;
            for (int c = 0; c < 16; c++) {
                if (used[c] >= cap[c]) continue;
                int v = 0;
                for (int o = 0; o < nOff; o++) v += w[o][(o - c) & 15];
                if (bestC < 0 || v < bestV) { bestC = c; bestV = v; }
            }
            used[bestC]++;
            place(x, bestC);
        }
    }

    /* incidence lists from a visitor that calls f(group, entity, offset); returns the group count */
    template <class V>
    int collect(int nEnt, V &&visitAll)
    {
        tmpG.clear(); tmpX.clear(); tmpO.clear();
        incPtr.assign((size_t)nEnt + 1, 0);
        const int nbGroups = visitAll([&](int g, int x, int off) {
            tmpG.push_back(g); tmpX.push_back(x); tmpO.push_back((uint8_t)off);
            incPtr[(size_t)x + 1]++;
        });
        for (int x = 0; x < nEnt; x++) incPtr[(size_t)x + 1] += incPtr[(size_t)x];
        incGroup.resize(tmpG.size());
        incNode.resize(tmpG.size());
        fill.assign(incPtr.begin(), incPtr.end() - 1);
        for (size_t i = 0; i < tmpG.size(); i++) {
            const int q = fill[(size_t)tmpX[i]]++;
            incGroup[(size_t)q] = tmpG[i];
            incNode[(size_t)q] = tmpO[i];
        }
        return nbGroups;
    }

    void run(const UnitOut &out, int nSlots, std::vector<int> &newSlot)
    {
        newSlot.resize((size_t)nSlots);
        for (int s = 0; s < nSlots; s++) newSlot[(size_t)s] = s;
        if (nSlots <= 16) return;
        const int nbGroups = collect(nSlots, [&](auto &&f) { return visit(out, nSlots, f); });
        /* the greedy works on keys (offset - class) mod 16; model 1 reads bank q - 5*slot, so its
         * class is 5*(slot mod 16) and the slot residue of a class is 13*class (5*13 = 1 mod 16) */
        const int mul = model == 1 ? 13 : 1;
        int cap[16], used[16];
        for (int c = 0; c < 16; c++) { used[c] = 0; cap[c] = (nSlots - ((mul * c) & 15) + 15) / 16; }
        greedy(nSlots, nbGroups, nullptr, cap, newSlot);
        /* inside a class, keep the incoming (element) order: the 16 consecutive slots that one
         * half-warp turns into records then hold elements close to each other, which share nodes
         * (their coordinate reads are broadcasts) */
        for (int s = 0; s < nSlots; s++) {
            const int c = (mul * newSlot[(size_t)s]) & 15;
            newSlot[(size_t)s] = c + 16 * used[c]++;
        }
    }

    /* The same for the record build: the 16 threads of a half-warp read, per instruction, one
     * coordinate of node j of 16 consecutive slots; the bank is 3*tableIndex + q (mod 16), so two
     * different nodes collide iff their table indices are equal mod 16.  Own nodes sit where the
     * bulk copy of the coordinate block puts them; the halo nodes' table positions are free.
     * nodeOf[slot*4+j] = entity (0..nEnt-1), fixedCls as in greedy(); positions of the free
     * entities (table index mod 16 = (16 - class) mod 16) are handed out in haloPos. */
    void run_nodes(const std::vector<int> &nodeOf, int nSlots, int nEnt, const std::vector<int> &fixedCls,
                   int ownSpan, int nHalo, std::vector<int> &cls)
    {
        const int nbGroups = collect(nEnt, [&](auto &&f) {
            int g = 0;
            for (int s0 = 0; s0 < nSlots; s0 += 16)
                for (int j = 0; j < 4; j++, g++) {
                    int seen[16], nseen = 0;
                    for (int l = 0; l < 16 && s0 + l < nSlots; l++) {
                        const int x = nodeOf[(size_t)(s0 + l) * 4 + j];
                        bool dup = false;
                        for (int q = 0; q < nseen; q++) dup |= seen[q] == x;
                        if (dup) continue;
                        seen[nseen++] = x;
                        f(g, x, 0);
                    }
                }
            return g;
        });
        int cap[16];
        for (int c = 0; c < 16; c++) cap[c] = 0;
        for (int p = 0; p < nHalo; p++) cap[(16 - ((ownSpan + p) & 15)) & 15]++;
        const int keep = nOff;
        nOff = 1;                                   /* every entity is read at offset 0 */
        greedy(nEnt, nbGroups, fixedCls.data(), cap, cls);
        nOff = keep;
    }
};

/* per-thread work arrays of build_unit */
struct Scratch {
    IntMap elemSlot, colEdge, nodeEnt;
    std::vector<int> pairK;            /* contributions in visiting order: edge (relative) ... */
    std::vector<uint16_t> pairItem;    /* ... and item */
    std::vector<int> edgePtr;          /* items of edge k: edgeItems[edgePtr[k] .. edgePtr[k+1]) */
    std::vector<uint16_t> edgeItems;
    std::vector<int> fill;
    std::vector<char> isDiag;
    std::vector<int> outside, nodeOf, fixedCls, cls, tableIdx, newSlot, regroup;
    std::vector<int32_t> movedElems;
    std::vector<uint8_t> residue;
    double phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};   /* seconds per build_unit phase (DC_PLAN_VERBOSE) */
    inline int count(int k) const { return edgePtr[(size_t)k + 1] - edgePtr[(size_t)k]; }
    inline const uint16_t *items(int k) const { return edgeItems.data() + edgePtr[(size_t)k]; }
};

/* edge index of (rowNode, colNode) */
inline int find_edge(const Mesh &m, int rowNode, int colNode)
{
    const int target = colNode + m.colBase;
    for (int k = m.row[rowNode]; k < m.row[rowNode + 1]; k++)
        if (m.col[k] == target) return k;
    return -1;
}

inline int slot_of(const UnitSeed &s, int e)
{
    if (e >= s.ownFirst && e < s.ownFirst + s.ownCount) return e - s.ownFirst;
    return s.ownCount + (int)(std::lower_bound(s.halo.begin(), s.halo.end(), e) - s.halo.begin());
}

/* items of every edge of the unit, grouped per edge in ascending element order */
void build_tasks(const Mesh &m, const UnitSeed &s, int32_t *diagRel, UnitOut &out,
                 Scratch &w, int &error, bool symmetric)
{
    const int64_t edgeFirst = m.row[s.nodeFirst];
    const int edgeCount = m.row[s.nodeFirst + s.nodeCount] - (int)edgeFirst;
    w.isDiag.assign((size_t)edgeCount, 0);
    w.edgePtr.assign((size_t)edgeCount + 1, 0);
    w.pairK.clear();
    w.pairItem.clear();
    /* padding of the item rows: slot nSlots is the kernel's all-zero record */
    const uint16_t padItem = (uint16_t)((s.ownCount + (int)s.halo.size()) << 4);
    w.elemSlot.reset(s.halo.size());
    for (size_t h = 0; h < s.halo.size(); h++) w.elemSlot.put(s.halo[h], s.ownCount + (int)h);

    /* a row is visited once, its incident elements in ascending order: the contributions of an
     * edge come out in ascending element order */
    for (int n = s.nodeFirst; n < s.nodeFirst + s.nodeCount; n++) {
        const int r0 = m.row[n], r1 = m.row[n + 1];
        diagRel[n] = -1;
        w.colEdge.reset((size_t)(r1 - r0));
        for (int k = r1 - 1; k >= r0; k--) {               /* the first of equal columns wins, as a forward search would */
            const int c = m.col[k] - m.colBase;
            w.colEdge.put(c, (int)(k - edgeFirst));
            if (c == n) w.isDiag[(size_t)(k - edgeFirst)] = 1;
        }
        for (int k = r0; k < r1; k++)
            if (m.col[k] - m.colBase == n) diagRel[n] = (int)(k - edgeFirst);
        for (int64_t q = m.n2ePtr[n]; q < m.n2ePtr[(size_t)n + 1]; q++) {
            const int e = m.n2eVal[(size_t)q];
            const int *c = m.conn + (int64_t)e * 4;
            int a = 0;
            while (a < 4 && c[a] - 1 != n) a++;
            const int slot = (e >= s.ownFirst && e < s.ownFirst + s.ownCount) ? e - s.ownFirst : w.elemSlot.get(e);
            for (int b = 0; b < 4; b++) {
                if (symmetric && c[b] - 1 < n) continue;  /* lower edges are written as transposes */
                const int rel = w.colEdge.get(c[b] - 1);
                if (rel < 0) { error = 1; continue; }     /* pattern lacks an element edge */
                w.pairK.push_back(rel);
                w.pairItem.push_back((uint16_t)((slot << 4) | (a << 2) | b));
                w.edgePtr[(size_t)rel + 1]++;
            }
        }
    }
    for (int k = 0; k < edgeCount; k++) w.edgePtr[(size_t)k + 1] += w.edgePtr[(size_t)k];
    w.edgeItems.resize(w.pairK.size());
    w.fill.assign(w.edgePtr.begin(), w.edgePtr.end() - 1);
    for (size_t i = 0; i < w.pairK.size(); i++) w.edgeItems[(size_t)w.fill[(size_t)w.pairK[i]]++] = w.pairItem[i];
    const std::vector<char> &isDiag = w.isDiag;
    if (symmetric) {
        /* owned edges = diagonal + upper edges (col > row) of the unit's rows.  Each becomes one
         * lane; an upper-edge lane also owns the transposed edge (col,row), wherever that row
         * lives.  Lanes are grouped 32 per task after a stable sort by contribution count
         * (descending), so that the lanes of a task run the same number of iterations --
         * diagonals (one contribution per incident element) end up in tasks of their own. */
        struct Owned { int rel; int k; int kt; int count; };
        std::vector<Owned> owned;
        for (int n = s.nodeFirst; n < s.nodeFirst + s.nodeCount; n++)
            for (int k = m.row[n]; k < m.row[n + 1]; k++) {
                const int c = m.col[k] - m.colBase;
                if (c < n) continue;
                int kt = -1;
                if (c > n) {
                    kt = find_edge(m, c, n);
                    if (kt < 0) { error = 1; continue; }      /* pattern is not structurally symmetric */
                }
                const int rel = (int)(k - edgeFirst);
                owned.push_back(Owned{rel, k, kt, w.count(rel)});
            }
        /* diagonals first (their own task kind: only one gradient per contribution, symmetric
         * block), then upper edges; inside each class by descending contribution count */
        std::stable_sort(owned.begin(), owned.end(), [](const Owned &x, const Owned &y) {
            const bool dx = x.kt < 0, dy = y.kt < 0;
            if (dx != dy) return dx;
            return x.count > y.count;
        });
        for (const Owned &o : owned) {
            out.tgt.push_back(o.k);
            out.tgt.push_back(o.kt);
        }
        size_t nbDiag = 0;
        while (nbDiag < owned.size() && owned[nbDiag].kt < 0) nbDiag++;
        for (size_t base = 0; base < owned.size(); base += 32) {
            const bool diagTask = base < nbDiag;
            const size_t classEnd = diagTask ? nbDiag : owned.size();
            dc_task_t t;
            t.kind = diagTask ? DC_TASK_OWNDIAG : DC_TASK_UPPER;
            t.first = (uint32_t)base;
            t.mask = 0;
            t.itemRel = (uint32_t)out.items.size();
            size_t iters = 0;
            if (diagTask && base + 32 > nbDiag) {
                /* last diagonal task: stop at the class boundary, the next task starts there */
            }
            for (size_t l = 0; l < 32 && base + l < classEnd; l++) {
                t.mask |= 1u << l;
                iters = std::max(iters, (size_t)owned[base + l].count);
            }
            t.iters = (uint16_t)iters;
            out.items.resize(out.items.size() + iters * 32, padItem);
            uint16_t *dst = out.items.data() + t.itemRel;
            for (size_t l = 0; l < 32 && base + l < classEnd; l++) {
                const int rel = owned[base + l].rel, nv = w.count(rel);
                const uint16_t *v = w.items(rel);
                for (int i = 0; i < nv; i++) dst[(size_t)i * 32 + l] = v[i];
                out.useful += nv;
            }
            out.tasks.push_back(t);
            if (diagTask && base + 32 > nbDiag) base = nbDiag - 32;      /* next task begins at the boundary */
        }
        return;
    }
    /* edge tasks: 32 consecutive edges, item rows [iter][lane]; diagonal lanes are
     * left to the diagonal tasks */
    for (int base = 0; base < edgeCount; base += 32) {
        dc_task_t t;
        t.kind = DC_TASK_EDGES;
        t.first = (uint32_t)base;
        t.mask = 0;
        t.itemRel = (uint32_t)out.items.size();
        size_t iters = 0;
        for (int l = 0; l < 32 && base + l < edgeCount; l++)
            if (isDiag[(size_t)(base + l)]) t.mask |= 1u << l;
        iters = 0;
        for (int l = 0; l < 32 && base + l < edgeCount; l++)
            if (!((t.mask >> l) & 1u)) iters = std::max(iters, (size_t)w.count(base + l));
        t.iters = (uint16_t)iters;
        out.items.resize(out.items.size() + iters * 32, padItem);
        uint16_t *dst = out.items.data() + t.itemRel;
        for (int l = 0; l < 32 && base + l < edgeCount; l++) {
            if ((t.mask >> l) & 1u) continue;
            const int nv = w.count(base + l);
            const uint16_t *v = w.items(base + l);
            for (int i = 0; i < nv; i++) dst[(size_t)i * 32 + l] = v[i];
            out.useful += nv;
        }
        out.tasks.push_back(t);
    }
    std::vector<dc_task_t> diagTasks;
    /* diagonal tasks: 32 consecutive nodes */
    for (int base = 0; base < s.nodeCount; base += 32) {
        dc_task_t t;
        t.kind = DC_TASK_DIAG;
        t.first = (uint32_t)base;
        t.mask = 0;
        t.itemRel = (uint32_t)out.items.size();
        size_t iters = 0;
        for (int l = 0; l < 32 && base + l < s.nodeCount; l++) {
            int d = diagRel[s.nodeFirst + base + l];
            if (d < 0) continue;
            t.mask |= 1u << l;
            iters = std::max(iters, (size_t)w.count(d));
        }
        t.iters = (uint16_t)iters;
        out.items.resize(out.items.size() + iters * 32, padItem);
        uint16_t *dst = out.items.data() + t.itemRel;
        for (int l = 0; l < 32 && base + l < s.nodeCount; l++) {
            if (!((t.mask >> l) & 1u)) continue;
            const int d = diagRel[s.nodeFirst + base + l], nv = w.count(d);
            const uint16_t *v = w.items(d);
            for (int i = 0; i < nv; i++) dst[(size_t)i * 32 + l] = v[i];
            out.useful += nv;
        }
        if (t.mask) diagTasks.push_back(t); else out.items.resize(t.itemRel);
    }
    /* longest tasks first: the kernel hands tasks to warps dynamically */
    out.tasks.insert(out.tasks.begin(), diagTasks.begin(), diagTasks.end());
}


/* slots renumbered for the banks, then the node table: own interval (from the even node at or
 * below nodeFirst, padded to an even count so that the block of coordinates is a 16-byte granule
 * copy), then the halo nodes; slotNodes in slot order */
void build_unit(const Mesh &m, const UnitSeed &s, int32_t *diagRel, UnitOut &out,
                Scratch &w, int &error, bool symmetric, BankOpt &opt, bool bankAware, int regroupPasses)
{
    double tp = omp_get_wtime();
    auto lap = [&](int phase) { const double t = omp_get_wtime(); w.phase[phase] += t - tp; tp = t; };
    build_tasks(m, s, diagRel, out, w, error, symmetric);
    if (error) return;
    lap(0);
    const int nSlots = s.ownCount + (int)s.halo.size();
    std::vector<int> &newSlot = w.newSlot;
    if (bankAware) opt.run(out, nSlots, newSlot);
    else { newSlot.resize((size_t)nSlots); for (int q = 0; q < nSlots; q++) newSlot[(size_t)q] = q; }
    lap(1);
    out.slotElems.assign((size_t)nSlots, -1);
    for (int sl = 0; sl < nSlots; sl++)
        out.slotElems[(size_t)newSlot[(size_t)sl]] =
            sl < s.ownCount ? s.ownFirst + sl : s.halo[(size_t)(sl - s.ownCount)];
    const int alignedFirst = s.nodeFirst & ~1;
    const int ownSpan = ((s.nodeFirst + s.nodeCount + 1) & ~1) - alignedFirst;
    /* distinct nodes outside the own interval, ascending */
    std::vector<int> &outside = w.outside;
    outside.clear();
    w.nodeEnt.reset((size_t)nSlots * 2 + 16);
    for (int sl = 0; sl < nSlots; sl++) {
        const int e = out.slotElems[(size_t)sl];
        for (int j = 0; j < 4; j++) {
            const int nd = m.conn[(int64_t)e * 4 + j] - 1;
            if (nd >= s.nodeFirst && nd < s.nodeFirst + s.nodeCount) continue;
            const size_t h = w.nodeEnt.probe(nd);
            if (w.nodeEnt.key[h] == nd) continue;
            w.nodeEnt.key[h] = nd;
            outside.push_back(nd);
        }
    }
    std::sort(outside.begin(), outside.end());
    const int nHalo = (int)outside.size();
    if (ownSpan + nHalo > 65535) { error = 1; return; }
    for (int x = 0; x < nHalo; x++) w.nodeEnt.val[w.nodeEnt.probe(outside[(size_t)x])] = ownSpan + x;
    /* entity = own node (table index fixed) or halo node (position chosen for the banks) */
    std::vector<int> &nodeOf = w.nodeOf, &fixedCls = w.fixedCls, &cls = w.cls, &tableIdx = w.tableIdx;
    fixedCls.assign((size_t)(ownSpan + nHalo), -1);
    for (int x = 0; x < ownSpan; x++) fixedCls[(size_t)x] = (16 - (x & 15)) & 15;
    /* entities of the four nodes of every slot, in the current slot order */
    auto entities = [&]() {
        nodeOf.resize((size_t)nSlots * 4);
        for (int sl = 0; sl < nSlots; sl++) {
            const int e = out.slotElems[(size_t)sl];
            for (int j = 0; j < 4; j++) {
                const int nd = m.conn[(int64_t)e * 4 + j] - 1;
                if (nd >= s.nodeFirst && nd < s.nodeFirst + s.nodeCount) nodeOf[(size_t)sl * 4 + j] = nd - alignedFirst;
                else nodeOf[(size_t)sl * 4 + j] = w.nodeEnt.get(nd);
            }
        }
    };
    /* table positions of the halo nodes for the current slot order */
    auto place_halo = [&](bool optimise) {
        tableIdx.resize((size_t)(ownSpan + nHalo));
        for (int x = 0; x < ownSpan + nHalo; x++) tableIdx[(size_t)x] = x;
        if (optimise && bankAware && nHalo > 16) {
            opt.run_nodes(nodeOf, nSlots, ownSpan + nHalo, fixedCls, ownSpan, nHalo, cls);
            /* positions ownSpan + p with (ownSpan + p) mod 16 == (16 - class) mod 16, in node order */
            int nextPos[16];
            for (int r = 0; r < 16; r++) nextPos[r] = ((r - ownSpan) % 16 + 16) % 16;      /* first p with that residue */
            for (int x = ownSpan; x < ownSpan + nHalo; x++) {
                const int r = (16 - cls[(size_t)x]) & 15;
                tableIdx[(size_t)x] = ownSpan + nextPos[r];
                nextPos[r] += 16;
            }
        }
    };
    const bool regroup = bankAware && nHalo > 16 && nSlots > 32 && regroupPasses > 0;
    entities();
    lap(2);
    place_halo(!regroup);          /* the re-deal starts from the natural halo positions; they are optimised after it */
    if (regroup)
        for (int pass = 0; pass < regroupPasses; pass++) {
            /* The record reads fix a slot's CLASS (slot mod 16), not its rank inside the class, and
             * the record build reads the coordinates of 16 consecutive slots -- one of every class
             * -- per instruction.  Re-deal the ranks: group after group, every class contributes,
             * out of its next few slots in element order (neighbouring elements share nodes, which
             * are broadcasts), the one whose four nodes collide least with the table positions the
             * group already reads; then the halo nodes are placed again for the new groups. */
            std::vector<int> &perm = w.regroup;           /* current slot -> new slot */
            perm.assign((size_t)nSlots, -1);
            int head[16];
            for (int c = 0; c < 16; c++) head[c] = c;      /* first slot of the class not yet dealt */
            const int window = 8;
            std::vector<uint8_t> &res = w.residue;        /* bank residue of the table position of every (slot, j) */
            res.resize((size_t)nSlots * 4);
            for (size_t q = 0; q < (size_t)nSlots * 4; q++) res[q] = (uint8_t)(tableIdx[(size_t)nodeOf[q]] & 15);
            for (int base = 0; base < nSlots; base += 16) {
                /* distinct table entities the group reads per node position j and bank residue */
                int occ[4][16][16], nocc[4][16];
                for (int j = 0; j < 4; j++)
                    for (int r = 0; r < 16; r++) nocc[j][r] = 0;
                for (int c = 0; c < 16 && base + c < nSlots; c++) {
                    while (head[c] < nSlots && perm[(size_t)head[c]] >= 0) head[c] += 16;
                    int best = -1, bestCost = 1 << 30, seen = 0;
                    for (int cand = head[c]; cand < nSlots && seen < window; cand += 16) {
                        if (perm[(size_t)cand] >= 0) continue;
                        seen++;
                        int cost = 0;
                        for (int j = 0; j < 4; j++) {
                            const int r = res[(size_t)cand * 4 + j], have = nocc[j][r];
                            if (!have) continue;
                            const int ent = nodeOf[(size_t)cand * 4 + j];
                            bool present = false;
                            for (int q = 0; q < have; q++) present |= occ[j][r][q] == ent;
                            if (!present) cost += have;             /* one more address on a bank that has `have` already */
                        }
                        if (cost < bestCost) { bestCost = cost; best = cand; }
                        if (cost == 0) break;
                    }
                    perm[(size_t)best] = base + c;
                    for (int j = 0; j < 4; j++) {
                        const int ent = nodeOf[(size_t)best * 4 + j], r = res[(size_t)best * 4 + j];
                        bool present = false;
                        for (int q = 0; q < nocc[j][r]; q++) present |= occ[j][r][q] == ent;
                        if (!present) occ[j][r][nocc[j][r]++] = ent;
                    }
                }
            }
            std::vector<int32_t> &moved = w.movedElems;
            moved.assign((size_t)nSlots, -1);
            for (int sl = 0; sl < nSlots; sl++) moved[(size_t)perm[(size_t)sl]] = out.slotElems[(size_t)sl];
            out.slotElems.swap(moved);
            for (int sl = 0; sl < nSlots; sl++) newSlot[(size_t)sl] = perm[(size_t)newSlot[(size_t)sl]];
            lap(3);
            entities();
            place_halo(true);
        }
    lap(4);
    for (uint16_t &it : out.items) {
        const int slot = it >> 4;
        if (slot < nSlots) it = (uint16_t)((newSlot[(size_t)slot] << 4) | (it & 15));
    }
    out.haloNodes.assign((size_t)nHalo, 0);
    for (int x = ownSpan; x < ownSpan + nHalo; x++) out.haloNodes[(size_t)(tableIdx[(size_t)x] - ownSpan)] = outside[(size_t)(x - ownSpan)];
    out.slotNodes.resize((size_t)nSlots * 4);
    for (size_t q = 0; q < (size_t)nSlots * 4; q++) out.slotNodes[q] = (uint16_t)tableIdx[(size_t)nodeOf[q]];
    while (out.slotNodes.size() % 8) out.slotNodes.push_back(0);    /* 16-byte granules */
    while (out.slotElems.size() % 2) out.slotElems.push_back(-1);
    while (out.haloNodes.size() % 4) out.haloNodes.push_back(0);
    lap(5);
}

int g_bankModel = 0;

template <typename T>
T *dup(const std::vector<T> &v)
{
    T *p = (T *)malloc(sizeof(T) * (v.size() ? v.size() : 1));
    if (v.size()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

}  // namespace

extern "C" int dc_contrib_build(dc_contrib_list_t *out, const int *conn, int nbElem, int nbNodes,
                                const int *row, const int *col, int colBase, const int *rows,
                                int64_t nbRows, int symmetric)
{
    /* list items are elem << 4 | a << 2 | b in 32 bits: the element must fit 28 bits */
    if (nbElem >= (1 << 28)) return -5;
    Mesh m{conn, nbElem, nbNodes, row, col, colBase, {}, nullptr, {}};
    build_node_to_elem(m);
    std::vector<int64_t> edge, ptr, edgeT;
    std::vector<uint32_t> item;
    ptr.push_back(0);
    int error = 0;
    std::vector<std::vector<uint32_t>> perEdge;
    for (int64_t r = 0; r < (rows ? nbRows : (int64_t)nbNodes); r++) {
        const int n = rows ? rows[r] : (int)r;
        const int r0 = row[n], r1 = row[n + 1];
        perEdge.assign((size_t)(r1 - r0), {});
        for (int64_t q = m.n2ePtr[n]; q < m.n2ePtr[(size_t)n + 1]; q++) {
            const int e = m.n2eVal[(size_t)q];
            const int *c = conn + (int64_t)e * 4;
            int a = 0;
            while (a < 4 && c[a] - 1 != n) a++;
            for (int b = 0; b < 4; b++) {
                if (symmetric && c[b] - 1 < n) continue;
                int k = r0;
                while (k < r1 && col[k] != c[b] - 1 + colBase) k++;
                if (k == r1) { error = 1; continue; }
                perEdge[(size_t)(k - r0)].push_back(((uint32_t)e << 4) | (uint32_t)(a << 2) | (uint32_t)b);
            }
        }
        for (int k = r0; k < r1; k++) {
            const int c = col[k] - colBase;
            if (symmetric && c < n) continue;            /* written as a transpose by the owner of c */
            edge.push_back(k);
            if (symmetric) {
                int64_t kt = -1;
                if (c > n) {
                    kt = find_edge(m, c, n);
                    if (kt < 0) error = 1;
                }
                edgeT.push_back(kt);
            }
            item.insert(item.end(), perEdge[(size_t)(k - r0)].begin(), perEdge[(size_t)(k - r0)].end());
            ptr.push_back((int64_t)item.size());
        }
    }
    out->nbEdges = (int64_t)edge.size();
    out->nbItems = (int64_t)item.size();
    out->edge = dup(edge);
    out->ptr = dup(ptr);
    out->item = dup(item);
    out->edgeT = symmetric ? dup(edgeT) : nullptr;
    return error ? -2 : 0;
}

extern "C" int dc_plan_build(dc_plan_t *plan, const int *conn, int nbElem, int nbNodes,
                             const int *row, const int *col, int colBase, const int *treeRec,
                             int64_t nbTreeRec, int maxSlots, int nbThreads, int symmetric)
{
    memset(plan, 0, sizeof *plan);
    const bool verbose = getenv("DC_PLAN_VERBOSE") != nullptr;
    double t0 = omp_get_wtime();
    auto lap = [&](const char *what) {
        if (verbose) { double t1 = omp_get_wtime(); fprintf(stderr, "[dc_plan] %-12s %.2fs\n", what, t1 - t0); t0 = t1; }
    };
    if (maxSlots < 16 || maxSlots > DC_SLOT_LIMIT) return -1;
    Mesh m{conn, nbElem, nbNodes, row, col, colBase, {}, nullptr, {}};
    build_node_to_elem(m);

    lap("node->elem");
    const int nt = nbThreads > 0 ? nbThreads : omp_get_max_threads();
    Selector sel{m, maxSlots, {}, {}, {}, {}, true, 0};
    sel.seedsT.resize((size_t)nt);
    sel.bigT.resize((size_t)nt);
    sel.seenT.resize((size_t)nt);
    if (const char *g = getenv("DC_PLAN_GREEDY_SPLIT")) sel.evenSplit = !(*g && *g != '0');
    if (const char *g = getenv("DC_PLAN_MAX_ROWS")) sel.maxRows = atoi(g) > 0 ? atoi(g) : 0;
    if (treeRec && nbTreeRec > 0) {
        int64_t cursor = 0;
        index_tree(treeRec, nbTreeRec, cursor, sel.tree, true);
        const char *win = getenv("DC_PLAN_UNIT_NODES");
        if (win && atoi(win) > 0) {
            const int width = atoi(win);
            const int chunk = width * 64;
            #pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
            for (int c = 0; c < (nbNodes + chunk - 1) / chunk; c++)
                sel.window_rows(c * chunk, std::min(nbNodes, (c + 1) * chunk), width);
        } else if (getenv("DC_PLAN_SUBTREE_UNITS")) {
            #pragma omp parallel num_threads(nt)
            #pragma omp single
            sel.select(0);
        } else {
            std::vector<std::pair<int, int>> leaves;           /* (firstNode, nodeCount) in ascending node order */
            for (const TreeIdx &t : sel.tree)
                if (!t.isSep && t.isLeaf) leaves.emplace_back(t.firstNode, t.lastNode - t.firstNode + 1);
            std::sort(leaves.begin(), leaves.end());
            const size_t nbChunks = std::min<size_t>(leaves.size(), (size_t)nt * 16);
            #pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
            for (size_t c = 0; c < nbChunks; c++)
                sel.pack_leaves(leaves, leaves.size() * c / nbChunks, leaves.size() * (c + 1) / nbChunks);
        }
    } else {
        sel.split_rows(0, nbNodes);          /* no tree: plain row blocks */
    }
    lap("select");
    std::vector<UnitSeed> seeds;
    std::vector<int> bigRows;
    for (int t = 0; t < nt; t++) {
        for (UnitSeed &u : sel.seedsT[(size_t)t]) seeds.push_back(std::move(u));
        bigRows.insert(bigRows.end(), sel.bigT[(size_t)t].begin(), sel.bigT[(size_t)t].end());
    }
    std::sort(seeds.begin(), seeds.end(), [](const UnitSeed &a, const UnitSeed &b) { return a.nodeFirst < b.nodeFirst; });
    std::sort(bigRows.begin(), bigRows.end());

    /* every row must belong to exactly one unit or to the fallback list */
    {
        std::vector<char> covered((size_t)nbNodes, 0);
        for (const UnitSeed &s : seeds)
            for (int n = s.nodeFirst; n < s.nodeFirst + s.nodeCount; n++) covered[n]++;
        for (int n : bigRows) covered[n]++;
        for (int n = 0; n < nbNodes; n++) if (covered[n] != 1) return -3;
    }

    lap("merge+check");
    const int64_t nbUnits = (int64_t)seeds.size();
    std::vector<dc_unit_t> units((size_t)nbUnits);
    std::vector<int32_t> diagRel((size_t)nbNodes, -1);
    std::vector<std::vector<dc_task_t>> tTasks((size_t)nt);
    std::vector<std::vector<uint16_t>> tItems((size_t)nt);
    std::vector<std::vector<int32_t>> tTgt((size_t)nt);
    std::vector<std::vector<uint16_t>> tSlot((size_t)nt);
    std::vector<std::vector<int32_t>> tSlotElem((size_t)nt);
    const char *noBank = getenv("DC_PLAN_NO_BANK_NUMBERING");
    const bool bankAware = !(noBank && *noBank && *noBank != '0');
    /* passes of the rank re-deal for the record build (DC_PLAN_REGROUP=0: slots keep the element
     * order inside a class).  Modelled on the plan (tools/smem_model.py, 64^3 elasticity): coordinate
     * reads of the record build 3.10 -> 2.81 wavefronts per element with one pass, no gain from more. */
    const char *rg = getenv("DC_PLAN_REGROUP");
    const int regroupPasses = rg ? atoi(rg) : 1;
    std::vector<std::vector<int32_t>> tHnode((size_t)nt);
    std::vector<int> unitThread((size_t)nbUnits);
    int error = 0;
    int64_t useful = 0;

    #pragma omp parallel num_threads(nt) reduction(|:error) reduction(+:useful)
    {
        const int tid = omp_get_thread_num();
        UnitOut out;
        BankOpt opt;
        opt.model = g_bankModel;
        opt.nOff = g_bankModel == 1 ? 10 : 4;
        Scratch work;
        #pragma omp for schedule(dynamic, 64)
        for (int64_t u = 0; u < nbUnits; u++) {
            const UnitSeed &s = seeds[(size_t)u];
            out.tasks.clear();
            out.items.clear();
            out.tgt.clear();
            out.slotNodes.clear();
            out.haloNodes.clear();
            out.slotElems.clear();
            out.useful = 0;
            build_unit(m, s, diagRel.data(), out, work, error, symmetric != 0, opt, bankAware, regroupPasses);
            const int32_t nbUpper = (int32_t)(out.tgt.size() / 2);
            while (out.tgt.size() % 4) out.tgt.push_back(0);      /* 16-byte granules for the bulk copy */
            dc_unit_t &d = units[(size_t)u];
            memset(&d, 0, sizeof d);
            d.edgeFirst = row[s.nodeFirst];
            d.edgeCount = row[s.nodeFirst + s.nodeCount] - row[s.nodeFirst];
            d.nodeFirst = s.nodeFirst;
            d.nodeCount = s.nodeCount;
            d.ownElemFirst = s.ownFirst;
            d.ownElemCount = s.ownCount;
            d.haloCount = (int32_t)s.halo.size();
            d.taskFirst = (int32_t)tTasks[tid].size();     /* thread-local, fixed up below */
            d.taskCount = (int32_t)out.tasks.size();
            d.itemFirst = (int64_t)tItems[tid].size();
            d.itemCount = (int32_t)out.items.size();
            d.tgtFirst = (int64_t)(tTgt[tid].size() / 2);
            d.tgtCount = nbUpper;
            d.slotFirst = (int64_t)(tSlot[tid].size() / 4);
            d.hnodeFirst = (int64_t)tHnode[tid].size();
            {
                int hn = (int)out.haloNodes.size();          /* padded; the true count is recovered below */
                d.hnodeCount = hn;
            }
            tSlot[tid].insert(tSlot[tid].end(), out.slotNodes.begin(), out.slotNodes.end());
            tSlotElem[tid].insert(tSlotElem[tid].end(), out.slotElems.begin(), out.slotElems.end());
            tHnode[tid].insert(tHnode[tid].end(), out.haloNodes.begin(), out.haloNodes.end());
            tTgt[tid].insert(tTgt[tid].end(), out.tgt.begin(), out.tgt.end());
            unitThread[(size_t)u] = tid;
            tTasks[tid].insert(tTasks[tid].end(), out.tasks.begin(), out.tasks.end());
            tItems[tid].insert(tItems[tid].end(), out.items.begin(), out.items.end());
            useful += out.useful;
        }
        #pragma omp critical
        if (verbose && tid == 0)
            fprintf(stderr, "[dc_plan] thread 0: contributions %.2fs slot classes %.2fs node tables %.2fs re-deal %.2fs "
                    "halo placement %.2fs output %.2fs\n", work.phase[0], work.phase[1], work.phase[2], work.phase[3],
                    work.phase[4], work.phase[5]);
    }
    if (error) return -2;
    lap("units");

    std::vector<int64_t> taskBase((size_t)nt + 1, 0), itemBase((size_t)nt + 1, 0), tgtBase((size_t)nt + 1, 0);
    std::vector<int64_t> slotBase((size_t)nt + 1, 0), hnodeBase((size_t)nt + 1, 0);
    for (int t = 0; t < nt; t++) {
        taskBase[(size_t)t + 1] = taskBase[t] + (int64_t)tTasks[t].size();
        itemBase[(size_t)t + 1] = itemBase[t] + (int64_t)tItems[t].size();
        tgtBase[(size_t)t + 1] = tgtBase[t] + (int64_t)(tTgt[t].size() / 2);
        slotBase[(size_t)t + 1] = slotBase[t] + (int64_t)(tSlot[t].size() / 4);
        hnodeBase[(size_t)t + 1] = hnodeBase[t] + (int64_t)tHnode[t].size();
    }
    plan->nbSlotNodes = slotBase[nt] * 4;
    plan->slotNodes = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)(plan->nbSlotNodes ? plan->nbSlotNodes : 1));
    plan->nbHaloNodes = hnodeBase[nt];
    plan->haloNodes = (int32_t *)malloc(sizeof(int32_t) * (size_t)(plan->nbHaloNodes ? plan->nbHaloNodes : 1));
    plan->slotElems = (int32_t *)malloc(sizeof(int32_t) * (size_t)(slotBase[nt] ? slotBase[nt] : 1));
    plan->tgt = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(tgtBase[nt] ? tgtBase[nt] : 1));
    plan->tasks = (dc_task_t *)malloc(sizeof(dc_task_t) * (size_t)(taskBase[nt] ? taskBase[nt] : 1));
    plan->items = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)(itemBase[nt] ? itemBase[nt] : 1));
    /* every thread's pieces go to their place in the final arrays (disjoint ranges: copied in parallel) */
    #pragma omp parallel for schedule(static, 1) num_threads(nt)
    for (int t = 0; t < nt; t++) {
        if (!tSlot[t].empty()) memcpy(plan->slotNodes + 4 * slotBase[t], tSlot[t].data(), sizeof(uint16_t) * tSlot[t].size());
        if (!tSlotElem[t].empty()) memcpy(plan->slotElems + slotBase[t], tSlotElem[t].data(), sizeof(int32_t) * tSlotElem[t].size());
        if (!tHnode[t].empty()) memcpy(plan->haloNodes + hnodeBase[t], tHnode[t].data(), sizeof(int32_t) * tHnode[t].size());
        if (!tTgt[t].empty()) memcpy(plan->tgt + 2 * tgtBase[t], tTgt[t].data(), sizeof(int32_t) * tTgt[t].size());
        if (!tTasks[t].empty()) memcpy(plan->tasks + taskBase[t], tTasks[t].data(), sizeof(dc_task_t) * tTasks[t].size());
        if (!tItems[t].empty()) memcpy(plan->items + itemBase[t], tItems[t].data(), sizeof(uint16_t) * tItems[t].size());
        std::vector<int32_t>().swap(tSlotElem[t]);
        std::vector<uint16_t>().swap(tSlot[t]);
        std::vector<int32_t>().swap(tHnode[t]);
        std::vector<int32_t>().swap(tTgt[t]);
        std::vector<dc_task_t>().swap(tTasks[t]);
        std::vector<uint16_t>().swap(tItems[t]);
    }
    plan->symmetric = symmetric ? 1 : 0;
    plan->nbTgt = tgtBase[nt];
    if (taskBase[nt] >= INT32_MAX) return -4;
    plan->nbTasks = taskBase[nt];
    plan->nbItems = itemBase[nt];
    int64_t haloTotal = 0, slotsTotal = 0;
    int maxUsed = 0, maxHalo = 0, maxTasks = 0, maxItems = 0, maxTgt = 0, maxNodes = 0;
    for (int64_t u = 0; u < nbUnits; u++) {
        dc_unit_t &d = units[(size_t)u];
        const int t = unitThread[(size_t)u];
        d.taskFirst += (int32_t)taskBase[t];
        d.itemFirst += itemBase[t];
        d.tgtFirst += tgtBase[t];
        d.slotFirst += slotBase[t];
        d.hnodeFirst += hnodeBase[t];
        maxTgt = std::max(maxTgt, (d.tgtCount + 1) & ~1);
        maxNodes = std::max(maxNodes, ((d.nodeFirst + d.nodeCount + 1) & ~1) - (d.nodeFirst & ~1));
        haloTotal += d.haloCount;
        slotsTotal += d.ownElemCount + d.haloCount;
        maxUsed = std::max(maxUsed, d.ownElemCount + d.haloCount);
        maxHalo = std::max(maxHalo, d.hnodeCount);
        maxTasks = std::max(maxTasks, d.taskCount);
        maxItems = std::max(maxItems, d.itemCount);
    }
    plan->maxHalo = maxHalo;
    plan->maxTasks = maxTasks;
    plan->maxItems = maxItems;
    plan->maxTgt = maxTgt;
    plan->maxNodes = maxNodes;
    lap("concat");
    plan->nbHalo = haloTotal;
    plan->nbUnits = nbUnits;
    plan->units = dup(units);
    plan->diagRel = dup(diagRel);
    plan->maxSlots = maxUsed;
    plan->nbSlotsTotal = slotsTotal;
    plan->nbItemsUseful = useful;
    if (!bigRows.empty()) {
        /* hub rows go to the per-edge lists, whose items hold the element in 28 bits: -5 beyond that */
        int rc = dc_contrib_build(&plan->big, conn, nbElem, nbNodes, row, col, colBase,
                                  bigRows.data(), (int64_t)bigRows.size(), symmetric);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int64_t dc_plan_order_units(dc_plan_t *plan, const unsigned char *nodeFlag)
{
    if (!plan || !nodeFlag || plan->nbUnits <= 0) return 0;
    std::vector<dc_unit_t> front, rest;
    for (int64_t u = 0; u < plan->nbUnits; u++) {
        const dc_unit_t &d = plan->units[u];
        bool hit = false;
        for (int n = d.nodeFirst; n < d.nodeFirst + d.nodeCount && !hit; n++) hit = nodeFlag[n] != 0;
        (hit ? front : rest).push_back(d);
    }
    std::copy(front.begin(), front.end(), plan->units);
    std::copy(rest.begin(), rest.end(), plan->units + front.size());
    return (int64_t)front.size();
}

extern "C" void dc_plan_set_bank_model(int model) { g_bankModel = model == 1 ? 1 : 0; }

extern "C" void dc_contrib_free(dc_contrib_list_t *c)
{
    free(c->edge); free(c->ptr); free(c->item); free(c->edgeT);
    memset(c, 0, sizeof *c);
}

extern "C" void dc_plan_free(dc_plan_t *plan)
{
    free(plan->units); free(plan->tasks); free(plan->items); free(plan->slotElems); free(plan->diagRel);
    free(plan->tgt); free(plan->slotNodes); free(plan->haloNodes);
    dc_contrib_free(&plan->big);
    memset(plan, 0, sizeof *plan);
}
