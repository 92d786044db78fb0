/* dc_interface.cc -- from the reference's interface layout to CSR edges.
 *
 * The reference describes what two domains share by NODES: per neighbour i the slice
 * intfNodes[intfIndex[i] .. intfIndex[i+1]) of 1-based local node ids, listed in the same order on
 * both sides (include/DC.h:51-53,141-143; src/tree_creation.cc:48-111 cuts these lists per leaf).
 * The assembled values that have to be summed across the two domains are the CSR entries (r, c)
 * with both nodes in that list.  Ordering them by (position of r, position of c) in the list gives
 * both sides the same order without any further agreement. */
#include <algorithm>
#include <cstdint>
#include <unordered_map>
#include <vector>
#include "dc_cuda.h"

extern "C" int64_t dc_interface_edges(const int *h_row, const int *h_col, int colBase, const int *nodes, int nbListed,
                                      int64_t *edges, int64_t *keys)
{
    if (!h_row || !h_col || !nodes || nbListed <= 0) return 0;
    std::unordered_map<int, int> posOf;
    posOf.reserve((size_t)nbListed * 2);
    for (int p = 0; p < nbListed; p++) posOf.emplace(nodes[p] - 1, p);
    std::vector<std::pair<int64_t, int64_t>> found;       /* (key, edge) */
    for (int p = 0; p < nbListed; p++) {
        const int r = nodes[p] - 1;
        for (int k = h_row[r]; k < h_row[r + 1]; k++) {
            const auto it = posOf.find(h_col[k] - colBase);
            if (it != posOf.end()) found.emplace_back((int64_t)p * nbListed + it->second, (int64_t)k);
        }
    }
    std::sort(found.begin(), found.end());
    for (size_t i = 0; i < found.size(); i++) {
        if (keys) keys[i] = found[i].first;
        if (edges) edges[i] = found[i].second;
    }
    return (int64_t)found.size();
}
