/* dc_perm.cc -- permutation primitives of the DC-lib API.
 *
 * Semantics follow reference src/permutations.cc: every permute is a SCATTER,
 * new[perm[i]] = old[i] (permutations.cc:22-104), renumbering maps values through
 * nodePerm (permutations.cc:107-115) and DC_create_permutation is the stable
 * counting order of a partition vector (permutations.cc:169-180).  The reference
 * chases cycles serially with O(size*nbPart) / O(nbItem) scans; here the scatter
 * goes through one out-of-place pass (parallel over items) and the partition
 * order is a counting sort, which give the same arrays bit for bit.
 */
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include "dc_internal.h"

namespace dc {

void fatal(const char *msg)
{
    fprintf(stderr, "%s\n", msg);
    exit(EXIT_FAILURE);
}

/* perm[j] = rank of j in the order (part[j], j). */
void stable_partition_perm(int *perm, const int *part, int size, int nbPart)
{
    std::vector<int64_t> start((size_t)nbPart + 1, 0);
    for (int j = 0; j < size; j++) {
        int p = part[j];
        if (p >= 0 && p < nbPart) start[(size_t)p + 1]++;
    }
    for (int p = 0; p < nbPart; p++) start[(size_t)p + 1] += start[p];
    for (int j = 0; j < size; j++) {
        int p = part[j];
        /* entries whose part is outside [0,nbPart) are left untouched, as in the
         * reference's nested scan */
        if (p >= 0 && p < nbPart) perm[j] = (int)start[p]++;
    }
}

template <typename T>
void scatter_rows_inplace(T *tab, const int *perm, int64_t nbItem, int dimItem)
{
    if (nbItem <= 0 || dimItem <= 0) return;
    std::vector<T> src(tab, tab + nbItem * dimItem);
    #pragma omp parallel for schedule(static) if (nbItem > (1 << 16))
    for (int64_t i = 0; i < nbItem; i++) {
        const T *s = &src[(size_t)(i * dimItem)];
        T *d = tab + (int64_t)perm[i] * dimItem;
        for (int j = 0; j < dimItem; j++) d[j] = s[j];
    }
}
template void scatter_rows_inplace<int>(int *, const int *, int64_t, int);
template void scatter_rows_inplace<double>(double *, const int *, int64_t, int);

}  // namespace dc

void DC_permute_double_2d_array (double *tab, int nbItem, int dimItem)
{
    if (!nodePerm) dc::fatal("Error: DC_permute_double_2d_array called without a node permutation.");
    dc::scatter_rows_inplace<double>(tab, nodePerm, nbItem, dimItem);
}

void DC_permute_int_2d_array (int *tab, int *perm, int nbItem, int dimItem, int offset)
{
    if (perm == nullptr) perm = elemPerm;
    if (!perm) dc::fatal("Error: DC_permute_int_2d_array called without a permutation.");
    dc::scatter_rows_inplace<int>(tab + (int64_t)offset * dimItem, perm, nbItem, dimItem);
}

void DC_permute_int_1d_array (int *tab, int size)
{
    if (!nodePerm) dc::fatal("Error: DC_permute_int_1d_array called without a node permutation.");
    dc::scatter_rows_inplace<int>(tab, nodePerm, size, 1);
}

void DC_renumber_int_array (int *tab, int size, bool isFortran)
{
    if (!nodePerm) dc::fatal("Error: DC_renumber_int_array called without a node permutation.");
    const int base = isFortran ? 1 : 0;
    #pragma omp parallel for schedule(static) if (size > (1 << 16))
    for (int i = 0; i < size; i++) tab[i] = nodePerm[tab[i] - base] + base;
}

void DC_create_permutation (int *perm, int *part, int size, int nbPart)
{
    dc::stable_partition_perm(perm, part, size, nbPart);
}
