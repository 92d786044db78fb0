/* dc_meshgen.cc -- user-side helpers for large synthetic runs (C ABI, OpenMP):
 * the jittered Kuhn cube of SURVEY.md 8(d) and the CSR pattern of a P1 mesh.  In the
 * reference these belong to the application (Mini-FEM), not to DC-lib; they live here so
 * that 10^8-element benchmarks do not spend minutes in numpy.  Results are identical to
 * dc-lib_b200/meshgen.py (tests/test_meshgen.py). */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <omp.h>

extern "C" {

/* conn: [6 n^3][4] 1-based; coord: [(n+1)^3][3].  xOffset shifts the cube by that many
 * cells along x inside a longer bar of cubes: coordinates and the jitter hash then use the
 * GLOBAL node id ((i+xOffset)*(n+1)+j)*(n+1)+k, so that neighbouring slabs agree on the
 * nodes of their common plane (multi-GPU decomposition). */
void dcx_kuhn_slab(int n, int *conn, double *coord, double amplitude, int seed, int xOffset);

void dcx_kuhn_cube(int n, int *conn, double *coord, double amplitude, int seed)
{
    dcx_kuhn_slab(n, conn, coord, amplitude, seed, 0);
}

void dcx_kuhn_slab(int n, int *conn, double *coord, double amplitude, int seed, int xOffset)
{
    static const int order[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    const int64_t m = n + 1, cells = (int64_t)n * n * n;
    if (conn) {
        #pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < cells; c++) {
            const int64_t i = c / ((int64_t)n * n), j = (c / n) % n, k = c % n;
            for (int p = 0; p < 6; p++) {
                int64_t cur[3] = {i, j, k};
                int *row = conn + (c * 6 + p) * 4;
                row[0] = (int)((cur[0] * m + cur[1]) * m + cur[2]) + 1;
                for (int s = 0; s < 3; s++) {
                    cur[order[p][s]]++;
                    row[s + 1] = (int)((cur[0] * m + cur[1]) * m + cur[2]) + 1;
                }
            }
        }
    }
    if (coord) {
        const int64_t N = m * m * m;
        #pragma omp parallel for schedule(static)
        for (int64_t id = 0; id < N; id++) {
            const double grid[3] = {(double)(id / (m * m) + xOffset), (double)((id / m) % m), (double)(id % m)};
            const int64_t gid = id + (int64_t)xOffset * m * m;
            for (int comp = 0; comp < 3; comp++) {
                int64_t h = (gid * 3 + comp + seed) & 0x7FFFFFFF;
                for (int r = 0; r < 3; r++) {
                    h = (h * 1103515245LL + 12345) & 0x7FFFFFFF;
                    h ^= h >> 13;
                }
                coord[id * 3 + comp] = grid[comp] + ((double)h / 2147483648.0 - 0.5) * amplitude;
            }
        }
    }
}

/* node -> incident elements (ascending), conn 1-based */
static void node_to_elem(const int *conn, int64_t nbElem, int nbNodes, std::vector<int64_t> &ptr,
                         std::vector<int> &val)
{
    ptr.assign((size_t)nbNodes + 1, 0);
    #pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nbElem * 4; k++) {
        #pragma omp atomic
        ptr[(size_t)conn[k]]++;
    }
    for (int n = 0; n < nbNodes; n++) ptr[(size_t)n + 1] += ptr[n];
    val.resize((size_t)(nbElem * 4));
    std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
    #pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < nbElem; e++)
        for (int j = 0; j < 4; j++) {
            int64_t slot;
            #pragma omp atomic capture
            slot = fill[(size_t)conn[e * 4 + j] - 1]++;
            val[(size_t)slot] = (int)e;
        }
    #pragma omp parallel for schedule(dynamic, 4096)
    for (int n = 0; n < nbNodes; n++) std::sort(val.begin() + ptr[n], val.begin() + ptr[(size_t)n + 1]);
}

/* CSR pattern: edge (i,j) iff an element holds both; columns ascending, diagonal included.
 * Call with col == NULL to get row[] (and nnz as return value), then again to fill col. */
int64_t dcx_build_csr(const int *conn, int nbElem, int nbNodes, int *row, int *col, int colBase)
{
    std::vector<int64_t> ptr;
    std::vector<int> val;
    node_to_elem(conn, nbElem, nbNodes, ptr, val);
    std::vector<int> count((size_t)nbNodes + 1, 0);
    if (!col) {
        #pragma omp parallel
        {
            std::vector<int> nb;
            #pragma omp for schedule(dynamic, 4096)
            for (int n = 0; n < nbNodes; n++) {
                nb.clear();
                for (int64_t q = ptr[n]; q < ptr[(size_t)n + 1]; q++)
                    for (int j = 0; j < 4; j++) nb.push_back(conn[(int64_t)val[(size_t)q] * 4 + j] - 1);
                std::sort(nb.begin(), nb.end());
                count[(size_t)n + 1] = (int)(std::unique(nb.begin(), nb.end()) - nb.begin());
            }
        }
        int64_t total = 0;
        row[0] = 0;
        for (int n = 0; n < nbNodes; n++) {
            total += count[(size_t)n + 1];
            if (total > INT32_MAX) return -1;
            row[n + 1] = (int)total;
        }
        return total;
    }
    #pragma omp parallel
    {
        std::vector<int> nb;
        #pragma omp for schedule(dynamic, 4096)
        for (int n = 0; n < nbNodes; n++) {
            nb.clear();
            for (int64_t q = ptr[n]; q < ptr[(size_t)n + 1]; q++)
                for (int j = 0; j < 4; j++) nb.push_back(conn[(int64_t)val[(size_t)q] * 4 + j] - 1);
            std::sort(nb.begin(), nb.end());
            const int k = (int)(std::unique(nb.begin(), nb.end()) - nb.begin());
            for (int q = 0; q < k; q++) col[row[n] + q] = nb[q] + colBase;
        }
    }
    return row[nbNodes];
}

}
