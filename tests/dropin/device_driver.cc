/* tests/dropin/device_driver.cc -- the same USER program as driver.cc after the switch to the
 * device path: the per-step DC_tree_traversal with host callbacks is kept as the program's own
 * check, and the assembly itself goes through DC_device_set_mesh / DC_assembly_device
 * (include/DC_device.h) with the built-in P1 Laplacian.  Prints the largest relative difference
 * between the two (the host callback here is plain C++ arithmetic, so it differs from the device
 * operator's fixed rounding sequence in the last bits).
 *
 * usage: device_driver <n> <out-prefix>
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>
#include "DC.h"
#include "DC_device.h"

struct User {
    int *elemToNode;
    double *coord;
    int *row, *col;
    double *values;
};

/* user callback: P1 Laplacian, element matrix then scatter-add (plain C++, no library code) */
static void assemble_range(void *args, DCargs_t *dc, bool reset)
{
    User *u = (User *)args;
    if (reset && !dc->isSep)
        for (int k = dc->firstEdge; k <= dc->lastEdge; k++) u->values[k] = 0.0;
    for (int e = dc->firstElem; e <= dc->lastElem; e++) {
        const int *c = u->elemToNode + 4 * e;
        double x[4][3];
        for (int a = 0; a < 4; a++)
            for (int k = 0; k < 3; k++) x[a][k] = u->coord[3 * (c[a] - 1) + k];
        double m[3][3], g[4][3];
        for (int k = 0; k < 3; k++) { m[0][k] = x[1][k] - x[0][k]; m[1][k] = x[2][k] - x[0][k]; m[2][k] = x[3][k] - x[0][k]; }
        const double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                           m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
        g[1][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) / det; g[1][1] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) / det;
        g[1][2] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) / det;
        g[2][0] = (m[2][1] * m[0][2] - m[2][2] * m[0][1]) / det; g[2][1] = (m[2][2] * m[0][0] - m[2][0] * m[0][2]) / det;
        g[2][2] = (m[2][0] * m[0][1] - m[2][1] * m[0][0]) / det;
        g[3][0] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) / det; g[3][1] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) / det;
        g[3][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) / det;
        for (int k = 0; k < 3; k++) g[0][k] = -(g[1][k] + g[2][k] + g[3][k]);
        const double vol = std::fabs(det) / 6.0;
        for (int a = 0; a < 4; a++)
            for (int b = 0; b < 4; b++) {
                const double ke = vol * (g[a][0] * g[b][0] + g[a][1] * g[b][1] + g[a][2] * g[b][2]);
                const int i = c[a] - 1, j = c[b];            /* columns are 1-based */
                int k = u->row[i];
                while (u->col[k] != j) k++;
                u->values[k] += ke;
            }
    }
}
static void seq_fct(void *args, DCargs_t *dc)
{
#ifdef DC_VEC
    assemble_range(args, dc, false);      /* the vector part ran first and did the reset */
#else
    assemble_range(args, dc, true);
#endif
}
static void vec_fct(void *args, DCargs_t *dc) { assemble_range(args, dc, true); }

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    const int n = atoi(argv[1]), m = n + 1;
    const std::string prefix = argv[2];
    const int N = m * m * m, E = 6 * n * n * n;
    std::vector<double> coord((size_t)N * 3);
    std::vector<int> conn((size_t)E * 4);
    unsigned long long s = 88172645463325252ULL;
    for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++)
            for (int k = 0; k < m; k++) {
                const int id = (i * m + j) * m + k;
                const int ijk[3] = {i, j, k};
                for (int q = 0; q < 3; q++) {
                    s ^= s << 13; s ^= s >> 7; s ^= s << 17;                /* xorshift64 */
                    coord[3 * (size_t)id + q] = ijk[q] + 0.3 * ((double)(s % 100001) / 100000.0 - 0.5);
                }
            }
    /* Kuhn subdivision: the six monotone paths from (i,j,k) to (i+1,j+1,k+1) */
    static const int perm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    size_t e = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            for (int k = 0; k < n; k++)
                for (int p = 0; p < 6; p++) {
                    int v[3] = {i, j, k};
                    conn[4 * e] = (v[0] * m + v[1]) * m + v[2] + 1;
                    for (int q = 0; q < 3; q++) {
                        v[perm[p][q]]++;
                        conn[4 * e + q + 1] = (v[0] * m + v[1]) * m + v[2] + 1;
                    }
                    e++;
                }

    DC_create_tree(conn.data(), E, 4, N);
    DC_permute_double_2d_array(coord.data(), N, 3);
    DC_renumber_int_array(conn.data(), 4 * E, true);

    /* user-side CSR pattern: 0-based row pointers, 1-based ascending columns, diagonal included */
    std::vector<std::vector<int>> nb((size_t)N);
    for (int el = 0; el < E; el++)
        for (int a = 0; a < 4; a++)
            for (int b = 0; b < 4; b++) nb[(size_t)conn[4 * (size_t)el + a] - 1].push_back(conn[4 * (size_t)el + b]);
    std::vector<int> row((size_t)N + 1, 0), col;
    for (int i = 0; i < N; i++) {
        std::sort(nb[i].begin(), nb[i].end());
        nb[i].erase(std::unique(nb[i].begin(), nb[i].end()), nb[i].end());
        row[(size_t)i + 1] = row[i] + (int)nb[i].size();
        col.insert(col.end(), nb[i].begin(), nb[i].end());
    }
    int nbDCcomm = 0;
    DC_finalize_tree(row.data(), conn.data(), nullptr, nullptr, nullptr, &nbDCcomm, E, 4, 1, 0, 0);

    std::vector<double> values(col.size(), NAN);
    User u{conn.data(), coord.data(), row.data(), col.data(), values.data()};
    const double t0 = DC_get_time();
    for (int step = 0; step < 2; step++)            /* two time steps: the reset must make them equal */
        DC_tree_traversal(seq_fct, vec_fct, nullptr, &u, nullptr);
    const double t1 = DC_get_time();
    if (!(t1 >= t0)) return 3;

    /* the switch: device operator instead of host callbacks (columns here are 1-based) */
    std::vector<double> dvalues(col.size(), NAN);
    DC_device_init(0);
    DC_device_set_mesh(coord.data(), conn.data(), row.data(), col.data(), E, N, 1);
    for (int step = 0; step < 2; step++)
        DC_assembly_device(DC_OP_LAPLACIAN, nullptr, 0, dvalues.data(), 1);
    double scale = 0, worst = 0;
    for (double v : values) scale = std::max(scale, std::fabs(v));
    for (size_t k = 0; k < values.size(); k++) worst = std::max(worst, std::fabs(values[k] - dvalues[k]) / scale);
    /* elasticity: every 3x3 block row sums to zero over a row (rigid translations) */
    std::vector<double> ev(col.size() * 9, NAN);
    const double lame[2] = {0.5769230769230769, 0.3846153846153846};
    DC_assembly_device(DC_OP_ELASTICITY, lame, 2, ev.data(), 3);
    double escale = 0, eworst = 0;
    for (double v : ev) escale = std::max(escale, std::fabs(v));
    for (int i = 0; i < N; i++) {
        double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = row[i]; k < row[(size_t)i + 1]; k++)
            for (int q = 0; q < 9; q++) acc[q] += ev[(size_t)k * 9 + q];
        for (int q = 0; q < 9; q++) eworst = std::max(eworst, std::fabs(acc[q]) / escale);
    }
    DC_device_finalize();
    printf("device_vs_host_rel=%.3e elasticity_rowsum_rel=%.3e\n", worst, eworst);

    std::string treePath = prefix + ".tree";
    FILE *f = fopen((prefix + ".bin").c_str(), "wb");
    if (!f) return 4;
    fwrite(coord.data(), sizeof(double), coord.size(), f);
    fwrite(conn.data(), sizeof(int), conn.size(), f);
    fwrite(row.data(), sizeof(int), row.size(), f);
    fwrite(values.data(), sizeof(double), values.size(), f);
    fclose(f);
    DC_store_tree(treePath, E, N, 0, 0);
    double sum = 0;
    for (double v : values) sum += v;
    printf("E=%d N=%d nnz=%zu sum=%.17g\n", E, N, col.size(), sum);
    return 0;
}
