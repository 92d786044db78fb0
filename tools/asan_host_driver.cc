#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include "DC.h"
#include "DC_device.h"
#include "dc_plan.h"
extern "C" void dcx_kuhn_cube(int n, int *conn, double *coord, double amplitude, int seed);
extern "C" int64_t dcx_build_csr(const int *conn, int nbElem, int nbNodes, int *row, int *col, int colBase);
int main(int argc, char **argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 20, cap = argc > 2 ? atoi(argv[2]) : 169, slots = argc > 3 ? atoi(argv[3]) : 512;
    int geo = argc > 4 ? atoi(argv[4]) : 1, model = argc > 5 ? atoi(argv[5]) : 0, sym = argc > 6 ? atoi(argv[6]) : 1;
    int E = 6 * n * n * n, N = (n + 1) * (n + 1) * (n + 1);
    std::vector<int> conn((size_t)E * 4);
    std::vector<double> coord((size_t)N * 3);
    dcx_kuhn_cube(n, conn.data(), coord.data(), 0.3, 12345);
    DC_set_max_elem_per_part(cap);
    if (geo) DC_create_tree_geometric(conn.data(), E, 4, N, coord.data(), 3);
    else DC_create_tree(conn.data(), E, 4, N);
    DC_permute_double_2d_array(coord.data(), N, 3);
    DC_renumber_int_array(conn.data(), 4 * E, true);
    std::vector<int> row((size_t)N + 1);
    int64_t nnz = dcx_build_csr(conn.data(), E, N, row.data(), nullptr, 0);
    std::vector<int> col((size_t)nnz);
    dcx_build_csr(conn.data(), E, N, row.data(), col.data(), 0);
    DC_finalize_tree(row.data(), conn.data(), N);
    int64_t nrec = dc_flatten_tree(nullptr, 0);
    std::vector<int> rec((size_t)nrec * 8);
    dc_flatten_tree(rec.data(), nrec);
    dc_plan_t plan;
    dc_plan_set_bank_model(model);
    int rc = dc_plan_build(&plan, conn.data(), E, N, row.data(), col.data(), 0, rec.data(), nrec, slots, 0, sym);
    printf("n=%d cap=%d slots=%d geo=%d model=%d sym=%d rc=%d units=%lld items=%lld big=%lld\n", n, cap, slots, geo, model, sym, rc,
           (long long)plan.nbUnits, (long long)plan.nbItems, (long long)plan.big.nbEdges);
    dc_plan_free(&plan);
    DC_free_tree();
    return rc;
}
