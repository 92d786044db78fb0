/* tests/dropin/multi_driver.cc -- a C++ USER program of the multi-domain device path: one process
 * per domain (the reference's users run one MPI/GASPI rank per domain), one GPU per process.
 *
 * Domain `rank` of `world` is one jittered Kuhn cube of a bar of cubes stacked along x.  The program
 * follows the reference's call order: DC_create_tree -> DC_permute_* /
 * DC_renumber_int_array (the interface node lists are renumbered like any node array) -> its own CSR
 * pattern -> DC_finalize_tree (interface arguments in the reference's layout, include/DC.h:141-143)
 * -> DC_device_set_interface / DC_device_set_mesh -> handles swapped through files (an MPI program
 * would use MPI_Sendrecv) -> DC_assembly_device per step.  It writes, for every local CSR entry, the
 * global (row, column) key and the assembled 3 x 3 block; the test compares them with the
 * single-domain oracle.
 *
 * usage: multi_driver <rank> <world> <n> <directory>
 */
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>
#include "DC.h"
#include "DC_device.h"

extern "C" void dcx_kuhn_slab(int n, int *conn, double *coord, double amplitude, int seed, int x_offset);

static bool read_file(const std::string &path, void *buf, size_t bytes)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    const size_t got = fread(buf, 1, bytes, f);
    fclose(f);
    return got == bytes;
}
static void write_file(const std::string &path, const void *buf, size_t bytes)
{
    const std::string tmp = path + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    fwrite(buf, 1, bytes, f);
    fclose(f);
    rename(tmp.c_str(), path.c_str());          /* readers never see a partial file */
}
static void wait_for(const std::string &path, void *buf, size_t bytes)
{
    for (int tries = 0; tries < 6000; tries++) {
        if (read_file(path, buf, bytes)) return;
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
    fprintf(stderr, "multi_driver: timed out waiting for %s\n", path.c_str());
    exit(3);
}

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    const int rank = atoi(argv[1]), world = atoi(argv[2]), n = atoi(argv[3]), m = n + 1;
    const std::string dir = argv[4];
    const int N = m * m * m, E = 6 * n * n * n;
    std::vector<double> coord((size_t)N * 3);
    std::vector<int> conn((size_t)E * 4);
    dcx_kuhn_slab(n, conn.data(), coord.data(), 0.3, 12345, rank * n);

    /* interface in the reference's layout: per neighbour the nodes of the shared plane, 1-based, both
     * sides in plane order (j, k) */
    std::vector<int> nbRank, intfIndex(1, 0), intfNodes;
    for (int side = 0; side < 2; side++) {
        const int nb = side == 0 ? rank - 1 : rank + 1;
        if (nb < 0 || nb >= world) continue;
        nbRank.push_back(nb);
        const int i = side == 0 ? 0 : n;
        for (int p = 0; p < m * m; p++) intfNodes.push_back(i * m * m + p + 1);
        intfIndex.push_back((int)intfNodes.size());
    }
    const int nbIntf = (int)nbRank.size();

    DC_create_tree_geometric(conn.data(), E, 4, N, coord.data(), 3);
    std::vector<int> nodePerm((size_t)N);
    dc_get_node_perm(nodePerm.data(), N);
    DC_permute_double_2d_array(coord.data(), N, 3);
    DC_renumber_int_array(conn.data(), 4 * E, true);
    if (!intfNodes.empty()) DC_renumber_int_array(intfNodes.data(), (int)intfNodes.size(), true);

    std::vector<std::vector<int>> adj((size_t)N);
    for (int el = 0; el < E; el++)
        for (int a = 0; a < 4; a++)
            for (int b = 0; b < 4; b++) adj[(size_t)conn[4 * (size_t)el + a] - 1].push_back(conn[4 * (size_t)el + b] - 1);
    std::vector<int> row((size_t)N + 1, 0), col;
    for (int i = 0; i < N; i++) {
        std::sort(adj[i].begin(), adj[i].end());
        adj[i].erase(std::unique(adj[i].begin(), adj[i].end()), adj[i].end());
        row[(size_t)i + 1] = row[i] + (int)adj[i].size();
        col.insert(col.end(), adj[i].begin(), adj[i].end());
    }
    const int nnz = row[N];
    int nbDCcomm = -1;
    std::vector<int> dst(intfNodes.size(), 0);
    DC_finalize_tree(row.data(), conn.data(), intfIndex.data(), intfNodes.data(), dst.data(), &nbDCcomm, E, 4, 1, nbIntf, rank);

    DC_device_init(rank);                        /* one GPU per domain */
    DC_device_set_interface(nbIntf, nbRank.data(), intfIndex.data(), intfNodes.data());
    DC_device_set_mesh(coord.data(), conn.data(), row.data(), col.data(), E, N, 0);
    DC_device_prepare_exchange(3);
    for (int i = 0; i < nbIntf; i++) {
        unsigned char h[128];
        DC_device_exchange_handle(i, h);
        write_file(dir + "/handle_" + std::to_string(rank) + "_to_" + std::to_string(nbRank[i]) + ".bin", h, sizeof h);
    }
    for (int i = 0; i < nbIntf; i++) {
        unsigned char h[128];
        wait_for(dir + "/handle_" + std::to_string(nbRank[i]) + "_to_" + std::to_string(rank) + ".bin", h, sizeof h);
        DC_device_exchange_connect(i, h);
    }

    const double prm[2] = {0.5769230769230769, 0.3846153846153846};
    std::vector<double> values((size_t)nnz * 9, -1.0), first;
    for (int step = 0; step < 3; step++) {
        DC_assembly_device(DC_OP_ELASTICITY, prm, 2, values.data(), 3);
        if (step == 0) first = values;
        else if (first != values) { fprintf(stderr, "multi_driver: step %d differs from step 0\n", step); return 4; }
    }

    /* global keys: node `old` of this cube is node old + rank * n * m * m of the bar */
    std::vector<int64_t> gid((size_t)N);
    for (int old = 0; old < N; old++) gid[(size_t)nodePerm[old]] = (int64_t)old + (int64_t)rank * n * m * m;
    const int64_t Nglob = (int64_t)(world * n + 1) * m * m;
    std::vector<int64_t> keys((size_t)nnz);
    for (int i = 0; i < N; i++)
        for (int k = row[i]; k < row[(size_t)i + 1]; k++) keys[(size_t)k] = gid[(size_t)i] * Nglob + gid[(size_t)col[(size_t)k]];
    write_file(dir + "/keys_" + std::to_string(rank) + ".bin", keys.data(), keys.size() * sizeof(int64_t));
    write_file(dir + "/values_" + std::to_string(rank) + ".bin", values.data(), values.size() * sizeof(double));
    /* nobody tears its buffers down while a neighbour may still push into them */
    write_file(dir + "/done_" + std::to_string(rank) + ".bin", &rank, sizeof rank);
    for (int r = 0; r < world; r++) { int x; wait_for(dir + "/done_" + std::to_string(r) + ".bin", &x, sizeof x); }
    DC_device_finalize();
    printf("RANK_OK %d E=%d N=%d nnz=%d interfaces=%d nbDCcomm=%d\n", rank, E, N, nnz, nbIntf, nbDCcomm);
    return 0;
}
