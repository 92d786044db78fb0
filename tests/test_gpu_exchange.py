"""dc_cuda_exchange_* (include/dc_cuda.h): interface exchange over peer memory, and the pieces of the
overlapped multi-GPU step (front units first, range launches), with fewer GPUs than ranks: all logical
ranks live in this process and share one stream, every rank packs before any rank unpacks (a kernel
that waits for a flag must never be queued ahead of the kernel that raises it).  The real
one-process-per-GPU run is tests/test_gpu_multi.py (needs two GPUs)."""
import numpy as np
import pytest

import dc_lib_b200 as dc
from common import LAMBDA, MU, oracle_serial

pytestmark = pytest.mark.gpu
REL_TOL = 1e-12


def _check_against_global(parts, vals, key_ref, want, N_glob, per):
    for p, v in zip(parts, vals):
        rows = np.repeat(np.arange(p["N"]), np.diff(p["row"]))
        key = p["gid_of_new"][rows] * N_glob + p["gid_of_new"][p["col"]]
        pos = np.searchsorted(key_ref, key)
        assert np.array_equal(key_ref[pos], key)
        got = v.cpu().numpy().reshape(-1, per)
        assert np.abs(got - want[pos]).max() <= REL_TOL * np.abs(want).max()


@pytest.mark.parametrize("front_first", [False, True])
def test_exchange_of_slabs_in_one_process(front_first):
    """Two slabs: per-rank values are bit-exact partial sums, after the exchange every entry equals the
    oracle on the merged mesh; with front_first the units that own interface rows are assembled by a
    first launch, the exchange is packed, and the other units follow (the order of the overlapped step)."""
    import torch
    from dc_lib_b200 import multigpu
    from test_multirank_gloo import _global_reference
    n, world, dim = 6, 2, 3
    stream = torch.cuda.current_stream().cuda_stream
    slabs, vals, ctxs, xs, fronts = [], [], [], {}, []
    for r in range(world):
        s = multigpu.build_slab(n, r, world, builder="geometric")
        plan = dc.HostPlan(s["conn"], s["N"], s["row"], s["col"], 0, s["rec"], 288)
        flag = np.zeros(s["N"], dtype=np.uint8)
        for e in s["intf"].values():
            flag[np.searchsorted(s["row"], e, side="right") - 1] = 1
        nb_front = plan.order_units(flag) if front_first else 0
        assert not front_first or 0 < nb_front < plan.c.nbUnits
        ctx = dc.Context(0)
        ctx.set_mesh(s["conn"], s["coord"], s["row"], s["col"], 0)
        ctx.upload_plan(plan)
        slabs.append(s); ctxs.append(ctx); fronts.append((nb_front, int(plan.c.nbUnits)))
        vals.append(torch.full((s["nnz"] * 9,), float("nan"), dtype=torch.float64, device="cuda:0"))
        xs[r] = dc.DeviceExchange(s["intf"], 9, 0)
    dc.DeviceExchange.connect_local(xs)
    for step in range(3):                                   # both parities of the buffers, and their reuse
        for r in range(world):
            nf, nu = fronts[r]
            if front_first:
                ctxs[r].assemble_units(dc.OP_ELASTICITY, [LAMBDA, MU], vals[r], 0, nf, stream=stream)
            else:
                ctxs[r].assemble(dc.OP_ELASTICITY, [LAMBDA, MU], dc.MODE_DETERMINISTIC, vals[r], stream)
        for r in range(world):
            xs[r].pack(vals[r], stream)
        for r in range(world):
            nf, nu = fronts[r]
            if front_first:
                ctxs[r].assemble_units(dc.OP_ELASTICITY, [LAMBDA, MU], vals[r], nf, nu - nf, reserve=4, skip_big=True, stream=stream)
        if step == 0:
            torch.cuda.synchronize()
        for r in range(world):
            xs[r].unpack(vals[r], stream)
        torch.cuda.synchronize()
        row, col, want = _global_reference(n, world, dim)
        N_glob = row.shape[0] - 1
        key_ref = np.repeat(np.arange(N_glob, dtype=np.int64), np.diff(row)) * N_glob + col
        _check_against_global(slabs, vals, key_ref, want, N_glob, 9)
    assert xs[0].bytes_per_step == 2 * 8 * 9 * slabs[0]["intf"][1].shape[0]


def test_exchange_of_a_metis_split_with_three_way_edges():
    """BASELINE config 4 at test size through the C exchange: four parts of the graded / relabelled /
    shuffled mesh (edges shared by three ranks included: every pairwise buffer is packed before any is
    added); neighbours are added in ascending rank order, so a rerun gives the same bits."""
    import torch
    from dc_lib_b200 import multigpu
    from test_multirank_gloo import _irregular_mesh
    n, world, dim = 8, 4, 3
    conn_g, coord_g = _irregular_mesh(n)
    N_glob = coord_g.shape[0]
    part = dc.DC_partition_mesh(conn_g, N_glob, world)
    stream = torch.cuda.current_stream().cuda_stream
    parts, vals, ctxs, xs = [], [], [], {}
    for r in range(world):
        p = multigpu.build_part(conn_g, coord_g, part, r, world, builder="metis")
        plan = dc.HostPlan(p["conn"], p["N"], p["row"], p["col"], 0, p["rec"], 320)
        ctx = dc.Context(0)
        ctx.set_mesh(p["conn"], p["coord"], p["row"], p["col"], 0)
        ctx.upload_plan(plan)
        parts.append(p); ctxs.append(ctx)
        vals.append(torch.empty(p["nnz"] * 9, dtype=torch.float64, device="cuda:0"))
        xs[r] = dc.DeviceExchange(p["intf"], 9, 0)
    dc.DeviceExchange.connect_local(xs)
    runs = []
    for _ in range(2):
        for r in range(world):
            ctxs[r].assemble(dc.OP_ELASTICITY, [LAMBDA, MU], dc.MODE_DETERMINISTIC, vals[r], stream)
        for r in range(world):
            xs[r].pack(vals[r], stream)
        for r in range(world):
            xs[r].unpack(vals[r], stream)
        torch.cuda.synchronize()
        runs.append([v.cpu().numpy().copy() for v in vals])
    for a, b in zip(*runs):
        assert a.tobytes() == b.tobytes()
    row, col = dc.build_csr(conn_g, N_glob)
    mesh = dict(conn=conn_g, coord=coord_g, row=row, col=col, E=conn_g.shape[0], N=N_glob, nnz=int(row[-1]))
    want = oracle_serial(mesh, dim).reshape(-1, 9)
    key_ref = np.repeat(np.arange(N_glob, dtype=np.int64), np.diff(row)) * N_glob + col
    _check_against_global(parts, vals, key_ref, want, N_glob, 9)


def test_interface_edges_from_the_reference_node_lists():
    """dc_interface_edges: the reference describes an interface by node lists (intfNodes, DC.h:141-143);
    the edges among the nodes of a slab's plane, ordered by list positions, are the edges plane_edges
    finds, and both sides of the plane get the same keys."""
    from dc_lib_b200 import multigpu
    n, world = 5, 2
    s0 = multigpu.build_slab(n, 0, world, builder="geometric")
    s1 = multigpu.build_slab(n, 1, world, builder="geometric")
    m = n + 1
    plane = np.arange(m * m)
    nodes0 = s0["node_perm"][n * m * m + plane] + 1           # plane x = n of slab 0, 1-based local ids
    nodes1 = s1["node_perm"][0 * m * m + plane] + 1           # plane x = 0 of slab 1, same order
    e0, k0 = dc.interface_edges(s0["row"], s0["col"], nodes0)
    e1, k1 = dc.interface_edges(s1["row"], s1["col"], nodes1)
    assert np.array_equal(k0, k1) and e0.shape[0] > 0
    assert np.array_equal(np.sort(e0), np.sort(s0["intf"][1])) and np.array_equal(np.sort(e1), np.sort(s1["intf"][0]))
    assert np.array_equal(e0, s0["intf"][1])                  # the very order of multigpu.plane_edges
