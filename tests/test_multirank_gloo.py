"""world_size-2 gloo test (CPU) of the multi-GPU host logic: slab decomposition, agreement of
the interface edge lists on both sides of the shared plane, and the exchange itself.  The
local partial values come from the CPU oracle here (this file is a test; on GPUs they come
from dc_cuda_assemble) and the summed result is compared with the oracle run on the merged
global mesh."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _global_reference(n, world, dim):
    """Oracle on the merged bar of `world` cubes, keyed by global (row, col) node ids."""
    import common
    import dc_lib_b200 as dc
    m = n + 1
    conns, N_glob = [], (world * n + 1) * m * m
    coord = np.zeros((N_glob, 3))
    for r in range(world):
        c, x, _ = dc.kuhn_cube_fast(n, x_offset=r * n)
        gid = np.arange(m ** 3, dtype=np.int64) + r * n * m * m
        conns.append(gid[c - 1] + 1)
        coord[gid] = x
    conn = np.ascontiguousarray(np.concatenate(conns), dtype=np.int32)
    row, col = dc.build_csr(conn, N_glob)
    mesh = dict(conn=conn, coord=np.ascontiguousarray(coord), row=row, col=col, E=conn.shape[0], N=N_glob, nnz=int(row[-1]))
    vals = common.oracle_serial(mesh, dim)
    return row, col, vals.reshape(mesh["nnz"], dim * dim)


def _worker(rank, world, n, dim, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import common
    import dc_lib_b200 as dc
    from dc_lib_b200 import multigpu
    slab = multigpu.build_slab(n, rank, world, builder="geometric")
    part = torch.from_numpy(common.oracle_serial(slab, dim))            # local partial sums
    before = part.clone()
    ex = multigpu.InterfaceExchange(slab["intf"], dim * dim, "cpu", dist)
    ex.run(part)
    # both sides of a plane must list the same (global row, global col) pairs in the same order
    rows = np.repeat(np.arange(slab["N"]), np.diff(slab["row"]))
    for nb, edges in slab["intf"].items():
        pairs = np.stack([slab["gid_of_new"][rows[edges]], slab["gid_of_new"][slab["col"][edges]]], 1)
        theirs = [None, None]
        dist.all_gather_object(theirs, (rank, nb, pairs))
        for (r2, nb2, p2) in theirs:
            if r2 == nb and nb2 == rank:
                assert np.array_equal(pairs, p2)
    changed = int((part != before).sum())
    grow = slab["gid_of_new"][rows]
    gcol = slab["gid_of_new"][slab["col"]]
    torch.save(dict(grow=grow, gcol=gcol, vals=part.numpy().reshape(slab["nnz"], dim * dim), changed=changed),
               os.path.join(out, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,dim", [(6, 1), (5, 3)])
def test_two_rank_interface_exchange(tmp_path, n, dim):
    world, port = 2, 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, n, dim, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    row, col, want = _global_reference(n, world, dim)
    N_glob = row.shape[0] - 1
    key_ref = np.repeat(np.arange(N_glob, dtype=np.int64), np.diff(row)) * N_glob + col
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"), weights_only=False)
        assert d["changed"] > 0                                   # the exchange did something
        pos = np.searchsorted(key_ref, d["grow"] * N_glob + d["gcol"])
        assert np.array_equal(key_ref[pos], d["grow"] * N_glob + d["gcol"])
        # rows of nodes on a shared plane are complete only for edges both ranks hold; every local
        # edge must now carry the global value
        scale = np.abs(want).max()
        assert np.abs(d["vals"] - want[pos]).max() <= 1e-12 * scale


# ---- general element partition (BASELINE config 4: irregular mesh, METIS-split) --------------

def _irregular_mesh(n):
    """Graded, relabelled, shuffled cube (config C4's shape at test size), original numbering."""
    import common
    import dc_lib_b200 as dc
    conn, _ = dc.kuhn_cube(n)
    coord = dc.jitter_coords(n)
    return common.scramble_mesh(conn, coord)


def _part_worker(rank, world, n, dim, method, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import common
    import dc_lib_b200 as dc
    from dc_lib_b200 import multigpu
    conn_g, coord_g = _irregular_mesh(n)
    part = dc.DC_partition_mesh(conn_g, coord_g.shape[0], world, coord_g if method == "geometric" else None)
    p = multigpu.build_part(conn_g, coord_g, part, rank, world, builder="metis")
    vals = torch.from_numpy(common.oracle_serial(p, dim))               # local partial sums
    before = vals.clone()
    multigpu.InterfaceExchange(p["intf"], dim * dim, "cpu", dist).run(vals)
    rows = np.repeat(np.arange(p["N"]), np.diff(p["row"]))
    torch.save(dict(grow=p["gid_of_new"][rows], gcol=p["gid_of_new"][p["col"]], vals=vals.numpy().reshape(p["nnz"], dim * dim),
                    changed=int((vals != before).sum()), nb=sorted(p["intf"]), E=p["E"],
                    lens={q: int(e.shape[0]) for q, e in p["intf"].items()}),
               os.path.join(out, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,dim,method", [(2, 6, 3, "metis"), (3, 6, 1, "metis"), (2, 5, 3, "geometric")])
def test_general_partition_interface_exchange(tmp_path, world, n, dim, method):
    """An irregular mesh cut by DC_partition_mesh into `world` parts: every rank assembles its own
    elements, the interface exchange sums the partial values of the edges several ranks hold, and
    every local CSR entry then carries the value of the single-domain assembly (<= 1e-12)."""
    port = 29500 + ((os.getpid() + 7 * world + n) % 500)
    mp.spawn(_part_worker, args=(world, n, dim, method, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    import dc_lib_b200 as dc
    conn_g, coord_g = _irregular_mesh(n)
    N_glob = coord_g.shape[0]
    row, col = dc.build_csr(conn_g, N_glob)
    mesh = dict(conn=conn_g, coord=coord_g, row=row, col=col, E=conn_g.shape[0], N=N_glob, nnz=int(row[-1]))
    want = common.oracle_serial(mesh, dim).reshape(mesh["nnz"], dim * dim)
    key_ref = np.repeat(np.arange(N_glob, dtype=np.int64), np.diff(row)) * N_glob + col
    scale = np.abs(want).max()
    total, lens = 0, {}
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"), weights_only=False)
        assert d["changed"] > 0 and d["nb"]
        total += d["E"]
        lens[r] = d["lens"]
        key = d["grow"] * N_glob + d["gcol"]
        pos = np.searchsorted(key_ref, key)
        assert np.array_equal(key_ref[pos], key)
        assert np.abs(d["vals"] - want[pos]).max() <= 1e-12 * scale
    assert total == conn_g.shape[0]                                   # every element assembled exactly once
    for r in range(world):
        for q, ln in lens[r].items():
            assert lens[q][r] == ln                                   # both sides list the same edges
