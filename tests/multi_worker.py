"""Worker of tests/test_gpu_multi.py: one rank of a two-GPU run (torchrun).  `slabs`: two jittered cubes
stacked along x; `metis`: the graded / relabelled / shuffled mesh of BASELINE config 4 cut by
DC_partition_mesh.  Every rank assembles with the overlapped step and compares its local values with the
single-domain oracle (tolerance 1e-12: interface entries are summed in another order)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dc_lib_b200 as dc  # noqa: E402
from common import LAMBDA, MU, oracle_serial  # noqa: E402
from dc_lib_b200 import multigpu  # noqa: E402


def main():
    case = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                    # only the 128-byte handles travel through it
    dc.DC_set_num_threads(0)
    if case == "slabs":
        from test_multirank_gloo import _global_reference
        n = 10
        p = multigpu.build_slab(n, rank, world, builder="geometric")
        row, col, want = _global_reference(n, world, 3)
        N_glob = row.shape[0] - 1
    else:
        from test_multirank_gloo import _irregular_mesh
        conn_g, coord_g = _irregular_mesh(12)
        N_glob = coord_g.shape[0]
        part = dc.DC_partition_mesh(conn_g, N_glob, world)
        p = multigpu.build_part(conn_g, coord_g, part, rank, world, builder="metis")
        row, col = dc.build_csr(conn_g, N_glob)
        mesh = dict(conn=conn_g, coord=coord_g, row=row, col=col, E=conn_g.shape[0], N=N_glob, nnz=int(row[-1]))
        want = oracle_serial(mesh, 3).reshape(-1, 9)
    plan = dc.HostPlan(p["conn"], p["N"], p["row"], p["col"], 0, p["rec"], 512)
    flag = np.zeros(p["N"], dtype=np.uint8)
    for e in p["intf"].values():
        flag[np.searchsorted(p["row"], e, side="right") - 1] = 1
    nb_front = plan.order_units(flag)
    ctx = dc.Context(local)
    ctx.set_mesh(p["conn"], p["coord"], p["row"], p["col"], 0)
    ctx.upload_plan(plan)
    ctx.set_front_units(nb_front)
    x = dc.DeviceExchange(p["intf"], 9, local)
    x.connect(dist, rank, world)
    vals = torch.empty(p["nnz"] * 9, dtype=torch.float64, device=f"cuda:{local}")
    stream = torch.cuda.current_stream().cuda_stream
    first = None
    for step in range(4):
        vals.fill_(float("nan"))
        ctx.assemble_exchange(x, dc.OP_ELASTICITY, [LAMBDA, MU], vals, stream)
        torch.cuda.synchronize()
        got = vals.cpu().numpy()
        if first is None:
            first = got.copy()
        assert got.tobytes() == first.tobytes(), "steps differ"
    key_ref = np.repeat(np.arange(N_glob, dtype=np.int64), np.diff(row)) * N_glob + col
    rows = np.repeat(np.arange(p["N"]), np.diff(p["row"]))
    key = p["gid_of_new"][rows] * N_glob + p["gid_of_new"][p["col"]]
    pos = np.searchsorted(key_ref, key)
    assert np.array_equal(key_ref[pos], key)
    err = np.abs(first.reshape(-1, 9) - want.reshape(-1, 9)[pos]).max() / np.abs(want).max()
    assert err <= 1e-12, err
    dist.barrier()
    print(f"RANK_OK {rank} case {case} elements {p['E']} front units {nb_front}/{plan.c.nbUnits} interface bytes {x.bytes_per_step} err {err:.2e}", flush=True)
    x.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
