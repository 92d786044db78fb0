"""Small deterministic + atomic + list assembly, the tree passes, the CSR build and the permutation
kernels, for compute-sanitizer runs (compute-sanitizer --tool memcheck python tools/sanitize_case.py)."""
import sys, os, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dc_lib_b200 as dc, common
with tempfile.TemporaryDirectory() as tmp:
    mesh = common.prepare_mesh(common.MineLib(4), 9, tmp)
for sym in (True, False):
    plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, common.tree_records(mesh), 160, symmetric=sym)
    ctx = dc.Context(0); ctx.set_mesh(mesh["conn"], mesh["coord"], mesh["row"], mesh["col"], 0); ctx.upload_plan(plan)
    ctx.upload_lists(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0)
    for dim, op, prm in ((1, dc.OP_LAPLACIAN, None), (3, dc.OP_ELASTICITY, [common.LAMBDA, common.MU])):
        want = common.oracle_serial(mesh, dim)
        for mode in (dc.MODE_DETERMINISTIC, dc.MODE_LIST, dc.MODE_ATOMIC):
            v = torch.full((mesh["nnz"] * dim * dim,), float("nan"), dtype=torch.float64, device="cuda")
            ctx.assemble(op, prm, mode, v); torch.cuda.synchronize()
            got = v.cpu().numpy()
            ok = got.tobytes() == want.tobytes() if mode != dc.MODE_ATOMIC else np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
            print("sym", sym, "dim", dim, "mode", mode, "ok", bool(ok), flush=True)
            assert ok
with tempfile.TemporaryDirectory() as tmp:
    for vec, geometric in ((0, False), (4, True)):
        a = common.prepare_mesh(common.MineLib(vec, geometric), 9, tmp, tag="host")
        b = common.prepare_mesh(common.MineLib(vec, geometric, device=True), 9, tmp, tag="dev")
        assert a["tree"] == b["tree"] and a["conn"].tobytes() == b["conn"].tobytes()
        print("tree on the device", vec, geometric, "ok", flush=True)
row, col = dc.build_csr_device(torch.from_numpy(mesh["conn"]).cuda(), mesh["N"], 0)
assert np.array_equal(row.cpu().numpy(), mesh["row"]) and np.array_equal(col.cpu().numpy(), mesh["col"])
print("done")
