"""Per-phase cycle counts of gather_units (clock64 stamps of thread 0, debug hook)."""
import sys, os, tempfile, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dc_lib_b200 as dc
import common
n = int(sys.argv[1]); slots = int(sys.argv[2]); threads = int(sys.argv[3]); geo = len(sys.argv) > 4 and sys.argv[4] == 'geo'
with tempfile.TemporaryDirectory() as tmp:
    mesh = common.prepare_mesh(common.MineLib(0, geometric=geo), n, tmp)
plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, common.tree_records(mesh), slots)
ctx = dc.Context(0); ctx.set_mesh(mesh["conn"], mesh["coord"], mesh["row"], mesh["col"], 0); ctx.upload_plan(plan)
ctx.set_option("unit_threads", threads)
ctx.set_option("persistent", int(sys.argv[5]) if len(sys.argv) > 5 else 1)
vals = torch.empty(mesh["nnz"] * 9, dtype=torch.float64, device="cuda")
prm = [common.LAMBDA, common.MU]
for _ in range(3): ctx.assemble(dc.OP_ELASTICITY, prm, dc.MODE_DETERMINISTIC, vals)
L = dc.lib(); L.dc_cuda_debug_phases.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
L.dc_cuda_debug_phases(ctx.h, None)
ctx.assemble(dc.OP_ELASTICITY, prm, dc.MODE_DETERMINISTIC, vals)
nu = plan.c.nbUnits
out = np.zeros((nu, 8), dtype=np.int64)
L.dc_cuda_debug_phases(ctx.h, out.ctypes.data)
persistent = int(sys.argv[5]) if len(sys.argv) > 5 else 1
out = out[out[:, 1] > 0]          # stamp 0 only exists for the first unit of each CTA
d = np.diff(out[:, 1:7], axis=1)
names = ["records (P1b) + sync A", "wait schedule", "accumulate+store (warp 0)", "halo gather of next unit", "sync B (other warps)"]
tot = (out[:, 6] - out[:, 1])
print(f"units {nu}  elements {mesh['E']}  own elems/unit {mesh['E']/nu:.0f}  mean cycles/unit {tot.mean():.0f}")
for i, nm in enumerate(names):
    print(f"  {nm:24s} mean {d[:, i].mean():8.0f}  median {np.median(d[:, i]):8.0f}  share {d[:, i].sum()/tot.sum()*100:5.1f}%")
