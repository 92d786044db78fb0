"""Per-phase cycle counts of gather_units (clock64 stamps, debug hook dc_cuda_debug_phases): thread 0's
phases and the end of the tasks of every warp, on the bench geometry.
usage: python tools/phase_times.py <cube> <slots> <threads> [leaf_nodes=62] [op=elasticity|laplacian]"""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import dc_lib_b200 as dc
import bench
n = int(sys.argv[1]); slots = int(sys.argv[2]); threads = int(sys.argv[3])
leaf_nodes = int(sys.argv[4]) if len(sys.argv) > 4 else 62
lap = len(sys.argv) > 5 and sys.argv[5].startswith("lap")
dc.DC_set_max_elem_per_part(bench.leaf_capacity(n, leaf_nodes) if leaf_nodes > 0 else 200)
w = bench.build_workload(dc, n, 0, "geometric")
plan = dc.HostPlan(w["conn"], w["N"], w["row"], w["col"], 0, w["rec"], slots, bank_model=1 if lap else 0)
ctx = dc.Context(0); ctx.set_mesh(w["conn"], w["coord"], w["row"], w["col"], 0); ctx.upload_plan(plan)
ctx.set_option("unit_threads", threads)
vals = torch.empty(w["nnz"] * (1 if lap else 9), dtype=torch.float64, device="cuda")
prm = None if lap else [bench.LAMBDA, bench.MU]
OP = dc.OP_LAPLACIAN if lap else dc.OP_ELASTICITY
for _ in range(3): ctx.assemble(OP, prm, dc.MODE_DETERMINISTIC, vals)
L = dc.lib(); L.dc_cuda_debug_phases.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
L.dc_cuda_debug_phases(ctx.h, None)
ctx.assemble(OP, prm, dc.MODE_DETERMINISTIC, vals)
nu = plan.c.nbUnits
out = np.zeros((nu, 24), dtype=np.int64)
L.dc_cuda_debug_phases(ctx.h, out.ctypes.data)
out = out[out[:, 1] > 0]
d = np.diff(out[:, 1:7], axis=1)
names = ["records (P1b) + sync A", "issue copies of next unit", "tasks (warp 0)", "halo gather of next unit", "sync B (wait for others)"]
tot = (out[:, 6] - out[:, 1])
st = plan.stats()
print(f"units {nu}  elements {w['E']}  elems/unit {w['E']/nu:.0f}  tasks/unit {st['tasks']/nu:.2f}  mean cycles/unit {tot.mean():.0f}")
for i, nm in enumerate(names):
    print(f"  {nm:28s} mean {d[:, i].mean():8.0f}  median {np.median(d[:, i]):8.0f}  share {d[:, i].sum()/tot.sum()*100:5.1f}%")
nw = threads // 32
ends = out[:, 8:8 + nw] - out[:, 3:4]          # end of every warp's tasks, relative to the start of the task phase
print("end of tasks per warp (cycles after the start of the task phase), mean:")
print("  " + " ".join(f"w{i}:{ends[:, i].mean():.0f}" for i in range(nw)))
srt = np.sort(ends, axis=1)
print(f"  first warp done {srt[:, 0].mean():.0f}  median warp {srt[:, nw // 2].mean():.0f}  last warp {srt[:, -1].mean():.0f}  "
      f"last is warp 0 in {np.mean(np.argmax(ends, axis=1) == 0) * 100:.0f}% of units; mean idle per warp before sync B "
      f"{(srt[:, -1:] - ends).mean():.0f}")
