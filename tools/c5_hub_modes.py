"""tools/c5_hub_modes.py [n=96] -- BASELINE config 5 on one GPU: the hub-collapsed cube (valence ~100 hubs,
skewed node degree; dc-lib_b200/meshgen.py hub_collapse), tree by DC_create_tree (METIS, as the reference
builds it), elasticity: deterministic mode against the oracle's serial loop (bitwise) and against the
order-free red.global mode (<= 1e-12), with the kernel times of both.  One JSON line."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import common
import dc_lib_b200 as dc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
dc.DC_set_num_threads(0)
t0 = time.time()
with tempfile.TemporaryDirectory() as tmp:
    mesh = common.prepare_mesh(common.MineLib(0, device=True), n, tmp, hub=True)
t_mesh = time.time() - t0
deg = np.diff(mesh["row"])
t0 = time.time()
plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, common.tree_records(mesh), 512)
t_plan = time.time() - t0
ctx = dc.Context(0)
ctx.set_mesh(mesh["conn"], mesh["coord"], mesh["row"], mesh["col"], 0)
ctx.upload_plan(plan)
prm = [common.LAMBDA, common.MU]
vals = {m: torch.empty(mesh["nnz"] * 9, dtype=torch.float64, device="cuda") for m in ("det", "atomic")}
ms = {}
for name, mode in (("det", dc.MODE_DETERMINISTIC), ("atomic", dc.MODE_ATOMIC)):
    for _ in range(3):
        ctx.assemble(dc.OP_ELASTICITY, prm, mode, vals[name])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ctx.assemble(dc.OP_ELASTICITY, prm, mode, vals[name])
    e1.record()
    torch.cuda.synchronize()
    ms[name] = e0.elapsed_time(e1) / 20
det, atom = vals["det"].cpu().numpy(), vals["atomic"].cpu().numpy()
want = common.oracle_serial(mesh, 3)
st = plan.stats()
print(json.dumps({
    "workload": f"c5: hub-collapsed jittered Kuhn cube {n}^3 ({mesh['E']} tets, {mesh['N']} nodes), METIS tree (element passes on the GPU), elasticity",
    "max_row_length": int(deg.max()), "mean_row_length": float(deg.mean()), "rows_on_the_fallback_lists_edges": int(st.get("big_edges", 0)),
    "deterministic_ms": ms["det"], "deterministic_G_elem_s": mesh["E"] / ms["det"] / 1e6,
    "atomic_ms": ms["atomic"], "atomic_G_elem_s": mesh["E"] / ms["atomic"] / 1e6,
    "deterministic_bitwise_equal_to_oracle_serial": bool(det.tobytes() == want.tobytes()),
    "atomic_vs_deterministic_max_rel": float(np.abs(atom - det).max() / np.abs(det).max()),
    "setups_per_element": st["slots_total"] / mesh["E"], "setup_s": {"mesh_tree_csr": t_mesh, "plan": t_plan}}))
