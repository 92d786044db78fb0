"""Scratch timing of the assembly kernels on one GPU (not the bench contract)."""
import sys, os, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dc_lib_b200 as dc
import common

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
slots_list = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [832]
threads_list = [int(s) for s in sys.argv[3].split(",")] if len(sys.argv) > 3 else [128, 256]
sym_list = [int(s) for s in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 0]
vec = 0
t0 = time.time()
with tempfile.TemporaryDirectory() as tmp:
    mesh = common.prepare_mesh(common.MineLib(vec), n, tmp)
print(f"mesh n={n} E={mesh['E']} N={mesh['N']} nnz={mesh['nnz']} prep {time.time()-t0:.1f}s nproc={os.cpu_count()}", flush=True)
rec = common.tree_records(mesh)
E, N, nnz = mesh["E"], mesh["N"], mesh["nnz"]
first = True
for slots, sym in [(a, b) for a in slots_list for b in sym_list]:
    t0 = time.time()
    plan = dc.HostPlan(mesh["conn"], N, mesh["row"], mesh["col"], 0, rec, slots, symmetric=bool(sym))
    st = plan.stats()
    print(f"slots={slots} sym={sym} plan {time.time()-t0:.2f}s units={st['units']} pad_eff={st['items_useful']/max(st['items'],1):.3f} slots/elem={st['slots_total']/E:.3f} maxItems={plan.c.maxItems} maxTasks={plan.c.maxTasks} maxHalo={plan.c.maxHalo}", flush=True)
    for threads in threads_list:
        ctx = dc.Context(0)
        ctx.set_mesh(mesh["conn"], mesh["coord"], mesh["row"], mesh["col"], 0)
        ctx.upload_plan(plan)
        ctx.set_option("unit_threads", threads)
        for dim, op, prm in ((1, dc.OP_LAPLACIAN, None), (3, dc.OP_ELASTICITY, [common.LAMBDA, common.MU])):
            vals = torch.empty(nnz * dim * dim, dtype=torch.float64, device="cuda")
            balg = 16 * E + 24 * N + 4 * (N + 1) + 4 * nnz + 8 * dim * dim * nnz
            modes = [(dc.MODE_DETERMINISTIC, "det")] + ([(dc.MODE_ATOMIC, "atomic")] if first else [])
            for mode, name in modes:
                for _ in range(3):
                    ctx.assemble(op, prm, mode, vals)
                torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                ev0.record()
                for _ in range(reps):
                    ctx.assemble(op, prm, mode, vals)
                ev1.record(); torch.cuda.synchronize()
                ms = ev0.elapsed_time(ev1) / reps
                print(f"  thr={threads} d={dim} {name:7s} {ms:8.3f} ms  {E/ms/1e6:8.2f} Gelem/s  B_alg {balg/ms/1e6:7.1f} GB/s  frac {balg/ms/1e6/6550.4:.3f}", flush=True)
        first = False
        ctx.close()
