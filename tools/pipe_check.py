"""Scratch: unit-kernel variants against each other: bitwise equality of the values and timing.
usage: pipe_check.py [n] [slots]   (plans: default, DC_PLAN_DIAG_SPLIT=3)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dc_lib_b200 as dc
from dc_lib_b200 import multigpu

n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 512
LAMBDA, MU = 0.5769230769230769, 0.3846153846153846
ref = {}
for split in ("", "3"):
    if split:
        os.environ["DC_PLAN_DIAG_SPLIT"] = split
    else:
        os.environ.pop("DC_PLAN_DIAG_SPLIT", None)
    ctx, info = multigpu.setup_device_slabs(n, 0, 1, "cuda:0", builder="geometric", vec=0, slots=slots, dist=None)
    E, nnz, st = info["E"], info["nnz"], info["stats"]
    print(f"split={split or 'no'} n={n} E={E} units={st['units']} setups/elem={st['slots_total']/E:.3f} item_eff={st['items_useful']/max(st['items'],1):.3f}", flush=True)
    stream = torch.cuda.current_stream().cuda_stream
    for dim, op, prm in ((3, dc.OP_ELASTICITY, [LAMBDA, MU]), (1, dc.OP_LAPLACIAN, None)):
        for pers, thr in ((1, 256), (2, 256)):
            ctx.set_option("unit_threads", thr)
            ctx.set_option("persistent", pers)
            vals = torch.full((nnz * dim * dim,), float("nan"), dtype=torch.float64, device="cuda:0")
            for _ in range(3):
                ctx.assemble(op, prm, dc.MODE_DETERMINISTIC, vals, stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ctx.assemble(op, prm, dc.MODE_DETERMINISTIC, vals, stream)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            if dim not in ref:
                ref[dim] = vals.clone()
            same = torch.equal(vals, ref[dim])
            print(f"  dim={dim} persistent={pers} threads={thr}: {ms:7.3f} ms {E/ms/1e6:7.2f} Gelem/s  bitwise equal to the first: {same}  nan: {int(torch.isnan(vals).sum())}", flush=True)
            del vals
    ctx.close()
