"""Knock-out timing of the unit kernel (scratch helper, experiment build only):
  make -C dc-lib_b200 variant NAME=ko DEFS=-DDC_KO
  DC_B200_LIB=dc-lib_b200/libdc_b200_ko.so python tools/ko_sweep.py [n] [slots]
Each DC_KO bit switches one phase off (results are wrong); the time difference to the full kernel is
what that phase costs.  Also times the plan built with DC_PLAN_UNIT_NODES (fixed row windows)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dc_lib_b200 as dc
from dc_lib_b200 import multigpu

import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 512
kos = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 3, 4, 8, 16, 32, 36, 19, 27, 63]
windows = sys.argv[4].split(",") if len(sys.argv) > 4 else ["", "32", "64"]
LAMBDA, MU = 0.5769230769230769, 0.3846153846153846
DIM = int(os.environ.get("DC_SWEEP_DIM", "3"))
THREADS = int(os.environ.get("DC_SWEEP_THREADS", "256"))
SYM = int(os.environ.get("DC_SWEEP_SYM", "1"))
OP, PRM = (dc.OP_ELASTICITY, [LAMBDA, MU]) if DIM == 3 else (dc.OP_LAPLACIAN, None)
names = {1: "transposed stores", 2: "direct stores", 4: "diag tasks", 8: "record build", 16: "accumulation loops", 32: "upper tasks"}
dc.DC_set_max_elem_per_part(bench.leaf_capacity(n, int(os.environ.get("DC_SWEEP_LEAF_NODES", "62"))))   # the bench geometry
for w in windows:
    if w:
        os.environ["DC_PLAN_UNIT_NODES"] = w
    else:
        os.environ.pop("DC_PLAN_UNIT_NODES", None)
    t0 = time.time()
    ctx, info = multigpu.setup_device_slabs(n, 0, 1, "cuda:0", builder="geometric", vec=0, slots=slots, dist=None, symmetric=bool(SYM))
    ctx.set_option("unit_threads", THREADS)
    ctx.set_option("persistent", int(os.environ.get("DC_SWEEP_PERSISTENT", "1")))
    E, nnz, st = info["E"], info["nnz"], info["stats"]
    print(f"window={w or 'subtree'} n={n} E={E} units={st['units']} setups/elem={st['slots_total']/E:.3f} "
          f"item_eff={st['items_useful']/max(st['items'],1):.3f} setup {time.time()-t0:.1f}s", flush=True)
    vals = torch.empty(nnz * DIM * DIM, dtype=torch.float64, device="cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    base = None
    for ko in (kos if not w else [0]):
        os.environ["DC_KO"] = str(ko)
        for _ in range(3):
            ctx.assemble(OP, PRM, dc.MODE_DETERMINISTIC, vals, stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ctx.assemble(OP, PRM, dc.MODE_DETERMINISTIC, vals, stream)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if ko == 0:
            base = ms
        off = " + ".join(v for k, v in names.items() if ko & k) or "nothing"
        print(f"  KO={ko:3d} ({off} off): {ms:7.3f} ms  {E/ms/1e6:7.2f} Gelem/s  saves {100*(base-ms)/base:5.1f}%", flush=True)
    ctx.close()
    del vals
