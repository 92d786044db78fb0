"""Scratch: the north-star "fast mode" (reset_values + scatter_atomic, red.global.add.f64) next to the
deterministic unit kernel on the same mesh, for an ncu capture and a CUDA-event timing.
usage: atomic_profile.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dc_lib_b200 as dc
from dc_lib_b200 import multigpu

n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
LAMBDA, MU = 0.5769230769230769, 0.3846153846153846
ctx, info = multigpu.setup_device_slabs(n, 0, 1, "cuda:0", builder="geometric", vec=0, slots=512, dist=None)
E, nnz = info["E"], info["nnz"]
stream = torch.cuda.current_stream().cuda_stream
for dim, op, prm in ((3, dc.OP_ELASTICITY, [LAMBDA, MU]), (1, dc.OP_LAPLACIAN, None)):
    vals = torch.empty(nnz * dim * dim, dtype=torch.float64, device="cuda:0")
    balg = 16 * E + 24 * info["N"] + 4 * (info["N"] + 1) + 4 * nnz + 8 * dim * dim * nnz
    for mode, name in ((dc.MODE_DETERMINISTIC, "deterministic"), (dc.MODE_ATOMIC, "atomic")):
        for _ in range(2):
            ctx.assemble(op, prm, mode, vals, stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ctx.assemble(op, prm, mode, vals, stream)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"dim={dim} {name:13s} {ms:8.3f} ms  {E/ms/1e6:7.2f} Gelem/s  B_alg {balg/ms/1e6:7.1f} GB/s", flush=True)
    del vals
