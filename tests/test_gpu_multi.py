"""One process per GPU (torchrun), the overlapped step of the bench: front units, pack over NVLink ||
interior units, unpack-add (dc_cuda_assemble_exchange).  Needs at least two GPUs; run with
    gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu -q
The worker is tests/multi_worker.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("case", ["slabs", "metis"])
def test_overlapped_step_on_two_gpus(case, tmp_path):
    if _gpus() < 2:
        pytest.skip("needs two GPUs (kernels that wait for another rank must not share a GPU)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "multi_worker.py"), case]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("RANK_OK") == 2, out.stdout[-3000:]


def test_cpp_program_one_process_per_domain(tmp_path):
    """tests/dropin/multi_driver.cc: a C++ user program (DC.h + DC_device.h, no Python, no torch) run as two
    processes, one GPU each: interface in the reference's node-list layout, handles swapped through files,
    DC_assembly_device with the overlapped exchange; every local entry equals the single-domain oracle."""
    if _gpus() < 2:
        pytest.skip("needs two GPUs (kernels that wait for another rank must not share a GPU)")
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dc_lib_b200 as dc
    from test_multirank_gloo import _global_reference
    exe = os.path.join(str(tmp_path), "multi_driver")
    libdir, libfile = os.path.split(dc.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "dropin", "multi_driver.cc"), "-o", exe, "-L" + libdir,
                           "-l:" + libfile, "-Wl,-rpath," + libdir, "-pthread"])
    n, world = 10, 2
    procs = [subprocess.Popen([exe, str(r), str(world), str(n), str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True) for r in range(world)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0 and "RANK_OK" in o, o + e
    row, col, want = _global_reference(n, world, 3)
    N_glob = row.shape[0] - 1
    key_ref = np.repeat(np.arange(N_glob, dtype=np.int64), np.diff(row)) * N_glob + col
    for r in range(world):
        keys = np.fromfile(os.path.join(str(tmp_path), f"keys_{r}.bin"), dtype=np.int64)
        vals = np.fromfile(os.path.join(str(tmp_path), f"values_{r}.bin"), dtype=np.float64).reshape(-1, 9)
        pos = np.searchsorted(key_ref, keys)
        assert np.array_equal(key_ref[pos], keys)
        assert np.abs(vals - want.reshape(-1, 9)[pos]).max() <= 1e-12 * np.abs(want).max()
