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
