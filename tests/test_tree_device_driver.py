"""Host side of DC_create_tree_device (the breadth-first bookkeeping around the dc_cuda_tree_* passes,
dc-lib_b200/host/dc_tree.cc) against the recursive host builder, on a machine WITHOUT a GPU: a child
process preloads tests/mock/tree_split_mock.cc -- plain loops written from the contract in
include/dc_cuda.h, test infrastructure only -- in place of the CUDA entry points.  Tree file,
connectivity and coordinates must be byte-identical.  (The kernels: tests/test_gpu_tree_device.py.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import common
from common import MineLib, prepare_mesh
import dc_lib_b200 as dc
cases = [(7, 0, False, False, False), (12, 0, False, False, False), (12, 4, False, False, False),
         (16, 0, True, False, False), (14, 0, False, True, False), (11, 0, False, False, True),
         (11, 2, True, False, True), (20, 0, True, False, False)]
for n, vec, geometric, hub, scramble in cases:
    for cap in (200, 61):
        dc.DC_set_max_elem_per_part(cap)
        a = prepare_mesh(MineLib(vec, geometric), n, {tmp!r}, hub=hub, tag="host", scramble=scramble)
        b = prepare_mesh(MineLib(vec, geometric, device=True), n, {tmp!r}, hub=hub, tag="dev", scramble=scramble)
        assert a["tree"] == b["tree"], (n, vec, geometric, hub, scramble, cap)
        assert a["conn"].tobytes() == b["conn"].tobytes()
        assert a["coord"].tobytes() == b["coord"].tobytes()
print("OK", len(cases))
"""


def test_breadth_first_driver_equals_recursive_builder(tmp_path):
    mock = os.path.join(str(tmp_path), "libtree_split_mock.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", mock, os.path.join(ROOT, "tests", "mock", "tree_split_mock.cc")])
    env = dict(os.environ, LD_PRELOAD=mock)
    out = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT, tmp=str(tmp_path))], env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().startswith("OK")


def test_without_gpu_the_device_builder_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    code = ("import sys; sys.path.insert(0, %r); import dc_lib_b200 as dc; conn, N = dc.kuhn_cube(8); "
            "dc.DC_create_tree_device(conn, conn.shape[0], 4, N)" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CUDA device" in out.stderr
