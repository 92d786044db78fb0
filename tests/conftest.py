import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the checkers (oracle restatement, and the reference itself when /root/reference
    # exists) and the product library are built on demand
    if not os.path.exists(os.path.join(ROOT, "oracle", "libdc_oracle.so")) or \
       not os.path.exists(os.path.join(ROOT, "dc-lib_b200", "libdc_b200.so")):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
