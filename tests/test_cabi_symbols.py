"""The C-ABI library loads without a GPU and exports every symbol the public headers
declare (include/dc_cuda.h, include/DC_f.h, DC_device.h's extern "C" block)."""
import ctypes
import os
import re

import pytest

import dc_lib_b200 as dc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_symbols(header, pattern):
    with open(os.path.join(ROOT, "include", header)) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(pattern, text)))


def test_library_loads_and_exports_c_abi():
    L = dc.lib()
    names = _c_symbols("dc_cuda.h", r"\b(dc_cuda_\w+)\s*\(")
    assert len(names) >= 15
    for name in names:
        assert hasattr(L, name), f"{name} declared in dc_cuda.h but not exported"


def test_library_exports_fortran_abi():
    L = dc.lib()
    names = _c_symbols("DC_f.h", r"\b(dc_\w+_)\s*\(") + _c_symbols("DC_device.h", r"\b(dc_\w+_)\s*\(")
    assert "dc_tree_traversal_" in names and "dc_assembly_device_" in names
    for name in names:
        assert hasattr(L, name), f"{name} declared but not exported"


def test_cpp_api_symbols_present():
    """C++-linkage names of include/DC.h (mangled) that a program linked against DC-lib needs."""
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", "-C", dc.LIB_PATH]).decode()
    for sig in ["DC_create_tree(int*, int, int, int)", "DC_tree_traversal(", "DC_permute_double_2d_array(double*, int, int)",
                "DC_permute_int_2d_array(int*, int*, int, int, int)", "DC_permute_int_1d_array(int*, int)",
                "DC_renumber_int_array(int*, int, bool)", "DC_create_permutation(int*, int*, int, int)",
                "DC_finalize_tree(int*, int*, int*, int*, int*, int*, int, int, int, int, int)", "DC_finalize_tree(int*, int*, int)",
                "DC_store_tree(std::", "DC_read_tree(std::", "DC_assembly(", "DC_create_nodeToElem(", "DC_create_elemToElem(",
                "DC_get_time()", "DC_get_cycles()", "DC_timer::start_time()", "DC_assembly_device(", "DC_device_set_mesh("]:
        assert sig in out, f"missing C++ symbol {sig}"


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        dc.Context(0)
