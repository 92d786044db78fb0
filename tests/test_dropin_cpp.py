"""Drop-in at the C++ boundary: ONE user program (tests/dropin/driver.cc, written against the
reference's include/DC.h and nothing else) is compiled and linked twice -- against the unmodified
reference (oracle/_ref) and against libdc_b200.so -- and the files the two executables write
(tree file, permuted coordinates, renumbered connectivity, row pointers, values assembled by
DC_tree_traversal with the program's own host callbacks) must be identical byte for byte."""
import os
import subprocess

import pytest

import dc_lib_b200 as dc
from common import ROOT, have_ref, ref_path

REF_INC = "/root/reference/include"
SRC = os.path.join(ROOT, "tests", "dropin", "driver.cc")


def _build(tmp_path, name, include, lib, defines):
    exe = os.path.join(str(tmp_path), name)
    libdir, libfile = os.path.split(lib)
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-ffp-contract=off"] + defines + ["-I" + include, SRC, "-o", exe,
                           "-L" + libdir, "-l:" + libfile, "-Wl,-rpath," + libdir, "-fopenmp"])
    return exe


def _run(exe, n, prefix):
    out = subprocess.check_output([exe, str(n), prefix], env=dict(os.environ, OMP_NUM_THREADS="4")).decode()
    with open(prefix + ".tree", "rb") as f:
        tree = f.read()
    with open(prefix + ".bin", "rb") as f:
        data = f.read()
    return out, tree, data


@pytest.mark.parametrize("n,vec", [(7, 0), (12, 0), (10, 4)])
def test_same_user_program_against_reference_and_product(tmp_path, n, vec):
    if not (have_ref(vec) and os.path.isdir(REF_INC)):
        pytest.skip("needs /root/reference and oracle/_ref (this container only)")
    defines = ["-DOMP", "-DTREE_CREATION", "-DBULK_SYNCHRONOUS"] + (["-DDC_VEC", f"-DVEC_SIZE={vec}"] if vec else [])
    # both builds use the REFERENCE's header: the product has to satisfy it as is
    ref_exe = _build(tmp_path, "driver_ref", REF_INC, ref_path(vec), defines)
    my_exe = _build(tmp_path, "driver_b200", REF_INC, dc.LIB_PATH, defines)
    env_vec = dict(os.environ)
    a = _run(ref_exe, n, os.path.join(str(tmp_path), "ref"))
    os.environ["DC_VEC_SIZE"] = str(vec)
    try:
        b = _run(my_exe, n, os.path.join(str(tmp_path), "mine"))
    finally:
        os.environ.clear()
        os.environ.update(env_vec)
    assert a[1] == b[1], "tree files differ"
    assert a[2] == b[2], "user arrays / assembled values differ"
    assert a[0] == b[0]


def test_user_program_builds_against_own_header(tmp_path):
    """Same program against the product's own include/DC.h (the only header on the GPU box)."""
    exe = _build(tmp_path, "driver_own", os.path.join(ROOT, "include"), dc.LIB_PATH, [])
    out, tree, data = _run(exe, 6, os.path.join(str(tmp_path), "own"))
    assert out.startswith("E=1296 N=343") and len(tree) > 4 + 4 * 1296 + 4 * 343
