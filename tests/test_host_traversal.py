"""Host-callback entry points of the product library (API drop-in for arbitrary host loops):
DC_tree_traversal (reference: src/tree_traversal.cc:106-117) and the README-generation
DC_assembly (README.md:120-140), driven with the oracle's callbacks.  These are NOT the
accelerated path (that is DC_assembly_device / dc_cuda_assemble, tested in the gpu files); the
test only checks that the drop-in schedules leaves like the reference does."""
import ctypes
from ctypes import byref, c_int, c_void_p

import numpy as np
import pytest

import common
import dc_lib_b200 as dc
from common import MineLib, make_user, oracle, oracle_serial, prepare_mesh


@pytest.mark.parametrize("n,vec,dim,threads", [(8, 0, 1, 1), (10, 0, 3, 4), (10, 4, 3, 4), (12, 2, 1, 8)])
def test_tree_traversal_with_host_callbacks(tmp_path, n, vec, dim, threads):
    lib = MineLib(vec)
    mesh = prepare_mesh(lib, n, tmp_path)
    dc.DC_set_num_threads(threads)
    values = np.full(mesh["nnz"] * dim * dim, np.nan)
    u = make_user(mesh, dim, values, reset_in_seq=0 if vec else 1)
    O = oracle()
    dc.lib().dc_tree_traversal_(ctypes.cast(O.dco_seq_callback, c_void_p), ctypes.cast(O.dco_vec_callback, c_void_p),
                                None, ctypes.cast(byref(u), c_void_p), None)
    assert values.tobytes() == oracle_serial(mesh, dim).tobytes()
    dc.DC_set_num_threads(0)


@pytest.mark.parametrize("n,vec,dim", [(8, 0, 1), (10, 4, 3)])
def test_readme_generation_assembly(tmp_path, n, vec, dim):
    """DC_assembly(userSeqFct, userVecFct, userArgs, nodeToNodeValue, operatorDim): the library
    zeroes each non-separator leaf's edge interval (operatorDim^2 doubles per edge) itself."""
    lib = MineLib(vec)
    mesh = prepare_mesh(lib, n, tmp_path)
    values = np.full(mesh["nnz"] * dim * dim, np.nan)
    u = make_user(mesh, dim, values)
    cb = ctypes.cast(oracle().dco_range_callback, c_void_p)
    dc.lib().dc_assembly_(cb, cb, ctypes.cast(byref(u), c_void_p), c_void_p(values.ctypes.data), byref(c_int(dim)))
    assert not np.isnan(values).any()
    assert values.tobytes() == oracle_serial(mesh, dim).tobytes()
