"""CSR pattern and node -> element lists built on the GPU (dc_cuda_csr_*, csrc/dc_csr.cuh) against the
host builder and a numpy restatement: bit-exact index arrays (the user-side
step before DC_finalize_tree; reference helper DC_create_nodeToElem, src/coloring.cc:79-112)."""
import numpy as np
import pytest

import dc_lib_b200 as dc
from common import MineLib, prepare_mesh

pytestmark = pytest.mark.gpu


def _numpy_csr(conn, N, colBase):
    """Independent restatement: edge (i, j) iff an element holds both nodes (diagonal included),
    columns ascending inside a row -- the 16 node pairs of every element, sorted and made unique."""
    c = conn.astype(np.int64) - 1
    key = np.unique((np.repeat(c, 4, axis=1) * N + np.tile(c, (1, 4))).reshape(-1))
    row = np.concatenate(([0], np.cumsum(np.bincount(key // N, minlength=N)))).astype(np.int32)
    return row, (key % N + colBase).astype(np.int32)


def _check(conn, N, colBase):
    import torch
    want_row, want_col = dc.build_csr(conn, N, colBase) if conn.shape[0] < 20000 else dc.build_csr_fast(conn, N, colBase)
    np_row, np_col = _numpy_csr(conn, N, colBase)            # the checker of both builders
    assert np.array_equal(want_row, np_row) and np.array_equal(want_col, np_col)
    d_conn = torch.from_numpy(conn).cuda()
    row, col, ptr, val = dc.build_csr_device(d_conn, N, colBase, node_to_elem=True)
    assert np.array_equal(row.cpu().numpy(), want_row)
    assert np.array_equal(col.cpu().numpy(), want_col)
    # node -> element lists: ascending elements per node
    flat = conn.reshape(-1).astype(np.int64) - 1
    order = np.argsort(flat, kind="stable")
    assert np.array_equal(ptr.cpu().numpy(), np.concatenate(([0], np.cumsum(np.bincount(flat, minlength=N)))))
    assert np.array_equal(val.cpu().numpy(), (order // 4).astype(np.int32))


@pytest.mark.parametrize("n,colBase", [(1, 0), (3, 1), (10, 0), (24, 0), (40, 1)])
def test_kuhn_cube_pattern(n, colBase):
    conn, N = dc.kuhn_cube(n)
    _check(conn, N, colBase)


def test_renumbered_tree_order_mesh_and_hubs(tmp_path):
    mesh = prepare_mesh(MineLib(0), 12, tmp_path, scramble=True)        # irregular numbering
    _check(mesh["conn"], mesh["N"], 0)
    assert np.array_equal(dc.build_csr_fast(mesh["conn"], mesh["N"])[1], mesh["col"])
    hub = prepare_mesh(MineLib(0), 18, tmp_path, hub=True)               # valence ~100: rows beyond the staging area
    _check(hub["conn"], hub["N"], 0)


def test_empty_and_bad_input():
    import torch
    row, col = dc.build_csr_device(torch.zeros((0, 4), dtype=torch.int32, device="cuda"), 5)
    assert row.cpu().tolist() == [0] * 6 and col.numel() == 0
    with pytest.raises(RuntimeError):
        dc.build_csr_device(torch.tensor([[1, 2, 3, 9]], dtype=torch.int32, device="cuda"), 4)
