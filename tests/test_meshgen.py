"""The C++ mesh / CSR helpers used for large runs reproduce the numpy definitions."""
import numpy as np
import pytest

import dc_lib_b200 as dc


@pytest.mark.parametrize("n", [1, 3, 7, 12])
def test_fast_kuhn_cube_matches_numpy(n):
    conn, N = dc.kuhn_cube(n)
    coord = dc.jitter_coords(n)
    conn2, coord2, N2 = dc.kuhn_cube_fast(n)
    assert N == N2 and conn.tobytes() == conn2.tobytes() and coord.tobytes() == coord2.tobytes()
    # every tetrahedron has positive volume magnitude and 4 distinct nodes
    c = conn - 1
    assert (np.sort(c, axis=1)[:, 1:] != np.sort(c, axis=1)[:, :-1]).all()
    assert conn.shape[0] == 6 * n ** 3


@pytest.mark.parametrize("n", [2, 5, 9])
def test_fast_csr_matches_numpy(n):
    conn, N = dc.kuhn_cube(n)
    rng = np.random.default_rng(n)
    relabel = rng.permutation(N).astype(np.int32)
    conn = np.ascontiguousarray(relabel[conn - 1] + 1, dtype=np.int32)      # scramble node ids
    row, col = dc.build_csr(conn, N)
    row2, col2 = dc.build_csr_fast(conn, N)
    assert row.tobytes() == row2.tobytes() and col.tobytes() == col2.tobytes()
    # Kuhn cube: nnz = N + 2 * #edges
    assert row[-1] == N + 2 * (3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3)
