"""Pins the ORACLE: the plain-C restatement (oracle/dc_oracle.c) against the unmodified
reference compiled into oracle/_ref (only where /root/reference existed at build time;
the committed golden vectors cover the GPU box, see test_golden.py)."""
import numpy as np
import pytest

import common
from common import RefLib, have_ref, oracle, oracle_serial, oracle_traverse, parse_tree, prepare_mesh

needs_ref = pytest.mark.skipif(not have_ref(0), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
@pytest.mark.parametrize("n,vec,dim", [(6, 0, 1), (10, 0, 1), (10, 0, 3), (10, 4, 3), (12, 2, 1), (16, 4, 1), (12, 8, 3)])
def test_traversal_restatement_matches_reference(tmp_path, n, vec, dim):
    """tree_traversal.cc:27-117 run by the reference itself (OpenMP tasks) with the oracle's
    callbacks == the oracle's serial restatement of the walk == the ascending element loop."""
    ref = RefLib(vec)
    mesh = prepare_mesh(ref, n, tmp_path)
    got_ref = ref.traversal(mesh, dim)
    nodes, cnt, _, _ = parse_tree(mesh["tree"], mesh["E"], mesh["N"])
    got_walk = oracle_traverse(mesh, dim, nodes, split=vec > 0)
    got_loop = oracle_serial(mesh, dim)
    assert not np.isnan(got_ref).any()
    assert got_ref.tobytes() == got_walk.tobytes()
    assert got_ref.tobytes() == got_loop.tobytes()


@needs_ref
def test_cilk_serial_elision_matches_openmp(tmp_path):
    if not have_ref(0, "cilkser"):
        pytest.skip("cilk serial elision build missing")
    a = prepare_mesh(RefLib(0, "omp"), 10, tmp_path, tag="omp")
    b = prepare_mesh(RefLib(0, "cilkser"), 10, tmp_path, tag="cilk")
    assert a["tree"] == b["tree"]
    assert RefLib(0, "cilkser").L is not None


@needs_ref
@pytest.mark.parametrize("n,vec", [(6, 0), (10, 4)])
def test_tree_file_parser_and_finalize(tmp_path, n, vec):
    """IO.cc format + tree_creation.cc:185-186 edge intervals, restated."""
    ref = RefLib(vec)
    mesh = prepare_mesh(ref, n, tmp_path)
    nodes, cnt, elemPerm, nodePerm = parse_tree(mesh["tree"], mesh["E"], mesh["N"])
    assert sorted(elemPerm.tolist()) == list(range(mesh["E"]))
    assert sorted(nodePerm.tolist()) == list(range(mesh["N"]))
    stored = [(nodes[i].firstEdge, nodes[i].lastEdge) for i in range(cnt)]
    oracle().dco_finalize(nodes, cnt, mesh["row"].ctypes.data)
    again = [(nodes[i].firstEdge, nodes[i].lastEdge) for i in range(cnt)]
    assert stored == again
    # leaves are exactly the nodes without children and cover all elements once
    covered = np.zeros(mesh["E"], dtype=np.int32)
    for i in range(cnt):
        if nodes[i].isLeaf:
            covered[nodes[i].firstElem:nodes[i].lastElem + 1] += 1
    assert (covered == 1).all()


@needs_ref
def test_permutation_restatements_match_reference(tmp_path):
    """permutations.cc:22-180, literally restated, against the reference's own functions."""
    ref = RefLib(0)
    mesh = prepare_mesh(ref, 6, tmp_path)          # leaves the permutations freed (IO.cc:157) ...
    ref.read(mesh["tree_path"], mesh["E"], mesh["N"])   # ... so load them back
    _, _, elemPerm, nodePerm = parse_tree(mesh["tree"], mesh["E"], mesh["N"])
    rng = np.random.default_rng(3)
    N, E = mesh["N"], mesh["E"]
    O = oracle()

    d = rng.standard_normal((N, 3))
    a, b = d.copy(), d.copy()
    ref.permute_double_2d(a, N, 3)
    O.dco_permute_double_2d(b.ctypes.data, nodePerm.ctypes.data, N, 3)
    assert a.tobytes() == b.tobytes()
    assert np.array_equal(a[nodePerm], d)            # new[perm[i]] = old[i]

    t = rng.integers(0, 1000, size=(E, 4)).astype(np.int32)
    a, b = t.copy(), t.copy()
    ref.permute_int_2d(a, E, 4)
    O.dco_permute_int_2d(b.ctypes.data, elemPerm.ctypes.data, E, 4, 0)
    assert a.tobytes() == b.tobytes()

    v = rng.integers(0, 1000, size=N).astype(np.int32)
    a, b = v.copy(), v.copy()
    ref.permute_int_1d(a, N)
    O.dco_permute_int_1d(b.ctypes.data, nodePerm.ctypes.data, N)
    assert a.tobytes() == b.tobytes()

    r = rng.integers(1, N + 1, size=5000).astype(np.int32)
    a, b = r.copy(), r.copy()
    ref.renumber(a, a.size)
    O.dco_renumber(b.ctypes.data, nodePerm.ctypes.data, b.size, 1)
    assert a.tobytes() == b.tobytes()

    part = rng.integers(0, 7, size=3000).astype(np.int32)
    a = ref.create_permutation(part, 7)
    b = np.full(part.size, -1, dtype=np.int32)
    O.dco_create_permutation(b.ctypes.data, part.ctypes.data, part.size, 7)
    assert a.tobytes() == b.tobytes()
