"""Product host library (tree, permutations, colouring, tree file) against the golden
vectors generated from the reference (tests/golden/golden.json, oracle/gen_golden.py) and,
where oracle/_ref exists, against the reference run side by side.  Bit-exact throughout."""
import json
import os

import numpy as np
import pytest

import common
import dc_lib_b200 as dc
from common import GOLDEN, MineLib, RefLib, have_ref, md5, oracle, parse_tree, prepare_mesh

with open(os.path.join(GOLDEN, "golden.json")) as f:
    GOLD = json.load(f)

FAST_KEYS = [k for k, v in GOLD.items() if v["n"] <= 20]


@pytest.mark.parametrize("key", FAST_KEYS)
def test_tree_file_matches_golden(tmp_path, key):
    g = GOLD[key]
    mesh = prepare_mesh(MineLib(g["vec"]), g["n"], tmp_path, hub=g["hub"])
    assert (mesh["E"], mesh["N"], mesh["nnz"]) == (g["E"], g["N"], g["nnz"])
    assert md5(mesh["tree"]) == g["tree_md5"]          # elemPerm + nodePerm + every tree record
    assert md5(mesh["conn"]) == g["conn_md5"]          # element order and renumbering
    assert md5(mesh["coord"]) == g["coord_md5"]        # node permutation applied to doubles
    raw = os.path.join(GOLDEN, f"{key}_tree.bin")
    if os.path.exists(raw):
        with open(raw, "rb") as f:
            assert f.read() == mesh["tree"]


@pytest.mark.parametrize("key", ["kuhn40_vec0", "kuhn40_vec4", "kuhn64_vec0", "hub64_vec0"])
def test_tree_file_matches_golden_c1_c2(tmp_path, key):
    """BASELINE configs C1 / C2: the 40^3 cube, pure D&C and D&C + colouring (VEC_SIZE 4); the 64^3 cube
    (1.57 M tets, the largest size SURVEY.md 4 asks a structural golden for) and the 64^3 hub mesh of
    config C5 (valence ~100 hubs): tree file, element order and renumbering equal the reference's bytes."""
    g = GOLD[key]
    mesh = prepare_mesh(MineLib(g["vec"]), g["n"], tmp_path, hub=g["hub"])
    assert (mesh["E"], mesh["N"], mesh["nnz"]) == (g["E"], g["N"], g["nnz"])
    assert md5(mesh["tree"]) == g["tree_md5"]
    assert md5(mesh["conn"]) == g["conn_md5"]


@pytest.mark.skipif(not have_ref(0), reason="oracle/_ref not built")
@pytest.mark.parametrize("n,vec", [(4, 0), (5, 0), (5, 4), (7, 2), (9, 8), (13, 4)])
def test_tree_file_matches_reference_side_by_side(tmp_path, n, vec):
    # (meshes of <= MAX_ELEM_PER_PART elements make the reference call METIS with one part,
    #  which leaves its partition vector uninitialised -- undefined there, so not compared)
    a = prepare_mesh(RefLib(vec), n, tmp_path, tag="ref")
    b = prepare_mesh(MineLib(vec), n, tmp_path, tag="mine")
    assert a["tree"] == b["tree"]
    assert a["conn"].tobytes() == b["conn"].tobytes()
    assert a["coord"].tobytes() == b["coord"].tobytes()


@pytest.mark.parametrize("n,vec", [(9, 0), (10, 4)])
def test_graded_shuffled_mesh_matches_reference(tmp_path, n, vec):
    """BASELINE config C4's irregular mesh at test size (graded, relabelled, shuffled): tree file,
    permuted connectivity and coordinates byte-identical to the unmodified reference."""
    if not have_ref(vec):
        pytest.skip("reference not built (oracle/_ref)")
    a = prepare_mesh(RefLib(vec), n, tmp_path, tag="ref", scramble=True)
    b = prepare_mesh(MineLib(vec), n, tmp_path, tag="mine", scramble=True)
    assert a["tree"] == b["tree"]
    assert a["conn"].tobytes() == b["conn"].tobytes()
    assert a["coord"].tobytes() == b["coord"].tobytes()


def test_store_read_store_roundtrip(tmp_path):
    lib = MineLib(4)
    mesh = prepare_mesh(lib, 8, tmp_path)
    lib.read(mesh["tree_path"], mesh["E"], mesh["N"])
    again = os.path.join(str(tmp_path), "again.bin")
    lib.store(again, mesh["E"], mesh["N"])
    with open(again, "rb") as f:
        assert f.read() == mesh["tree"]


def test_read_tree_then_permute_equals_create_tree(tmp_path):
    """Appendix A, alternative path: DC_read_tree + DC_permute_int_2d_array(elemPerm) reproduces
    the element order DC_create_tree left in elemToNode."""
    lib = MineLib(0)
    mesh = prepare_mesh(lib, 8, tmp_path)
    conn, N = dc.kuhn_cube(8)
    E = conn.shape[0]
    lib.read(mesh["tree_path"], E, N)
    lib.permute_int_2d(conn, E, 4)
    flat = conn.reshape(-1)
    lib.renumber(flat, 4 * E)
    assert conn.tobytes() == mesh["conn"].tobytes()


def test_permutation_family_matches_oracle(tmp_path):
    lib = MineLib(0)
    mesh = prepare_mesh(lib, 6, tmp_path)
    lib.read(mesh["tree_path"], mesh["E"], mesh["N"])
    _, _, elemPerm, nodePerm = parse_tree(mesh["tree"], mesh["E"], mesh["N"])
    rng = np.random.default_rng(11)
    N, E, O = mesh["N"], mesh["E"], oracle()
    for dim in (1, 3, 5):
        d = rng.standard_normal((N, dim))
        a, b = d.copy(), d.copy()
        lib.permute_double_2d(a, N, dim)
        O.dco_permute_double_2d(b.ctypes.data, nodePerm.ctypes.data, N, dim)
        assert a.tobytes() == b.tobytes()
    t = rng.integers(0, 1000, size=(E + 5, 4)).astype(np.int32)
    a, b = t.copy(), t.copy()
    lib.permute_int_2d(a, E, 4, offset=0)
    O.dco_permute_int_2d(b.ctypes.data, elemPerm.ctypes.data, E, 4, 0)
    assert a.tobytes() == b.tobytes()
    v = rng.integers(0, 1000, size=N).astype(np.int32)
    a, b = v.copy(), v.copy()
    lib.permute_int_1d(a, N)
    O.dco_permute_int_1d(b.ctypes.data, nodePerm.ctypes.data, N)
    assert a.tobytes() == b.tobytes()
    r = rng.integers(1, N + 1, size=4000).astype(np.int32)
    a, b = r.copy(), r.copy()
    lib.renumber(a, a.size)
    O.dco_renumber(b.ctypes.data, nodePerm.ctypes.data, b.size, 1)
    assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("size,nbPart", [(0, 3), (1, 1), (17, 4), (5000, 37)])
def test_create_permutation_matches_oracle(size, nbPart):
    rng = np.random.default_rng(size + nbPart)
    part = rng.integers(0, nbPart, size=size).astype(np.int32)
    a = dc.DC_create_permutation(part, nbPart)
    b = np.full(size, -1, dtype=np.int32)
    oracle().dco_create_permutation(b.ctypes.data, part.ctypes.data, size, nbPart)
    assert a.tobytes() == b.tobytes()


def test_tiny_meshes(tmp_path):
    """Meshes smaller than one partition: the tree is a single leaf."""
    for n in (1, 2, 3):
        mesh = prepare_mesh(MineLib(0), n, tmp_path)
        nodes, cnt, ep, npm = parse_tree(mesh["tree"], mesh["E"], mesh["N"])
        if mesh["E"] <= 200:
            assert cnt == 1 and nodes[0].isLeaf
            assert (nodes[0].firstEdge, nodes[0].lastEdge) == (0, mesh["nnz"] - 1)


@pytest.mark.parametrize("n,vec", [(6, 0), (12, 0), (12, 4), (20, 0)])
def test_geometric_tree_is_a_valid_dc_tree(tmp_path, n, vec):
    """DC_create_tree_geometric (coordinate bisection instead of METIS) must obey the D&C
    invariants the traversal relies on: leaves cover every element once; the elements of a
    non-separator node only touch that node's node interval; left/right node intervals split
    the parent's; a separator subtree inherits its parent's node interval."""
    mesh = prepare_mesh(MineLib(vec, geometric=True), n, tmp_path)
    nodes, cnt, elemPerm, nodePerm = parse_tree(mesh["tree"], mesh["E"], mesh["N"])
    assert sorted(elemPerm.tolist()) == list(range(mesh["E"]))
    assert sorted(nodePerm.tolist()) == list(range(mesh["N"]))
    conn = mesh["conn"] - 1
    covered = np.zeros(mesh["E"], dtype=np.int32)
    for i in range(cnt):
        t = nodes[i]
        if t.isLeaf:
            covered[t.firstElem:t.lastElem + 1] += 1
            assert t.lastElem - t.firstElem + 1 <= 200 or t.lastSep == t.lastElem
        if t.lastSep >= t.firstElem:
            sub = conn[t.firstElem:t.lastSep + 1]
            assert sub.min() >= t.firstNode and sub.max() <= t.lastNode
        if not t.isLeaf:
            l, r = nodes[t.left], nodes[t.right]
            if not t.isSep:
                assert (l.firstNode, r.lastNode) == (t.firstNode, t.lastNode) and l.lastNode + 1 == r.firstNode
            if t.sep >= 0:
                s = nodes[t.sep]
                assert s.isSep and (s.firstNode, s.lastNode) == (t.firstNode, t.lastNode)
                assert (s.firstElem, s.lastSep) == (t.lastElem + 1, t.lastSep)
    assert (covered == 1).all()
    # the reference's traversal semantics hold on it: restated walk == ascending loop
    walk = common.oracle_traverse(mesh, 1, nodes, split=vec > 0)
    assert walk.tobytes() == common.oracle_serial(mesh, 1).tobytes()


@pytest.mark.parametrize("nbPart,geometric", [(1, False), (2, False), (4, False), (3, True), (8, True)])
def test_partition_mesh(nbPart, geometric):
    """DC_partition_mesh: every element gets a part in range, no part is empty, the result is
    reproducible, one part means part 0, and an element lies in a part that holds at least as
    many of its nodes as any other part (the majority rule over the node partition)."""
    conn, _ = dc.kuhn_cube(7)
    coord = dc.jitter_coords(7)
    N = coord.shape[0]
    part = dc.DC_partition_mesh(conn, N, nbPart, coord if geometric else None)
    again = dc.DC_partition_mesh(conn, N, nbPart, coord if geometric else None)
    assert part.dtype == np.int32 and part.shape == (conn.shape[0],)
    assert np.array_equal(part, again)
    assert part.min() == 0 and part.max() == nbPart - 1
    counts = np.bincount(part, minlength=nbPart)
    assert counts.min() > 0
    assert counts.max() <= 2.0 * conn.shape[0] / nbPart          # recursive bisection: roughly balanced
    if nbPart == 1:
        assert not part.any()
    # a node partition consistent with the element partition exists: nodes whose elements all lie
    # in one part pin that part; every element then has at least one node pinned to its own part
    # unless it is an interface element
    node_parts = [set() for _ in range(N)]
    for e, p in enumerate(part):
        for n in conn[e] - 1:
            node_parts[n].add(int(p))
    interior = np.array([len(s) == 1 for s in node_parts])
    assert interior.sum() > 0.5 * N or nbPart >= 8


@pytest.mark.parametrize("geometric", [True, False])
def test_tree_and_schedule_do_not_depend_on_the_thread_count(tmp_path, geometric):
    """Tree builder (task-parallel recursion over shared scratch) and schedule builder (units built by
    whichever thread takes them): one thread and all threads give the same tree file bytes and, per unit,
    the same schedule (tools/plan_hash.canonical_md5)."""
    import sys
    sys.path.insert(0, os.path.join(common.ROOT, "tools"))
    from plan_hash import canonical_md5
    out = []
    for threads in (1, 0):
        dc.DC_set_num_threads(threads)
        try:
            mesh = prepare_mesh(MineLib(0, geometric=geometric), 14, tmp_path, tag=f"thr{threads}")
        finally:
            dc.DC_set_num_threads(0)
        plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, common.tree_records(mesh), 160,
                           nbThreads=1 if threads == 1 else 0)
        out.append((mesh["tree"], mesh["conn"].tobytes(), canonical_md5(plan, mesh["N"])))
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1]
    assert out[0][2] == out[1][2]


def test_geometric_tree_of_an_irregular_part_terminates():
    """Regression: one METIS part of the graded / relabelled / shuffled cube (BASELINE config 4) has separators
    in which every element straddles a coordinate cut; the recursion used to cut the same set again for ever
    (stack overflow at 80^3).  Such a set now stays one leaf, and the tree still covers every element once."""
    from dc_lib_b200 import multigpu
    conn, coord, N = dc.kuhn_cube_fast(80)
    conn, coord = dc.scramble_mesh(conn, coord)
    part = dc.DC_partition_mesh(conn, N, 4)
    dc.DC_set_max_elem_per_part(182)
    try:
        p = multigpu.build_part(conn, coord, part, 1, 4, builder="geometric")
    finally:
        dc.DC_set_max_elem_per_part(200)
    rec = p["rec"]
    leaves = rec[rec[:, 6] == 1]
    sizes = leaves[:, 1] - leaves[:, 0] + 1
    assert sizes.sum() == p["E"] and sizes.min() >= 0
    covered = np.zeros(p["E"], dtype=np.int32)
    for first, last in leaves[:, :2]:
        covered[first:last + 1] += 1
    assert (covered == 1).all()
    dc.DC_free_tree()
