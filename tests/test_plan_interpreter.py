"""The device schedule executed on the CPU, lane by lane, with the ORACLE's element arithmetic.

gather_units (csrc/dc_kernels.cuh) does, per lane of a task, exactly this: start from 0.0, apply
the lane's items in row order (item = slot<<4 | a<<2 | b; padding items add exact zeros), store the
block at the lane's edge and -- symmetric schedule -- its transpose at the transposed edge.  This
test interprets the schedule that way with the oracle's operators (oracle/dc_oracle.c: the numeric
contract shared with the device operators) and compares the result BITWISE with the reference
order, i.e. the oracle's serial loop over the elements (identical to the
reference's parallel traversal, tree_traversal.cc:27-103).  It pins, without a GPU, that a schedule
-- unit selection, even leaf split, bank-aware slot numbering, symmetric ownership -- reproduces the
reference's bits; the GPU tests then only have to show that the kernel follows the schedule.
"""
import ctypes

import numpy as np
import pytest

import common
import dc_lib_b200 as dc
from common import LAMBDA, MU, MineLib, oracle_serial, prepare_mesh, tree_records


def _dp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def run_schedule(mesh, plan, dim):
    O = common.oracle()
    O.dco_laplacian_ke.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    O.dco_elasticity_h.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    O.dco_elasticity_add.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    c = plan.c
    E, N, nnz = mesh["E"], mesh["N"], mesh["nnz"]
    conn = mesh["conn"].astype(np.int64) - 1
    X = np.ascontiguousarray(mesh["coord"].reshape(N, 3)[conn])           # [E,4,3]
    per = dim * dim
    if dim == 1:
        rec = np.zeros((E, 16))
        for e in range(E):
            O.dco_laplacian_ke(_dp(X[e]), _dp(rec[e]))
    else:
        rec = np.zeros((E, 4, 3))
        for e in range(E):
            O.dco_elasticity_h(_dp(X[e]), _dp(rec[e]))
    values = np.full((nnz, per), np.nan)
    written = np.zeros(nnz, dtype=np.int64)
    v9 = np.zeros(9)

    def lane_sum(contribs):
        """contribs: (element, a, b) in schedule order -> the block the lane stores"""
        if dim == 1:
            v = 0.0
            for e, a, b in contribs:
                v = v + rec[e, a * 4 + b]
            return np.array([v])
        v9[:] = 0.0
        for e, a, b in contribs:
            O.dco_elasticity_add(_dp(rec[e, a]), _dp(rec[e, b]), LAMBDA, MU, _dp(v9))
        return v9.copy()

    def store(k, kT, block):
        values[k] = block
        written[k] += 1
        if kT >= 0:
            values[kT] = block.reshape(dim, dim).T.reshape(-1)
            written[kT] += 1

    items = np.ctypeslib.as_array(c.items, shape=(max(c.nbItems, 1),))
    slot_elems = np.ctypeslib.as_array(c.slotElems, shape=(max(c.nbSlotNodes // 4, 1),))
    tgt = np.ctypeslib.as_array(c.tgt, shape=(max(c.nbTgt, 1), 2))
    diag_rel = plan.diag_rel(N)
    for ui in range(c.nbUnits):
        u = c.units[ui]
        nslots = u.ownElemCount + u.haloCount
        elems = slot_elems[u.slotFirst:u.slotFirst + nslots]
        for ti in range(u.taskFirst, u.taskFirst + u.taskCount):
            t = c.tasks[ti]
            block = items[u.itemFirst + t.itemRel:u.itemFirst + t.itemRel + 32 * t.iters].reshape(t.iters, 32)
            for lane in range(32):
                kT = -1
                if t.kind == 0:                                         # 32 consecutive edges
                    if (t.mask >> lane) & 1 or t.first + lane >= u.edgeCount:
                        continue
                    k = u.edgeFirst + t.first + lane
                elif t.kind == 1:                                       # diagonals of 32 consecutive nodes
                    if not (t.mask >> lane) & 1:
                        continue
                    k = u.edgeFirst + diag_rel[u.nodeFirst + t.first + lane]
                else:                                                   # owned edges of the symmetric schedule
                    if not (t.mask >> lane) & 1:
                        continue
                    k, kT = (int(x) for x in tgt[u.tgtFirst + t.first + lane])
                col = [(int(elems[it >> 4]), (it >> 2) & 3, it & 3) for it in block[:, lane].tolist() if (it >> 4) < nslots]
                store(k, kT, lane_sum(col))
    big = c.big
    for q in range(big.nbEdges):                                        # rows that did not fit a unit
        contribs = [(int(it >> 4), (int(it) >> 2) & 3, int(it) & 3) for it in
                    np.ctypeslib.as_array(big.item, shape=(big.nbItems,))[big.ptr[q]:big.ptr[q + 1]]]
        store(int(big.edge[q]), int(big.edgeT[q]) if c.symmetric else -1, lane_sum(contribs))
    assert (written == 1).all()              # every value written exactly once: the reset is fused
    return values.reshape(-1)


CASES = [
    # n, colouring vector, slots, hub mesh, geometric tree, nodes per leaf (0 = capacity 200)
    (8, 0, 832, False, False, 0),
    (10, 4, 200, False, False, 0),
    (12, 0, 512, False, True, 62),          # bench.py's elasticity geometry: even split of the leaves
    (12, 0, 64, False, True, 0),
    (18, 0, 120, True, False, 0),           # hub rows: fallback lists
]


@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("dim", [1, 3])
@pytest.mark.parametrize("n,vec,slots,hub,geometric,leaf_nodes", CASES)
def test_schedule_reproduces_reference_bits(tmp_path, n, vec, slots, hub, geometric, leaf_nodes, dim, sym):
    import bench
    if hub and dim == 3 and not sym:
        pytest.skip("covered by the symmetric hub case; keeps the CPU suite short")
    dc.DC_set_max_elem_per_part(bench.leaf_capacity(n, leaf_nodes))
    try:
        mesh = prepare_mesh(MineLib(vec, geometric=geometric), n, tmp_path, hub=hub)
    finally:
        dc.DC_set_max_elem_per_part(200)
    # the slot numbering is chosen for the operator's record layout, as bench.py does
    plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, tree_records(mesh), slots, symmetric=sym,
                       bank_model=1 if dim == 1 else 0)
    got = run_schedule(mesh, plan, dim)
    want = oracle_serial(mesh, dim)
    assert got.tobytes() == want.tobytes()


@pytest.mark.skipif(not common.have_ref(0), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("dim", [1, 3])
def test_reference_on_the_bench_tree_equals_the_schedule(tmp_path, dim):
    """The loop closed for bench.py's geometry: a coordinate-bisection tree with CTA-sized leaves,
    built by the product, is handed to the UNMODIFIED reference as a tree file (what the bench's CPU
    arm does); the reference's own DC_tree_traversal over it, the oracle's serial loop and the device
    schedule interpreted on the CPU give the same bits."""
    import bench
    n = 14
    dc.DC_set_max_elem_per_part(bench.leaf_capacity(n, 62))
    try:
        mesh = prepare_mesh(MineLib(0, geometric=True), n, tmp_path)        # stores the tree file
    finally:
        dc.DC_set_max_elem_per_part(200)
    ref = common.RefLib(0)
    ref.read(mesh["tree_path"], mesh["E"], mesh["N"])
    got_ref = ref.traversal(mesh, dim)
    assert not np.isnan(got_ref).any()
    want = oracle_serial(mesh, dim)
    assert got_ref.tobytes() == want.tobytes()
    plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, tree_records(mesh), 512,
                       bank_model=1 if dim == 1 else 0)
    assert run_schedule(mesh, plan, dim).tobytes() == got_ref.tobytes()
