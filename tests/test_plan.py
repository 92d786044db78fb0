"""Invariants of the device schedule built on the host (host/dc_plan.cc)."""
import ctypes

import numpy as np
import pytest

import common
import dc_lib_b200 as dc
from common import MineLib, prepare_mesh, tree_records


def _plan_arrays(p):
    c = p.c
    units = [c.units[i] for i in range(c.nbUnits)]
    tasks = [c.tasks[i] for i in range(c.nbTasks)]
    items = np.ctypeslib.as_array(c.items, shape=(max(c.nbItems, 1),))[:c.nbItems]
    slot_elems = np.ctypeslib.as_array(c.slotElems, shape=(max(c.nbSlotNodes // 4, 1),))[:c.nbSlotNodes // 4]
    return units, tasks, items, slot_elems


@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("n,vec,slots,hub", [(8, 0, 832, False), (10, 4, 832, False), (10, 0, 200, False),
                                             (12, 0, 64, False), (18, 0, 832, True), (18, 0, 120, True)])
def test_plan_covers_every_contribution_once(tmp_path, n, vec, slots, hub, sym):
    mesh = prepare_mesh(MineLib(vec), n, tmp_path, hub=hub)
    rec = tree_records(mesh)
    plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, rec, slots, symmetric=sym)
    check_plan(mesh, plan, slots, sym)


@pytest.mark.parametrize("sym", [True, False])
def test_leaves_sized_for_the_unit_budget_split_evenly(tmp_path, sym):
    """bench.py's elasticity trees: leaf capacity chosen so that a leaf has about 62 nodes (two node
    parts); such a leaf exceeds the 512-slot budget and is cut into two pieces of at most 32 rows --
    one full diagonal task per unit in the symmetric schedule -- instead of a greedy 40 + 22."""
    import bench
    n = 20
    dc.DC_set_max_elem_per_part(bench.leaf_capacity(n, 62))
    try:
        mesh = prepare_mesh(MineLib(0, geometric=True), n, tmp_path)
    finally:
        dc.DC_set_max_elem_per_part(200)
    rec = tree_records(mesh)
    plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, rec, 512, symmetric=sym)
    check_plan(mesh, plan, 512, sym)
    units, tasks, _, _ = _plan_arrays(plan)
    leaves = rec[(rec[:, 6] != 0) & (rec[:, 5] == 0)]
    assert ((leaves[:, 4] - leaves[:, 3] + 1) <= 64).mean() > 0.9          # the capacity does what bench.py wants
    rows = np.array([u.nodeCount for u in units])
    assert (rows <= 32).mean() > 0.9
    if sym:
        diag_tasks = sum(1 for t in tasks if t.kind == 3)
        assert diag_tasks <= 1.1 * len(units)
    # the greedy cut (DC_PLAN_GREEDY_SPLIT=1) is still a valid schedule, with more padding
    import os
    os.environ["DC_PLAN_GREEDY_SPLIT"] = "1"
    try:
        greedy = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, rec, 512, symmetric=sym)
    finally:
        del os.environ["DC_PLAN_GREEDY_SPLIT"]
    if sym:
        assert greedy.c.nbItems > plan.c.nbItems and greedy.c.nbItemsUseful == plan.c.nbItemsUseful


def check_plan(mesh, plan, slots, sym):
    units, tasks, items, slot_elems = _plan_arrays(plan)
    tgt = np.ctypeslib.as_array(plan.c.tgt, shape=(max(plan.c.nbTgt, 1), 2))
    slotNodes = np.ctypeslib.as_array(plan.c.slotNodes, shape=(max(plan.c.nbSlotNodes, 1),))
    haloNodes = np.ctypeslib.as_array(plan.c.haloNodes, shape=(max(plan.c.nbHaloNodes, 1),))
    rows_of = np.repeat(np.arange(mesh["N"]), np.diff(mesh["row"]))
    E, N, nnz = mesh["E"], mesh["N"], mesh["nnz"]
    conn = mesh["conn"].astype(np.int64) - 1
    row, col = mesh["row"], mesh["col"]
    # expected number of contributions per edge
    want = np.zeros(nnz, dtype=np.int64)
    pos = {}
    for i in range(N):
        for k in range(row[i], row[i + 1]):
            pos[(i, int(col[k]))] = k
    for e in range(E):
        for a in range(4):
            for b in range(4):
                want[pos[(int(conn[e, a]), int(conn[e, b]))]] += 1
    got = np.zeros(nnz, dtype=np.int64)
    rows_seen = np.zeros(N, dtype=np.int64)
    for u in units:
        rows_seen[u.nodeFirst:u.nodeFirst + u.nodeCount] += 1
        assert u.edgeFirst == row[u.nodeFirst] and u.edgeCount == row[u.nodeFirst + u.nodeCount] - row[u.nodeFirst]
        nslots = u.ownElemCount + u.haloCount
        assert nslots <= slots
        # slots are numbered for the shared-memory banks: a permutation of own elements + halo
        elems = slot_elems[u.slotFirst:u.slotFirst + nslots].tolist()
        own = set(range(u.ownElemFirst, u.ownElemFirst + u.ownElemCount))
        assert len(set(elems)) == nslots and own <= set(elems)
        assert all(e in own or any(u.nodeFirst <= nd < u.nodeFirst + u.nodeCount for nd in conn[e]) for e in elems)
        # node table: own interval from the even node at or below nodeFirst, then the halo nodes
        aligned = u.nodeFirst & ~1
        span = ((u.nodeFirst + u.nodeCount + 1) & ~1) - aligned
        hn = haloNodes[u.hnodeFirst:u.hnodeFirst + u.hnodeCount]
        sn = slotNodes[u.slotFirst * 4:(u.slotFirst + nslots) * 4].reshape(nslots, 4).astype(np.int64)
        resolved = np.where(sn < span, sn + aligned, hn[np.clip(sn - span, 0, max(len(hn) - 1, 0))] if len(hn) else sn + aligned)
        assert np.array_equal(resolved, conn[elems])
        assert u.slotFirst % 2 == 0 and u.hnodeFirst % 4 == 0 and u.hnodeCount % 4 == 0
        for t in tasks[u.taskFirst:u.taskFirst + u.taskCount]:
            block = items[u.itemFirst + t.itemRel:u.itemFirst + t.itemRel + 32 * t.iters].reshape(t.iters, 32)
            assert (t.kind in (2, 3)) == sym
            for lane in range(32):
                kT = -1
                if t.kind == 0:
                    if (t.mask >> lane) & 1 or t.first + lane >= u.edgeCount:
                        assert (block[:, lane] == (nslots << 4)).all()
                        continue
                    k = u.edgeFirst + t.first + lane
                elif t.kind in (2, 3):
                    if not (t.mask >> lane) & 1:
                        assert (block[:, lane] == (nslots << 4)).all()
                        continue
                    assert t.first + lane < u.tgtCount
                    k, kT = (int(v) for v in tgt[u.tgtFirst + t.first + lane])
                    assert u.edgeFirst <= k < u.edgeFirst + u.edgeCount and col[k] >= rows_of[k]
                    assert (t.kind == 3) == (col[k] == rows_of[k])              # diagonals have their own tasks
                    if col[k] == rows_of[k]:
                        assert kT == -1                                          # a diagonal edge
                    else:
                        assert rows_of[kT] == col[k] and col[kT] == rows_of[k]  # the transposed edge
                else:
                    if not (t.mask >> lane) & 1:
                        assert (block[:, lane] == (nslots << 4)).all()
                        continue
                    node = u.nodeFirst + t.first + lane
                    k = u.edgeFirst + plan.c.diagRel[node]
                    assert col[k] == node
                i_row = int(np.searchsorted(row, k, side="right") - 1)
                last = -1
                for it in block[:, lane]:
                    if it >> 4 == nslots:               # padding: the kernel's all-zero record
                        continue
                    e = elems[it >> 4]
                    a, b = (it >> 2) & 3, it & 3
                    assert conn[e, a] == i_row and conn[e, b] == col[k]
                    assert e > last                      # ascending element order inside an edge
                    last = e
                    got[k] += 1
                    if kT >= 0:
                        got[kT] += 1
    big = plan.c.big
    for q in range(big.nbEdges):
        k = big.edge[q]
        got[k] += big.ptr[q + 1] - big.ptr[q]
        if sym and big.edgeT[q] >= 0:
            got[big.edgeT[q]] += big.ptr[q + 1] - big.ptr[q]
    covered = rows_seen.copy()
    for q in range(big.nbEdges):
        covered[int(np.searchsorted(row, big.edge[q], side="right") - 1)] = 1
    assert (covered == 1).all()
    assert np.array_equal(got, want)
    assert plan.c.nbItemsUseful + big.nbItems == (10 if sym else 16) * E


@pytest.mark.parametrize("sym", [True, False])
def test_plan_numbered_for_the_laplacian_records(tmp_path, sym):
    """bank_model=1 (11-word records of 4x4 symmetric entries): same units and contributions, other slot numbers."""
    mesh = prepare_mesh(MineLib(0), 10, tmp_path)
    rec = tree_records(mesh)
    # one builder thread: the position of a unit's slices in the arrays then does not depend on scheduling
    plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, rec, 832, nbThreads=1, symmetric=sym, bank_model=1)
    check_plan(mesh, plan, 832, sym)
    ref = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, rec, 832, nbThreads=1, symmetric=sym)
    assert plan.c.nbUnits == ref.c.nbUnits and plan.c.nbItems == ref.c.nbItems
    a, b = plan.arrays(), ref.arrays()
    assert not np.array_equal(a["items"], b["items"])
    assert np.array_equal(a["items"] & 15, b["items"] & 15)            # same (a, b) pairs, other slots


def test_plan_without_tree_uses_row_blocks(tmp_path):
    mesh = prepare_mesh(MineLib(0), 8, tmp_path)
    plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, None, 300, symmetric=False)
    assert plan.c.nbUnits > 1 and plan.c.nbItemsUseful == 16 * mesh["E"]


def test_list_items_refuse_meshes_beyond_28_bit_element_indices():
    """Fallback / DC_MODE_LIST items are elem << 4 | a << 2 | b in 32 bits: dc_contrib_build must refuse 2^28 or
    more elements instead of truncating the element index (checked before any array is touched)."""
    import ctypes
    cl = dc.ContribList()
    conn = np.ones((1, 4), dtype=np.int32)
    row = np.zeros(2, dtype=np.int32)
    col = np.zeros(1, dtype=np.int32)
    rc = dc.lib().dc_contrib_build(ctypes.byref(cl), conn.ctypes.data, 1 << 28, 1, row.ctypes.data, col.ctypes.data, 0,
                                   None, 0, 1)
    assert rc == -5
