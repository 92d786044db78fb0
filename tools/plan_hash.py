#!/usr/bin/env python
"""tools/plan_hash.py -- canonical md5 of the device schedule (host/dc_plan.cc) of a set of meshes.

The position of a unit's slices inside the schedule arrays depends on which host thread built the
unit, so the raw arrays differ from run to run; the CONTENT the kernel sees per unit does not.  The
hash walks the units in order and digests, per unit, its descriptor without the array offsets and
its slices of tasks / items / slotNodes / slotElems / haloNodes / tgt, then diagRel, the fallback
lists and the plan-wide maxima.  A change of the builder that is meant to be a pure speed-up must
leave every line of this tool's output unchanged (run it before and after).

    python tools/plan_hash.py            # the standard set (seconds)
    python tools/plan_hash.py 96 512 1   # one cube: n, unit_slots, symmetric
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dc_lib_b200 as dc  # noqa: E402


def canonical_md5(plan, nbNodes):
    c = plan.c
    a = plan.arrays()
    h = hashlib.md5()
    units = a["units"].view(np.int64).reshape(-1, 12) if c.nbUnits else np.zeros((0, 12), np.int64)
    u32 = a["units"].view(np.int32).reshape(-1, 24) if c.nbUnits else np.zeros((0, 24), np.int32)
    tasks = a["tasks"].reshape(-1, 16)
    slot_elems = np.ctypeslib.as_array(c.slotElems, shape=(max(c.nbSlotNodes // 4, 1),))
    for i in range(c.nbUnits):
        item0, tgt0, slot0, hn0 = (int(units[i, k]) for k in (1, 2, 3, 4))
        w = u32[i]
        # int32 fields: 12 edgeCount 13 nodeFirst 14 nodeCount 15 ownElemFirst 16 ownElemCount 17 haloCount
        #               18 taskFirst 19 taskCount 20 itemCount 21 tgtCount 22 hnodeCount
        h.update(units[i, 0:1].tobytes())
        h.update(w[[12, 13, 14, 15, 16, 17, 19, 20, 21, 22]].tobytes())
        nslots = int(w[16] + w[17])
        h.update(tasks[int(w[18]):int(w[18] + w[19])].tobytes())
        h.update(a["items"][item0:item0 + int(w[20])].tobytes())
        h.update(a["slotNodes"][slot0 * 4:(slot0 + nslots) * 4].tobytes())
        h.update(slot_elems[slot0:slot0 + nslots].tobytes())
        h.update(a["haloNodes"][hn0:hn0 + int(w[22])].tobytes())
        if c.symmetric:
            h.update(a["tgt"][tgt0:tgt0 + int(w[21])].tobytes())
    h.update(plan.diag_rel(nbNodes).tobytes())
    big = c.big
    if big.nbEdges:
        h.update(np.ctypeslib.as_array(big.edge, shape=(big.nbEdges,)).tobytes())
        h.update(np.ctypeslib.as_array(big.ptr, shape=(big.nbEdges + 1,)).tobytes())
        h.update(np.ctypeslib.as_array(big.item, shape=(big.nbItems,)).tobytes())
        if c.symmetric:
            h.update(np.ctypeslib.as_array(big.edgeT, shape=(big.nbEdges,)).tobytes())
    h.update(np.array([c.maxSlots, c.maxHalo, c.maxNodes, c.maxTasks, c.maxItems, c.maxTgt, c.symmetric,
                       c.nbUnits, c.nbTasks, c.nbItems, c.nbHalo, c.nbTgt, c.nbSlotNodes, c.nbHaloNodes,
                       c.nbSlotsTotal, c.nbItemsUseful, big.nbEdges, big.nbItems], dtype=np.int64).tobytes())
    return h.hexdigest()


def cube(n, slots, sym, builder="geometric", vec=0, hub=False, tree=True):
    import common
    if builder == "geometric":
        conn, coord, N = dc.kuhn_cube_fast(n, seed=12345)
        E = conn.shape[0]
        dc.DC_set_vec_size(vec)
        dc.DC_create_tree_geometric(conn, E, 4, N, coord)
        dc.DC_permute_double_2d_array(coord, N, 3)
        dc.DC_renumber_int_array(conn.reshape(-1), 4 * E)
        row, col = dc.build_csr_fast(conn, N)
        dc.DC_finalize_tree(row, conn, E, 4)
        rec = dc.flatten_tree()
    else:
        with tempfile.TemporaryDirectory() as tmp:
            from pathlib import Path
            mesh = common.prepare_mesh(common.MineLib(vec), n, Path(tmp), hub=hub)
        conn, N, row, col = mesh["conn"], mesh["N"], mesh["row"], mesh["col"]
        rec = common.tree_records(mesh)
    plan = dc.HostPlan(conn, N, row, col, 0, rec if tree else None, slots, symmetric=sym)
    return canonical_md5(plan, N), plan.stats()


STANDARD = [
    dict(n=48, slots=512, sym=True),
    dict(n=48, slots=512, sym=False),
    dict(n=40, slots=1536, sym=True),
    dict(n=32, slots=200, sym=True),
    dict(n=20, slots=64, sym=False),
    dict(n=12, slots=832, sym=True, builder="metis", vec=4),
    dict(n=18, slots=120, sym=True, builder="metis", hub=True),
    dict(n=18, slots=120, sym=False, builder="metis", hub=True),
    dict(n=10, slots=300, sym=False, builder="metis", tree=False),
]

if __name__ == "__main__":
    if len(sys.argv) > 1:
        cases = [dict(n=int(sys.argv[1]), slots=int(sys.argv[2]) if len(sys.argv) > 2 else 512,
                      sym=bool(int(sys.argv[3])) if len(sys.argv) > 3 else True)]
    else:
        cases = STANDARD
    for case in cases:
        md5, st = cube(**case)
        print(md5, case, "units", st["units"], "items", st["items"], "big", st["big_edges"])
