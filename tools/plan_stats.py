"""tools/plan_stats.py -- schedule statistics of the n^3 jittered Kuhn cube (coordinate-bisection tree,
symmetric schedule) for a list of leaf capacities: units, rows and slots per unit, element setups per
element, item (padding) efficiency, tasks and diagonal tasks per unit, gradient reads per element
(ideal 16) and the per-unit maxima the kernel sizes its shared memory from.  CPU only; source of the
tables in profiles/README.md "Round-1f".     python tools/plan_stats.py <n> <unit_slots> <capacity> [<capacity> ...]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dc_lib_b200 as dc
n = int(sys.argv[1]); slots = int(sys.argv[2])
for mep in [int(x) for x in sys.argv[3:]]:
    conn, coord, N = dc.kuhn_cube_fast(n, seed=12345)
    E = conn.shape[0]
    dc.DC_set_vec_size(0)
    dc.DC_set_max_elem_per_part(mep)
    dc.DC_create_tree_geometric(conn, E, 4, N, coord)
    dc.DC_permute_double_2d_array(coord, N, 3)
    dc.DC_renumber_int_array(conn.reshape(-1), 4 * E)
    row, col = dc.build_csr_fast(conn, N)
    dc.DC_finalize_tree(row, conn, E, 4)
    rec = dc.flatten_tree()
    plan = dc.HostPlan(conn, N, row, col, 0, rec, slots, symmetric=True)
    c = plan.c
    a = plan.arrays()
    u32 = a["units"].view(np.int32).reshape(-1, 24)
    nodes = u32[:, 14]; nslots = u32[:, 16] + u32[:, 17]; tasks = u32[:, 19]; items = u32[:, 20]
    T = a["tasks"].view(np.uint16).reshape(-1, 8)
    iters = T[:, 2].astype(np.int64); kind = T[:, 3]
    # read rounds per element: upper 2 sides, diag 1 side  (x3 components)
    rr = (iters[kind == 2].sum() * 2 + iters[kind == 3].sum()) * 32 / E
    print(f"mep {mep}: units {c.nbUnits} nodes/unit {nodes.mean():.1f} slots/unit {nslots.mean():.0f} (max {nslots.max()}) setups/elem {c.nbSlotsTotal/E:.3f} item_eff {c.nbItemsUseful/c.nbItems:.3f} "
          f"tasks/unit {tasks.mean():.2f} diag tasks/unit {(kind==3).sum()/c.nbUnits:.2f} gradient-reads/elem {rr:.2f} maxTasks {c.maxTasks} maxItems {c.maxItems} maxTgt {c.maxTgt} maxHalo {c.maxHalo} maxNodes {c.maxNodes}")
