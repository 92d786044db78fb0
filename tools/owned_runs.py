"""tools/owned_runs.py -- how contiguous are the CSR blocks a work unit owns?  (CPU only)

Evidence for DESIGN.md section 9 (b): per row of a unit, the blocks with col >= the unit's first node
(in-unit lower, diagonal, upper) form one contiguous run; only the blocks with col < nodeFirst are
written by other units.  Prints, per element, the 128-byte lines touched when every 72-byte block is
stored on its own (what gather_units does today; equals the measured global-store wavefronts) and
when the owned runs are stored contiguously.   python tools/owned_runs.py [n]
"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import dc_lib_b200 as dc, bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
conn, coord, N = dc.kuhn_cube_fast(n, seed=12345)
E = conn.shape[0]
dc.DC_set_max_elem_per_part(bench.leaf_capacity(n, 62))
dc.DC_create_tree_geometric(conn, E, 4, N, coord)
dc.DC_renumber_int_array(conn.reshape(-1), 4 * E)
row, col = dc.build_csr_fast(conn, N)
dc.DC_finalize_tree(row, conn, E, 4)
rec = dc.flatten_tree()
plan = dc.HostPlan(conn, N, row, col, 0, rec, 512, symmetric=True)
a = plan.arrays()
u32 = a["units"].view(np.int32).reshape(-1, 24)
nf, nc = u32[:, 13].astype(np.int64), u32[:, 14].astype(np.int64)
unit_of = np.repeat(np.arange(len(nf)), nc)           # units sorted by nodeFirst and cover all nodes
assert unit_of.shape[0] == N
rows_of = np.repeat(np.arange(N), np.diff(row))
c = col.astype(np.int64)
first = nf[unit_of[rows_of]]; last = first + nc[unit_of[rows_of]] - 1
before = c < first; inside = (c >= first) & (c <= last); after = c > last
print("edges per row", len(c) / N, "col<nodeFirst", before.sum() / N, "in unit", inside.sum() / N, "col>nodeLast", after.sum() / N)
owned_bytes = (inside.sum() + after.sum()) * 72 / N
print("owned run per row: blocks", (inside.sum() + after.sum()) / N, "bytes", owned_bytes)
# lines touched if each block is stored alone vs runs stored contiguously
k = np.arange(len(c), dtype=np.int64)
lo, hi = k * 72 // 128, (k * 72 + 71) // 128
print("lines per block stored alone", (hi - lo + 1).mean())
own = ~before
start = np.r_[True, (np.diff(rows_of) != 0) | (~own[:-1])] & own
run_id = np.cumsum(start) - 1
rs = np.zeros(run_id.max() + 1, dtype=np.int64); re = np.zeros_like(rs)
np.minimum.at(rs := np.full(run_id.max() + 1, 1 << 62), run_id[own], k[own] * 72)
np.maximum.at(re, run_id[own], k[own] * 72 + 71)
lines_runs = (re // 128 - rs // 128 + 1).sum()
lines_out = (hi - lo + 1)[before].sum()
print("per element: lines today", (hi - lo + 1).sum() / E, "owned runs", lines_runs / E, "+ out-of-unit transposes", lines_out / E)
