"""Host-side setup timing at scale (CPU only): mesh, geometric tree, CSR, plan."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dc_lib_b200 as dc
n = int(sys.argv[1]); slots = int(sys.argv[2]) if len(sys.argv) > 2 else 288
T = time.time
t = T(); conn, coord, N = dc.kuhn_cube_fast(n); E = conn.shape[0]; print(f"mesh {T()-t:.1f}s E={E} N={N}", flush=True)
t = T(); dc.DC_create_tree_geometric(conn, E, 4, N, coord); print(f"tree {T()-t:.1f}s", flush=True)
t = T(); dc.DC_permute_double_2d_array(coord, N, 3); dc.DC_renumber_int_array(conn.reshape(-1), 4 * E); print(f"perm {T()-t:.1f}s", flush=True)
t = T(); row, col = dc.build_csr_fast(conn, N); print(f"csr {T()-t:.1f}s nnz={row[-1]}", flush=True)
t = T(); dc.DC_finalize_tree(row, conn, E, 4); rec = dc.flatten_tree(); print(f"finalize+flatten {T()-t:.1f}s treeNodes={rec.shape[0]}", flush=True)
t = T(); plan = dc.HostPlan(conn, N, row, col, 0, rec, slots); print(f"plan {T()-t:.1f}s {plan.stats()}", flush=True)
st = plan.stats(); print(f"pad_eff={st['items_useful']/st['items']:.3f} slots/elem={st['slots_total']/E:.3f}")
