"""Host (OpenMP) vs device construction of the CSR pattern on a renumbered Kuhn cube (scratch timing)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import dc_lib_b200 as dc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
conn, coord, N = dc.kuhn_cube_fast(n); E = conn.shape[0]
dc.DC_create_tree_geometric(conn, E, 4, N, coord)
dc.DC_renumber_int_array(conn.reshape(-1), 4 * E)
t = time.time(); row, col = dc.build_csr_fast(conn, N); t_host = time.time() - t
torch.cuda.synchronize()
t = time.time(); d_conn = torch.from_numpy(conn).cuda(); torch.cuda.synchronize(); t_up = time.time() - t
dc.build_csr_device(d_conn[:1000].contiguous(), N)      # warm-up (context, module load)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); d_row, d_col = dc.build_csr_device(d_conn, N); ev1.record(); torch.cuda.synchronize()
t = time.time(); r2, c2 = d_row.cpu().numpy(), d_col.cpu().numpy(); t_down = time.time() - t
same = bool(np.array_equal(r2, row) and np.array_equal(c2, col))
print(f"n={n} E={E} N={N} nnz={row[-1]}: host {t_host:.2f}s ({os.cpu_count()} threads)  device {ev0.elapsed_time(ev1)/1e3:.3f}s "
      f"(+ upload {t_up:.2f}s, download {t_down:.2f}s)  identical={same}")
