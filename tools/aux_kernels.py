"""Timing of the auxiliary kernels of the path at BASELINE config-3 sizes (256^3 cube: 16 974 593
nodes, 100 663 296 tets, 253 036 801 CSR entries): coordinate / connectivity permutation, renumbering,
CSR reset, SpMV on the assembled matrix.  Bytes are the algorithmic ones of SURVEY.md 8d
(permute: 2 x array + 4 B per item of the permutation; reset: the bytes written)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dc_lib_b200 as dc

N, E, nnz = 16974593, 100663296, 253036801
dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(1)
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def report(name, ms, nbytes):
    print(f"{name:42s} {ms:8.3f} ms  {nbytes/ms/1e6:8.1f} GB/s  {nbytes/ms/1e6/6550.4:.2f} of the measured copy bandwidth", flush=True)
# a tree-like permutation: local shuffles inside blocks of 4096 items, blocks in random order
def perm_like_tree(n):
    blocks = (n + 4095) // 4096
    border = torch.randperm(blocks, device=dev, generator=g)
    inner = torch.argsort(torch.rand(blocks, 4096, device=dev, generator=g), dim=1)
    p = (border[:, None] * 4096 + inner).reshape(-1)
    p = p[p < n]
    # compress to a permutation of 0..n-1 keeping the order
    return torch.argsort(torch.argsort(p)).to(torch.int32)
nperm = perm_like_tree(N); eperm = perm_like_tree(E)
coord = torch.rand(N, 3, dtype=torch.float64, device=dev, generator=g); out = torch.empty_like(coord)
report("permute_rows<double> coord [N][3]", timed(lambda: dc.permute_double_2d(coord, out, nperm, 3)), 2 * coord.numel() * 8 + 4 * N)
conn = torch.randint(1, N + 1, (E, 4), dtype=torch.int32, device=dev, generator=g); outc = torch.empty_like(conn)
report("permute_rows_int4 elemToNode [E][4]", timed(lambda: dc.permute_int_2d(conn, outc, eperm, 4)), 2 * conn.numel() * 4 + 4 * E)
ninv = torch.empty_like(nperm); ninv[nperm.long()] = torch.arange(N, dtype=torch.int32, device=dev)
einv = torch.empty_like(eperm); einv[eperm.long()] = torch.arange(E, dtype=torch.int32, device=dev)
report("gather_rows<double> coord [N][3] (by destination)", timed(lambda: dc.gather_double_2d(coord, out, ninv, 3)), 2 * coord.numel() * 8 + 4 * N)
report("gather_rows_int4 elemToNode [E][4] (by destination)", timed(lambda: dc.gather_int_2d(conn, outc, einv, 4)), 2 * conn.numel() * 4 + 4 * E)
report("renumber elemToNode (4E entries), random node ids", timed(lambda: dc.renumber(outc, nperm, True)), 2 * conn.numel() * 4)
# node ids as a tree-ordered mesh has them: the nodes of an element lie within a few thousand ids of each other
base_id = (torch.arange(E, device=dev) * (N / E)).long()
local = (base_id[:, None] + torch.randint(-2000, 2000, (E, 4), device=dev, generator=g)).clamp_(0, N - 1).to(torch.int32) + 1
report("renumber elemToNode (4E entries), local node ids", timed(lambda: dc.renumber(local, nperm, True)), 2 * local.numel() * 4)
del local, base_id
# the connectivity of a real tree-ordered mesh (160^3 cube after DC_create_tree_geometric / permute / renumber):
# neighbouring elements share nodes, so the lanes of a load share sectors
import bench
dc.DC_set_max_elem_per_part(bench.leaf_capacity(160, 62))
w = bench.build_workload(dc, 160, 0, "geometric")
real = torch.from_numpy(w["conn"]).to(dev)
rperm = perm_like_tree(w["N"])
report("renumber elemToNode (4E entries), 160^3 tree-ordered mesh", timed(lambda: dc.renumber(real, rperm, True)), 2 * real.numel() * 4)
del real, rperm, w
del conn, outc, coord, out
for dim in (1, 3):
    vals = torch.empty(nnz * dim * dim, dtype=torch.float64, device=dev)
    report(f"reset_values d={dim} (whole CSR)", timed(lambda: dc.reset_edges(vals, 0, nnz - 1, dim * dim)), vals.numel() * 8)
    del vals
