"""Multi-GPU decomposition of the assembly: one process per GPU, each owns a
slab of the mesh with its own D&C tree and device schedule; the only exchange per assembly is
the sum of the partial values of the CSR edges whose two nodes lie on a plane shared with a
neighbouring slab.  The per-neighbour edge lists play the role of the reference's
intfIndex / intfNodes / intfDst triple (include/DC.h:51-53, 141-143), at edge granularity, and
are ordered identically on both sides of a plane so that buffers can be added position by
position.  Transport is torch.distributed send/recv (NCCL over NVLink on GPUs, gloo in the
CPU tests); packing and the unpack-add are the kernels behind dc_cuda_pack_edges /
dc_cuda_unpack_add_edges.
"""
import numpy as np

from . import (DC_create_tree, DC_create_tree_device, DC_create_tree_geometric, DC_create_tree_geometric_device, DC_device_select, DC_finalize_tree, DC_partition_mesh, DC_permute_double_2d_array,
               DC_renumber_int_array, DC_set_num_threads, DC_set_vec_size, build_csr_fast, flatten_tree, get_node_perm,
               kuhn_cube_fast, pack_edges, unpack_add_edges)


def global_node_ids(n, rank):
    """Global id of every local (original-numbering) node of slab `rank`."""
    m = n + 1
    ids = np.arange(m ** 3, dtype=np.int64)
    return ids + np.int64(rank) * n * m * m


def plane_edges(node_perm, row, col, n, side):
    """Local CSR edges whose two nodes lie on the plane x = side (0 or n) of the cube, ordered by
    (plane index of the row node, plane index of the column node) -- the same on both slabs."""
    m = n + 1
    N = node_perm.shape[0]
    plane_old = side * m * m + np.arange(m * m, dtype=np.int64)          # plane index p = j*(n+1)+k
    plane_new = node_perm[plane_old].astype(np.int64)
    plane_of = -np.ones(N, dtype=np.int64)
    plane_of[plane_new] = np.arange(m * m)
    starts, lens = row[plane_new].astype(np.int64), (row[plane_new + 1] - row[plane_new]).astype(np.int64)
    owner = np.repeat(np.arange(m * m), lens)
    idx = np.repeat(starts - np.concatenate(([0], np.cumsum(lens)[:-1])), lens) + np.arange(lens.sum())
    other = plane_of[col[idx]]
    keep = other >= 0
    key = owner[keep] * (m * m) + other[keep]
    return np.ascontiguousarray(idx[keep][np.argsort(key, kind="stable")], dtype=np.int64)


def _create_tree(conn, E, N, coord, builder, device):
    """DC_create_tree / DC_create_tree_geometric; with a CUDA device the element passes of the tree
    proper run there (DC_create_tree_device, include/DC_device.h) -- same bytes either way."""
    if device is not None:
        import torch
        DC_device_select(torch.device(device).index or 0)
    if builder == "geometric":
        (DC_create_tree_geometric_device if device is not None else DC_create_tree_geometric)(conn, E, 4, N, coord)
    else:
        (DC_create_tree_device if device is not None else DC_create_tree)(conn, E, 4, N)


def build_slab(n, rank, world, builder="geometric", vec=0, seed=12345, csr_device=None):
    """Appendix-A driver for slab `rank` of a bar of `world` jittered Kuhn cubes stacked along x.
    Returns the user-side arrays, the flattened tree and, per neighbour, the local CSR edges
    shared with it (both nodes on the common plane) in an order both sides agree on."""
    conn, coord, N = kuhn_cube_fast(n, seed=seed, x_offset=rank * n)
    E = conn.shape[0]
    DC_set_vec_size(vec)
    _create_tree(conn, E, N, coord, builder, csr_device)
    node_perm = get_node_perm(N)
    DC_permute_double_2d_array(coord, N, 3)
    DC_renumber_int_array(conn.reshape(-1), 4 * E)
    import time
    t_csr = time.time()
    if csr_device is not None:
        # CSR pattern on the GPU (dc_cuda_csr_*): connectivity up, row pointers and columns back
        # (the host needs them for DC_finalize_tree and the schedule builder)
        import torch
        from . import build_csr_device
        d_row, d_col = build_csr_device(torch.from_numpy(conn).to(csr_device), N)
        row, col = d_row.cpu().numpy(), d_col.cpu().numpy()
        del d_row, d_col
    else:
        row, col = build_csr_fast(conn, N)
    t_csr = time.time() - t_csr
    DC_finalize_tree(row, conn, E, 4)
    rec = flatten_tree()

    intf = {}
    for side, nb in ((0, rank - 1), (n, rank + 1)):
        if 0 <= nb < world:
            intf[nb] = plane_edges(node_perm, row, col, n, side)
    gid_of_new = np.empty(N, dtype=np.int64)
    gid_of_new[node_perm] = global_node_ids(n, rank)
    return dict(conn=conn, coord=coord, row=row, col=col, E=E, N=N, nnz=int(row[-1]), rec=rec, n=n,
                intf=intf, gid_of_new=gid_of_new, node_perm=node_perm, rank=rank, world=world, csr_s=t_csr)


def _pair_keys(conn_g, n_glob, keep):
    """Sorted unique keys row*n_glob+col of the CSR entries (all 16 node pairs of every element,
    diagonal included) whose two GLOBAL nodes both satisfy keep[node]."""
    c = conn_g.astype(np.int64) - 1
    c = c[keep[c].sum(axis=1) >= 1]
    a = np.repeat(c, 4, axis=1).reshape(-1)
    b = np.tile(c, (1, 4)).reshape(-1)
    ok = keep[a] & keep[b]
    return np.unique(a[ok] * np.int64(n_glob) + b[ok])


def build_part(conn_g, coord_g, elem_part, rank, world, builder="metis", vec=0, csr_device=None):
    """Appendix-A driver for part `rank` of a GENERAL element partition (DC_partition_mesh) of one
    global mesh (BASELINE config 4: an irregular mesh split over the GPUs): local renumbering, own
    D&C tree, CSR pattern of the local elements, and per neighbouring rank the local CSR edges
    that the neighbour holds too -- listed by ascending (global row, global column), so both sides
    agree on the order.  Every rank holds the global connectivity here (host memory only); the
    device only ever sees the local part.  All passes are O(global elements) array operations
    (no per-rank sort of the whole mesh).  Returns the same dictionary as build_slab."""
    import time
    n_glob = coord_g.shape[0]
    mine = np.flatnonzero(elem_part == rank)
    conn_mine_g = conn_g[mine]
    mine_nodes = np.zeros(n_glob, dtype=bool)
    mine_nodes[conn_mine_g.reshape(-1) - 1] = True
    nodes = np.flatnonzero(mine_nodes)                                   # global ids of the local nodes, ascending
    local_of = np.zeros(n_glob, dtype=np.int32)
    local_of[nodes] = np.arange(1, nodes.shape[0] + 1, dtype=np.int32)
    conn = np.ascontiguousarray(local_of[conn_mine_g - 1], dtype=np.int32)
    del local_of
    coord = np.ascontiguousarray(coord_g[nodes])
    E, N = conn.shape[0], nodes.shape[0]
    DC_set_vec_size(vec)
    _create_tree(conn, E, N, coord, builder, csr_device)
    node_perm = get_node_perm(N)
    DC_permute_double_2d_array(coord, N, 3)
    DC_renumber_int_array(conn.reshape(-1), 4 * E)
    t_csr = time.time()
    if csr_device is not None:
        import torch
        from . import build_csr_device
        d_row, d_col = build_csr_device(torch.from_numpy(conn).to(csr_device), N)
        row, col = d_row.cpu().numpy(), d_col.cpu().numpy()
        del d_row, d_col
    else:
        row, col = build_csr_fast(conn, N)
    t_csr = time.time() - t_csr
    DC_finalize_tree(row, conn, E, 4)
    rec = flatten_tree()
    gid_of_new = np.empty(N, dtype=np.int64)
    gid_of_new[node_perm] = nodes
    intf = {}
    rows = None
    for q in range(world):
        if q == rank:
            continue
        theirs = conn_g[elem_part == q]
        if theirs.shape[0] == 0:
            continue
        shared = np.zeros(n_glob, dtype=bool)
        shared[theirs.reshape(-1) - 1] = True
        shared &= mine_nodes
        if not shared.any():
            continue
        # local CSR entries whose two nodes are shared with q, keyed by global (row, col) ...
        if rows is None:
            rows = np.repeat(np.arange(N, dtype=np.int64), np.diff(row))
        cand = np.flatnonzero(shared[gid_of_new[rows]] & shared[gid_of_new[col]])
        key_cand = gid_of_new[rows[cand]] * np.int64(n_glob) + gid_of_new[col[cand]]
        order = np.argsort(key_cand, kind="stable")
        # ... of which q holds those that one of ITS elements produces
        theirs_keys = _pair_keys(theirs, n_glob, shared)
        pos = np.searchsorted(key_cand[order], theirs_keys)
        pos[pos >= order.shape[0]] = 0
        hit = key_cand[order][pos] == theirs_keys if order.shape[0] else np.zeros(0, dtype=bool)
        if not hit.any():
            continue
        intf[q] = np.ascontiguousarray(cand[order[pos[hit]]], dtype=np.int64)       # ascending global key
    return dict(conn=conn, coord=coord, row=row, col=col, E=E, N=N, nnz=int(row[-1]), rec=rec, intf=intf,
                gid_of_new=gid_of_new, node_perm=node_perm, rank=rank, world=world, elems=mine, csr_s=t_csr)


def setup_device_parts(n, rank, world, device, slots=512, dist=None, seed=2024, partition="metis", bank_model=0):
    """BASELINE config 4 on the device: ONE graded / relabelled / shuffled Kuhn cube n^3 (scramble_mesh), cut into
    `world` parts by DC_partition_mesh on rank 0 (METIS recursive bisection of the nodal graph, the call
    DC_create_tree makes, partitioning.cc:167-228; "geometric" = coordinate bisection) and broadcast; every
    rank extracts its part (build_part), builds its own D&C tree (coordinate bisection), CSR pattern (on the
    GPU) and schedule with the interface units first, and uploads them.  Returns (ctx, info) like
    setup_device_slabs; info["intf"] holds the per-neighbour edge lists."""
    import time

    import torch

    from . import Context, HostPlan, kuhn_cube_fast, scramble_mesh
    t0 = time.time()
    DC_set_num_threads(0)
    conn_g, coord_g, n_glob = kuhn_cube_fast(n)
    conn_g, coord_g = scramble_mesh(conn_g, coord_g, seed=seed)
    t_mesh = time.time() - t0
    t1 = time.time()
    if world > 1:
        part_t = torch.empty(conn_g.shape[0], dtype=torch.int32, device=device)
        if rank == 0:
            # DC_PARTITION_CACHE=<dir>: the element partition (a host-only step: METIS on the nodal graph)
            # is kept there per (mesh, parts), so that a run on GPUs does not leave them idle behind it
            import os
            cache = os.environ.get("DC_PARTITION_CACHE")
            path = os.path.join(cache, f"c4_part_n{n}_w{world}_{partition}_s{seed}.npy") if cache else None
            if path and os.path.exists(path):
                part = np.load(path).astype(np.int32)
                assert part.shape[0] == conn_g.shape[0]
            else:
                part = DC_partition_mesh(conn_g, n_glob, world, coord_g if partition == "geometric" else None)
                if path:
                    os.makedirs(cache, exist_ok=True)
                    np.save(path, part.astype(np.int8))
            part_t.copy_(torch.from_numpy(part))
        dist.broadcast(part_t, src=0)
        part = part_t.cpu().numpy()
        del part_t
    else:
        part = np.zeros(conn_g.shape[0], dtype=np.int32)
    t_part = time.time() - t1
    t1 = time.time()
    p = build_part(conn_g, coord_g, part, rank, world, builder="geometric", csr_device=device)
    E_glob = conn_g.shape[0]
    del conn_g, coord_g, part
    t_tree = time.time() - t1
    t1 = time.time()
    plan = HostPlan(p["conn"], p["N"], p["row"], p["col"], 0, p["rec"], slots, bank_model=bank_model)
    nb_front = 0
    if p["intf"]:
        flag = np.zeros(p["N"], dtype=np.uint8)
        for e in p["intf"].values():
            flag[np.searchsorted(p["row"], e, side="right") - 1] = 1
        nb_front = plan.order_units(flag)
    t_plan = time.time() - t1
    ctx = Context(torch.device(device).index or 0)
    ctx.set_mesh(p["conn"], p["coord"], p["row"], p["col"], 0)
    ctx.upload_plan(plan)
    ctx.set_front_units(nb_front)
    st = plan.stats()
    info = dict(E=p["E"], N=p["N"], nnz=p["nnz"], E_global=E_glob, stats=st, tree_s=t_tree, plan_s=t_plan, csr_s=p["csr_s"],
                mesh_s=t_mesh, partition_s=t_part, nb_front=nb_front, intf=p["intf"], row=p["row"],
                coord=torch.from_numpy(p["coord"]), setup_s=time.time() - t0,
                interface_edges=int(sum(e.shape[0] for e in p["intf"].values())))
    return ctx, info


class InterfaceExchange:
    """Sums the interface-edge partial values with the neighbouring ranks."""

    def __init__(self, intf, per_edge, device, dist=None):
        import torch
        self.torch, self.dist, self.per = torch, dist, per_edge
        self.on_gpu = torch.device(device).type == "cuda"
        self.nbs = sorted(intf)
        self.edges = {nb: torch.from_numpy(intf[nb]).to(device) for nb in self.nbs}
        self.send = {nb: torch.empty(intf[nb].shape[0] * per_edge, dtype=torch.float64, device=device) for nb in self.nbs}
        self.recv = {nb: torch.empty_like(self.send[nb]) for nb in self.nbs}
        self.bytes_per_step = sum(2 * 8 * self.send[nb].numel() for nb in self.nbs)
        self.launches_per_step = 2 * len(self.nbs) if self.on_gpu else 0

    def _pack(self, values, nb, stream):
        if self.on_gpu:
            pack_edges(values, self.edges[nb], self.per, self.send[nb], stream)
        else:
            self.send[nb].copy_(values.view(-1, self.per)[self.edges[nb]].reshape(-1))

    def _unpack_add(self, values, nb, stream):
        if self.on_gpu:
            unpack_add_edges(values, self.edges[nb], self.per, self.recv[nb], stream)
        else:
            values.view(-1, self.per)[self.edges[nb]] += self.recv[nb].view(-1, self.per)

    def run(self, values, stream=None):
        if not self.nbs:
            return
        dist = self.dist
        if self.on_gpu and stream is None:
            # NCCL orders its send / recv against torch's CURRENT stream: pack and unpack-add must run there too
            stream = self.torch.cuda.current_stream().cuda_stream
        ops = []
        for nb in self.nbs:
            self._pack(values, nb, stream)
            ops.append(dist.P2POp(dist.isend, self.send[nb], nb))
            ops.append(dist.P2POp(dist.irecv, self.recv[nb], nb))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for nb in self.nbs:
            self._unpack_add(values, nb, stream)


def setup_device_slabs(n, rank, world, device, builder="geometric", vec=0, slots=400, dist=None, seed=12345,
                       symmetric=True, bank_model=0):
    """Device-resident slab for this rank.  Every slab of the bar has the same topology, so
    rank 0 builds tree, CSR pattern and schedule once on its host cores and broadcasts the
    arrays GPU-to-GPU (NCCL over NVLink); each rank generates its own coordinates and permutes
    them on the device.  Returns (ctx, info) with the context ready for dc_cuda_assemble."""
    import time

    import torch

    from . import Context, HostPlan, gather_double_2d

    names = ["conn", "row", "col", "node_perm", "units", "tasks", "items", "slotNodes", "haloNodes", "diagRel", "tgt",
             "bigEdge", "bigPtr", "bigItem", "bigEdgeT", "intf_left", "intf_right"]
    t0 = time.time()
    tensors, meta = {}, None
    if rank == 0:
        DC_set_num_threads(0)                      # all host cores (torchrun exports OMP_NUM_THREADS=1)
        slab = build_slab(n, 0, 1, builder=builder, vec=vec, seed=seed, csr_device=device)
        node_perm = slab["node_perm"]
        t_tree = time.time() - t0
        t1 = time.time()
        plan = HostPlan(slab["conn"], slab["N"], slab["row"], slab["col"], 0, slab["rec"], slots, symmetric=symmetric,
                        bank_model=bank_model)
        intf_l = plane_edges(node_perm, slab["row"], slab["col"], n, 0) if world > 1 else np.zeros(0, np.int64)
        intf_r = plane_edges(node_perm, slab["row"], slab["col"], n, n) if world > 1 else np.zeros(0, np.int64)
        nb_front = 0
        if world > 1:
            # units that own a row of either interface plane go first: their partial sums travel while the
            # other units are assembled (every slab has the same topology, so both planes are flagged)
            flag = np.zeros(slab["N"], dtype=np.uint8)
            for e in (intf_l, intf_r):
                flag[np.searchsorted(slab["row"], e, side="right") - 1] = 1
            nb_front = plan.order_units(flag)
        t_plan = time.time() - t1
        arr = plan.arrays()
        host = dict(conn=slab["conn"], row=slab["row"], col=slab["col"], node_perm=node_perm, units=arr["units"],
                    tasks=arr["tasks"], items=arr["items"].view(np.int16), slotNodes=arr["slotNodes"].view(np.int16),
                    haloNodes=arr["haloNodes"],
                    diagRel=plan.diag_rel(slab["N"]), tgt=arr["tgt"],
                    bigEdge=arr["bigEdge"], bigPtr=arr["bigPtr"], bigItem=arr["bigItem"].view(np.int32), bigEdgeT=arr["bigEdgeT"],
                    intf_left=intf_l, intf_right=intf_r)
        for k in names:
            tensors[k] = torch.from_numpy(np.ascontiguousarray(host[k])).to(device)
        st = plan.stats()
        meta = dict(shapes={k: (tuple(tensors[k].shape), str(tensors[k].dtype)) for k in names},
                    ints={k: arr[k] for k in ("nbUnits", "maxSlots", "maxHalo", "maxNodes", "maxTasks", "maxItems", "maxTgt", "symmetric", "bigEdges")},
                    E=slab["E"], N=slab["N"], nnz=slab["nnz"], stats=st, tree_s=t_tree, plan_s=t_plan, nb_front=nb_front,
                    csr_s=slab["csr_s"])
        del plan, slab, host, arr
    if world > 1:
        box = [meta]
        dist.broadcast_object_list(box, src=0)
        meta = box[0]
        for k in names:
            shape, dtype = meta["shapes"][k]
            if rank != 0:
                tensors[k] = torch.empty(shape, dtype=getattr(torch, dtype.split(".")[1]), device=device)
            if tensors[k].numel():
                dist.broadcast(tensors[k].view(torch.uint8), src=0)       # raw bytes: NCCL has no int16
    # own coordinates: generated in the original numbering, permuted on the device
    _, coord_old, _ = kuhn_cube_fast(n, seed=seed, x_offset=rank * n, coords_only=True)
    d_old = torch.from_numpy(coord_old).to(device)
    d_coord = torch.zeros((d_old.shape[0] + 2, 3), dtype=torch.float64, device=device)[:d_old.shape[0]]   # 2 spare nodes
    # told by destination (new position -> old position): whole sectors are written
    inv = torch.empty_like(tensors["node_perm"])
    inv[tensors["node_perm"].long()] = torch.arange(inv.numel(), dtype=inv.dtype, device=inv.device)
    gather_double_2d(d_old, d_coord, inv, 3, torch.cuda.current_stream().cuda_stream)
    del inv
    torch.cuda.synchronize()
    del d_old, coord_old
    ctx = Context(torch.device(device).index or 0)
    ctx.set_mesh_device(tensors["conn"], d_coord, tensors["row"], tensors["col"], meta["nnz"], 0)
    plan_t = dict(units=tensors["units"], tasks=tensors["tasks"], items=tensors["items"],
                  slotNodes=tensors["slotNodes"], haloNodes=tensors["haloNodes"], diagRel=tensors["diagRel"],
                  tgt=tensors["tgt"], bigEdge=tensors["bigEdge"], bigPtr=tensors["bigPtr"], bigItem=tensors["bigItem"],
                  bigEdgeT=tensors["bigEdgeT"], **meta["ints"])
    ctx.set_plan_device(plan_t)
    ctx.set_front_units(meta.get("nb_front", 0))
    intf = {}
    if rank > 0:
        intf[rank - 1] = tensors["intf_left"]
    if rank < world - 1:
        intf[rank + 1] = tensors["intf_right"]
    info = dict(meta, coord=d_coord, intf=intf, setup_s=time.time() - t0)
    return ctx, info


def device_exchange(intf, per_edge, device, dist, rank, world):
    """dc_cuda_exchange over the per-neighbour edge lists (numpy or CUDA tensors): peer-memory
    transport inside libdc_b200.so; torch.distributed only carries the 128-byte handles once."""
    import torch

    from . import DeviceExchange
    x = DeviceExchange(intf, per_edge, torch.device(device).index or 0)
    x.connect(dist, rank, world)
    return x
