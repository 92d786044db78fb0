"""User-side synthetic meshes and CSR patterns (what the reference leaves to the
application).  numpy only; sizes up to a few million elements.

kuhn_cube(n): n^3 cells, node (i,j,k) -> (i*(n+1)+j)*(n+1)+k, each cell cut into
the 6 tetrahedra that follow the 6 monotone axis-order paths from (i,j,k) to
(i+1,j+1,k+1).  Connectivity is 1-based int32 (Fortran numbering, as the reference
expects at its API, partitioning.cc:169-177).
"""
import itertools

import numpy as np


def kuhn_cube(n):
    idx = np.arange(n ** 3, dtype=np.int64)
    i, j, k = idx // (n * n), (idx // n) % n, idx % n

    def nid(a, b, c):
        return (a * (n + 1) + b) * (n + 1) + c

    tets = []
    for order in itertools.permutations(range(3)):
        cur = [i.copy(), j.copy(), k.copy()]
        nodes = [nid(*cur)]
        for axis in order:
            cur[axis] = cur[axis] + 1
            nodes.append(nid(*cur))
        tets.append(np.stack(nodes, axis=1))
    conn = np.stack(tets, axis=1).reshape(-1, 4)
    return np.ascontiguousarray(conn + 1, dtype=np.int32), (n + 1) ** 3


def node_hash(ids, comp, seed=12345):
    """31-bit LCG hash of (node id, component); identical in numpy / torch / C."""
    h = (ids.astype(np.int64) * 3 + comp + seed) & 0x7FFFFFFF
    for _ in range(3):
        h = (h * 1103515245 + 12345) & 0x7FFFFFFF
        h ^= h >> 13
    return h


def jitter_coords(n, amplitude=0.3, seed=12345):
    """Grid coordinates + deterministic jitter in [-amplitude/2, amplitude/2).
    Jitter is mandatory: on integer coordinates every summation order is exact."""
    N = (n + 1) ** 3
    ids = np.arange(N, dtype=np.int64)
    grid = np.stack([ids // ((n + 1) * (n + 1)), (ids // (n + 1)) % (n + 1), ids % (n + 1)], axis=1).astype(np.float64)
    jit = np.stack([node_hash(ids, c, seed) for c in range(3)], axis=1).astype(np.float64)
    jit = (jit / 2147483648.0 - 0.5) * amplitude
    return np.ascontiguousarray(grid + jit)


def build_csr(conn, nbNodes, colBase=0):
    """CSR pattern of the P1 operator: edge (i,j) iff some element holds both nodes;
    columns sorted ascending, diagonal included.  conn is 1-based."""
    c = conn.astype(np.int64) - 1
    keys = []
    for a in range(4):
        for b in range(4):
            keys.append(c[:, a] * nbNodes + c[:, b])
    keys = np.unique(np.concatenate(keys))
    rows = keys // nbNodes
    cols = keys - rows * nbNodes
    row = np.zeros(nbNodes + 1, dtype=np.int64)
    np.add.at(row, rows + 1, 1)
    row = np.cumsum(row)
    assert row[-1] < 2 ** 31
    return row.astype(np.int32), (cols + colBase).astype(np.int32)


def hub_collapse(conn, coord, n, block=3, stride=9):
    """High-valence variant: merge every `block`^3 group of nodes anchored on a
    `stride` lattice into one hub node, drop tetrahedra that become degenerate.
    Returns (conn 1-based int32, coord, nbNodes)."""
    N = coord.shape[0]
    ids = np.arange(N, dtype=np.int64)
    i, j, k = ids // ((n + 1) * (n + 1)), (ids // (n + 1)) % (n + 1), ids % (n + 1)

    def anchor(v):
        return (v // stride) * stride + 1

    inblock = np.ones(N, dtype=bool)
    for v in (i, j, k):
        inblock &= (v - anchor(v) >= 0) & (v - anchor(v) < block) & (anchor(v) + block <= n)
    target = ids.copy()
    hub = (anchor(i) * (n + 1) + anchor(j)) * (n + 1) + anchor(k)
    target[inblock] = hub[inblock]
    c = target[conn.astype(np.int64) - 1]
    s = np.sort(c, axis=1)
    keep = (s[:, 0] != s[:, 1]) & (s[:, 1] != s[:, 2]) & (s[:, 2] != s[:, 3])
    c = c[keep]
    used = np.unique(c)
    remap = -np.ones(N, dtype=np.int64)
    remap[used] = np.arange(used.shape[0])
    c = remap[c]
    return np.ascontiguousarray(c + 1, dtype=np.int32), np.ascontiguousarray(coord[used]), int(used.shape[0])


def scramble_mesh(conn, coord, seed=2024):
    """BASELINE config C4's irregular mesh: graded coordinates (smooth monotone warp x -> x^1.5 of the
    unit cube), a seeded random relabelling of the nodes and a shuffle of the elements, applied BEFORE
    DC_create_tree / DC_partition_mesh (the library has to find the locality again)."""
    rng = np.random.default_rng(seed)
    N, E = coord.shape[0], conn.shape[0]
    lo = coord.min(axis=0)
    span = coord.max(axis=0) - lo
    coord = np.ascontiguousarray(((coord - lo) / span) ** 1.5 * span)
    relabel = rng.permutation(N).astype(np.int32)          # old node -> new node
    new_coord = np.empty_like(coord)
    new_coord[relabel] = coord
    new_conn = relabel[conn - 1] + 1
    new_conn = np.ascontiguousarray(new_conn[rng.permutation(E)], dtype=np.int32)
    return new_conn, np.ascontiguousarray(new_coord)
