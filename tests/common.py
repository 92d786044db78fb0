"""Shared test plumbing: loaders for the CHECKERS (oracle restatement and, when it was
built, the unmodified reference in oracle/_ref) and the driver sequence a DC-lib user
runs around the library.  The product is only reached through
dc_lib_b200 (ctypes over libdc_b200.so)."""
import ctypes
import hashlib
import os
from ctypes import POINTER, byref, c_double, c_int, c_int64, c_void_p

import numpy as np

import dc_lib_b200 as dc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
I = c_int


def _ip(v):
    return byref(c_int(int(v)))


# ------------------------------------------------------------------ oracle restatement

class DcoUser(ctypes.Structure):
    _fields_ = [("elemToNode", c_void_p), ("coord", c_void_p), ("row", c_void_p), ("col", c_void_p),
                ("values", c_void_p), ("colBase", c_int), ("opDim", c_int), ("lam", c_double),
                ("mu", c_double), ("resetInSeq", c_int), ("opKind", c_int), ("w", c_double * 3)]


class DcoNode(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("firstElem", "lastElem", "lastSep", "firstNode", "lastNode", "firstEdge",
                                     "lastEdge", "vecOffset", "isSep", "isLeaf", "left", "right", "sep")]


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        L = ctypes.CDLL(os.path.join(ROOT, "oracle", "libdc_oracle.so"), mode=os.RTLD_LOCAL)
        L.dco_traverse.argtypes = [c_void_p, c_int, POINTER(DcoUser), c_int]
        L.dco_assemble_serial.argtypes = [POINTER(DcoUser), c_int, c_int64]
        L.dco_parse_tree.argtypes = [c_void_p, c_int64, c_int, c_int, POINTER(c_int), c_void_p, c_void_p, c_void_p, c_int64]
        L.dco_parse_tree.restype = c_int64
        L.dco_finalize.argtypes = [c_void_p, c_int64, c_void_p]
        L.dco_permute_double_2d.argtypes = [c_void_p, c_void_p, c_int, c_int]
        L.dco_permute_int_2d.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int]
        L.dco_permute_int_1d.argtypes = [c_void_p, c_void_p, c_int]
        L.dco_renumber.argtypes = [c_void_p, c_void_p, c_int, c_int]
        L.dco_create_permutation.argtypes = [c_void_p, c_void_p, c_int, c_int]
        _oracle = L
    return _oracle


LAMBDA, MU = 0.5769230769230769, 0.3846153846153846     # E = 1, nu = 0.3


def make_user(mesh, dim, values, reset_in_seq=1, op_kind=0, w=(0.0, 0.0, 0.0)):
    u = DcoUser()
    u.opKind = op_kind
    u.w[0], u.w[1], u.w[2] = w
    u.elemToNode = mesh["conn"].ctypes.data
    u.coord = mesh["coord"].ctypes.data
    u.row = mesh["row"].ctypes.data
    u.col = mesh["col"].ctypes.data
    u.values = values.ctypes.data
    u.colBase = 0
    u.opDim = dim
    u.lam, u.mu = LAMBDA, MU
    u.resetInSeq = reset_in_seq
    return u


def oracle_serial(mesh, dim, op_kind=0, w=(0.0, 0.0, 0.0)):
    """Secondary oracle: zero everything, then elements in ascending (tree) order."""
    values = np.full(mesh["nnz"] * dim * dim, np.nan)
    u = make_user(mesh, dim, values, op_kind=op_kind, w=w)
    oracle().dco_assemble_serial(byref(u), mesh["E"], mesh["nnz"])
    return values


def parse_tree(buf, E, N):
    buf = np.frombuffer(buf, dtype=np.uint8)
    L = oracle()
    n = L.dco_parse_tree(buf.ctypes.data, buf.size, E, N, None, None, None, None, 0)
    assert n > 0, "tree file did not parse"
    nodes = (DcoNode * n)()
    elemPerm = np.zeros(E, dtype=np.int32)
    nodePerm = np.zeros(N, dtype=np.int32)
    comm = c_int(0)
    L.dco_parse_tree(buf.ctypes.data, buf.size, E, N, byref(comm), elemPerm.ctypes.data, nodePerm.ctypes.data,
                     ctypes.cast(nodes, c_void_p), n)
    return nodes, n, elemPerm, nodePerm


def oracle_traverse(mesh, dim, nodes, split):
    values = np.full(mesh["nnz"] * dim * dim, np.nan)
    u = make_user(mesh, dim, values)
    oracle().dco_traverse(ctypes.cast(nodes, c_void_p), 0, byref(u), 1 if split else 0)
    return values


# ------------------------------------------------------------------ the reference itself

def ref_path(vec=0, runtime="omp"):
    name = f"libDC_ref_{runtime}" + (f"_vec{vec}" if vec else "") + ".so"
    return os.path.join(ROOT, "oracle", "_ref", name)


def have_ref(vec=0, runtime="omp"):
    return os.path.exists(ref_path(vec, runtime))


class RefLib:
    """The UNMODIFIED reference (oracle/_ref), driven through its Fortran ABI
    (src/fortran_wrapper.cc:64-146)."""

    def __init__(self, vec=0, runtime="omp"):
        self.L = ctypes.CDLL(ref_path(vec, runtime), mode=os.RTLD_LOCAL)
        self.vec = vec

    def create_tree(self, conn, E, N):
        self.L.dc_create_tree_(c_void_p(conn.ctypes.data), _ip(E), _ip(4), _ip(N))

    def permute_double_2d(self, tab, n, dim):
        self.L.dc_permute_double_2d_array_(c_void_p(tab.ctypes.data), _ip(n), _ip(dim))

    def permute_int_2d(self, tab, n, dim, offset=0):
        self.L.dc_permute_int_2d_array_(c_void_p(tab.ctypes.data), _ip(n), _ip(dim), _ip(offset))

    def permute_int_1d(self, tab, n):
        self.L.dc_permute_int_1d_array_(c_void_p(tab.ctypes.data), _ip(n))

    def renumber(self, tab, n):
        self.L.dc_renumber_int_array_(c_void_p(tab.ctypes.data), _ip(n))

    def create_permutation(self, part, nbPart):
        perm = np.full(part.shape[0], -1, dtype=np.int32)
        self.L.dc_create_permutation_(c_void_p(perm.ctypes.data), c_void_p(part.ctypes.data), _ip(part.shape[0]), _ip(nbPart))
        return perm

    def finalize(self, row, conn, E):
        dummy = c_int(0)
        self.L.dc_finalize_tree_(c_void_p(row.ctypes.data), c_void_p(conn.ctypes.data), None, None, None,
                                 byref(dummy), _ip(E), _ip(4), _ip(1), _ip(0), _ip(0))

    def store(self, path, E, N):
        self.L.dc_store_tree_(str(path).encode(), _ip(E), _ip(N), _ip(0), _ip(0))

    def read(self, path, E, N):
        comm = c_int(0)
        self.L.dc_read_tree_(str(path).encode(), _ip(E), _ip(N), _ip(0), byref(comm))

    def traversal(self, mesh, dim):
        """The reference's own DC_tree_traversal (OpenMP tasks or the Cilk serial elision)
        running the oracle's callbacks."""
        values = np.full(mesh["nnz"] * dim * dim, np.nan)
        u = make_user(mesh, dim, values, reset_in_seq=0 if self.vec else 1)
        O = oracle()
        seq = ctypes.cast(O.dco_seq_callback, c_void_p)
        vecf = ctypes.cast(O.dco_vec_callback, c_void_p)
        self.L.dc_tree_traversal_(seq, vecf, None, byref(u), None)
        return values


class MineLib:
    """Same driver surface over the product's host library."""

    def __init__(self, vec=0, geometric=False, device=False):
        self.vec = vec
        self.geometric = geometric
        self.device = device          # element passes of the tree proper on the GPU (DC_device.h)
        dc.DC_set_vec_size(vec)

    def create_tree(self, conn, E, N, coord=None):
        dc.DC_set_vec_size(self.vec)
        if self.geometric:
            (dc.DC_create_tree_geometric_device if self.device else dc.DC_create_tree_geometric)(conn, E, 4, N, coord)
        else:
            (dc.DC_create_tree_device if self.device else dc.DC_create_tree)(conn, E, 4, N)

    def permute_double_2d(self, tab, n, dim):
        dc.DC_permute_double_2d_array(tab, n, dim)

    def permute_int_2d(self, tab, n, dim, offset=0):
        dc.DC_permute_int_2d_array(tab, n, dim, offset)

    def permute_int_1d(self, tab, n):
        dc.DC_permute_int_1d_array(tab, n)

    def renumber(self, tab, n):
        dc.DC_renumber_int_array(tab, n)

    def create_permutation(self, part, nbPart):
        return dc.DC_create_permutation(part, nbPart)

    def finalize(self, row, conn, E):
        dc.DC_finalize_tree(row, conn, E, 4)

    def store(self, path, E, N):
        dc.DC_store_tree(path, E, N)

    def read(self, path, E, N):
        dc.DC_read_tree(path, E, N)


scramble_mesh = dc.scramble_mesh          # BASELINE config C4's irregular mesh (dc-lib_b200/meshgen.py)


def prepare_mesh(lib, n, tmp_path, hub=False, tag="t", scramble=False):
    """Appendix-A driver: mesh -> create_tree -> permute/renumber -> CSR -> finalize -> store.
    Returns the user-side arrays plus the stored tree file bytes."""
    conn, N = dc.kuhn_cube(n)
    coord = dc.jitter_coords(n)
    if hub:
        conn, coord, N = dc.hub_collapse(conn, coord, n)
    if scramble:
        conn, coord = scramble_mesh(conn, coord)
    E = conn.shape[0]
    if getattr(lib, "geometric", False):
        lib.create_tree(conn, E, N, coord)
    else:
        lib.create_tree(conn, E, N)
    lib.permute_double_2d(coord, N, 3)
    flat = conn.reshape(-1)
    lib.renumber(flat, 4 * E)
    row, col = dc.build_csr(conn, N)
    lib.finalize(row, conn, E)
    path = os.path.join(str(tmp_path), f"tree_{tag}_{n}.bin")
    lib.store(path, E, N)
    with open(path, "rb") as f:
        tree_bytes = f.read()
    return dict(conn=conn, coord=coord, row=row, col=col, E=E, N=N, nnz=int(row[-1]), tree=tree_bytes,
                tree_path=path, n=n)


def md5(b):
    return hashlib.md5(b if isinstance(b, (bytes, bytearray)) else np.ascontiguousarray(b).tobytes()).hexdigest()


def tree_records(mesh):
    """Pre-order [n,8] records {firstElem,lastElem,lastSep,firstNode,lastNode,isSep,isLeaf,hasSep}
    for dc_plan_build, decoded from the stored tree file by the oracle's parser."""
    nodes, n, _, _ = parse_tree(mesh["tree"], mesh["E"], mesh["N"])
    rec = np.zeros((n, 8), dtype=np.int32)
    for i in range(n):
        t = nodes[i]
        rec[i] = (t.firstElem, t.lastElem, t.lastSep, t.firstNode, t.lastNode, t.isSep, t.isLeaf, 1 if t.sep >= 0 else 0)
    return rec
