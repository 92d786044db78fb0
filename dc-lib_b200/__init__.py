"""dc-lib_b200 -- Python mirror of the DC-lib interface over libdc_b200.so.

The product is the shared library (host C++ API of include/DC.h + the C ABI of
include/dc_cuda.h + the sm_100a kernels).  This module only binds it with ctypes so
that tests and bench.py can drive it the way a C/C++/Fortran caller would; function
names follow the reference (include/DC.h:95-146).  Nothing here computes assembly
values and there is no CPU fallback: if the library is missing or no CUDA device
is present the device entry points raise.

The directory name contains '-', so import it through the repo-root shim:
    import dc_lib_b200 as dc
"""
import ctypes
import os
import subprocess
from ctypes import POINTER, byref, c_double, c_int, c_int32, c_int64, c_uint16, c_uint32, c_void_p, c_char_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DC_B200_LIB: an experiment build of the same library (dc-lib_b200/Makefile, target `variant`)
LIB_PATH = os.environ.get("DC_B200_LIB") or os.path.join(_HERE, "libdc_b200.so")

OP_LAPLACIAN, OP_ELASTICITY = 0, 1
MODE_DETERMINISTIC, MODE_ATOMIC, MODE_LIST = 0, 1, 2
OP_DIM = {OP_LAPLACIAN: 1, OP_ELASTICITY: 3}


def build_library(force=False):
    """Compile libdc_b200.so in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", _HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _HERE, "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(dc-lib_b200 has no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH, mode=os.RTLD_LOCAL)
        _declare(_lib)
    return _lib


class Unit(ctypes.Structure):
    _fields_ = [("edgeFirst", c_int64), ("itemFirst", c_int64), ("tgtFirst", c_int64), ("slotFirst", c_int64),
                ("hnodeFirst", c_int64), ("reserved_", c_int64),
                ("edgeCount", c_int32), ("nodeFirst", c_int32), ("nodeCount", c_int32),
                ("ownElemFirst", c_int32), ("ownElemCount", c_int32), ("haloCount", c_int32),
                ("taskFirst", c_int32), ("taskCount", c_int32), ("itemCount", c_int32),
                ("tgtCount", c_int32), ("hnodeCount", c_int32), ("pad_", c_int32)]


class Task(ctypes.Structure):
    _fields_ = [("itemRel", c_uint32), ("iters", c_uint16), ("kind", c_uint16),
                ("mask", c_uint32), ("first", c_uint32)]


class ContribList(ctypes.Structure):
    _fields_ = [("nbEdges", c_int64), ("nbItems", c_int64), ("edge", POINTER(c_int64)),
                ("ptr", POINTER(c_int64)), ("item", POINTER(c_uint32)), ("edgeT", POINTER(c_int64))]


class Plan(ctypes.Structure):
    _fields_ = [("maxSlots", c_int32), ("maxHalo", c_int32), ("maxNodes", c_int32), ("maxTasks", c_int32), ("maxItems", c_int32),
                ("maxTgt", c_int32), ("symmetric", c_int32), ("nbTgt", c_int64), ("tgt", POINTER(c_int32)),
                ("nbUnits", c_int64), ("nbTasks", c_int64),
                ("nbItems", c_int64), ("nbHalo", c_int64), ("units", POINTER(Unit)),
                ("tasks", POINTER(Task)), ("items", POINTER(c_uint16)),
                ("slotElems", POINTER(c_int32)), ("nbSlotNodes", c_int64), ("slotNodes", POINTER(c_uint16)),
                ("nbHaloNodes", c_int64), ("haloNodes", POINTER(c_int32)), ("diagRel", POINTER(c_int32)),
                ("big", ContribList), ("nbSlotsTotal", c_int64), ("nbItemsUseful", c_int64)]


def _declare(L):
    vp, ip = c_void_p, POINTER(c_int)
    L.dc_cuda_last_error.restype = c_char_p
    L.dc_cuda_create.argtypes = [POINTER(c_void_p), c_int]
    L.dc_cuda_destroy.argtypes = [vp]
    L.dc_cuda_destroy.restype = None
    L.dc_cuda_set_mesh.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, c_int]
    L.dc_cuda_set_mesh_device.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, c_int64, c_int]
    L.dc_cuda_update_coords.argtypes = [vp, vp, vp]
    L.dc_cuda_upload_plan.argtypes = [vp, POINTER(Plan)]
    L.dc_cuda_set_plan_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int]
    L.dc_cuda_set_plan_big_device.argtypes = [vp, vp, vp, vp, vp, c_int64]
    L.dc_cuda_download_range.argtypes = [vp, c_int64, c_int64, vp, vp]
    L.dc_cuda_upload_lists.argtypes = [vp, vp, vp, vp, c_int64, c_int64]
    L.dc_cuda_assemble.argtypes = [vp, c_int, vp, c_int, c_int, vp, vp]
    L.dc_cuda_assemble_with.argtypes = [vp, vp, vp, c_int, c_int, vp, vp]
    L.dc_cuda_values.argtypes = [vp, c_int, POINTER(c_void_p)]
    L.dc_cuda_download_values.argtypes = [vp, c_int, vp, vp]
    L.dc_cuda_permute_double_2d.argtypes = [vp, vp, vp, c_int64, c_int, vp]
    L.dc_cuda_permute_int_2d.argtypes = [vp, vp, vp, c_int64, c_int, vp]
    L.dc_cuda_renumber.argtypes = [vp, vp, c_int64, c_int, vp]
    L.dc_cuda_gather_double_2d.argtypes = [vp, vp, vp, c_int64, c_int, vp]
    L.dc_cuda_gather_int_2d.argtypes = [vp, vp, vp, c_int64, c_int, vp]
    L.dc_cuda_reset_edges.argtypes = [vp, c_int64, c_int64, c_int, vp]
    L.dc_cuda_launch_count.argtypes = [vp]
    L.dc_cuda_launch_count.restype = c_int64
    L.dc_cuda_sync.argtypes = [vp]
    L.dc_cuda_set_option.argtypes = [vp, c_char_p, c_int]
    L.dc_plan_build.argtypes = [POINTER(Plan), vp, c_int, c_int, vp, vp, c_int, vp, c_int64, c_int, c_int, c_int]
    L.dc_plan_set_bank_model.argtypes = [c_int]
    L.dc_plan_set_bank_model.restype = None
    L.dc_contrib_build.argtypes = [POINTER(ContribList), vp, c_int, c_int, vp, vp, c_int, vp, c_int64, c_int]
    L.dc_plan_free.argtypes = [POINTER(Plan)]
    L.dc_plan_free.restype = None
    L.dc_contrib_free.argtypes = [POINTER(ContribList)]
    L.dc_contrib_free.restype = None
    L.dcx_kuhn_cube.argtypes = [c_int, vp, vp, c_double, c_int]
    L.dcx_kuhn_cube.restype = None
    L.dcx_kuhn_slab.argtypes = [c_int, vp, vp, c_double, c_int, c_int]
    L.dcx_kuhn_slab.restype = None
    L.dc_get_node_perm.argtypes = [vp, c_int]
    L.dc_get_elem_perm.argtypes = [vp, c_int]
    L.dc_cuda_spmv.argtypes = [vp, c_int, vp, vp, vp, vp]
    L.dc_cuda_block_jacobi.argtypes = [vp, c_int, vp, vp, POINTER(c_int), vp]
    L.dc_cuda_pack_edges.argtypes = [vp, vp, c_int64, c_int, vp, vp]
    L.dc_cuda_unpack_add_edges.argtypes = [vp, vp, c_int64, c_int, vp, vp]
    L.dc_cuda_exchange_create.argtypes = [POINTER(c_void_p), c_int, c_int, c_int, vp, vp, vp]
    L.dc_cuda_exchange_create_device.argtypes = [POINTER(c_void_p), c_int, c_int, c_int, vp, vp, vp]
    L.dc_cuda_exchange_neighbours.argtypes = [vp]
    L.dc_cuda_exchange_export.argtypes = [vp, c_int, vp]
    L.dc_cuda_exchange_import.argtypes = [vp, c_int, vp]
    L.dc_cuda_exchange_pack.argtypes = [vp, vp, vp]
    L.dc_cuda_exchange_unpack.argtypes = [vp, vp, vp]
    L.dc_cuda_exchange_run.argtypes = [vp, vp, vp]
    L.dc_cuda_exchange_bytes.argtypes = [vp]
    L.dc_cuda_exchange_bytes.restype = c_int64
    L.dc_cuda_exchange_set_blocks.argtypes = [vp, c_int]
    L.dc_cuda_exchange_destroy.argtypes = [vp]
    L.dc_cuda_exchange_destroy.restype = None
    L.dc_cuda_exchange_last_error.restype = c_char_p
    L.dc_cuda_set_front_units.argtypes = [vp, c_int64]
    L.dc_cuda_assemble_units.argtypes = [vp, c_int, vp, c_int, vp, vp, c_int64, c_int64, c_int, c_int]
    L.dc_cuda_assemble_exchange.argtypes = [vp, vp, c_int, vp, c_int, vp, vp]
    L.dc_cuda_assemble_exchange_with.argtypes = [vp, vp, vp, vp, c_int, vp, vp]
    L.dc_interface_edges.argtypes = [vp, vp, c_int, vp, c_int, vp, vp]
    L.dc_interface_edges.restype = c_int64
    L.dc_plan_order_units.argtypes = [POINTER(Plan), vp]
    L.dc_plan_order_units.restype = c_int64
    L.dc_cuda_csr_begin.argtypes = [vp, c_int, c_int, vp, POINTER(c_int64), POINTER(c_void_p), vp]
    L.dc_cuda_csr_columns.argtypes = [vp, vp, c_int, vp]
    L.dc_cuda_csr_node_to_elem.argtypes = [vp, vp, vp, vp]
    L.dc_cuda_csr_end.argtypes = [vp]
    L.dc_cuda_csr_end.restype = None
    L.dcx_build_csr.argtypes = [vp, c_int, c_int, vp, vp, c_int]
    L.dcx_build_csr.restype = c_int64
    L.dc_flatten_tree.argtypes = [vp, c_int64]
    L.dc_flatten_tree.restype = c_int64
    L.dc_get_time_.restype = c_double
    L.dc_create_tree_.argtypes = [vp, ip, ip, ip]
    L.dc_create_tree_geometric_.argtypes = [vp, ip, ip, ip, vp, ip]
    L.dc_create_tree_device_.argtypes = [vp, ip, ip, ip]
    L.dc_device_select_.argtypes = [ip]
    L.dc_create_tree_geometric_device_.argtypes = [vp, ip, ip, ip, vp, ip]
    L.dc_partition_mesh_.argtypes = [vp, ip, ip, ip, ip, vp, ip, vp]
    L.dc_finalize_tree_.argtypes = [vp, vp, vp, vp, vp, vp, ip, ip, ip, ip, ip]
    L.dc_store_tree_.argtypes = [c_char_p, ip, ip, ip, ip]
    L.dc_read_tree_.argtypes = [c_char_p, ip, ip, ip, ip]
    L.dc_permute_double_2d_array_.argtypes = [vp, ip, ip]
    L.dc_permute_int_2d_array_.argtypes = [vp, ip, ip, ip]
    L.dc_permute_int_1d_array_.argtypes = [vp, ip]
    L.dc_renumber_int_array_.argtypes = [vp, ip]
    L.dc_create_permutation_.argtypes = [vp, vp, ip, ip]
    L.dc_tree_traversal_.argtypes = [vp, vp, vp, vp, vp]
    L.dc_assembly_.argtypes = [vp, vp, vp, vp, ip]
    L.dc_set_vec_size_.argtypes = [ip]
    L.dc_set_max_elem_per_part_.argtypes = [ip]
    L.dc_set_num_threads_.argtypes = [ip]


def _p(a):
    """Raw address of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()          # torch tensor


def _i(v):
    return byref(c_int(int(v)))


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a


# --------------------------------------------------------------------------- host API
# Same names and argument meaning as the reference (include/DC.h:95-146); arrays are
# numpy arrays modified in place where the reference modifies them in place.

def DC_set_vec_size(v):
    lib().dc_set_vec_size_(_i(v))


def DC_set_max_elem_per_part(v):
    lib().dc_set_max_elem_per_part_(_i(v))


def DC_set_num_threads(v):
    lib().dc_set_num_threads_(_i(v))


def DC_create_tree(elemToNode, nbElem, dimElem, nbNodes):
    assert elemToNode.dtype == np.int32 and elemToNode.flags.c_contiguous
    lib().dc_create_tree_(_p(elemToNode), _i(nbElem), _i(dimElem), _i(nbNodes))


def DC_create_tree_geometric(elemToNode, nbElem, dimElem, nbNodes, coord):
    """DC_create_tree with recursive coordinate bisection instead of METIS (coord in the
    ORIGINAL node numbering)."""
    assert elemToNode.dtype == np.int32 and elemToNode.flags.c_contiguous
    assert coord.dtype == np.float64 and coord.flags.c_contiguous
    lib().dc_create_tree_geometric_(_p(elemToNode), _i(nbElem), _i(dimElem), _i(nbNodes), _p(coord), _i(coord.shape[1]))


def DC_device_select(device):
    """CUDA device of DC_create_tree_device / DC_create_tree_geometric_device (include/DC_device.h)."""
    lib().dc_device_select_(_i(device))


def DC_create_tree_device(elemToNode, nbElem, dimElem, nbNodes):
    """DC_create_tree with the level-wise element passes of the tree proper on the GPU
    (include/DC_device.h); same bytes as DC_create_tree."""
    assert elemToNode.dtype == np.int32 and elemToNode.flags.c_contiguous
    lib().dc_create_tree_device_(_p(elemToNode), _i(nbElem), _i(dimElem), _i(nbNodes))


def DC_create_tree_geometric_device(elemToNode, nbElem, dimElem, nbNodes, coord):
    """DC_create_tree_geometric with the element passes of the tree proper on the GPU."""
    assert elemToNode.dtype == np.int32 and elemToNode.flags.c_contiguous
    assert coord.dtype == np.float64 and coord.flags.c_contiguous
    lib().dc_create_tree_geometric_device_(_p(elemToNode), _i(nbElem), _i(dimElem), _i(nbNodes), _p(coord),
                                           _i(coord.shape[1]))


def DC_partition_mesh(elemToNode, nbNodes, nbPart, coord=None):
    """Element partition for nbPart devices: METIS on the nodal graph (coord None) or coordinate
    bisection of the nodes; an element follows the majority of its nodes.  Returns int32[nbElem]."""
    assert elemToNode.dtype == np.int32 and elemToNode.flags.c_contiguous
    nbElem, dimElem = elemToNode.shape
    part = np.zeros(nbElem, dtype=np.int32)
    if coord is not None:
        assert coord.dtype == np.float64 and coord.flags.c_contiguous
    lib().dc_partition_mesh_(_p(elemToNode), _i(nbElem), _i(dimElem), _i(nbNodes), _i(nbPart),
                             _p(coord) if coord is not None else None, _i(coord.shape[1] if coord is not None else 3),
                             _p(part))
    return part


def DC_finalize_tree(nodeToNodeRow, elemToNode, nbElem, dimElem):
    assert nodeToNodeRow.dtype == np.int32
    dummy = c_int(0)
    lib().dc_finalize_tree_(_p(nodeToNodeRow), _p(elemToNode), None, None, None,
                            ctypes.cast(byref(dummy), c_void_p), _i(nbElem), _i(dimElem), _i(1), _i(0), _i(0))


def DC_store_tree(path, nbElem, nbNodes):
    lib().dc_store_tree_(str(path).encode(), _i(nbElem), _i(nbNodes), _i(0), _i(0))


def DC_read_tree(path, nbElem, nbNodes):
    comm = c_int(0)
    lib().dc_read_tree_(str(path).encode(), _i(nbElem), _i(nbNodes), _i(0), byref(comm))
    return comm.value


def DC_permute_double_2d_array(tab, nbItem, dimItem):
    assert tab.dtype == np.float64 and tab.flags.c_contiguous
    lib().dc_permute_double_2d_array_(_p(tab), _i(nbItem), _i(dimItem))


def DC_permute_int_2d_array(tab, nbItem, dimItem, offset=0):
    assert tab.dtype == np.int32 and tab.flags.c_contiguous
    lib().dc_permute_int_2d_array_(_p(tab), _i(nbItem), _i(dimItem), _i(offset))


def DC_permute_int_1d_array(tab, size):
    assert tab.dtype == np.int32 and tab.flags.c_contiguous
    lib().dc_permute_int_1d_array_(_p(tab), _i(size))


def DC_renumber_int_array(tab, size):
    assert tab.dtype == np.int32 and tab.flags.c_contiguous
    lib().dc_renumber_int_array_(_p(tab), _i(size))


def DC_create_permutation(part, nbPart):
    part = _i32(part)
    perm = np.full(part.shape[0], -1, dtype=np.int32)
    lib().dc_create_permutation_(_p(perm), _p(part), _i(part.shape[0]), _i(nbPart))
    return perm


def DC_free_tree():
    lib().dc_free_tree_()


def flatten_tree():
    """Pre-order records of the current tree: [nbTreeNodes, 8] int32
    {firstElem,lastElem,lastSep,firstNode,lastNode,isSep,isLeaf,hasSep}."""
    n = lib().dc_flatten_tree(None, 0)
    out = np.zeros((n, 8), dtype=np.int32)
    lib().dc_flatten_tree(_p(out), n)
    return out


# --------------------------------------------------------------------------- schedule

class HostPlan:
    """Owns a dc_plan_t built by dc_plan_build()."""

    def __init__(self, conn, nbNodes, row, col, colBase, treeRec, maxSlots=400, nbThreads=0, symmetric=True,
                 bank_model=0):
        """bank_model: record layout the slot numbering is optimised for (dc_plan_set_bank_model):
        0 = node gradients (elasticity, default), 1 = 4x4 symmetric entries (P1 Laplacian).
        symmetric=True: every off-diagonal block is computed once, by the lane of the upper edge
        (row < col), which also stores its transpose (operators with bitwise symmetric element
        matrices); False: one lane per edge, every unit writes exactly its own edge interval."""
        self.c = Plan()
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        nrec = 0 if treeRec is None else treeRec.shape[0]
        lib().dc_plan_set_bank_model(int(bank_model))
        try:
            rc = lib().dc_plan_build(byref(self.c), _p(conn), conn.shape[0], nbNodes, _p(row), _p(col), colBase,
                                     _p(treeRec), nrec, maxSlots, nbThreads, 1 if symmetric else 0)
        finally:
            lib().dc_plan_set_bank_model(0)
        if rc:
            raise RuntimeError(f"dc_plan_build failed with code {rc}")

    def arrays(self):
        """Zero-copy numpy views of the schedule arrays (valid while this object lives)."""
        c = self.c
        as_np = np.ctypeslib.as_array

        def view(ptr, count, dtype, width=1):
            if count == 0:
                return np.zeros((0,) if width == 1 else (0, width), dtype=dtype)
            a = as_np(ctypes.cast(ptr, POINTER(ctypes.c_uint8)), shape=(count * width * np.dtype(dtype).itemsize,))
            a = a.view(dtype)
            return a if width == 1 else a.reshape(count, width)

        return dict(units=view(c.units, c.nbUnits, np.uint8, 96), tasks=view(c.tasks, c.nbTasks, np.uint8, 16),
                    items=view(c.items, c.nbItems, np.uint16), slotNodes=view(c.slotNodes, c.nbSlotNodes, np.uint16),
                    haloNodes=view(c.haloNodes, c.nbHaloNodes, np.int32), tgt=view(c.tgt, c.nbTgt, np.int32, 2),
                    nbUnits=int(c.nbUnits), maxSlots=int(c.maxSlots), maxHalo=int(c.maxHalo), maxNodes=int(c.maxNodes),
                    maxTasks=int(c.maxTasks), maxItems=int(c.maxItems), maxTgt=int(c.maxTgt),
                    symmetric=int(c.symmetric), bigEdges=int(c.big.nbEdges),
                    bigEdge=view(c.big.edge, c.big.nbEdges, np.int64), bigPtr=view(c.big.ptr, c.big.nbEdges + 1 if c.big.nbEdges else 0, np.int64),
                    bigItem=view(c.big.item, c.big.nbItems, np.uint32),
                    bigEdgeT=view(c.big.edgeT, c.big.nbEdges if c.symmetric else 0, np.int64))

    def order_units(self, node_flag):
        """Units that own a flagged row (uint8 per node: the interface nodes) move to the front of the
        schedule (dc_plan_order_units); returns their number, for Context.set_front_units."""
        flag = np.ascontiguousarray(node_flag, dtype=np.uint8)
        return int(lib().dc_plan_order_units(byref(self.c), _p(flag)))

    def diag_rel(self, nbNodes):
        return np.ctypeslib.as_array(self.c.diagRel, shape=(max(nbNodes, 1),))[:nbNodes]

    def stats(self):
        c = self.c
        return dict(units=c.nbUnits, tasks=c.nbTasks, items=c.nbItems, items_useful=c.nbItemsUseful, symmetric=c.symmetric,
                    halo=c.nbHalo, slots_total=c.nbSlotsTotal, max_slots=c.maxSlots, big_edges=c.big.nbEdges)

    def __del__(self):
        try:
            lib().dc_plan_free(byref(self.c))
        except Exception:
            pass


class Context:
    """Thin wrapper of the C ABI (include/dc_cuda.h).  Array arguments are numpy arrays
    (host entry points) or torch CUDA tensors (device entry points)."""

    def __init__(self, device=0):
        self.h = c_void_p()
        self._keep = []
        if lib().dc_cuda_create(byref(self.h), device):
            raise RuntimeError("dc_cuda_create: " + lib().dc_cuda_last_error().decode())

    def _ck(self, rc, what):
        if rc:
            raise RuntimeError(f"{what}: {lib().dc_cuda_last_error().decode()} (code {rc})")

    def set_mesh(self, conn, coord, row, col, colBase=0):
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        coord = np.ascontiguousarray(coord, dtype=np.float64)
        row, col = _i32(row), _i32(col)
        self.nbElem, self.nbNodes, self.nnz = conn.shape[0], coord.shape[0], int(row[-1])
        self._ck(lib().dc_cuda_set_mesh(self.h, _p(conn), _p(coord), _p(row), _p(col), self.nbElem,
                                        self.nbNodes, colBase), "dc_cuda_set_mesh")

    def set_mesh_device(self, conn, coord, row, col, nnz, colBase=0):
        self._keep = [conn, coord, row, col]
        self.nbElem, self.nbNodes, self.nnz = conn.shape[0], coord.shape[0], int(nnz)
        self._ck(lib().dc_cuda_set_mesh_device(self.h, _p(conn), _p(coord), _p(row), _p(col), self.nbElem,
                                               self.nbNodes, self.nnz, colBase), "dc_cuda_set_mesh_device")

    def update_coords(self, coord, stream=None):
        self._ck(lib().dc_cuda_update_coords(self.h, _p(coord), stream), "dc_cuda_update_coords")

    def upload_plan(self, plan):
        self._ck(lib().dc_cuda_upload_plan(self.h, byref(plan.c)), "dc_cuda_upload_plan")

    def set_plan_device(self, t):
        """t: dict of torch CUDA tensors {units,tasks,items,slotNodes,haloNodes,diagRel,tgt} + the plan's ints."""
        self._keep_plan = t
        self._ck(lib().dc_cuda_set_plan_device(self.h, _p(t["units"]), _p(t["tasks"]), _p(t["items"]),
                                               _p(t["slotNodes"]), _p(t["haloNodes"]), _p(t["diagRel"]), _p(t["tgt"]),
                                               t["nbUnits"], t["maxSlots"], t["maxHalo"], t["maxNodes"], t["maxTasks"],
                                               t["maxItems"], t["maxTgt"], t["symmetric"]), "dc_cuda_set_plan_device")
        if t.get("bigEdges"):
            self._ck(lib().dc_cuda_set_plan_big_device(self.h, _p(t["bigEdge"]), _p(t["bigPtr"]), _p(t["bigItem"]),
                                                       _p(t["bigEdgeT"]) if t["symmetric"] else None, t["bigEdges"]),
                     "dc_cuda_set_plan_big_device")

    def download_range(self, d_values, first, count, h_buf, stream=None):
        self._ck(lib().dc_cuda_download_range(_p(d_values), first, count, _p(h_buf), stream), "dc_cuda_download_range")

    def sync(self, stream=None):
        self._ck(lib().dc_cuda_sync(stream), "dc_cuda_sync")

    def upload_lists(self, conn, nbNodes, row, col, colBase=0):
        cl = ContribList()
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        rc = lib().dc_contrib_build(byref(cl), _p(conn), conn.shape[0], nbNodes, _p(row), _p(col), colBase, None, 0, 0)
        if rc:
            raise RuntimeError(f"dc_contrib_build failed with code {rc}")
        try:
            self._ck(lib().dc_cuda_upload_lists(self.h, ctypes.cast(cl.edge, c_void_p), ctypes.cast(cl.ptr, c_void_p),
                                                ctypes.cast(cl.item, c_void_p), cl.nbEdges, cl.nbItems),
                     "dc_cuda_upload_lists")
        finally:
            lib().dc_contrib_free(byref(cl))

    def assemble(self, op, params, mode, values, stream=None):
        """values: torch CUDA float64 tensor with nnz*dim*dim entries."""
        prm = np.ascontiguousarray(params if params is not None else [], dtype=np.float64)
        self._ck(lib().dc_cuda_assemble(self.h, op, _p(prm) if prm.size else None, prm.size, mode, _p(values), stream),
                 "dc_cuda_assemble")

    def assemble_with(self, launcher, params, mode, values, stream=None):
        """launcher: address of a function defined with DC_DEVICE_OPERATOR in a user .cu file."""
        prm = np.ascontiguousarray(params if params is not None else [], dtype=np.float64)
        self._ck(lib().dc_cuda_assemble_with(self.h, launcher, _p(prm) if prm.size else None, prm.size, mode,
                                             _p(values), stream), "dc_cuda_assemble_with")

    def set_front_units(self, n):
        self._ck(lib().dc_cuda_set_front_units(self.h, int(n)), "dc_cuda_set_front_units")

    def assemble_units(self, op, params, values, first, count, reserve=0, skip_big=False, stream=None):
        prm = np.ascontiguousarray(params if params is not None else [], dtype=np.float64)
        self._ck(lib().dc_cuda_assemble_units(self.h, op, _p(prm) if prm.size else None, prm.size, _p(values), stream,
                                              int(first), int(count), int(reserve), 1 if skip_big else 0),
                 "dc_cuda_assemble_units")

    def assemble_exchange(self, exchange, op, params, values, stream=None):
        """One step of a multi-GPU rank: front units, pack || other units, unpack (dc_cuda_assemble_exchange)."""
        prm = np.ascontiguousarray(params if params is not None else [], dtype=np.float64)
        self._ck(lib().dc_cuda_assemble_exchange(self.h, exchange.h if exchange is not None else None, op,
                                                 _p(prm) if prm.size else None, prm.size, _p(values), stream),
                 "dc_cuda_assemble_exchange")

    def assemble_to_host(self, op, params, mode, out, stream=None):
        """The call a reference user makes: values land in the HOST array `out`."""
        dim = OP_DIM[op]
        d = c_void_p()
        self._ck(lib().dc_cuda_values(self.h, dim, byref(d)), "dc_cuda_values")
        prm = np.ascontiguousarray(params if params is not None else [], dtype=np.float64)
        self._ck(lib().dc_cuda_assemble(self.h, op, _p(prm) if prm.size else None, prm.size, mode, d, stream),
                 "dc_cuda_assemble")
        self._ck(lib().dc_cuda_download_values(self.h, dim, _p(out), stream), "dc_cuda_download_values")
        self._ck(lib().dc_cuda_sync(stream), "dc_cuda_sync")

    def spmv(self, dim, values, x, y, stream=None):
        """y = A x with the assembled block-CSR values (torch CUDA tensors)."""
        self._ck(lib().dc_cuda_spmv(self.h, dim, _p(values), _p(x), _p(y), stream), "dc_cuda_spmv")

    def block_jacobi(self, dim, values, inv_diag, stream=None):
        """inv_diag[N][dim*dim] = inverse diagonal blocks of the assembled matrix; returns the number of bad rows."""
        bad = c_int(0)
        self._ck(lib().dc_cuda_block_jacobi(self.h, dim, _p(values), _p(inv_diag), byref(bad), stream), "dc_cuda_block_jacobi")
        return bad.value

    def set_option(self, name, value):
        self._ck(lib().dc_cuda_set_option(self.h, name.encode(), int(value)), "dc_cuda_set_option")

    def launch_count(self):
        return int(lib().dc_cuda_launch_count(self.h))

    def close(self):
        if self.h:
            lib().dc_cuda_destroy(self.h)
            self.h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceExchange:
    """Binding of dc_cuda_exchange_* (include/dc_cuda.h): interface exchange over peer memory.
    intf: {neighbour rank: local CSR edges shared with it (numpy int64 or torch CUDA int64), in the
    order both sides agree on}.  Handles travel through `dist` (torch.distributed, one rank per
    process) or are swapped directly with connect_local() when the ranks live in one process."""

    HANDLE_BYTES = 128

    def __init__(self, intf, per_edge, device_index):
        import torch
        self.nbs = sorted(intf)
        self.per = per_edge
        counts = [int(intf[nb].shape[0]) if hasattr(intf[nb], "shape") else len(intf[nb]) for nb in self.nbs]
        ptr = np.zeros(len(self.nbs) + 1, dtype=np.int64)
        ptr[1:] = np.cumsum(counts)
        ranks = np.ascontiguousarray(self.nbs, dtype=np.int32)
        self.h = c_void_p()
        on_device = any(isinstance(intf[nb], torch.Tensor) and intf[nb].is_cuda for nb in self.nbs)
        if on_device:
            self._edges = (torch.cat([intf[nb].to(torch.int64).reshape(-1) for nb in self.nbs]) if self.nbs
                           else torch.zeros(0, dtype=torch.int64, device=f"cuda:{device_index}"))
            rc = lib().dc_cuda_exchange_create_device(byref(self.h), device_index, per_edge, len(self.nbs), _p(ranks), _p(ptr),
                                                      _p(self._edges))
        else:
            edges = (np.ascontiguousarray(np.concatenate([np.asarray(intf[nb], dtype=np.int64) for nb in self.nbs]))
                     if self.nbs else np.zeros(0, dtype=np.int64))
            rc = lib().dc_cuda_exchange_create(byref(self.h), device_index, per_edge, len(self.nbs), _p(ranks), _p(ptr), _p(edges))
        if rc:
            raise RuntimeError("dc_cuda_exchange_create: " + lib().dc_cuda_exchange_last_error().decode())
        self.bytes_per_step = int(lib().dc_cuda_exchange_bytes(self.h))
        self.launches_per_step = 2 * len(self.nbs)

    def _ck(self, rc, what):
        if rc:
            raise RuntimeError(f"{what}: {lib().dc_cuda_exchange_last_error().decode()} (code {rc})")

    def export(self, nb):
        buf = ctypes.create_string_buffer(self.HANDLE_BYTES)
        self._ck(lib().dc_cuda_exchange_export(self.h, self.nbs.index(nb), buf), "dc_cuda_exchange_export")
        return buf.raw

    def import_(self, nb, handle):
        self._ck(lib().dc_cuda_exchange_import(self.h, self.nbs.index(nb), handle), "dc_cuda_exchange_import")

    def connect(self, dist, rank, world):
        """Swap the handles with the neighbours through torch.distributed (any backend)."""
        mine = {(rank, nb): self.export(nb) for nb in self.nbs}
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        for nb in self.nbs:
            self.import_(nb, everyone[nb][(nb, rank)])

    @staticmethod
    def connect_local(exchanges):
        """exchanges: {rank: DeviceExchange} living in this process."""
        for r, x in exchanges.items():
            for nb in x.nbs:
                x.import_(nb, exchanges[nb].export(r))

    def pack(self, values, stream=None):
        self._ck(lib().dc_cuda_exchange_pack(self.h, _p(values), stream), "dc_cuda_exchange_pack")

    def unpack(self, values, stream=None):
        self._ck(lib().dc_cuda_exchange_unpack(self.h, _p(values), stream), "dc_cuda_exchange_unpack")

    def run(self, values, stream=None):
        self._ck(lib().dc_cuda_exchange_run(self.h, _p(values), stream), "dc_cuda_exchange_run")

    def close(self):
        if self.h:
            lib().dc_cuda_exchange_destroy(self.h)
            self.h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def interface_edges(row, col, nodes, colBase=0):
    """CSR edges among the listed interface nodes (1-based, the reference's intfNodes slice of one
    neighbour), ordered by (position of the row node, position of the column node): (edges, keys)."""
    row, col, nodes = _i32(row), _i32(col), _i32(nodes)
    n = lib().dc_interface_edges(_p(row), _p(col), colBase, _p(nodes), nodes.shape[0], None, None)
    edges, keys = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
    lib().dc_interface_edges(_p(row), _p(col), colBase, _p(nodes), nodes.shape[0], _p(edges), _p(keys))
    return edges, keys


def permute_double_2d(d_in, d_out, d_perm, dimItem, stream=None):
    rc = lib().dc_cuda_permute_double_2d(_p(d_in), _p(d_out), _p(d_perm), d_perm.shape[0], dimItem, stream)
    if rc:
        raise RuntimeError(lib().dc_cuda_last_error().decode())


def permute_int_2d(d_in, d_out, d_perm, dimItem, stream=None):
    rc = lib().dc_cuda_permute_int_2d(_p(d_in), _p(d_out), _p(d_perm), d_perm.shape[0], dimItem, stream)
    if rc:
        raise RuntimeError(lib().dc_cuda_last_error().decode())


def gather_double_2d(d_in, d_out, d_inv, dimItem, stream=None):
    """d_out[d] = d_in[inv[d]] (inv = inverse permutation: new position -> old position)."""
    rc = lib().dc_cuda_gather_double_2d(_p(d_in), _p(d_out), _p(d_inv), d_inv.shape[0], dimItem, stream)
    if rc:
        raise RuntimeError(lib().dc_cuda_last_error().decode())


def gather_int_2d(d_in, d_out, d_inv, dimItem, stream=None):
    rc = lib().dc_cuda_gather_int_2d(_p(d_in), _p(d_out), _p(d_inv), d_inv.shape[0], dimItem, stream)
    if rc:
        raise RuntimeError(lib().dc_cuda_last_error().decode())


def renumber(d_tab, d_perm, isFortran=True, stream=None):
    rc = lib().dc_cuda_renumber(_p(d_tab), _p(d_perm), d_tab.numel(), 1 if isFortran else 0, stream)
    if rc:
        raise RuntimeError(lib().dc_cuda_last_error().decode())


def reset_edges(d_values, firstEdge, lastEdge, perEdge, stream=None):
    rc = lib().dc_cuda_reset_edges(_p(d_values), firstEdge, lastEdge, perEdge, stream)
    if rc:
        raise RuntimeError(lib().dc_cuda_last_error().decode())


def kuhn_cube_fast(n, amplitude=0.3, seed=12345, x_offset=0, coords_only=False):
    """C++/OpenMP version of kuhn_cube + jitter_coords (identical arrays).  x_offset places
    the cube that many cells along x inside a bar of cubes (shared planes get equal nodes)."""
    E, N = 6 * n ** 3, (n + 1) ** 3
    conn = None if coords_only else np.empty((E, 4), dtype=np.int32)
    coord = np.empty((N, 3), dtype=np.float64)
    lib().dcx_kuhn_slab(n, _p(conn), _p(coord), amplitude, seed, x_offset)
    return conn, coord, N


def get_node_perm(nbNodes):
    out = np.empty(nbNodes, dtype=np.int32)
    if lib().dc_get_node_perm(_p(out), nbNodes):
        raise RuntimeError("no node permutation (call between DC_create_tree/DC_read_tree and DC_store_tree)")
    return out


def pack_edges(d_values, d_edge, perEdge, d_buf, stream=None):
    rc = lib().dc_cuda_pack_edges(_p(d_values), _p(d_edge), d_edge.numel(), perEdge, _p(d_buf), stream)
    if rc:
        raise RuntimeError(lib().dc_cuda_last_error().decode())


def unpack_add_edges(d_values, d_edge, perEdge, d_buf, stream=None):
    rc = lib().dc_cuda_unpack_add_edges(_p(d_values), _p(d_edge), d_edge.numel(), perEdge, _p(d_buf), stream)
    if rc:
        raise RuntimeError(lib().dc_cuda_last_error().decode())


def build_csr_device(d_conn, nbNodes, colBase=0, node_to_elem=False, stream=None):
    """CSR pattern of the renumbered mesh built on the GPU (include/dc_cuda.h, dc_cuda_csr_*).
    d_conn: torch int32 CUDA tensor [nbElem, 4], 1-based.  Returns (row, col) as torch CUDA tensors,
    plus (n2e_ptr, n2e_val) when node_to_elem is set; identical to build_csr_fast's arrays."""
    import torch
    assert d_conn.is_cuda and d_conn.dtype == torch.int32 and d_conn.is_contiguous()
    E = d_conn.shape[0]
    row = torch.empty(nbNodes + 1, dtype=torch.int32, device=d_conn.device)
    nnz, h = c_int64(0), c_void_p()
    if lib().dc_cuda_csr_begin(_p(d_conn), E, nbNodes, _p(row), byref(nnz), byref(h), stream):
        raise RuntimeError("dc_cuda_csr_begin: " + lib().dc_cuda_last_error().decode())
    try:
        col = torch.empty(nnz.value, dtype=torch.int32, device=d_conn.device)
        if lib().dc_cuda_csr_columns(h, _p(col), colBase, stream):
            raise RuntimeError("dc_cuda_csr_columns: " + lib().dc_cuda_last_error().decode())
        out = (row, col)
        if node_to_elem:
            ptr = torch.empty(nbNodes + 1, dtype=torch.int64, device=d_conn.device)
            val = torch.empty(4 * E, dtype=torch.int32, device=d_conn.device)
            if lib().dc_cuda_csr_node_to_elem(h, _p(ptr), _p(val), stream):
                raise RuntimeError("dc_cuda_csr_node_to_elem: " + lib().dc_cuda_last_error().decode())
            out = (row, col, ptr, val)
        if lib().dc_cuda_sync(stream):
            raise RuntimeError("dc_cuda_sync: " + lib().dc_cuda_last_error().decode())
    finally:
        lib().dc_cuda_csr_end(h)
    return out


def build_csr_fast(conn, nbNodes, colBase=0):
    """C++/OpenMP version of build_csr (identical arrays)."""
    row = np.empty(nbNodes + 1, dtype=np.int32)
    nnz = lib().dcx_build_csr(_p(conn), conn.shape[0], nbNodes, _p(row), None, colBase)
    if nnz < 0:
        raise RuntimeError("CSR pattern exceeds 2^31-1 edges")
    col = np.empty(nnz, dtype=np.int32)
    lib().dcx_build_csr(_p(conn), conn.shape[0], nbNodes, _p(row), _p(col), colBase)
    return row, col


from .meshgen import kuhn_cube, jitter_coords, build_csr, hub_collapse, scramble_mesh  # noqa: E402,F401
