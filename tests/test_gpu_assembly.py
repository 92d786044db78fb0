"""Parity tests proper: the CUDA assembly, called through the C ABI (include/dc_cuda.h),
against the CPU oracle on the same seeded meshes and against the golden vectors produced by
the reference itself.  Deterministic / list modes: bit-exact.  Atomic mode: max relative
error <= 1e-12 (the north-star tolerance; order of the red.global adds is free)."""
import json
import os

import numpy as np
import pytest

import common
import dc_lib_b200 as dc
from common import GOLDEN, LAMBDA, MU, MineLib, md5, oracle_serial, prepare_mesh, tree_records

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN, "golden.json")) as f:
    GOLD = json.load(f)

OPS = {1: (dc.OP_LAPLACIAN, None), 3: (dc.OP_ELASTICITY, [LAMBDA, MU])}
REL_TOL = 1e-12


def _ctx(mesh, slots=832, lists=False, with_tree=True, sym=True, bank_model=0):
    ctx = dc.Context(0)
    ctx.set_mesh(mesh["conn"], mesh["coord"], mesh["row"], mesh["col"], 0)
    rec = tree_records(mesh) if with_tree else None
    plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, rec, slots, symmetric=sym,
                       bank_model=bank_model)
    ctx.upload_plan(plan)
    if lists:
        ctx.upload_lists(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0)
    return ctx, plan


def _run(ctx, mesh, dim, mode):
    import torch
    op, prm = OPS[dim]
    vals = torch.full((mesh["nnz"] * dim * dim,), float("nan"), dtype=torch.float64, device="cuda:0")
    ctx.assemble(op, prm, mode, vals)
    torch.cuda.synchronize()
    return vals.cpu().numpy()


def _rel_err(got, want):
    scale = np.abs(want).max()
    return np.abs(got - want).max() / scale


@pytest.mark.parametrize("key", [k for k, v in GOLD.items() if v["n"] <= 20])
@pytest.mark.parametrize("dim", [1, 3])
def test_deterministic_matches_reference_golden(tmp_path, key, dim):
    """Values produced by the reference's own traversal (golden md5) == GPU deterministic mode."""
    g = GOLD[key]
    mesh = prepare_mesh(MineLib(g["vec"]), g["n"], tmp_path, hub=g["hub"])
    ctx, _ = _ctx(mesh)
    got = _run(ctx, mesh, dim, dc.MODE_DETERMINISTIC)
    assert md5(got) == g[f"values_md5_d{dim}"]
    ctx2, _ = _ctx(mesh, slots=300, sym=False)                    # general (non-symmetric) schedule
    assert _run(ctx2, mesh, dim, dc.MODE_DETERMINISTIC).tobytes() == got.tobytes()
    raw = os.path.join(GOLDEN, f"{key}_values_d1.npy")
    if dim == 1 and os.path.exists(raw):
        assert np.load(raw).tobytes() == got.tobytes()


@pytest.mark.parametrize("key,dim", [("kuhn40_vec0", 1), ("kuhn40_vec4", 3), ("kuhn40_vec4", 1)])
def test_baseline_configs_c1_c2(tmp_path, key, dim):
    """BASELINE configs: 40^3 cube, Laplacian on the pure tree (C1) and elasticity on the
    D&C + colouring tree (C2)."""
    g = GOLD[key]
    mesh = prepare_mesh(MineLib(g["vec"]), g["n"], tmp_path)
    ctx, _ = _ctx(mesh)
    got = _run(ctx, mesh, dim, dc.MODE_DETERMINISTIC)
    assert md5(got) == g[f"values_md5_d{dim}"]
    fast = _run(ctx, mesh, dim, dc.MODE_ATOMIC)
    assert _rel_err(fast, got) <= REL_TOL


@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("n,vec,dim,slots,hub", [
    (8, 0, 1, 832, False), (8, 4, 3, 832, False), (12, 0, 3, 200, False), (12, 2, 1, 64, False),
    (16, 8, 3, 832, False), (18, 0, 3, 832, True), (18, 4, 1, 100, True), (18, 0, 3, 100, True),
    (3, 0, 3, 832, False), (1, 0, 1, 832, False)])
def test_all_modes_against_oracle(tmp_path, n, vec, dim, slots, hub, sym):
    """Small slot budgets force leaf splitting and the big-row fallback; hub meshes have
    nodes of valence ~100 (BASELINE config C5)."""
    mesh = prepare_mesh(MineLib(vec), n, tmp_path, hub=hub)
    want = oracle_serial(mesh, dim)
    ctx, plan = _ctx(mesh, slots=slots, lists=True, sym=sym)
    det = _run(ctx, mesh, dim, dc.MODE_DETERMINISTIC)
    assert det.tobytes() == want.tobytes()
    lst = _run(ctx, mesh, dim, dc.MODE_LIST)
    assert lst.tobytes() == want.tobytes()
    fast = _run(ctx, mesh, dim, dc.MODE_ATOMIC)
    assert _rel_err(fast, want) <= REL_TOL
    if slots <= 100 and hub:
        assert plan.c.big.nbEdges > 0           # the fallback really ran


@pytest.mark.parametrize("persistent", [0, 1])
@pytest.mark.parametrize("threads", [128, 256, 512])
def test_kernel_variants_bit_exact(tmp_path, threads, persistent):
    """Every instantiation of the unit kernel (CTA size, persistent grid or one CTA per unit), on a
    plan with several units per resident CTA so that the persistent loop really cycles."""
    mesh = prepare_mesh(MineLib(0), 16, tmp_path)
    ctx, plan = _ctx(mesh, slots=64)
    assert plan.c.nbUnits > 2 * 4 * 148        # four 128-thread CTAs per SM at most
    ctx.set_option("unit_threads", threads)
    ctx.set_option("persistent", persistent)
    for dim in (1, 3):
        assert _run(ctx, mesh, dim, dc.MODE_DETERMINISTIC).tobytes() == oracle_serial(mesh, dim).tobytes()


@pytest.mark.parametrize("dim,vec", [(3, 0), (1, 4)])
def test_graded_shuffled_mesh_bit_exact(tmp_path, dim, vec):
    """BASELINE config C4's mesh at test size: graded coordinates, random node labels, shuffled
    elements before DC_create_tree (METIS tree)."""
    mesh = prepare_mesh(MineLib(vec), 12, tmp_path, scramble=True)
    want = oracle_serial(mesh, dim)
    for sym in (True, False):
        ctx, _ = _ctx(mesh, slots=400, sym=sym)
        assert _run(ctx, mesh, dim, dc.MODE_DETERMINISTIC).tobytes() == want.tobytes()
        assert _rel_err(_run(ctx, mesh, dim, dc.MODE_ATOMIC), want) <= REL_TOL


@pytest.mark.parametrize("n,dim", [(12, 3), (20, 1)])
def test_geometric_tree_bit_exact(tmp_path, n, dim):
    """Trees cut by coordinate bisection (the at-scale builder) feed the same device path."""
    mesh = prepare_mesh(MineLib(0, geometric=True), n, tmp_path)
    ctx, _ = _ctx(mesh, slots=288)
    assert _run(ctx, mesh, dim, dc.MODE_DETERMINISTIC).tobytes() == oracle_serial(mesh, dim).tobytes()


def test_plan_without_tree_is_still_bit_exact(tmp_path):
    mesh = prepare_mesh(MineLib(0), 10, tmp_path)
    ctx, _ = _ctx(mesh, slots=400, with_tree=False, sym=False)
    for dim in (1, 3):
        assert _run(ctx, mesh, dim, dc.MODE_DETERMINISTIC).tobytes() == oracle_serial(mesh, dim).tobytes()


def test_host_array_entry_point_and_coordinate_update(tmp_path):
    """The call a reference user makes: host values array out, new coordinates in."""
    mesh = prepare_mesh(MineLib(0), 8, tmp_path)
    ctx, _ = _ctx(mesh)
    out = np.full(mesh["nnz"] * 9, np.nan)
    ctx.assemble_to_host(dc.OP_ELASTICITY, [LAMBDA, MU], dc.MODE_DETERMINISTIC, out)
    assert out.tobytes() == oracle_serial(mesh, 3).tobytes()
    moved = dict(mesh)
    moved["coord"] = np.ascontiguousarray(mesh["coord"] * 1.25 + 0.1)
    ctx.update_coords(moved["coord"])
    ctx.assemble_to_host(dc.OP_ELASTICITY, [LAMBDA, MU], dc.MODE_DETERMINISTIC, out)
    assert out.tobytes() == oracle_serial(moved, 3).tobytes()
    assert ctx.launch_count() == 2


def test_repeatable_and_physical_properties_at_larger_size(tmp_path):
    """Size-independent properties on a mesh the oracle does not need to finish: run-to-run
    bitwise repeatability, zero row sums of the Laplacian (constants are in its kernel),
    block symmetry K_ij = K_ji^T, rigid translations in the kernel of elasticity."""
    n = 48
    mesh = prepare_mesh(MineLib(4), n, tmp_path)
    ctx, _ = _ctx(mesh)
    a = _run(ctx, mesh, 1, dc.MODE_DETERMINISTIC)
    b = _run(ctx, mesh, 1, dc.MODE_DETERMINISTIC)
    assert a.tobytes() == b.tobytes()
    row, col = mesh["row"], mesh["col"]
    rows = np.repeat(np.arange(mesh["N"]), np.diff(row))
    rowsum = np.zeros(mesh["N"])
    np.add.at(rowsum, rows, a)
    assert np.abs(rowsum).max() <= 1e-10 * np.abs(a).max()
    import scipy.sparse as sp
    A = sp.csr_matrix((a, col, row), shape=(mesh["N"], mesh["N"]))
    assert abs(A - A.T).max() == 0.0                       # bitwise symmetric
    e = _run(ctx, mesh, 3, dc.MODE_DETERMINISTIC).reshape(-1, 3, 3)
    for comp in range(3):
        acc = np.zeros((mesh["N"], 3))
        np.add.at(acc, rows, e[:, :, comp])
        assert np.abs(acc).max() <= 1e-10 * np.abs(e).max()
    fast = _run(ctx, mesh, 3, dc.MODE_ATOMIC)
    assert _rel_err(fast, e.reshape(-1)) <= REL_TOL


def test_two_logical_ranks_on_one_gpu(tmp_path):
    """Multi-GPU path with fewer GPUs than ranks: both slabs are assembled on this GPU one after
    the other and the interface reduce runs on-device with the pack / unpack-add kernels; the
    result is compared with the oracle on the merged mesh."""
    import torch
    from dc_lib_b200 import multigpu
    from test_multirank_gloo import _global_reference
    n, world, dim = 6, 2, 3
    slabs, vals, ctxs = [], [], []
    for r in range(world):
        s = multigpu.build_slab(n, r, world, builder="geometric")
        plan = dc.HostPlan(s["conn"], s["N"], s["row"], s["col"], 0, s["rec"], 288)
        ctx = dc.Context(0)
        ctx.set_mesh(s["conn"], s["coord"], s["row"], s["col"], 0)
        ctx.upload_plan(plan)
        v = torch.empty(s["nnz"] * 9, dtype=torch.float64, device="cuda:0")
        ctx.assemble(dc.OP_ELASTICITY, [LAMBDA, MU], dc.MODE_DETERMINISTIC, v)
        assert v.cpu().numpy().tobytes() == oracle_serial(s, 3).tobytes()        # local partial sums
        slabs.append(s); vals.append(v); ctxs.append(ctx)
    bufs = {}
    for r in range(world):
        for nb, edges in slabs[r]["intf"].items():
            e = torch.from_numpy(edges).cuda()
            b = torch.empty(edges.shape[0] * 9, dtype=torch.float64, device="cuda:0")
            dc.pack_edges(vals[r], e, 9, b)
            bufs[(r, nb)] = b
    for r in range(world):
        for nb, edges in slabs[r]["intf"].items():
            dc.unpack_add_edges(vals[r], torch.from_numpy(edges).cuda(), 9, bufs[(nb, r)])
    torch.cuda.synchronize()
    row, col, want = _global_reference(n, world, dim)
    N_glob = row.shape[0] - 1
    key_ref = np.repeat(np.arange(N_glob, dtype=np.int64), np.diff(row)) * N_glob + col
    for r in range(world):
        s = slabs[r]
        rows = np.repeat(np.arange(s["N"]), np.diff(s["row"]))
        key = s["gid_of_new"][rows] * N_glob + s["gid_of_new"][s["col"]]
        pos = np.searchsorted(key_ref, key)
        got = vals[r].cpu().numpy().reshape(-1, 9)
        assert np.abs(got - want[pos]).max() <= REL_TOL * np.abs(want).max()


def test_metis_split_irregular_mesh_on_one_gpu(tmp_path):
    """BASELINE config 4 at test size: a graded / relabelled / shuffled mesh cut by DC_partition_mesh
    (METIS) into four parts; every part is assembled on this GPU by the unit kernel (bit-exact against
    the oracle on the part), the interface partial sums are combined on the device with the pack /
    unpack-add kernels, and every local entry then equals the single-domain oracle within 1e-12."""
    import torch
    from dc_lib_b200 import multigpu
    from test_multirank_gloo import _irregular_mesh
    n, world, dim = 8, 4, 3
    conn_g, coord_g = _irregular_mesh(n)
    N_glob = coord_g.shape[0]
    part = dc.DC_partition_mesh(conn_g, N_glob, world)
    assert np.bincount(part, minlength=world).min() > 0
    parts, vals, ctxs = [], [], []
    for r in range(world):
        p = multigpu.build_part(conn_g, coord_g, part, r, world, builder="metis")
        plan = dc.HostPlan(p["conn"], p["N"], p["row"], p["col"], 0, p["rec"], 320)
        ctx = dc.Context(0)
        ctx.set_mesh(p["conn"], p["coord"], p["row"], p["col"], 0)
        ctx.upload_plan(plan)
        v = torch.empty(p["nnz"] * 9, dtype=torch.float64, device="cuda:0")
        ctx.assemble(dc.OP_ELASTICITY, [LAMBDA, MU], dc.MODE_DETERMINISTIC, v)
        assert v.cpu().numpy().tobytes() == oracle_serial(p, 3).tobytes()        # local partial sums, bit for bit
        parts.append(p); vals.append(v); ctxs.append(ctx)
    bufs = {}
    for r in range(world):
        for nb, edges in parts[r]["intf"].items():
            b = torch.empty(edges.shape[0] * 9, dtype=torch.float64, device="cuda:0")
            dc.pack_edges(vals[r], torch.from_numpy(edges).cuda(), 9, b)
            bufs[(r, nb)] = b
    torch.cuda.synchronize()
    for r in range(world):
        for nb, edges in parts[r]["intf"].items():
            assert bufs[(nb, r)].numel() == edges.shape[0] * 9
            dc.unpack_add_edges(vals[r], torch.from_numpy(edges).cuda(), 9, bufs[(nb, r)])
    torch.cuda.synchronize()
    row, col = dc.build_csr(conn_g, N_glob)
    mesh = dict(conn=conn_g, coord=coord_g, row=row, col=col, E=conn_g.shape[0], N=N_glob, nnz=int(row[-1]))
    want = oracle_serial(mesh, dim).reshape(-1, 9)
    key_ref = np.repeat(np.arange(N_glob, dtype=np.int64), np.diff(row)) * N_glob + col
    for r in range(world):
        p = parts[r]
        rows = np.repeat(np.arange(p["N"]), np.diff(p["row"]))
        key = p["gid_of_new"][rows] * N_glob + p["gid_of_new"][p["col"]]
        pos = np.searchsorted(key_ref, key)
        assert np.array_equal(key_ref[pos], key)
        got = vals[r].cpu().numpy().reshape(-1, 9)
        assert np.abs(got - want[pos]).max() <= REL_TOL * np.abs(want).max()


def test_device_resident_setup_path(tmp_path):
    """setup_device_slabs (mesh + schedule adopted from device tensors, coordinates permuted by
    the CUDA permute kernel) gives the same bits as the host-upload path."""
    import torch
    from dc_lib_b200 import multigpu
    ctx, info = multigpu.setup_device_slabs(8, 0, 1, "cuda:0", builder="geometric", slots=288)
    v = torch.empty(info["nnz"] * 9, dtype=torch.float64, device="cuda:0")
    ctx.assemble(dc.OP_ELASTICITY, [LAMBDA, MU], dc.MODE_DETERMINISTIC, v)
    torch.cuda.synchronize()
    s = multigpu.build_slab(8, 0, 1, builder="geometric")
    assert v.cpu().numpy().tobytes() == oracle_serial(s, 3).tobytes()


def _custom_ops():
    import ctypes
    path = os.path.join(os.path.dirname(__file__), "custom_op", "libcustom_ops.so")
    if not os.path.exists(path):
        pytest.fail("tests/custom_op/libcustom_ops.so missing: run __graft_entry__.build()")
    return ctypes.CDLL(path)


@pytest.mark.parametrize("sym", [True, False])
def test_user_defined_symmetric_operator(tmp_path, sym):
    """Device-operator hook: a mass matrix written in a user .cu file (KeOperator convenience
    form), run through both schedules and all three modes."""
    import ctypes
    import torch
    ops = _custom_ops()
    mesh = prepare_mesh(MineLib(0), 10, tmp_path)
    want = oracle_serial(mesh, 1, op_kind=1)
    ctx, _ = _ctx(mesh, slots=300, lists=True, sym=sym)
    launcher = ctypes.cast(ops.user_mass_op, ctypes.c_void_p)
    for mode in (dc.MODE_DETERMINISTIC, dc.MODE_LIST, dc.MODE_ATOMIC):
        v = torch.full((mesh["nnz"],), float("nan"), dtype=torch.float64, device="cuda:0")
        ctx.assemble_with(launcher, None, mode, v)
        torch.cuda.synchronize()
        got = v.cpu().numpy()
        if mode == dc.MODE_ATOMIC:
            assert _rel_err(got, want) <= REL_TOL
        else:
            assert got.tobytes() == want.tobytes()


def test_user_defined_non_symmetric_operator(tmp_path):
    """A non-symmetric operator (advection) must use the general schedule; the symmetric one
    refuses it loudly."""
    import ctypes
    import torch
    ops = _custom_ops()
    w = (0.3, -1.1, 0.7)
    mesh = prepare_mesh(MineLib(4), 10, tmp_path)
    want = oracle_serial(mesh, 1, op_kind=2, w=w)
    launcher = ctypes.cast(ops.user_advection_op, ctypes.c_void_p)
    ctx, _ = _ctx(mesh, slots=300, sym=False)
    v = torch.full((mesh["nnz"],), float("nan"), dtype=torch.float64, device="cuda:0")
    ctx.assemble_with(launcher, list(w), dc.MODE_DETERMINISTIC, v)
    torch.cuda.synchronize()
    assert v.cpu().numpy().tobytes() == want.tobytes()
    ctx2, _ = _ctx(mesh, slots=300, sym=True)
    with pytest.raises(RuntimeError, match="non-symmetric"):
        ctx2.assemble_with(launcher, list(w), dc.MODE_DETERMINISTIC, v)


def test_reference_facing_device_entry_points(tmp_path):
    """DC_device_set_mesh / DC_assembly_device through the Fortran ABI (include/DC_device.h): the call
    a DC-lib user makes after the switch; values land in the host array, bit-identical to the oracle."""
    import ctypes
    from ctypes import byref, c_int
    mesh = prepare_mesh(MineLib(0), 10, tmp_path)       # the tree stays in the library after store
    dc.DC_read_tree(mesh["tree_path"], mesh["E"], mesh["N"])
    L = dc.lib()
    vp = ctypes.c_void_p
    for name in ("dc_device_init_", "dc_device_set_mesh_", "dc_assembly_device_", "dc_device_update_coords_",
                 "dc_device_set_mode_", "dc_device_finalize_"):
        getattr(L, name).restype = None
    L.dc_device_init_(byref(c_int(0)))
    L.dc_device_set_mesh_(vp(mesh["coord"].ctypes.data), vp(mesh["conn"].ctypes.data), vp(mesh["row"].ctypes.data),
                          vp(mesh["col"].ctypes.data), byref(c_int(mesh["E"])), byref(c_int(mesh["N"])), byref(c_int(0)))
    prm = np.array([LAMBDA, MU])
    for op, dim, p, nb in ((dc.OP_LAPLACIAN, 1, None, 0), (dc.OP_ELASTICITY, 3, vp(prm.ctypes.data), 2)):
        out = np.full(mesh["nnz"] * dim * dim, np.nan)
        L.dc_assembly_device_(byref(c_int(op)), p, byref(c_int(nb)), vp(out.ctypes.data), byref(c_int(dim)))
        assert out.tobytes() == oracle_serial(mesh, dim).tobytes()
    # atomic mode through the same entry point, then new coordinates
    L.dc_device_set_mode_(byref(c_int(1)))
    out = np.full(mesh["nnz"], np.nan)
    L.dc_assembly_device_(byref(c_int(dc.OP_LAPLACIAN)), None, byref(c_int(0)), vp(out.ctypes.data), byref(c_int(1)))
    assert _rel_err(out, oracle_serial(mesh, 1)) <= REL_TOL
    L.dc_device_set_mode_(byref(c_int(0)))
    moved = dict(mesh, coord=np.ascontiguousarray(mesh["coord"] * 1.25))
    L.dc_device_update_coords_(vp(moved["coord"].ctypes.data))
    L.dc_assembly_device_(byref(c_int(dc.OP_LAPLACIAN)), None, byref(c_int(0)), vp(out.ctypes.data), byref(c_int(1)))
    assert out.tobytes() == oracle_serial(moved, 1).tobytes()
    L.dc_device_finalize_()


def test_user_cpp_program_on_the_device_path(tmp_path):
    """tests/dropin/device_driver.cc: a C++ user program (DC.h + DC_device.h) built with g++ and linked
    to libdc_b200.so; its own host-callback traversal agrees with DC_assembly_device within 1e-12."""
    import os
    import subprocess
    from common import ROOT
    exe = os.path.join(str(tmp_path), "device_driver")
    libdir, libfile = os.path.split(dc.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "dropin", "device_driver.cc"), "-o", exe, "-L" + libdir,
                           "-l:" + libfile, "-Wl,-rpath," + libdir, "-fopenmp"])
    out = subprocess.check_output([exe, "12", os.path.join(str(tmp_path), "dev")]).decode()
    vals = dict(kv.split("=") for kv in out.split() if "=" in kv)
    assert float(vals["device_vs_host_rel"]) <= REL_TOL and float(vals["elasticity_rowsum_rel"]) <= REL_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [1, 3])
def test_spmv_on_the_assembled_matrix(tmp_path, dim):
    """dc_cuda_spmv (the consumer step after the assembly): y = A x with the block-CSR values left
    on the device, against a float64 numpy evaluation of the same sums (tolerance 1e-13 of the
    largest term: the device uses one fma per entry); constants are in the kernel of both operators."""
    import torch
    mesh = prepare_mesh(MineLib(0), 10, tmp_path)
    plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, tree_records(mesh), 400)
    ctx = dc.Context(0)
    ctx.set_mesh(mesh["conn"], mesh["coord"], mesh["row"], mesh["col"], 0)
    ctx.upload_plan(plan)
    N, nnz = mesh["N"], mesh["nnz"]
    op, prm = (dc.OP_ELASTICITY, [LAMBDA, MU]) if dim == 3 else (dc.OP_LAPLACIAN, None)
    vals = torch.empty(nnz * dim * dim, dtype=torch.float64, device="cuda:0")
    ctx.assemble(op, prm, dc.MODE_DETERMINISTIC, vals)
    rng = np.random.default_rng(7)
    x = rng.standard_normal((N, dim))
    y = torch.full((N * dim,), float("nan"), dtype=torch.float64, device="cuda:0")
    ctx.spmv(dim, vals, torch.from_numpy(x).cuda(), y)
    torch.cuda.synchronize()
    v = vals.cpu().numpy().reshape(nnz, dim, dim)
    rows = np.repeat(np.arange(N), np.diff(mesh["row"]))
    want = np.zeros((N, dim))
    np.add.at(want, rows, np.einsum("kij,kj->ki", v, x[mesh["col"]]))
    got = y.cpu().numpy().reshape(N, dim)
    scale = np.abs(v).max() * np.abs(x).max() * np.diff(mesh["row"]).max() * dim
    assert np.abs(got - want).max() <= 1e-13 * scale
    ones = torch.ones(N * dim, dtype=torch.float64, device="cuda:0")
    ctx.spmv(dim, vals, ones, y)
    torch.cuda.synchronize()
    assert float(y.abs().max()) <= 1e-10 * np.abs(v).max()            # K * (rigid translation) = 0
    ctx.close()


def test_more_halo_nodes_than_threads(tmp_path):
    """Large units on small CTAs: a unit's halo-node list is longer than the CTA (832-slot units,
    128 threads), so the coordinates of the next unit are copied by the per-thread asynchronous copies
    AND by the chunk loop of the warps that run out of tasks; several units per resident CTA."""
    mesh = prepare_mesh(MineLib(0, geometric=True), 32, tmp_path)
    ctx, plan = _ctx(mesh, slots=832)
    assert plan.c.maxHalo > 128 and plan.c.nbUnits > 2 * 148
    ctx.set_option("unit_threads", 128)
    for dim in (3, 1):
        assert _run(ctx, mesh, dim, dc.MODE_DETERMINISTIC).tobytes() == oracle_serial(mesh, dim).tobytes()


def test_device_resident_plan_with_hub_rows(tmp_path):
    """dc_cuda_set_plan_device + dc_cuda_set_plan_big_device: the schedule arrays adopted from device memory (the
    path multi-GPU runs use after receiving the schedule from another rank), on a hub mesh whose high-valence
    rows exceed the unit budget and go through the fallback lists: bit-identical to the oracle."""
    import torch
    mesh = prepare_mesh(MineLib(0), 18, tmp_path, hub=True)
    plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, tree_records(mesh), 100)
    arr = plan.arrays()
    assert arr["bigEdges"] > 0
    dev = "cuda:0"
    t = {k: torch.from_numpy(np.ascontiguousarray(arr[k])).to(dev) for k in ("units", "tasks", "tgt", "haloNodes", "bigEdge", "bigPtr", "bigEdgeT")}
    t["items"] = torch.from_numpy(arr["items"].view(np.int16).copy()).to(dev)
    t["slotNodes"] = torch.from_numpy(arr["slotNodes"].view(np.int16).copy()).to(dev)
    t["bigItem"] = torch.from_numpy(arr["bigItem"].view(np.int32).copy()).to(dev)
    t["diagRel"] = torch.from_numpy(plan.diag_rel(mesh["N"]).copy()).to(dev)
    for k in ("nbUnits", "maxSlots", "maxHalo", "maxNodes", "maxTasks", "maxItems", "maxTgt", "symmetric", "bigEdges"):
        t[k] = arr[k]
    d_coord = torch.zeros((mesh["N"] + 2, 3), dtype=torch.float64, device=dev)
    d_coord[:mesh["N"]] = torch.from_numpy(mesh["coord"]).to(dev)
    ctx = dc.Context(0)
    ctx.set_mesh_device(torch.from_numpy(mesh["conn"]).to(dev), d_coord[:mesh["N"]], torch.from_numpy(mesh["row"]).to(dev),
                        torch.from_numpy(mesh["col"]).to(dev), mesh["nnz"], 0)
    ctx.set_plan_device(t)
    for dim, op, prm in ((1, dc.OP_LAPLACIAN, None), (3, dc.OP_ELASTICITY, [LAMBDA, MU])):
        vals = torch.full((mesh["nnz"] * dim * dim,), float("nan"), dtype=torch.float64, device=dev)
        ctx.assemble(op, prm, dc.MODE_DETERMINISTIC, vals)
        torch.cuda.synchronize()
        assert vals.cpu().numpy().tobytes() == oracle_serial(mesh, dim).tobytes()


@pytest.mark.parametrize("key,dim", [("kuhn64_vec0", 1), ("hub64_vec0", 1), ("hub64_vec0", 3)])
def test_parity_at_64_cubed_against_reference_golden(tmp_path, key, dim):
    """1.5 M tets: the values of the reference's own traversal over the reference's own tree (golden md5) ==
    GPU deterministic mode, bit for bit; hub64 is BASELINE config C5 (high-valence mesh, nodes of valence
    ~100): the order-free red.global mode agrees within 1e-12."""
    g = GOLD[key]
    mesh = prepare_mesh(MineLib(g["vec"]), g["n"], tmp_path, hub=g["hub"])
    ctx, _ = _ctx(mesh, slots=512)
    got = _run(ctx, mesh, dim, dc.MODE_DETERMINISTIC)
    assert md5(got) == g[f"values_md5_d{dim}"]
    if g["hub"]:
        fast = _run(ctx, mesh, dim, dc.MODE_ATOMIC)
        assert _rel_err(fast, got) <= REL_TOL


def test_parity_at_100_cubed_metis_tree_against_the_oracle(tmp_path):
    """6 M tets with the tree the reference would build (METIS node partitions, capacity 200): GPU
    deterministic elasticity == the oracle's serial ascending-element loop (= the reference's traversal
    order, tests/test_oracle_vs_reference.py), bit for bit."""
    mesh = prepare_mesh(MineLib(0), 100, tmp_path)
    ctx, plan = _ctx(mesh, slots=512)
    got = _run(ctx, mesh, 3, dc.MODE_DETERMINISTIC)
    want = oracle_serial(mesh, 3)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("dim", [1, 3])
def test_block_jacobi_of_the_assembled_matrix(tmp_path, dim):
    """dc_cuda_block_jacobi (second half of the consumer step): the inverse of the
    diagonal block of every row of the matrix the assembly left on the device; block * inverse = identity
    within 1e-12 (elasticity and Laplacian diagonal blocks are well conditioned), no bad rows."""
    import torch
    mesh = prepare_mesh(MineLib(0), 10, tmp_path)
    plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, tree_records(mesh), 400)
    ctx = dc.Context(0)
    ctx.set_mesh(mesh["conn"], mesh["coord"], mesh["row"], mesh["col"], 0)
    ctx.upload_plan(plan)
    N, nnz = mesh["N"], mesh["nnz"]
    op, prm = (dc.OP_ELASTICITY, [LAMBDA, MU]) if dim == 3 else (dc.OP_LAPLACIAN, None)
    vals = torch.empty(nnz * dim * dim, dtype=torch.float64, device="cuda:0")
    ctx.assemble(op, prm, dc.MODE_DETERMINISTIC, vals)
    inv = torch.full((N * dim * dim,), float("nan"), dtype=torch.float64, device="cuda:0")
    assert ctx.block_jacobi(dim, vals, inv) == 0
    v = vals.cpu().numpy().reshape(nnz, dim, dim)
    rows = np.repeat(np.arange(N), np.diff(mesh["row"]))
    diag = v[np.flatnonzero(mesh["col"] == rows)]
    assert diag.shape[0] == N
    got = inv.cpu().numpy().reshape(N, dim, dim)
    err = np.abs(np.einsum("nij,njk->nik", diag, got) - np.eye(dim)).max()
    assert err <= 1e-12
    ctx.close()
