"""bench.py's unit geometry on the GPU, bit for bit against the oracle (runs after the other GPU
parity tests: the file name sorts last)."""
import pytest

import dc_lib_b200 as dc
from common import MineLib, oracle_serial, prepare_mesh
from test_gpu_assembly import _ctx, _run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,slots,threads", [(3, 512, 256), (1, 1536, 512)])
def test_bench_unit_geometry_bit_exact(tmp_path, dim, slots, threads):
    """The geometry bench.py runs: coordinate-bisection tree whose leaf capacity gives ~62-node
    leaves (bench.leaf_capacity), each cut evenly into two units of at most 32 rows (one full
    diagonal task); 512 slots x 256 threads for elasticity, 1536 x 512 for the Laplacian."""
    import bench
    n = 32
    dc.DC_set_max_elem_per_part(bench.leaf_capacity(n, 62))
    try:
        mesh = prepare_mesh(MineLib(0, geometric=True), n, tmp_path)
    finally:
        dc.DC_set_max_elem_per_part(200)
    ctx, plan = _ctx(mesh, slots=slots, bank_model=1 if dim == 1 else 0)
    ctx.set_option("unit_threads", threads)
    assert plan.c.big.nbEdges == 0
    got = _run(ctx, mesh, dim, dc.MODE_DETERMINISTIC)
    assert got.tobytes() == oracle_serial(mesh, dim).tobytes()
