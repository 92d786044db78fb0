"""Seeded sweep of small meshes through the schedule builder: tree builder (METIS / coordinate
bisection), leaf capacity, slot budget (down to budgets that push rows into the fallback lists),
colouring, symmetric or general schedule and bank model are drawn at random; every schedule must
cover every contribution once (test_plan.check_plan) and, interpreted on the CPU with the oracle's
arithmetic (test_plan_interpreter.run_schedule), reproduce the reference order bit for bit."""
import random

import pytest

import common
import dc_lib_b200 as dc
from common import MineLib, oracle_serial, prepare_mesh, tree_records
from test_plan import check_plan
from test_plan_interpreter import run_schedule


def _cases(count=14, seed=7):
    rng = random.Random(seed)
    out = []
    while len(out) < count:
        n = rng.choice([3, 4, 5, 6, 7, 8, 9])
        geo = rng.random() < 0.6
        if not geo and n < 6:
            continue          # the reference's METIS path reads uninitialised memory on <= 200-element meshes
        out.append((n, geo, rng.choice([200, 60, 90, 120, 150]), rng.choice([16, 24, 40, 64, 100, 160, 300, 512]),
                    rng.choice([0, 0, 2, 4]), rng.random() < 0.6, rng.choice([0, 1])))
    return out


@pytest.mark.parametrize("n,geometric,capacity,slots,vec,sym,bank_model", _cases())
def test_random_schedule_is_complete_and_bit_exact(tmp_path, n, geometric, capacity, slots, vec, sym, bank_model):
    dc.DC_set_max_elem_per_part(capacity)
    try:
        mesh = prepare_mesh(MineLib(vec, geometric=geometric), n, tmp_path)
    finally:
        dc.DC_set_max_elem_per_part(200)
    plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, tree_records(mesh), slots,
                       symmetric=sym, bank_model=bank_model)
    check_plan(mesh, plan, slots, sym)
    dim = 1 if bank_model else 3
    assert run_schedule(mesh, plan, dim).tobytes() == oracle_serial(mesh, dim).tobytes()
