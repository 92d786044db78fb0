"""Device permutation / renumber / reset kernels (replace src/permutations.cc and the
callback-side CSR reset) against the oracle's literal restatements."""
import numpy as np
import pytest

import dc_lib_b200 as dc
from common import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,dim", [(0, 3), (1, 3), (1000, 3), (77777, 3), (5000, 1), (4097, 5)])
def test_permute_double(n, dim):
    import torch
    rng = np.random.default_rng(n + dim)
    perm = rng.permutation(n).astype(np.int32)
    tab = rng.standard_normal((n, dim))
    want = tab.copy()
    oracle().dco_permute_double_2d(want.ctypes.data, perm.ctypes.data, n, dim)
    d_in = torch.from_numpy(tab).cuda()
    d_out = torch.full_like(d_in, float("nan"))
    dc.permute_double_2d(d_in, d_out, torch.from_numpy(perm).cuda(), dim)
    torch.cuda.synchronize()
    assert d_out.cpu().numpy().tobytes() == want.tobytes()
    # the same permutation told by destination (inverse permutation, whole-sector writes)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(n, dtype=np.int32)
    d_out2 = torch.full_like(d_in, float("nan"))
    dc.gather_double_2d(d_in, d_out2, torch.from_numpy(inv).cuda(), dim)
    torch.cuda.synchronize()
    assert d_out2.cpu().numpy().tobytes() == want.tobytes()


@pytest.mark.parametrize("n,dim", [(1, 4), (12345, 4), (999, 1), (3000, 3), (100000, 4)])
def test_permute_int_and_renumber(n, dim):
    import torch
    rng = np.random.default_rng(n * 7 + dim)
    perm = rng.permutation(n).astype(np.int32)
    tab = rng.integers(1, n + 1, size=(n, dim)).astype(np.int32)
    want = tab.copy()
    oracle().dco_permute_int_2d(want.ctypes.data, perm.ctypes.data, n, dim, 0)
    d_in = torch.from_numpy(tab).cuda()
    d_out = torch.zeros_like(d_in)
    d_perm = torch.from_numpy(perm).cuda()
    dc.permute_int_2d(d_in, d_out, d_perm, dim)
    torch.cuda.synchronize()
    assert d_out.cpu().numpy().tobytes() == want.tobytes()
    inv = np.empty_like(perm)
    inv[perm] = np.arange(n, dtype=np.int32)
    d_out2 = torch.zeros_like(d_in)
    dc.gather_int_2d(d_in, d_out2, torch.from_numpy(inv).cuda(), dim)
    torch.cuda.synchronize()
    assert d_out2.cpu().numpy().tobytes() == want.tobytes()
    ren = want.reshape(-1).copy()
    oracle().dco_renumber(ren.ctypes.data, perm.ctypes.data, ren.size, 1)
    dc.renumber(d_out, d_perm, isFortran=True)
    torch.cuda.synchronize()
    assert d_out.cpu().numpy().reshape(-1).tobytes() == ren.tobytes()
    if ren.size > 9:          # a view that starts off a 16-byte boundary: scalar head and tail of the vector kernel
        flat = torch.from_numpy(want.reshape(-1).copy()).cuda()
        part = flat[3:ren.size - 2]
        dc.renumber(part, d_perm, isFortran=True)
        torch.cuda.synchronize()
        assert flat.cpu().numpy()[3:ren.size - 2].tobytes() == ren[3:ren.size - 2].tobytes()
        assert flat.cpu().numpy()[:3].tobytes() == want.reshape(-1)[:3].tobytes()


@pytest.mark.parametrize("first,last,per", [(0, 0, 1), (3, 10, 9), (5, 4, 9), (1, 100001, 1), (7, 7777, 9)])
def test_reset_edge_interval(first, last, per):
    import torch
    total = (max(last, first) + 5) * per
    v = torch.ones(total, dtype=torch.float64, device="cuda")
    dc.reset_edges(v, first, last, per)
    torch.cuda.synchronize()
    want = np.ones(total)
    if last >= first:
        want[first * per:(last + 1) * per] = 0.0
    assert np.array_equal(v.cpu().numpy(), want)
