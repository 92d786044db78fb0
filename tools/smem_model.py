"""Shared-memory wavefront model of gather_units (CPU only): counts, per element, the data-pipe
wavefronts of every shared-memory access of the kernel from the plan itself, so that layout and
numbering choices can be compared without a GPU.

A 64-bit access is served per half-warp (16 lanes x 8 B = 128 B = one wavefront when the lanes
hit 16 different 8-byte banks); wavefronts of a half-warp = max over banks of the number of
DISTINCT addresses in that bank.
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import common
import dc_lib_b200 as dc


def wavefronts64(addr, active=None):
    """addr: [R, 16] int array of 8-byte word addresses per half-warp row; active: bool mask.
    Returns the wavefront count per row."""
    R = addr.shape[0]
    if active is None:
        active = np.ones_like(addr, dtype=bool)
    a = np.where(active, addr, -1 - np.arange(16)[None, :] * 0 - 1)
    order = np.argsort(a, axis=1, kind="stable")
    s = np.take_along_axis(a, order, axis=1)
    uniq = np.ones_like(s, dtype=bool)
    uniq[:, 1:] = s[:, 1:] != s[:, :-1]
    uniq &= s >= 0
    bank = np.where(s >= 0, s % 16, 0)
    cnt = np.zeros((R, 16), dtype=np.int32)
    rows = np.repeat(np.arange(R), 16)
    np.add.at(cnt, (rows, bank.ravel()), uniq.ravel().astype(np.int32))
    w = cnt.max(axis=1)
    return np.where(active.any(axis=1), np.maximum(w, 1), 0)


def model_unit(U, T, items, slotN, RS, NS, threads, DD):
    """Returns dict phase -> wavefronts for one unit."""
    out = dict(acc_fetch=0, acc_fetch_ideal=0, acc_items=0, build_ld=0, build_ld_ideal=0, build_st=0, stage=0)
    nSlots = int(U["own"]) + int(U["halo"])
    # ---- accumulation: per task, per iteration, lanes read recs[slot*RS + 3a + k] and [.. 3b + k]
    for t in T:
        iters, rel, kind = int(t["iters"]), int(t["itemRel"]), int(t["kind"])
        if iters == 0:
            continue
        it = items[rel:rel + iters * 32].reshape(iters, 32).astype(np.int64)
        slot, a, b = it >> 4, (it >> 2) & 3, it & 3
        out["acc_items"] += iters          # 32 x 2 B = 64 B: one wavefront per row
        halves = [(slice(0, 16)), (slice(16, 32))]
        for h in halves:
            s_, a_, b_ = slot[:, h], a[:, h], b[:, h]
            if DD == 9:
                reads = [a_] if kind == 3 else [a_, b_]
                for which in reads:
                    for k in range(3):
                        out["acc_fetch"] += int(wavefronts64(s_ * RS + 3 * which + k).sum())
                        out["acc_fetch_ideal"] += iters
            else:
                q = np.array([0, 1, 2, 3, 1, 4, 5, 6, 2, 5, 7, 8, 3, 6, 8, 9])[(a_ * 4 + b_)]
                out["acc_fetch"] += int(wavefronts64(s_ * RS + q).sum())
                out["acc_fetch_ideal"] += iters
    # ---- record build: thread tid handles slots tid, tid+THREADS (pairs); reads 12 coords per slot
    sn = slotN[:nSlots]
    for s0 in range(0, nSlots, 16):
        blk = sn[s0:s0 + 16]
        n = blk.shape[0]
        pad = np.zeros((16, 4), dtype=np.int64)
        pad[:n] = blk
        act = np.zeros(16, dtype=bool)
        act[:n] = True
        for j in range(4):
            for q in range(3):
                out["build_ld"] += int(wavefronts64((pad[:, j] * 3 + q)[None, :], act[None, :]).sum())
                out["build_ld_ideal"] += 1
        out["build_st"] += NS              # stride-RS stores: conflict free
        out["build_ld"] += 1               # slotN
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    slots = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    dim = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    # optional: tree builder ("metis" as the reference, "geometric" as the bench) and leaf capacity
    builder = sys.argv[4] if len(sys.argv) > 4 else "metis"
    if len(sys.argv) > 5:
        dc.DC_set_max_elem_per_part(int(sys.argv[5]))
    with tempfile.TemporaryDirectory() as tmp:
        mesh = common.prepare_mesh(common.MineLib(0, geometric=builder == "geometric"), n, tmp)
    rec = common.tree_records(mesh)
    bank_model = int(os.environ.get("DC_BANK_MODEL", "1" if dim == 1 else "0"))
    plan = dc.HostPlan(mesh["conn"], mesh["N"], mesh["row"], mesh["col"], 0, rec, slots, symmetric=True, bank_model=bank_model)
    A = plan.arrays()
    units = np.frombuffer(A["units"].tobytes(), dtype=np.dtype(
        [("edgeFirst", "i8"), ("itemFirst", "i8"), ("tgtFirst", "i8"), ("slotFirst", "i8"), ("hnodeFirst", "i8"),
         ("reserved_", "i8"), ("edgeCount", "i4"), ("nodeFirst", "i4"), ("nodeCount", "i4"), ("ownElemFirst", "i4"),
         ("own", "i4"), ("halo", "i4"), ("taskFirst", "i4"), ("taskCount", "i4"), ("itemCount", "i4"),
         ("tgtCount", "i4"), ("hnodeCount", "i4"), ("pad_", "i4")]))
    tasks = np.frombuffer(A["tasks"].tobytes(), dtype=np.dtype(
        [("itemRel", "u4"), ("iters", "u2"), ("kind", "u2"), ("mask", "u4"), ("first", "u4")]))
    items = A["items"]
    slotNodes = A["slotNodes"].reshape(-1, 4).astype(np.int64)
    NS, DD = (12, 9) if dim == 3 else (10, 1)
    RS = NS + 1 if NS % 2 == 0 else NS
    tot = {}
    rng = np.random.default_rng(0)
    sample = rng.choice(len(units), size=min(len(units), 200), replace=False)
    own = 0
    for u in sample:
        U = units[u]
        T = tasks[U["taskFirst"]:U["taskFirst"] + U["taskCount"]]
        it = items[U["itemFirst"]:U["itemFirst"] + U["itemCount"]]
        sn = slotNodes[U["slotFirst"]:U["slotFirst"] + U["own"] + U["halo"]]
        r = model_unit(U, T, it, sn, RS, NS, 256, DD)
        # output staging: per owned edge 2 flushes (block + transpose) for upper, 1 for diag; each flush of 32
        # blocks = STS 2*DD + LDS 2*DD (+ tgt) wavefronts
        nd = U["nodeCount"]
        upper = U["tgtCount"] - nd
        if DD > 1:
            r["stage"] = int(np.ceil(upper / 32) * 2 * (4 * DD + 2) + np.ceil(nd / 32) * (4 * DD + 2))
        for k, v in r.items():
            tot[k] = tot.get(k, 0) + v
        own += U["own"] + 0
    # elements per unit (by ownership of rows): use E/units average
    elems = mesh["E"] * len(sample) / len(units)
    print(f"n={n} slots={slots} dim={dim} units={len(units)} sampled={len(sample)} elems/unit={mesh['E']/len(units):.1f}")
    s = 0
    for k, v in tot.items():
        print(f"  {k:16s} {v/elems:7.3f} wavefronts/elem")
        if not k.endswith("ideal"):
            s += v / elems
    print(f"  {'sum':16s} {s:7.3f}")


if __name__ == "__main__":
    main()
