"""tools/store_model.py -- 128-byte lines touched by the value stores of gather_units (CPU only).

The flush of a task stores the 32 staged blocks with DD instructions per direction (block and
transpose): instruction r stores words r*32 .. r*32+31 of the staged [32 blocks][DD] array, word
idx going to block idx // DD.  A store instruction costs one data-pipe wavefront per distinct
128-byte line, so the order of the lanes inside a task decides the cost (the set of lanes of a
task, hence the padding, stays the same).  Prints lines per element for the schedule as built and
for the lanes of every task re-sorted by a key.   python tools/store_model.py [n] [slots]
"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import dc_lib_b200 as dc, bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 512
DD = 9
conn, coord, N = dc.kuhn_cube_fast(n, seed=12345)
E = conn.shape[0]
dc.DC_set_max_elem_per_part(bench.leaf_capacity(n, 62))
dc.DC_create_tree_geometric(conn, E, 4, N, coord)
dc.DC_renumber_int_array(conn.reshape(-1), 4 * E)
row, col = dc.build_csr_fast(conn, N)
dc.DC_finalize_tree(row, conn, E, 4)
rec = dc.flatten_tree()
plan = dc.HostPlan(conn, N, row, col, 0, rec, slots, symmetric=True)
A = plan.arrays()
units = np.frombuffer(A["units"].tobytes(), dtype=np.dtype(
    [("edgeFirst", "i8"), ("itemFirst", "i8"), ("tgtFirst", "i8"), ("slotFirst", "i8"), ("hnodeFirst", "i8"),
     ("reserved_", "i8"), ("edgeCount", "i4"), ("nodeFirst", "i4"), ("nodeCount", "i4"), ("ownElemFirst", "i4"),
     ("own", "i4"), ("halo", "i4"), ("taskFirst", "i4"), ("taskCount", "i4"), ("itemCount", "i4"),
     ("tgtCount", "i4"), ("hnodeCount", "i4"), ("pad_", "i4")]))
tasks = np.frombuffer(A["tasks"].tobytes(), dtype=np.dtype(
    [("itemRel", "u4"), ("iters", "u2"), ("kind", "u2"), ("mask", "u4"), ("first", "u4")]))
tgt = A["tgt"].reshape(-1, 2).astype(np.int64)
TR = np.array([(c % 3) * 3 + c // 3 for c in range(DD)])


def lines_of_task(k, kt):
    """k, kt: [32] target edges (-1 = none).  Lines touched by the 2*DD store instructions."""
    idx = np.arange(32 * DD).reshape(DD, 32)
    ent, comp = idx // DD, idx % DD
    tot = 0
    for r in range(DD):
        a = k[ent[r]]
        ad = (a * DD + comp[r]) * 8 // 128
        tot += len(np.unique(ad[a >= 0]))
        b = kt[ent[r]]
        bd = (b * DD + TR[comp[r]]) * 8 // 128
        tot += len(np.unique(bd[b >= 0]))
    return tot


rng = np.random.default_rng(0)
sample = rng.choice(len(units), size=min(len(units), 300), replace=False)
res = {}
for u in sample:
    U = units[u]
    T = tasks[U["taskFirst"]:U["taskFirst"] + U["taskCount"]]
    tg = tgt[U["tgtFirst"]:U["tgtFirst"] + U["tgtCount"]]
    for t in T:
        f = int(t["first"])
        k = np.full(32, -1, dtype=np.int64); kt = k.copy()
        m = np.array([(int(t["mask"]) >> l) & 1 for l in range(32)], dtype=bool)
        cnt = int(m.sum())
        k[:cnt] = tg[f:f + cnt, 0]; kt[:cnt] = tg[f:f + cnt, 1]
        for name, key in (("as built", None), ("by edge", lambda: k[:cnt]), ("by transposed edge", lambda: kt[:cnt]),
                          ("by col block then edge", None)):
            if name == "as built":
                o = np.arange(cnt)
            elif name == "by col block then edge":
                # group the lanes by the row of the transposed edge (= the column), 8 rows per group
                colnode = np.searchsorted(row, kt[:cnt], side="right") - 1
                o = np.lexsort((k[:cnt], colnode // 8))
            else:
                o = np.argsort(key(), kind="stable")
            kk, kk2 = k.copy(), kt.copy()
            kk[:cnt] = k[:cnt][o]; kk2[:cnt] = kt[:cnt][o]
            res[name] = res.get(name, 0) + lines_of_task(kk, kk2)
elems = E * len(sample) / len(units)
print(f"n={n} slots={slots} units={len(units)} elems/unit={E / len(units):.1f}  ideal lines/elem {2.0 * len(col) * 72 / 128 / E / 2:.2f}")
for k_, v in res.items():
    print(f"  lanes {k_:24s} {v / elems:6.3f} lines per element")

# ---- second experiment: the upper-edge lanes of a unit dealt into tasks in another order ----
items = A["items"]
res2 = {}
for u in sample:
    U = units[u]
    T = tasks[U["taskFirst"]:U["taskFirst"] + U["taskCount"]]
    tg = tgt[U["tgtFirst"]:U["tgtFirst"] + U["tgtCount"]]
    it = items[U["itemFirst"]:U["itemFirst"] + U["itemCount"]]
    pad = (int(U["own"]) + int(U["halo"])) << 4
    K, KT, C = [], [], []
    for t in T:
        if int(t["kind"]) == 3:
            continue
        f, iters, rel = int(t["first"]), int(t["iters"]), int(t["itemRel"])
        cnt = bin(int(t["mask"])).count("1")
        blk = it[rel:rel + iters * 32].reshape(iters, 32)
        K += list(tg[f:f + cnt, 0]); KT += list(tg[f:f + cnt, 1])
        C += list((blk[:, :cnt] != pad).sum(axis=0))
    K, KT, C = np.array(K), np.array(KT), np.array(C)
    rown = np.searchsorted(row, K, side="right") - 1
    coln = np.searchsorted(row, KT, side="right") - 1
    orders = {
        "count desc (as built)": np.argsort(-C, kind="stable"),
        "row-major": np.argsort(K, kind="stable"),
        "column-major": np.lexsort((rown, coln)),
        "4-row blocks, col-major inside": np.lexsort((rown, coln, (rown - rown.min()) // 4)),
        "8-row blocks, col-major inside": np.lexsort((rown, coln, (rown - rown.min()) // 8)),
        "8-row blocks, count inside": np.lexsort((-C, (rown - rown.min()) // 8)),
        "count class (>=6 / <6), row-major inside": np.lexsort((K, C < 6)),
    }
    for name, o in orders.items():
        lines = 0; padded = 0
        for b in range(0, len(o), 32):
            sel = o[b:b + 32]
            k = np.full(32, -1, dtype=np.int64); kt = k.copy()
            k[:len(sel)] = K[sel]; kt[:len(sel)] = KT[sel]
            lines += lines_of_task(k, kt)
            padded += 32 * C[sel].max()
        r = res2.setdefault(name, [0, 0, 0])
        r[0] += lines; r[1] += padded; r[2] += C.sum()
print("upper-edge lanes of a unit dealt into tasks of 32 in another order (diagonal tasks not included):")
for name, (lines, padded, useful) in res2.items():
    print(f"  {name:42s} {lines / elems:6.3f} lines/elem   item efficiency {useful / padded:5.3f}   padded iterations/elem {padded / 32 / elems:5.3f}")
