"""Generates tests/golden/ from the UNMODIFIED reference (oracle/_ref) -- run here, where
/root/reference exists; the GPU box only sees the committed outputs.

For every case: the reference builds the tree (DC_create_tree), the driver permutes /
renumbers / finalizes exactly as a user would, the reference stores the tree file
(DC_store_tree) and the reference's own OpenMP traversal runs the oracle's callbacks.
Recorded: md5 of the tree file, of the permuted connectivity and of the value arrays;
for the smallest cases the raw tree file and the Laplacian values as well.

    python oracle/gen_golden.py
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402

CASES = [(4, 0), (6, 0), (6, 4), (8, 0), (8, 4), (10, 0), (10, 2), (10, 4), (10, 8), (16, 0), (16, 4),
         (20, 0), (20, 4), (40, 0), (40, 4), (64, 0)]
HUB_CASES = [(18, 0), (18, 4), (64, 0)]
RAW_LIMIT = 8       # raw tree file + values kept up to this cube size


def main():
    out = {}
    os.makedirs(common.GOLDEN, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        for hub, cases in ((False, CASES), (True, HUB_CASES)):
            for n, vec in cases:
                ref = common.RefLib(vec)
                mesh = common.prepare_mesh(ref, n, tmp, hub=hub)
                key = f"{'hub' if hub else 'kuhn'}{n}_vec{vec}"
                entry = dict(n=n, vec=vec, hub=hub, E=mesh["E"], N=mesh["N"], nnz=mesh["nnz"],
                             tree_md5=common.md5(mesh["tree"]), conn_md5=common.md5(mesh["conn"]),
                             coord_md5=common.md5(mesh["coord"]))
                for dim in (1, 3):
                    if n > 20 and dim == 3 and vec == 0 and not hub:
                        continue
                    vals = ref.traversal(mesh, dim)
                    entry[f"values_md5_d{dim}"] = common.md5(vals)
                    entry[f"values_sum_d{dim}"] = float(np.sum(vals))
                    if n <= RAW_LIMIT and dim == 1:
                        np.save(os.path.join(common.GOLDEN, f"{key}_values_d1.npy"), vals)
                if n <= RAW_LIMIT:
                    with open(os.path.join(common.GOLDEN, f"{key}_tree.bin"), "wb") as f:
                        f.write(mesh["tree"])
                out[key] = entry
                print(key, entry["tree_md5"], flush=True)
    with open(os.path.join(common.GOLDEN, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
