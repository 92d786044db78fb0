"""tools/c4_partition.py n world [partition] -- the element partition of bench.py --workload c4 (graded,
relabelled, shuffled Kuhn cube cut by DC_partition_mesh), computed on the host and kept in the directory
DC_PARTITION_CACHE (default tools/_cache), where multigpu.setup_device_parts looks for it: METIS on the
nodal graph of 10^8 elements takes minutes during which the GPUs of a bench call would sit idle."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dc_lib_b200 as dc

n, world = int(sys.argv[1]), int(sys.argv[2])
partition = sys.argv[3] if len(sys.argv) > 3 else "metis"
seed = 2024
cache = os.environ.get("DC_PARTITION_CACHE", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cache"))
dc.DC_set_num_threads(0)
t = time.time()
conn, coord, N = dc.kuhn_cube_fast(n)
conn, coord = dc.scramble_mesh(conn, coord, seed=seed)
print("mesh", round(time.time() - t, 1), "s", flush=True)
t = time.time()
part = dc.DC_partition_mesh(conn, N, world, coord if partition == "geometric" else None)
print("partition", round(time.time() - t, 1), "s", np.bincount(part).tolist(), flush=True)
os.makedirs(cache, exist_ok=True)
np.save(os.path.join(cache, f"c4_part_n{n}_w{world}_{partition}_s{seed}.npy"), part.astype(np.int8))
