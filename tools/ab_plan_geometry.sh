#!/bin/bash
# tools/ab_plan_geometry.sh -- same-box A/B of the schedule-only steps that were made without GPU time
# (profiles/README.md "Round-1f"): leaf capacity, even leaf split, rank re-deal, Laplacian bank model.
# Every line is a bench.py JSON line (device-resident value + roofline); about 40 s per line at 160^3.
#   gpurun --timeout 900 -- 'bash tools/ab_plan_geometry.sh > gpurun_out/ab_plan_geometry.jsonl 2>gpurun_out/ab_plan_geometry.err'
set -u
B="python bench.py --no-cpu-baseline --steps 20 --warmup 3"
run() { echo "# $*" ; env "$@" $B --workload "$W" $EXTRA | head -1 ; }
W=e160 EXTRA=""
run DC_AB=default                              # capacity ~182, even split, re-deal
run DC_PLAN_REGROUP=0                          # without the rank re-deal
run DC_PLAN_GREEDY_SPLIT=1                     # 62-node leaves, greedy row cut
run DC_PLAN_MAX_ROWS=32                        # no unit above 32 rows: smaller per-unit maxima, ~92 KB of shared memory per CTA
EXTRA="--leaf-nodes 0"
run DC_AB=capacity200                          # the reference's leaf capacity (the kernel numbers of round 1)
EXTRA="--slots 800 --threads 512"
run DC_AB=whole_leaf_units                     # one 62-node leaf per unit, one 512-thread CTA per SM
EXTRA="--slots 1280 --threads 512"
run DC_AB=two_leaf_units                       # the largest elasticity unit that fits 227 KB
W=l160 EXTRA=""
run DC_AB=default                              # Laplacian: bank model 1, re-deal
run DC_BENCH_BANK_MODEL=0                      # gradient bank model (round 1)
run DC_PLAN_REGROUP=0
EXTRA="--leaf-nodes 62"
run DC_AB=leaf62                               # Laplacian with 62-node leaves
