#!/bin/bash
# tools/asan_host.sh -- the host library (tree, colouring, CSR helper, schedule builder) under
# AddressSanitizer + UBSan on a few small cubes; CUDA entry points stay unresolved (never called).
#   bash tools/asan_host.sh          (about a minute; prints one line per case, sanitizer reports if any)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${TMPDIR:-/tmp}/dc_asan_host
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fopenmp -w \
    -I"$ROOT/include" -I"$ROOT/dc-lib_b200/host" -I"$ROOT/dc-lib_b200/csrc" \
    "$ROOT/tools/asan_host_driver.cc" "$ROOT"/dc-lib_b200/host/*.cc \
    /usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a -Wl,--unresolved-symbols=ignore-all -o "$OUT"
# n, leaf capacity, unit slots, geometric tree, bank model, symmetric schedule
for args in "20 169 512 1 0 1" "20 169 1536 1 1 1" "16 200 64 0 0 0" "12 60 24 1 1 1" "14 200 300 0 0 1" "10 200 16 1 0 1"; do
    OMP_NUM_THREADS=4 "$OUT" $args
done
DC_VEC_SIZE=4 OMP_NUM_THREADS=8 "$OUT" 10 90 16 0 1 0
