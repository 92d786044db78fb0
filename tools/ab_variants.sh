#!/bin/bash
# tools/ab_variants.sh -- same-box A/B of kernel experiment builds (make -C dc-lib_b200 variant NAME=x DEFS=...):
# one bench.py line per library.   usage: bash tools/ab_variants.sh <workload> <name> [<name> ...]   ("default" = libdc_b200.so)
W=$1; shift
for v in "$@"; do
  lib=dc-lib_b200/libdc_b200.so; [ "$v" != default ] && lib=dc-lib_b200/libdc_b200_$v.so
  echo "# $v"
  DC_B200_LIB=$PWD/$lib python bench.py --workload $W --no-cpu-baseline --steps 20 --warmup 3 ${EXTRA:-} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: j = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(json.dumps({'value_G': round(j['value']/1e9, 3), 'kernel_ms': round(j['roofline']['kernel_ms'], 4), 'frac': round(j['roofline']['frac'], 4), 'setups': round(j['setup']['element_setups_per_element'], 3), 'unit_kernel': j.get('unit_kernel')}))
"
done
