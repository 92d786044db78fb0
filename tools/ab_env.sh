#!/bin/bash
# tools/ab_env.sh -- same-box A/B of bench.py under different environments (one line per setting).
#   usage: bash tools/ab_env.sh <workload> "VAR=1 VAR2=x" "VAR=0" ...      ("-" = no extra environment)
W=$1; shift
for e in "$@"; do
  echo "# $e"
  [ "$e" = "-" ] && e="DC_AB=none"
  env $e python bench.py --workload $W --no-cpu-baseline --steps 20 --warmup 3 ${EXTRA:-} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: j = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(json.dumps({'value_G': round(j['value']/1e9, 3), 'kernel_ms': round(j['roofline']['kernel_ms'], 4), 'frac': round(j['roofline']['frac'], 4), 'setups': round(j['setup']['element_setups_per_element'], 3), 'unit_kernel': j.get('unit_kernel')}))
"
done
