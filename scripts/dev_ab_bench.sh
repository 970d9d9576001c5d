#!/bin/bash
# A/B of a tuning knob on ONE box: scripts/dev_ab_bench.sh <key> [steps]   (alternates key=0 / key=1 twice)
key=$1; steps=${2:-40}
for v in 0 1 0 1; do
  CLK_TUNING="$key=$v" python bench.py --steps $steps --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read())
print('$key=$v', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 4), 'ms/step', d['clocks']['sm_mhz'], 'MHz conv fwd+dgrad',
      round(d['roofline']['achieved'], 1), 'TF/s; all kernels serialised', round(d['roofline']['all_kernels_ms_per_step'], 3), 'ms')"
done
