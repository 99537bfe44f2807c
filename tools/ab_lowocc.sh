#!/bin/bash
# Same-node A/B of the monolithic kernel's low-occupancy instantiation (MPPI_NO_LOWOCC=1: throughput instantiation).
B="python bench.py --no-cpu-baseline --no-extras --no-closed-loop --latency-steps 50 --workload C3 --steps 100 --variant mono"
show() { python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('$1', 'us_per_launch', round(d['ms_per_step']*1e3,2), 'e2e_p50', round(d['e2e']['p50_us'],2))"; }
for K in 12288 16384 24576 32768 40960; do
  for i in 1 2; do
    $B --K $K 2>/dev/null | show "lowocc     K=$K"
    MPPI_NO_LOWOCC=1 $B --K $K 2>/dev/null | show "throughput K=$K"
  done
done
