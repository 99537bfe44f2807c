#!/bin/bash
# Same-node A/B of programmatic dependent launch (MPPI_NO_PDL=1 turns the launch attribute off): back-to-back steps
# (warm_l2), the device-resident closed loop, and the flushed headline (which it must not move).
B="python bench.py --no-cpu-baseline --no-extras --steps 300 --latency-steps 300"
show() { python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); c=d.get('closed_loop') or {}
print('$1', 'flushed_us', round(d['ms_per_step']*1e3,2), 'warm_l2_us', round(d['warm_l2']['ms_per_step']*1e3,2), 'e2e_p50_us', round(d['e2e']['p50_us'],2), 'closed_loop_us_per_iter', round(c.get('device_resident_us_per_iteration',0),2), 'host_driven', round(c.get('host_driven_us_per_iteration',0),2))"; }
for i in 1 2 3; do
  $B 2>/dev/null | show pdl
  MPPI_NO_PDL=1 $B 2>/dev/null | show no_pdl
done
