#!/bin/bash
# Same-node A/B of a monolithic-kernel change: this build vs a baseline library (BASE=path), alternated, on the
# throughput workloads at full and at partial occupancy (K per GPU of the strong-scaling end of C3) and on C2.
BASE=${BASE:-$PWD/husky-rover-mppi-isaacsim_b200/libmppi_b200_base.so}
B="python bench.py --no-cpu-baseline --no-extras --no-closed-loop --latency-steps 50"
show() { python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('$1', 'us_per_launch', round(d['ms_per_step']*1e3,2), 'value %.4g' % d['value'])"; }
run() {  # label args...
  L=$1; shift
  for i in 1 2; do
    $B "$@" 2>/dev/null | show "new  $L"
    MPPI_B200_LIB=$BASE $B "$@" 2>/dev/null | show "base $L"
  done
}
run C3_K262144 --workload C3 --steps 60
run C3_K32768 --workload C3 --K 32768 --steps 100
run C3_K16384 --workload C3 --K 16384 --steps 100
run C4 --workload C4 --steps 40
run C5ext --workload C5 --steps 60
run C5many --workload C5many --steps 40
run C2 --steps 300 --latency-steps 300
