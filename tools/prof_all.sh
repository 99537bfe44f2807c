#!/bin/bash
# The ncu captures committed under profiles/ (one gpurun call; every profiled command first exits 0 without ncu).
P="python tools/prof_run.py"
O=gpurun_out
cap() { # name regex extra-ncu-flags workload...
  n=$1; k=$2; x=$3; shift 3
  $P "$@" > $O/r2_plain_$n.log 2>&1 && \
  ncu --set full --clock-control none $x -k regex:$k -s 3 -c 1 -f -o $O/r2_prof_$n $P "$@" > $O/r2_ncu_$n.log 2>&1
  echo "$n rc=$?"
}
cap c2 mppi_fused_pipe "--import-source on" --workload C2
cap c3 "mppi_fused_kernel" "" --workload C3
cap c4 "mppi_fused_kernel" "" --workload C4
cap c5 "mppi_fused_kernel" "" --workload C5
cap c5many "mppi_fused_kernel" "" --workload C5many
B="python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-closed-loop --latency-steps 10"
$B > $O/r2_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_c2.csv $B > $O/r2_ncu_bench.log 2>&1
echo "launches rc=$?"
ls -la $O/r2_prof_* $O/r2_launches_c2.csv
