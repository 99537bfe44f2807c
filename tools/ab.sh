#!/bin/bash
# usage: ab.sh <lib_base> <args...>  -> alternates current / base three times
BASE=$1; shift
for i in 1 2 3; do
  python bench.py --no-cpu-baseline --no-closed-loop "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('current', round(d['ms_per_step']*1e3,2), round(d['latency_us']['p50'],2))"
  MPPI_B200_LIB=$BASE python bench.py --no-cpu-baseline --no-closed-loop "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('base   ', round(d['ms_per_step']*1e3,2), round(d['latency_us']['p50'],2))"
done
