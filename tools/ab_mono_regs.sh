B="python bench.py --no-cpu-baseline --no-extras --no-closed-loop --latency-steps 50 --workload C3 --steps 100"
show() { python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('$1', 'us_per_launch', round(d['ms_per_step']*1e3,2))"; }
D=$PWD/husky-rover-mppi-isaacsim_b200
for K in 8192 16384 32768 65536; do
  $B --K $K --variant mono 2>/dev/null | show "swp128 K=$K"
  MPPI_B200_LIB=$D/libmppi_b200_mb1.so $B --K $K --variant mono 2>/dev/null | show "swp166 K=$K"
  MPPI_B200_LIB=$D/libmppi_b200_base.so $B --K $K --variant mono 2>/dev/null | show "base   K=$K"
done
