P="python tools/prof_run.py --workload C3 --K 16384"
O=gpurun_out
$P > $O/r2_plain_lowocc.log 2>&1 && ncu --set full --clock-control none -k regex:mppi_fused_kernel -s 3 -c 1 -f -o $O/r2_prof_c3k16384_lowocc $P > $O/r2_ncu_lowocc.log 2>&1; echo "lowocc rc=$?"
MPPI_NO_LOWOCC=1 ncu --set full --clock-control none -k regex:mppi_fused_kernel -s 3 -c 1 -f -o $O/r2_prof_c3k16384_throughput $P > $O/r2_ncu_thr.log 2>&1; echo "thr rc=$?"
ls -la $O/r2_prof_c3k16384*
