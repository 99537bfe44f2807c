L=husky-rover-mppi-isaacsim_b200
run() { env $3 MPPI_B200_LIB=$L/$2 python tools/timeline.py 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); r=d['rollout_us_per_block']; p=d['phases_us']
print('$1', 'event', round(d['event_us']['median'],2), 'step_ns', round(r['per_step_ns_median'],1), 'rollout', round(r['min'],2), round(r['median'],2), round(r['max'],2), 'start', p['rollout_start']['last'], 'end_last', p['rollout_end']['last'], 'hdr_last', p.get('header_published',p.get('partial_published'))['last'], 'cmd', p.get('updater_g_cmd', p.get('g_cmd')), 'done', p.get('updater_update_done', p.get('update_done')))"
}
for i in 1 2; do
run r1 libmppi_b200_r1.so A=1
run dry_wheels libmppi_b200.so A=1
run dry_obst libmppi_b200_dryobst.so A=1
run dry_chain libmppi_b200_drychain.so A=1
run dry_filter libmppi_b200_dryfilter.so A=1
done
for v in "" _dryobst _drychain _dryfilter _r1; do MPPI_B200_LIB=$L/libmppi_b200$v.so python bench.py --no-cpu-baseline --no-closed-loop --no-extras --steps 300 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('lib$v', round(d['ms_per_step']*1e3,2), round(d['latency_us']['p50'],2), round(d['e2e']['p50_us'],2), round(d['warm_l2']['p50_us'],2))"; done
