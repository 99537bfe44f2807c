#!/bin/bash
# Same-node A/B of library variants (make variant EXTRA=... VOUT=../libmppi_b200_<name>.so): kernel-internal timeline and
# bench means, alternating.   VARIANTS="r1 cur <name> ..." bash tools/ab_variants.sh
L=husky-rover-mppi-isaacsim_b200
run() { env MPPI_B200_LIB=$L/$2 python tools/timeline.py 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); r=d['rollout_us_per_block']; p=d['phases_us']; u=d.get('updater_cycles_since_poll_start') or {}
print('$1', 'event', round(d['event_us']['median'],2), 'end_last', p['rollout_end']['last'], 'hdr_last', p.get('header_published',p.get('partial_published'))['last'], 'min', p.get('updater_g_min', p.get('g_min')), 'cmd', p.get('updater_g_cmd', p.get('g_cmd')), 'done', p.get('updater_update_done', p.get('update_done')), 'upd_cycles min->cmd', round(u.get('command_stats',0)-u.get('block_min',0)), 'rec', round(u.get('recurrence(warp1)',0)-u.get('nominal',0)))"
}
for i in 1 2; do
for v in $VARIANTS; do
f=libmppi_b200_$v.so; [ $v = cur ] && f=libmppi_b200.so
run $v $f
done; done
for i in 1 2; do for v in $VARIANTS; do f=libmppi_b200_$v.so; [ $v = cur ] && f=libmppi_b200.so; MPPI_B200_LIB=$L/$f python bench.py --no-cpu-baseline --no-closed-loop --no-extras --steps 500 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('lib_$v', round(d['ms_per_step']*1e3,2), round(d['latency_us']['mean'],2), round(d['e2e']['p50_us'],2), round(d['warm_l2']['ms_per_step']*1e3,2))"; done; done
