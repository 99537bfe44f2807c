#!/bin/bash
# The bench lines committed under profiles/r2_bench/ (one B200): python bench.py variants, one JSON line each.
O=gpurun_out/r2_bench; mkdir -p $O
python bench.py > $O/c2_strict_default.json 2> $O/c2_strict_default.err; echo "default rc=$?"
python bench.py --math fast --no-extras --no-cpu-baseline > $O/c2_fast.json 2>/dev/null
python bench.py --start rocks --no-extras --no-cpu-baseline --no-closed-loop > $O/c2_strict_rocks.json 2>/dev/null
python bench.py --dem-noise 0.02 --no-extras --no-cpu-baseline --no-closed-loop > $O/c2_strict_noisy_dem.json 2>/dev/null
python bench.py --workload C1 --no-extras --cpu-seconds 5 > $O/c1_strict.json 2>/dev/null
python bench.py --workload C3 --steps 100 --no-extras --no-cpu-baseline --no-closed-loop > $O/c3_strict.json 2>/dev/null
python bench.py --workload C3 --math fast --steps 100 --no-extras --no-cpu-baseline --no-closed-loop > $O/c3_fast.json 2>/dev/null
python bench.py --workload C4 --steps 50 > $O/c4_strict.json 2>/dev/null
python bench.py --workload C5 --steps 100 --no-extras --no-cpu-baseline --no-closed-loop > $O/c5_strict_ext.json 2>/dev/null
python bench.py --workload C5 --critics reference --steps 100 --no-extras --no-cpu-baseline --no-closed-loop > $O/c5_strict_ref4.json 2>/dev/null
python bench.py --workload C5many --steps 50 > $O/c5many_strict_ext.json 2>/dev/null
python bench.py --impl reference --steps 20 --warmup 3 > $O/c2_reference_arm.json 2>/dev/null
ls -la $O
