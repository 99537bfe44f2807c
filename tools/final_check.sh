set -x
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) 2>&1
( time python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3 ) 2>&1
( time python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err ) 2>&1; tail -c 600 gpurun_out/final_ref.json
( time python bench.py --steps 20 --warmup 3 > gpurun_out/final_b200_s20.json 2> gpurun_out/final_b200_s20.err ) 2>&1; tail -c 300 gpurun_out/final_b200_s20.json
( time bash tools/bench_all.sh ) 2>&1 | tail -25
