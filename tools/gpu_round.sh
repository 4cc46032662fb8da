#!/bin/bash
# One gpurun call: GPU parity tests, smoke, 1-GPU bench (both arms), ncu launch list, ncu --set full of the solver kernels.
# usage: tools/gpu_round.sh <tag>   (outputs under gpurun_out/<tag>_*)
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/${TAG}_clocks.csv 2>/dev/null &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $O/${TAG}_smoke.log
timeout 600 python bench.py --steps 100 --warmup 10 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 600 python bench.py --workload sand_pile_4m --steps 40 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_sand.json 2> $O/${TAG}_bench_sand.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fluid|k_build|k_table|k_scatter|k_scan|k_predict' -s 40 -c 12 \
    -o $O/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_full.log 2>&1
tail -3 $O/${TAG}_pytest.log; tail -2 $O/${TAG}_smoke.log; cat $O/${TAG}_bench.json | cut -c1-600
