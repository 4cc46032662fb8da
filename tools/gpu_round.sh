#!/bin/bash
# One gpurun call (1 GPU): GPU parity tests, smoke, bench (both arms), ncu launch list, ncu --set full of the hot kernels.
# usage: tools/gpu_round.sh <tag> [quick]   (outputs under gpurun_out/<tag>_*; summarise here with tools/ncu_summary.py / ncu_source_stalls.py)
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "teacher_forced and 12" > $O/${TAG}_first.log 2>&1; echo "exit $?" >> $O/${TAG}_first.log
if ! grep -q "exit 0" $O/${TAG}_first.log; then tail -20 $O/${TAG}_first.log; exit 1; fi
if [ "$2" != "quick" ]; then
  timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
  timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $O/${TAG}_smoke.log
  timeout 900 python bench.py --steps 100 --warmup 10 > $O/${TAG}_bench_dam_break_1m.json 2> $O/${TAG}_bench.err
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_ref.err
else
  timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > $O/${TAG}_bench_dam_break_1m.json 2> $O/${TAG}_bench.err
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fluid|k_build|k_scatter|k_scan|k_predict|k_reorder|k_brick' -s 40 -c 12 \
    -o $O/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/${TAG}_ncu_full.log 2>&1
if [ "$2" != "quick" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sand_iteration' -s 12 -c 1 \
      -o $O/${TAG}_prof_sand -f python bench.py --workload sand_pile_4m --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/${TAG}_ncu_sand.log 2>&1
fi
tail -3 $O/${TAG}_pytest.log 2>/dev/null; tail -2 $O/${TAG}_smoke.log 2>/dev/null; cut -c1-600 $O/${TAG}_bench_dam_break_1m.json; tail -5 $O/${TAG}_bench.err
