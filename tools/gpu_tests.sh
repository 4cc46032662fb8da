#!/bin/bash
# One gpurun call (1 GPU): a quick canary test first (a dead-locked pipeline must not eat the budget), then the GPU suite.
# usage: tools/gpu_tests.sh <tag> [pytest -k expression]
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "teacher_forced and 12" > $O/${TAG}_first.log 2>&1; echo "exit $?" >> $O/${TAG}_first.log
tail -12 $O/${TAG}_first.log
if ! grep -q "exit 0" $O/${TAG}_first.log; then exit 1; fi
if [ -n "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -q -k "$2" > $O/${TAG}_pytest.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1
fi
echo "pytest exit $?" >> $O/${TAG}_pytest.log
grep -E "passed|failed|error" $O/${TAG}_pytest.log | tail -5
grep -E "^FAILED|^ERROR" $O/${TAG}_pytest.log | head -40
