#!/bin/bash
# A/B of environment toggles on the default bench: tools/ab.sh <tag> "ENV=.. ENV=.." "ENV=.." ...
TAG=$1; shift
O=gpurun_out; mkdir -p $O; : > $O/${TAG}_ab.txt
n=0
for e in "$@"; do
  n=$((n+1))
  env $e timeout 300 python bench.py --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline --no-e2e ${BENCH_ARGS} > $O/${TAG}_ab$n.json 2> $O/${TAG}_ab$n.err
  python - "$e" $O/${TAG}_ab$n.json >> $O/${TAG}_ab.txt <<'PY'
import json, sys
try:
    d = [json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")][-1]
    ph = d["roofline_step"]["phases_ms"]
    print("%-28s ms/step %.4f  %s  ovf %s" % (sys.argv[1], d["ms_per_step"], " ".join("%s=%.3f" % (k[:10], v) for k, v in ph.items()), d["config"].get("table_overflows")))
except Exception as e:
    print("%-28s FAILED %r" % (sys.argv[1], e))
PY
done
cat $O/${TAG}_ab.txt
