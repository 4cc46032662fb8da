#!/bin/bash
# tuning sweep: bench.py (device-timed only) once per liblgpu.so variant given as arguments (directory names under lustrine_b200/)
# usage: tools/sweep_libs.sh <tag> lib lib_t256s3072 ...   -> gpurun_out/<tag>_sweep.txt
TAG=$1; shift
O=gpurun_out; mkdir -p $O
: > $O/${TAG}_sweep.txt
for d in "$@"; do
  LGPU_LIB=$PWD/lustrine_b200/$d/liblgpu.so timeout 300 python bench.py --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline --no-e2e ${BENCH_ARGS} > $O/${TAG}_$d.json 2> $O/${TAG}_$d.err
  python - "$d" $O/${TAG}_$d.json >> $O/${TAG}_sweep.txt <<'PY'
import json, sys
try:
    d = [json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")][-1]
    ph = d["roofline_step"]["phases_ms"]
    print("%-16s ms/step %.4f  %s  ovf %s" % (sys.argv[1], d["ms_per_step"], " ".join("%s=%.3f" % (k[:10], v) for k, v in ph.items()), d["config"].get("table_overflows")))
except Exception as e:
    print("%-16s FAILED %r" % (sys.argv[1], e))
PY
done
cat $O/${TAG}_sweep.txt
