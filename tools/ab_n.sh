#!/bin/bash
# A/B of environment toggles on the N-GPU bench: tools/ab_n.sh <tag> <N> "ENV=.." "ENV=.." ...
TAG=$1; N=$2; shift; shift
O=gpurun_out; mkdir -p $O; : > $O/${TAG}_ab.txt
n=0
for e in "$@"; do
  n=$((n+1))
  env $e timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $N --steps ${STEPS:-40} --warmup 5 ${BENCH_ARGS} > $O/${TAG}_ab$n.json 2> $O/${TAG}_ab$n.err
  python - "$e" $O/${TAG}_ab$n.json >> $O/${TAG}_ab.txt <<'PY'
import json, sys
try:
    d = [json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")][-1]
    ph = d["roofline_step"]["phases_ms_rank0"]
    print("%-22s ms/step %.4f  %s  e2e_ms %s" % (sys.argv[1], d["ms_per_step"], " ".join("%s=%.3f" % (k[:9], v) for k, v in ph.items()), (d.get("e2e") or {}).get("ms_per_step")))
except Exception as e:
    print("%-22s FAILED %r" % (sys.argv[1], e))
PY
done
cat $O/${TAG}_ab.txt
