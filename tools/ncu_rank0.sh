#!/bin/bash
# torchrun --no-python wrapper: rank 0 runs bench.py under ncu (one kernel launch captured), the other ranks run it plainly.
# usage (inside torchrun): tools/ncu_rank0.sh <kernel regex> <skip> <out> -- bench args...
RX=$1; SKIP=$2; OUT=$3; shift 4
if [ "$LOCAL_RANK" = "0" ]; then
  exec ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c 1 -o $OUT -f python bench.py "$@"
else
  exec python bench.py "$@"
fi
