#!/usr/bin/env python
"""Neighbour-list length distribution of the free-running 1M dam break (tuning aid): after N substeps dump the
list lengths and the particles-per-cell histogram."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
from lustrine_b200 import lgpu
import scenes

side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
domain, sand = scenes.dam_break(side)
G = lgpu.Context(domain, capacity_sand=len(sand))
G.upload_sand(sand)
p = lgpu.default_step_params(iterations=4, exact_math=0, literal_lambda_index=0)
done = 0
for target in (1, 20, 50, 100, 200, 400):
    while done < target:
        G.step_fluid(p); done += 1
    G.sync()
    cnt = G.dump(lgpu.DUMP_NBR_COUNT)
    cs = G.dump(lgpu.DUMP_CELL_START)
    ppc = np.diff(cs)
    ppc = ppc[ppc > 0]
    q = np.percentile(cnt, [50, 90, 99, 99.9, 100])
    print("substep %4d  ms %.3f  list len mean %.1f  p50 %d p90 %d p99 %d p99.9 %d max %d  >32: %.1f%%  >48: %.1f%% >64: %.1f%%  | ppc mean %.2f p99 %d max %d  occupied cells %d"
          % (done, G.last_step_ms(0), cnt.mean(), *q, 100 * (cnt > 32).mean(), 100 * (cnt > 48).mean(), 100 * (cnt > 64).mean(), ppc.mean(), np.percentile(ppc, 99), ppc.max(), len(ppc)))

# phase split in the collapsed state
names = {1: "predict", 2: "scan", 3: "reorder", 4: "table+lambda1", 6: "lambda", 7: "deltap"}
G.set_phase_timing(True)
c0 = G.dump(lgpu.DUMP_COUNTERS)
acc = {k: 0.0 for k in names}
tot = 0.0
N = 10
for _ in range(N):
    G.step_fluid(p); G.sync()
    tot += G.last_step_ms(0)
    for k in names: acc[k] += G.last_step_ms(k)
c1 = G.dump(lgpu.DUMP_COUNTERS)
print("phases at substep %d: total %.3f ms | " % (done, tot / N) + "  ".join("%s %.3f" % (names[k], acc[k] / N) for k in names) + " | re-walked rows per substep %.0f" % ((int(c1[1]) - int(c0[1])) / N))
