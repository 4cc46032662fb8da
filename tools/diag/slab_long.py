#!/usr/bin/env python
"""Long free-running slab run on ONE GPU (virtual ranks) against the single-context run: prints max|dx| every 50 substeps.
usage: slab_long.py [side] [world] [steps] [replan_every] [capacity_factor] [iterations]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
from lustrine_b200 import lgpu, slabs
import scenes

side = int(sys.argv[1]) if len(sys.argv) > 1 else 48
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 500
every = int(sys.argv[4]) if len(sys.argv) > 4 else 25
factor = float(sys.argv[5]) if len(sys.argv) > 5 else 1.2
K = int(sys.argv[6]) if len(sys.argv) > 6 else 4
domain, pos = scenes.dam_break(side)
if os.environ.get("WIDE"):  # a domain the splash never leaves: no predicted position outside the grid
    domain = (8 * side, 3 * side, 6 * side)
    pos = pos + np.array([0.0, 0.0, 2.0 * side], np.float32)
kw = dict(dt=0.01, iterations=K, literal_lambda_index=0, exact_math=1)
G = lgpu.Context(domain, capacity_sand=len(pos)); G.upload_sand(pos)
V = slabs.VirtualSlabs(domain, pos, world, capacity_factor=factor)
for step in range(steps):
    G.step_fluid(**kw); V.step(1, **kw)
    if every and (step + 1) % every == 0:
        if V.replan_if_needed(): print("  step %d re-plan -> %s" % (step + 1, V.counts()[0]))
    lo, hi = [int(x) for x in os.environ.get("DETAIL", "0,0").split(",")]
    if (step + 1) % 50 == 0 or step < 3 or lo <= step + 1 <= hi:
        rp, rv, _ = G.download(); sp, sv, _ = V.gather()
        print("step %4d  max|dx| %.3e  max|dv| %.3e  counters %s" % (step + 1, np.abs(sp - rp).max(), np.abs(sv - rv).max(), G.dump(lgpu.DUMP_COUNTERS)[:3]))
