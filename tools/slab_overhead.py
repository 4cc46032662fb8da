#!/usr/bin/env python
"""Single-GPU check of what slab mode costs by itself: the same scene stepped by a plain context and by
ONE slab context that owns the whole grid (no neighbours: arena buffers and slab code paths, no pushes)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
from lustrine_b200 import lgpu, slabs

side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
domain, pos = scenes.dam_break(side)
p = lgpu.default_step_params(dt=0.01, iterations=4, literal_lambda_index=0, exact_math=0)

def phases(G, step):
    G.set_phase_timing(True)
    acc = np.zeros(9)
    for k in range(12):
        step()
        G.sync()
        if k >= 2:
            acc += [G.last_step_ms(ph) for ph in range(9)]
    return acc / 10

G = lgpu.Context(domain, capacity_sand=len(pos))
G.upload_sand(pos)
a = phases(G, lambda: G.step_fluid(p))
G.close()
V = slabs.VirtualSlabs(domain, pos, 1)
S = V.ctx[0].G
b = phases(S, lambda: V.step(1, p))
V.close()
names = ["step", "predict", "scan", "reorder", "table", "solver", "lambda", "deltap", "halo"]
print("phase      plain    slab(world=1)")
for n, x, y in zip(names, a, b):
    print("%-8s %8.3f %8.3f" % (n, x, y))
