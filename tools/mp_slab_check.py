#!/usr/bin/env python
"""Multi-process slab check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/mp_slab_check.py [--side 40] [--steps 8]

Every rank owns one x-slab of a dam break (real CUDA-IPC arenas between processes, stores over
NVLink); after `steps` substeps the particles are gathered in id order and rank 0 compares them with
a single-context run of the same scene on its own GPU.  Exact arithmetic: the two runs differ only by
the summation order of ghost neighbours, so the bound is the parity tolerance of the GPU tests.
Prints one JSON line and exits non-zero on a mismatch.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=40)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--kind", default="fluid", choices=["fluid"])
    ap.add_argument("--replan-every", type=int, default=0, help="free run: check the load every R substeps and re-plan the slab boundaries when it has drifted")
    ap.add_argument("--no-push", action="store_true", help="no initial sideways velocity: a plain free-running dam break")
    ap.add_argument("--capacity-factor", type=float, default=1.5)
    ap.add_argument("--wide", action="store_true", help="a domain the splash never leaves (8 x 3 x 6 sides): no predicted position outside the grid, "
                    "where a slab bins per axis and the single context wraps like the reference (DESIGN.md §6)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import scenes
    from lustrine_b200 import lgpu, slabs

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world = dist.get_world_size()

    domain, pos = scenes.dam_break(args.side)
    if args.wide:
        domain = (8 * args.side, 3 * args.side, 6 * args.side)
        pos = pos + np.array([0.0, 0.0, 2.0 * args.side], np.float32)
    vel0 = np.zeros_like(pos)
    if not args.no_push:
        vel0[:, 0] = 12.0 * np.sin(pos[:, 2])  # pushes particles across the slab boundaries: migration every substep
    solids = scenes.floor_plate(domain[0], domain[2]) if args.kind == "sand" else None
    if args.kind == "fluid":
        kw = dict(dt=0.01, iterations=4, literal_lambda_index=0, exact_math=1)
        mode = 1
    else:
        kw = dict(dt=0.016, iterations=4, exact_math=1)
        mode = 2
    params = lgpu.default_step_params(**kw)

    gw = int(os.environ.get("LGPU_GHOST_COLUMNS", "2" if mode == 1 else "1"))
    S = slabs.DistributedSlab(domain, pos, solids=solids, vel=vel0, device=local_rank, capacity_factor=args.capacity_factor, ghost_columns=gw)
    worst_imbalance = 1.0
    for k in range(args.steps):
        S.step(mode, params)
        if args.replan_every and (k + 1) % args.replan_every == 0:
            S.G.sync()
            worst_imbalance = max(worst_imbalance, slabs.imbalance(S.counts()[0]))
            S.replan_if_needed()
    S.G.sync()
    final_owned = S.counts()[0]
    info = S.G.slab_info()
    sp, sv, _ = S.gather()
    ok = True
    line = None
    if rank == 0:
        G = lgpu.Context(domain, capacity_sand=len(pos), capacity_solid=0 if solids is None else len(solids), device=local_rank)
        G.upload_sand(pos, vel0)
        if solids is not None:
            G.upload_solids(solids)
        for _ in range(args.steps):
            (G.step_fluid if mode == 1 else G.step_sand)(params)
        rp, rv, _ = G.download()
        err = float(np.abs(sp - rp).max())
        verr = float(np.abs(sv - rv).max())
        tol = 2e-4
        ok = err <= tol
        line = {"check": "mp_slab", "kind": args.kind, "world": world, "particles": len(pos), "steps": args.steps,
                "max_abs_dx_vs_single_context": err, "max_abs_dv": verr, "tolerance": tol, "ok": ok, "slabs": [list(s) for s in S.slabs],
                "rank0_owned": info["owned"], "rank0_ghosts": info["ghosts"], "replans": S.replans, "owned_per_rank": final_owned,
                "worst_imbalance_seen": worst_imbalance}
        print(json.dumps(line), flush=True)
        G.close()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.broadcast(flag, 0)
    S.close()
    dist.barrier()
    dist.destroy_process_group()
    return int(flag.item())


if __name__ == "__main__":
    sys.exit(main())
