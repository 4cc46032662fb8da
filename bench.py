#!/usr/bin/env python
"""bench.py — particle-substeps/s of the Lustrine particle step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one substep (predict + grid build + `iterations` solver iterations + commit) of the
synthetic scene named in config.workload.  Prints ONE JSON line (rank 0).

  value     whole-job particle-substeps/s with the state resident in HBM, timed on the device with
            CUDA events around every substep (max over ranks), L2 flushed between substeps.
  e2e       the same metric through the drop-in boundary with HOST buffers: per step the
            positions/velocities/flags are copied host->device from pinned memory, the substep runs,
            and the new positions/velocities/flags are copied device->host (what
            Lustrine::Simulation::simulate_fun does when the caller owns the arrays).
  roofline  achieved algorithmic HBM bytes/s of the dominant kernel vs MEASURED_PEAKS.json.
  cpu_baseline  the reference's own CPU step (oracle/_ref, unmodified sources, Release flags) timed on
            this box's host cores on a bounded sample (rank 0, N=1 only).
"""
import argparse
import contextlib
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

# algorithmic bytes per particle-substep (SURVEY §8d / DESIGN.md §Roofline)
def fluid_bytes_per_particle(K, cells_per_particle):
    return 116 + 44 * K + (12 + 8 * K) * cells_per_particle

def sand_bytes_per_particle(K, cells_per_particle):
    return 48 + 60 + 36 * K + 24 + (12 + 8 * K) * cells_per_particle

# per-kernel algorithmic bytes per particle
KERNEL_BYTES = {"fluid_lambda": 16.0, "fluid_deltap": 28.0, "sand_iteration": 36.0}

WORKLOADS = {
    # name: (kind, cube side, iterations, dt)
    "dam_break_1m": ("fluid", 100, 4, 0.01),    # BASELINE.json configs[1]
    "dam_break_64k": ("fluid", 40, 4, 0.01),
    "dam_break_1m_k1": ("fluid", 100, 1, 0.01),  # one solver iteration (the reference's own setting)
    "dam_break_2m": ("fluid", 126, 4, 0.01),    # weak scaling: N x 1M particles on N GPUs
    "dam_break_4m": ("fluid", 159, 4, 0.01),
    "dam_break_8m": ("fluid", 200, 4, 0.01),
    "dam_break_16m": ("fluid", 252, 4, 0.01),   # BASELINE.json configs[3]
    "sand_pile_4m": ("sand", 160, 4, 0.016),    # BASELINE.json configs[2]
    "sand_pile_262k": ("sand", 64, 4, 0.016),
}


def load_traffic(workload, kernel):
    """DRAM bytes per launch of `kernel` on `workload` from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return float(t[workload][kernel]), t.get("source")
    except Exception:
        return None, None


def load_pipes(kernel):
    """Utilisation of the fp32 (FMA) pipe, the shared-memory data pipe and the issue slots of `kernel` from the same committed ncu
    capture (SURVEY §8d asks for the fp32 pipe beside the HBM figure), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)["pipes"][kernel]
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_scene(kind, side):
    import scenes
    if kind == "fluid":
        domain, sand = scenes.dam_break(side)
        return domain, sand, None
    # sand pile: side^3 block dropped on a solid voxel floor with obstacle boxes
    D = (3 * side, side + side // 2, 3 * side)
    sand = scenes.lattice(side, side, side, origin=(side + 0.5, 8.5, side + 0.5), jitter=0.05, seed=777)
    floor = scenes.lattice(3 * side, 2, 3 * side, origin=(0.5, 0.5, 0.5), jitter=0.0)
    boxes = [scenes.lattice(side // 4, 4, side // 4, origin=(side + 0.5 + k * side // 2, 2.5, side + 0.5 + k * side // 3), jitter=0.0)
             for k in range(2)]
    return D, sand, np.concatenate([floor] + boxes).astype(np.float32)


def step_kwargs(kind, K, dt, exact, literal=0):
    kw = dict(dt=dt, iterations=K, exact_math=int(exact))
    if kind == "fluid":
        # lambdas[neighbour] by default; --literal 1 = the reference's lambdas[loop counter], which is
        # only defined on a single GPU and collapses the column within a few hundred substeps
        kw["literal_lambda_index"] = int(literal)
    return kw


# ------------------------------------------------------------------------------------------------
def workload_config(workload, kind, n, n_solid, domain, grid_cells, K, dt, literal, reset_every):
    """The keys that NAME the workload: identical in the B200 arm and in the reference arm."""
    return {"workload": workload, "kind": kind, "particles": int(n), "solid_particles": int(n_solid), "domain": [int(x) for x in domain],
            "grid_cells": int(grid_cells), "solver_iterations": int(K), "dt": dt,
            "literal_lambda_index": int(literal) if kind == "fluid" else None, "scene_reset_every": int(reset_every)}


def grid_cells_of(domain, radius=0.5, scale=3.1):
    """(int)(D / h) + 1 per axis, h = 3.1 r in fp32 (reference src/Lustrine.cpp:253-266)."""
    h = np.float32(scale) * np.float32(radius)
    g = [int(np.float32(d) / h) + 1 for d in domain]
    return g[0] * g[1] * g[2]


def run_cpu_reference(kind, side, K, dt, steps, warmup, budget_s, literal=0, stock=False):
    """Times the reference's own CPU step.  Uses the compiled unmodified reference (oracle/_ref/libref_time.so, the
    reference's Release flags) when present, else the plain-C port.  Single-threaded: the reference has no threading.
    The scene is the workload's own (exactly `side`^3 particles) whenever the whole run fits the budget; otherwise
    a smaller cube of the same scene family.  stock=True: the unmodified simulate_fluid (one in-place Gauss-Seidel
    iteration, lambdas[loop counter]) instead of the K-iteration Jacobi loop of oracle/ref_harness.cpp."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    # K=4 fluid: ~3.8e5 particle-substeps/s on one core of the GPU boxes; sand: ~5e5 (BENCH_r01.json)
    est = (3.5e5 if K > 1 else 4.5e5) if kind == "fluid" else 4.5e5
    full_n = side ** 3
    sample_side = side
    # (the reference sizes every array by the domain volume, SURVEY F17: scenes above ~110^3 need tens of GB of host memory)
    if full_n * (steps + warmup) / est > budget_s or side > 110:
        n_target = est * budget_s / max(steps + warmup, 1)
        sample_side = int(max(8, min(side, 100, int(n_target ** (1.0 / 3.0)))))
    domain, sand, solids = make_scene(kind, sample_side)
    if kind == "sand" and solids is not None and len(solids) > 60000:
        # one Bullet box per solid voxel is created by the reference's init: keep the floor modest
        solids = solids[:60000]
    n = len(sand)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)  # the reference prints banners on stdout
    try:
        if O.have_ref("time"):
            kind_name = "reference"
            R = O.RefSim(*domain, n_sand=n, n_solid=0 if solids is None else len(solids), which="time")
            R.set_sand(sand)
            if solids is not None:
                R.set_solid(solids)
            if kind == "fluid":
                if stock:
                    R.set_fun(R.FLUID)
                else:
                    R.set_fun(R.FLUID_JACOBI, K, bool(literal))
            else:
                R.set_fun(R.SAND)
            R.step(dt, warmup)
            seconds = R.step(dt, steps)
            R.close()
        else:
            kind_name = "port"
            P = O.PortSim(*domain, capacity=n, n_solid=0 if solids is None else len(solids))
            P.set_sand(sand)
            if solids is not None:
                P.set_solid(solids)
            def one():
                if kind == "fluid":
                    P.L.lo_step_fluid(P.p, dt, 1 if stock else K, 0 if stock else 1, 1 if stock else int(literal))
                else:
                    P.L.lo_step_sand(P.p, dt, K, 0)
            for _ in range(warmup):
                one()
            t0 = time.perf_counter()
            for _ in range(steps):
                one()
            seconds = time.perf_counter() - t0
            P.close()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
    value = n * steps / seconds
    how = ("unmodified reference sources, -O3 -mavx2 -ffast-math -mfma -march=skylake" if kind_name == "reference" else "plain-C port of the reference, -O2")
    if kind == "fluid":
        how += ("; stock simulate_fluid (src/Simulate.cpp:27-115): ONE in-place iteration, lambdas[loop counter]" if stock else
                "; K-iteration loop = oracle/ref_harness.cpp simulate_fluid_jacobi over the reference's own neighbour search and kernels, lambdas[%s]"
                % ("loop counter" if literal else "neighbour"))
    sample = ("%s scene, %d^3 = %d particles%s, %d substeps (%d warm-up), %d solver iterations, dt=%g; %s"
              % (kind, sample_side, n, "" if sample_side == side else " (the workload has %d^3: bounded sample)" % side, steps, warmup,
                 1 if stock else K, dt, how))
    return {"value": value, "unit": "particle-substeps/s", "cores": 1, "kind": kind_name, "sample": sample,
            "seconds": seconds, "host_cores_available": os.cpu_count(), "exact_workload_size": sample_side == side}, n, seconds, (domain, solids)


def run_slabs(args, workload, kind, side, K, dt, steps, warmup, hbm_gbs, peak_src, domain, sand, solids, rank, world, local_rank):
    """N > 1: the scene partitioned into spatial slabs, one per GPU (lustrine_b200/slabs.py)."""
    import torch
    import torch.distributed as dist
    from lustrine_b200 import lgpu, slabs
    n_total = len(sand)
    # Weak-scaling reference measured in this same run: the scene with 1/N of the particles (cube of side / N^(1/3)) on ONE
    # GPU — every rank runs it on its own device at the same time, the slowest rank counts.
    single = None
    if kind == "fluid" and not args.no_extras:
        s1 = int(round(side / world ** (1.0 / 3.0)))
        name1 = "dam_break_1gpu_%d" % s1
        WORKLOADS[name1] = (kind, s1, K, dt)
        fl = Flusher()
        r1, G1, _ = single_gpu_workload(name1, args, local_rank, fl, hbm_gbs, peak_src, min(steps, 20), 3)
        G1.close()
        del fl
        torch.cuda.empty_cache()
        t1 = torch.tensor([r1["ms_per_step"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t1, op=dist.ReduceOp.MAX)
        single = {"workload": name1, "particles": s1 ** 3, "ms_per_step": float(t1.item()), "value": s1 ** 3 / (float(t1.item()) * 1e-3)}
    # one-column ghost layer.  (LGPU_GHOST_COLUMNS=2, fluid: the inner ghosts' lambdas are computed locally, K - 1 ghost
    # refreshes per substep instead of 2K - 1 — measured 1.3 % SLOWER on 2 and 8 B200s: the second ghost column's work
    # costs what the four saved handshakes gain.)
    gw = int(os.environ.get("LGPU_GHOST_COLUMNS", "1"))
    # the plan is cropped to the occupied cell columns + a margin (LGPU_SLAB_MARGIN=none: the whole grid, whose empty two
    # thirds then land on the last rank)
    m_env = os.environ.get("LGPU_SLAB_MARGIN", str(slabs.DEFAULT_MARGIN))
    margin = None if m_env.lower() == "none" else int(m_env)
    S = slabs.DistributedSlab(domain, sand, solids=solids, device=local_rank, ghost_columns=gw, margin=margin)
    G = S.G
    mode = 1 if kind == "fluid" else 2
    params = lgpu.default_step_params(**step_kwargs(kind, K, dt, args.exact, 0))
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush_buf.zero_()
        torch.cuda.synchronize()

    def maybe_reset(k):
        if args.reset_every and k % args.reset_every == 0:
            S.reset()

    for k in range(warmup):
        maybe_reset(k)
        S.step(mode, params)
    G.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = G.launch_count()
    dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    owned_sum = 0
    for k in range(steps):
        maybe_reset(k)
        flush_l2()
        dist.barrier()          # all ranks start the substep together (outside the event bracket)
        torch.cuda.synchronize()
        S.step(mode, params)
        G.sync()
        dev_ms += G.last_step_ms(0)
        owned_sum += G.n
    dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    launches = G.launch_count() - launches0
    sampler.stop_flag = True
    sampler.join()
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps                      # max over ranks of the summed device time
    cnt = torch.tensor([float(launches), float(owned_sum) / steps], dtype=torch.float64, device="cuda")
    gathered = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(gathered, cnt)
    value = n_total / (ms_per_step * 1e-3)

    # per-kernel timing pass
    G.set_phase_timing(True)
    phase = np.zeros(9)
    PH = 10
    S.reset()
    for _ in range(PH):
        flush_l2()
        dist.barrier()
        S.step(mode, params)
        G.sync()
        for ph in range(9):
            phase[ph] += G.last_step_ms(ph)
    phase /= PH
    G.set_phase_timing(False)
    info = G.slab_info()
    # every rank's phases and grid size (rank 0 is an edge slab: one neighbour, and without the crop not the critical path)
    pr = torch.tensor(list(phase) + [float(info["local_cells"]), float(info["owned"]), float(info["ghosts"])], dtype=torch.float64, device="cuda")
    pr_all = [torch.zeros_like(pr) for _ in range(world)]
    dist.all_gather(pr_all, pr)
    pr_all = np.array([x.cpu().numpy() for x in pr_all])
    n_local = info["owned"]
    cells_per_particle = info["local_cells"] / max(n_local, 1)
    step_bytes = fluid_bytes_per_particle(K, cells_per_particle) if kind == "fluid" else sand_bytes_per_particle(K, cells_per_particle)
    if kind == "fluid":
        cand = {"fluid_deltap": phase[7] / K}
        if K > 1:
            cand["fluid_lambda"] = phase[6] / (K - 1)   # (the first lambda pass runs inside the table build)
    else:
        cand = {"sand_iteration": phase[7] / K}
    dom = max(cand, key=lambda k: cand[k])
    dom_bytes = KERNEL_BYTES[dom] * (info["owned"] + info["ghosts"])
    achieved = dom_bytes / (cand[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                "traffic": None, "peak_source": peak_src, "kernel_ms": cand[dom], "algorithmic_bytes_per_launch": dom_bytes,
                "note": "rank 0's slab; per-GPU figures"}
    step_achieved = step_bytes * n_total / world / (ms_per_step * 1e-3) / 1e9
    roofline_step = {"bound": "hbm", "achieved": step_achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": step_achieved / hbm_gbs,
                     "algorithmic_bytes_per_particle_substep": step_bytes, "per": "GPU",
                     "phases_ms_rank0": {"predict_key_hist_halo": phase[1], "scan": phase[2], "scatter_reorder": phase[3], "neighbour_table": phase[4],
                                         "lambda_total": phase[6], "deltap_or_contact_total": phase[7], "halo_refresh_total": phase[8]},
                     "phases_ms_per_rank": {"predict_key_hist_halo": pr_all[:, 1].tolist(), "scan": pr_all[:, 2].tolist(), "scatter_reorder": pr_all[:, 3].tolist(),
                                            "neighbour_table": pr_all[:, 4].tolist(), "lambda_total": pr_all[:, 6].tolist(),
                                            "deltap_or_contact_total": pr_all[:, 7].tolist(), "halo_refresh_total": pr_all[:, 8].tolist()},
                     "local_cells_per_rank": [int(x) for x in pr_all[:, 9]], "owned_per_rank": [int(x) for x in pr_all[:, 10]],
                     "ghosts_per_rank": [int(x) for x in pr_all[:, 11]],
                     # what a rank receives per substep over NVLink (its neighbours send the same amounts): 64-byte halo records
                     # for its ghosts after predict, then 4 bytes (lambda) / 16 bytes (x*) per ghost after every pass but the last
                     "nvlink_bytes_in_per_substep_per_rank": [int(g) * (64 + (4 * K + 16 * (K - 1) if kind == "fluid" and gw < 2 else 16 * (K - 1)))
                                                              for g in pr_all[:, 11]]}

    # e2e: host buffers in, host buffers out, every substep, on every rank
    e2e = None
    if not args.no_e2e:
        S.reset()
        # pinned host buffers (two sets: the step's input and its output), sized for the slab's capacity
        cap = int(G.n * 1.5) + 4096
        def pinned_set():
            return (torch.empty((cap, 3), dtype=torch.float32, pin_memory=True).numpy(), torch.empty((cap, 3), dtype=torch.float32, pin_memory=True).numpy(),
                    torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy(), torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy())
        bufs = [pinned_set(), pinned_set()]
        pos, vel, flags, ids = G.slab_download(out=bufs[0])
        e_steps = min(steps, args.reset_every or 20)
        turn = [1]
        def one(pos, vel, flags, ids):
            G.slab_upload(pos, ids, vel, flags)
            S.step(mode, params)
            out = G.slab_download(out=bufs[turn[0]])
            turn[0] ^= 1
            return out
        for _ in range(2):
            pos, vel, flags, ids = one(pos, vel, flags, ids)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        moved = 0
        for _ in range(e_steps):
            moved += len(ids) * 32 * 2
            pos, vel, flags, ids = one(pos, vel, flags, ids)
        torch.cuda.synchronize()
        e_t = torch.tensor([time.perf_counter() - t0, float(moved)], dtype=torch.float64, device="cuda")
        e_max = e_t.clone()
        dist.all_reduce(e_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(e_t, op=dist.ReduceOp.SUM)
        e_ms = float(e_max[0].item()) * 1e3 / e_steps
        e2e = {"value": n_total / (e_ms * 1e-3), "unit": "particle-substeps/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(float(e_t[1].item()) / e_steps / 2), "d2h_bytes_per_step": int(float(e_t[1].item()) / e_steps / 2), "steps": e_steps,
               "path": "lgpu_slab_upload + lgpu_step + lgpu_slab_download with pinned host buffers on every rank"}
    plan = S.slabs
    line = {"metric": "particle-substeps/s", "value": value, "unit": "particle-substeps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(workload, kind, n_total, 0 if solids is None else len(solids), domain, grid_cells_of(domain), K, dt, 0, args.reset_every),
            "notes": {"scaling_note": "every line is throughput (particle-substeps/s), so value(N) / (N * value(1)) is a weak-scaling efficiency; the 1-GPU "
                                      "line is the 1M scene of BASELINE configs[1], the N-GPU lines (N > 1) the 16M scene of configs[3] (%d particles per GPU "
                                      "here); weak_efficiency below compares like with like" % (n_total // world),
                      "particles_per_gpu": [int(g[1].item()) for g in gathered],
                      "partition": "x-slabs of whole cell columns, %d-column ghost layer, migration + ghost refresh written peer-to-peer over NVLink" % gw,
                      "slabs": [list(x) for x in plan], "collective": "none on the data path",
                      "arithmetic": "exact (reference op order)" if args.exact else "fast (FMA + approx rsqrt, parity-tested to 1e-5)",
                      "l2": "flushed between substeps (256 MiB write, outside the event bracket)",
                      "timing": "CUDA events on the launching stream around every substep (barrier before each), summed, max over ranks",
                      "wall_ms_per_step_incl_flush_and_barriers": t_wall * 1e3 / steps},
            "roofline": roofline, "roofline_step": roofline_step, "cpu_baseline": None, "e2e": e2e,
            "gpu_launches": int(sum(g[0].item() for g in gathered)), "clocks": sampler.result()}
    if single is not None:
        line["single_gpu_same_load"] = single
        line["weak_efficiency"] = single["ms_per_step"] / ms_per_step   # T(1 GPU, particles / N) / T(N GPUs, particles)
    S.close()
    return line


class Flusher:
    """Writes a buffer larger than the 126 MB L2 between timed substeps (outside the event bracket)."""

    def __init__(self):
        import torch
        self.torch = torch
        self.buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def __call__(self):
        self.buf.zero_()
        self.torch.cuda.synchronize()


def timed_substeps(G, step, params, steps, warmup, flush, reset=None, reset_every=0):
    """`warmup` untimed substeps, then `steps` substeps each bracketed by CUDA events on the launching stream (summed),
    L2 flushed between them.  reset(): puts the scene back (outside the bracket) every `reset_every` substeps."""
    for k in range(warmup):
        if reset and reset_every and k % reset_every == 0:
            reset()
        step(params)
    G.sync()
    dev_ms = 0.0
    timed_substeps.launches = 0
    for k in range(steps):
        if reset and reset_every and k % reset_every == 0:
            reset()
        flush()
        l0 = G.launch_count()
        step(params)
        G.sync()
        timed_substeps.launches += G.launch_count() - l0   # kernels of the substeps themselves (not the scene re-uploads)
        dev_ms += G.last_step_ms(0)
    return dev_ms / steps


def phase_times(G, step, params, flush, reset, n_phases=9, reps=20):
    """Per-kernel-kind device times (ms per substep): event marks around every launch (a separate short pass)."""
    G.set_phase_timing(True)
    phase = np.zeros(n_phases)
    reset()
    for _ in range(reps):
        flush()
        step(params)
        G.sync()
        for ph in range(n_phases):
            phase[ph] += G.last_step_ms(ph)
    G.set_phase_timing(False)
    return phase / reps


def rooflines(kind, K, n, num_cells, phase, ms_per_step, hbm_gbs, peak_src, workload):
    """roofline of the dominant solver kernel and of the whole substep (algorithmic bytes: SURVEY §8d)."""
    cells_per_particle = num_cells / max(n, 1)
    if kind == "fluid":
        step_bytes = fluid_bytes_per_particle(K, cells_per_particle)
        # the first density + lambda pass runs inside the table build (phase 4): K - 1 separate lambda launches
        cand = {"fluid_deltap": phase[7] / K}
        if K > 1:
            cand["fluid_lambda"] = phase[6] / (K - 1)
    else:
        step_bytes = sand_bytes_per_particle(K, cells_per_particle)
        cand = {"sand_iteration": phase[7] / K}
    dom = max(cand, key=lambda k: cand[k])
    dom_ms = cand[dom]
    dom_bytes = KERNEL_BYTES[dom] * n
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = load_traffic(workload, "k_" + dom)
    roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s",
                "frac": achieved / hbm_gbs, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel_ms": dom_ms, "algorithmic_bytes_per_launch": dom_bytes,
                # % of peak of the fp32 (FMA) pipe, the shared-memory data pipe and the issue slots, from the same ncu capture
                "pipes_ncu": load_pipes("k_" + dom),
                "note": "sparse stencil: ~19 neighbour interactions per particle per launch are gathered from a shared-memory stage of the "
                        "particle's brick; the kernel is bound by that gather (LDS wavefronts + issue slots), not by HBM (DESIGN.md §5)"}
    step_achieved = step_bytes * n / (ms_per_step * 1e-3) / 1e9
    roofline_step = {"bound": "hbm", "achieved": step_achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": step_achieved / hbm_gbs,
                     "algorithmic_bytes_per_particle_substep": step_bytes,
                     "phases_ms": {"predict_key_hist": phase[1], "scan": phase[2], "scatter_reorder": phase[3],
                                   "neighbour_table" + ("_and_first_lambda" if kind == "fluid" else ""): phase[4],
                                   "lambda_total": phase[6], "deltap_or_contact_total": phase[7],
                                   "sum": float(phase[1:5].sum() + phase[6] + phase[7])}}
    return roofline, roofline_step


def single_gpu_workload(workload, args, device, flush, hbm_gbs, peak_src, steps, warmup, free_run_steps=0):
    """One workload on one GPU through the C ABI with the state resident in HBM.  Returns (dict of results, context, scene)."""
    from lustrine_b200 import lgpu
    kind, side, K, dt = WORKLOADS[workload]
    domain, sand, solids = make_scene(kind, side)
    n = len(sand)
    G = lgpu.Context(domain, capacity_sand=n, capacity_solid=0 if solids is None else len(solids), device=device)
    G.upload_sand(sand)
    if solids is not None:
        G.upload_solids(solids)
    params = lgpu.default_step_params(**step_kwargs(kind, K, dt, args.exact, args.literal))
    step = G.step_fluid if kind == "fluid" else G.step_sand
    reset = lambda: G.upload_sand(sand)
    # the substep's launches are captured once and replayed as ONE CUDA graph per substep (lgpu_set_use_graph)
    G.set_use_graph(0 if args.no_graph else 1)
    ms = timed_substeps(G, step, params, steps, warmup, flush, reset, args.reset_every)
    launches = timed_substeps.launches
    G.set_use_graph(0)
    phase = phase_times(G, step, params, flush, reset)
    counters = G.dump(lgpu.DUMP_COUNTERS)
    roofline, roofline_step = rooflines(kind, K, n, G.num_cells, phase, ms, hbm_gbs, peak_src, workload)
    out = {"ms_per_step": ms, "value": n / (ms * 1e-3), "roofline": roofline, "roofline_step": roofline_step,
           "table_overflows": int(counters[1]), "key_violations": int(counters[0]), "launches": int(launches),
           "config": workload_config(workload, kind, n, 0 if solids is None else len(solids), domain, G.num_cells, K, dt,
                                     args.literal if kind == "fluid" else 0, args.reset_every)}
    if free_run_steps:
        # the scene left to itself (no re-upload): the column collapses, lists lengthen, warps diverge
        reset()
        G.sync()
        fr_ms = timed_substeps(G, step, params, free_run_steps, 0, flush)
        c2 = G.dump(lgpu.DUMP_COUNTERS)
        # the second half alone: the collapsed regime
        tail_ms = timed_substeps(G, step, params, max(free_run_steps // 4, 1), 0, flush)
        out["free_run"] = {"substeps": free_run_steps, "ms_per_step": fr_ms, "value": n / (fr_ms * 1e-3),
                           "ms_per_step_after": tail_ms, "table_overflows": int(c2[1]) - int(counters[1]),
                           "key_violations": int(c2[0]) - int(counters[0]),
                           "note": "reset-every 0: %d free-running substeps from the initial scene (SURVEY §8d config 2), then %d more "
                                   "(ms_per_step_after); table_overflows = particle-substeps that re-walked the stencil (list longer than 64 entries, or a "
                                   "brick part that does not fit a block's shared memory even when cut down to one cell layer)"
                                   % (free_run_steps, max(free_run_steps // 4, 1))}
    return out, G, (domain, sand, solids, params, step, kind, K, dt)


def wrapper_e2e(workload, steps, sync_mode):
    """The same metric through the reference-facing plugin boundary with HOST buffers: LustrineWrapper's C entry points
    on liblustrine_b200.so — init_grid_box + init_simulation, then per step Wrapper::simulate(dt, attract, blow)
    followed by simulation_bind_positions_copy into a pinned caller buffer (reference src/LustrineWrapper.cpp:343-379).
    sync_mode 0 = SYNC_FULL: the host arrays of Lustrine::Simulation stay authoritative, every step uploads
    positions / velocities / attracted and downloads them again; 2 = SYNC_LAZY: the state stays on the device and only
    the positions come back (what the game does)."""
    import torch
    from wrapper_driver import Wrapper
    kind, side, K, dt = WORKLOADS[workload]
    lib = os.path.join(ROOT, "lustrine_b200", "lib", "liblustrine_b200.so")
    if not os.path.exists(lib):
        raise RuntimeError("liblustrine_b200.so is missing: run __graft_entry__.build()")
    os.environ.setdefault("LUSTRINE_B200_QUIET", "1")
    W = Wrapper(lib)
    domain = (3 * side, 2 * side, 2 * side) if kind == "fluid" else (3 * side, side + side // 2, 3 * side)
    sand_pos = (1.0, 1.0, 1.0) if kind == "fluid" else (float(side), 8.0, float(side))
    solids = [] if kind == "fluid" else [((3 * side, 2, 3 * side), (0.0, 0.0, 0.0), 2)]
    data = W.init(domain, 0.5, sand=[((side, side, side), sand_pos)], solids=solids, subdivision=1)
    n = data.num_sand_particles
    if kind == "fluid":
        W.L.b200_set_simulate_function(2)       # Simulation::simulate_fun = simulate_fluid (set through the public field)
        W.L.b200_set_solver_options(K, 0, 0)    # K iterations, lambdas[neighbour], throughput arithmetic
    else:
        W.L.b200_set_solver_options(1, 0, 0)
    W.L.b200_set_host_sync(sync_mode)
    out = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    for _ in range(3):
        W.L.simulate(dt, False, False)
        W.L.simulation_bind_positions_copy(out.data_ptr())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        W.L.simulate(dt, False, False)
        W.L.simulation_bind_positions_copy(out.data_ptr())
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    checksum = float(out.double().sum())
    W.L.cleanup_simulation()
    # per particle, SYNC_FULL: up positions 12 + velocities 12 (+ attracted 4 for a sand step: simulate_fluid never touches
    # it); down the same plus positions_star 12 (the reference leaves it equal to positions) plus the 12 of
    # simulation_bind_positions_copy — all by the copy engine, the Simulation's host arrays being page-locked
    fl = 0 if kind == "fluid" else 4
    h2d = n * (24 + fl) if sync_mode == 0 else 0
    d2h = n * (24 + fl + 12 + 12) if sync_mode == 0 else n * 12
    return {"value": n / (ms * 1e-3), "unit": "particle-substeps/s", "ms_per_step": ms, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "steps": steps, "particles": int(n), "position_checksum": checksum,
            "path": "LustrineWrapper C API on liblustrine_b200.so: simulate(dt, attract, blow) + simulation_bind_positions_copy(pinned buffer), "
                    + ("SYNC_FULL (host arrays authoritative: positions, velocities (and attracted around a sand step) uploaded and downloaded "
                       "every step, positions_star refreshed; the Simulation's host arrays are page-locked by the drop-in)" if sync_mode == 0
                       else "SYNC_LAZY (state resident on the device, positions copied out every step)")}


@contextlib.contextmanager
def quiet_stdout():
    """The reference (and the kept API, like it) prints banners on stdout; main() has already sent fd 1 to stderr, this
    drops them altogether for the chatty set-up calls."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)


def fluid_demo_config1(calls, with_cpu):
    """BASELINE configs[0]: the reference's own fluid experiment (experiments/fluid/fluid.cpp:46-118: 40x40x80 domain,
    level1_physical.vox solids, 27 sand particles, four sources, one sink), headless, `calls` frames of
    Lustrine::simulate with simulate_fun = simulate_fluid, dt = 0.01.  The caller is tests/cpp/fluid_demo.cpp, ONE
    C++ source compiled against the drop-in and against the reference; the scene holds a few thousand particles, so
    the GPU side is bound by launch latency and the host<->device round trip of every frame (SYNC_FULL), not by bandwidth."""
    from fluid_demo_driver import DEMO_B200, DEMO_REF, Demo
    if not os.path.exists(DEMO_B200):
        raise RuntimeError("tests/cpp/_build/libfluid_demo_b200.so is missing: run __graft_entry__.build()")
    out = {"config": "experiments/fluid scene, %d calls of Lustrine::simulate, simulate_fluid, dt=0.01" % calls}
    with quiet_stdout():
        d = Demo(DEMO_B200, 2)
        d.run(20)
        counts, _, seconds = d.run(calls)
        d.close()
    out["b200"] = {"calls_per_s": calls / seconds, "value": float(counts.sum()) / seconds, "unit": "particle-substeps/s",
                   "live_particles_last_frame": int(counts[-1]), "seconds": seconds}
    if with_cpu and os.path.exists(DEMO_REF):
        with quiet_stdout():
            r = Demo(DEMO_REF, 2)
            r.run(20)
            rcounts, _, rseconds = r.run(calls)
            r.close()
        out["cpu_reference"] = {"calls_per_s": calls / rseconds, "value": float(rcounts.sum()) / rseconds, "unit": "particle-substeps/s",
                                "live_particles_last_frame": int(rcounts[-1]), "seconds": rseconds, "cores": 1, "kind": "reference",
                                "sample": "the same %d frames on the unmodified reference (-O2 -ffp-contract=off build, fluid loop double-buffered: SURVEY F5)" % calls}
    return out


def mixed_scene_config5(steps):
    """BASELINE configs[4]: mixed scene through the LustrineWrapper C API (reference experiments/bullet/bullet.cpp:71-121
    scaled up): 126^3 = 2 000 376 sand particles on voxel solids, a capsule player + boxes on the HOST rigid-body side,
    init_simulation_extra_parameters(subdivision 1, kernel_radius_scale 3.1, no credits), simulate(dt = 0.016, attract,
    blow) with attraction on steps 100..199 and blowing on 300..319, particle bounding boxes around the player enabled,
    simulation_bind_positions_copy every step (timed separately).  The rigid bodies run on the host as in the
    reference; in this repository they are integrated by host/HostBodies.cpp, a stand-in for Bullet."""
    import torch
    from wrapper_driver import Vec3, Wrapper
    lib = os.path.join(ROOT, "lustrine_b200", "lib", "liblustrine_b200.so")
    side = 126
    domain = (3 * side, 2 * side, 3 * side)
    os.environ.setdefault("LUSTRINE_B200_QUIET", "1")
    os.environ["LUSTRINE_B200_MAX_SAND"] = str(side ** 3 + 65536)
    W = Wrapper(lib)
    L = W.L
    L.add_capsule.argtypes = [Vec3, C.c_float, C.c_float]
    L.add_box.argtypes = [Vec3, C.c_bool, Vec3]
    L.set_player_box_scale.argtypes = [Vec3]
    solids = [((3 * side, 2, 3 * side), (0.0, 0.0, 0.0), 2), ((40, 8, 40), (float(side) + 20.0, 2.0, float(side) + 20.0), 2),
              ((30, 12, 30), (float(side) + 70.0, 2.0, float(side) + 60.0), 2)]
    with quiet_stdout():
        data = W.init(domain, 0.5, sand=[((side, side, side), (float(side), 3.0, float(side)))], solids=solids, subdivision=1, extra=(3.1, 0))
        n = data.num_sand_particles
        top = 3.0 + side
        player = L.add_capsule(Vec3(1.5 * side, top + 6.0, 1.5 * side), 2.0, 3.0)
        L.add_box(Vec3(1.5 * side + 1.0, top + 3.0, 1.5 * side + 1.0), True, Vec3(1.0, 1.0, 1.0))
        L.add_box(Vec3(1.5 * side - 1.0, top + 3.0, 1.5 * side - 1.0), True, Vec3(1.0, 1.0, 1.0))
        L.add_box(Vec3(1.5 * side - 5.0, top + 2.0, 1.5 * side - 5.0), True, Vec3(3.0, 1.0, 1.0))
        L.add_box(Vec3(1.5 * side, -0.5, 1.5 * side), False, Vec3(1.5 * side, 1.0, 1.5 * side))
        L.set_player_id(player)
        L.set_player_box_scale(Vec3(2.0, 2.0, 2.0))
        L.enable_particles_bounding_boxes()
        L.set_attract_blow_parameters(12.0, 9.0, 1000.0, 500.0)
        L.b200_set_solver_options(1, 0, 0)
        L.b200_set_host_sync(2)  # the state stays on the device; positions are copied out on request (what the game does)
    out = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    t_sim = t_copy = 0.0
    for s in range(steps):
        attract, blow = 100 <= s < 200, 300 <= s < 320
        t0 = time.perf_counter()
        L.simulate(0.016, attract, blow)
        t1 = time.perf_counter()
        L.simulation_bind_positions_copy(out.data_ptr())
        t2 = time.perf_counter()
        if s >= 5:
            t_sim += t1 - t0; t_copy += t2 - t1
    timed = steps - 5
    checksum = float(out.double().sum())
    with quiet_stdout():
        L.cleanup_simulation()
    return {"value": n * timed / t_sim, "unit": "particle-substeps/s", "particles": int(n), "solid_particles": int(data.num_solid_particles),
            "steps": steps, "ms_per_step_simulate": 1e3 * t_sim / timed, "ms_per_step_positions_copy": 1e3 * t_copy / timed,
            "attract_steps": [100, 200], "blow_steps": [300, 320], "position_checksum": checksum,
            "d2h_bytes_per_step": int(n) * 12, "h2d_bytes_per_step": 0,
            "path": "LustrineWrapper C API on liblustrine_b200.so (init_grid_box, init_simulation_extra_parameters, add_capsule, add_box, set_player_id, "
                    "enable_particles_bounding_boxes, simulate, simulation_bind_positions_copy), SYNC_LAZY",
            "note": "wall clock of the C calls on the host (the step is synchronous at return); the rigid bodies are stepped on the host by "
                    "host/HostBodies.cpp, a stand-in for Bullet; no CPU reference at this size (the reference needs ~35 s per 2M-particle step, see secondary.sand_pile_4m.cpu_baseline)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--exact", type=int, default=0, help="1 = parity arithmetic (every fp32 op separately rounded)")
    ap.add_argument("--literal", type=int, default=0, help="fluid: 1 = lambdas[loop counter] as in the reference (single GPU only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches (with programmatic dependent launch) instead of one CUDA graph per substep")
    ap.add_argument("--no-extras", action="store_true", help="skip free_run / stock_pair / secondary workloads (profiling runs)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--free-run", type=int, default=200, help="free-running substeps of the free_run leg")
    ap.add_argument("--reset-every", type=int, default=20,
                    help="re-upload the initial scene every R substeps (outside the timed bracket); 0 = free run")
    args = ap.parse_args()

    # stdout carries exactly ONE line (the JSON): anything a library prints there (NCCL's version banner,
    # torchrun notices, the wrapper's init banner) is sent to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global emit
    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = args.steps, max(args.warmup, 3)

    # BASELINE.json: 1 GPU = the 1M dam break (configs[1]); 2 / 4 / 8 GPUs = the 16M dam break partitioned
    # into 2 / 4 / 8 slabs (configs[3])
    workload = args.workload or ("dam_break_1m" if args.gpus <= 1 else "dam_break_16m")
    kind, side, K, dt = WORKLOADS[workload]
    metric = "particle-substeps/s"

    # ---------------- reference arm: the reference's own CPU step ----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, n, seconds, (domain, solids) = run_cpu_reference(kind, side, K, dt, steps, warmup, budget_s=420.0, literal=args.literal)
        full_domain = (3 * side, 2 * side, 2 * side) if kind == "fluid" else (3 * side, side + side // 2, 3 * side)
        n_solid_full = 0 if kind == "fluid" else 3 * side * 2 * 3 * side + 2 * (side // 4) * 4 * (side // 4)   # make_scene: floor + two boxes
        line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": "particle-substeps/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * seconds / steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(workload, kind, side ** 3, n_solid_full, full_domain,
                                          grid_cells_of(full_domain), K, dt, args.literal, args.reset_every),
                "notes": {"sample": cb["sample"], "exact_workload_size": cb["exact_workload_size"],
                          "timing": "steady_clock around the reference's simulate_fun on one host core (the reference has no threading)"},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "particle-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    # ---------------- B200 arm ----------------
    import torch
    import torch.distributed as dist
    from lustrine_b200 import lgpu
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the particle step has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    hbm_gbs, peak_src = load_peaks()
    if world > 1:
        domain, sand, solids = make_scene(kind, side)
        result = run_slabs(args, workload, kind, side, K, dt, steps, warmup, hbm_gbs, peak_src, domain, sand, solids, rank, world, local_rank)
        if rank == 0:
            emit(result)
        dist.barrier()
        dist.destroy_process_group()
        return 0

    flush = Flusher()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wall0 = time.perf_counter()
    main_res, G, scene = single_gpu_workload(workload, args, local_rank, flush, hbm_gbs, peak_src, steps, warmup,
                                             free_run_steps=0 if args.no_extras else args.free_run)
    sampler.stop_flag = True
    sampler.join()
    domain, sand, solids, params, step, kind, K, dt = scene
    n = len(sand)
    ms_per_step, value = main_res["ms_per_step"], main_res["value"]

    # back-to-back throughput without the flush (what a game loop sees), eager and as a CUDA graph; reported, not the headline
    def pipelined(use_graph):
        G.set_use_graph(use_graph)
        G.upload_sand(sand)
        for _ in range(3):
            step(params)
        G.sync()
        reps = min(steps, args.reset_every or steps)
        t0 = time.perf_counter()
        for _ in range(reps):
            step(params)
        G.sync()
        G.set_use_graph(0)
        return (time.perf_counter() - t0) * 1e3 / reps
    pipelined_ms = pipelined(0)
    pipelined_graph_ms = pipelined(1)

    extras = {}
    if not args.no_extras and kind == "fluid":
        # stock pair: the reference's OWN semantics — one solver iteration, lambdas[loop counter] — in parity
        # arithmetic and in throughput arithmetic (the CPU side of the pair is timed with cpu_baseline below)
        sp = {}
        for name, exact in (("b200_exact", 1), ("b200_fast", 0)):
            p1 = lgpu.default_step_params(dt=dt, iterations=1, literal_lambda_index=1, exact_math=exact)
            G.upload_sand(sand)
            ms1 = timed_substeps(G, step, p1, 20, 3, flush, lambda: G.upload_sand(sand), args.reset_every)
            sp[name] = {"ms_per_step": ms1, "value": n / (ms1 * 1e-3)}
        extras["stock_pair"] = {"config": "K=1, literal_lambda_index=1 (src/Simulate.cpp:27-115 as shipped), %d particles" % n, **sp}
    G.close()

    if not args.no_extras and workload == "dam_break_1m":
        # BASELINE configs[2]: the 4M sand pile on voxel solids (driver-visible record of the sand solver)
        sec_steps = max(10, min(steps, 40))
        sec, G2, scene2 = single_gpu_workload("sand_pile_4m", args, local_rank, flush, hbm_gbs, peak_src, sec_steps, 5)
        G2.close()
        sec_line = {"value": sec["value"], "unit": "particle-substeps/s", "ms_per_step": sec["ms_per_step"], "steps": sec_steps,
                    "config": sec["config"], "roofline": sec["roofline"], "roofline_step": sec["roofline_step"],
                    "table_overflows": sec["table_overflows"], "cpu_baseline": None}
        if not args.no_cpu_baseline:
            cbs, _, _, _ = run_cpu_reference("sand", 160, 4, 0.016, steps=2, warmup=1, budget_s=args.cpu_budget)
            sec_line["cpu_baseline"] = cbs
        extras["secondary"] = {"sand_pile_4m": sec_line}
        # BASELINE configs[0] and configs[4] through the kept API (C++ caller / C wrapper)
        extras["secondary"]["fluid_demo"] = fluid_demo_config1(600, not args.no_cpu_baseline)
        extras["secondary"]["mixed_2m"] = mixed_scene_config5(330)

    # ---------------- e2e through the plugin boundary with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        e_steps = min(steps, 20)
        e2e = wrapper_e2e(workload, e_steps, 0)
        if not args.no_extras:
            extras["e2e_device_resident"] = wrapper_e2e(workload, e_steps, 2)

    cpu_baseline = None
    if not args.no_cpu_baseline:
        cpu_baseline, _, _, _ = run_cpu_reference(kind, side, K, dt, steps=3, warmup=1, budget_s=max(args.cpu_budget, 14.0), literal=args.literal)
        if "stock_pair" in extras:
            cs, _, _, _ = run_cpu_reference(kind, side, 1, dt, steps=3, warmup=1, budget_s=max(args.cpu_budget, 14.0), stock=True)
            extras["stock_pair"]["cpu_reference"] = cs

    line = {"metric": metric, "value": value, "unit": "particle-substeps/s", "n_gpus": 1, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": main_res["config"],
            "notes": {"arithmetic": "exact (reference op order)" if args.exact else "fast (FMA + approx rsqrt, parity-tested to 1e-5)",
                      "l2": "flushed between substeps (256 MiB write, outside the event bracket)",
                      "timing": "CUDA events on the launching stream around every substep, summed",
                      "launch": "eager launches" if args.no_graph else "one CUDA graph replay per substep (captured once: lgpu_set_use_graph)",
                      "pipelined_ms_per_step_no_flush": pipelined_ms, "pipelined_ms_per_step_no_flush_cuda_graph": pipelined_graph_ms,
                      "table_overflows": main_res["table_overflows"], "key_violations": main_res["key_violations"]},
            "roofline": main_res["roofline"], "roofline_step": main_res["roofline_step"], "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": main_res["launches"], "clocks": sampler.result()}
    if "free_run" in main_res:
        line["free_run"] = main_res["free_run"]
    line.update(extras)
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
