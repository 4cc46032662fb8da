"""Spatial slabs: host logic on CPU (incl. a world_size-2 gloo run), and on the GPU the
slab protocol itself with several "virtual ranks" on one device against the single-context run."""
import os
import subprocess
import sys

import numpy as np
import pytest

import scenes
from lustrine_b200 import lgpu, slabs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ host logic (CPU)
def test_plan_covers_grid_and_balances():
    domain, pos = scenes.dam_break(20)
    cs = slabs.cell_size()
    grid = slabs.grid_dims(domain, cs)
    assert grid == (39, 26, 26)  # N=20 -> grid=39x26x26
    cx = slabs.cell_x(pos, cs)
    for world in (1, 2, 3, 4, 8):
        plan = slabs.plan_slabs(cx, grid[0], world)
        assert plan[0][0] == 0 and plan[-1][1] == grid[0]
        assert all(a[1] == b[0] for a, b in zip(plan, plan[1:])) and all(hi > lo for lo, hi in plan)
        owned = slabs.deal(pos, plan, cs)
        assert sum(len(o) for o in owned) == len(pos) and len(np.unique(np.concatenate(owned))) == len(pos)
        counts = np.array([len(o) for o in owned])
        # whole columns only: the imbalance is bounded by the fullest column
        assert counts.max() - counts.min() <= 2 * np.bincount(cx).max()
    with pytest.raises(ValueError):
        slabs.plan_slabs(cx, 3, 4)


def test_plan_edge_cases():
    # all particles in one column: every rank still gets at least one column
    plan = slabs.plan_slabs(np.full(100, 5, np.int32), 12, 4)
    assert plan[0][0] == 0 and plan[-1][1] == 12 and all(hi > lo for lo, hi in plan)
    # no particles
    plan = slabs.plan_slabs(np.zeros(0, np.int32), 8, 2)
    assert plan[0][0] == 0 and plan[-1][1] == 8
    # out-of-grid columns are clipped into the edge slabs
    pos = np.array([[-3.0, 1, 1], [1000.0, 1, 1], [5.0, 1, 1]], np.float32)
    owned = slabs.deal(pos, [(0, 4), (4, 9)], slabs.cell_size())
    assert owned[0].tolist() == [0, 2] or owned[0].tolist() == [0] and owned[1].tolist() == [1, 2]
    assert sorted(np.concatenate(owned).tolist()) == [0, 1, 2]


def test_cropped_plan_leaves_the_empty_part_of_the_domain_out():
    """The 16 M dam break fills the first third of its domain: planned over the whole grid, the last of 8 ranks owns
    two thirds of the cell columns (and their scan / brick descriptors / histogram); a cropped plan stops `margin`
    columns past the outermost occupied column."""
    grid_x = 488
    cols = np.repeat(np.arange(0, 164), 1000)  # columns 0..163 occupied (252 lattice planes of spacing 1 in cells of 1.55)
    full = slabs.plan_slabs(cols, grid_x, 8)
    assert full[0][0] == 0 and full[-1][1] == grid_x and full[-1][1] - full[-1][0] > 300
    plan = slabs.plan_slabs(cols, grid_x, 8, margin=32)
    assert plan[0][0] == 0 and plan[-1][1] == 164 + 32
    assert all(a[1] == b[0] for a, b in zip(plan, plan[1:])) and all(hi > lo for lo, hi in plan)
    assert [p[0] for p in plan[1:]] == [p[0] for p in full[1:]]          # the inner boundaries are the balanced ones
    assert max(hi - lo for lo, hi in plan) <= 21 + 32
    # the occupied block in the middle of the grid: cropped on both sides; never beyond the grid
    mid = slabs.plan_slabs(cols + 100, grid_x, 4, margin=16)
    assert mid[0][0] == 100 - 16 and mid[-1][1] == 264 + 16
    assert slabs.plan_slabs(cols + 100, 270, 4, margin=16)[-1][1] == 270
    # a block narrower than the ranks need: the plan is widened to min_columns per rank
    thin = slabs.plan_slabs(np.full(50, 7), 40, 4, min_columns=2, margin=0)
    assert all(hi - lo >= 2 for lo, hi in thin) and thin[0][0] >= 0 and thin[-1][1] <= 40
    owned = slabs.deal(np.array([[7 * 1.55 + 0.1, 1, 1]] * 50, np.float32), thin, slabs.cell_size())
    assert sum(len(o) for o in owned) == 50
    assert slabs.guard_columns(None) == 0 and slabs.guard_columns(32) == 16 and slabs.guard_columns(1) == 1
    # particles in the guard columns of an open end ask for a new plan
    assert not slabs.needs_replan([100, 100], [120, 120], 1000)
    assert slabs.needs_replan([100, 100], [120, 120], 1000, near_edge=3)


def test_c_planner_equals_the_python_planner():
    """include/lgpu.h offers the planning to a C / C++ host (lgpu_plan_slabs, lgpu_slab_capacity, lgpu_slab_guard_columns);
    it must decide exactly what lustrine_b200/slabs.py decides, on the BASELINE layouts and on random histograms."""
    rng = np.random.default_rng(7)
    cs = slabs.cell_size()
    cases = []
    domain, pos = scenes.dam_break(20)
    grid = slabs.grid_dims(domain, cs)
    cases.append((slabs.cell_x(pos, cs), grid[0]))
    cases.append((np.repeat(np.arange(0, 164), 50), 488))                       # the 16 M dam break's column layout
    cases.append((np.repeat(np.arange(0, 164), 50) + 100, 488))
    for _ in range(60):
        gx = int(rng.integers(8, 300))
        lo = int(rng.integers(0, gx)); hi = int(rng.integers(lo, gx))
        n = int(rng.integers(1, 4000))
        cols = rng.integers(lo, hi + 1, n)
        if rng.random() < 0.3:
            cols = np.concatenate([cols, np.full(int(rng.integers(1, 3000)), int(rng.integers(lo, hi + 1)))])  # one crowded column
        cases.append((cols, gx))
    checked = 0
    for cols, gx in cases:
        hist = np.bincount(np.clip(cols, 0, gx - 1), minlength=gx)
        for world in (1, 2, 3, 4, 8):
            for min_columns in (1, 2):
                if gx < world * min_columns:
                    continue
                for margin in (None, 0, 3, 32):
                    want = slabs.plan_slabs(cols, gx, world, min_columns, margin)
                    got = lgpu.plan_slabs_c(hist, world, min_columns, margin)
                    assert got == want, (gx, world, min_columns, margin, got, want)
                    checked += 1
                    # capacity: the Python helper works from positions, the C one from the histogram
                    owned = [int(hist[(0 if k == 0 else lo):(gx if k == world - 1 else hi)].sum()) for k, (lo, hi) in enumerate(want)]
                    need = 0
                    for (lo, hi), o in zip(want, owned):
                        g = int(hist[max(lo - min_columns, 0):lo].sum()) + int(hist[hi:min(hi + min_columns, gx)].sum())
                        need = max(need, o + 2 * g)
                    assert lgpu.slab_capacity_c(hist, want, min_columns, 1.5) == int(need * 1.5) + 4096
    assert checked > 1000
    assert [lgpu.lib().lgpu_slab_guard_columns(m) for m in (-1, 0, 1, 2, 32)] == [slabs.guard_columns(m) for m in (None, 0, 1, 2, 32)]
    with pytest.raises(lgpu.LgpuError):
        lgpu.plan_slabs_c(np.ones(3, np.int64), 4)


def test_merge_by_id_detects_loss_and_duplicates():
    a = (np.ones((2, 3), np.float32), np.zeros((2, 3), np.float32), np.zeros(2, np.int32), np.array([0, 2], np.int32))
    b = (np.ones((1, 3), np.float32) * 2, np.zeros((1, 3), np.float32), np.ones(1, np.int32), np.array([1], np.int32))
    pos, vel, flags = slabs.merge_by_id([a, b], 3)
    assert pos[:, 0].tolist() == [1, 2, 1] and flags.tolist() == [0, 1, 0]
    with pytest.raises(RuntimeError):
        slabs.merge_by_id([a], 3)
    with pytest.raises(RuntimeError):
        slabs.merge_by_id([a, a, b], 3)


def test_local_to_global_keys():
    grid = (39, 26, 26)
    info = dict(local_grid_x=7, x_off=9)
    cy, cx, cz = 3, 11, 20
    local = cy * 7 * 26 + (cx - 9) * 26 + cz
    assert slabs.local_to_global_keys(np.array([local]), info, grid)[0] == cy * 39 * 26 + cx * 26 + cz


GLOO_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch.distributed as dist
import scenes
from lustrine_b200 import slabs

class FakeG:
    """Stands in for the CUDA context: remembers what it was told, returns its particles unchanged."""
    def __init__(self, rank): self.rank, self.connected, self.state = rank, {{}}, None
    def slab_export(self): return (bytes([self.rank]) * 64, 0, 0)
    def slab_connect(self, side, handle=None, same_process_ptr=None): self.connected[side] = handle[0]
    def slab_upload(self, pos, ids, vel=None, flags=None): self.state = (pos.copy(), ids.copy())
    def slab_download(self):
        pos, ids = self.state
        return pos, np.zeros_like(pos), np.zeros(len(ids), np.int32), ids
    def slab_info(self): return {{"owned": len(self.state[1]), "ghosts": 0}}
    def slab_edge(self): return (0, False)
    def close(self): pass

class FakeCtx:
    def __init__(self, domain, slab, capacity, solids, device, halo_capacity, **kw):
        self.G, self.slab, self.capacity = FakeG(dist.get_rank()), slab, capacity
    def close(self): pass

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
domain, pos = scenes.dam_break(16)
S = slabs.DistributedSlab(domain, pos, context_factory=FakeCtx)
assert S.slabs == slabs.plan_slabs(slabs.cell_x(pos, slabs.cell_size()), S.grid[0], world, 1, slabs.DEFAULT_MARGIN)
# neighbours wired left/right with the right handles
want = {{}}
if rank > 0: want[0] = rank - 1
if rank + 1 < world: want[1] = rank + 1
assert S.G.connected == want, (S.G.connected, want)
# every particle is owned exactly once and comes back in id order
got, vel, flags = S.gather()
assert np.array_equal(got, pos)
counts = [None] * world
dist.all_gather_object(counts, len(S.initial[1]))
assert sum(counts) == len(pos) and S.ctx.capacity >= max(counts)
assert not S.replan_if_needed(), "a balanced deal must not be re-planned"
# the particles drift to +x: after the migrations of many substeps the last slab owns most of them
moved = pos + np.array([0.45 * domain[0], 0.0, 0.0], np.float32)
mine = slabs.deal(moved, S.slabs, slabs.cell_size())[rank]
S.G.state = (moved[mine].copy(), mine.copy())
before = list(S.slabs)
owned, live = S.counts()
assert sum(owned) == len(pos) and slabs.imbalance(owned) > 1.25, owned
assert S.replan_if_needed() and S.replans == 1
assert S.slabs != before and S.slabs == slabs.plan_slabs(slabs.cell_x(moved, slabs.cell_size()), S.grid[0], world, 1, slabs.DEFAULT_MARGIN)
assert S.G.connected == want, "neighbours re-wired after the re-plan"
owned2, _ = S.counts()
assert sum(owned2) == len(pos) and slabs.imbalance(owned2) < 1.1, owned2
got2, _, _ = S.gather()
assert np.array_equal(got2, moved)
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("slab-rank-%d-ok %s %s\n" % (rank, counts, owned2)); sys.stdout.flush()
'''


def test_distributed_host_logic_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("slab-rank-") == 2 and out.stdout.count("-ok") == 2, out.stdout


def test_slab_capacity_counts_the_ghost_columns_twice():
    """Storage of a substep = previous slots (owned + dead ghost copies) + incoming ghost copies and migrants."""
    cs = slabs.cell_size()
    domain, pos = scenes.dam_break(16)
    grid = slabs.grid_dims(domain, cs)
    for world in (2, 4, 8):
        plan = slabs.plan_slabs(slabs.cell_x(pos, cs), grid[0], world)
        owned = slabs.deal(pos, plan, cs)
        cap = slabs.slab_capacity(pos, plan, owned, cs, grid[0], factor=1.0)
        hist = np.bincount(slabs.cell_x(pos, cs), minlength=grid[0])
        worst = 0
        for (lo, hi), o in zip(plan, owned):
            ghosts = (hist[lo - 1] if lo > 0 else 0) + (hist[hi] if hi < grid[0] else 0)
            worst = max(worst, len(o) + 2 * ghosts)
        assert cap == worst + 4096
        assert cap >= max(len(o) for o in owned)
    # thin slabs (a few columns each): the ghost columns dominate and a plain 1.5 x owned would not hold a substep
    plan = slabs.plan_slabs(slabs.cell_x(pos, cs), grid[0], 8)
    owned = slabs.deal(pos, plan, cs)
    assert slabs.slab_capacity(pos, plan, owned, cs, grid[0], factor=1.5) > int(1.5 * max(len(o) for o in owned)) + 4096


# ------------------------------------------------------------------ the protocol on one GPU
REL_TOL, ABS_TOL = 1e-5, 1e-5


def close(a, b, what, rtol=REL_TOL, atol=ABS_TOL):
    err = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))
    bound = atol + rtol * np.abs(np.asarray(b, np.float64))
    worst = float((err / bound).max()) if err.size else 0.0
    print("  %-10s max|err| %.3e  worst err/bound %.3f" % (what, err.max() if err.size else 0.0, worst))
    assert worst <= 1.0, what


@pytest.mark.gpu
@pytest.mark.parametrize("world,gw", [(2, 1), (3, 1), (2, 2), (3, 2)])
def test_fluid_slabs_match_single_context(world, gw):
    """P slabs == 1 slab == no slabs, same tolerances as the oracle parity (summation order only).  gw = width of the
    ghost layer in cell columns: with 2 the inner ghosts' lambdas are computed locally (their lists are the owners'
    lists, their neighbours all present) and no lambda travels between the slabs."""
    domain, pos = scenes.dam_break(16)
    solids = scenes.floor_plate(30, 20)
    vel0 = np.zeros_like(pos)
    vel0[:, 0] = 12.0 * np.sin(pos[:, 2])  # sideways motion: particles cross the slab boundaries
    kw = dict(dt=0.01, iterations=3, literal_lambda_index=0, exact_math=1)
    with lgpu.Context(domain, capacity_sand=len(pos), capacity_solid=len(solids)) as G:
        G.upload_sand(pos, vel0)
        G.upload_solids(solids)
        V = slabs.VirtualSlabs(domain, pos, world, solids=solids, vel=vel0, ghost_columns=gw)
        moved = 0
        for step in range(6):
            G.step_fluid(**kw)
            V.step(1, **kw)
            rp, rv, _ = G.download()
            sp, sv, _ = V.gather()
            print("step", step, [c.G.slab_info()["owned"] for c in V.ctx], [c.G.slab_info()["ghosts"] for c in V.ctx])
            close(sp, rp, "position")
            close(sv, rv, "velocity", atol=1e-3)
            owners = np.concatenate([np.full(c.G.n, k) for k, c in enumerate(V.ctx)])
            moved = max(moved, abs(int((owners == 0).sum()) - len(slabs.deal(pos, V.slabs, slabs.cell_size())[0])))
        assert moved > 0, "the test scene must make particles migrate"
        assert sum(c.G.n for c in V.ctx) == len(pos)
        V.close()


@pytest.mark.gpu
@pytest.mark.parametrize("gw", [1, 2])
def test_sand_slabs_match_single_context(gw):
    domain, sand, solids = scenes.sand_pile(12, drop=1.0)
    vel0 = np.zeros_like(sand)
    vel0[:, 0] = 6.0 * np.cos(sand[:, 1])
    kw = dict(dt=0.016, iterations=4, exact_math=1)
    with lgpu.Context(domain, capacity_sand=len(sand), capacity_solid=len(solids)) as G:
        G.upload_sand(sand, vel0)
        G.upload_solids(solids)
        V = slabs.VirtualSlabs(domain, sand, 3, solids=solids, vel=vel0, ghost_columns=gw)
        ids = np.arange(len(sand))
        for step in range(5):
            G.step_sand(**kw)
            V.step(2, **kw)
            # the single context permutes its storage like the reference: follow the particles
            ids = ids[G.dump(lgpu.DUMP_PERM)]
            rp, rv, _ = G.download()
            sp, sv, _ = V.gather()
            close(sp[ids], rp, "position")
            close(sv[ids], rv, "velocity", atol=1e-3)
        V.close()


@pytest.mark.gpu
@pytest.mark.parametrize("gw", [1, 2])
def test_slab_keys_and_neighbour_counts_are_the_single_gpu_ones(gw):
    domain, pos = scenes.dam_break(12)
    kw = dict(dt=0.01, iterations=1, literal_lambda_index=0, exact_math=1)
    with lgpu.Context(domain, capacity_sand=len(pos)) as G:
        G.upload_sand(pos)
        G.step_fluid(**kw)
        keys = G.dump(lgpu.DUMP_KEYS); orig = G.dump(lgpu.DUMP_ORIG); cnt = G.dump(lgpu.DUMP_NBR_COUNT)
        ref_key = np.zeros(len(pos), np.int64); ref_key[orig] = keys
        ref_cnt = np.zeros(len(pos), np.int64); ref_cnt[orig] = cnt
    V = slabs.VirtualSlabs(domain, pos, 2, ghost_columns=gw)
    V.step(1, **kw)
    seen = 0
    for c in V.ctx:
        info = c.G.slab_info()
        n_live = info["owned"] + info["ghosts"]
        # dumps cover the sorted live particles (owned + ghosts); ghosts have no list of their own
        k = np.zeros(n_live, np.int32); o = np.zeros(n_live, np.int32); w = np.zeros(n_live, np.int32)
        for what, arr in ((lgpu.DUMP_KEYS, k), (lgpu.DUMP_ORIG, o), (lgpu.DUMP_NBR_COUNT, w)):
            lgpu._check(c.G.L.lgpu_dump(c.G._h, what, arr.ctypes.data, arr.nbytes), "lgpu_dump")
        gk = slabs.local_to_global_keys(k.astype(np.int64), info, V.grid)
        assert np.array_equal(gk, ref_key[o]), "cell keys of a slab are the global ones"
        col = (gk // V.grid[2]) % V.grid[0]
        own = (col >= info["x_lo"]) & (col < info["x_hi"])
        assert own.sum() == info["owned"]
        assert np.array_equal(w[own], ref_cnt[o[own]]), "neighbour counts of owned particles"
        seen += int(own.sum())
    assert seen == len(pos)
    V.close()


@pytest.mark.gpu
def test_literal_lambda_index_is_refused_on_slabs():
    domain, pos = scenes.dam_break(8)
    V = slabs.VirtualSlabs(domain, pos, 2)
    with pytest.raises(lgpu.LgpuError):
        V.step(1, dt=0.01, iterations=1, literal_lambda_index=1)
    V.close()


@pytest.mark.gpu
@pytest.mark.parametrize("gw", [1, 2])
def test_free_running_slabs_replan_instead_of_overflowing(gw):
    """A free-running dam break drains the slabs on the left into the slab on the right.  With the boundaries planned
    once the right slab outgrows its capacity (LGPU_ERR_CAPACITY); re-planning from the current per-column histogram
     keeps the run going, and the particles stay those of the single-context run."""
    domain, pos = scenes.dam_break(24)
    kw = dict(dt=0.01, iterations=2, literal_lambda_index=0, exact_math=1)
    steps, every = 240, 10
    with lgpu.Context(domain, capacity_sand=len(pos)) as G:
        G.upload_sand(pos)
        fixed = slabs.VirtualSlabs(domain, pos, 4, capacity_factor=1.1, ghost_columns=gw)
        V = slabs.VirtualSlabs(domain, pos, 4, capacity_factor=1.1, ghost_columns=gw)
        overflowed = None
        for step in range(steps):
            G.step_fluid(**kw)
            V.step(1, **kw)
            if fixed is not None:
                try:
                    fixed.step(1, **kw)
                    fixed.sync()
                except lgpu.LgpuError as e:
                    overflowed = (step, str(e))
                    fixed = None
            if (step + 1) % every == 0 and V.replan_if_needed():
                print("step", step + 1, "re-planned:", V.slabs, V.counts()[0])
        owned, _ = V.counts()
        fixed_owned = fixed.counts()[0] if fixed is not None else None
        print("re-plans", V.replans, "final owned", owned, "| fixed plan:", fixed_owned, "overflowed at", overflowed)
        assert V.replans >= 1 and slabs.imbalance(owned) < 1.6
        # the plan made once either ran out of capacity or ended far more lopsided than the re-planned one
        if overflowed is not None:
            assert "capacity" in overflowed[1].lower()
        else:
            assert slabs.imbalance(fixed_owned) > 1.5 * slabs.imbalance(owned), (fixed_owned, owned)
            fixed.close()
        rp, rv, _ = G.download()
        sp, sv, _ = V.gather()
        close(sp, rp, "position")
        close(sv, rv, "velocity", atol=1e-3)
        V.close()


@pytest.mark.gpu
def test_cropped_plan_is_renewed_before_the_front_reaches_its_end():
    """A plan cropped to the occupied columns (+ a small margin here) on a free-running dam break: the contexts count
    the particles that enter the guard columns at the open end, replan_if_needed renews the plan, and the run stays the
    single-context run.  Without re-planning the step fails loudly instead of binning particles into the wrong cells."""
    domain, pos = scenes.dam_break(20)
    kw = dict(dt=0.01, iterations=2, literal_lambda_index=0, exact_math=1)
    grid_x = slabs.grid_dims(domain, slabs.cell_size())[0]
    with lgpu.Context(domain, capacity_sand=len(pos)) as G:
        G.upload_sand(pos)
        V = slabs.VirtualSlabs(domain, pos, 3, capacity_factor=1.3, margin=6)
        first_plan = list(V.slabs)
        assert first_plan[-1][1] < grid_x and V.ctx[-1].G.slab_edge() == (0, True) and V.ctx[0].G.slab_edge() == (0, False)
        cells_cropped = sum(c.G.slab_info()["local_cells"] for c in V.ctx)
        full = slabs.VirtualSlabs(domain, pos, 3, capacity_factor=1.3, margin=None)
        assert sum(c.G.slab_info()["local_cells"] for c in full.ctx) > cells_cropped
        full.close()
        for step in range(120):
            G.step_fluid(**kw)
            V.step(1, **kw)
            if (step + 1) % 4 == 0 and V.replan_if_needed():
                print("step", step + 1, "re-planned:", V.slabs, "near edge", V.near_edge())
        assert V.replans >= 1 and V.slabs[-1][1] > first_plan[-1][1]
        rp, rv, _ = G.download()
        sp, sv, _ = V.gather()
        close(sp, rp, "position")
        close(sv, rv, "velocity", atol=1e-3)
        V.close()
    # never re-planned: the front runs out of the planned columns and the step says so
    V = slabs.VirtualSlabs(domain, pos, 2, capacity_factor=2.0, margin=2)
    with pytest.raises(lgpu.LgpuError, match="left the planned cell columns"):
        for step in range(400):
            V.step(1, **kw)
            V.sync()
    V.close()


def test_plan_keeps_slabs_as_wide_as_the_ghost_layer():
    """With a two-column ghost layer every slab owns at least two columns (a narrower slab could not fill its
    neighbours' outer ghost column), and the capacity estimate counts both ghost columns."""
    rng = np.random.default_rng(3)
    cols = np.concatenate([np.full(5000, 7), rng.integers(0, 40, 300)])  # nearly everything in one column
    for world in (2, 4, 8):
        plan = slabs.plan_slabs(cols, 40, world, min_columns=2)
        assert plan[0][0] == 0 and plan[-1][1] == 40
        assert all(hi - lo >= 2 for lo, hi in plan), plan
        assert all(plan[k][1] == plan[k + 1][0] for k in range(world - 1))
    with pytest.raises(ValueError):
        slabs.plan_slabs(cols, 7, 4, min_columns=2)
    domain, pos = scenes.dam_break(16)
    cs = slabs.cell_size()
    grid = slabs.grid_dims(domain, cs)
    plan = slabs.plan_slabs(slabs.cell_x(pos, cs), grid[0], 3, min_columns=2)
    owned = slabs.deal(pos, plan, cs)
    assert slabs.slab_capacity(pos, plan, owned, cs, grid[0], ghost_columns=2) > slabs.slab_capacity(pos, plan, owned, cs, grid[0], ghost_columns=1)
