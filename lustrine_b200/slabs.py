"""Spatial slabs: the particle domain partitioned across the GPUs of one box.

Each rank (one process per GPU) owns the cell columns [x_lo, x_hi) of the reference's uniform grid
and runs the same substep as the single-GPU path on its particles plus a ghost layer of one or two cell
columns (`ghost_columns`; with two, the lambdas of the inner ghost column are computed locally and a fluid substep needs
K - 1 ghost refreshes instead of 2K - 1).
Per substep the neighbouring slabs exchange, with no collective on the data path:
  * after predict: migrants (particles whose predicted cell column left the slab) and ghost copies
    of the particles in the first / last owned column;
  * after every solver pass but the last: the refreshed x* / lambda of the ghosts.
The messages are written by the sender's CUDA kernels straight into the receiver's memory over
NVLink (CUDA IPC-mapped arena) — see lustrine_b200/csrc/lgpu_slab.cu.  torch.distributed is only the
plumbing: it ships the 64-byte IPC handles at start-up, the barriers and the max-over-ranks timing.

This module holds the host logic: choosing the slab boundaries, dealing the particles, wiring the
neighbours (multi-process through IPC handles, or several "virtual ranks" on one device for the
single-GPU tests), gathering the result in particle-id order.
"""
import numpy as np

from . import lgpu

CELL_SCALE = np.float32(3.1)  # kernelRadius = 3.1f * particleRadius = cell size (reference src/Lustrine.cpp:253-254)


def cell_size(radius=0.5, kernel_radius_scale=3.1):
    return np.float32(np.float32(kernel_radius_scale) * np.float32(radius))


def grid_dims(domain, cs):
    """gridN = (int)(domainN / cell_size) + 1 in fp32 (reference src/Lustrine.cpp:261-266)."""
    return tuple(int(np.float32(d) / cs) + 1 for d in domain)


def cell_x(pos, cs):
    """Global cell column of each position: (int)(x / cell_size), IEEE fp32 division, truncation."""
    return (np.ascontiguousarray(pos[:, 0], np.float32) / cs).astype(np.int32)


DEFAULT_MARGIN = 32   # cell columns a cropped plan keeps free beyond the outermost occupied column on either side


def plan_slabs(columns, grid_x, world, min_columns=1, margin=None):
    """Slab boundaries [x_lo, x_hi) per rank, whole cell columns, chosen from the per-column particle histogram so
    that the ranks hold (nearly) equal particle counts.  `columns` = cell column of every particle (any order).

    margin=None: the slabs cover the whole grid [0, grid_x).  margin=m: the plan is CROPPED to the occupied columns
    plus m free columns on either side (never beyond the grid).  A slab's scan, brick descriptors and histogram cost
    per CELL, and the empty part of the domain used to land on the last rank: 36.7 M cells against 2.3 M on the
    others for the 16 M dam break on 8 ranks (the fluid fills the first third of the domain), which made that rank
    the critical path of every substep.  A cropped plan must be re-planned before the particles reach its end:
    guard_columns(margin) is the width of the warning zone a slab context counts (lgpu_slab_edge)."""
    if world < 1 or grid_x < world * min_columns:
        raise ValueError("need at least %d cell column(s) per rank (grid_x=%d, world=%d)" % (min_columns, grid_x, world))
    hist = np.bincount(np.clip(columns, 0, grid_x - 1), minlength=grid_x).astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(hist)])
    total = int(cum[-1])
    first, last = 0, grid_x
    if margin is not None and total > 0:
        occ = np.nonzero(hist)[0]
        first = max(int(occ[0]) - int(margin), 0)
        last = min(int(occ[-1]) + 1 + int(margin), grid_x)
        short = world * min_columns - (last - first)   # (every rank still owns min_columns columns)
        if short > 0:
            first = max(first - short, 0)
            last = min(first + world * min_columns, grid_x)
    bounds = [first]
    for k in range(1, world):
        target = total * k / world
        x = int(np.searchsorted(cum, target, side="left"))  # first boundary with cum >= target
        # choose the closer of x-1 / x, keep at least one column per rank on both sides
        if x > 0 and abs(cum[x - 1] - target) <= abs(cum[min(x, grid_x)] - target):
            x -= 1
        x = max(x, bounds[-1] + min_columns)        # (a slab owns at least as many columns as the ghost layer is wide)
        x = min(x, last - (world - k) * min_columns)
        bounds.append(x)
    bounds.append(last)
    return [(bounds[k], bounds[k + 1]) for k in range(world)]


def guard_columns(margin):
    """Width of the warning zone at the open end of a cropped plan: half the margin (the plan is renewed when the front
    has covered half of the free columns; the other half is what the particles may cover until the next check)."""
    return 0 if margin is None else max(int(margin) // 2, 1)


def deal(pos, slabs, cs):
    """Index arrays of the particles each slab owns at start (by the cell column of their position)."""
    cx = cell_x(pos, cs)
    out = []
    for k, (lo, hi) in enumerate(slabs):
        lo_eff = -(1 << 30) if k == 0 else lo
        hi_eff = (1 << 30) if k == len(slabs) - 1 else hi
        out.append(np.nonzero((cx >= lo_eff) & (cx < hi_eff))[0].astype(np.int32))
    return out


def slab_capacity(pos, slabs, owned, cs, grid_x, factor=1.5, ghost_columns=1):
    """Particle capacity of every slab context (equal on all ranks).  The storage of a substep holds the
    previous substep's slots (owned + ghost copies, the latter dead by then) plus the incoming ghost copies
    and migrants, so the ghost columns on both sides count twice; `factor` is the head room for
    particles gathering in a slab before the boundaries are re-planned."""
    hist = np.bincount(np.clip(cell_x(pos, cs), 0, grid_x - 1), minlength=grid_x).astype(np.int64)
    need = 0
    for (lo, hi), o in zip(slabs, owned):
        ghosts = int(hist[max(lo - ghost_columns, 0):lo].sum()) + int(hist[hi:min(hi + ghost_columns, grid_x)].sum())
        need = max(need, len(o) + 2 * ghosts)
    return int(need * factor) + 4096


def imbalance(owned_counts):
    """max / mean of the slabs' particle counts (1.0 = perfectly balanced)."""
    c = np.asarray(owned_counts, np.float64)
    return float(c.max() / max(c.mean(), 1.0))


def needs_replan(owned_counts, live_counts, capacity, max_imbalance=1.25, fill=0.8, near_edge=0):
    """Re-plan when a slab holds `max_imbalance` times the mean, when a slab's storage (owned + ghosts, counted
    twice like slab_capacity does) approaches its capacity, or when particles have entered the guard columns at the
    open end of a cropped plan — whichever comes first."""
    return imbalance(owned_counts) > max_imbalance or max(live_counts) > fill * capacity or near_edge > 0


def merge_by_id(parts, n_total):
    """Reassembles (pos, vel, flags, ids) tuples of all slabs into arrays indexed by particle id."""
    pos = np.zeros((n_total, 3), np.float32)
    vel = np.zeros((n_total, 3), np.float32)
    flags = np.zeros(n_total, np.int32)
    seen = np.zeros(n_total, np.int32)
    for p, v, f, ids in parts:
        pos[ids] = p
        vel[ids] = v
        flags[ids] = f
        np.add.at(seen, ids, 1)
    if not np.all(seen == 1):
        raise RuntimeError("slab gather: %d particles missing, %d duplicated" % (int((seen == 0).sum()), int((seen > 1).sum())))
    return pos, vel, flags


def local_to_global_keys(keys, info, grid):
    """Cell ids of a slab context (local grid) as ids of the reference's global grid."""
    gX, gY, gZ = grid
    lgx = info["local_grid_x"]
    cy, rem = np.divmod(keys, lgx * gZ)
    cxl, cz = np.divmod(rem, gZ)
    return cy * (gX * gZ) + (cxl + info["x_off"]) * gZ + cz


class SlabContext:
    """One slab = one lgpu context plus what the host needs to know about it."""

    def __init__(self, domain, slab, capacity, solids=None, device=-1, halo_capacity=0, **ctx_kw):
        n_solid = 0 if solids is None else len(solids)
        self.G = lgpu.Context(domain, capacity_sand=int(capacity), capacity_solid=n_solid, device=device, slab=slab,
                              halo_capacity=halo_capacity, **ctx_kw)
        if n_solid:
            self.G.upload_solids(solids)  # replicated; the context keeps the solids inside its columns
        self.slab = slab

    def close(self):
        self.G.close()


class VirtualSlabs:
    """P slabs on ONE device, driven by one host thread (tests, single-GPU debugging): same kernels,
    same messages, neighbours' arenas addressed by plain device pointers instead of IPC handles."""

    def __init__(self, domain, pos, world, solids=None, vel=None, flags=None, device=-1, capacity_factor=1.5, halo_capacity=0,
                 slabs=None, margin=DEFAULT_MARGIN, **ctx_kw):
        pos = np.ascontiguousarray(pos, np.float32)
        self.n_total = len(pos)
        self.world = world
        self.margin = margin
        ctx_kw = dict(ctx_kw, guard_columns=guard_columns(margin))
        self._args = (domain, solids, device, capacity_factor, halo_capacity, ctx_kw)
        self.grid = grid_dims(domain, cell_size())
        self.ctx = []
        self.replans = 0
        self._build(pos, vel, flags, None, slabs)

    def _build(self, pos, vel, flags, ids, slabs=None):
        domain, solids, device, capacity_factor, halo_capacity, ctx_kw = self._args
        cs = cell_size()
        gw = max(int(ctx_kw.get("ghost_columns", 0)), 1)
        self.slabs = slabs or plan_slabs(cell_x(pos, cs), self.grid[0], self.world, gw, self.margin)
        owned = deal(pos, self.slabs, cs)
        self.capacity = slab_capacity(pos, self.slabs, owned, cs, self.grid[0], capacity_factor, gw)
        self.ctx = [SlabContext(domain, s, self.capacity, solids, device, halo_capacity, **ctx_kw) for s in self.slabs]
        exports = [c.G.slab_export() for c in self.ctx]
        for k, c in enumerate(self.ctx):
            if k > 0:
                c.G.slab_connect(0, same_process_ptr=exports[k - 1][1])
            if k + 1 < len(self.ctx):
                c.G.slab_connect(1, same_process_ptr=exports[k + 1][1])
        for c, idx in zip(self.ctx, owned):
            c.G.slab_upload(pos[idx], idx if ids is None else ids[idx], None if vel is None else vel[idx], None if flags is None else flags[idx])

    def step(self, mode, params=None, **kw):
        p = params if params is not None else lgpu.default_step_params(**kw)
        for c in self.ctx:
            c.G.slab_step_begin(p, mode)
        for c in self.ctx:
            c.G.slab_step_end()

    def counts(self):
        info = [c.G.slab_info() for c in self.ctx]
        return [i["owned"] for i in info], [i["owned"] + 2 * i["ghosts"] for i in info]

    def near_edge(self):
        """Particles inside the guard columns at the open ends of a cropped plan (last substep)."""
        return sum(c.G.slab_edge()[0] for c in self.ctx)

    def replan_if_needed(self, max_imbalance=1.25, fill=0.8):
        """Re-plans the slab boundaries from the current per-column histogram when the particles have gathered in a few
        slabs or approach the open end of a cropped plan: all particles are collected, dealt again and
        uploaded into new contexts.  A rare, slow step (the boundaries are fixed between re-plans); returns True if it
        happened."""
        owned, live = self.counts()
        if not needs_replan(owned, live, self.capacity, max_imbalance, fill, self.near_edge()):
            return False
        pos, vel, flags = self.gather()
        for c in self.ctx:
            c.close()
        self._build(pos, vel, flags, np.arange(self.n_total, dtype=np.int32))
        self.replans += 1
        return True

    def sync(self):
        for c in self.ctx:
            c.G.sync()

    def gather(self):
        return merge_by_id([c.G.slab_download() for c in self.ctx], self.n_total)

    def close(self):
        for c in self.ctx:
            c.close()


class DistributedSlab:
    """This rank's slab of a torch.distributed job (one process per GPU, NCCL or gloo for the plumbing)."""

    def __init__(self, domain, pos, solids=None, vel=None, flags=None, device=0, capacity_factor=1.5, halo_capacity=0,
                 context_factory=None, margin=DEFAULT_MARGIN, **ctx_kw):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        pos = np.ascontiguousarray(pos, np.float32)
        self.n_total = len(pos)
        self.margin = margin
        ctx_kw = dict(ctx_kw, guard_columns=guard_columns(margin))
        self._args = (domain, solids, device, capacity_factor, halo_capacity, context_factory or SlabContext, ctx_kw)
        self.grid = grid_dims(domain, cell_size())
        self.ctx = None
        self.replans = 0
        # every rank holds the same synthetic scene and derives the same plan: no broadcast needed
        idx = self._build(pos, vel, flags, None)
        self.initial = (pos[idx].copy(), idx.copy())
        dist.barrier()

    def _build(self, pos, vel, flags, ids):
        """Plans the slabs from `pos` (the same array on every rank), creates this rank's context, wires the neighbours
        through CUDA IPC handles and uploads this rank's particles.  Returns the indices this rank owns."""
        domain, solids, device, capacity_factor, halo_capacity, factory, ctx_kw = self._args
        cs = cell_size()
        gw = max(int(ctx_kw.get("ghost_columns", 0)), 1)
        self.slabs = plan_slabs(cell_x(pos, cs), self.grid[0], self.world, gw, self.margin)
        owned = deal(pos, self.slabs, cs)
        self.capacity = slab_capacity(pos, self.slabs, owned, cs, self.grid[0], capacity_factor, gw)
        self.ctx = factory(domain, self.slabs[self.rank], self.capacity, solids, device, halo_capacity, **ctx_kw)
        handle, _, _ = self.ctx.G.slab_export()
        handles = [None] * self.world
        self.dist.all_gather_object(handles, handle)
        if self.rank > 0:
            self.ctx.G.slab_connect(0, handle=handles[self.rank - 1])
        if self.rank + 1 < self.world:
            self.ctx.G.slab_connect(1, handle=handles[self.rank + 1])
        idx = owned[self.rank]
        self.ctx.G.slab_upload(pos[idx], idx if ids is None else ids[idx], None if vel is None else vel[idx], None if flags is None else flags[idx])
        return idx

    @property
    def G(self):
        return self.ctx.G

    def reset(self):
        """Puts the slab back to its initial particles (all ranks must call it together; not after a re-plan)."""
        self.ctx.G.slab_upload(self.initial[0], self.initial[1])

    def step(self, mode, params):
        if mode == 1:
            self.ctx.G.step_fluid(params)
        else:
            self.ctx.G.step_sand(params)

    def counts(self, with_edge=False):
        """(owned, storage need[, particles near the open end of a cropped plan]) of every rank — one small all_gather,
        no device data."""
        info = self.ctx.G.slab_info()
        mine = (int(info["owned"]), int(info["owned"] + 2 * info["ghosts"]), int(self.ctx.G.slab_edge()[0]))
        allc = [None] * self.world
        self.dist.all_gather_object(allc, mine)
        out = [c[0] for c in allc], [c[1] for c in allc]
        return out + (sum(c[2] for c in allc),) if with_edge else out

    def replan_if_needed(self, max_imbalance=1.25, fill=0.8):
        """Collective: re-plans the slab boundaries from the current particle columns when the load has drifted
        (SURVEY §8e "rebalanced every R steps from the per-x-plane histogram").  Every rank contributes its particles
        (all_gather over the process group), derives the same new plan, re-creates its context and re-wires its
        neighbours.  Rare and slow by design: between re-plans the boundaries are fixed and the data path has no
        collective.  Returns True if it happened."""
        owned, live, edge = self.counts(with_edge=True)
        if not needs_replan(owned, live, self.capacity, max_imbalance, fill, edge):
            return False
        pos, vel, flags = self.gather()
        self.dist.barrier()        # nobody re-maps a neighbour's arena while it is still being read
        self.ctx.close()
        self._build(pos, vel, flags, np.arange(self.n_total, dtype=np.int32))
        self.replans += 1
        self.dist.barrier()
        return True

    def gather(self):
        """All particles in id order on every rank."""
        parts = [None] * self.world
        self.dist.all_gather_object(parts, self.ctx.G.slab_download())
        return merge_by_id(parts, self.n_total)

    def close(self):
        self.ctx.close()
