"""ctypes binding of the C ABI in include/lgpu.h (liblgpu.so, hand-written CUDA for sm_100a).

This is the Python-side mirror used by the tests, bench.py and the multi-GPU driver.  There
is no CPU fallback: importing works anywhere (so that symbol/ABI checks can run without a GPU)
but every compute call goes to the CUDA library and raises ``LgpuError`` if it fails.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LGPU_LIB") or os.path.join(_HERE, "lib", "liblgpu.so")  # LGPU_LIB: tuning builds only

c_f, c_i = C.c_float, C.c_int


class LgpuError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("domain", c_i * 3), ("particle_radius", c_f), ("particle_diameter", c_f),
                ("kernel_radius_scale", c_f), ("capacity_sand", c_i), ("capacity_solid", c_i),
                ("max_neighbors", c_i), ("device", c_i), ("slab_x_lo", c_i), ("slab_x_hi", c_i),
                ("stream", C.c_void_p), ("halo_capacity", c_i), ("slab_ghost_columns", c_i), ("slab_guard_columns", c_i)]


class StepParams(C.Structure):
    _fields_ = [("dt", c_f), ("gravity", c_f * 3),
                ("rest_density", c_f), ("mass", c_f), ("relaxation_epsilon", c_f),
                ("s_corr_dq", c_f), ("s_corr_k", c_f), ("s_corr_n", c_f),
                ("iterations", c_i), ("literal_lambda_index", c_i), ("exact_math", c_i), ("sph_kernel", c_i),
                ("player_position", c_f * 3),
                ("attract_flag", c_i), ("blow_flag", c_i), ("prev_attract_flag", c_i),
                ("attract_radius", c_f), ("blow_radius", c_f), ("attract_coeff", c_f), ("blow_coeff", c_f),
                ("collision_coeff", c_f), ("friction_coeff", c_f), ("mu_s", c_f), ("mu_k", c_f),
                ("credits", c_i)]


class GridInfo(C.Structure):
    _fields_ = [("grid", c_i * 3), ("num_cells", c_i), ("cell_size", c_f), ("kernel_radius", c_f),
                ("cubic_k", c_f), ("cubic_l", c_f)]


DUMP_KEYS, DUMP_PERM, DUMP_ORIG, DUMP_NBR_COUNT, DUMP_NBR, DUMP_DENSITY, DUMP_LAMBDA, DUMP_PSTAR, \
    DUMP_CELL_START, DUMP_COUNTERS = range(10)

# every symbol include/lgpu.h declares
SYMBOLS = [
    "lgpu_default_step_params", "lgpu_create", "lgpu_destroy", "lgpu_last_error", "lgpu_get_grid",
    "lgpu_upload_sand", "lgpu_upload_solids", "lgpu_append_sand", "lgpu_download_sand", "lgpu_download_sand2",
    "lgpu_host_register", "lgpu_host_unregister", "lgpu_host_is_pinned", "lgpu_num_sand",
    "lgpu_num_solids", "lgpu_step_fluid", "lgpu_step_sand", "lgpu_sync", "lgpu_last_step_ms",
    "lgpu_launch_count", "lgpu_set_phase_timing", "lgpu_set_use_graph", "lgpu_graph_stats", "lgpu_set_stage_slots", "lgpu_set_generic_kernels", "lgpu_cell_count",
    "lgpu_remove_in_cells", "lgpu_aabb_first_k", "lgpu_dump", "lgpu_eval_kernel", "lgpu_counting_sort",
    "lgpu_plan_slabs", "lgpu_slab_guard_columns", "lgpu_slab_capacity", "lgpu_slab_export", "lgpu_slab_connect", "lgpu_slab_info", "lgpu_slab_edge", "lgpu_slab_upload", "lgpu_slab_download",
    "lgpu_slab_step_begin", "lgpu_slab_step_end",
]

_lib = None


def lib():
    """Loads liblgpu.so.  Raises if it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LgpuError("liblgpu.so is missing (%s): build it with `python -c 'import __graft_entry__ as g; "
                            "g.build()'` or `make -C lustrine_b200/csrc`; there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.lgpu_last_error.restype = C.c_char_p
        L.lgpu_default_step_params.argtypes = [C.POINTER(StepParams)]
        L.lgpu_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.lgpu_destroy.argtypes = [vp]
        L.lgpu_destroy.restype = None
        L.lgpu_get_grid.argtypes = [vp, C.POINTER(GridInfo)]
        L.lgpu_upload_sand.argtypes = [vp, c_i, vp, vp, vp]
        L.lgpu_append_sand.argtypes = [vp, c_i, vp, vp, vp]
        L.lgpu_upload_solids.argtypes = [vp, c_i, vp]
        L.lgpu_download_sand.argtypes = [vp, vp, vp, vp]
        L.lgpu_download_sand2.argtypes = [vp, vp, vp, vp, vp]
        L.lgpu_host_register.argtypes = [vp, C.c_size_t]
        L.lgpu_host_unregister.argtypes = [vp]
        L.lgpu_host_is_pinned.argtypes = [vp]
        L.lgpu_num_sand.argtypes = [vp]
        L.lgpu_num_solids.argtypes = [vp]
        L.lgpu_step_fluid.argtypes = [vp, C.POINTER(StepParams)]
        L.lgpu_step_sand.argtypes = [vp, C.POINTER(StepParams)]
        L.lgpu_sync.argtypes = [vp]
        L.lgpu_last_step_ms.argtypes = [vp, c_i, C.POINTER(c_f)]
        L.lgpu_launch_count.argtypes = [vp]
        L.lgpu_launch_count.restype = C.c_long
        L.lgpu_set_phase_timing.argtypes = [vp, c_i]
        L.lgpu_set_use_graph.argtypes = [vp, c_i]
        L.lgpu_graph_stats.argtypes = [vp, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.lgpu_set_stage_slots.argtypes = [vp, c_i]
        L.lgpu_set_generic_kernels.argtypes = [vp, c_i]
        L.lgpu_cell_count.argtypes = [vp, C.POINTER(c_i * 3), C.POINTER(c_i * 3), c_i, C.POINTER(c_i)]
        L.lgpu_remove_in_cells.argtypes = [vp, vp, c_i, C.POINTER(c_i)]
        L.lgpu_aabb_first_k.argtypes = [vp, C.POINTER(c_f * 3), C.POINTER(c_f * 3), c_i, vp, C.POINTER(c_i)]
        L.lgpu_dump.argtypes = [vp, c_i, vp, C.c_size_t]
        L.lgpu_eval_kernel.argtypes = [vp, C.POINTER(StepParams), c_i, vp, c_i, vp]
        L.lgpu_counting_sort.argtypes = [vp, c_i, c_i, vp, c_i]
        L.lgpu_plan_slabs.argtypes = [vp, c_i, c_i, c_i, c_i, vp]
        L.lgpu_slab_guard_columns.argtypes = [c_i]
        L.lgpu_slab_capacity.argtypes = [vp, c_i, vp, c_i, c_i, C.c_double]
        L.lgpu_slab_capacity.restype = C.c_longlong
        L.lgpu_slab_export.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.lgpu_slab_connect.argtypes = [vp, c_i, vp, vp]
        L.lgpu_slab_info.argtypes = [vp, C.POINTER(c_i * 8)]
        L.lgpu_slab_edge.argtypes = [vp, C.POINTER(c_i * 2)]
        L.lgpu_slab_upload.argtypes = [vp, c_i, vp, vp, vp, vp]
        L.lgpu_slab_download.argtypes = [vp, vp, vp, vp, vp, C.POINTER(c_i)]
        L.lgpu_slab_step_begin.argtypes = [vp, C.POINTER(StepParams), c_i]
        L.lgpu_slab_step_end.argtypes = [vp]
        _lib = L
    return _lib


def _check(status, what):
    if status != 0:
        raise LgpuError("%s failed with status %d: %s" % (what, status, lib().lgpu_last_error().decode()))


def plan_slabs_c(column_hist, world, min_columns=1, margin=None):
    """lgpu_plan_slabs (the C host's planner): [(x_lo, x_hi)] per rank from the per-column particle histogram."""
    hist = np.ascontiguousarray(column_hist, np.int64)
    bounds = np.zeros(world + 1, np.int32)
    _check(lib().lgpu_plan_slabs(_ptr(hist), len(hist), int(world), int(min_columns), -1 if margin is None else int(margin), _ptr(bounds)), "lgpu_plan_slabs")
    return [(int(bounds[k]), int(bounds[k + 1])) for k in range(world)]


def slab_capacity_c(column_hist, slabs_plan, ghost_columns=1, factor=1.5):
    hist = np.ascontiguousarray(column_hist, np.int64)
    bounds = np.ascontiguousarray([s[0] for s in slabs_plan] + [slabs_plan[-1][1]], np.int32)
    return int(lib().lgpu_slab_capacity(_ptr(hist), len(hist), _ptr(bounds), len(slabs_plan), int(ghost_columns), float(factor)))


def default_step_params(**overrides):
    p = StepParams()
    lib().lgpu_default_step_params(C.byref(p))
    for k, v in overrides.items():
        if k in ("gravity", "player_position"):
            for a in range(3):
                getattr(p, k)[a] = float(v[a])
        else:
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data


def counting_sort(keys, num_cells, device=-1):
    keys = np.ascontiguousarray(keys, np.int32)
    out = np.zeros(keys.shape[0], np.int32)
    _check(lib().lgpu_counting_sort(_ptr(keys), keys.shape[0], int(num_cells), _ptr(out), device), "lgpu_counting_sort")
    return out


class Context:
    """One device context = one Lustrine simulation's particle state on one GPU."""

    def __init__(self, domain, radius=0.5, diameter=1.0, capacity_sand=0, capacity_solid=0,
                 kernel_radius_scale=3.1, max_neighbors=0, device=-1, stream=None, slab=None, halo_capacity=0, ghost_columns=0,
                 guard_columns=0):
        L = lib()
        cfg = Config()
        for a in range(3):
            cfg.domain[a] = int(domain[a])
        cfg.particle_radius = radius
        cfg.particle_diameter = diameter
        cfg.kernel_radius_scale = kernel_radius_scale
        cfg.capacity_sand = int(capacity_sand)
        cfg.capacity_solid = int(capacity_solid)
        cfg.max_neighbors = int(max_neighbors)
        cfg.device = device
        cfg.stream = stream
        if slab is not None:
            cfg.slab_x_lo, cfg.slab_x_hi = int(slab[0]), int(slab[1])
        cfg.halo_capacity = int(halo_capacity)
        cfg.slab_ghost_columns = int(ghost_columns)
        cfg.slab_guard_columns = int(guard_columns)
        self._h = C.c_void_p()
        _check(L.lgpu_create(C.byref(cfg), C.byref(self._h)), "lgpu_create")
        self.L = L
        gi = GridInfo()
        _check(L.lgpu_get_grid(self._h, C.byref(gi)), "lgpu_get_grid")
        self.grid = tuple(gi.grid)
        self.num_cells = gi.num_cells
        self.cell_size, self.kernel_radius, self.cubic_k, self.cubic_l = gi.cell_size, gi.kernel_radius, gi.cubic_k, gi.cubic_l

    def close(self):
        if getattr(self, "_h", None):
            self.L.lgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- state ----
    @property
    def n(self):
        return self.L.lgpu_num_sand(self._h)

    @property
    def n_solid(self):
        return self.L.lgpu_num_solids(self._h)

    def upload_sand(self, pos, vel=None, flags=None, append=False):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32).reshape(-1, 3)
        flags = None if flags is None else np.ascontiguousarray(flags, np.int32)
        f = self.L.lgpu_append_sand if append else self.L.lgpu_upload_sand
        _check(f(self._h, pos.shape[0], _ptr(pos), _ptr(vel), _ptr(flags)), "lgpu_upload_sand")

    def upload_solids(self, pos):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        _check(self.L.lgpu_upload_solids(self._h, pos.shape[0], _ptr(pos)), "lgpu_upload_solids")

    def download(self, pos=True, vel=True, flags=True, out=None):
        n = self.n
        if out is not None:
            p, v, f = out
        else:
            p = np.zeros((n, 3), np.float32) if pos else None
            v = np.zeros((n, 3), np.float32) if vel else None
            f = np.zeros(n, np.int32) if flags else None
        _check(self.L.lgpu_download_sand(self._h, _ptr(p), _ptr(v), _ptr(f)), "lgpu_download_sand")
        return p, v, f

    def download_into(self, pos_ptr, vel_ptr=None, flags_ptr=None):
        """Raw-pointer variant (pinned host buffers owned by the caller)."""
        _check(self.L.lgpu_download_sand(self._h, pos_ptr, vel_ptr, flags_ptr), "lgpu_download_sand")

    def upload_from(self, n, pos_ptr, vel_ptr=None, flags_ptr=None):
        _check(self.L.lgpu_upload_sand(self._h, int(n), pos_ptr, vel_ptr, flags_ptr), "lgpu_upload_sand")

    # ---- stepping ----
    def step_fluid(self, params=None, **kw):
        p = params if params is not None else default_step_params(**kw)
        _check(self.L.lgpu_step_fluid(self._h, C.byref(p)), "lgpu_step_fluid")

    def step_sand(self, params=None, **kw):
        if params is None:
            kw.setdefault("iterations", 4)
            kw.setdefault("dt", 0.016)
        p = params if params is not None else default_step_params(**kw)
        _check(self.L.lgpu_step_sand(self._h, C.byref(p)), "lgpu_step_sand")

    def sync(self):
        _check(self.L.lgpu_sync(self._h), "lgpu_sync")

    def last_step_ms(self, phase=0):
        ms = c_f()
        _check(self.L.lgpu_last_step_ms(self._h, phase, C.byref(ms)), "lgpu_last_step_ms")
        return ms.value

    def launch_count(self):
        return self.L.lgpu_launch_count(self._h)

    def set_phase_timing(self, on):
        _check(self.L.lgpu_set_phase_timing(self._h, int(on)), "lgpu_set_phase_timing")

    def set_stage_slots(self, slots):
        _check(self.L.lgpu_set_stage_slots(self._h, int(slots)), "lgpu_set_stage_slots")

    def set_generic_kernels(self, on):
        _check(self.L.lgpu_set_generic_kernels(self._h, int(on)), "lgpu_set_generic_kernels")

    def set_use_graph(self, on):
        _check(self.L.lgpu_set_use_graph(self._h, int(on)), "lgpu_set_use_graph")

    def graph_stats(self):
        """(steps captured into a CUDA graph, steps replayed from it)."""
        a, b = C.c_long(), C.c_long()
        _check(self.L.lgpu_graph_stats(self._h, C.byref(a), C.byref(b)), "lgpu_graph_stats")
        return a.value, b.value

    # ---- spatial slabs (one context per GPU; see lustrine_b200/slabs.py) ----
    def slab_export(self):
        """(64-byte CUDA IPC handle, device pointer, bytes) of this context's peer-visible arena."""
        handle = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        nbytes = C.c_size_t()
        _check(self.L.lgpu_slab_export(self._h, handle, C.byref(ptr), C.byref(nbytes)), "lgpu_slab_export")
        return bytes(handle), ptr.value, nbytes.value

    def slab_connect(self, side, handle=None, same_process_ptr=None):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle) if handle is not None else None
        _check(self.L.lgpu_slab_connect(self._h, int(side), buf, same_process_ptr), "lgpu_slab_connect")

    def slab_info(self):
        out = (c_i * 8)()
        _check(self.L.lgpu_slab_info(self._h, C.byref(out)), "lgpu_slab_info")
        return dict(zip(("x_lo", "x_hi", "local_grid_x", "local_cells", "owned", "ghosts", "halo_capacity", "x_off"), list(out)))

    def slab_edge(self):
        """(particles of the last substep inside the guard columns of an open side of a cropped plan, slab has an open side)"""
        out = (c_i * 2)()
        _check(self.L.lgpu_slab_edge(self._h, C.byref(out)), "lgpu_slab_edge")
        return int(out[0]), bool(out[1])

    def slab_upload(self, pos, ids, vel=None, flags=None):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        ids = np.ascontiguousarray(ids, np.int32)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32).reshape(-1, 3)
        flags = None if flags is None else np.ascontiguousarray(flags, np.int32)
        _check(self.L.lgpu_slab_upload(self._h, pos.shape[0], _ptr(pos), _ptr(vel), _ptr(flags), _ptr(ids)), "lgpu_slab_upload")

    def slab_download(self, out=None):
        """(pos, vel, flags, ids) of the owned particles.  out = caller-owned arrays of at least
        self.n rows (e.g. views of pinned memory) to avoid the allocation and the pageable copy."""
        n = self.n
        if out is not None:
            pos, vel, flags, ids = out
            if len(pos) < n or len(vel) < n or len(flags) < n or len(ids) < n:
                raise LgpuError("slab_download: output buffers hold fewer than %d particles" % n)
        else:
            pos = np.zeros((max(n, 1), 3), np.float32); vel = np.zeros((max(n, 1), 3), np.float32)
            flags = np.zeros(max(n, 1), np.int32); ids = np.zeros(max(n, 1), np.int32)
        got = c_i()
        _check(self.L.lgpu_slab_download(self._h, _ptr(pos), _ptr(vel), _ptr(flags), _ptr(ids), C.byref(got)), "lgpu_slab_download")
        return pos[:got.value], vel[:got.value], flags[:got.value], ids[:got.value]

    def slab_step_begin(self, params, mode):
        _check(self.L.lgpu_slab_step_begin(self._h, C.byref(params), int(mode)), "lgpu_slab_step_begin")

    def slab_step_end(self):
        _check(self.L.lgpu_slab_step_end(self._h), "lgpu_slab_step_end")

    # ---- grid queries ----
    def cell_count(self, lo, hi, include_solid):
        a = (c_i * 3)(*[int(x) for x in lo])
        b = (c_i * 3)(*[int(x) for x in hi])
        out = c_i()
        _check(self.L.lgpu_cell_count(self._h, C.byref(a), C.byref(b), int(include_solid), C.byref(out)), "lgpu_cell_count")
        return out.value

    def remove_in_cells(self, cell_ids):
        cell_ids = np.ascontiguousarray(cell_ids, np.int32)
        out = c_i()
        _check(self.L.lgpu_remove_in_cells(self._h, _ptr(cell_ids), cell_ids.shape[0], C.byref(out)), "lgpu_remove_in_cells")
        return out.value

    def aabb_first_k(self, center, half, k):
        a = (c_f * 3)(*[float(x) for x in center])
        b = (c_f * 3)(*[float(x) for x in half])
        out = np.zeros((k, 3), np.float32)
        n = c_i()
        _check(self.L.lgpu_aabb_first_k(self._h, C.byref(a), C.byref(b), int(k), _ptr(out), C.byref(n)), "lgpu_aabb_first_k")
        return out[: n.value]

    # ---- stage dumps (tests) ----
    def dump(self, what):
        n = self.n
        if what in (DUMP_KEYS, DUMP_PERM, DUMP_ORIG, DUMP_NBR_COUNT):
            out = np.zeros(n, np.int32)
        elif what in (DUMP_DENSITY, DUMP_LAMBDA):
            out = np.zeros(n, np.float32)
        elif what == DUMP_PSTAR:
            out = np.zeros((n, 3), np.float32)
        elif what == DUMP_CELL_START:
            out = np.zeros(self.num_cells + 1, np.int32)
        elif what == DUMP_COUNTERS:
            out = np.zeros(4, np.uint64)
        elif what == DUMP_NBR:
            cnt = self.dump(DUMP_NBR_COUNT)
            out = np.zeros(int(cnt.sum()), np.int32)
        else:
            raise ValueError(what)
        if out.nbytes:
            _check(self.L.lgpu_dump(self._h, what, _ptr(out), out.nbytes), "lgpu_dump")
        return out

    def neighbors(self):
        cnt = self.dump(DUMP_NBR_COUNT)
        off = np.zeros(cnt.shape[0] + 1, np.int64)
        np.cumsum(cnt, out=off[1:])
        return off, self.dump(DUMP_NBR)

    def eval_kernel(self, which, r, exact=True, **kw):
        r = np.ascontiguousarray(r, np.float32)
        out = np.zeros_like(r)
        n = r.shape[0]
        p = default_step_params(exact_math=int(exact), **kw)
        _check(self.L.lgpu_eval_kernel(self._h, C.byref(p), which, _ptr(r), n, _ptr(out)), "lgpu_eval_kernel")
        return out
