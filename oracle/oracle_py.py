"""oracle/oracle_py.py — TEST INFRASTRUCTURE ONLY.

ctypes bindings for (a) the plain-C port ``liblustrine_oracle.so`` and (b) the compiled
unmodified reference ``_ref/libref_{lit,time}.so`` (see oracle/Makefile).  Only tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
import this module; nothing under ``lustrine_b200/`` does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liblustrine_oracle.so")
REF_LIT_SO = os.path.join(HERE, "_ref", "libref_lit.so")
REF_TIME_SO = os.path.join(HERE, "_ref", "libref_time.so")

c_f = C.c_float
c_i = C.c_int
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)
lp = C.POINTER(C.c_long)


def build(ref=True):
    """Compile the port (always) and the reference libraries (when /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])


def have_ref(which="lit"):
    return os.path.exists(REF_LIT_SO if which == "lit" else REF_TIME_SO)


class LoSim(C.Structure):
    _fields_ = [
        ("domainX", c_f), ("domainY", c_f), ("domainZ", c_f),
        ("particleRadius", c_f), ("particleDiameter", c_f),
        ("kernelRadius", c_f), ("kernelFactor", c_f), ("cell_size", c_f),
        ("cubic_kernel_k", c_f), ("cubic_kernel_l", c_f),
        ("gridX", c_i), ("gridY", c_i), ("gridZ", c_i), ("num_grid_cells", c_i),
        ("rest_density", c_f), ("mass", c_f), ("relaxation_epsilon", c_f),
        ("s_corr_dq", c_f), ("s_corr_k", c_f), ("s_corr_n", c_f),
        ("gravity", c_f * 3), ("time_step", c_f),
        ("attract_radius", c_f), ("blow_radius", c_f), ("attract_coeff", c_f), ("blow_coeff", c_f),
        ("player_position", c_f * 3),
        ("attract_flag", c_i), ("blow_flag", c_i), ("prev_attract_flag", c_i),
        ("n_sand", c_i), ("n_solid", c_i), ("capacity", c_i),
        ("positions", fp), ("positions_star", fp), ("positions_tmp", fp), ("velocities", fp),
        ("attracted", ip), ("lambdas", fp), ("densities", fp),
        ("nbr_offsets", lp), ("nbr", ip), ("nbr_capacity", C.c_long),
        ("keys", ip), ("sorted_index", ip),
        ("cell_counts", ip), ("cell_start", ip), ("cell_items", ip),
        ("solid_cell_start", ip), ("solid_cell_items", ip), ("solid_grid_built", c_i),
        ("scratch3", fp), ("scratchi", ip), ("violations", C.c_long),
        ("sph_kernel", c_i),
    ]


_port = None


def port_lib():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build(ref=False)
        L = C.CDLL(PORT_SO)
        L.lo_create.restype = C.POINTER(LoSim)
        L.lo_create.argtypes = [c_i, c_i, c_i, c_f, c_f, c_f, c_i, c_i]
        P = C.POINTER(LoSim)
        L.lo_destroy.argtypes = [P]
        L.lo_set_sand.argtypes = [P, c_i, C.c_void_p, C.c_void_p, C.c_void_p]
        L.lo_set_solid.argtypes = [P, C.c_void_p]
        L.lo_cell_id.argtypes = [P, c_f, c_f, c_f]
        L.lo_cell_id.restype = c_i
        L.lo_counting_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_void_p]
        L.lo_cubic_kernel.argtypes = [P, c_f]
        L.lo_cubic_kernel.restype = c_f
        L.lo_poly6_kernel.argtypes = [P, c_f]
        L.lo_poly6_kernel.restype = c_f
        L.lo_s_coor.argtypes = [P, c_f]
        L.lo_s_coor.restype = c_f
        L.lo_cubic_kernel_grad.argtypes = [P, C.c_void_p, C.c_void_p]
        L.lo_spiky_kernel.argtypes = [P, C.c_void_p, C.c_void_p]
        for name in ("lo_find_neighbors_v0", "lo_find_neighbors_v1", "lo_fluid_lambda", "lo_fluid_commit", "lo_sand_commit"):
            getattr(L, name).argtypes = [P]
        L.lo_fluid_predict.argtypes = [P, c_f]
        L.lo_fluid_deltap.argtypes = [P, c_i, c_i]
        L.lo_step_fluid.argtypes = [P, c_f, c_i, c_i, c_i]
        L.lo_sand_predict.argtypes = [P, c_f, c_i]
        L.lo_sand_iteration.argtypes = [P, c_i]
        L.lo_step_sand.argtypes = [P, c_f, c_i, c_i]
        _port = L
    return _port


def _arr(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).reshape(shape)


class PortSim:
    """The plain-C port.  Sand indices are [0, n_sand); solids are reported as n_sand + k."""

    def __init__(self, X, Y, Z, radius=0.5, diameter=1.0, kernel_radius_scale=3.1, capacity=0, n_solid=0):
        self.L = port_lib()
        self.p = self.L.lo_create(X, Y, Z, radius, diameter, kernel_radius_scale, max(capacity, 1), n_solid)
        self.s = self.p.contents

    def close(self):
        if self.p is not None:
            self.L.lo_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_sand(self, pos, vel=None, attracted=None):
        pos = np.ascontiguousarray(pos, np.float32)
        assert pos.shape[0] <= self.s.capacity
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32)
        attracted = None if attracted is None else np.ascontiguousarray(attracted, np.int32)
        self.L.lo_set_sand(self.p, pos.shape[0], pos.ctypes.data,
                           None if vel is None else vel.ctypes.data,
                           None if attracted is None else attracted.ctypes.data)

    def set_solid(self, pos):
        pos = np.ascontiguousarray(pos, np.float32)
        assert pos.shape[0] == self.s.n_solid
        self.L.lo_set_solid(self.p, pos.ctypes.data)

    n = property(lambda self: self.s.n_sand)
    positions = property(lambda self: _arr(self.s.positions, (self.n, 3), np.float32))
    positions_star = property(lambda self: _arr(self.s.positions_star, (self.n, 3), np.float32))
    velocities = property(lambda self: _arr(self.s.velocities, (self.n, 3), np.float32))
    attracted = property(lambda self: _arr(self.s.attracted, (self.n,), np.int32))
    lambdas = property(lambda self: _arr(self.s.lambdas, (self.n,), np.float32))
    densities = property(lambda self: _arr(self.s.densities, (self.n,), np.float32))
    keys = property(lambda self: _arr(self.s.keys, (self.n,), np.int32))
    sorted_index = property(lambda self: _arr(self.s.sorted_index, (self.n,), np.int32))

    def neighbors(self):
        off = _arr(self.s.nbr_offsets, (self.n + 1,), np.int64).copy()
        flat = _arr(self.s.nbr, (int(off[-1]),), np.int32).copy()
        return off, flat

    def cell_ids(self, pos):
        pos = np.ascontiguousarray(pos, np.float32)
        return np.array([self.L.lo_cell_id(self.p, float(a), float(b), float(c)) for a, b, c in pos], np.int32)


_ref = {}


def ref_lib(which="lit"):
    if which not in _ref:
        path = REF_LIT_SO if which == "lit" else REF_TIME_SO
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build it here with `make -C oracle ref`; needs /root/reference)")
        L = C.CDLL(path)
        vp = C.c_void_p
        L.refh_create.restype = vp
        L.refh_create.argtypes = [c_i, c_i, c_i, c_f, c_f, c_i, c_i, c_i, c_i, c_f, c_i]
        L.refh_destroy.argtypes = [vp]
        L.refh_info.argtypes = [vp, vp, vp]
        L.refh_set_sand.argtypes = [vp, vp, vp, vp]
        L.refh_set_solid.argtypes = [vp, vp]
        L.refh_get_sand.argtypes = [vp, vp, vp, vp, vp]
        L.refh_get_solid.argtypes = [vp, vp]
        L.refh_set_fun.argtypes = [vp, c_i, c_i, c_i]
        L.refh_set_scalars.argtypes = [vp, vp, c_f, c_f, c_f, c_f, c_f, c_f]
        L.refh_set_kernel.argtypes = [vp, c_i]
        L.refh_set_player.argtypes = [vp, vp, c_i, c_i, c_f, c_f, c_f, c_f]
        L.refh_step.restype = C.c_double
        L.refh_step.argtypes = [vp, c_f, c_i, c_i]
        L.refh_find_neighbors.argtypes = [vp, c_i]
        L.refh_get_lambdas.argtypes = [vp, vp]
        L.refh_neighbor_counts.restype = C.c_long
        L.refh_neighbor_counts.argtypes = [vp, vp]
        L.refh_neighbors.argtypes = [vp, vp]
        L.refh_sorted_cell_ids.argtypes = [vp, vp]
        L.refh_cell_ids.argtypes = [vp, vp, c_i, vp]
        L.refh_counting_sort.argtypes = [vp, vp, C.c_long, C.c_long, vp]
        L.refh_scalar_kernel.argtypes = [vp, c_i, vp, c_i, vp]
        L.refh_vector_kernel.argtypes = [vp, c_i, vp, c_i, vp]
        L.refh_query_cell_num_particles.argtypes = [vp, vp, vp, c_i]
        L.refh_add_sink.argtypes = [vp, vp, vp, c_f]
        L.refh_add_source_box.argtypes = [vp, c_i, c_i, c_i, vp, vp, c_f, c_i]
        _ref[which] = L
    return _ref[which]


class RefSim:
    """The unmodified reference behind oracle/ref_harness.cpp.  Sand indices are [0, n_sand);
    neighbour indices of solids are rebased to n_sand + k like PortSim."""

    SAND, FLUID, FLUID_JACOBI, SAND_CREDITS = 0, 1, 2, 3

    def __init__(self, X, Y, Z, radius=0.5, diameter=1.0, n_sand=0, n_solid=0, which="lit",
                 use_extra=False, kernel_radius_scale=3.1, with_credits=False, subdivision=1):
        self.L = ref_lib(which)
        # the reference prints banners on std::cout; keep them out of JSON lines
        self.h = self.L.refh_create(X, Y, Z, radius, diameter, n_sand, n_solid, subdivision,
                                    int(use_extra), kernel_radius_scale, int(with_credits))
        self.refresh()

    def refresh(self):
        iv = (c_i * 16)()
        fv = (c_f * 20)()
        self.L.refh_info(self.h, iv, fv)
        (self.n_sand, self.n_solid, self.ptr_sand_start, self.ptr_sand_end, self.ptr_solid_start,
         self.ptr_solid_end, self.total_allocated, self.gridX, self.gridY, self.gridZ, self.num_grid_cells,
         self.num_remaining) = list(iv)[:12]
        f = list(fv)
        self.kernelRadius, self.cell_size, self.cubic_k, self.cubic_l = f[0:4]
        self.player_position = np.array(f[14:17], np.float32)
        self.time_step = f[9]

    def close(self):
        if self.h:
            self.L.refh_destroy(self.h)
            self.h = None

    def set_sand(self, pos=None, vel=None, attracted=None):
        a = [None if x is None else np.ascontiguousarray(x, t) for x, t in
             ((pos, np.float32), (vel, np.float32), (attracted, np.int32))]
        self.L.refh_set_sand(self.h, *[None if x is None else x.ctypes.data for x in a])

    def set_solid(self, pos):
        pos = np.ascontiguousarray(pos, np.float32)
        assert pos.shape[0] == self.n_solid
        self.L.refh_set_solid(self.h, pos.ctypes.data)

    def get_sand(self):
        self.refresh()
        n = self.ptr_sand_end - self.ptr_sand_start
        pos = np.zeros((n, 3), np.float32); star = np.zeros((n, 3), np.float32)
        vel = np.zeros((n, 3), np.float32); att = np.zeros(n, np.int32)
        self.L.refh_get_sand(self.h, pos.ctypes.data, star.ctypes.data, vel.ctypes.data, att.ctypes.data)
        return pos, star, vel, att

    def set_fun(self, which, iterations=1, literal_lambda_index=True):
        self.L.refh_set_fun(self.h, which, iterations, int(literal_lambda_index))

    def set_kernel(self, which):
        """0 = cubic spline (the reference's wiring), 1 = Simulation::W / gradW pointed at poly6_kernel / spiky_kernel."""
        self.L.refh_set_kernel(self.h, int(which))

    def set_player(self, pos=None, attract=False, blow=False, attract_radius=1.5, blow_radius=2.0,
                   attract_coeff=1000.0, blow_coeff=500.0):
        p = None if pos is None else np.ascontiguousarray(pos, np.float32)
        self.L.refh_set_player(self.h, None if p is None else p.ctypes.data, int(attract), int(blow),
                               attract_radius, blow_radius, attract_coeff, blow_coeff)

    def step(self, dt, steps=1, through_simulate=False):
        return self.L.refh_step(self.h, dt, steps, int(through_simulate))

    def find_neighbors(self, which):
        self.L.refh_find_neighbors(self.h, which)

    def lambdas(self):
        self.refresh()
        out = np.zeros(self.ptr_sand_end - self.ptr_sand_start, np.float32)
        self.L.refh_get_lambdas(self.h, out.ctypes.data)
        return out

    def neighbors(self):
        self.refresh()
        n = self.ptr_sand_end - self.ptr_sand_start
        counts = np.zeros(n, np.int32)
        total = self.L.refh_neighbor_counts(self.h, counts.ctypes.data)
        flat = np.zeros(total, np.int32)
        self.L.refh_neighbors(self.h, flat.ctypes.data)
        off = np.zeros(n + 1, np.int64)
        np.cumsum(counts, out=off[1:])
        solid = flat >= self.ptr_solid_start
        flat[solid] = flat[solid] - self.ptr_solid_start + n
        return off, flat

    def sorted_cell_ids(self):
        out = np.zeros(self.n_sand, np.int32)
        self.L.refh_sorted_cell_ids(self.h, out.ctypes.data)
        return out

    def cell_ids(self, pos):
        pos = np.ascontiguousarray(pos, np.float32)
        out = np.zeros(pos.shape[0], np.int32)
        self.L.refh_cell_ids(self.h, pos.ctypes.data, pos.shape[0], out.ctypes.data)
        return out

    def scalar_kernel(self, which, r):
        r = np.ascontiguousarray(r, np.float32)
        out = np.zeros_like(r)
        self.L.refh_scalar_kernel(self.h, which, r.ctypes.data, r.shape[0], out.ctypes.data)
        return out

    def vector_kernel(self, which, r):
        r = np.ascontiguousarray(r, np.float32)
        out = np.zeros_like(r)
        self.L.refh_vector_kernel(self.h, which, r.ctypes.data, r.shape[0], out.ctypes.data)
        return out


def counting_sort_ref(keys, num_cells, which="lit"):
    keys = np.ascontiguousarray(keys, np.int32)
    counts = np.zeros(num_cells + 1, np.int32)
    out = np.zeros(keys.shape[0], np.int32)
    ref_lib(which).refh_counting_sort(counts.ctypes.data, keys.ctypes.data, keys.shape[0], num_cells, out.ctypes.data)
    return out


def counting_sort_port(keys, num_cells):
    keys = np.ascontiguousarray(keys, np.int32)
    counts = np.zeros(num_cells + 1, np.int32)
    out = np.zeros(keys.shape[0], np.int32)
    port_lib().lo_counting_sort(counts.ctypes.data, keys.ctypes.data, keys.shape[0], num_cells, out.ctypes.data)
    return out
