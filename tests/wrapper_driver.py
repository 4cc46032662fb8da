"""One ctypes driver for the LustrineWrapper C API (reference src/LustrineWrapper.hpp:30-188).
The SAME calls are issued against the reference's own library (oracle/_ref/libref_lit.so, which
exports these symbols from the unmodified src/LustrineWrapper.cpp) and against
lustrine_b200/lib/liblustrine_b200.so — the drop-in test compares what the two return."""
import ctypes as C

import numpy as np


class SimulationParameters(C.Structure):
    _fields_ = [("X", C.c_int), ("Y", C.c_int), ("Z", C.c_int), ("particleRadius", C.c_float), ("particleDiameter", C.c_float)]


class SimulationData(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("num_sand_particles", "num_solid_particles", "start_sand_index", "end_sand_index",
                                       "start_solid_index", "end_solid_index")]


class Color(C.Structure):
    _fields_ = [(n, C.c_float) for n in "rgba"]


class Vec3(C.Structure):
    _fields_ = [(n, C.c_float) for n in "xyz"]


class GridWrapper(C.Structure):
    _fields_ = [("cells", C.POINTER(C.c_int)), ("colors", C.POINTER(Color)), ("color", Color), ("position", Vec3),
                ("has_one_color_per_cell", C.c_bool), ("X", C.c_int), ("Y", C.c_int), ("Z", C.c_int),
                ("num_grid_cells", C.c_int), ("num_occupied_grid_cells", C.c_int), ("type", C.c_int)]


WRAPPER_SYMBOLS = """init_simulation init_simulation_extra_parameters simulate simulate_no_flags simulation_bind_positions_copy
cleanup_simulation init_grid_box read_vox_scene free_string create_grid init_grid_magikavoxel get_num_sand_particles
get_grid_cell_size get_gravity set_gravity add_box add_capsule add_detector_block check_collision do_collide
do_collide_except_for check_collisions get_num_bodies apply_impulse get_position get_velocity set_velocity set_position
add_velocity set_body_no_rotation set_body_frixion get_body_frixion set_body_damping get_body_damping set_player_id
set_player_box_scale is_grounded set_attract_blow_parameters add_particle_source add_particle_sink set_source_state
set_sink_state get_source_spawned get_sink_despawned set_simulate_function set_body_gravity set_body_no_collision_response
collide_with_player enable_particles_bounding_boxes disable_particles_bounding_boxes
set_player_particles_bounding_spheres_radius_placement query_cell_num_particles test_allocate_1gb test_deallocate_1gb
get_num_observation get_cycles get_duration""".split()


class Wrapper:
    def __init__(self, path):
        L = C.CDLL(path)
        self.L = L
        L.init_grid_box.argtypes = [C.POINTER(SimulationParameters), C.POINTER(GridWrapper), C.c_int, C.c_int, C.c_int, Vec3, Color, C.c_int]
        L.init_simulation.argtypes = [C.POINTER(SimulationParameters), C.POINTER(SimulationData), C.POINTER(GridWrapper), C.c_int,
                                      C.POINTER(GridWrapper), C.c_int, C.c_int]
        L.init_simulation_extra_parameters.argtypes = L.init_simulation.argtypes + [C.c_float, C.c_int]
        L.simulate.argtypes = [C.c_float, C.c_bool, C.c_bool]
        L.simulate_no_flags.argtypes = [C.c_float]
        L.simulation_bind_positions_copy.argtypes = [C.c_void_p]
        L.get_num_sand_particles.restype = C.c_int
        L.set_attract_blow_parameters.argtypes = [C.c_float] * 4
        L.add_particle_source.argtypes = [C.POINTER(GridWrapper), Vec3, C.c_float, C.c_int]
        L.add_particle_sink.argtypes = [Vec3, Vec3, C.c_float]
        L.query_cell_num_particles.argtypes = [Vec3, Vec3, C.c_bool]
        L.set_simulate_function.argtypes = [C.c_int]
        L.get_source_spawned.argtypes = [C.c_int]
        L.set_player_id.argtypes = [C.c_int]
        L.get_position.argtypes = [C.c_int]
        L.get_position.restype = Vec3
        L.get_duration.argtypes = [C.c_int]
        L.get_duration.restype = C.c_double
        # rigid bodies (host side; src/LustrineWrapper.hpp:108-160)
        L.add_box.argtypes = [Vec3, C.c_bool, Vec3]
        L.add_capsule.argtypes = [Vec3, C.c_float, C.c_float]
        L.add_detector_block.argtypes = [Vec3, Vec3]
        L.set_player_box_scale.argtypes = [Vec3]
        L.set_velocity.argtypes = [C.c_int, Vec3]
        L.add_velocity.argtypes = [C.c_int, Vec3]
        L.apply_impulse.argtypes = [C.c_int, Vec3, Vec3]
        L.set_body_gravity.argtypes = [C.c_int, Vec3]
        L.set_body_frixion.argtypes = [C.c_int, C.c_float]
        if hasattr(L, "get_body_frixion"):  # (declared by the reference's header, defined only by the drop-in)
            L.get_body_frixion.argtypes = [C.c_int]
            L.get_body_frixion.restype = C.c_float
        L.set_body_damping.argtypes = [C.c_int, C.c_float, C.c_float]
        L.get_body_damping.argtypes = [C.c_int]
        L.get_body_damping.restype = C.c_float
        for f in (L.is_grounded, L.do_collide, L.set_body_no_rotation, L.set_body_no_collision_response, L.collide_with_player):
            f.argtypes = [C.c_int]
        L.check_collision.argtypes = [C.c_int, C.c_int]
        L.do_collide_except_for.argtypes = [C.c_int, C.c_int]
        self.params = None

    def grid_box(self, params, dims, position, gtype, cell_value=1):
        g = GridWrapper()
        self.L.init_grid_box(C.byref(params), C.byref(g), dims[0], dims[1], dims[2], Vec3(*position), Color(1, 1, 1, 1), gtype)
        if cell_value != 1:
            for i in range(g.num_grid_cells):
                g.cells[i] = cell_value
        return g

    def init(self, domain, radius, sand, solids, subdivision=1, extra=None):
        """sand / solids: lists of (dims, position[, cell_value])"""
        p = SimulationParameters(domain[0], domain[1], domain[2], radius, 2.0 * radius)
        self.params = p
        sg = (GridWrapper * max(len(sand), 1))(*[self.grid_box(p, s[0], s[1], 1) for s in sand])
        og = (GridWrapper * max(len(solids), 1))(*[self.grid_box(p, s[0], s[1], 0, s[2] if len(s) > 2 else 1) for s in solids])
        self._keep = (sg, og)
        data = SimulationData()
        if extra is None:
            self.L.init_simulation(C.byref(p), C.byref(data), sg, len(sand), og, len(solids), subdivision)
        else:
            self.L.init_simulation_extra_parameters(C.byref(p), C.byref(data), sg, len(sand), og, len(solids), subdivision, extra[0], extra[1])
        return data

    def positions(self, pinned=False):
        n = self.L.get_num_sand_particles()
        if pinned:  # a page-locked caller buffer (the drop-in fills it with the copy engine)
            import torch
            buf = torch.zeros((max(n, 1), 3), dtype=torch.float32).pin_memory()
            if n:
                self.L.simulation_bind_positions_copy(buf.data_ptr())
            return buf.numpy()[:n].copy()
        out = np.zeros((n, 3), np.float32)
        if n:
            self.L.simulation_bind_positions_copy(out.ctypes.data)
        return out

    def add_source(self, dims, position, direction, freq, capacity):
        g = self.grid_box(self.params, dims, position, 1)
        return self.L.add_particle_source(C.byref(g), Vec3(*direction), freq, capacity)

    def add_sink(self, lo, hi, freq):
        return self.L.add_particle_sink(Vec3(*lo), Vec3(*hi), freq)

    def query(self, lo, hi, include_solid):
        return self.L.query_cell_num_particles(Vec3(*lo), Vec3(*hi), include_solid)
