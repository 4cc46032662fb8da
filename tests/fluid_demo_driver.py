"""ctypes driver of tests/cpp/fluid_demo.cpp (the reference's fluid experiment, headless) — shared by
tests/test_fluid_demo.py and bench.py's config-1 line.  TEST INFRASTRUCTURE."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO_B200 = os.path.join(ROOT, "tests", "cpp", "_build", "libfluid_demo_b200.so")
DEMO_REF = os.path.join(ROOT, "oracle", "_ref", "libfluid_demo_ref.so")
OURS = os.path.join(ROOT, "lustrine_b200", "lib", "liblustrine_b200.so")
FIXTURE = os.path.join(ROOT, "tests", "golden", "level1_physical_cells.npz")
VOX = "/root/reference/experiments/resources/level1_physical.vox"


class Demo:
    def __init__(self, path, fun, iterations=1, solids=True):
        L = C.CDLL(path)
        L.demo_create.restype = C.c_void_p
        L.demo_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.demo_run.restype = C.c_double
        L.demo_run.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        L.demo_info.argtypes = [C.c_void_p, C.c_void_p]
        L.demo_positions.restype = C.c_int
        L.demo_positions.argtypes = [C.c_void_p, C.c_void_p]
        L.demo_destroy.argtypes = [C.c_void_p]
        self.L = L
        fx = np.load(FIXTURE)
        dims, cells = fx["dims"], np.ascontiguousarray(fx["cells"].astype(np.int32))
        self.h = L.demo_create(cells.ctypes.data if solids else None, int(dims[0]), int(dims[1]), int(dims[2]), fun, iterations)
        info = np.zeros(4, np.int32)
        L.demo_info(self.h, info.ctypes.data)
        self.info = tuple(int(x) for x in info)

    def run(self, calls, dt=0.01):
        counts = np.zeros(calls, np.int32)
        sums = np.zeros(calls, np.float64)
        seconds = self.L.demo_run(self.h, calls, dt, counts.ctypes.data, sums.ctypes.data)
        return counts, sums, seconds

    def positions(self):
        out = np.zeros((self.info[3], 3), np.float32)
        n = self.L.demo_positions(self.h, out.ctypes.data)
        return out[:n].copy()

    def close(self):
        self.L.demo_destroy(self.h)
