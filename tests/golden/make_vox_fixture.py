"""tests/golden/level1_physical_cells.npz: the cell values of the reference's experiments/resources/level1_physical.vox
as the UNMODIFIED reference loader lays them out (src/VoxelLoader.cpp:47-100, through the wrapper's
init_grid_magikavoxel of oracle/_ref/libref_lit.so).  The .vox asset itself stays in /root/reference; the GPU box only
sees this array (3 KB).  Run in the build container:

    python tests/golden/make_vox_fixture.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py as O  # noqa: E402
from wrapper_driver import GridWrapper, Vec3  # noqa: E402

VOX = "/root/reference/experiments/resources/level1_physical.vox"


def load_cells(lib_path, vox=VOX):
    L = C.CDLL(lib_path)
    L.init_grid_magikavoxel.argtypes = [C.POINTER(GridWrapper), C.c_char_p, Vec3]
    g = GridWrapper()
    L.init_grid_magikavoxel(C.byref(g), vox.encode(), Vec3(0.0, 0.0, 0.0))
    n = g.X * g.Y * g.Z
    cells = np.ctypeslib.as_array(g.cells, shape=(n,)).copy()
    return (g.X, g.Y, g.Z), cells, g.num_occupied_grid_cells


if __name__ == "__main__":
    dims, cells, occ = load_cells(O.REF_LIT_SO)
    assert occ == int((cells != 0).sum())
    np.savez_compressed(os.path.join(HERE, "level1_physical_cells.npz"), dims=np.array(dims, np.int32), cells=cells.astype(np.uint8))
    print("level1_physical.vox: %s cells, %d occupied" % (dims, occ))
