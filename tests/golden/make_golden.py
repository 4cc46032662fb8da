"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/libref_lit.so, built from /root/reference by oracle/Makefile with
-O2 -ffp-contract=off).  Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The reference stores no golden vectors of its own for this path, so these files are
"outputs of the reference itself run here" (seeded inputs + stage outputs), kept small.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py as O  # noqa: E402
import scenes  # noqa: E402


def fluid_case(n_side, steps, literal_gs, sph_kernel=0):
    domain, sand = scenes.dam_break(n_side)
    solids = scenes.floor_plate(3 * n_side, 2 * n_side)
    R = O.RefSim(*domain, n_sand=len(sand), n_solid=len(solids))
    R.set_sand(sand); R.set_solid(solids)
    R.set_kernel(sph_kernel)  # 1: Simulation::W / gradW = poly6_kernel / spiky_kernel (src/Kernels.cpp:43-67)
    R.set_fun(R.FLUID if literal_gs else R.FLUID_JACOBI, 1, True)
    out = {"domain": np.array(domain, np.int32), "sand": sand, "solids": solids}
    for s in range(steps):
        R.step(0.01)
        pos, star, vel, att = R.get_sand()
        off, flat = R.neighbors()
        out["pos_%d" % s] = pos; out["vel_%d" % s] = vel; out["lambda_%d" % s] = R.lambdas()
        out["nbr_off_%d" % s] = off.astype(np.int32); out["nbr_%d" % s] = flat
        out["keys_%d" % s] = R.cell_ids(star)
    R.close()
    return out


def sand_case(n_side, steps):
    domain, sand, solids = scenes.sand_pile(n_side, drop=1.0)
    R = O.RefSim(*domain, n_sand=len(sand), n_solid=len(solids))
    R.set_sand(sand); R.set_solid(solids)
    R.set_fun(R.SAND)
    out = {"domain": np.array(domain, np.int32), "sand": sand, "solids": solids}
    for s in range(steps):
        R.step(0.016)
        pos, star, vel, att = R.get_sand()
        off, flat = R.neighbors()
        out["pos_%d" % s] = pos; out["vel_%d" % s] = vel
        out["nbr_off_%d" % s] = off.astype(np.int32); out["nbr_%d" % s] = flat
        out["keys_%d" % s] = R.sorted_cell_ids()
    R.close()
    return out


def kernel_tables():
    R = O.RefSim(60, 40, 40, n_sand=1)
    h = R.kernelRadius
    r = np.linspace(0.0, 2.0 * h, 1024).astype(np.float32)
    rng = np.random.default_rng(5)
    d = rng.normal(size=(1024, 3)).astype(np.float32)
    d *= (r / np.linalg.norm(d, axis=1))[:, None].astype(np.float32)
    d[0] = 0.0
    out = {"r": r, "d": d, "W": R.scalar_kernel(0, r), "poly6": R.scalar_kernel(1, r), "s_coor": R.scalar_kernel(2, r),
           "gradW": R.vector_kernel(0, d), "spiky": R.vector_kernel(1, d),
           "consts": np.array([R.kernelRadius, R.cell_size, R.cubic_k, R.cubic_l], np.float32),
           "grid": np.array([R.gridX, R.gridY, R.gridZ, R.num_grid_cells], np.int32)}
    R.close()
    return out


def counting_sort_case():
    rng = np.random.default_rng(2022)
    keys = rng.integers(0, 1987, 20000).astype(np.int32)  # sizes of experiments/unit_tests/main.cpp:46-89
    return {"keys": keys, "sorted": O.counting_sort_ref(keys, 1987)}


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "fluid_literal_8.npz"), **fluid_case(8, 3, True))
    np.savez_compressed(os.path.join(HERE, "fluid_jacobi_8.npz"), **fluid_case(8, 3, False))
    np.savez_compressed(os.path.join(HERE, "fluid_poly6_jacobi_8.npz"), **fluid_case(8, 3, False, sph_kernel=1))
    np.savez_compressed(os.path.join(HERE, "sand_8.npz"), **sand_case(8, 3))
    np.savez_compressed(os.path.join(HERE, "kernel_tables.npz"), **kernel_tables())
    np.savez_compressed(os.path.join(HERE, "counting_sort.npz"), **counting_sort_case())
    print("golden vectors written to", HERE)
