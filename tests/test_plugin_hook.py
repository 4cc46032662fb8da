"""The reference-side binding of INTEGRATION.md §A, exercised: tests/cpp/plugin_hook.cpp is compiled INSIDE the reference's
build (its headers and objects, oracle/_ref) and provides two `Simulate_fun`s (src/Simulation.hpp:22) on the C ABI of
include/lgpu.h.  The unmodified reference then runs its own frame function (Lustrine::simulate, src/Lustrine.cpp:802-836)
with its stock step and with the plugged-in step; the host arrays it exposes to its callers must agree frame by frame."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "tests", "cpp", "_build", "libref_plugin_hook.so")


def run(lib, which, side, steps):
    out = np.zeros((steps, side ** 3, 3), np.float32)
    n = C.c_int(0)
    lib.plugin_run(which, side, steps, out.ctypes.data_as(C.c_void_p), C.byref(n))
    assert n.value == side ** 3
    return out


def test_plugin_library_is_built_where_the_reference_is():
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("build container only")
    assert os.path.exists(PLUGIN), "tests/cpp/Makefile builds libref_plugin_hook.so where /root/reference exists"


@pytest.mark.gpu
@pytest.mark.parametrize("kind,side,steps", [("sand", 10, 40), ("fluid", 12, 25)])
def test_reference_with_the_step_plugged_in_matches_the_stock_reference(kind, side, steps):
    if not O.have_ref() or not os.path.exists(PLUGIN):
        pytest.skip("oracle/_ref or tests/cpp/_build/libref_plugin_hook.so not built")
    lib = C.CDLL(PLUGIN)
    lib.plugin_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    stock = run(lib, 0 if kind == "sand" else 2, side, steps)
    plugged = run(lib, 1 if kind == "sand" else 3, side, steps)
    err = np.abs(plugged.astype(np.float64) - stock)
    bound = 1e-5 + 1e-5 * np.abs(stock.astype(np.float64))
    moved = float(np.abs(stock[-1] - stock[0]).max())
    print("  %s through Simulation::simulate_fun: max|dx| over %d frames %.3e (worst err/bound %.3f), particles moved up to %.2f"
          % (kind, steps, err.max(), float((err / bound).max()), moved))
    assert moved > 0.5, "the scene must move"
    assert np.all(err <= bound)
