"""CPU tests (-m "not gpu") of the drop-in boundary: the C-ABI library loads, exports every symbol
include/lgpu.h declares, its POD structs match the ctypes mirror, and — there being no CPU
fallback — every compute entry point fails loudly on a box without a CUDA device."""
import os
import re
import subprocess
import tempfile

import pytest

from lustrine_b200 import lgpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lgpu.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"LGPU_API\s+[\w\s\*]+?\b(lgpu_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    L = lgpu.lib()
    decl = declared_symbols()
    assert len(decl) >= 20
    assert sorted(lgpu.SYMBOLS) == decl, "lgpu.SYMBOLS must list exactly what include/lgpu.h declares"
    for s in decl:
        assert hasattr(L, s), "liblgpu.so does not export %s" % s
    out = subprocess.check_output(["nm", "-D", "--defined-only", lgpu.LIB_PATH]).decode()
    exported = set(re.findall(r" T (lgpu_\w+)", out))
    assert set(decl) <= exported
    assert not [s for s in exported if s not in decl], "undeclared symbols leak from liblgpu.so"


def test_struct_layout_matches_c():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "lgpu.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(lgpu_config), sizeof(lgpu_step_params), sizeof(lgpu_grid_info),
         offsetof(lgpu_config, stream), offsetof(lgpu_step_params, player_position), offsetof(lgpu_step_params, credits),
         offsetof(lgpu_step_params, iterations));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    import ctypes as C
    assert sizes == [C.sizeof(lgpu.Config), C.sizeof(lgpu.StepParams), C.sizeof(lgpu.GridInfo), lgpu.Config.stream.offset,
                     lgpu.StepParams.player_position.offset, lgpu.StepParams.credits.offset, lgpu.StepParams.iterations.offset]


def test_default_step_params_match_the_reference_defaults():
    p = lgpu.default_step_params()
    # src/Simulation.hpp:147-171,229-232 and src/Simulate.cpp:159-163
    assert (p.rest_density, p.mass, p.relaxation_epsilon) == (24.0, 5.0, 10.0)
    assert tuple(p.gravity) == (0.0, -10.0, 0.0)
    assert (p.s_corr_dq, p.s_corr_k, p.s_corr_n) == (0.5, 1.0, 4.0)
    assert (p.attract_radius, p.blow_radius, p.attract_coeff, p.blow_coeff) == (1.5, 2.0, 1000.0, 500.0)
    assert abs(p.collision_coeff - 0.8) < 1e-7 and abs(p.mu_s - 0.95) < 1e-7 and abs(p.mu_k - 0.9) < 1e-7
    assert p.literal_lambda_index == 1 and p.exact_math == 1 and p.iterations == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lgpu.LgpuError):
        lgpu.Context((20, 20, 20), capacity_sand=8)
    with pytest.raises(lgpu.LgpuError):
        import numpy as np
        lgpu.counting_sort(np.array([1, 0], np.int32), 2)


def test_product_never_imports_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "lustrine_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"oracle_py|lustrine_oracle|libref_|import oracle", text):
                    bad.append(f)
    assert not bad, "product files reference the oracle: %s" % bad
