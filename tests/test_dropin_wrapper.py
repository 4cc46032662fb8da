"""Drop-in test of the kept API: the same LustrineWrapper C calls are driven against the reference's
own compiled library and against liblustrine_b200.so; particle counts, positions and grid queries must
agree (sand steps in parity arithmetic are bit-identical)."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_py as O
from wrapper_driver import WRAPPER_SYMBOLS, Wrapper

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "lustrine_b200", "lib", "liblustrine_b200.so")
# the drop-in built against the reference's BulletPhysics.hpp and linked with the reference's own BulletPhysics.cpp + Bullet 3.21
# (tests/cpp/Makefile; test infrastructure, only where /root/reference was available at build time)
OURS_REAL_BULLET = os.path.join(ROOT, "tests", "cpp", "_build", "liblustrine_b200_realbullet.so")


def test_wrapper_symbols_exported():
    out = subprocess.check_output(["nm", "-D", "--defined-only", OURS]).decode()
    exported = set(re.findall(r" T (\w+)", out))
    missing = [s for s in WRAPPER_SYMBOLS if s not in exported]
    assert not missing, "liblustrine_b200.so lacks C entry points of the reference wrapper: %s" % missing
    assert len(WRAPPER_SYMBOLS) >= 56  # the wrapper entry points + 3 profiling getters
    hdr = "/root/reference/src/LustrineWrapper.hpp"
    if os.path.exists(hdr):  # only in the build container
        ref = set(re.findall(r'^\s*extern "C" LUSTRINE_WRAPPER_EXPORT [\w:]+\s+(\w+)\(', open(hdr).read(), re.M))
        assert ref <= exported, sorted(ref - exported)
    for cxx in ("init_simulation", "init_simulation_extra_parameters", "clean_simulation", "init_chunk_from_grid", "init_grid_box",
                "init_grid_box_random", "init_grid_from_magika_voxel", "add_particle_source", "add_particle_sink",
                "query_cell_num_particles", "simulate", "simulate_fluid", "simulate_sand", "simulate_sand_credits", "simulate_sand_v3"):
        assert re.search(r"_ZN8Lustrine\d+%s" % cxx, out), "C++ API function %s not exported" % cxx


def run_scenario(w, steps, flags=None, with_source_sink=False, host_sync=0, pinned_out=False):
    data = w.init((30, 30, 30), 0.5, sand=[((10, 10, 10), (5.0, 8.0, 5.0))],
                  solids=[((24, 1, 24), (0.0, 0.0, 0.0), 2), ((4, 3, 4), (8.0, 1.0, 8.0), 2)], subdivision=1)
    if host_sync:
        # 1 = SYNC_RESIDENT (device state authoritative, host arrays refreshed after every step), 2 = SYNC_LAZY
        # (nothing downloaded until simulation_bind_positions_copy) — lustrine_b200 only, lustrine/Lustrine.hpp
        w.L.b200_set_host_sync(host_sync)
    trace = {"n0": (data.num_sand_particles, data.num_solid_particles, data.start_solid_index, data.end_solid_index), "pos": [], "n": [], "q": []}
    if with_source_sink:
        w.add_source((2, 2, 2), (14.0, 22.0, 14.0), (0.0, -1.0, 0.0), 0.05, 48)
        w.add_sink((0.0, 0.0, 0.0), (30.0, 2.5, 12.0), 0.0)
    w.L.set_attract_blow_parameters(12.0, 9.0, 1000.0, 500.0)
    for s in range(steps):
        att, blow = flags[s] if flags else (False, False)
        w.L.simulate(0.016, att, blow)
        trace["pos"].append(w.positions(pinned_out))
        trace["n"].append(w.L.get_num_sand_particles())
        trace["q"].append((w.query((4.0, 0.0, 4.0), (16.0, 12.0, 16.0), False), w.query((0.0, 0.0, 0.0), (30.0, 4.0, 30.0), True)))
    w.L.cleanup_simulation()
    return trace


@pytest.mark.gpu
@pytest.mark.parametrize("scenario,host_sync", [("sand", 0), ("attract_blow", 0), ("source_sink", 0),
                                                ("attract_blow", 1), ("source_sink", 1), ("attract_blow", 2), ("source_sink", 2)])
def test_wrapper_matches_reference(scenario, host_sync):
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    steps = 24
    flags = None
    if scenario == "attract_blow":
        flags = [(False, False)] * 4 + [(True, False)] * 6 + [(False, False)] * 4 + [(False, True)] * 4 + [(False, False)] * 6
    kw = dict(steps=steps, flags=flags, with_source_sink=scenario == "source_sink")
    ref = run_scenario(Wrapper(O.REF_LIT_SO), **kw)
    got = run_scenario(Wrapper(OURS), host_sync=host_sync, **kw)
    assert got["n0"] == ref["n0"]
    assert got["n"] == ref["n"], "particle counts per step"
    assert got["q"] == ref["q"], "query_cell_num_particles per step"
    worst = 0.0
    for a, b in zip(got["pos"], ref["pos"]):
        assert a.shape == b.shape
        worst = max(worst, float(np.abs(a - b).max()) if a.size else 0.0)
    print("  %s: max|dx| over %d steps = %.3e" % (scenario, steps, worst))
    assert worst <= 1e-5
    if scenario == "source_sink":
        assert ref["n"][-1] != ref["n0"][0], "sources/sinks must have changed the particle count"


@pytest.mark.gpu
@pytest.mark.parametrize("host_sync", [0, 1, 2])
def test_device_capacity_grows_with_the_sources(host_sync, monkeypatch):
    """The drop-in sizes its device context for the live particles (not for the X*Y*Z slots of the reference's host
    arrays) and re-creates it larger when the particle sources outgrow it: with the capacity pinned to the initial
    count, the first spawn forces a growth — results must not change."""
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    kw = dict(steps=24, flags=None, with_source_sink=True)
    ref = run_scenario(Wrapper(O.REF_LIT_SO), **kw)
    monkeypatch.setenv("LUSTRINE_B200_MAX_SAND", str(ref["n0"][0]))
    got = run_scenario(Wrapper(OURS), host_sync=host_sync, **kw)
    assert got["n"] == ref["n"] and max(ref["n"]) > ref["n0"][0], "the sources must have outgrown the initial capacity"
    assert got["q"] == ref["q"]
    worst = max(float(np.abs(a - b).max()) for a, b in zip(got["pos"], ref["pos"]))
    print("  growth under host_sync %d: max|dx| %.3e" % (host_sync, worst))
    assert worst <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("pin_host,pinned_out", [("0", False), ("1", True), ("0", True)])
def test_host_arrays_pageable_or_page_locked(pin_host, pinned_out, monkeypatch):
    """The drop-in page-locks the live part of the Simulation's host arrays (positions, positions_star, velocities,
    attracted) and lets the copy engine fill positions_star and a page-locked simulation_bind_positions_copy
    destination; LUSTRINE_B200_PIN_HOST=0 keeps everything pageable.  Same frames either way, sources growing the
    registered range included."""
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    kw = dict(steps=24, flags=None, with_source_sink=True)
    ref = run_scenario(Wrapper(O.REF_LIT_SO), **kw)
    monkeypatch.setenv("LUSTRINE_B200_PIN_HOST", pin_host)
    monkeypatch.setenv("LUSTRINE_B200_MAX_SAND", str(ref["n0"][0]))  # the first spawn re-creates the context and re-registers
    got = run_scenario(Wrapper(OURS), pinned_out=pinned_out, **kw)
    assert got["n"] == ref["n"] and got["q"] == ref["q"]
    worst = max(float(np.abs(a - b).max()) for a, b in zip(got["pos"], ref["pos"]))
    assert worst <= 1e-5


def run_rigid_body_scenario(w, steps):
    """The mixed scene of experiments/bullet/bullet.cpp:71-121 through the C wrapper: sand on a voxel floor, a capsule
    player that drops into it, dynamic boxes, a static ground box, a detector block; the <= 100 proxy boxes around the
    player switched on; attract and blow windows.  Returns what a game would read back every frame."""
    from wrapper_driver import Vec3
    L = w.L
    data = w.init((30, 25, 30), 0.5, sand=[((9, 7, 9), (10.0, 2.0, 10.0))], solids=[((24, 1, 24), (0.0, 0.0, 0.0), 2)], subdivision=1)
    player = L.add_capsule(Vec3(15.0, 11.0, 15.0), 2.0, 3.0)
    boxes = [L.add_box(Vec3(16.0, 15.0, 16.0), True, Vec3(0.5, 0.5, 0.5)), L.add_box(Vec3(14.0, 15.0, 14.0), True, Vec3(0.5, 0.5, 0.5)),
             L.add_box(Vec3(6.0, 2.0, 6.0), True, Vec3(3.0, 1.0, 1.0))]
    ground = L.add_box(Vec3(15.0, -0.5, 15.0), False, Vec3(15.0, 1.0, 15.0))
    detector = L.add_detector_block(Vec3(15.0, 6.0, 15.0), Vec3(2.0, 2.0, 2.0))
    L.set_body_no_rotation(player)
    L.set_body_frixion(player, 0.4)
    L.set_body_damping(boxes[2], 0.1, 0.2)
    L.set_player_id(player)
    L.set_player_box_scale(Vec3(4.0, 7.0, 4.0))
    L.enable_particles_bounding_boxes()
    L.set_attract_blow_parameters(12.0, 9.0, 1000.0, 500.0)
    trace = {"n0": (data.num_sand_particles, data.num_solid_particles), "bodies": L.get_num_bodies(), "pos": [], "body_pos": [], "flags": [],
             "props": (L.get_body_damping(boxes[2]), L.get_body_damping(player))}
    for s in range(steps):
        if s == 30:
            L.apply_impulse(boxes[2], Vec3(0.0, 12.0, 3.0), Vec3(0.0, 0.0, 0.0))
        L.simulate(0.016, 18 <= s < 30, 38 <= s < 44)
        trace["pos"].append(w.positions())
        bp = []
        for b in [player] + boxes + [ground, detector]:
            v = L.get_position(b)
            bp.append((v.x, v.y, v.z))
        trace["body_pos"].append(bp)
        trace["flags"].append((L.is_grounded(player), L.check_collision(detector, player), L.do_collide(player), L.collide_with_player(boxes[2]),
                               L.do_collide_except_for(player, ground)))
    L.cleanup_simulation()
    return trace


@pytest.mark.gpu
def test_mixed_scene_with_the_real_bullet_world_matches_the_reference():
    """BASELINE config 5 in small: the drop-in with a REAL Bullet world on the host (the reference's own BulletPhysics.cpp
    and Bullet 3.21 linked in, north_star keeps them there) against the unmodified reference, frame by frame: particles,
    every rigid body's position, the ground / detector / collision answers.  The particle step runs on the GPU; the proxy
    boxes the reference parks on the <= 100 particles around the player come from the host arrays the step hands back."""
    if not O.have_ref() or not os.path.exists(OURS_REAL_BULLET):
        pytest.skip("oracle/_ref or tests/cpp/_build/liblustrine_b200_realbullet.so not built")
    steps = 70
    ref = run_rigid_body_scenario(Wrapper(O.REF_LIT_SO), steps)
    got = run_rigid_body_scenario(Wrapper(OURS_REAL_BULLET), steps)
    assert got["n0"] == ref["n0"] and got["bodies"] == ref["bodies"] and got["props"] == ref["props"]
    worst_p = max(float(np.abs(a - b).max()) for a, b in zip(got["pos"], ref["pos"]))
    worst_b = max(float(np.abs(np.array(a) - np.array(b)).max()) for a, b in zip(got["body_pos"], ref["body_pos"]))
    moved = float(np.abs(np.array(ref["body_pos"][-1]) - np.array(ref["body_pos"][0])).max())
    grounded = [f[0] for f in ref["flags"]]
    print("  particles max|dx| %.3e, rigid bodies max|dx| %.3e over %d frames (bodies moved up to %.2f; player grounded in %d frames, "
          "in the detector in %d)" % (worst_p, worst_b, steps, moved, sum(grounded), sum(f[1] for f in ref["flags"])))
    assert worst_p <= 1e-5 and worst_b <= 1e-4
    assert got["flags"] == ref["flags"], "is_grounded / check_collision / do_collide per frame"
    assert moved > 3.0 and 0 < sum(grounded) < steps, "the scene must exercise the rigid bodies"


@pytest.mark.gpu
def test_stand_in_bodies_keep_the_wrapper_api_alive():
    """The default build answers the ~30 rigid-body pass-throughs from host/HostBodies.cpp (a stand-in): the
    particle side of the mixed scene must not depend on which engine moves the player — with the player pinned by
    set_position every frame, the particles are those of the reference."""
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    from wrapper_driver import Vec3

    def run(w):
        L = w.L
        w.init((30, 25, 30), 0.5, sand=[((9, 7, 9), (10.0, 2.0, 10.0))], solids=[((24, 1, 24), (0.0, 0.0, 0.0), 2)], subdivision=1)
        player = L.add_capsule(Vec3(15.0, 6.0, 15.0), 2.0, 3.0)
        L.set_body_gravity(player, Vec3(0.0, 0.0, 0.0))
        L.set_player_id(player)
        L.set_player_box_scale(Vec3(4.0, 7.0, 4.0))
        L.set_attract_blow_parameters(12.0, 9.0, 1000.0, 500.0)
        out = []
        for s in range(30):
            L.set_position(player, Vec3(15.0 + 0.1 * s, 6.0, 15.0))
            L.set_velocity(player, Vec3(0.0, 0.0, 0.0))
            L.simulate(0.016, 8 <= s < 16, 20 <= s < 24)
            out.append(w.positions())
        L.cleanup_simulation()
        return out
    ref, got = run(Wrapper(O.REF_LIT_SO)), run(Wrapper(OURS))
    worst = max(float(np.abs(a - b).max()) for a, b in zip(got, ref))
    print("  particles max|dx| %.3e over 30 frames with the player moved by the caller" % worst)
    assert worst <= 1e-5
