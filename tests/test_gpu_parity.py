"""GPU parity tests: the CUDA path, called through the C ABI (include/lgpu.h), against the oracle.

Contract: cell keys, sort permutation and neighbour lists (order included) are
bit-exact; densities, lambdas, positions and velocities agree within REL_TOL = 1e-5 relative
(ABS_TOL = 1e-5 absolute floor) per teacher-forced substep.  With exact_math=1 the device
performs the reference's fp32 operations one by one, so most arrays are in fact bit-identical;
the tests print the observed maximum error next to the asserted tolerance.
"""
import numpy as np
import pytest

import oracle_py as O
import scenes
from lustrine_b200 import lgpu

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5
ABS_TOL = 1e-5


def close(a, b, what, rtol=REL_TOL, atol=ABS_TOL):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    err = np.abs(a - b)
    bound = atol + rtol * np.abs(b)
    worst = float((err / bound).max()) if err.size else 0.0
    exact = float((a == b).mean()) if err.size else 1.0
    print("  %-10s max|err| %.3e  worst err/bound %.3f  bit-exact %.4f" % (what, err.max() if err.size else 0.0, worst, exact))
    assert worst <= 1.0, "%s outside tolerance: max|err| %.3e" % (what, err.max())


def make_pair(domain, sand, solids=None, **ctx_kw):
    n_solid = 0 if solids is None else len(solids)
    P = O.PortSim(*domain, capacity=len(sand) + 64, n_solid=n_solid)
    P.set_sand(sand)
    G = lgpu.Context(domain, capacity_sand=len(sand) + 64, capacity_solid=n_solid, **ctx_kw)
    G.upload_sand(sand)
    if n_solid:
        P.set_solid(solids)
        G.upload_solids(solids)
    return P, G


def force_state(P, G):
    """Teacher forcing: the GPU state is reset to the oracle's state before a compared substep."""
    G.upload_sand(P.positions.copy(), P.velocities.copy(), P.attracted.copy())


# ----------------------------------------------------------------------------------------------
def test_counting_sort_known_answer():
    # restated from the reference's only unit test, experiments/unit_tests/main.cpp:46-89
    rng = np.random.default_rng(2022)
    keys = rng.integers(0, 1987, 20000).astype(np.int32)
    got = lgpu.counting_sort(keys, 1987)
    assert np.array_equal(got, np.argsort(keys, kind="stable").astype(np.int32))
    assert np.array_equal(got, O.counting_sort_port(keys, 1987))
    assert np.array_equal(keys[got], np.sort(keys))
    # edge cases: empty, single, all-equal, maximum key
    assert lgpu.counting_sort(np.zeros(0, np.int32), 5).shape == (0,)
    assert np.array_equal(lgpu.counting_sort(np.array([3], np.int32), 5), [0])
    assert np.array_equal(lgpu.counting_sort(np.full(1000, 4, np.int32), 5), np.arange(1000))
    big = rng.integers(0, 3_000_000, 100000).astype(np.int32)
    assert np.array_equal(lgpu.counting_sort(big, 3_000_000), np.argsort(big, kind="stable"))


def test_kernel_tables():
    P = O.PortSim(60, 40, 40, capacity=1)
    with lgpu.Context((60, 40, 40), capacity_sand=1) as G:
        h = G.kernel_radius
        assert (G.grid, G.num_cells) == ((P.s.gridX, P.s.gridY, P.s.gridZ), P.s.num_grid_cells)
        assert (G.kernel_radius, G.cubic_k, G.cubic_l) == (P.s.kernelRadius, P.s.cubic_kernel_k, P.s.cubic_kernel_l)
        r = np.linspace(0.0, 2.0 * h, 4096).astype(np.float32)
        rng = np.random.default_rng(5)
        d = rng.normal(size=(4096, 3)).astype(np.float32)
        d *= (r / np.linalg.norm(d, axis=1))[:, None].astype(np.float32)
        d[0] = 0.0
        L = P.L
        W = np.array([L.lo_cubic_kernel(P.p, float(x)) for x in r], np.float32)
        p6 = np.array([L.lo_poly6_kernel(P.p, float(x)) for x in r], np.float32)
        sc = np.array([L.lo_s_coor(P.p, float(x)) for x in r], np.float32)
        gW = np.zeros_like(d); sp = np.zeros_like(d)
        for i in range(len(d)):
            L.lo_cubic_kernel_grad(P.p, d[i].ctypes.data, gW[i].ctypes.data)
            L.lo_spiky_kernel(P.p, d[i].ctypes.data, sp[i].ctypes.data)

        def ulps(a, b):
            a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
            spacing = np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32))
            return float((np.abs(a.astype(np.float64) - b) / np.maximum(spacing, np.float32(1e-45))).max())

        # On the hot path neighbours are cut at r <= h, i.e. q <= 0.5: that branch of W and
        # the whole of gradW contain no libm call and must be bit-exact.  The outer branch of W
        # (std::pow(1-q, 3.0f)), poly6/spiky (std::pow in double) and s_coor (powf) go through libm,
        # where the device's pow may round differently: <= 4 ulp.
        inner = r <= h
        got = G.eval_kernel(0, r, exact=True)
        print("  W      exact: inner %.1f ulp (tolerance 0), outer %.1f ulp (tolerance 4)" % (ulps(got[inner], W[inner]), ulps(got[~inner], W[~inner])))
        assert np.array_equal(got[inner], W[inner]) and ulps(got[~inner], W[~inner]) <= 4
        for name, which, ref, arg, tol in (("gradW", 1, gW, d, 0), ("poly6", 2, p6, r, 4), ("spiky", 3, sp, d, 4)):
            got = G.eval_kernel(which, arg, exact=True)
            u = ulps(got, ref)
            print("  %-6s exact: %.1f ulp (tolerance %d), bit-exact fraction %.4f" % (name, u, tol, (got == ref).mean()))
            assert u <= tol, name
        # s_coor = -k (W(r)/W(dq))^4: powf on the hot path (<= 4 ulp); beyond r > h the 1-ulp
        # difference of W's outer branch is amplified by the 4th power (<= 16 ulp, off the hot path)
        got = G.eval_kernel(4, r, exact=True)
        print("  s_coor exact: inner %.1f ulp (tolerance 4), outer %.1f ulp (tolerance 16), bit-exact fraction %.4f"
              % (ulps(got[inner], sc[inner]), ulps(got[~inner], sc[~inner]), (got == sc).mean()))
        assert ulps(got[inner], sc[inner]) <= 4 and ulps(got[~inner], sc[~inner]) <= 16
        for name, which, ref, arg in (("W", 0, W, r), ("gradW", 1, gW, d), ("s_coor", 4, sc, r)):
            got = G.eval_kernel(which, arg, exact=False)
            scale = np.abs(ref).max()
            err = np.abs(got.astype(np.float64) - ref).max() / scale
            print("  %-6s fast: max err / max|ref| = %.2e" % (name, err))
            assert err <= 1e-5, name


# ----------------------------------------------------------------------------------------------
def compare_fluid_substep(P, G, iterations, literal, exact, check_lists=True, sph=0):
    n = P.n
    force_state(P, G)
    P.s.sph_kernel = sph  # 1: Simulation::W / gradW = poly6_kernel / spiky_kernel (src/Kernels.cpp:43-67)
    P.L.lo_step_fluid(P.p, 0.01, iterations, 1, int(literal))
    G.step_fluid(dt=0.01, iterations=iterations, literal_lambda_index=int(literal), exact_math=int(exact), sph_kernel=sph)
    orig = G.dump(lgpu.DUMP_ORIG)
    assert np.array_equal(np.sort(orig), np.arange(n))
    # keys and sort order: bit-exact
    assert np.array_equal(G.dump(lgpu.DUMP_KEYS), P.keys[orig]), "cell keys"
    assert np.array_equal(G.dump(lgpu.DUMP_PERM), orig)
    keys = G.dump(lgpu.DUMP_KEYS)
    assert np.all(np.diff(keys) >= 0), "storage sorted by cell"
    same = np.diff(keys) == 0
    assert np.all(np.diff(orig)[same] > 0), "stable: ascending reference slot inside a cell"
    if check_lists:
        # neighbour lists: bit-exact, list order included (sand slots mapped back to reference slots)
        goff, gflat = G.neighbors()
        poff, pflat = P.neighbors()
        gmap = np.where(gflat < n, orig[np.minimum(gflat, n - 1)], gflat)
        assert np.array_equal(goff[1:] - goff[:-1], (poff[1:] - poff[:-1])[orig]), "neighbour counts"
        order = np.concatenate([np.arange(poff[o], poff[o + 1]) for o in orig]) if n else np.zeros(0, np.int64)
        assert np.array_equal(gmap, pflat[order]), "neighbour lists (order included)"
    close(G.dump(lgpu.DUMP_DENSITY), P.densities[orig], "density")
    close(G.dump(lgpu.DUMP_LAMBDA), P.lambdas[orig], "lambda", atol=1e-7)
    pos, vel, _ = G.download()
    close(pos, P.positions, "position")
    close(vel, P.velocities, "velocity", atol=1e-3)  # v = dx/dt amplifies the position tolerance by 1/dt


@pytest.mark.parametrize("n_side,with_solids", [(12, False), (20, True), (40, False)])
def test_fluid_teacher_forced(n_side, with_solids):
    domain, sand = scenes.dam_break(n_side)
    solids = scenes.floor_plate(min(3 * n_side, 40), min(2 * n_side, 30)) if with_solids else None
    P, G = make_pair(domain, sand, solids)
    for step in range(4):
        print("substep", step)
        compare_fluid_substep(P, G, iterations=1, literal=True, exact=True)
    c = G.dump(lgpu.DUMP_COUNTERS)
    assert c[0] == 0 and P.s.violations == 0
    G.close(); P.close()


@pytest.mark.parametrize("iterations,literal,exact", [(4, False, True), (4, True, True), (1, True, False), (4, False, False)])
def test_fluid_modes(iterations, literal, exact):
    domain, sand = scenes.dam_break(16)
    P, G = make_pair(domain, sand, scenes.floor_plate(30, 20))
    for step in range(3):
        print("substep", step)
        compare_fluid_substep(P, G, iterations, literal, exact)
    G.close(); P.close()


@pytest.mark.parametrize("iterations,literal,with_solids", [(1, True, False), (1, True, True), (3, False, True)])
def test_fluid_poly6_spiky_solver(iterations, literal, with_solids):
    """sph_kernel=1: the solver with W = poly6_kernel(float), gradW = spiky_kernel (src/Kernels.cpp:43-67; selectable,
    never wired by the reference itself) against the port's poly6 step, which tests/test_oracle.py pins
    against the compiled reference with Simulation::W / gradW pointed at those functions."""
    domain, sand = scenes.dam_break(14)
    P, G = make_pair(domain, sand, scenes.floor_plate(30, 20) if with_solids else None)
    for step in range(3):
        print("substep", step)
        compare_fluid_substep(P, G, iterations, literal, True, sph=1)
    G.close(); P.close()


@pytest.mark.parametrize("with_solids,generic,slots", [(False, False, 0), (True, False, 0), (True, True, 0), (True, False, 600), (False, False, 40)])
def test_fluid_fast_kernels(with_solids, generic, slots):
    """Throughput arithmetic (exact_math=0), 4 iterations, lambdas[neighbour]: the specialised
    k_fluid_*_fast pair (branch-free inner spline + out-of-line correction of the neighbours that drifted
    beyond q = 0.5) and the generic kernels (test hook) both stay within the parity tolerance, also when
    the blocks run in virtual-slot / re-walk mode."""
    domain, sand = scenes.dam_break(16)
    P, G = make_pair(domain, sand, scenes.floor_plate(30, 20) if with_solids else None)
    G.set_generic_kernels(generic)
    if slots:
        G.set_stage_slots(slots)
    for step in range(4):
        print("substep", step)
        compare_fluid_substep(P, G, 4, False, False)
    G.close(); P.close()


def test_fluid_fast_matches_generic_on_a_moving_scene():
    """Fast vs generic kernels on a scene with sideways motion (many neighbours cross q = 0.5 inside a
    substep, which exercises the out-of-line correction): 5 substeps, the fast context teacher-forced to
    the generic context's state before each one, same tolerance as the oracle parity."""
    domain, sand = scenes.dam_break(24)
    vel = np.zeros_like(sand)
    vel[:, 0] = 6.0 * np.sin(sand[:, 2]); vel[:, 2] = 6.0 * np.cos(sand[:, 0])
    kw = dict(dt=0.01, iterations=4, literal_lambda_index=0, exact_math=0)
    with lgpu.Context(domain, capacity_sand=len(sand)) as A, lgpu.Context(domain, capacity_sand=len(sand)) as B:
        B.set_generic_kernels(1)
        pos = sand
        for step in range(5):
            A.upload_sand(pos, vel); B.upload_sand(pos, vel)
            A.step_fluid(**kw); B.step_fluid(**kw)
            pa, va, _ = A.download(); pos, vel, _ = B.download()
            print("substep", step)
            close(A.dump(lgpu.DUMP_LAMBDA), B.dump(lgpu.DUMP_LAMBDA), "lambda", atol=1e-7)
            close(pa, pos, "position")


def test_fluid_table_overflow_falls_back_to_walk():
    domain, sand = scenes.dam_break(12)
    P, G = make_pair(domain, sand, max_neighbors=6)
    for step in range(2):
        compare_fluid_substep(P, G, 2, True, True)
    assert G.dump(lgpu.DUMP_COUNTERS)[1] > 0, "the narrow table must have overflowed"
    G.close(); P.close()


@pytest.mark.parametrize("spacing", [0.74, 0.6])
def test_dense_scenes_spill_chunks_and_split_bricks(spacing):
    """Packings denser than the rest lattice: lists longer than the 32-entry table row continue in a spill chunk
    (33..64 entries; beyond that the particle re-walks), columns with more than 32 candidates are tested in rounds,
    and bricks whose neighbourhood exceeds a block's shared memory are cut along z.  Same parity contract."""
    n = 20
    sand = scenes.lattice(n, n, n, origin=(1.5, 1.5, 1.5), spacing=spacing, jitter=0.03, seed=4242)
    domain = (24, 24, 24)
    P, G = make_pair(domain, sand)
    for step in range(2):
        compare_fluid_substep(P, G, 2, False, False)
        if step == 0:
            cnt = G.dump(lgpu.DUMP_NBR_COUNT)
            print("spacing %.2f: list length mean %.1f max %d, longer than 32: %.0f%%, longer than 64: %.0f%%, re-walked rows %d"
                  % (spacing, cnt.mean(), cnt.max(), 100.0 * (cnt > 32).mean(), 100.0 * (cnt > 64).mean(), G.dump(lgpu.DUMP_COUNTERS)[1]))
            assert (cnt > 32).mean() > 0.3, "the scene must exercise the spill chunks"
        compare_fluid_substep(P, G, 1, True, True)
    G.close(); P.close()
    if spacing < 0.7:
        return  # (sand grains of diameter 1 at this spacing overlap by 40 %: the push-out is an explosion, not a parity scene)
    solids = scenes.floor_plate(24, 24)
    P, G = make_pair(domain, sand, solids)
    for step in range(2):
        compare_sand_substep(P, G, 4, step % 2 == 0)
    G.close(); P.close()


@pytest.mark.parametrize("slots", [600, 40])
def test_small_stage_uses_virtual_slots_or_walk(slots):
    # a stage too small for the neighbourhoods: blocks switch to virtual slots (same codes, neighbours
    # read through L1/L2) and, beyond the 16-bit code space, to the stencil re-walk; results unchanged
    domain, sand = scenes.dam_break(16)
    P, G = make_pair(domain, sand, scenes.floor_plate(30, 20))
    G.set_stage_slots(slots)
    for step in range(2):
        compare_fluid_substep(P, G, 2, False, True)
        compare_fluid_substep(P, G, 1, True, False)
    G.close(); P.close()
    domain, sand, solids = scenes.sand_pile(12, drop=1.0)
    P, G = make_pair(domain, sand, solids)
    G.set_stage_slots(slots)
    for step in range(3):
        compare_sand_substep(P, G, 4, step % 2 == 0)
    G.close(); P.close()


@pytest.mark.parametrize("n_side", [20, 40])
def test_fluid_free_running_horizon(n_side):
    # 10 free-running substeps (no teacher forcing), K=1 literal vs the Jacobi oracle; 8 000 and 64 000 particles
    domain, sand = scenes.dam_break(n_side)
    P, G = make_pair(domain, sand)
    for step in range(10):
        P.L.lo_step_fluid(P.p, 0.01, 1, 1, 1)
        G.step_fluid(dt=0.01, iterations=1, literal_lambda_index=1, exact_math=1)
    pos, vel, _ = G.download()
    d = np.abs(pos - P.positions)
    print("  horizon 10: max|dx| %.3e mean|dx| %.3e" % (d.max(), d.mean()))
    assert d.mean() <= 1e-3 * 1.0  # 1e-3 * diameter
    # neighbour sets of the last grid build
    orig = G.dump(lgpu.DUMP_ORIG)
    goff, gflat = G.neighbors()
    poff, pflat = P.neighbors()
    n = P.n
    gmap = np.where(gflat < n, orig[np.minimum(gflat, n - 1)], gflat)
    gi = np.repeat(orig, np.diff(goff))
    pi = np.repeat(np.arange(n), np.diff(poff))
    gset = set(zip(gi.tolist(), gmap.tolist()))
    pset = set(zip(pi.tolist(), pflat.tolist()))
    jac = len(gset & pset) / max(len(gset | pset), 1)
    print("  neighbour-set Jaccard %.6f" % jac)
    assert jac >= 0.999
    G.close(); P.close()


# ----------------------------------------------------------------------------------------------
def compare_sand_substep(P, G, iterations, exact, credits=False, dt=0.016, player=None, attract=False, blow=False):
    n = P.n
    force_state(P, G)
    prev = P.s.prev_attract_flag
    if player is not None:
        for a in range(3):
            P.s.player_position[a] = float(player[a])
    P.s.attract_flag, P.s.blow_flag = int(attract), int(blow)
    before_pos = P.positions.copy()
    P.L.lo_step_sand(P.p, dt, iterations, int(credits))
    kw = dict(dt=dt, iterations=iterations, exact_math=int(exact), credits=int(credits), attract_flag=int(attract),
              blow_flag=int(blow), prev_attract_flag=int(prev))
    if credits:
        kw.update(mu_s=0.8, mu_k=0.7)
    if player is not None:
        kw["player_position"] = player
    G.step_sand(**kw)
    # sort permutation: bit-exact against Sorting::counting_sort
    assert np.array_equal(G.dump(lgpu.DUMP_PERM), P.sorted_index), "sort permutation"
    assert np.array_equal(G.dump(lgpu.DUMP_KEYS), P.keys), "cell keys"
    assert np.array_equal(G.dump(lgpu.DUMP_ORIG), np.arange(n))
    # neighbour lists: the oracle keeps the reference's two self entries (F7); the device drops them
    goff, gflat = G.neighbors()
    poff, pflat = P.neighbors()
    owner = np.repeat(np.arange(n), np.diff(poff))
    keep = pflat != owner
    pcount = np.bincount(owner[keep], minlength=n)
    assert np.array_equal(np.diff(goff), pcount), "neighbour counts"
    assert np.array_equal(gflat, pflat[keep]), "neighbour lists (order included)"
    pos, vel, flags = G.download()
    close(pos, P.positions, "position")
    close(vel, P.velocities, "velocity", atol=1e-3)
    assert np.array_equal(flags, P.attracted), "attracted flags"
    return before_pos


@pytest.mark.parametrize("n_side", [8, 16])
def test_sand_teacher_forced(n_side):
    domain, sand, solids = scenes.sand_pile(n_side)
    P, G = make_pair(domain, sand, solids)
    for step in range(6):
        print("step", step)
        compare_sand_substep(P, G, 4, True)
    assert G.dump(lgpu.DUMP_COUNTERS)[0] == 0
    G.close(); P.close()


@pytest.mark.parametrize("iterations,exact", [(1, True), (2, True), (3, True), (4, False)])
def test_sand_iterations_and_fast_math(iterations, exact):
    domain, sand, solids = scenes.sand_pile(12, drop=1.0)
    P, G = make_pair(domain, sand, solids)
    for step in range(4):
        print("step", step)
        compare_sand_substep(P, G, iterations, exact)
    G.close(); P.close()


def test_sand_attract_blow_and_credits():
    domain, sand, solids = scenes.sand_pile(10, drop=1.0)
    player = (15.0, 4.0, 15.0)
    P, G = make_pair(domain, sand, solids)
    P.s.attract_radius, P.s.blow_radius = 6.0, 5.0
    schedule = [(False, False), (True, False), (True, False), (False, False), (False, True), (False, False)]
    for step, (att, blow) in enumerate(schedule):
        print("step", step, att, blow)
        force_state(P, G)
        prev = P.s.prev_attract_flag
        for a in range(3):
            P.s.player_position[a] = player[a]
        P.s.attract_flag, P.s.blow_flag = int(att), int(blow)
        P.L.lo_step_sand(P.p, 0.016, 4, 0)
        G.step_sand(dt=0.016, iterations=4, exact_math=1, attract_flag=int(att), blow_flag=int(blow),
                    prev_attract_flag=int(prev), player_position=player, attract_radius=6.0, blow_radius=5.0)
        pos, vel, flags = G.download()
        close(pos, P.positions, "position")
        close(vel, P.velocities, "velocity", atol=1e-3)
        assert np.array_equal(flags, P.attracted)
        if att:
            assert flags.sum() > 0, "somebody must be attracted in this scene"
    G.close(); P.close()
    # credits: bit 1 = no gravity until first contact/attraction/blow (src/Simulate.cpp:362-364,463-470)
    P, G = make_pair(domain, sand, solids)
    flags0 = np.full(len(sand), 2, np.int32)
    P.set_sand(sand, None, flags0)
    for step in range(4):
        print("credits step", step)
        compare_sand_substep(P, G, 4, True, credits=True)
    G.close(); P.close()


@pytest.mark.parametrize("n_side", [14, 40])
def test_sand_free_running_horizon(n_side):
    # 10 free-running steps; 2 744 and 64 000 particles
    domain, sand, solids = scenes.sand_pile(n_side, drop=2.0)
    P, G = make_pair(domain, sand, solids)
    for step in range(10):
        P.L.lo_step_sand(P.p, 0.016, 4, 0)
        G.step_sand(dt=0.016, iterations=4, exact_math=1)
    pos, vel, _ = G.download()
    d = np.abs(pos - P.positions)
    print("  horizon 10: max|dx| %.3e mean|dx| %.3e bit-exact %.4f" % (d.max(), d.mean(), (pos == P.positions).mean()))
    assert d.mean() <= 1e-3
    G.close(); P.close()


# ----------------------------------------------------------------------------------------------
def aabb_first_k_numpy(pos, center, half, k):
    """particle_collide_with_player + the first-k loop of set_particles_box_colliders_positions
    (src/BulletPhysics.cpp:602-606,623-642): particles in index order with |p - player| <= half on every axis,
    stopping once k proxy boxes are placed."""
    d = np.abs(pos - np.asarray(center, np.float32)[None, :])
    inside = np.all(d <= np.asarray(half, np.float32)[None, :], axis=1)
    return pos[np.nonzero(inside)[0][:k]]


@pytest.mark.parametrize("mode", ["sand", "fluid"])
def test_aabb_first_k_matches_the_reference_scan(mode):
    """lgpu_aabb_first_k (the feed of the Bullet proxy boxes): same particles, same order, same
    truncation as the reference's host loop over simulation.positions — after a sand step (storage permuted
    by the sort) and after fluid steps (storage order kept, device order sorted)."""
    domain, sand, solids = scenes.sand_pile(16, drop=1.0)
    with lgpu.Context(domain, capacity_sand=len(sand), capacity_solid=len(solids)) as G:
        G.upload_sand(sand); G.upload_solids(solids)
        for _ in range(3):
            if mode == "sand":
                G.step_sand(dt=0.016, iterations=4, exact_math=1)
            else:
                G.step_fluid(dt=0.01, iterations=1, literal_lambda_index=1, exact_math=1)
        pos, _, _ = G.download()  # the reference's storage order
        center = pos.mean(axis=0) + np.float32(0.3)
        scale = np.array([4.0, 5.0, 4.0], np.float32) + np.float32(0.5 * 4.0)  # player_box_scale + 4 r (:604)
        half = scale * np.float32(0.5)
        hits = len(aabb_first_k_numpy(pos, center, half, 10 ** 9))
        assert hits > 100, "the scene must have more candidates than proxy boxes"
        for k in (100, 7, 1, hits, hits + 50):
            want = aabb_first_k_numpy(pos, center, half, k)
            got = G.aabb_first_k(center, half, k)
            assert got.shape == want.shape and np.array_equal(got, want), "k=%d" % k
        far = G.aabb_first_k((-50.0, -50.0, -50.0), half, 100)
        assert far.shape == (0, 3)
        assert G.aabb_first_k(center, half, 0).shape == (0, 3)


def test_queries_on_a_dense_domain_with_capacity_equal_to_n():
    """More particles than grid cells (n > C + 16k) and capacity == n: the scans of lgpu_remove_in_cells and
    lgpu_aabb_first_k run over the PARTICLES (scan status words and scratch were sized by the cell
    count).  Run under compute-sanitizer by tools/gpu_round.sh."""
    domain = (40, 40, 40)
    sand = scenes.lattice(38, 38, 38, origin=(1.0, 1.0, 1.0), jitter=0.02)
    with lgpu.Context(domain, capacity_sand=len(sand)) as G:
        assert len(sand) > G.num_cells + 16384
        G.upload_sand(sand)
        G.step_sand(dt=0.016, iterations=1, exact_math=1)
        pos, _, _ = G.download()
        center, half = (20.0, 20.0, 20.0), (6.0, 5.0, 4.0)
        want = aabb_first_k_numpy(pos, center, half, 100)
        assert np.array_equal(G.aabb_first_k(center, half, 100), want)
        # sink over the bottom cell layers: every particle whose cell is listed goes, the others stay (as a set)
        cs = np.float32(G.cell_size)
        cell = (pos / cs).astype(np.int32)  # get_cell_id: IEEE division, truncation
        gx, gy, gz = G.grid
        ids = cell[:, 1] * gx * gz + cell[:, 0] * gz + cell[:, 2]
        sink = np.array([y * gx * gz + x * gz + z for y in range(2) for x in range(gx) for z in range(gz)], np.int32)
        doomed = np.isin(ids, sink)
        assert G.cell_count((0, 0, 0), (gx - 1, 1, gz - 1), False) == int(doomed.sum())
        removed = G.remove_in_cells(sink)
        assert removed == int(doomed.sum()) and G.n == len(sand) - removed
        left, _, _ = G.download()
        a = np.sort(left.view([("x", "f4"), ("y", "f4"), ("z", "f4")]).ravel())
        b = np.sort(pos[~doomed].view([("x", "f4"), ("y", "f4"), ("z", "f4")]).ravel())
        assert np.array_equal(a, b), "survivors"
        G.step_sand(dt=0.016, iterations=1, exact_math=1)  # the state is still steppable
        assert G.download()[0].shape == (len(sand) - removed, 3)


def test_edge_cases():
    # empty, single particle, ragged block, particles on the domain faces, coincident particles
    with lgpu.Context((20, 20, 20), capacity_sand=16) as G:
        G.upload_sand(np.zeros((0, 3), np.float32))
        G.step_fluid(dt=0.01)
        G.step_sand()
        assert G.download()[0].shape == (0, 3)
    for mode in ("fluid", "sand"):
        pts = np.array([[10.0, 10.0, 10.0]], np.float32)
        P, G = make_pair((20, 20, 20), pts)
        if mode == "fluid":
            compare_fluid_substep(P, G, 1, True, True)
        else:
            compare_sand_substep(P, G, 4, True)
        G.close(); P.close()
    ragged = scenes.lattice(7, 3, 11, origin=(0.5, 0.5, 0.5))
    corners = np.array([[0.5, 0.5, 0.5], [19.5, 19.5, 19.5], [0.5, 19.5, 0.5], [19.4, 0.6, 19.4]], np.float32)
    coincident = np.array([[5.0, 9.0, 5.0], [5.0, 9.0, 5.0], [5.0, 9.0, 5.0]], np.float32)
    pts = np.concatenate([ragged, corners, coincident]).astype(np.float32)
    for mode in ("fluid", "sand"):
        P, G = make_pair((20, 20, 20), pts)
        for step in range(3):
            if mode == "fluid":
                compare_fluid_substep(P, G, 1, True, True)
            else:
                compare_sand_substep(P, G, 4, True)
        G.close(); P.close()


def test_against_compiled_reference_if_present():
    """Same comparison straight against the unmodified reference (oracle/_ref, built where
    /root/reference exists and shipped to the GPU box as a binary)."""
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    domain, sand = scenes.dam_break(14)
    solids = scenes.floor_plate(20, 20)
    # fluid: keys / lists / lambdas vs the literal reference, positions vs O-jac
    R = O.RefSim(*domain, n_sand=len(sand), n_solid=len(solids))
    R.set_sand(sand); R.set_solid(solids)
    R.set_fun(R.FLUID_JACOBI, 1, True)
    with lgpu.Context(domain, capacity_sand=len(sand), capacity_solid=len(solids)) as G:
        G.upload_sand(sand); G.upload_solids(solids)
        for step in range(3):
            rp, _, rv, ra = R.get_sand()
            G.upload_sand(rp, rv, ra)
            R.step(0.01)
            G.step_fluid(dt=0.01, iterations=1, literal_lambda_index=1, exact_math=1)
            orig = G.dump(lgpu.DUMP_ORIG)
            close(G.dump(lgpu.DUMP_LAMBDA), R.lambdas()[orig], "lambda", atol=1e-7)
            roff, rflat = R.neighbors()
            goff, gflat = G.neighbors()
            n = len(sand)
            gmap = np.where(gflat < n, orig[np.minimum(gflat, n - 1)], gflat)
            order = np.concatenate([np.arange(roff[o], roff[o + 1]) for o in orig])
            assert np.array_equal(gmap, rflat[order])
            rp, _, rv, _ = R.get_sand()
            pos, vel, _ = G.download()
            close(pos, rp, "position")
    R.close()
    # sand vs the unmodified simulate_sand
    domain, sand, solids = scenes.sand_pile(10, drop=1.0)
    R = O.RefSim(*domain, n_sand=len(sand), n_solid=len(solids))
    R.set_sand(sand); R.set_solid(solids)
    R.set_fun(R.SAND)
    with lgpu.Context(domain, capacity_sand=len(sand), capacity_solid=len(solids)) as G:
        G.upload_solids(solids)
        for step in range(4):
            rp, _, rv, ra = R.get_sand()
            G.upload_sand(rp, rv, ra)
            R.step(0.016)
            G.step_sand(dt=0.016, iterations=4, exact_math=1)
            assert np.array_equal(G.dump(lgpu.DUMP_KEYS), R.sorted_cell_ids())
            rp, _, rv, _ = R.get_sand()
            pos, vel, _ = G.download()
            close(pos, rp, "position")
            close(vel, rv, "velocity", atol=1e-3)
    R.close()


# ----------------------------------------------------------------------------------------------
def test_full_size_properties_1m():
    """BASELINE size (configs[1]: 100^3 = 1M particles, K = 4), where the oracle is too slow to run per
    substep: size-independent properties of one substep.
      * keys = get_cell_id of the predicted positions (recomputed here in numpy fp32), storage sorted by
        key, stable inside a cell, cell offsets consistent with the keys;
      * neighbour lists: self included exactly once, every entry within h of the build-time x*,
        symmetric (j in N(i) <=> i in N(j)), ascending sorted slot; counts equal to a brute-force count
        on a random sample of particles;
      * throughput arithmetic vs exact arithmetic from the same state: positions within the parity
        tolerance; fast kernels vs generic kernels likewise."""
    domain, sand = scenes.dam_break(100)
    n = len(sand)
    kw = dict(dt=0.01, iterations=4, literal_lambda_index=0)
    out = {}
    with lgpu.Context(domain, capacity_sand=n) as G:
        G.upload_sand(sand)
        G.step_fluid(exact_math=1, **kw)
        keys = G.dump(lgpu.DUMP_KEYS).astype(np.int64)
        orig = G.dump(lgpu.DUMP_ORIG)
        cell_start = G.dump(lgpu.DUMP_CELL_START).astype(np.int64)
        off, flat = G.neighbors()
        out["exact"] = G.download()[0]
        grid = {"grid": tuple(G.grid), "cell_size": G.cell_size, "kernel_radius": G.kernel_radius}
    # predicted positions of the substep as the reference computes them (src/Simulate.cpp:49-50, zero initial
    # velocity): v = gravity * mass * dt, x* = x + v * dt, every operation rounded to fp32
    p = lgpu.default_step_params(**kw)
    dt = np.float32(min(max(p.dt, 0.001), 0.01))
    xs = sand.copy()
    for a in range(3):
        va = np.float32(np.float32(np.float32(p.gravity[a]) * np.float32(p.mass)) * dt)
        xs[:, a] = xs[:, a] + np.float32(va * dt)
    cs = np.float32(grid["cell_size"])
    c = (xs / cs).astype(np.int64)
    gX, gY, gZ = grid["grid"]
    ref_keys = c[:, 1] * gX * gZ + c[:, 0] * gZ + c[:, 2]
    assert np.array_equal(keys, ref_keys[orig]), "cell keys at 1M"
    assert np.all(np.diff(keys) >= 0)
    assert np.all(np.diff(orig)[np.diff(keys) == 0] > 0), "stable inside a cell"
    assert np.array_equal(cell_start[:-1], np.searchsorted(keys, np.arange(len(cell_start) - 1))), "cell offsets"
    assert cell_start[-1] == n
    # neighbour lists
    cnt = np.diff(off)
    owner = np.repeat(np.arange(n), cnt)
    assert flat.min() >= 0 and flat.max() < n
    assert int((flat == owner).sum()) == n, "self exactly once in every list"
    xs_sorted = xs[orig]
    d = xs_sorted[owner] - xs_sorted[flat]
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    h = np.float32(grid["kernel_radius"])
    assert np.all(d2 <= h * h), "every list entry within h"
    seg_sorted = np.ones(len(flat), bool)
    seg_sorted[1:] = (np.diff(flat) > 0) | (np.diff(owner) != 0)
    assert seg_sorted.all(), "lists in ascending sorted slot"
    pair = owner.astype(np.int64) * n + flat
    rev = flat.astype(np.int64) * n + owner
    assert np.array_equal(np.sort(pair), np.sort(rev)), "neighbour relation is symmetric"
    rng = np.random.default_rng(7)
    for i in rng.integers(0, n, 64):
        dd = xs_sorted - xs_sorted[i]
        r2 = (dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]) + dd[:, 2] * dd[:, 2]
        assert int((r2 <= h * h).sum()) == cnt[i], "brute-force neighbour count"
    # arithmetic policies and kernel variants from the same state
    for name, exact, generic in (("fast", 0, 0), ("generic", 0, 1)):
        with lgpu.Context(domain, capacity_sand=n) as G:
            G.set_generic_kernels(generic)
            G.upload_sand(sand)
            G.step_fluid(exact_math=exact, **kw)
            out[name] = G.download()[0]
    close(out["fast"], out["exact"], "fast vs exact")
    close(out["generic"], out["exact"], "generic vs exact")


def test_full_size_properties_sand_4m():
    """BASELINE size (configs[2]: 160^3 = 4M sand particles on a voxel floor with obstacle boxes): the
    reference permutes its storage into the stable cell-sorted order, so after a step the storage is
    sorted by the cell of the predicted positions, the permutation is a bijection that keeps the
    previous order inside every cell, neighbour counts are symmetric in total, and throughput
    arithmetic agrees with exact arithmetic within the parity tolerance."""
    domain, sand, solids = scenes.sand_pile(160)
    n = len(sand)
    kw = dict(dt=0.016, iterations=4)
    out = {}
    for name, exact in (("exact", 1), ("fast", 0)):
        with lgpu.Context(domain, capacity_sand=n, capacity_solid=len(solids)) as G:
            G.upload_sand(sand)
            G.upload_solids(solids)
            G.step_sand(exact_math=exact, **kw)
            out[name] = G.download()[0]
            if exact:
                keys = G.dump(lgpu.DUMP_KEYS).astype(np.int64)
                perm = G.dump(lgpu.DUMP_PERM)
                cell_start = G.dump(lgpu.DUMP_CELL_START).astype(np.int64)
                cnt = G.dump(lgpu.DUMP_NBR_COUNT).astype(np.int64)
                counters = G.dump(lgpu.DUMP_COUNTERS)
    assert counters[0] == 0, "no particle left the grid"
    assert np.array_equal(np.sort(perm), np.arange(n)), "the sort permutation is a bijection"
    assert np.all(np.diff(keys) >= 0), "storage sorted by cell"
    assert np.all(np.diff(perm)[np.diff(keys) == 0] > 0), "stable: previous order kept inside a cell"
    assert np.array_equal(cell_start[:-1], np.searchsorted(keys, np.arange(len(cell_start) - 1))) and cell_start[-1] == n
    assert cnt.min() >= 0 and cnt.max() < 64 and cnt.sum() > 10 * n, "plausible list lengths"
    # Coulomb friction switches between its static and kinetic branch on `d * mu_s > |x_tan|`
    # (src/Simulate.cpp:260-266): a contact that sits on the threshold takes the other branch under FMA
    # contraction, which moves that particle by a fraction of the tangential slip.  So the bulk agrees to the
    # parity tolerance and the rest stays within a hundredth of a radius.
    err = np.abs(out["fast"].astype(np.float64) - out["exact"])
    bound = ABS_TOL + REL_TOL * np.abs(out["exact"].astype(np.float64))
    inside = float((err <= bound).mean())
    print("  fast vs exact: %.5f of the coordinates inside the tolerance, max|err| %.3e" % (inside, err.max()))
    assert inside >= 0.999 and err.max() <= 5e-3
    r = np.float32(0.5)
    hi = np.array(domain, np.float32) - r
    assert np.all(out["exact"] >= r) and np.all(out["exact"] <= hi), "positions clamped to the box (src/Simulate.cpp:307)"


def test_cuda_graph_replay_is_identical():
    """lgpu_set_use_graph: the substep captured with cudaStreamBeginCapture and replayed with cudaGraphLaunch gives
    bit-identical results (fluid and sand), and the replay really happens (lgpu_graph_stats)."""
    domain, sand = scenes.dam_break(20)
    res = []
    for graph in (0, 1):
        with lgpu.Context(domain, capacity_sand=len(sand)) as G:
            G.set_use_graph(graph)
            G.upload_sand(sand)
            for _ in range(5):
                G.step_fluid(dt=0.01, iterations=3, literal_lambda_index=0, exact_math=0)
            res.append(G.download())
            captures, replays = G.graph_stats()
            print("  fluid graph=%d: %d captures, %d replays" % (graph, captures, replays))
            assert (captures, replays) == ((1, 4) if graph else (0, 0))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    domain, sand, solids = scenes.sand_pile(12, drop=1.0)
    res = []
    for graph in (0, 1):
        with lgpu.Context(domain, capacity_sand=len(sand), capacity_solid=len(solids)) as G:
            G.set_use_graph(graph)
            G.upload_sand(sand); G.upload_solids(solids)
            for k in range(6):
                # the player moves on steps 2 and 3 (the parameters change: the instantiated graph is updated in
                # place), then stands still (replays)
                player = (18.0 + min(k, 3), 4.0, 18.0)
                G.step_sand(dt=0.016, iterations=4, exact_math=0, attract_flag=1, attract_radius=6.0, player_position=player)
            res.append(G.download())
            captures, replays = G.graph_stats()
            print("  sand graph=%d: %d captures, %d replays" % (graph, captures, replays))
            assert (captures, replays) == ((4, 2) if graph else (0, 0))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(res[0][2], res[1][2]), "attracted flags"
