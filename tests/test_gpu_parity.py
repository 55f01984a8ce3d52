"""GPU parity tests (pytest -m gpu): the CUDA path through the C-ABI against the oracle on the same seeded
inputs and against the committed golden fixtures recorded from the reference.

Bar: integer observables (broadphase pair set, per-pair contact counts, island labels, the four dynamic
iteration counters, dRand seed) identical; floating observables BIT-IDENTICAL wherever the path uses only +,-,*,/,sqrt (the build
uses -fmad=false) or single-precision atan2 (cullPoints, hinge / universal / motor angles), which the library computes with the
host libm's own algorithm (fdlibm atan2f restated in odeb_math.cuh, compared bit for bit with the host's atan2f in tests/test_capi.py).
Measured on B200 (tools/measure_tolerances.py, profiles/r2_tolerances.txt): in SINGLE precision every scene -- box stacks, chains, free
boxes, both piles, the capsule ragdoll, the full 100k-box wall -- reproduces the oracle / the recorded reference trajectories bit for bit,
free-running.  In DOUBLE precision atan2 is CUDA's libm (<= 2 ulp from glibc's): everything but the ragdoll is bit-identical, the ragdoll has
stated tolerances = 4x the measured maxima: teacher-forced single step 6.6e-14 absolute on body state (measured 1.64e-14), golden trajectory
with every <= 16-step segment re-synchronised to the recorded reference state 4.2e-10 on state and 4.1e-12 on contact geometry (measured
1.05e-10 / 1.0e-12): one ulp in a joint-limit error is amplified by the contact dynamics within a segment.
"""
import os
import numpy as np
import pytest
from parity_util import B, ROOT, REAL, gpu_lib, orc_lib, compare_step
import golden_cases as G
from ode_b200 import scenes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
PRECS = ("single", "double")
# 4x the maxima measured on B200 (tools/measure_tolerances.py); only scenes with hinge / universal angles need them
TOL_TF = {"single": dict(contact=2.4e-5, state=6.7e-5), "double": dict(contact=4.1e-12, state=6.6e-14)}
# joint-family tests further down (hinge2 stops, motors in Euler mode, rolling friction, kinematic bodies, rays, cylinders): an upper bound per
# teacher-forced step from round 1, not re-measured per scene; used in DOUBLE precision only (CUDA's atan2) -- in single these tests run
# free and are compared bit-exactly
TOL = {"single": dict(contact=2e-5, state=2e-5), "double": dict(contact=1e-12, state=1e-12)}
TOL_FREE = {"single": dict(contact=2.4e-5, state=3.2e-3), "double": dict(contact=4.1e-12, state=4.2e-10)}
# Every scene is compared bit-exactly, free-running (measured deviation: 0), except the ragdoll in DOUBLE precision: its hinge / universal
# angles go through atan2, which the single build computes exactly like the reference's host libm (odeb_math.cuh: the fdlibm atan2f restated
# operation by operation, tests/test_capi.py) and the double build takes from CUDA's libm (within 2 ulp of glibc's).
def EXACT(name, prec):
    return name != "ragdoll" or prec == "single"


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", sorted(G.TRAJ_SCENES))
def test_golden_trajectories(prec, name):
    mk, h, nsteps, every = G.TRAJ_SCENES[name]
    gold = np.load(os.path.join(GOLD, "traj_%s_%s.npz" % (name, prec)))
    bad = G.compare_traj(B.Batch(gpu_lib(prec), mk()), gold, h, exact=EXACT(name, prec), tol=TOL_FREE[prec], resync=not EXACT(name, prec))
    assert not bad, bad


@pytest.mark.parametrize("prec", PRECS)
def test_free_running_vs_oracle_bit_exact(prec):
    """Box stacks (both world-option sets), chains and free boxes: every observable identical, every step."""
    for mk, h, n in ((lambda: scenes.box_stack(nworlds=5), 0.02, 150),
                     (lambda: scenes.box_stack(nworlds=3, demo_world_options=False), 0.02, 150),
                     (lambda: scenes.chain(4), 0.05, 150),
                     (lambda: scenes.free_boxes(3, 16, grid=4), 0.01, 100)):
        sc = mk()
        a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
        for s in range(n):
            a.step(h)
            b.step(h)
            bad = compare_step(a, b, sc.nworlds)
            assert not bad, (s, bad)


@pytest.mark.parametrize("prec", PRECS)
def test_teacher_forced_single_step(prec):
    """SURVEY 8(d) protocol (i): upload the oracle's state at steps {0,10,100,300}, step once on both,
    compare. Scenes with transcendental calls (pile: cullPoints/atan2, ragdoll: hinge angles) included."""
    for mk, h in ((lambda: scenes.pile(nbodies=125), 0.01), (lambda: scenes.ragdoll(3), 0.01),
                  (lambda: scenes.box_stack(nworlds=2, demo_world_options=False), 0.02)):   # no auto-disable: its history is not part of the uploaded state
        sc = mk()
        a = B.Batch(orc_lib(prec), sc)
        b = B.Batch(gpu_lib(prec), sc)
        done = 0
        for target in (0, 10, 100, 300):
            a.step(h, target - done)
            done = target
            st = a.get_state()
            b.set_state(**st)
            b.set_seeds(a.get_seeds())
            a.set_state(**st)        # same normalise + R rebuild on both sides
            a.step(h)
            b.step(h)
            done += 1
            exact = prec == "single" or sc.nbody != scenes.ragdoll(1).nbody   # bit-exact but for the ragdoll in double (hinge angles through CUDA's atan2)
            bad = compare_step(a, b, sc.nworlds, exact_float=exact, tol=TOL_TF[prec], what=("pairs", "contacts", "islands", "seeds", "state"))
            assert not bad, (target, bad)


@pytest.mark.parametrize("prec", PRECS)
def test_full_size_properties(prec):
    """BASELINE config sizes, size-independent properties: replicated worlds stay replicated (4096 stacks built
    from 8 distinct seeds -> only 8 distinct trajectories), results do not depend on batch position or batch
    size, state stays finite, the stack neither sinks nor explodes."""
    W = 4096
    sc = scenes.box_stack(nworlds=W, demo_world_options=False)
    for k in sc.state:
        sc.state[k] = np.ascontiguousarray(np.tile(sc.state[k][:8], (W // 8, 1, 1)))
    sc.seeds = np.tile(sc.seeds[:8], W // 8)
    b = B.Batch(gpu_lib(prec), sc)
    b.step(0.02, 60)
    st = b.get_state()
    for k in st:
        assert np.isfinite(st[k]).all()
        v = st[k].reshape(W // 8, 8, -1)
        assert np.array_equal(v, np.broadcast_to(v[:1], v.shape)), "%s differs between replicas" % k
    seeds = b.get_seeds().reshape(W // 8, 8)
    assert np.array_equal(seeds, np.broadcast_to(seeds[:1], seeds.shape))
    # the same 8 worlds alone give the same answer (no dependence on batch size / placement)
    sc8 = scenes.box_stack(nworlds=8, demo_world_options=False)
    b8 = B.Batch(gpu_lib(prec), sc8)
    b8.step(0.02, 60)
    st8 = b8.get_state()
    for k in st:
        assert np.array_equal(st[k][:8], st8[k])
    z = st["pos"][:8, :, 2]
    assert (z > 0.2).all() and (z < 9.0).all()
    # and the oracle agrees on those 8 worlds bit for bit
    o = B.Batch(orc_lib(prec), sc8)
    o.step(0.02, 60)
    so = o.get_state()
    for k in st8:
        assert np.array_equal(so[k], st8[k])


@pytest.mark.parametrize("prec", PRECS)
def test_large_island_global_path(prec):
    """A pile big enough that its island exceeds the shared-memory solve budget takes the global-memory
    solve path; islands, pair sets and contact counts stay identical to the oracle, state within tolerance."""
    sc = scenes.pile(nbodies=343)
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    for s in range(70):
        a.step(0.01)
        b.step(0.01)
        if s % 10 == 9:
            st = a.get_state()
            b.set_state(**st)
            a.set_state(**st)
            b.set_seeds(a.get_seeds())
    a.step(0.01)
    b.step(0.01)
    bad = compare_step(a, b, 1, exact_float=prec == "single", tol=TOL[prec], what=("pairs", "contacts", "islands", "seeds", "state"))
    assert not bad, bad
    n, lab = b.get_islands(0)
    assert n >= 1


def test_capacity_overflow_is_reported():
    """Too small a contact capacity must make the step fail loudly, never drop contacts silently."""
    os.environ["ODEB_MAX_CONTACTS"] = "4"
    try:
        b = B.Batch(gpu_lib("single"), scenes.box_stack(nworlds=2, nboxes=6))
    finally:
        del os.environ["ODEB_MAX_CONTACTS"]
    with pytest.raises(RuntimeError):
        b.step(0.02, 40)


def test_capacity_overflow_in_the_async_loop():
    """in the documented asynchronous loop (odeb_add_force, odeb_step_async, odeb_get_state) nothing used to read the
    overflow flag.  Now the blocking getter reports it, the capacity comes from OdebWorldParams (no environment variable), and the
    failing step plus every step queued behind it leave the body state as the last complete step wrote it."""
    sc = scenes.box_stack(nworlds=2, nboxes=6)
    sc.wp.max_contacts_per_world = 4
    b = B.Batch(gpu_lib("single"), sc)
    b.step_async(0.02, 60)
    with pytest.raises(RuntimeError, match="capacity overflow"):
        b.get_state()
    s1 = b.get_state()                       # the flag was consumed: reading the (frozen) state works
    for k in s1:
        assert np.isfinite(s1[k]).all()
    b.step_async(0.02, 5)                    # still too small: overflows again in the first of these steps, before anything moves
    with pytest.raises(RuntimeError, match="capacity overflow"):
        b.sync()
    s2 = b.get_state()
    for k in s1:
        assert np.array_equal(s1[k], s2[k]), k
    # with room for the contacts the same scene steps, and agrees with the oracle
    sc.wp.max_contacts_per_world = 0
    g, o = B.Batch(gpu_lib("single"), sc), B.Batch(orc_lib("single"), sc)
    g.step_async(0.02, 60)
    o.step(0.02, 60)
    sg, so = g.get_state(), o.get_state()
    for k in sg:
        assert np.array_equal(sg[k], so[k]), k


# ---------------------------------------------------------------------------------------------------------------
# large-world path (ODEB_MODE_CANONICAL): sort/scan/union-find pipeline + coloured tile sweeps against the oracle run in the
# same mode. Integer observables identical; floats bit-identical for scenes without libm calls.
def _canon_pair(prec, sc):
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    a.set_solver_mode(1)
    b.set_solver_mode(1)
    return a, b


@pytest.mark.parametrize("prec", PRECS)
def test_canonical_mode_bit_exact(prec):
    cases = ((lambda: scenes.wall(12, 8, max_contacts=8), 0.05, 40),                     # one island, SAP space, cannon ball hits at ~step 8
             (lambda: scenes.wall(9, 5, max_contacts=8, space_type=B.SPACE_HASH, ball=False), 0.05, 25),
             (lambda: scenes.free_boxes(1, 100, grid=10), 0.01, 40),                      # 100 one-body islands
             (lambda: scenes.box_stack(nworlds=1, nboxes=16), 0.02, 120),                 # auto-disable + re-enable path
             (lambda: scenes.chain(1), 0.05, 60))                                         # permanent joints
    for mk, h, n in cases:
        sc = mk()
        a, b = _canon_pair(prec, sc)
        for s in range(n):
            a.step(h)
            b.step(h)
            bad = compare_step(a, b, 1)
            assert not bad, (s, bad)


@pytest.mark.parametrize("maxgrid", (1, 3))
def test_canonical_mode_more_tiles_than_warps(maxgrid, monkeypatch):
    """The persistent sweep kernel gives every warp the tiles t, t + (all warps), ... of a colour; up to 100k boxes a B200 has more warps
    than the largest colour has tiles, so that loop (and the grid-stride loops of the colouring rounds, the body check and the island
    control) only runs once.  ODEB_LW_MAXGRID caps the cooperative grids at 1 / 3 blocks: several tiles per warp and colour, same bits."""
    monkeypatch.setenv("ODEB_LW_MAXGRID", str(maxgrid))
    for prec in PRECS:
        for mk, h, n in ((lambda: scenes.wall(24, 12, max_contacts=8), 0.05, 14), (lambda: scenes.free_boxes(1, 400, grid=20), 0.01, 12),
                         (lambda: scenes.pile(nbodies=343), 0.01, 25)):
            sc = mk()
            a, b = _canon_pair(prec, sc)
            for s in range(n):
                a.step(h)
                b.step(h)
                bad = compare_step(a, b, 1, exact_float=prec == "single" or sc.nbody != 343, tol=TOL[prec])
                assert not bad, (prec, sc.nbody, s, bad)
            b.close()


def _platform_scene(nside):
    """a dynamic platform on the plane carrying nside x nside unit boxes: nside^2 + 1 row groups act on the platform's body"""
    sc = B.Scene(B.default_world_params(gravity=(0, 0, -9.81), max_contacts=4, surf_mode=B.CONTACT_APPROX1, mu=0.5), 1)
    sc.add_geom(B.PLANE, (0, 0, 1, 0))
    side = 2.0 * nside + 2.0
    m, I = B.box_mass(1.0, side, side, 0.5)
    p = sc.add_body(m, I, (0, 0, 0.25))
    sc.add_geom(B.BOX, (side, side, 0.5), body=p)
    m, I = B.box_mass(1.0, 1, 1, 1)
    for i in range(nside * nside):
        bb = sc.add_body(m, I, ((i % nside) * 2.0 - (nside - 1.0), (i // nside) * 2.0 - (nside - 1.0), 1.0))
        sc.add_geom(B.BOX, (1, 1, 1), body=bb)
    n = nside * nside + 1
    pos = np.asarray(sc.body_pos, dtype=np.float64)[None]
    quat = np.tile(np.array([1.0, 0, 0, 0])[None, None], (1, n, 1))
    sc.state = dict(pos=pos, quat=quat, lvel=np.zeros_like(pos), avel=np.zeros_like(pos))
    sc.seeds = np.array([7], dtype=np.uint32)
    return sc


def test_canonical_mode_many_groups_on_one_body():
    """A body that carries many contact pairs needs as many colours (every group on it conflicts with every other one): 82 groups on the
    platform of a 9 x 9 load -> more than 82 colours, bit-identical to the oracle; beyond ODEB_CANON_COLOURS - 2 = 254 groups on one body
    the step is refused loudly and the state stays the one of the last complete step."""
    sc = _platform_scene(9)
    a, b = _canon_pair("single", sc)
    for s in range(12):
        a.step(0.01)
        b.step(0.01)
        bad = compare_step(a, b, 1)
        assert not bad, (s, bad)
    b.close()
    sc = _platform_scene(16)
    b = B.Batch(gpu_lib("single"), sc)
    b.set_solver_mode(1)
    before = b.get_state()
    with pytest.raises(RuntimeError, match="row groups on one body"):
        b.step(0.01, 2)
        b.get_state()
    after = b.get_state()
    assert np.array_equal(before["pos"], after["pos"])
    b.close()


def test_canonical_mode_full_size_wall():
    """BASELINE configs[4] at its full size (500 x 200 bricks + cannon ball = 100 001 bodies, ~3.05 M rows, 1.14 M contacts, sweep-and-prune
    space): three steps of the CUDA large-world path against the oracle in the same mode, every observable bit for bit (single precision;
    the oracle needs ~20 s per step on one core).  From step 1 on the bricks are no longer axis-aligned and box pairs with more than 4
    candidate points go through cullPoints' atan2: with CUDA's own atan2f, 5 of the 1 137 774 contacts of step 1 kept a different corner
    (3 pairs of bricks touching edge-on, two candidates a rounding error apart in angle); with the host libm's algorithm restated in
    odeb_math.cuh there is no difference left."""
    sc = scenes.wall(500, 200)
    a, b = _canon_pair("single", sc)
    for s in range(3):
        a.step(0.05)
        b.step(0.05)
        bad = compare_step(a, b, 1)
        assert not bad, (s, bad)
    assert b.get_totals()[2] > 3000000


@pytest.mark.parametrize("prec", PRECS)
def test_canonical_mode_joint_feedback(prec):
    """Joint feedback on the large-world path: the lambdas live in the tile layout during the sweeps and go back to row order for
    Stage 4b (k_lwt_lambda_out); permanent joints + contacts, every step identical to the oracle in the same mode."""
    from test_oracle import compare_feedback
    for mk, h, n in ((lambda: scenes.chain(1), 0.05, 40), (lambda: scenes.wall(6, 4, max_contacts=8), 0.05, 20)):
        sc = mk()
        a, b = _canon_pair(prec, sc)
        a.enable_feedback()
        b.enable_feedback()
        for s in range(n):
            a.step(h)
            b.step(h)
            bad = compare_step(a, b, 1) + compare_feedback(a, b, 1, True, 0)
            assert not bad, (s, bad)


@pytest.mark.parametrize("prec", PRECS)
def test_canonical_mode_teacher_forced(prec):
    """Scenes with cullPoints/atan2 (<= 4 contacts per box pair) and hinge angles: teacher-forced single steps."""
    for mk, h, targets in ((lambda: scenes.pile(nbodies=1000), 0.01, (0, 40, 80)),
                           (lambda: scenes.pile(nbodies=512, space_type=B.SPACE_SAP), 0.01, (0, 60)),
                           (lambda: scenes.wall(30, 20), 0.05, (0, 12)),
                           (lambda: scenes.ragdoll(1), 0.01, (0, 50))):
        sc = mk()
        a, b = _canon_pair(prec, sc)
        done = 0
        for target in targets:
            a.step(h, target - done)
            done = target
            st = a.get_state()
            b.set_state(**st)
            b.set_seeds(a.get_seeds())
            a.set_state(**st)
            a.step(h)
            b.step(h)
            done += 1
            bad = compare_step(a, b, 1, exact_float=prec == "single", tol=TOL[prec], what=("pairs", "contacts", "islands", "seeds", "state"))
            assert not bad, (target, bad)


@pytest.mark.parametrize("prec", PRECS)
def test_canonical_mode_vs_compiled_reference(prec):
    """ODEB_MODE_CANONICAL pinned to the UNMODIFIED reference (oracle/_ref), not only to the oracle run in the same invented mode:
    the canonical order only changes the order in which rows are relaxed, so everything a teacher-forced step decides before the
    solver runs must equal what the reference computes from the same state -- broadphase pair set (dxSAPSpace / dxHashSpace
    semantics), per-pair contact counts and contact geometry, island partition and numbering.  Scenes: the wall of
    BASELINE configs[4] at 30 x 20 and at its full 500 x 200 = 100k boxes (one step; single precision only: the reference needs
    ~45 s of CPU for it), the 1000-body pile of configs[0] with the sweep-and-prune space."""
    from parity_util import ref_lib
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("reference build (oracle/_ref) not present")
    cases = [(lambda: scenes.wall(30, 20), 0.05, (0, 12)),
             (lambda: scenes.pile(nbodies=1000, space_type=B.SPACE_SAP), 0.01, (0, 40))]
    if prec == "single":
        cases.append((lambda: scenes.wall(500, 200), 0.05, (0,)))
    for mk, h, targets in cases:
        sc = mk()
        a = B.Batch(ref, sc)
        b = B.Batch(gpu_lib(prec), sc)
        b.set_solver_mode(1)
        done = 0
        for target in targets:
            a.step(h, target - done)
            done = target
            st = a.get_state()
            b.set_state(**st)
            b.set_seeds(a.get_seeds())
            a.set_state(**st)
            a.step(h)
            b.step(h)
            done += 1
            bad = compare_step(a, b, 1, exact_float=prec == "single", tol=TOL[prec], what=("pairs", "contacts", "islands"))
            assert not bad, (sc.nbody, target, bad)
            npairs = len(b.get_pairs(0, 1 << 21))
            assert npairs > sc.nbody // 2
        b.close()
        a.close()


def _sample_worlds(full_scene, worlds, small_scene):
    """the sub-batch made of the given worlds of a full-size scene (worlds are independent: same state + seed -> same trajectory)"""
    for k in full_scene.state:
        small_scene.state[k] = np.ascontiguousarray(full_scene.state[k][worlds])
    small_scene.seeds = np.ascontiguousarray(full_scene.seeds[worlds])
    return small_scene


def test_full_size_pile64_batch_sampled_worlds():
    """The north-star shape at the bench's size: 4096 worlds of a 64-body pile (boxes + spheres, ~13 islands per world, walker solver
    k_solve6).  After 120 free-running steps (the piles have landed and are settling: several islands per world, reorders, early exits) 48
    worlds sampled across the batch carry exactly the state, dRand seed and iteration counters the oracle computes for those worlds alone
    (single precision: bit for bit)."""
    W = 4096
    sc = scenes.pile(nworlds=W, nbodies=64, vary=0.05)
    worlds = np.unique(np.concatenate([np.arange(0, W, W // 40), [1, 2, 3, 5, 2047, 2048, 4094, 4095]]))[:48]
    b = B.Batch(gpu_lib("single"), sc)
    b.step(0.01, 120)
    st, seeds = b.get_state(), b.get_seeds()
    a = B.Batch(orc_lib("single"), _sample_worlds(sc, worlds, scenes.pile(nworlds=len(worlds), nbodies=64, vary=0.05)))
    a.step(0.01, 120)
    so = a.get_state()
    for k in ("pos", "quat", "lvel", "avel"):
        assert np.array_equal(st[k][worlds], so[k]), k
    assert np.array_equal(seeds[worlds], a.get_seeds())
    for j, w in enumerate(worlds):
        assert np.array_equal(b.get_pairs(int(w)), a.get_pairs(j))
        assert np.array_equal(b.get_contacts(int(w))[1], a.get_contacts(j)[1])
        na, la = a.get_islands(j)
        nb_, lb = b.get_islands(int(w))
        assert na == nb_ and np.array_equal(la, lb)
        assert np.array_equal(b.get_stats(int(w)), a.get_stats(j))
    b.close()


@pytest.mark.parametrize("prec", PRECS)
def test_full_size_chain_batch_sampled_worlds(prec):
    """BASELINE configs[2] at its full size (65536 worlds of the 10-link chain on one GPU): 64 worlds sampled across the batch
    must carry, after 40 steps, exactly the state and dRand seed the oracle computes for those worlds alone (bit-exact scene)."""
    W = 65536
    sc = scenes.chain(W)
    worlds = np.unique(np.concatenate([np.arange(0, W, W // 48), [1, 15, 16, 17, 4095, 4096, 32767, 32768, 65534, 65535]]))[:64]
    b = B.Batch(gpu_lib(prec), sc)
    b.step(0.05, 40)
    st, seeds = b.get_state(), b.get_seeds()
    small = _sample_worlds(sc, worlds, scenes.chain(len(worlds)))
    a = B.Batch(orc_lib(prec), small)
    a.step(0.05, 40)
    so = a.get_state()
    for k in ("pos", "quat", "lvel", "avel"):
        assert np.array_equal(st[k][worlds], so[k]), k
    assert np.array_equal(seeds[worlds], a.get_seeds())
    for j, w in enumerate(worlds[:8]):
        assert np.array_equal(b.get_pairs(int(w)), a.get_pairs(j))
        assert np.array_equal(b.get_stats(int(w)), a.get_stats(j))
    b.close()


@pytest.mark.parametrize("prec", PRECS)
def test_full_size_ragdoll_batch_sampled_worlds(prec):
    """BASELINE configs[3] per-GPU size (2048 capsule ragdolls): 64 sampled worlds against the oracle after 25 free-running steps.
    Single precision: the sampled worlds' state bit for bit.  Double: hinge / universal angles go through CUDA's atan2, so the state is
    compared to the free-running tolerance (4.2e-10); pair sets, contact counts, island labels and the dRand seed of the last step must
    be identical in both."""
    W = 2048
    sc = scenes.ragdoll(W)
    worlds = np.unique(np.concatenate([np.arange(0, W, W // 56), [1, 3, 4, 5, 2046, 2047]]))[:64]
    b = B.Batch(gpu_lib(prec), sc)
    b.step(0.01, 25)
    st, seeds = b.get_state(), b.get_seeds()
    a = B.Batch(orc_lib(prec), _sample_worlds(sc, worlds, scenes.ragdoll(len(worlds))))
    a.step(0.01, 25)
    so = a.get_state()
    tol = TOL_FREE[prec]["state"]
    for k in ("pos", "quat", "lvel", "avel"):
        if prec == "single":
            assert np.array_equal(st[k][worlds], so[k]), k
        else:
            assert np.abs(st[k][worlds].astype(np.float64) - so[k]).max() <= tol, k
    assert np.array_equal(seeds[worlds], a.get_seeds())
    for j, w in enumerate(worlds):
        assert np.array_equal(b.get_pairs(int(w)), a.get_pairs(j))
        assert np.array_equal(b.get_contacts(int(w))[1], a.get_contacts(j)[1])
        na, la = a.get_islands(j)
        nb_, lb = b.get_islands(int(w))
        assert na == nb_ and np.array_equal(la, lb)
    b.close()


@pytest.mark.parametrize("prec", PRECS)
def test_pile_1000_replay_mode_teacher_forced(prec):
    """BASELINE configs[0] (1000 boxes + spheres, hash space) in the default replay mode (the reference's own row order and dRand
    reorders, one island of ~7000 rows): teacher-forced single steps against the oracle."""
    sc = scenes.pile(nbodies=1000)
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    done = 0
    for target in (0, 45):
        a.step(0.01, target - done)
        done = target
        st = a.get_state()
        b.set_state(**st)
        b.set_seeds(a.get_seeds())
        a.set_state(**st)
        a.step(0.01)
        b.step(0.01)
        done += 1
        bad = compare_step(a, b, 1, what=("pairs", "contacts", "islands", "seeds", "state"))     # bit-exact (measured: tools/measure_tolerances.py)
        assert not bad, (target, bad)
    b.close()


def test_canonical_mode_needs_single_world():
    b = B.Batch(gpu_lib("single"), scenes.box_stack(nworlds=2, nboxes=4))
    with pytest.raises(RuntimeError):
        b.set_solver_mode(1)


SOLVERS = ("v4", "p2", "p4", "p8", "bl", "hy")


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("solver", SOLVERS)
def test_solver_kernels_bit_exact(prec, solver, monkeypatch):
    """Every solver kernel (k_solve, the P-processor static-schedule k_solve5<2/4/8>, the lane-per-body k_solve_bl) keeps
    the reference's sequential row semantics: all observables identical to the oracle, every step, including worlds that
    finish their sweeps early (double precision), several islands per world, ragged world counts in a warp and the
    dRand-driven reorders."""
    monkeypatch.setenv("ODEB_SOLVER", solver)
    for mk, h, n in ((lambda: scenes.box_stack(nworlds=5, nboxes=8), 0.02, 40),
                     (lambda: scenes.box_stack(nworlds=7), 0.02, 60),
                     (lambda: scenes.box_stack(nworlds=3, nboxes=24, demo_world_options=False), 0.02, 40),
                     (lambda: scenes.chain(3), 0.05, 60),
                     (lambda: scenes.free_boxes(2, 16, grid=4), 0.01, 40)):
        sc = mk()
        a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
        for s in range(n):
            a.step(h)
            b.step(h)
            bad = compare_step(a, b, sc.nworlds)
            assert not bad, (solver, s, bad)
        b.close()


def test_solver_kernels_agree_at_bench_size(monkeypatch):
    """Full-size property check (BASELINE configs[1] shape, 512 worlds): the three kernels produce the same bits after 40 steps."""
    ref = None
    for solver in ("v4", "p4", "bl", "hy"):
        monkeypatch.setenv("ODEB_SOLVER", solver)
        b = B.Batch(gpu_lib("single"), scenes.box_stack(nworlds=512, demo_world_options=False))
        b.step(0.02, 40)
        st = b.get_state()
        seeds = b.get_seeds()
        b.close()
        if ref is None:
            ref = (st, seeds)
        else:
            for k in ("pos", "quat", "lvel", "avel"):
                assert np.array_equal(ref[0][k], st[k]), (solver, k)
            assert np.array_equal(ref[1], seeds), solver


WALKERS = ("d1", "d2", "d4", "d8")


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("solver", WALKERS + ("auto",))
def test_walker_solver_bit_exact(prec, solver, monkeypatch):
    """k_solve6<P> (odeb_solve6.cuh: every world of a warp walks its own islands; shadow Fisher-Yates, window scheduler, world
    sorting): 64-body piles whose worlds differ in island structure (the north-star shape, >= 8 worlds, several islands of 3..400
    rows each), ragged small piles, many one-body islands, single-island stacks and chains, and a 70-box stack whose island
    exceeds nothing but exercises the 12-bit rows.  Every observable incl. the dRand seed and the four iteration counters
    identical to the oracle at every step."""
    if solver == "auto":
        monkeypatch.delenv("ODEB_SOLVER", raising=False)
    else:
        monkeypatch.setenv("ODEB_SOLVER", solver)
    for mk, h, n in ((lambda: scenes.pile(nworlds=9, nbodies=64, vary=0.05), 0.01, 200),
                     (lambda: scenes.pile(nworlds=21, nbodies=27, vary=0.1), 0.01, 120),
                     (lambda: scenes.free_boxes(3, 100, grid=10), 0.01, 20),
                     (lambda: scenes.box_stack(nworlds=7), 0.02, 60),
                     (lambda: scenes.chain(3), 0.05, 60),
                     (lambda: scenes.box_stack(nworlds=3, nboxes=70, demo_world_options=False), 0.02, 25)):
        sc = mk()
        a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
        for s in range(n):
            a.step(h)
            b.step(h)
            bad = compare_step(a, b, sc.nworlds)
            assert not bad, (solver, sc.nbody, s, bad[:4])
        b.close()


@pytest.mark.parametrize("prec", PRECS)
def test_hybrid_mixed_islands_keep_the_seed_order(prec, monkeypatch):
    """Hybrid solve (k_solve 8 sweeps + k_solve5): a world whose islands are partly above the hand-over threshold must not pause
    any island, or a later island would draw its reorders before an earlier, paused one.  ODEB_TEST_HY_ROWS
    lowers the threshold so that the small piles mix islands on both sides of it."""
    monkeypatch.setenv("ODEB_SOLVER", "hy")
    monkeypatch.setenv("ODEB_TEST_HY_ROWS", "20")
    for mk, h, n in ((lambda: scenes.pile(nworlds=21, nbodies=27, vary=0.1), 0.01, 120),
                     (lambda: scenes.pile(nworlds=5, nbodies=64, vary=0.05), 0.01, 160)):
        sc = mk()
        a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
        for s in range(n):
            a.step(h)
            b.step(h)
            bad = compare_step(a, b, sc.nworlds)
            assert not bad, (sc.nbody, s, bad[:4])
        b.close()


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("solver", ("v4", "p4", "bl", "hy", "d4", "d1"))
def test_joint_feedback_bit_exact(prec, solver, monkeypatch):
    """Joint feedback (dJointSetFeedback, quickstep.cpp:3108-3182) through the batch C-ABI: f1/t1/f2/t2 and the written /
    not-written state of every joint, CUDA path against the oracle, with every solver kernel (each writes lambda out)."""
    from test_oracle import compare_feedback
    monkeypatch.setenv("ODEB_SOLVER", solver)
    for name, mk, h, n, exact in (("stack", lambda: scenes.box_stack(nworlds=5, nboxes=6), 0.02, 50, True),
                                  ("chain", lambda: scenes.chain(3), 0.05, 50, True),
                                  ("ragdoll", lambda: scenes.ragdoll(2), 0.01, 30, prec == "single")):
        sc = mk()
        a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
        a.enable_feedback()
        b.enable_feedback()
        for s in range(n):
            a.step(h)
            b.step(h)
            if exact:
                bad = compare_step(a, b, sc.nworlds)
                assert not bad, (name, s, bad)
            bad = compare_feedback(a, b, sc.nworlds, exact, 2e-2 if prec == "single" else 1e-8)
            assert not bad, (name, solver, s, bad[:4])
        b.close()


@pytest.mark.parametrize("prec", PRECS)
def test_snapshot_restore_continues_bit_identically(prec):
    """odeb_snapshot / odeb_restore: the continuation from a restored checkpoint equals the uninterrupted run, in the same
    batch and in a freshly created one (auto-disable history, dRand seeds and statistics included)."""
    mk = lambda: scenes.box_stack(nworlds=6)            # demo world options: auto-disable + damping on
    a = B.Batch(gpu_lib(prec), mk())
    a.step(0.02, 70)
    snap = a.snapshot()
    a.step(0.02, 90)
    want, seeds, en = a.get_state(), a.get_seeds(), a.get_enabled()
    stats = [a.get_stats(w).copy() for w in range(6)]
    for target in (a, B.Batch(gpu_lib(prec), mk())):
        target.restore(snap)
        target.step(0.02, 90)
        got = target.get_state()
        for k in want:
            assert np.array_equal(want[k], got[k]), k
        assert np.array_equal(seeds, target.get_seeds()) and np.array_equal(en, target.get_enabled())
        assert all(np.array_equal(stats[w], target.get_stats(w)) for w in range(6))
    bad = B.Batch(gpu_lib(prec), scenes.box_stack(nworlds=5))
    with pytest.raises(RuntimeError):
        bad.restore(snap)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("solver", ("v4", "p4"))
def test_fixed_joints_bit_exact(prec, solver, monkeypatch):
    """Fixed joints (fixed.cpp:60-110, setFixedOrientation joint.cpp:228-284) on the GPU: every observable and the joint
    feedback identical to the oracle, every step."""
    from test_oracle import compare_feedback
    monkeypatch.setenv("ODEB_SOLVER", solver)
    sc = scenes.compound(5)
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    a.enable_feedback()
    b.enable_feedback()
    for s in range(120):
        a.step(0.02)
        b.step(0.02)
        bad = compare_step(a, b, sc.nworlds) + compare_feedback(a, b, sc.nworlds, True, 0)
        assert not bad, (solver, s, bad[:4])
    b.close()


@pytest.mark.parametrize("prec", PRECS)
def test_edge_cases_gpu(prec):
    """Edge cases of tests/test_oracle.py::test_edge_cases on the CUDA path: a single body, no pairs at all (free fall),
    category bits that filter everything, zero gravity; plus ragged and maximum sizes: world counts that do not fill a warp
    group (1, 3, 17, 33), dCollide contact limits 1..8, and body counts at the limits of the solver kernels' packed
    indices (62 bodies: last size k_solve5 takes; 63 / 70: k_solve only)."""
    cases = []
    cases.append((scenes.free_boxes(1, 1, grid=1), 0.01, 20))
    sc = scenes.free_boxes(1, 4, grid=2)
    sc.state["pos"][..., 2] += 5.0
    cases.append((sc, 0.01, 20))
    sc = scenes.free_boxes(2, 4, grid=2)
    for g in sc.geoms:
        g.collide_bits = 0
        g.category_bits = 0
    cases.append((sc, 0.01, 20))
    sc = scenes.box_stack(nworlds=1, nboxes=3)
    sc.wp.gravity[2] = 0.0
    cases.append((sc, 0.01, 20))
    for nw in (1, 3, 17, 33):
        cases.append((scenes.box_stack(nworlds=nw, nboxes=5), 0.02, 30))
    for mc in (1, 2, 3, 5, 8):
        sc = scenes.box_stack(nworlds=2, nboxes=4, demo_world_options=False)
        sc.wp.max_contacts = mc
        cases.append((sc, 0.02, 30))
    for nb in (62, 63, 70):
        cases.append((scenes.free_boxes(2, nb, grid=9), 0.01, 12))
    for sc, h, n in cases:
        a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
        for s in range(n):
            a.step(h)
            b.step(h)
            bad = compare_step(a, b, sc.nworlds)
            assert not bad, (sc.nworlds, sc.nbody, s, bad[:4])
            st = b.get_state()
            assert all(np.isfinite(st[k]).all() for k in st)
        b.close()


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("solver", ("v4", "p4"))
def test_slider_joints_bit_exact(prec, solver, monkeypatch):
    """Slider joints (slider.cpp:115-246, linear addLimot joint.cpp:596-780 with stops, bounce and a motor at its stop) on the
    GPU: every observable and the joint feedback identical to the oracle, every step."""
    from test_oracle import compare_feedback
    monkeypatch.setenv("ODEB_SOLVER", solver)
    sc = scenes.sliders(5)
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    a.enable_feedback()
    b.enable_feedback()
    for s in range(150):
        a.step(0.02)
        b.step(0.02)
        bad = compare_step(a, b, sc.nworlds) + compare_feedback(a, b, sc.nworlds, True, 0)
        assert not bad, (solver, s, bad[:4])
    b.close()


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("solver", ("v4", "p4"))
def test_hinge2_joints(prec, solver, monkeypatch):
    """Hinge2 joints (hinge2.cpp:110-209) on a demo_buggy-style vehicle. Without stops: every observable and the joint feedback
    identical to the oracle, every step. With the demo's steering stops (measureAngle1 -> atan2, CUDA libm vs glibc): integer
    observables identical, state within the stated tolerance per teacher-forced step."""
    from test_oracle import compare_feedback
    monkeypatch.setenv("ODEB_SOLVER", solver)
    sc = scenes.buggy(5, stops=False)
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    a.enable_feedback()
    b.enable_feedback()
    for s in range(150):
        a.step(0.05)
        b.step(0.05)
        bad = compare_step(a, b, sc.nworlds) + compare_feedback(a, b, sc.nworlds, True, 0)
        assert not bad, (solver, s, bad[:4])
    b.close()
    sc = scenes.buggy(3, stops=True)
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    for s in range(100):
        if prec == "double":
            b.set_state(**a.get_state())        # double: teacher-forced, the CUDA step starts from the oracle's state (CUDA's atan2); single: free-running, exact
        a.step(0.05)
        b.step(0.05)
        bad = compare_step(a, b, sc.nworlds, exact_float=prec == "single", tol=TOL[prec], what=("pairs", "islands", "state"))
        assert not bad, (solver, s, bad[:4])
    b.close()


@pytest.mark.parametrize("prec", PRECS)
def test_solve5_many_bodies_and_rows(prec, monkeypatch):
    """k_solve5's packed schedule entries (12-bit row, 8-bit body slots): a 70-box stack in ONE island (~850 rows, beyond k_solve's
    own row budget -> the case the host switches kernels for) and 100 separate bodies per world, against the oracle, bit-exact."""
    monkeypatch.setenv("ODEB_SOLVER", "p4")
    for mk, h, n in ((lambda: scenes.box_stack(nworlds=3, nboxes=70, demo_world_options=False), 0.02, 40),
                     (lambda: scenes.free_boxes(2, 100, grid=10), 0.01, 20)):
        sc = mk()
        a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
        rows = 0
        for s in range(n):
            a.step(h)
            b.step(h)
            bad = compare_step(a, b, sc.nworlds)
            assert not bad, (sc.nbody, s, bad[:4])
            rows = max(rows, int(b.get_totals()[2]) // sc.nworlds)
        b.close()
    monkeypatch.delenv("ODEB_SOLVER")
    # automatic choice: islands beyond k_solve's budget make the host pick k_solve5<4> after the first call
    sc = scenes.box_stack(nworlds=3, nboxes=70, demo_world_options=False)
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    for s in range(30):
        a.step(0.02)
        b.step(0.02)
        bad = compare_step(a, b, sc.nworlds)
        assert not bad, (s, bad[:4])
    b.close()


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("solver", ("v4", "p4"))
def test_rolling_friction(prec, solver, monkeypatch):
    """Rolling / spinning friction rows (contact.cpp:299-343, up to 6 rows per contact, friction index up to 5 rows back) on the GPU.
    Sphere contacts are libm-free: bit-exact; the box in the scene goes through cullPoints (atan2), so the comparison is teacher-forced
    with the stated tolerance and exact integer observables."""
    monkeypatch.setenv("ODEB_SOLVER", solver)
    for axis_dep in (False, True):
        sc = scenes.rolling(5, axis_dep=axis_dep)
        a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
        for s in range(120):
            if prec == "double":
                b.set_state(**a.get_state())    # double: teacher-forced with the stated tolerance; single: free-running, exact
            a.step(0.01)
            b.step(0.01)
            bad = compare_step(a, b, sc.nworlds, exact_float=prec == "single", tol=TOL[prec], what=("pairs", "contacts", "islands", "stats", "seeds", "state"))
            assert not bad, (solver, axis_dep, s, bad[:4])
        b.close()


@pytest.mark.parametrize("prec", PRECS)
def test_kinematic_bodies(prec):
    """Kinematic platform carrying boxes (box-box contacts go through cullPoints / atan2: teacher-forced, stated tolerance)."""
    sc = scenes.conveyor(5)
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    for s in range(120):
        if prec == "double":
            b.set_state(**a.get_state())        # double: teacher-forced with the stated tolerance; single: free-running, exact
        a.step(0.01)
        b.step(0.01)
        bad = compare_step(a, b, sc.nworlds, exact_float=prec == "single", tol=TOL[prec], what=("pairs", "contacts", "islands", "stats", "seeds", "state"))
        assert not bad, (s, bad[:4])
    b.close()


@pytest.mark.parametrize("prec", PRECS)
def test_geom_offsets(prec):
    """Composite bodies (geoms at offset poses on one body): spheres and capsules only, so nothing on the path calls libm: every
    observable identical to the oracle, every step."""
    sc = scenes.composite(5)
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    for s in range(200):
        a.step(0.01)
        b.step(0.01)
        bad = compare_step(a, b, sc.nworlds)
        assert not bad, (s, bad[:4])
    b.close()


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("space,levels", [(B.SPACE_HASH, None), (B.SPACE_HASH, (-1, 0)), (B.SPACE_HASH, (0, 4)), (B.SPACE_SIMPLE, None), (B.SPACE_SAP, None)])
def test_broadphase_spaces(prec, space, levels):
    """Callback stream of each reference space as a set, around the origin where the hash space's wrapped cell addresses drop pairs
    (oracle pinned to the compiled reference in test_oracle.py::test_broadphase_callback_stream_vs_reference): batch path and, for one
    world, the large-world sort + sweep path.  Boxes are in the scene, so floats are compared to tolerance and the sets exactly."""
    sc = scenes.scatter(24, space_type=space, levels=levels)
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    for s in range(12):
        a.step(0.02)
        b.step(0.02)
        bad = compare_step(a, b, sc.nworlds, exact_float=prec == "single", tol=TOL[prec], what=("pairs", "contacts", "islands", "state"))
        assert not bad, (s, bad[:4])
    b.close()
    sc = scenes.scatter(1, n=120, space_type=space, levels=levels, extent=1.5)
    a, b = _canon_pair(prec, sc)
    for s in range(6):
        a.step(0.02)
        b.step(0.02)
        bad = compare_step(a, b, 1, exact_float=prec == "single", tol=TOL[prec], what=("pairs", "contacts", "islands", "state"))
        assert not bad, (s, bad[:4])
    b.close()


@pytest.mark.parametrize("prec", PRECS)
def test_motor_joints(prec):
    """LMotor / AMotor rows (user + Euler mode, stops, powered at a stop, reversed attachment).  Euler angles go through atan2 (CUDA
    libm vs glibc), so teacher-forced single steps from the oracle's state: sets exact, floats to 10x the usual per-step tolerance
    (the motor-driven bodies of this scene spin at |w| ~ 10).  Dynamic iteration adjustment is switched off here (exactly 20 sweeps):
    with it a 1-ulp difference in a limit error can move one island's exit decision by a sweep, i.e. by ~1e-3 in a velocity, which is
    the reference's own sensitivity and says nothing about the motor rows; the other scenes cover the adjustment logic."""
    sc = scenes.motors(4, dynamic_iterations=False)
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    tol = dict(contact=TOL[prec]["contact"], state=10 * TOL[prec]["state"])
    for s in range(120):
        st = a.get_state()
        b.set_state(**st)
        b.set_seeds(a.get_seeds())
        a.set_state(**st)
        a.step(0.01)
        b.step(0.01)
        bad = compare_step(a, b, sc.nworlds, exact_float=prec == "single", tol=tol, what=("pairs", "contacts", "islands", "stats", "seeds", "state"))
        assert not bad, (s, bad[:4])
    b.close()


@pytest.mark.parametrize("prec", PRECS)
def test_ray_and_cylinder_colliders(prec):
    """Rays (sensor hits through odeb_get_ray_hits, no joints) and cylinders (plane / sphere colliders, AABB) on the GPU against the
    oracle.  No libm transcendental on these colliders, but the scene also holds box-box pairs (cullPoints' atan2), so trajectories are
    compared teacher-forced to the usual tolerance; pair sets, contact and hit lists (which geoms, how many) exactly, and the hit
    geometry of the step (computed from identical poses) bit for bit."""
    sc = scenes.sensors(8)
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    nhits = 0
    for s in range(100):
        st = a.get_state()
        b.set_state(**st)
        b.set_seeds(a.get_seeds())
        a.set_state(**st)
        a.step(0.01)
        b.step(0.01)
        bad = compare_step(a, b, sc.nworlds, exact_float=prec == "single", tol=TOL[prec], what=("pairs", "contacts", "islands", "stats", "seeds", "state"))
        for w in range(sc.nworlds):
            (ga, ia), (gb, ib) = a.get_ray_hits(w), b.get_ray_hits(w)
            if not np.array_equal(ia, ib) or not np.array_equal(ga, gb):
                bad.append("world %d: ray hits differ (%d vs %d)" % (w, len(ia), len(ib)))
            nhits += len(ia)
        assert not bad, (s, bad[:4])
    assert nhits > 3000
    # bulk read-out: nearest hit per ray and world, against the oracle's hit lists
    rng, hit = b.get_ray_ranges()
    rays = [i for i, g in enumerate(sc.geoms) if g.type == B.RAY]
    assert rng.shape == (sc.nworlds, len(rays)) and np.isfinite(rng).sum() > 20
    for w in range(sc.nworlds):
        ga, ia = a.get_ray_hits(w)
        for k, g in enumerate(rays):
            sel = [j for j in range(len(ia)) if g in ia[j]]
            if not sel:
                assert np.isinf(rng[w, k]) and hit[w, k] == -1
            else:
                j = min(sel, key=lambda j: ga[j, 6])
                assert rng[w, k] == ga[j, 6] and hit[w, k] == (ia[j][0] if ia[j][1] == g else ia[j][1])
    b.close()
    sc = scenes.sensors(1, n=60, extent=1.6)          # large-world path: same colliders behind the sort + sweep broadphase
    a, b = _canon_pair(prec, sc)
    for s in range(10):
        a.step(0.01)
        b.step(0.01)
        bad = compare_step(a, b, 1, exact_float=prec == "single", tol=TOL[prec], what=("pairs", "contacts", "islands", "state"))
        (ga, ia), (gb, ib) = a.get_ray_hits(0), b.get_ray_hits(0)
        assert not bad and np.array_equal(ia, ib), (s, bad[:4])
    b.close()


@pytest.mark.parametrize("prec", PRECS)
def test_cylinder_box_collider(prec):
    """collision_cylinder_box.cpp on the GPU: discs, rods and cylinders among boxes (all 40 candidate axes, both clipping routines).  No
    libm call at run time on this collider, but box-box pairs (cullPoints' atan2) share the scene: teacher-forced steps, sets exact,
    contact geometry of the cylinder-box pairs bit for bit."""
    sc = scenes.cylinders_and_boxes(16)
    types = [g.type for g in sc.geoms]
    a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
    ncb = 0
    for s in range(50):
        st = a.get_state()
        b.set_state(**st)
        b.set_seeds(a.get_seeds())
        a.set_state(**st)
        a.step(0.01)
        b.step(0.01)
        bad = compare_step(a, b, sc.nworlds, exact_float=prec == "single", tol=TOL[prec], what=("pairs", "contacts", "islands", "state"))
        assert not bad, (s, bad[:4])
        for w in range(sc.nworlds):
            (ga, ia), (gb, ib) = a.get_contacts(w), b.get_contacts(w)
            sel = np.array([{types[p[0]], types[p[1]]} == {1, 3} for p in ia], bool)
            assert np.array_equal(ga[sel], gb[sel]), (s, w)
            ncb += int(sel.sum())
    assert ncb > 2000
    b.close()


def test_page_locked_host_arrays_take_the_direct_path():
    """odeb_alloc_host arrays (transfers straight to / from the caller's memory) and pageable numpy arrays (staged) carry the same bytes"""
    sc = scenes.box_stack(nworlds=5, nboxes=6)
    a, b = B.Batch(gpu_lib("single"), sc), B.Batch(gpu_lib("single"), sc)
    f_pin = a.alloc_host((sc.nworlds, sc.nbody, 3))
    f_pin[:] = 0
    f_pin[:, -1, 0] = 0.3
    f_pag = np.array(f_pin)
    st_pin = a.alloc_state()
    for s in range(20):
        a.add_force(force=f_pin, torque=f_pin)
        b.add_force(force=f_pag, torque=f_pag)
        a.step(0.02)
        b.step(0.02)
        sa, sb = a.get_state(out=st_pin), b.get_state()
        for k in ("pos", "quat", "lvel", "avel"):
            assert np.array_equal(sa[k], sb[k]), (s, k)
    assert sa["pos"] is st_pin["pos"] and np.abs(sa["lvel"][:, -1, 0]).max() > 0
    a.close()
    b.close()


def test_device_side_gather_matches_host_state():
    """ode_b200.shard.DeviceGather: statistics + observations go from the pack kernel's device buffer through
    torch.distributed (NCCL) on a side stream, no host round trip; what arrives equals odeb_get_state / odeb_get_stats.  One rank
    here (the 2..8-rank runs are bench.py --gpus N); the batch runs on a torch stream via odeb_set_stream."""
    import torch
    import torch.distributed as dist
    from ode_b200.shard import DeviceGather
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29611")
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        sc = scenes.box_stack(nworlds=37, nboxes=8)
        b = B.Batch(gpu_lib("single"), sc)
        g = DeviceGather(b, dist)
        for it in range(4):
            b.step_async(0.02, 10)
            got = g.launch()
            b.step_async(0.02, 1)                 # the next step is queued while the gather runs on the side stream
            g.wait()
            if it == 3:
                pos, quat, lvel, avel, stats = g.unpack(got[0])
        # the gathered observation is the state after the 10-step call, one step behind the batch: replay on a second batch
        b2 = B.Batch(gpu_lib("single"), sc)
        b2.step(0.02, 43)
        st = b2.get_state()
        assert np.array_equal(pos, st["pos"]) and np.array_equal(quat, st["quat"]) and np.array_equal(lvel, st["lvel"]) and np.array_equal(avel, st["avel"])
        assert np.array_equal(stats, np.stack([b2.get_stats(w) for w in range(sc.nworlds)]))
        b.close()
        b2.close()
    finally:
        if created:
            dist.destroy_process_group()
