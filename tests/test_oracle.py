"""CPU tests (no GPU): the oracle restatement against the reference's golden vectors and, when the compiled
reference is present (oracle/_ref), against the reference itself, observable by observable and bit for bit."""
import ctypes as C
import os
import numpy as np
import pytest
from parity_util import B, ROOT, REAL, ref_lib, orc_lib, compare_step
import golden_cases as G
from ode_b200 import scenes

GOLD = os.path.join(ROOT, "tests", "golden")
PRECS = ("single", "double")


def test_rand_known_answers():
    """dTestRand's known answers (ode/src/misc.cpp:64-74) and recorded dRand/dRandInt streams."""
    lib = orc_lib("single")
    g = np.load(os.path.join(GOLD, "rand.npz"))
    fn = lib.lib.orc_rand_next
    fn.restype = C.c_ulong
    seed = C.c_uint32(0)
    stream = [fn(C.byref(seed)) for _ in range(64)]
    assert stream[:5] == [0x3c6ef35f, 0x47502932, 0xd1ccf6e9, 0xaaf95334, 0x6252e503]
    assert np.array_equal(np.array(stream, np.uint64), g["stream"])
    seed = C.c_uint32(12345)
    out = []
    for rep in range(8):
        for n in g["ns"]:
            out.append(lib.lib.orc_rand_int_(C.byref(seed), int(n)))
    assert np.array_equal(np.array(out, np.int64), g["randint"])
    assert seed.value == int(g["final_seed"])


@pytest.mark.parametrize("prec", PRECS)
def test_colliders_golden(prec):
    """dCollide for all 15 ordered pairs of {sphere, box, capsule, plane} + computeAABB: contact count exact,
    position / normal / depth bit-exact against the recorded reference outputs."""
    lib = orc_lib(prec)
    g = np.load(os.path.join(GOLD, "collide_%s.npz" % prec))
    n, g7, aabb = G.run_collide(lib, "orc_", G.collide_cases(REAL[prec]))
    assert np.array_equal(n, g["n"])
    assert (n > 0).sum() > 100 and (n > 1).sum() > 30    # the fixture really exercises multi-contact paths
    assert np.array_equal(aabb, g["aabb"])
    for i in range(len(n)):
        assert np.array_equal(g7[i][:n[i]], g["geom7"][i][:n[i]]), "case %d" % i


@pytest.mark.parametrize("prec", PRECS)
def test_contact_rows_golden(prec):
    """dxJointContact::getInfo1/getInfo2 (16-wide row layout), incl. the reference's own test vectors."""
    lib = orc_lib(prec)
    cases = G.contact_row_cases(REAL[prec])
    m, rows, fi = G.run_contact_rows(lib, "orc_", cases)
    # tests/friction.cpp:119-136 and :150-170 (expected J rows written out in the reference's test)
    assert m[0] == 2 and m[1] == 2
    np.testing.assert_allclose(rows[0][1][[0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13]], [0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 1, 0], atol=1e-6)
    np.testing.assert_allclose(rows[1][1][[0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13]], [0, 1, 0, 0, 0, 1, 0, -1, 0, 0, 0, 1], atol=1e-6)
    assert fi[0][1] == 0 and fi[1][1] == 0
    g = np.load(os.path.join(GOLD, "contact_rows_%s.npz" % prec))
    assert np.array_equal(m, g["m"]) and np.array_equal(fi, g["findex"])
    assert np.array_equal(rows, g["rows"])


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", sorted(G.TRAJ_SCENES))
def test_trajectories_golden(prec, name):
    """Whole-step observables on small scenes (pair set, contacts, island labels, iteration statistics,
    dRand seed, body state) against the recorded reference run -- all bit-exact."""
    mk, h, nsteps, every = G.TRAJ_SCENES[name]
    gold = np.load(os.path.join(GOLD, "traj_%s_%s.npz" % (name, prec)))
    bad = G.compare_traj(B.Batch(orc_lib(prec), mk()), gold, h)
    assert not bad, bad


@pytest.mark.parametrize("prec", PRECS)
def test_oracle_vs_compiled_reference(prec):
    """Direct run against oracle/_ref (skipped where the reference build is absent)."""
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("oracle/_ref not built here")
    for mk, h, n in ((lambda: scenes.box_stack(nworlds=2, nboxes=10), 0.02, 60), (lambda: scenes.pile(nbodies=40), 0.01, 60),
                     (lambda: scenes.ragdoll(1, seed0=3), 0.01, 120), (lambda: scenes.chain(1), 0.05, 60)):
        sc = mk()
        a, b = B.Batch(ref, sc), B.Batch(orc_lib(prec), sc)
        for s in range(n):
            a.step(h)
            b.step(h)
            bad = compare_step(a, b, sc.nworlds)
            assert not bad, (s, bad)


@pytest.mark.parametrize("prec", PRECS)
def test_edge_cases(prec):
    """Empty contact set (bodies in free fall), a world with a single body, geoms without bodies only,
    category bits that filter everything, zero gravity."""
    lib = orc_lib(prec)
    ref = ref_lib(prec)
    cases = []
    sc = scenes.free_boxes(1, 1, grid=1)
    cases.append(sc)
    sc = scenes.free_boxes(1, 4, grid=2)
    sc.state["pos"][..., 2] += 5.0          # far above the plane: no pairs at all
    cases.append(sc)
    sc = scenes.free_boxes(1, 4, grid=2)
    for g in sc.geoms:
        g.collide_bits = 0
        g.category_bits = 0
    cases.append(sc)
    sc = scenes.box_stack(nworlds=1, nboxes=3)
    sc.wp.gravity[2] = 0.0
    cases.append(sc)
    for sc in cases:
        b = B.Batch(lib, sc)
        a = B.Batch(ref, sc) if ref is not None else None
        for s in range(20):
            b.step(0.01)
            st = b.get_state()
            assert all(np.isfinite(st[k]).all() for k in st)
            if a is not None:
                a.step(0.01)
                bad = compare_step(a, b, sc.nworlds)
                assert not bad, bad


def _feedback_scenes():
    return (("stack", lambda: scenes.box_stack(nworlds=2, nboxes=6), 0.02, 60, True),
            ("chain", lambda: scenes.chain(2), 0.05, 60, True),
            ("ragdoll", lambda: scenes.ragdoll(2), 0.01, 40, False))


def compare_feedback(a, b, nworlds, exact, tol):
    bad = []
    for w in range(nworlds):
        (fa, sa), (fb, sb) = a.get_feedback(w), b.get_feedback(w)
        if fa.shape != fb.shape or not np.array_equal(sa, sb):
            bad.append("world %d: feedback joint count / state differ (%d vs %d)" % (w, len(sa), len(sb)))
            continue
        for i in range(len(sa)):
            k = 0 if sa[i] == 0 else 6 if sa[i] == 1 else 12
            if exact and not np.array_equal(fa[i, :k], fb[i, :k]):
                bad.append("world %d joint %d: feedback differs by %.3g" % (w, i, np.abs(fa[i, :k] - fb[i, :k]).max()))
            elif not exact and k and np.abs(fa[i, :k].astype(np.float64) - fb[i, :k]).max() > tol:
                bad.append("world %d joint %d: feedback differs by %.3g > %.3g" % (w, i, np.abs(fa[i, :k].astype(np.float64) - fb[i, :k]).max(), tol))
    return bad


@pytest.mark.parametrize("prec", PRECS)
def test_joint_feedback_vs_reference(prec):
    """dJointSetFeedback on every joint (permanent joints and the step's contact joints): f1/t1/f2/t2 and which of them the
    step wrote, restatement against the compiled reference. Bit-exact where the path has no libm call."""
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    for name, mk, h, n, exact in _feedback_scenes():
        sc = mk()
        a, b = B.Batch(ref, sc), B.Batch(orc_lib(prec), sc)
        a.enable_feedback()
        b.enable_feedback()
        seen = 0
        for s in range(n):
            a.step(h)
            b.step(h)
            bad = compare_feedback(a, b, sc.nworlds, exact, 1e-3 if prec == "single" else 1e-9)
            assert not bad, (name, s, bad[:4])
            seen += int((a.get_feedback(0)[1] > 0).sum())
        assert seen > 0, name


@pytest.mark.parametrize("prec", PRECS)
def test_fixed_joints_vs_reference(prec):
    """dJointCreateFixed / dJointSetFixed (joints/fixed.cpp) between bodies and to the environment: restatement against the
    compiled reference, every observable of every step bit-identical (no libm on this path), feedback included."""
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    sc = scenes.compound(3)
    a, b = B.Batch(ref, sc), B.Batch(orc_lib(prec), sc)
    a.enable_feedback()
    b.enable_feedback()
    for s in range(120):
        a.step(0.02)
        b.step(0.02)
        bad = compare_step(a, b, sc.nworlds) + compare_feedback(a, b, sc.nworlds, True, 0)
        assert not bad, (s, bad[:4])


@pytest.mark.parametrize("prec", PRECS)
def test_slider_joints_vs_reference(prec):
    """dJointCreateSlider / dJointSetSliderAxis / dJointSetSliderParam (joints/slider.cpp) with stops, bounce, a motor driven
    into its stop (dBodyAddForce path of the linear addLimot, joint.cpp:655-704) and a reversed attachment to the environment:
    restatement against the compiled reference, bit-identical."""
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    sc = scenes.sliders(3)
    a, b = B.Batch(ref, sc), B.Batch(orc_lib(prec), sc)
    a.enable_feedback()
    b.enable_feedback()
    six = 0
    for s in range(150):
        a.step(0.02)
        b.step(0.02)
        bad = compare_step(a, b, sc.nworlds) + compare_feedback(a, b, sc.nworlds, True, 0)
        assert not bad, (s, bad[:4])
    pos = b.get_state()["pos"]
    assert np.isfinite(pos).all()


@pytest.mark.parametrize("prec", PRECS)
def test_hinge2_joints_vs_reference(prec):
    """dJointCreateHinge2 / SetHinge2Anchor / SetHinge2Axes / SetHinge2Param (joints/hinge2.cpp, setBall2 with suspension ERP/CFM,
    both motors) on a demo_buggy-style vehicle: restatement against the compiled reference. Without stops nothing calls libm:
    bit-identical; with the demo's steering stops measureAngle1 goes through atan2 on both sides (glibc): still bit-identical."""
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    for stops in (False, True):
        sc = scenes.buggy(3, stops=stops)
        a, b = B.Batch(ref, sc), B.Batch(orc_lib(prec), sc)
        a.enable_feedback()
        b.enable_feedback()
        for s in range(150):
            a.step(0.05)
            b.step(0.05)
            bad = compare_step(a, b, sc.nworlds) + compare_feedback(a, b, sc.nworlds, True, 0)
            assert not bad, (stops, s, bad[:4])
        st = b.get_state()
        assert np.isfinite(st["pos"]).all() and np.abs(st["pos"][:, 0, :2]).max() > 0.05      # the buggy drove somewhere


@pytest.mark.parametrize("prec", PRECS)
def test_rolling_friction_vs_reference(prec):
    """dContactRolling (rolling about t1 / t2, spinning about the normal; proportional limits with Approx1; the AxisDep form with a
    zero coefficient, which the reference counts as a row and leaves empty): restatement against the compiled reference."""
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    for axis_dep in (False, True):
        sc = scenes.rolling(3, axis_dep=axis_dep)
        a, b = B.Batch(ref, sc), B.Batch(orc_lib(prec), sc)
        a.enable_feedback()
        b.enable_feedback()
        for s in range(150):
            a.step(0.01)
            b.step(0.01)
            bad = compare_step(a, b, sc.nworlds, exact_float=(prec == "double" or True)) + compare_feedback(a, b, sc.nworlds, True, 0)
            assert not bad, (axis_dep, s, bad[:4])
        st = b.get_state()
        assert np.abs(st["avel"]).max() < 6.0          # rolling / spinning friction slowed the bodies down


@pytest.mark.parametrize("prec", PRECS)
def test_kinematic_bodies_vs_reference(prec):
    """dBodySetKinematic: zero inverse mass / inertia, the body keeps its velocity under contact forces and gravity."""
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    sc = scenes.conveyor(2)
    a, b = B.Batch(ref, sc), B.Batch(orc_lib(prec), sc)
    v0 = b.get_state()["lvel"][:, 0].copy()
    for s in range(150):
        a.step(0.01)
        b.step(0.01)
        bad = compare_step(a, b, sc.nworlds)
        assert not bad, (s, bad[:4])
    st = b.get_state()
    assert np.array_equal(st["lvel"][:, 0], v0) and abs(st["pos"][0, 0, 2] - 0.5) < 1e-6      # the platform did not react
    assert st["pos"][0, 1, 0] > -0.6 + 0.2                                                    # and carried its load along


@pytest.mark.parametrize("prec", PRECS)
def test_geom_offsets_vs_reference(prec):
    """dGeomSetOffsetPosition / dGeomSetOffsetQuaternion: composite bodies (several geoms per body at offset poses), pair sets, contacts
    and trajectories of the restatement against the compiled reference."""
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    sc = scenes.composite(2)
    a, b = B.Batch(ref, sc), B.Batch(orc_lib(prec), sc)
    ncont = 0
    for s in range(200):
        a.step(0.01)
        b.step(0.01)
        bad = compare_step(a, b, sc.nworlds)
        assert not bad, (s, bad[:4])
        ncont += len(b.get_contacts(0)[1])
    assert ncont > 200


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("space,levels", [(B.SPACE_HASH, None), (B.SPACE_HASH, (-1, 0)), (B.SPACE_HASH, (0, 4)), (B.SPACE_SIMPLE, None), (B.SPACE_SAP, None)])
def test_broadphase_callback_stream_vs_reference(prec, space, levels):
    """Pair sets of the three spaces around the origin, where the hash space's negative-z cell addresses wrap (collision_space.cpp:499,
    :533) and it reports fewer pairs than the simple space: the restatement's set against the compiled reference's callback stream,
    with default and dHashSpaceSetLevels levels (min level clamp, big-box list)."""
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    sc = scenes.scatter(24, space_type=space, levels=levels)
    a, b = B.Batch(ref, sc), B.Batch(orc_lib(prec), sc)
    npairs = 0
    for s in range(12):
        a.step(0.02)
        b.step(0.02)
        bad = compare_step(a, b, sc.nworlds, what=("pairs", "contacts"))
        assert not bad, (s, bad[:4])
        npairs += sum(len(b.get_pairs(w)) for w in range(sc.nworlds))
    assert npairs > 2000


@pytest.mark.parametrize("prec", PRECS)
def test_hash_space_is_a_strict_subset_of_simple_space_here(prec):
    """The scatter scene does exercise the wrap: the hash space's set is smaller than the simple space's on the same geoms."""
    counts = {}
    for space in (B.SPACE_HASH, B.SPACE_SIMPLE):
        sc = scenes.scatter(24, space_type=space)
        b = B.Batch(orc_lib(prec), sc)
        b.step(0.02)
        counts[space] = [set(map(tuple, b.get_pairs(w))) for w in range(sc.nworlds)]
    assert all(h <= s for h, s in zip(counts[B.SPACE_HASH], counts[B.SPACE_SIMPLE]))
    assert sum(len(s) - len(h) for h, s in zip(counts[B.SPACE_HASH], counts[B.SPACE_SIMPLE])) > 0


@pytest.mark.parametrize("prec", PRECS)
def test_motor_joints_vs_reference(prec):
    """LMotor (lmotor.cpp) and AMotor (amotor.cpp, user and Euler mode, stops, bounce, powered at a stop, lo == hi, dJOINT_REVERSE):
    the restatement against the compiled reference, every observable, 300 steps."""
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    sc = scenes.motors(3)
    a, b = B.Batch(ref, sc), B.Batch(orc_lib(prec), sc)
    for s in range(300):
        a.step(0.01)
        b.step(0.01)
        bad = compare_step(a, b, sc.nworlds)
        assert not bad, (s, bad[:4])


def _ray_hits_equal(a, b, nworlds):
    bad = []
    for w in range(nworlds):
        (ga, ia), (gb, ib) = a.get_ray_hits(w), b.get_ray_hits(w)
        if not np.array_equal(ia, ib):
            bad.append("world %d: ray hit pairs differ (%d vs %d)" % (w, len(ia), len(ib)))
        elif not np.array_equal(ga, gb):
            bad.append("world %d: ray hit geometry differs by %.3g" % (w, np.abs(ga - gb).max()))
    return bad


@pytest.mark.parametrize("prec", PRECS)
def test_ray_and_cylinder_colliders_vs_reference(prec):
    """dCollideRaySphere / RayBox / RayCapsule / RayPlane / RayCylinder (ray.cpp), dCollideCylinderPlane, dCollideCylinderSphere and the
    ray / cylinder AABBs: the restatement against the compiled reference on drifting, tumbling bodies (rays at offset poses), every
    observable bit for bit incl. the ray hits, 150 steps x 8 worlds."""
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    sc = scenes.sensors(8)
    a, b = B.Batch(ref, sc), B.Batch(orc_lib(prec), sc)
    nhits = 0
    for s in range(150):
        a.step(0.01)
        b.step(0.01)
        bad = compare_step(a, b, sc.nworlds) + _ray_hits_equal(a, b, sc.nworlds)
        assert not bad, (s, bad[:4])
        nhits += sum(len(b.get_ray_hits(w)[1]) for w in range(sc.nworlds))
    assert nhits > 5000


@pytest.mark.parametrize("prec", PRECS)
def test_cylinder_box_collider_vs_reference(prec):
    """dCollideCylinderBox (collision_cylinder_box.cpp): the restatement, incl. the cap octagon's constant normals and dMatrix3Inv's double
    reciprocal, against the compiled reference on discs / rods / cylinders tumbling among boxes: every observable bit for bit."""
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    sc = scenes.cylinders_and_boxes(16)
    types = [g.type for g in sc.geoms]
    a, b = B.Batch(ref, sc), B.Batch(orc_lib(prec), sc)
    ncb = 0
    for s in range(60):
        a.step(0.01)
        b.step(0.01)
        bad = compare_step(a, b, sc.nworlds)
        assert not bad, (s, bad[:4])
        ncb += sum(1 for w in range(sc.nworlds) for p in b.get_contacts(w)[1] if {types[p[0]], types[p[1]]} == {1, 3})
    assert ncb > 3000


# ---------------------------------------------------------------------------------------------------------------
# canonical mode of the oracle (ODEB_MODE_CANONICAL, include/ode_b200.h): the large-world path's CPU restatement
@pytest.mark.parametrize("prec", PRECS)
def test_canonical_mode_of_the_oracle_vs_compiled_reference(prec):
    """The canonical order only changes the order in which rows are relaxed.  Teacher-forced from the compiled reference's state: pair set,
    per-pair contact counts, contact geometry and island partition of a step are the reference's, bit for bit; the body state agrees with
    the reference to the accuracy of a converged SOR solve (order effects only); two runs of the mode are identical."""
    ref = ref_lib(prec)
    if ref is None:
        pytest.skip("oracle/_ref not built here")
    for mk, h, targets in ((lambda: scenes.wall(10, 6), 0.05, (0, 8, 20)), (lambda: scenes.pile(nbodies=64), 0.01, (0, 40, 80)),
                           (lambda: scenes.chain(1), 0.05, (0, 30))):
        sc = mk()
        a = B.Batch(ref, sc)
        b, c = B.Batch(orc_lib(prec), sc), B.Batch(orc_lib(prec), sc)
        b.set_solver_mode(1)
        c.set_solver_mode(1)
        done = 0
        for target in targets:
            a.step(h, target - done)
            done = target
            st = a.get_state()
            for x in (a, b, c):
                x.set_state(**st)
            b.set_seeds(a.get_seeds())
            c.set_seeds(a.get_seeds())
            a.step(h)
            b.step(h)
            c.step(h)
            done += 1
            bad = compare_step(a, b, 1, what=("pairs", "contacts", "islands"))
            assert not bad, (sc.nbody, target, bad)
            bad = compare_step(b, c, 1)
            assert not bad, (sc.nbody, target, bad)
            sa, sb = a.get_state(), b.get_state()
            for k in ("pos", "lvel"):
                assert np.abs(sa[k].astype(np.float64) - sb[k]).max() < 0.05, (sc.nbody, target, k)


def test_canonical_colour_order_helpers():
    """odeb_canon_key through the C-ABI of the product library == the header's inline function the oracle uses (spot values), and the
    per-phase colour ranks are a permutation that changes with the phase."""
    lib = C.CDLL(os.path.join(ROOT, "ode_b200", "libode_b200_single.so"))
    lib.odeb_canon_key.restype = C.c_uint32
    lib.odeb_canon_key.argtypes = [C.c_uint32] * 4

    def key(seed, island, phase, row):
        x = (seed ^ (island * 0x9E3779B9) ^ (phase * 0x85EBCA6B) ^ (row * 0xC2B2AE35)) & 0xffffffff
        x ^= x >> 16
        x = (x * 0x85EBCA6B) & 0xffffffff
        x ^= x >> 13
        x = (x * 0xC2B2AE35) & 0xffffffff
        x ^= x >> 16
        return x
    for args in ((1, 0, 0, 0), (12345, 3, 2, 77), (0xdeadbeef, 0xffffffff, 4, 255)):
        assert lib.odeb_canon_key(*args) == key(*args)
    orders = []
    for phase in range(5):
        ks = sorted((key(7, 0xffffffff, phase, c), c) for c in range(10))
        orders.append(tuple(c for _, c in ks))
        assert sorted(orders[-1]) == list(range(10))
    assert len(set(orders)) > 1
