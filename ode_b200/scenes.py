"""Synthetic scenes of BASELINE.json's configs, as Scene descriptions.

Scene recipes follow the reference's demos (ode/demo/demo_boxstack.cpp, demo_chain2.cpp,
demo_crash.cpp); the ragdoll is authored here because the reference has none.
"""
import numpy as np
INF_ = float("inf")
from . import _binding as B


def _rng(seed):
    return np.random.RandomState(seed)


def box_stack(nworlds=1, nboxes=16, seed0=1000, jitter=1e-3, demo_world_options=True):
    """Config 2: nboxes boxes of side 0.5, density 5, stacked on the plane z=0.

    Surface parameters of demo_boxstack.cpp:144-151 (Bounce|SoftCFM, mu=inf, bounce 0.1,
    bounce_vel 0.1, soft_cfm 0.01), <= 8 contacts per pair, g = (0,0,-0.5), CFM 1e-5; world options of
    demo_boxstack.cpp:582-597 (auto-disable, damping, max angular speed, correcting vel, surface layer).
    """
    kw = dict(gravity=(0, 0, -0.5), cfm=1e-5, max_contacts=8,
              surf_mode=B.CONTACT_BOUNCE | B.CONTACT_SOFT_CFM, mu=B.INF, mu2=0.0,
              bounce=0.1, bounce_vel=0.1, soft_cfm=0.01, space_type=B.SPACE_HASH)
    if demo_world_options:
        kw.update(auto_disable=1, adis_samples=10, linear_damping=1e-5, angular_damping=0.005,
                  max_angular_speed=200.0, contact_max_vel=0.1, contact_surface_layer=0.001)
    sc = B.Scene(B.default_world_params(**kw), nworlds)
    sc.add_geom(B.PLANE, (0, 0, 1, 0))
    side = 0.5
    m, I = B.box_mass(5.0, side, side, side)
    for i in range(nboxes):
        pos = (0.01 * (i % 3), 0.005 * (i % 2), 0.25 + 0.5 * i + 0.001 * i)
        b = sc.add_body(m, I, pos)
        sc.add_geom(B.BOX, (side, side, side), body=b)
    pos = np.tile(np.asarray(sc.body_pos)[None], (nworlds, 1, 1))
    for w in range(nworlds):
        r = _rng(seed0 + w)
        pos[w, :, :2] += jitter * (r.rand(nboxes, 2) - 0.5)
    quat = np.tile(np.array([1.0, 0, 0, 0])[None, None], (nworlds, nboxes, 1))
    sc.state = dict(pos=pos, quat=quat, lvel=np.zeros_like(pos), avel=np.zeros_like(pos))
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def pile(nworlds=1, nbodies=1000, seed=12345, spacing=1.1, space_type=B.SPACE_HASH, max_contacts=4, vary=0.0):
    """Config 1: alternating unit boxes / r=0.5 spheres on a cubic lattice dropped on the plane z=0.

    dContactApprox1, mu=0.5, <= 4 contacts per pair, g=(0,0,-9.81), defaults otherwise.
    vary > 0: every world gets its own horizontal jitter of that size on top of the lattice (worlds then differ in island
    structure and contact counts, not only in their dRand seed).
    """
    sc = B.Scene(B.default_world_params(gravity=(0, 0, -9.81), max_contacts=max_contacts,
                                        surf_mode=B.CONTACT_APPROX1, mu=0.5, space_type=space_type), nworlds)
    sc.add_geom(B.PLANE, (0, 0, 1, 0))
    n = int(np.ceil(nbodies ** (1.0 / 3.0) - 1e-9))
    r = _rng(seed)
    mb, Ib = B.box_mass(1.0, 1, 1, 1)
    ms, Is = B.sphere_mass(1.0, 0.5)
    k = 0
    for iz in range(n):
        for iy in range(n):
            for ix in range(n):
                if k >= nbodies:
                    break
                jit = 0.01 * r.rand(2)
                pos = (ix * spacing + jit[0], iy * spacing + jit[1], 0.6 + iz * spacing)
                if k % 2 == 0:
                    b = sc.add_body(mb, Ib, pos)
                    sc.add_geom(B.BOX, (1, 1, 1), body=b)
                else:
                    b = sc.add_body(ms, Is, pos)
                    sc.add_geom(B.SPHERE, (0.5,), body=b)
                k += 1
    if vary > 0:
        pos = np.tile(np.asarray(sc.body_pos)[None], (nworlds, 1, 1))
        for w in range(nworlds):
            pos[w, :, :2] += vary * (_rng(seed + 7919 * (w + 1)).rand(sc.nbody, 2) - 0.5)
        quat = np.tile(np.array([1.0, 0, 0, 0])[None, None], (nworlds, sc.nbody, 1))
        sc.state = dict(pos=pos, quat=quat, lvel=np.zeros_like(pos), avel=np.zeros_like(pos))
    sc.seeds = (seed + np.arange(nworlds)).astype(np.uint32)
    return sc


def chain(nworlds=1, nlinks=10, seed0=7):
    """Config 3: demo_chain2.cpp:138-160 -- boxes of side 0.2, mass 1, ball joints between neighbours,
    plane, g=(0,0,-0.5), CFM 1e-5, contacts: 1 per pair, mode 0, mu=inf (demo_chain2.cpp:65-82)."""
    sc = B.Scene(B.default_world_params(gravity=(0, 0, -0.5), cfm=1e-5, max_contacts=1, surf_mode=0,
                                        mu=B.INF, skip_connected=0), nworlds)
    sc.add_geom(B.PLANE, (0, 0, 1, 0))
    side = 0.2
    _, I = B.box_mass(1.0, side, side, side)
    I = I / (side ** 3)  # dMassAdjust(m, MASS=1)
    for i in range(nlinks):
        k = i * side
        b = sc.add_body(1.0, I, (k, k, k + 0.4))
        sc.add_geom(B.BOX, (side, side, side), body=b)
    for i in range(nlinks - 1):
        k = (i + 0.5) * side
        sc.add_joint(B.JOINT_BALL, i, i + 1, (k, k, k + 0.4))
    pos = np.tile(np.asarray(sc.body_pos)[None], (nworlds, 1, 1))
    lvel = np.zeros_like(pos)
    for w in range(nworlds):
        r = _rng(seed0 + w)
        lvel[w] = 0.05 * (r.rand(nlinks, 3) - 0.5)
    quat = np.tile(np.array([1.0, 0, 0, 0])[None, None], (nworlds, nlinks, 1))
    sc.state = dict(pos=pos, quat=quat, lvel=lvel, avel=np.zeros_like(pos))
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def compound(nworlds=1, seed0=31):
    """Fixed joints (dJointCreateFixed + dJointSetFixed): per world two L-shaped compounds of three boxes welded together
    and one box welded to the environment, dropped on / resting over a plane, plus a ball joint between the compounds.
    Contacts: <= 4 per pair, mode 0, mu = inf (no libm on the path: bit-exact scenes)."""
    sc = B.Scene(B.default_world_params(gravity=(0, 0, -2.0), cfm=1e-5, max_contacts=4, surf_mode=0, mu=B.INF, skip_connected=1), nworlds)
    sc.add_geom(B.PLANE, (0, 0, 1, 0))
    m, I = B.box_mass(2.0, 0.3, 0.2, 0.2)
    centres = [(0.0, 0.0, 0.6), (0.3, 0.0, 0.6), (0.3, 0.0, 0.8), (1.2, 0.1, 0.5), (1.2, 0.4, 0.5), (1.5, 0.4, 0.5), (0.7, -0.8, 1.4)]
    for c in centres:
        b = sc.add_body(m, I, c)
        sc.add_geom(B.BOX, (0.3, 0.2, 0.2), body=b)
    for a, b in ((0, 1), (1, 2), (3, 4), (4, 5)):
        sc.add_joint(B.JOINT_FIXED, a, b, (0, 0, 0))
    sc.add_joint(B.JOINT_FIXED, -1, 6, (0, 0, 0))          # welded to the environment, attached with the bodies reversed
    sc.add_joint(B.JOINT_BALL, 2, 3, (0.75, 0.05, 0.65))
    nb = sc.nbody
    pos = np.tile(np.asarray(sc.body_pos)[None], (nworlds, 1, 1))
    quat = np.tile(np.array([1.0, 0, 0, 0])[None, None], (nworlds, nb, 1))
    lvel = np.zeros((nworlds, nb, 3))
    avel = np.zeros((nworlds, nb, 3))
    for w in range(nworlds):
        r = _rng(seed0 + w)
        lvel[w] = 0.2 * (r.rand(nb, 3) - 0.5)
        avel[w] = 0.3 * (r.rand(nb, 3) - 0.5)
    sc.state = dict(pos=pos, quat=quat, lvel=lvel, avel=avel)
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def sliders(nworlds=1, seed0=41):
    """Slider joints (joints/slider.cpp): a vertical piston between two boxes with stops and bounce, a horizontal slider to the
    environment (bodies reversed) driven by a motor into its stop, and a free slider between two boxes lying on the plane.
    Contacts <= 4 per pair, mode 0, mu = inf; nothing on the path calls libm."""
    sc = B.Scene(B.default_world_params(gravity=(0, 0, -3.0), cfm=1e-5, max_contacts=4, surf_mode=0, mu=B.INF, skip_connected=1), nworlds)
    sc.add_geom(B.PLANE, (0, 0, 1, 0))
    m, I = B.box_mass(2.0, 0.3, 0.3, 0.2)
    centres = [(0.0, 0.0, 0.1), (0.0, 0.0, 0.6), (1.5, 0.0, 0.8), (3.0, 0.0, 0.1), (3.6, 0.1, 0.1)]
    for c in centres:
        b = sc.add_body(m, I, c)
        sc.add_geom(B.BOX, (0.3, 0.3, 0.2), body=b)
    sc.add_joint(B.JOINT_SLIDER, 1, 0, (0, 0, 0), axis1=(0, 0, 1), lo_stop=(-0.15, -INF_), hi_stop=(0.25, INF_), bounce=(0.3, -1))
    sc.add_joint(B.JOINT_SLIDER, -1, 2, (0, 0, 0), axis1=(1, 0, 0), lo_stop=(-0.2, -INF_), hi_stop=(0.3, INF_), vel=(0.8, 0), fmax=(6.0, 0),
                 fudge_factor=(0.5, -1))
    sc.add_joint(B.JOINT_SLIDER, 3, 4, (0, 0, 0), axis1=(1, 0.2, 0))
    nb = sc.nbody
    pos = np.tile(np.asarray(sc.body_pos)[None], (nworlds, 1, 1))
    quat = np.tile(np.array([1.0, 0, 0, 0])[None, None], (nworlds, nb, 1))
    lvel = np.zeros((nworlds, nb, 3))
    avel = np.zeros((nworlds, nb, 3))
    for w in range(nworlds):
        r = _rng(seed0 + w)
        lvel[w] = 0.6 * (r.rand(nb, 3) - 0.5)
        avel[w] = 0.2 * (r.rand(nb, 3) - 0.5)
    sc.state = dict(pos=pos, quat=quat, lvel=lvel, avel=avel)
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def buggy(nworlds=1, seed0=51, stops=False):
    """demo_buggy.cpp:200-260: a box chassis on three sphere wheels held by hinge2 joints (axis 1 = steering / suspension (0,0,1),
    axis 2 = wheel axle (0,1,0), suspension ERP 0.4 / CFM 0.8), driven by the axis-2 motor of the front wheel and steered by the
    axis-1 motor; ground contacts as demo_buggy.cpp:96-108 (Slip1|Slip2|SoftERP|SoftCFM|Approx1, mu = inf).
    stops=False: no stops on axis 1 (nothing on the path calls libm: bit-exact scenes);
    stops=True: steering stops +-0.75 on the front wheel and the rear wheels locked with lo = hi = 0 as in the demo
    (measureAngle1 -> atan2: tolerance scenes)."""
    mode = B.CONTACT_SLIP1 | B.CONTACT_SLIP2 | B.CONTACT_SOFT_ERP | B.CONTACT_SOFT_CFM | B.CONTACT_APPROX1
    sc = B.Scene(B.default_world_params(gravity=(0, 0, -0.5), max_contacts=8, surf_mode=mode, mu=B.INF, slip1=0.1, slip2=0.1,
                                        soft_erp=0.5, soft_cfm=0.3, skip_connected=1), nworlds)
    sc.add_geom(B.PLANE, (0, 0, 1, 0), category=1, collide=2)
    L, W, H, R, Z = 0.7, 0.5, 0.2, 0.18, 0.5
    m, I = B.box_mass(1.0, L, W, H)
    I = I / m                                         # dMassAdjust(&m, CMASS = 1)
    chassis = sc.add_body(1.0, I, (0, 0, Z))
    sc.add_geom(B.BOX, (L, W, H), body=chassis, category=2, collide=1)
    ms, Is = B.sphere_mass(1.0, R)
    Is = Is * (0.2 / ms)                              # WMASS = 0.2
    s = np.sqrt(0.5)
    wheels = []
    for k, (x, y) in enumerate(((0.5 * L, 0.0), (-0.5 * L, 0.5 * W), (-0.5 * L, -0.5 * W))):
        b = sc.add_body(0.2, Is, (x, y, Z - 0.5 * H), (s, s, 0.0, 0.0))       # dQFromAxisAndAngle(q, 1,0,0, pi/2)
        sc.add_geom(B.SPHERE, (R,), body=b, category=2, collide=1)
        wheels.append(b)
        lo, hi = (-INF_, -INF_), (INF_, INF_)
        if stops:
            lo, hi = ((-0.75, -INF_), (0.75, INF_)) if k == 0 else ((0.0, -INF_), (0.0, INF_))
        vel, fmax = ((0.3, -1.5), (0.2, 0.1)) if k == 0 else ((0.0, 0.0), (0.0, 0.0))
        sc.add_joint(B.JOINT_HINGE2, chassis, b, (x, y, Z - 0.5 * H), axis1=(0, 0, 1), axis2=(0, 1, 0), lo_stop=lo, hi_stop=hi,
                     vel=vel, fmax=fmax, susp_erp=0.4, susp_cfm=0.8)
    nb = sc.nbody
    pos = np.tile(np.asarray(sc.body_pos)[None], (nworlds, 1, 1))
    quat = np.tile(np.asarray(sc.body_quat)[None], (nworlds, 1, 1))
    lvel = np.zeros((nworlds, nb, 3))
    avel = np.zeros((nworlds, nb, 3))
    for w in range(nworlds):
        r = _rng(seed0 + w)
        lvel[w] = 0.1 * (r.rand(nb, 3) - 0.5)
    sc.state = dict(pos=pos, quat=quat, lvel=lvel, avel=avel)
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def rolling(nworlds=1, seed0=61, axis_dep=False):
    """Rolling / spinning friction (dContactRolling, contact.cpp:73-116, :299-343): spheres and a box sliding, rolling and spinning
    on a plane and against each other. axis_dep=False: one rho for all three axes, Approx1 (limits proportional to the normal force);
    axis_dep=True: dContactAxisDep (= Mu2) with rho, rho2 = 0 (the reference counts a row for it and leaves it empty) and rhoN."""
    if axis_dep:
        mode = B.CONTACT_ROLLING | B.CONTACT_MU2
        wp = B.default_world_params(gravity=(0, 0, -9.81), max_contacts=4, surf_mode=mode, mu=0.8, mu2=0.4, rho=0.02, rho2=0.0, rhoN=0.1)
    else:
        mode = B.CONTACT_ROLLING | B.CONTACT_APPROX1
        wp = B.default_world_params(gravity=(0, 0, -9.81), max_contacts=4, surf_mode=mode, mu=1.0, rho=0.05)
    sc = B.Scene(wp, nworlds)
    sc.add_geom(B.PLANE, (0, 0, 1, 0))
    for k in range(5):
        ms, Is = B.sphere_mass(1.0, 0.2)
        b = sc.add_body(ms, Is, (0.45 * k, 0.05 * k, 0.2))
        sc.add_geom(B.SPHERE, (0.2,), body=b)
    m, I = B.box_mass(1.0, 0.4, 0.4, 0.2)
    b = sc.add_body(m, I, (0.9, 1.0, 0.1))
    sc.add_geom(B.BOX, (0.4, 0.4, 0.2), body=b)
    nb = sc.nbody
    pos = np.tile(np.asarray(sc.body_pos)[None], (nworlds, 1, 1))
    quat = np.tile(np.array([1.0, 0, 0, 0])[None, None], (nworlds, nb, 1))
    lvel = np.zeros((nworlds, nb, 3))
    avel = np.zeros((nworlds, nb, 3))
    for w in range(nworlds):
        r = _rng(seed0 + w)
        lvel[w, :, :2] = 1.5 * (r.rand(nb, 2) - 0.5)
        avel[w] = 6.0 * (r.rand(nb, 3) - 0.5)
    sc.state = dict(pos=pos, quat=quat, lvel=lvel, avel=avel)
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def conveyor(nworlds=1, seed0=71):
    """Kinematic bodies (dBodySetKinematic, ode.cpp:837-842): a platform that moves and turns at constant velocity whatever rests on
    it, carrying three boxes by friction above a ground plane."""
    sc = B.Scene(B.default_world_params(gravity=(0, 0, -9.81), max_contacts=4, surf_mode=B.CONTACT_APPROX1, mu=1.0), nworlds)
    sc.add_geom(B.PLANE, (0, 0, 1, 0))
    m, I = B.box_mass(1.0, 2.0, 1.0, 0.1)
    plat = sc.add_body(m, I, (0, 0, 0.5), flags=B.BODY_KINEMATIC)
    sc.add_geom(B.BOX, (2.0, 1.0, 0.1), body=plat)
    mb, Ib = B.box_mass(2.0, 0.3, 0.3, 0.3)
    for k in range(3):
        b = sc.add_body(mb, Ib, (-0.6 + 0.6 * k, 0.1 * k, 0.5 + 0.05 + 0.15 + 0.002))
        sc.add_geom(B.BOX, (0.3, 0.3, 0.3), body=b)
    nb = sc.nbody
    pos = np.tile(np.asarray(sc.body_pos)[None], (nworlds, 1, 1))
    quat = np.tile(np.array([1.0, 0, 0, 0])[None, None], (nworlds, nb, 1))
    lvel = np.zeros((nworlds, nb, 3))
    avel = np.zeros((nworlds, nb, 3))
    for w in range(nworlds):
        r = _rng(seed0 + w)
        lvel[w, 0] = (0.3 + 0.1 * r.rand(), 0.0, 0.0)
        avel[w, 0] = (0.0, 0.0, 0.2)
        lvel[w, 3] = (0.3, 0.0, 0.0)
    sc.state = dict(pos=pos, quat=quat, lvel=lvel, avel=avel)
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def composite(nworlds=1, seed0=81):
    """Composite bodies through geom offsets (dGeomSetOffsetPosition / dGeomSetOffsetQuaternion, collision_kernel.cpp:455-466): dumbbells =
    one body carrying two spheres at +-0.3 along its x axis and a capsule bar between them (offset rotation), tumbling onto a plane and
    onto each other; the demo_boxstack 'x' objects are built the same way."""
    sc = B.Scene(B.default_world_params(gravity=(0, 0, -9.81), max_contacts=4, surf_mode=B.CONTACT_APPROX1, mu=0.7), nworlds)
    sc.add_geom(B.PLANE, (0, 0, 1, 0))
    s = np.sqrt(0.5)
    for k in range(4):
        I = np.diag([0.02, 0.12, 0.12])
        b = sc.add_body(1.5, I, (0.25 * k, 0.3 * k, 0.5 + 0.55 * k), (np.cos(0.2 * k), 0.0, np.sin(0.2 * k), 0.0))
        sc.add_geom(B.SPHERE, (0.15,), body=b, offset_pos=(0.3, 0, 0))
        sc.add_geom(B.SPHERE, (0.15,), body=b, offset_pos=(-0.3, 0, 0))
        sc.add_geom(B.CAPSULE, (0.05, 0.5), body=b, offset_pos=(0, 0, 0), offset_quat=(s, 0.0, s, 0.0))      # capsule axis z -> body x
    nb = sc.nbody
    pos = np.tile(np.asarray(sc.body_pos)[None], (nworlds, 1, 1))
    quat = np.tile(np.asarray(sc.body_quat)[None], (nworlds, 1, 1))
    lvel = np.zeros((nworlds, nb, 3))
    avel = np.zeros((nworlds, nb, 3))
    for w in range(nworlds):
        r = _rng(seed0 + w)
        avel[w] = 2.0 * (r.rand(nb, 3) - 0.5)
    sc.state = dict(pos=pos, quat=quat, lvel=lvel, avel=avel)
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def motors(nworlds=1, seed0=91, dynamic_iterations=True):
    """Linear and angular motors (lmotor.cpp, amotor.cpp): a box driven along a global and a body-relative axis by an LMotor to the
    world; a two-link arm (ball joints) whose elbow carries an Euler-mode AMotor with stops, bounce and a powered middle axis; a
    ball-jointed pair with a user-mode AMotor (axes relative to body 1 / body 2 / global, one axis powered at its stop, one with
    lo == hi); an Euler AMotor and an LMotor attached as (world, body), i.e. with dJOINT_REVERSE; a 3-axis LMotor between two free boxes."""
    kw = dict(gravity=(0, 0, -9.81), max_contacts=4, surf_mode=B.CONTACT_APPROX1, mu=0.6, skip_connected=1)
    if not dynamic_iterations:      # dWorldSetQuickStepDynamicIterationParameters(w, 0, 0, ...): always exactly num_iterations sweeps
        kw.update(premature_exit_delta=0.0, max_extra_factor=0.0)
    sc = B.Scene(B.default_world_params(**kw), nworlds)
    sc.add_geom(B.PLANE, (0, 0, 1, 0))
    m, I = B.box_mass(2.0, 0.4, 0.3, 0.2)

    def box(pos):
        b = sc.add_body(m, I, pos)
        sc.add_geom(B.BOX, (0.4, 0.3, 0.2), body=b)
        return b
    b0 = box((0, 0, 0.1))
    sc.add_joint(B.JOINT_LMOTOR, b0, -1, motor_axes=((0, (1, 0, 0)), (1, (0, 1, 0))), vel=(0.5, -0.2), fmax=(8.0, 3.0))
    b1, b2 = box((2, 0, 1.0)), box((2.6, 0, 1.0))
    sc.add_joint(B.JOINT_BALL, b1, -1, (1.7, 0, 1.0))
    sc.add_joint(B.JOINT_BALL, b1, b2, (2.3, 0, 1.0))
    sc.add_joint(B.JOINT_AMOTOR, b1, b2, motor_mode=B.AMOTOR_EULER, motor_axes=((1, (1, 0, 0)), (0, (0, 1, 0)), (2, (0, 0, 1))),
                 lo_stop=(-0.3, -0.25, -0.4), hi_stop=(0.3, 0.25, 0.4), vel=(0, 0.5, 0), fmax=(0, 2.0, 0), bounce=(0.2, -1, 0.1))
    b3, b4 = box((4, 0, 1.0)), box((4, 0.6, 1.0))
    sc.add_joint(B.JOINT_BALL, b3, b4, (4, 0.3, 1.0))
    sc.add_joint(B.JOINT_AMOTOR, b3, b4, motor_mode=B.AMOTOR_USER, motor_axes=((1, (1, 0, 0)), (2, (0, 1, 0)), (0, (0, 0, 1))),
                 motor_angle=(0.5, 0.0, 0.0), lo_stop=(-0.2, -INF_, 0.0), hi_stop=(0.2, INF_, 0.0), vel=(0.3, -0.2, 0), fmax=(1.0, 1.0, 0),
                 fudge_factor=(0.5, -1, -1))
    b5 = box((6, 0, 1.5))
    sc.add_joint(B.JOINT_AMOTOR, -1, b5, motor_mode=B.AMOTOR_EULER, motor_axes=((1, (1, 0, 0)), (0, (0, 1, 0)), (2, (0, 0, 1))),
                 lo_stop=(-0.2, -0.2, -0.2), hi_stop=(0.2, 0.2, 0.2), stop_erp=(0.5, -1, -1), stop_cfm=(1e-3, -1, -1))
    b6, b7 = box((8, 0, 0.8)), box((8, 0.5, 1.1))
    sc.add_joint(B.JOINT_LMOTOR, b6, b7, motor_axes=((0, (0, 0, 1)), (1, (1, 0, 0)), (2, (0, 1, 0))), vel=(0.1, 0.3, -0.3), fmax=(5.0, 2.0, 2.0))
    b8 = box((10, 0, 0.6))
    sc.add_joint(B.JOINT_LMOTOR, -1, b8, motor_axes=((2, (1, 1, 0)),), vel=(0.4,), fmax=(6.0,))
    nb = sc.nbody
    pos = np.tile(np.asarray(sc.body_pos)[None], (nworlds, 1, 1))
    quat = np.tile(np.asarray(sc.body_quat)[None], (nworlds, 1, 1))
    lvel = np.zeros((nworlds, nb, 3))
    avel = np.zeros((nworlds, nb, 3))
    for w in range(nworlds):
        r = _rng(seed0 + w)
        avel[w] = 3.0 * (r.rand(nb, 3) - 0.5)
        lvel[w] = 0.3 * (r.rand(nb, 3) - 0.5)
    sc.state = dict(pos=pos, quat=quat, lvel=lvel, avel=avel)
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def sensors(nworlds=1, n=24, seed0=500, extent=1.2, space_type=B.SPACE_SIMPLE, cylinder_box=True):
    """Ray and cylinder colliders: n drifting bodies in zero gravity inside a cube of half-width `extent` above a ground plane, each
    carrying one of sphere / box / capsule / cylinder plus, on every second body, two ray geoms at offset poses (a forward "lidar" beam
    and a slanted one).  Rays report hits (odeb_get_ray_hits) and make no joints; cylinders collide with the plane and the spheres
    (collision_cylinder_plane.cpp, collision_cylinder_sphere.cpp) and the boxes (collision_cylinder_box.cpp; cylinder_box=False keeps
    boxes and cylinders apart through category bits); rays see everything."""
    kw = dict(gravity=(0, 0, -2.0), max_contacts=4, surf_mode=B.CONTACT_APPROX1, mu=0.5, space_type=space_type)
    sc = B.Scene(B.default_world_params(**kw), nworlds)
    CAT_BOX, CAT_CYL, CAT_OTHER, CAT_RAY = 1, 2, 4, 8
    sc.add_geom(B.PLANE, (0, 0, 1, -extent), category=CAT_OTHER)
    r = _rng(seed0)
    s = np.sqrt(0.5)
    for i in range(n):
        size = 0.25 + 0.25 * r.rand()
        kind = i % 4
        b = sc.add_body(1.0, np.eye(3) * 0.05, (0, 0, 0))
        if kind == 0:
            sc.add_geom(B.SPHERE, (0.5 * size,), body=b, category=CAT_OTHER)
        elif kind == 1:
            sc.add_geom(B.BOX, (size, 0.7 * size, 0.5 * size), body=b, category=CAT_BOX,
                        collide=CAT_BOX | CAT_OTHER | CAT_RAY | (CAT_CYL if cylinder_box else 0))
        elif kind == 2:
            sc.add_geom(B.CAPSULE, (0.2 * size, 0.8 * size), body=b, category=CAT_OTHER)
        else:
            sc.add_geom(B.CYLINDER, (0.4 * size, 0.6 * size), body=b, category=CAT_CYL,
                        collide=CAT_CYL | CAT_OTHER | CAT_RAY | (CAT_BOX if cylinder_box else 0))
        if i % 2 == 0:
            sc.add_geom(B.RAY, (1.5,), body=b, category=CAT_RAY, offset_pos=(0, 0, 0.6 * size))
            sc.add_geom(B.RAY, (1.0,), body=b, category=CAT_RAY, offset_pos=(0.1, 0, 0), offset_quat=(s, 0.0, s, 0.0))
    pos = np.zeros((nworlds, n, 3))
    quat = np.zeros((nworlds, n, 4))
    lvel = np.zeros((nworlds, n, 3))
    avel = np.zeros((nworlds, n, 3))
    for w in range(nworlds):
        rw = _rng(seed0 + 1 + w)
        pos[w] = extent * (2 * rw.rand(n, 3) - 1)
        q = rw.randn(n, 4)
        quat[w] = q / np.linalg.norm(q, axis=1, keepdims=True)
        lvel[w] = 0.8 * (rw.rand(n, 3) - 0.5)
        avel[w] = 2.0 * (rw.rand(n, 3) - 0.5)
    sc.state = dict(pos=pos, quat=quat, lvel=lvel, avel=avel)
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def cylinders_and_boxes(nworlds=1, n=30, seed0=700, extent=0.9):
    """Cylinder-box collider exerciser (collision_cylinder_box.cpp): flat discs, long rods and ordinary cylinders tumbling among boxes of
    mixed proportions in a small zero-gravity cube, so that face / edge / vertex / cap-edge separating axes and both clipping routines
    (cylinder side line against the box, box face against the cap octagon) all come up."""
    sc = B.Scene(B.default_world_params(gravity=(0, 0, 0), max_contacts=8, surf_mode=B.CONTACT_APPROX1, mu=0.3, space_type=B.SPACE_SIMPLE), nworlds)
    r = _rng(seed0)
    for i in range(n):
        b = sc.add_body(1.0, np.eye(3) * 0.05, (0, 0, 0))
        if i % 2 == 0:
            shape = (i // 2) % 3
            rad, length = ((0.35, 0.08), (0.08, 0.7), (0.2, 0.3))[shape]
            sc.add_geom(B.CYLINDER, (rad * (0.8 + 0.4 * r.rand()), length * (0.8 + 0.4 * r.rand())), body=b)
        else:
            sc.add_geom(B.BOX, tuple(0.15 + 0.45 * r.rand(3)), body=b)
    pos = np.zeros((nworlds, n, 3))
    quat = np.zeros((nworlds, n, 4))
    lvel = np.zeros((nworlds, n, 3))
    avel = np.zeros((nworlds, n, 3))
    for w in range(nworlds):
        rw = _rng(seed0 + 1 + w)
        pos[w] = extent * (2 * rw.rand(n, 3) - 1)
        q = rw.randn(n, 4)
        quat[w] = q / np.linalg.norm(q, axis=1, keepdims=True)
        if w % 4 == 0:                      # axis-aligned poses too: parallel axes, zero cross products
            quat[w, : n // 2] = (1, 0, 0, 0)
        lvel[w] = 0.6 * (rw.rand(n, 3) - 0.5)
        avel[w] = 1.5 * (rw.rand(n, 3) - 0.5)
    sc.state = dict(pos=pos, quat=quat, lvel=lvel, avel=avel)
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def scatter(nworlds=1, n=40, seed0=300, space_type=B.SPACE_HASH, extent=1.0, levels=None, plane=True):
    """Broadphase exerciser: n spheres / boxes / capsules of sizes 0.05 .. 1.5 (hash levels -3 .. 1) scattered in a cube of half-width
    `extent` around the origin, no gravity, slow drift.  Around the origin the hash space's cell addresses with negative z wrap
    (collision_space.cpp:499, :533), so its callback stream is a strict subset of the simple space's there; per world the positions
    and orientations differ, the sizes are the template's."""
    kw = dict(gravity=(0, 0, 0), max_contacts=2, surf_mode=B.CONTACT_APPROX1, mu=0.5, space_type=space_type)
    if levels is not None:
        kw.update(hash_levels_set=1, hash_minlevel=levels[0], hash_maxlevel=levels[1])
    sc = B.Scene(B.default_world_params(**kw), nworlds)
    if plane:
        sc.add_geom(B.PLANE, (0, 0, 1, -extent))
    r = _rng(seed0)
    for i in range(n):
        size = 0.05 * (30.0 ** r.rand())
        b = sc.add_body(1.0, np.eye(3) * 0.1, (0, 0, 0))
        kind = i % 3
        if kind == 0:
            sc.add_geom(B.SPHERE, (0.5 * size,), body=b)
        elif kind == 1:
            sc.add_geom(B.BOX, (size, size * (0.3 + 0.7 * r.rand()), size * (0.3 + 0.7 * r.rand())), body=b)
        else:
            sc.add_geom(B.CAPSULE, (0.15 * size, 0.7 * size), body=b)
    pos = np.zeros((nworlds, n, 3))
    quat = np.zeros((nworlds, n, 4))
    lvel = np.zeros((nworlds, n, 3))
    for w in range(nworlds):
        rw = _rng(seed0 + 1 + w)
        pos[w] = extent * (2 * rw.rand(n, 3) - 1)
        q = rw.randn(n, 4)
        quat[w] = q / np.linalg.norm(q, axis=1, keepdims=True)
        lvel[w] = 0.5 * (rw.rand(n, 3) - 0.5)
    sc.state = dict(pos=pos, quat=quat, lvel=lvel, avel=np.zeros_like(pos))
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def free_boxes(nworlds=1, nboxes=64, seed0=5, grid=8, spacing=1.5):
    """nboxes separate unit boxes resting/falling on the plane: many one-body islands per world
    (the scattered 64-body world of SURVEY.md 7.2(4))."""
    sc = B.Scene(B.default_world_params(gravity=(0, 0, -9.81), max_contacts=4, surf_mode=B.CONTACT_APPROX1,
                                        mu=0.5), nworlds)
    sc.add_geom(B.PLANE, (0, 0, 1, 0))
    m, I = B.box_mass(1.0, 1, 1, 1)
    for i in range(nboxes):
        b = sc.add_body(m, I, ((i % grid) * spacing, (i // grid) * spacing, 0.5 + 0.02))
        sc.add_geom(B.BOX, (1, 1, 1), body=b)
    pos = np.tile(np.asarray(sc.body_pos)[None], (nworlds, 1, 1))
    for w in range(nworlds):
        r = _rng(seed0 + w)
        pos[w, :, 2] += 0.05 * r.rand(nboxes)
    quat = np.tile(np.array([1.0, 0, 0, 0])[None, None], (nworlds, nboxes, 1))
    sc.state = dict(pos=pos, quat=quat, lvel=np.zeros_like(pos), avel=np.zeros_like(pos))
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def ragdoll(nworlds=1, seed0=11, drop=0.25, max_contacts=3):
    """Config 4: a 15-capsule humanoid (authored here; the reference ships no ragdoll), z up.

    14 joints: ball (neck, shoulders, hips), universal with stops on both axes (spine x2, wrists, ankles),
    hinge with stops (elbows, knees). Contacts: <= 3 per pair, dContactApprox1, mu = 1, bodies joined by a
    joint do not collide (dAreConnectedExcluding). g = -9.81, ERP 0.2, default CFM, 20 iterations.
    Per world: small random initial linear / angular velocities from seed0 + w.
    """
    sc = B.Scene(B.default_world_params(gravity=(0, 0, -9.81), max_contacts=max_contacts, surf_mode=B.CONTACT_APPROX1,
                                        mu=1.0, skip_connected=1), nworlds)
    sc.add_geom(B.PLANE, (0, 0, 1, 0))
    s = np.sqrt(0.5)
    QX = (s, 0.0, s, 0.0)      # capsule axis (local z) -> world x
    QY = (s, -s, 0.0, 0.0)     # local z -> world y
    QZ = (1.0, 0.0, 0.0, 0.0)
    idx = {}

    def cap(name, pos, r, length, q):
        m, I = B.capsule_mass(1000.0, r, length)
        b = sc.add_body(m, I, (pos[0], pos[1], pos[2] + drop), q)
        sc.add_geom(B.CAPSULE, (r, length), body=b)
        idx[name] = b

    cap("pelvis", (0, 0, 1.00), 0.11, 0.14, QX)
    cap("belly", (0, 0, 1.16), 0.10, 0.10, QX)
    cap("chest", (0, 0, 1.34), 0.12, 0.16, QX)
    cap("head", (0, 0, 1.64), 0.10, 0.05, QZ)
    for sgn, sd in ((1, "L"), (-1, "R")):
        cap("uarm" + sd, (sgn * 0.34, 0, 1.42), 0.05, 0.20, QX)
        cap("larm" + sd, (sgn * 0.63, 0, 1.42), 0.045, 0.20, QX)
        cap("hand" + sd, (sgn * 0.85, 0, 1.42), 0.04, 0.04, QX)
        cap("uleg" + sd, (sgn * 0.10, 0, 0.70), 0.07, 0.28, QZ)
        cap("lleg" + sd, (sgn * 0.10, 0, 0.29), 0.06, 0.28, QZ)
        cap("foot" + sd, (sgn * 0.10, 0.06, 0.055), 0.05, 0.12, QY)

    def J(jt, a, b, anchor, **kw):
        sc.add_joint(jt, idx[a], idx[b], (anchor[0], anchor[1], anchor[2] + drop), **kw)

    J(B.JOINT_UNIVERSAL, "pelvis", "belly", (0, 0, 1.08), axis1=(1, 0, 0), axis2=(0, 1, 0), lo_stop=(-0.4, -0.4), hi_stop=(0.4, 0.4))
    J(B.JOINT_UNIVERSAL, "belly", "chest", (0, 0, 1.24), axis1=(1, 0, 0), axis2=(0, 1, 0), lo_stop=(-0.4, -0.4), hi_stop=(0.4, 0.4))
    J(B.JOINT_BALL, "chest", "head", (0, 0, 1.52))
    for sgn, sd in ((1, "L"), (-1, "R")):
        J(B.JOINT_BALL, "chest", "uarm" + sd, (sgn * 0.19, 0, 1.42))
        J(B.JOINT_HINGE, "uarm" + sd, "larm" + sd, (sgn * 0.485, 0, 1.42), axis1=(0, 0, 1), lo_stop=(-2.0, 0), hi_stop=(0.05, 0))
        J(B.JOINT_UNIVERSAL, "larm" + sd, "hand" + sd, (sgn * 0.78, 0, 1.42), axis1=(0, 1, 0), axis2=(0, 0, 1),
          lo_stop=(-0.6, -0.6), hi_stop=(0.6, 0.6))
        J(B.JOINT_BALL, "pelvis", "uleg" + sd, (sgn * 0.10, 0, 0.92))
        J(B.JOINT_HINGE, "uleg" + sd, "lleg" + sd, (sgn * 0.10, 0, 0.495), axis1=(1, 0, 0), lo_stop=(-0.05, 0), hi_stop=(2.2, 0))
        J(B.JOINT_UNIVERSAL, "lleg" + sd, "foot" + sd, (sgn * 0.10, 0, 0.09), axis1=(1, 0, 0), axis2=(0, 1, 0),
          lo_stop=(-0.5, -0.3), hi_stop=(0.5, 0.3))
    nb = sc.nbody
    pos = np.tile(np.asarray(sc.body_pos)[None], (nworlds, 1, 1))
    quat = np.tile(np.asarray(sc.body_quat)[None], (nworlds, 1, 1))
    lvel = np.zeros((nworlds, nb, 3))
    avel = np.zeros((nworlds, nb, 3))
    for w in range(nworlds):
        r = _rng(seed0 + w)
        lvel[w] = 0.3 * (r.rand(nb, 3) - 0.5)
        avel[w] = 0.5 * (r.rand(nb, 3) - 0.5)
    sc.state = dict(pos=pos, quat=quat, lvel=lvel, avel=avel)
    sc.seeds = (seed0 + np.arange(nworlds)).astype(np.uint32)
    return sc


def wall(width=500, height=200, max_contacts=4, ball=True, seed=21, space_type=B.SPACE_SAP, mu=0.5, jitter=0.0):
    """Config 5: demo_crash.cpp:295-312-style brick wall scaled up, ONE world.

    width x height unit boxes (WBOXSIZE 1, mass 1), every other course shifted by half a brick, laid along x (the
    axis dSAP_AXES_XYZ sorts on) plus a cannon ball (sphere r=0.5, m=10, demo_crash.cpp:66-69) flying at the wall;
    dSweepAndPruneSpace, g=(0,0,-1.5), CFM 1e-5, ERP 0.8, contact parameters of demo_crash.cpp:133-141
    (Slip1|Slip2|SoftERP|SoftCFM|Approx1, mu=0.5, soft_erp 0.8, soft_cfm 0.01), <= 4 contacts per pair.
    """
    mode = B.CONTACT_SLIP1 | B.CONTACT_SLIP2 | B.CONTACT_SOFT_ERP | B.CONTACT_SOFT_CFM | B.CONTACT_APPROX1
    sc = B.Scene(B.default_world_params(gravity=(0, 0, -1.5), cfm=1e-5, erp=0.8, max_contacts=max_contacts, surf_mode=mode,
                                        mu=mu, slip1=0.0, slip2=0.0, soft_erp=0.8, soft_cfm=0.01, space_type=space_type), 1)
    sc.add_geom(B.PLANE, (0, 0, 1, 0))
    m, I = B.box_mass(1.0, 1, 1, 1)
    r = _rng(seed)
    for iz in range(height):
        for ix in range(width):
            j = jitter * (r.rand(2) - 0.5) if jitter else (0.0, 0.0)
            b = sc.add_body(m, I, (ix + 0.5 * (iz % 2) + j[0], j[1], 0.5 + iz))
            sc.add_geom(B.BOX, (1, 1, 1), body=b)
    nb = sc.nbody
    lvel = np.zeros((1, nb + (1 if ball else 0), 3))
    if ball:
        ms, Is = B.sphere_mass(1.0, 0.5)
        ms, Is = 10.0, Is * (10.0 / ms)          # dMassAdjust(&m, CANNON_BALL_MASS)
        b = sc.add_body(ms, Is, (0.5 * width, -6.0, 0.5 * min(height, 6) + 0.5))
        sc.add_geom(B.SPHERE, (0.5,), body=b)
        lvel[0, b] = (0.0, 30.0, 0.0)
        nb += 1
    pos = np.asarray(sc.body_pos)[None].copy()
    quat = np.tile(np.array([1.0, 0, 0, 0])[None, None], (1, nb, 1))
    sc.state = dict(pos=pos, quat=quat, lvel=lvel, avel=np.zeros_like(pos))
    sc.seeds = np.asarray([seed], dtype=np.uint32)
    return sc
