"""A small 'user application' written against the classic ODE C API through ctypes.

The same program runs on the unmodified reference (oracle/_ref/libode_ref_*.so exports the full ODE API) and on the
B200 library (ode_b200/libode_b200_*.so, include/ode_b200_classic.h); only the library path differs.  The
near-callback is the demo_boxstack.cpp:131-176 pattern with one addition SURVEY 7.2(1) describes: it buffers the
pairs it is handed and creates the contacts in canonical (geom index) order, so that constraint order does not depend
on the broadphase's internal traversal order.
"""
import ctypes as C
import math
import numpy as np

dJointTypeContact = 4
dContactBounce, dContactSoftCFM, dContactApprox1 = 0x004, 0x010, 0x7000
dParamLoStop, dParamHiStop, dParamLoStop2, dParamHiStop2 = 0, 1, 0x100, 0x101
dSAP_AXES_XYZ = (0) | (1 << 2) | (2 << 4)


def make_types(real):
    class dSurfaceParameters(C.Structure):
        _fields_ = [("mode", C.c_int), ("mu", real), ("mu2", real), ("rho", real), ("rho2", real), ("rhoN", real),
                    ("bounce", real), ("bounce_vel", real), ("soft_erp", real), ("soft_cfm", real),
                    ("motion1", real), ("motion2", real), ("motionN", real), ("slip1", real), ("slip2", real)]

    class dContactGeom(C.Structure):
        _fields_ = [("pos", real * 4), ("normal", real * 4), ("depth", real), ("g1", C.c_void_p), ("g2", C.c_void_p),
                    ("side1", C.c_int), ("side2", C.c_int)]

    class dContact(C.Structure):
        _fields_ = [("surface", dSurfaceParameters), ("geom", dContactGeom), ("fdir1", real * 4)]

    class dMass(C.Structure):
        _fields_ = [("mass", real), ("c", real * 4), ("I", real * 12)]

    class dJointFeedback(C.Structure):
        _fields_ = [("f1", real * 4), ("t1", real * 4), ("f2", real * 4), ("t2", real * 4)]

    class Stats(C.Structure):
        _fields_ = [("struct_size", C.c_uint), ("iteration_count", C.c_uint32), ("premature_exits", C.c_uint32),
                    ("prolonged_execs", C.c_uint32), ("full_extra_execs", C.c_uint32)]
    return dSurfaceParameters, dContactGeom, dContact, dMass, Stats, dJointFeedback


class Ode:
    """ctypes prototypes of the subset of the classic API the application uses."""

    def __init__(self, path, real):
        self.lib = L = C.CDLL(path)
        self.real = real
        self.np_real = np.float32 if real is C.c_float else np.float64
        self.dSurfaceParameters, self.dContactGeom, self.dContact, self.dMass, self.Stats, self.dJointFeedback = make_types(real)
        vp, r, i = C.c_void_p, real, C.c_int
        self.NearCallback = C.CFUNCTYPE(None, vp, vp, vp)
        sig = {
            "dInitODE2": (i, [C.c_uint]), "dCloseODE": (None, []), "dAllocateODEDataForThread": (i, [C.c_uint]),
            "dRandSetSeed": (None, [C.c_ulong]), "dRandGetSeed": (C.c_ulong, []),
            "dWorldCreate": (vp, []), "dWorldDestroy": (None, [vp]), "dWorldSetGravity": (None, [vp, r, r, r]),
            "dWorldSetCFM": (None, [vp, r]), "dWorldSetERP": (None, [vp, r]), "dWorldSetQuickStepNumIterations": (None, [vp, i]),
            "dWorldSetContactMaxCorrectingVel": (None, [vp, r]), "dWorldSetContactSurfaceLayer": (None, [vp, r]),
            "dWorldSetAutoDisableFlag": (None, [vp, i]), "dWorldQuickStep": (i, [vp, r]),
            "dWorldAttachQuickStepDynamicIterationStatisticsSink": (i, [vp, vp]),
            "dBodyCreate": (vp, [vp]), "dBodySetPosition": (None, [vp, r, r, r]), "dBodySetQuaternion": (None, [vp, C.POINTER(r)]),
            "dBodySetLinearVel": (None, [vp, r, r, r]), "dBodySetAngularVel": (None, [vp, r, r, r]),
            "dBodyGetPosition": (C.POINTER(r), [vp]), "dBodyGetQuaternion": (C.POINTER(r), [vp]),
            "dBodyGetLinearVel": (C.POINTER(r), [vp]), "dBodyGetAngularVel": (C.POINTER(r), [vp]), "dBodyGetRotation": (C.POINTER(r), [vp]),
            "dBodySetMass": (None, [vp, vp]), "dBodyAddForce": (None, [vp, r, r, r]), "dBodyIsEnabled": (i, [vp]),
            "dMassSetBox": (None, [vp, r, r, r, r]), "dMassSetSphere": (None, [vp, r, r]), "dMassSetCapsule": (None, [vp, r, i, r, r]),
            "dMassSetBoxTotal": (None, [vp, r, r, r, r]), "dMassAdjust": (None, [vp, r]),
            "dQFromAxisAndAngle": (None, [C.POINTER(r), r, r, r, r]),
            "dHashSpaceCreate": (vp, [vp]), "dSweepAndPruneSpaceCreate": (vp, [vp, i]), "dSimpleSpaceCreate": (vp, [vp]), "dSpaceDestroy": (None, [vp]),
            "dSpaceCollide": (None, [vp, vp, self.NearCallback]), "dCollide": (i, [vp, vp, i, vp, i]),
            "dCreateSphere": (vp, [vp, r]), "dCreateBox": (vp, [vp, r, r, r]), "dCreateCapsule": (vp, [vp, r, r]),
            "dCreatePlane": (vp, [vp, r, r, r, r]), "dGeomSetBody": (None, [vp, vp]), "dGeomGetBody": (vp, [vp]),
            "dGeomSetData": (None, [vp, vp]), "dGeomGetData": (vp, [vp]), "dGeomGetAABB": (None, [vp, C.POINTER(r)]),
            "dGeomSetPosition": (None, [vp, r, r, r]),
            "dJointGroupCreate": (vp, [i]), "dJointGroupEmpty": (None, [vp]), "dJointGroupDestroy": (None, [vp]),
            "dJointCreateContact": (vp, [vp, vp, vp]), "dJointCreateBall": (vp, [vp, vp]), "dJointCreateHinge": (vp, [vp, vp]),
            "dJointCreateUniversal": (vp, [vp, vp]), "dJointAttach": (None, [vp, vp, vp]),
            "dJointSetBallAnchor": (None, [vp, r, r, r]), "dJointSetHingeAnchor": (None, [vp, r, r, r]), "dJointSetHingeAxis": (None, [vp, r, r, r]),
            "dJointSetHingeParam": (None, [vp, i, r]), "dJointSetUniversalAnchor": (None, [vp, r, r, r]),
            "dJointSetUniversalAxis1": (None, [vp, r, r, r]), "dJointSetUniversalAxis2": (None, [vp, r, r, r]),
            "dJointSetUniversalParam": (None, [vp, i, r]),
            "dJointCreateLMotor": (vp, [vp, vp]), "dJointSetLMotorNumAxes": (None, [vp, i]), "dJointSetLMotorAxis": (None, [vp, i, i, r, r, r]),
            "dJointSetLMotorParam": (None, [vp, i, r]), "dJointCreateAMotor": (vp, [vp, vp]), "dJointSetAMotorMode": (None, [vp, i]),
            "dJointSetAMotorNumAxes": (None, [vp, i]), "dJointSetAMotorAxis": (None, [vp, i, i, r, r, r]), "dJointSetAMotorAngle": (None, [vp, i, r]),
            "dJointSetAMotorParam": (None, [vp, i, r]), "dJointGetAMotorAxisRel": (i, [vp, i]), "dJointGetAMotorNumAxes": (i, [vp]),
            "dGeomSetOffsetPosition": (None, [vp, r, r, r]), "dGeomSetOffsetQuaternion": (None, [vp, C.POINTER(r)]),
            "dHashSpaceSetLevels": (None, [vp, i, i]),
            "dJointSetFeedback": (None, [vp, vp]), "dJointGetFeedback": (vp, [vp]), "dAreConnectedExcluding": (i, [vp, vp, i]), "dAreConnected": (i, [vp, vp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
            setattr(self, name, f)


class App:
    """One world + one space + a contact group, stepped with collide -> quickstep -> empty."""

    def __init__(self, ode, space="hash", gravity=(0, 0, -9.81), max_contacts=8, surface="approx1", cfm=None, iters=20):
        self.o = o = ode
        o.dInitODE2(0)
        o.dAllocateODEDataForThread(0xffffffff)
        self.world = o.dWorldCreate()
        self.space = {"hash": lambda: o.dHashSpaceCreate(None), "simple": lambda: o.dSimpleSpaceCreate(None),
                      "sap": lambda: o.dSweepAndPruneSpaceCreate(None, dSAP_AXES_XYZ)}[space]()
        self.group = o.dJointGroupCreate(0)
        o.dWorldSetGravity(self.world, *gravity)
        if cfm is not None:
            o.dWorldSetCFM(self.world, cfm)
        o.dWorldSetQuickStepNumIterations(self.world, iters)
        self.stats = o.Stats()
        self.stats.struct_size = C.sizeof(o.Stats)
        assert o.dWorldAttachQuickStepDynamicIterationStatisticsSink(self.world, C.byref(self.stats)) == 1
        self.bodies, self.geoms = [], []
        self.max_contacts, self.surface = max_contacts, surface
        self.pairs = []
        self.ncontacts = 0
        self.contact_log = []
        self._cb = o.NearCallback(self._near)
        # dJointSetFeedback: structs for permanent joints (attach_feedback) and, when contact_feedback is set, for every
        # contact joint of the step; filled with a marker so that "not written by the step" is observable
        self.feedback = []
        self.contact_feedback = False
        self.contact_fb = []

    # ---- scene building
    def _add_geom(self, g, body):
        if body is not None:
            self.o.dGeomSetBody(g, body)
        self.o.dGeomSetData(g, C.c_void_p(len(self.geoms) + 1))
        self.geoms.append(g)
        return g

    def plane(self, a, b, c, d):
        return self._add_geom(self.o.dCreatePlane(self.space, a, b, c, d), None)

    def body(self, pos, quat=None):
        b = self.o.dBodyCreate(self.world)
        self.o.dBodySetPosition(b, *pos)
        if quat is not None:
            q = (self.o.real * 4)(*quat)
            self.o.dBodySetQuaternion(b, q)
        self.bodies.append(b)
        return b

    def box(self, pos, sides, density=1.0, quat=None):
        b = self.body(pos, quat)
        m = self.o.dMass()
        self.o.dMassSetBox(C.byref(m), density, *sides)
        self.o.dBodySetMass(b, C.byref(m))
        self._add_geom(self.o.dCreateBox(self.space, *sides), b)
        return b

    def sphere(self, pos, radius, density=1.0):
        b = self.body(pos)
        m = self.o.dMass()
        self.o.dMassSetSphere(C.byref(m), density, radius)
        self.o.dBodySetMass(b, C.byref(m))
        self._add_geom(self.o.dCreateSphere(self.space, radius), b)
        return b

    def capsule(self, pos, radius, length, density=1.0, quat=None):
        b = self.body(pos, quat)
        m = self.o.dMass()
        self.o.dMassSetCapsule(C.byref(m), density, 3, radius, length)
        self.o.dBodySetMass(b, C.byref(m))
        self._add_geom(self.o.dCreateCapsule(self.space, radius, length), b)
        return b

    # ---- the near-callback (user code)
    def _near(self, data, o1, o2):
        i1 = self.o.dGeomGetData(o1)
        i2 = self.o.dGeomGetData(o2)
        if i1 > i2:
            o1, o2, i1, i2 = o2, o1, i2, i1
        self.pairs.append((i1, i2, o1, o2))

    def _make_contacts(self):
        o = self.o
        N = self.max_contacts
        arr = (o.dContact * N)()
        for (i1, i2, o1, o2) in sorted(self.pairs, key=lambda t: (t[0], t[1])):
            b1, b2 = o.dGeomGetBody(o1), o.dGeomGetBody(o2)
            if b1 and b2 and o.dAreConnectedExcluding(b1, b2, dJointTypeContact):
                continue
            n = o.dCollide(o1, o2, N, C.byref(arr[0].geom), C.sizeof(o.dContact))
            for k in range(n):
                s = arr[k].surface
                if self.surface == "approx1":
                    s.mode, s.mu = dContactApprox1, 0.5
                elif self.surface == "boxstack":      # demo_boxstack.cpp:144-151
                    s.mode, s.mu, s.mu2, s.bounce, s.bounce_vel, s.soft_cfm = dContactBounce | dContactSoftCFM, float("inf"), 0.0, 0.1, 0.1, 0.01
                else:                                  # demo_chain2.cpp:74-80
                    s.mode, s.mu = 0, float("inf")
                j = o.dJointCreateContact(self.world, self.group, C.byref(arr[k]))
                o.dJointAttach(j, b1, b2)
                if self.contact_feedback:
                    self.contact_fb.append(self._new_feedback(j))
                g = arr[k].geom
                self.contact_log.append((i1, i2, tuple(g.pos)[:3], tuple(g.normal)[:3], g.depth))
                self.ncontacts += 1

    MARK = -4321.5

    def _new_feedback(self, joint):
        fb = self.o.dJointFeedback()
        for v in (fb.f1, fb.t1, fb.f2, fb.t2):
            for k in range(4):
                v[k] = self.MARK
        self.o.dJointSetFeedback(joint, C.byref(fb))
        return fb

    def attach_feedback(self, joint):
        self.feedback.append(self._new_feedback(joint))

    def feedback_values(self):
        """[n, 12] f1 t1 f2 t2 of the permanent joints with feedback, then of the step's contact joints"""
        out = []
        for fb in self.feedback + self.contact_fb:
            out.append([fb.f1[0], fb.f1[1], fb.f1[2], fb.t1[0], fb.t1[1], fb.t1[2], fb.f2[0], fb.f2[1], fb.f2[2], fb.t2[0], fb.t2[1], fb.t2[2]])
        return np.array(out, self.o.np_real).reshape(-1, 12)

    def step(self, h, seed=None):
        o = self.o
        if seed is not None:
            o.dRandSetSeed(seed)
        self.pairs = []
        self.contact_log = []
        self.contact_fb = []
        o.dSpaceCollide(self.space, None, self._cb)
        self._make_contacts()
        ok = o.dWorldQuickStep(self.world, h)
        o.dJointGroupEmpty(self.group)
        return ok

    def state(self):
        o = self.o
        out = np.zeros((len(self.bodies), 13), o.np_real)
        for k, b in enumerate(self.bodies):
            p, q, l, a = o.dBodyGetPosition(b), o.dBodyGetQuaternion(b), o.dBodyGetLinearVel(b), o.dBodyGetAngularVel(b)
            out[k] = [p[0], p[1], p[2], q[0], q[1], q[2], q[3], l[0], l[1], l[2], a[0], a[1], a[2]]
        return out

    def pair_set(self):
        return sorted((a, b) for (a, b, _, _) in self.pairs)

    def close(self):
        o = self.o
        o.dJointGroupDestroy(self.group)
        o.dSpaceDestroy(self.space)
        o.dWorldDestroy(self.world)
        o.dCloseODE()


def scene_mixed_pile(app, n=12):
    """plane + boxes, spheres and capsules dropped in a loose pile (collider coverage: every primitive pair type)"""
    app.plane(0, 0, 1, 0)
    for i in range(n):
        x, y, z = 0.13 * (i % 3) - 0.1, 0.11 * ((i // 3) % 2), 0.4 + 0.55 * i
        kind = i % 3
        if kind == 0:
            app.box((x, y, z), (0.5, 0.4, 0.3), density=2.0, quat=(math.cos(0.1 * i), 0.0, math.sin(0.1 * i), 0.0))
        elif kind == 1:
            app.sphere((x, y, z), 0.25, density=1.5)
        else:
            app.capsule((x, y, z), 0.15, 0.4, density=1.0, quat=(math.cos(0.3), math.sin(0.3), 0.0, 0.0))


def scene_stack(app, n=8):
    """demo_boxstack-style stack of boxes on a plane"""
    app.plane(0, 0, 1, 0)
    for i in range(n):
        app.box((0.01 * (i % 3), 0.005 * (i % 2), 0.25 + 0.501 * i), (0.5, 0.5, 0.5), density=5.0)


def scene_linkage(app):
    """boxes hanging from the environment: ball, hinge (with stops) and universal (with stops) joints + ground contacts"""
    o = app.o
    app.plane(0, 0, 1, 0)
    bs = [app.box((0.3 * i, 0.0, 1.0), (0.25, 0.1, 0.1), density=3.0) for i in range(6)]
    j = o.dJointCreateBall(app.world, None)
    o.dJointAttach(j, bs[0], None)
    o.dJointSetBallAnchor(j, -0.15, 0.0, 1.0)
    for i in range(5):
        a, b = bs[i], bs[i + 1]
        x = 0.3 * i + 0.15
        if i % 3 == 0:
            j = o.dJointCreateHinge(app.world, None)
            o.dJointAttach(j, a, b)
            o.dJointSetHingeAnchor(j, x, 0.0, 1.0)
            o.dJointSetHingeAxis(j, 0.0, 1.0, 0.0)
            o.dJointSetHingeParam(j, dParamLoStop, -0.4)
            o.dJointSetHingeParam(j, dParamHiStop, 0.4)
        elif i % 3 == 1:
            j = o.dJointCreateUniversal(app.world, None)
            o.dJointAttach(j, a, b)
            o.dJointSetUniversalAnchor(j, x, 0.0, 1.0)
            o.dJointSetUniversalAxis1(j, 0.0, 1.0, 0.0)
            o.dJointSetUniversalAxis2(j, 0.0, 0.0, 1.0)
            o.dJointSetUniversalParam(j, dParamLoStop, -0.5)
            o.dJointSetUniversalParam(j, dParamHiStop, 0.5)
            o.dJointSetUniversalParam(j, dParamLoStop2, -0.3)
            o.dJointSetUniversalParam(j, dParamHiStop2, 0.3)
        else:
            j = o.dJointCreateBall(app.world, None)
            o.dJointAttach(j, a, b)
            o.dJointSetBallAnchor(j, x, 0.0, 1.0)
        app.perm_joints = getattr(app, "perm_joints", []) + [j]
    return bs


def scene_motors(app):
    """a driven box (LMotor to the environment), a two-link arm whose elbow carries an Euler-mode AMotor with stops, a user-mode AMotor
    whose velocity target the application changes while it runs, and a composite dumbbell (geom offsets) dropped beside them"""
    o = app.o
    dParamVel, dParamFMax, dParamVel2, dParamFMax2, dParamLoStop3, dParamHiStop3 = 2, 5, 0x102, 0x105, 0x200, 0x201
    app.plane(0, 0, 1, 0)
    bs = [app.box((0.0, 0.0, 0.1), (0.4, 0.3, 0.2), density=2.0)]
    j = o.dJointCreateLMotor(app.world, None)
    o.dJointAttach(j, bs[0], None)
    o.dJointSetLMotorNumAxes(j, 2)
    o.dJointSetLMotorAxis(j, 0, 0, 1.0, 0.0, 0.0)
    o.dJointSetLMotorAxis(j, 1, 1, 0.0, 1.0, 0.0)
    o.dJointSetLMotorParam(j, dParamVel, 0.5)
    o.dJointSetLMotorParam(j, dParamFMax, 8.0)
    o.dJointSetLMotorParam(j, dParamVel2, -0.2)
    o.dJointSetLMotorParam(j, dParamFMax2, 3.0)
    a, b = app.box((2.0, 0.0, 1.0), (0.4, 0.3, 0.2), density=2.0), app.box((2.6, 0.0, 1.0), (0.4, 0.3, 0.2), density=2.0)
    bs += [a, b]
    j = o.dJointCreateBall(app.world, None)
    o.dJointAttach(j, a, None)
    o.dJointSetBallAnchor(j, 1.7, 0.0, 1.0)
    j = o.dJointCreateBall(app.world, None)
    o.dJointAttach(j, a, b)
    o.dJointSetBallAnchor(j, 2.3, 0.0, 1.0)
    j = o.dJointCreateAMotor(app.world, None)
    o.dJointAttach(j, a, b)
    o.dJointSetAMotorMode(j, 1)
    o.dJointSetAMotorAxis(j, 0, 1, 1.0, 0.0, 0.0)
    o.dJointSetAMotorAxis(j, 2, 2, 0.0, 0.0, 1.0)
    for grp in (0, 0x100, 0x200):
        o.dJointSetAMotorParam(j, dParamLoStop + grp, -0.3)
        o.dJointSetAMotorParam(j, dParamHiStop + grp, 0.3)
    assert o.dJointGetAMotorNumAxes(j) == 3 and o.dJointGetAMotorAxisRel(j, 2) == 2
    c, d = app.box((4.0, 0.0, 1.0), (0.4, 0.3, 0.2), density=2.0), app.box((4.0, 0.6, 1.0), (0.4, 0.3, 0.2), density=2.0)
    bs += [c, d]
    j = o.dJointCreateBall(app.world, None)
    o.dJointAttach(j, c, d)
    o.dJointSetBallAnchor(j, 4.0, 0.3, 1.0)
    j = o.dJointCreateAMotor(app.world, None)
    o.dJointAttach(j, c, d)
    o.dJointSetAMotorNumAxes(j, 2)
    o.dJointSetAMotorAxis(j, 0, 1, 1.0, 0.0, 0.0)
    o.dJointSetAMotorAxis(j, 1, 2, 0.0, 1.0, 0.0)
    o.dJointSetAMotorParam(j, dParamFMax, 1.0)
    o.dJointSetAMotorParam(j, dParamFMax2, 1.0)
    app.user_motor = j
    # composite dumbbell: one body, two offset spheres and an offset capsule
    e = app.body((6.0, 0.0, 0.8))
    m = o.dMass()
    o.dMassSetBoxTotal(C.byref(m), 1.5, 0.7, 0.2, 0.2)
    o.dBodySetMass(e, C.byref(m))
    for ox in (0.3, -0.3):
        g = app._add_geom(o.dCreateSphere(app.space, 0.15), e)
        o.dGeomSetOffsetPosition(g, ox, 0.0, 0.0)
    g = app._add_geom(o.dCreateCapsule(app.space, 0.05, 0.5), e)
    s = math.sqrt(0.5)
    o.dGeomSetOffsetQuaternion(g, (o.real * 4)(s, 0.0, s, 0.0))
    o.dBodySetAngularVel(e, 0.5, 1.0, -0.7)
    bs.append(e)
    return bs
