"""Python surface of the reference's own binding (bindings/python/ode.pyx: World, Body, Mass, JointGroup, the joint classes, the
geom classes, Space / SimpleSpace / HashSpace, Contact, collide(), areConnected()) over the classic C API of the B200 library
(include/ode_b200_classic.h) through ctypes -- SURVEY.md 8(f) rank 3.  A script written for the reference's `ode` module runs on it with

    from ode_b200 import ode

Class, method and constant names, argument order and return shapes follow ode.pyx (cited per class).  Every call goes to the C symbol
of the same name as in the reference; a method whose C function is outside the library's subset raises NotImplementedError naming it.
`use(path)` points the module at another library exporting the classic API (the parity tests load oracle/_ref's unmodified
reference that way and run the same script on both).  There is no CPU fallback: stepping without a CUDA device makes
World.quickStep raise.
"""
import ctypes as C
import os
import weakref

_HERE = os.path.dirname(os.path.abspath(__file__))

# ---- constants (ode.pyx:22-113)
paramLoStop, paramHiStop, paramVel, paramLoVel, paramHiVel, paramFMax, paramFudgeFactor, paramBounce, paramCFM, paramStopERP, paramStopCFM, \
    paramSuspensionERP, paramSuspensionCFM, paramERP = range(14)
for _g, _sfx in ((0, ""), (256, "2"), (512, "3")):
    for _i, _n in enumerate(("LoStop", "HiStop", "Vel", "LoVel", "HiVel", "FMax", "FudgeFactor", "Bounce", "CFM", "StopERP", "StopCFM",
                             "SuspensionERP", "SuspensionCFM", "ERP")):
        globals()["Param" + _n + _sfx] = _g + _i
ParamGroup = 256
ContactMu2 = ContactAxisDep = 0x001
ContactFDir1, ContactBounce, ContactSoftERP, ContactSoftCFM = 0x002, 0x004, 0x008, 0x010
ContactMotion1, ContactMotion2, ContactMotionN, ContactSlip1, ContactSlip2, ContactRolling = 0x020, 0x040, 0x080, 0x100, 0x200, 0x400
ContactApprox0, ContactApprox1_1, ContactApprox1_2, ContactApprox1_N, ContactApprox1 = 0x0000, 0x1000, 0x2000, 0x4000, 0x7000
AMotorUser, AMotorEuler = 0, 1
Infinity = float("inf")
environment = None

_geom_c2py_lut = weakref.WeakValueDictionary()     # ode.pyx:112


class _Lib(object):
    """The loaded classic-API library: ctypes prototypes are attached on first use of each symbol."""

    def __init__(self, path):
        self.path = path
        self.cdll = C.CDLL(path)
        chk = self.cdll.dCheckConfiguration
        chk.restype, chk.argtypes = C.c_int, [C.c_char_p]
        self.double = bool(chk(b"ODE_double_precision"))
        self.real = r = C.c_double if self.double else C.c_float
        self.vec3, self.vec4, self.mat3 = r * 4, r * 4, r * 12

        class dMass(C.Structure):
            _fields_ = [("mass", r), ("c", r * 4), ("I", r * 12)]

        class dSurfaceParameters(C.Structure):
            _fields_ = [("mode", C.c_int), ("mu", r), ("mu2", r), ("rho", r), ("rho2", r), ("rhoN", r), ("bounce", r), ("bounce_vel", r),
                        ("soft_erp", r), ("soft_cfm", r), ("motion1", r), ("motion2", r), ("motionN", r), ("slip1", r), ("slip2", r)]

        class dContactGeom(C.Structure):
            _fields_ = [("pos", r * 4), ("normal", r * 4), ("depth", r), ("g1", C.c_void_p), ("g2", C.c_void_p), ("side1", C.c_int), ("side2", C.c_int)]

        class dContact(C.Structure):
            _fields_ = [("surface", dSurfaceParameters), ("geom", dContactGeom), ("fdir1", r * 4)]

        class dJointFeedback(C.Structure):
            _fields_ = [("f1", r * 4), ("t1", r * 4), ("f2", r * 4), ("t2", r * 4)]
        self.dMass, self.dContactGeom, self.dContact, self.dJointFeedback = dMass, dContactGeom, dContact, dJointFeedback
        self.NearCallback = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)
        self._fn = {}

    def f(self, name, res, *args):
        fn = self._fn.get(name)
        if fn is None:
            try:
                fn = getattr(self.cdll, name)
            except AttributeError:
                raise NotImplementedError("%s is outside the subset exported by %s" % (name, os.path.basename(self.path)))
            fn.restype, fn.argtypes = res, list(args)
            self._fn[name] = fn
        return fn


_lib = None


def use(path=None, precision=None):
    """Select the library behind this module: default = the in-tree CUDA build (ODE_B200_PRECISION=single|double, default single)."""
    global _lib
    if path is None:
        precision = precision or os.environ.get("ODE_B200_PRECISION", "single")
        path = os.path.join(_HERE, "libode_b200_%s.so" % precision)
        if not os.path.exists(path):
            raise RuntimeError("CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    _lib = _Lib(path)
    InitODE()
    return _lib


def _L():
    if _lib is None:
        use()
    return _lib


VP, I_ = C.c_void_p, C.c_int


def _v3(p):
    return tuple(float(p[i]) for i in range(3))


def _bid(body):
    return body.bid if body is not None else None


# ------------------------------------------------------------------------------------------------ Mass (ode.pyx:116-416)
class Mass(object):
    def __init__(self):
        L = _L()
        self._mass = L.dMass()
        L.f("dMassSetZero", None, VP)(C.byref(self._mass))

    def setZero(self):
        _L().f("dMassSetZero", None, VP)(C.byref(self._mass))

    def setParameters(self, mass, cgx, cgy, cgz, I11, I22, I33, I12, I13, I23):
        r = _L().real
        _L().f("dMassSetParameters", None, VP, r, r, r, r, r, r, r, r, r, r)(C.byref(self._mass), mass, cgx, cgy, cgz, I11, I22, I33, I12, I13, I23)

    def setSphere(self, density, radius):
        r = _L().real
        _L().f("dMassSetSphere", None, VP, r, r)(C.byref(self._mass), density, radius)

    def setSphereTotal(self, total_mass, radius):
        r = _L().real
        _L().f("dMassSetSphereTotal", None, VP, r, r)(C.byref(self._mass), total_mass, radius)

    def setCapsule(self, density, direction, radius, length):
        r = _L().real
        _L().f("dMassSetCapsule", None, VP, r, I_, r, r)(C.byref(self._mass), density, direction, radius, length)

    def setCapsuleTotal(self, total_mass, direction, radius, length):
        r = _L().real
        _L().f("dMassSetCapsuleTotal", None, VP, r, I_, r, r)(C.byref(self._mass), total_mass, direction, radius, length)

    def setCylinder(self, density, direction, r_, h):
        r = _L().real
        _L().f("dMassSetCylinder", None, VP, r, I_, r, r)(C.byref(self._mass), density, direction, r_, h)

    def setBox(self, density, lx, ly, lz):
        r = _L().real
        _L().f("dMassSetBox", None, VP, r, r, r, r)(C.byref(self._mass), density, lx, ly, lz)

    def setBoxTotal(self, total_mass, lx, ly, lz):
        r = _L().real
        _L().f("dMassSetBoxTotal", None, VP, r, r, r, r)(C.byref(self._mass), total_mass, lx, ly, lz)

    def adjust(self, newmass):
        _L().f("dMassAdjust", None, VP, _L().real)(C.byref(self._mass), newmass)

    def translate(self, t):
        r = _L().real
        _L().f("dMassTranslate", None, VP, r, r, r)(C.byref(self._mass), t[0], t[1], t[2])

    def add(self, b):
        _L().f("dMassAdd", None, VP, VP)(C.byref(self._mass), C.byref(b._mass))

    @property
    def mass(self):
        return float(self._mass.mass)

    @mass.setter
    def mass(self, v):
        self.adjust(v)

    @property
    def c(self):
        return _v3(self._mass.c)

    @property
    def I(self):
        m = self._mass.I
        return ((m[0], m[1], m[2]), (m[4], m[5], m[6]), (m[8], m[9], m[10]))

    def __str__(self):
        m = self._mass.I
        return ("Mass=%s\nCg=(%s, %s, %s)\nI11=%s I22=%s I33=%s\nI12=%s I13=%s I23=%s"
                % (self.mass, self.c[0], self.c[1], self.c[2], m[0], m[5], m[10], m[1], m[2], m[6]))


# ------------------------------------------------------------------------------------------------ Contact (ode.pyx:419-747)
def _surface_prop(field):
    def get(self):
        return getattr(self._contact.surface, field)

    def set_(self, v):
        setattr(self._contact.surface, field, v)
    return get, set_


class Contact(object):
    def __init__(self):
        self._contact = _L().dContact()
        self._contact.surface.mode = ContactBounce
        self._contact.surface.mu = Infinity
        self._contact.surface.bounce = 0.1

    getMode, setMode = _surface_prop("mode")
    getMu, setMu = _surface_prop("mu")
    getMu2, setMu2 = _surface_prop("mu2")
    getBounce, setBounce = _surface_prop("bounce")
    getBounceVel, setBounceVel = _surface_prop("bounce_vel")
    getSoftERP, setSoftERP = _surface_prop("soft_erp")
    getSoftCFM, setSoftCFM = _surface_prop("soft_cfm")
    getMotion1, setMotion1 = _surface_prop("motion1")
    getMotion2, setMotion2 = _surface_prop("motion2")
    getSlip1, setSlip1 = _surface_prop("slip1")
    getSlip2, setSlip2 = _surface_prop("slip2")

    def getFDir1(self):
        return _v3(self._contact.fdir1)

    def setFDir1(self, fdir):
        for k in range(3):
            self._contact.fdir1[k] = fdir[k]

    def getContactGeomParams(self):
        g = self._contact.geom
        return _v3(g.pos), _v3(g.normal), float(g.depth), _geom_c2py_lut.get(g.g1), _geom_c2py_lut.get(g.g2)

    def setContactGeomParams(self, pos, normal, depth, g1=None, g2=None):
        g = self._contact.geom
        for k in range(3):
            g.pos[k] = pos[k]
            g.normal[k] = normal[k]
        g.depth = depth
        if g1 is not None:
            g.g1 = g1._id()
        if g2 is not None:
            g.g2 = g2._id()


# ------------------------------------------------------------------------------------------------ World (ode.pyx:750-1137)
def _world_prop(cname, ctype=None):
    def get(self):
        L = _L()
        return L.f("dWorldGet" + cname, ctype or L.real, VP)(self.wid)

    def set_(self, v):
        L = _L()
        L.f("dWorldSet" + cname, None, VP, ctype or L.real)(self.wid, v)
    return get, set_


class World(object):
    def __init__(self):
        self._lib = _L()
        self.wid = self._lib.f("dWorldCreate", VP)()

    def __del__(self):
        if getattr(self, "wid", None):
            self._lib.f("dWorldDestroy", None, VP)(self.wid)
            self.wid = None

    def setGravity(self, gravity):
        r = _L().real
        _L().f("dWorldSetGravity", None, VP, r, r, r)(self.wid, gravity[0], gravity[1], gravity[2])

    def getGravity(self):
        g = _L().vec3()
        _L().f("dWorldGetGravity", None, VP, VP)(self.wid, g)
        return _v3(g)

    getERP, setERP = _world_prop("ERP")
    getCFM, setCFM = _world_prop("CFM")
    getQuickStepNumIterations, setQuickStepNumIterations = _world_prop("QuickStepNumIterations", I_)
    getContactMaxCorrectingVel, setContactMaxCorrectingVel = _world_prop("ContactMaxCorrectingVel")
    getContactSurfaceLayer, setContactSurfaceLayer = _world_prop("ContactSurfaceLayer")
    getAutoDisableFlag, setAutoDisableFlag = _world_prop("AutoDisableFlag", I_)
    getAutoDisableLinearThreshold, setAutoDisableLinearThreshold = _world_prop("AutoDisableLinearThreshold")
    getAutoDisableAngularThreshold, setAutoDisableAngularThreshold = _world_prop("AutoDisableAngularThreshold")
    getAutoDisableSteps, setAutoDisableSteps = _world_prop("AutoDisableSteps", I_)
    getAutoDisableTime, setAutoDisableTime = _world_prop("AutoDisableTime")
    getLinearDamping, setLinearDamping = _world_prop("LinearDamping")
    getAngularDamping, setAngularDamping = _world_prop("AngularDamping")

    def step(self, stepsize):
        """dWorldStep (the big-matrix LCP stepper) is not on the hot path; the symbol is not exported."""
        if not _L().f("dWorldStep", I_, VP, _L().real)(self.wid, stepsize):
            raise RuntimeError("dWorldStep failed")

    def quickStep(self, stepsize):
        """ode.pyx:857-871 -> dWorldQuickStep"""
        if not _L().f("dWorldQuickStep", I_, VP, _L().real)(self.wid, stepsize):
            raise RuntimeError("dWorldQuickStep failed (no CUDA device, or out of memory)")

    def impulseToForce(self, stepsize, impulse):
        return tuple(i / stepsize for i in impulse)        # dWorldImpulseToForce ode.cpp:2183-2191


# ------------------------------------------------------------------------------------------------ Body (ode.pyx:1140-1742)
class Body(object):
    def __init__(self, world):
        self.world = world
        self._lib = _L()
        self.bid = self._lib.f("dBodyCreate", VP, VP)(world.wid)

    def __del__(self):
        if getattr(self, "bid", None) and getattr(self.world, "wid", None):      # a destroyed world has taken its bodies along (ode.cpp:1590-1623)
            self._lib.f("dBodyDestroy", None, VP)(self.bid)
            self.bid = None

    def _set3(self, name, v):
        r = _L().real
        _L().f(name, None, VP, r, r, r)(self.bid, v[0], v[1], v[2])

    def _get(self, name, n):
        p = _L().f(name, C.POINTER(_L().real), VP)(self.bid)
        return tuple(float(p[i]) for i in range(n))

    def setPosition(self, pos):
        self._set3("dBodySetPosition", pos)

    def getPosition(self):
        return self._get("dBodyGetPosition", 3)

    def setRotation(self, R):
        m = _L().mat3(R[0], R[1], R[2], 0, R[3], R[4], R[5], 0, R[6], R[7], R[8], 0)
        _L().f("dBodySetRotation", None, VP, VP)(self.bid, m)

    def getRotation(self):
        m = self._get("dBodyGetRotation", 12)
        return (m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10])

    def getQuaternion(self):
        return self._get("dBodyGetQuaternion", 4)

    def setQuaternion(self, q):
        _L().f("dBodySetQuaternion", None, VP, VP)(self.bid, _L().vec4(q[0], q[1], q[2], q[3]))

    def setLinearVel(self, vel):
        self._set3("dBodySetLinearVel", vel)

    def getLinearVel(self):
        return self._get("dBodyGetLinearVel", 3)

    def setAngularVel(self, vel):
        self._set3("dBodySetAngularVel", vel)

    def getAngularVel(self):
        return self._get("dBodyGetAngularVel", 3)

    def setMass(self, mass):
        _L().f("dBodySetMass", None, VP, VP)(self.bid, C.byref(mass._mass))

    def getMass(self):
        m = Mass()
        _L().f("dBodyGetMass", None, VP, VP)(self.bid, C.byref(m._mass))
        return m

    def addForce(self, f):
        self._set3("dBodyAddForce", f)

    def addTorque(self, t):
        self._set3("dBodyAddTorque", t)

    def addRelForce(self, f):
        self._set3("dBodyAddRelForce", f)

    def addRelTorque(self, t):
        self._set3("dBodyAddRelTorque", t)

    def getForce(self):
        return self._get("dBodyGetForce", 3)

    def getTorque(self):
        return self._get("dBodyGetTorque", 3)

    def setForce(self, f):
        self._set3("dBodySetForce", f)

    def setTorque(self, t):
        self._set3("dBodySetTorque", t)

    def vectorToWorld(self, v):
        R = self.getRotation()      # dBodyVectorToWorld ode.cpp:703-712
        return tuple(R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2] for i in range(3))

    def vectorFromWorld(self, v):
        R = self.getRotation()      # dBodyVectorFromWorld ode.cpp:715-724
        return tuple(R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2] for i in range(3))

    def getRelPointPos(self, p):
        q, o = self.vectorToWorld(p), self.getPosition()     # dBodyGetRelPointPos ode.cpp:640-652
        return (q[0] + o[0], q[1] + o[1], q[2] + o[2])

    def enable(self):
        _L().f("dBodyEnable", None, VP)(self.bid)

    def disable(self):
        _L().f("dBodyDisable", None, VP)(self.bid)

    def isEnabled(self):
        return _L().f("dBodyIsEnabled", I_, VP)(self.bid)

    def setFiniteRotationMode(self, mode):
        _L().f("dBodySetFiniteRotationMode", None, VP, I_)(self.bid, mode)

    def getFiniteRotationMode(self):
        return _L().f("dBodyGetFiniteRotationMode", I_, VP)(self.bid)

    def getNumJoints(self):
        return _L().f("dBodyGetNumJoints", I_, VP)(self.bid)

    def setGravityMode(self, mode):
        _L().f("dBodySetGravityMode", None, VP, I_)(self.bid, mode)

    def getGravityMode(self):
        return _L().f("dBodyGetGravityMode", I_, VP)(self.bid)

    def setDynamic(self):
        _L().f("dBodySetDynamic", None, VP)(self.bid)

    def setKinematic(self):
        _L().f("dBodySetKinematic", None, VP)(self.bid)

    def isKinematic(self):
        return _L().f("dBodyIsKinematic", I_, VP)(self.bid)

    def setMaxAngularSpeed(self, max_speed):
        _L().f("dBodySetMaxAngularSpeed", None, VP, _L().real)(self.bid, max_speed)


# ------------------------------------------------------------------------------------------------ joints (ode.pyx:1745-2920)
class JointGroup(object):
    def __init__(self):
        self._lib = _L()
        self.gid = self._lib.f("dJointGroupCreate", VP, I_)(0)
        self.jointlist = []

    def __del__(self):
        if getattr(self, "gid", None):
            for j in self.jointlist:
                j._destroyed()
            self._lib.f("dJointGroupDestroy", None, VP)(self.gid)
            self.gid = None

    def empty(self):
        _L().f("dJointGroupEmpty", None, VP)(self.gid)
        for j in self.jointlist:
            j._destroyed()
        self.jointlist = []

    def _addjoint(self, j):
        self.jointlist.append(j)


class Joint(object):
    _create = None
    _pfx = None

    def __init__(self, world, jointgroup=None):
        self.world, self.jointgroup = world, jointgroup
        self.body1 = self.body2 = None
        self.feedback = None
        self._lib = _L()
        self.jid = self._lib.f(self._create, VP, VP, VP)(world.wid, jointgroup.gid if jointgroup is not None else None)
        if jointgroup is not None:
            jointgroup._addjoint(self)

    def __del__(self):
        if getattr(self, "jid", None) and self.jointgroup is None and getattr(self.world, "wid", None):
            self._lib.f("dJointDestroy", None, VP)(self.jid)
            self.jid = None

    def _destroyed(self):
        self.jid = None

    def attach(self, body1, body2):
        self.body1, self.body2 = body1, body2
        _L().f("dJointAttach", None, VP, VP, VP)(self.jid, _bid(body1), _bid(body2))

    def getBody(self, index):
        if index == 0:
            return self.body1
        if index == 1:
            return self.body2
        raise IndexError()

    def setFeedback(self, flag=1):
        if flag:
            if self.feedback is None:
                self.feedback = _L().dJointFeedback()
            _L().f("dJointSetFeedback", None, VP, VP)(self.jid, C.byref(self.feedback))
        else:
            self.feedback = None
            _L().f("dJointSetFeedback", None, VP, VP)(self.jid, None)

    def getFeedback(self):
        fb = self.feedback
        if fb is None:
            return None
        return (_v3(fb.f1), _v3(fb.t1), _v3(fb.f2), _v3(fb.t2))

    # shared helpers
    def _set3(self, name, v):
        r = _L().real
        _L().f(name, None, VP, r, r, r)(self.jid, v[0], v[1], v[2])

    def _get3(self, name):
        out = _L().vec3()
        _L().f(name, None, VP, VP)(self.jid, out)
        return _v3(out)

    def _real(self, name):
        return float(_L().f(name, _L().real, VP)(self.jid))

    def setParam(self, param, value):
        _L().f("dJointSet%sParam" % self._pfx, None, VP, I_, _L().real)(self.jid, param, value)

    def getParam(self, param):
        return float(_L().f("dJointGet%sParam" % self._pfx, _L().real, VP, I_)(self.jid, param))


class BallJoint(Joint):
    _create, _pfx = "dJointCreateBall", "Ball"

    def setAnchor(self, pos):
        self._set3("dJointSetBallAnchor", pos)

    def getAnchor(self):
        return self._get3("dJointGetBallAnchor")

    def getAnchor2(self):
        return self._get3("dJointGetBallAnchor2")


class HingeJoint(Joint):
    _create, _pfx = "dJointCreateHinge", "Hinge"

    def setAnchor(self, pos):
        self._set3("dJointSetHingeAnchor", pos)

    def getAnchor(self):
        return self._get3("dJointGetHingeAnchor")

    def getAnchor2(self):
        return self._get3("dJointGetHingeAnchor2")

    def setAxis(self, axis):
        self._set3("dJointSetHingeAxis", axis)

    def getAxis(self):
        return self._get3("dJointGetHingeAxis")

    def getAngle(self):
        return self._real("dJointGetHingeAngle")

    def getAngleRate(self):
        return self._real("dJointGetHingeAngleRate")

    def addTorque(self, torque):
        _L().f("dJointAddHingeTorque", None, VP, _L().real)(self.jid, torque)


class SliderJoint(Joint):
    _create, _pfx = "dJointCreateSlider", "Slider"

    def setAxis(self, axis):
        self._set3("dJointSetSliderAxis", axis)

    def getAxis(self):
        return self._get3("dJointGetSliderAxis")

    def getPosition(self):
        return self._real("dJointGetSliderPosition")

    def getPositionRate(self):
        return self._real("dJointGetSliderPositionRate")

    def addForce(self, force):
        _L().f("dJointAddSliderForce", None, VP, _L().real)(self.jid, force)


class UniversalJoint(Joint):
    _create, _pfx = "dJointCreateUniversal", "Universal"

    def setAnchor(self, pos):
        self._set3("dJointSetUniversalAnchor", pos)

    def getAnchor(self):
        return self._get3("dJointGetUniversalAnchor")

    def getAnchor2(self):
        return self._get3("dJointGetUniversalAnchor2")

    def setAxis1(self, axis):
        self._set3("dJointSetUniversalAxis1", axis)

    def getAxis1(self):
        return self._get3("dJointGetUniversalAxis1")

    def setAxis2(self, axis):
        self._set3("dJointSetUniversalAxis2", axis)

    def getAxis2(self):
        return self._get3("dJointGetUniversalAxis2")

    def getAngle1(self):
        return self._real("dJointGetUniversalAngle1")

    def getAngle2(self):
        return self._real("dJointGetUniversalAngle2")

    def getAngle1Rate(self):
        return self._real("dJointGetUniversalAngle1Rate")

    def getAngle2Rate(self):
        return self._real("dJointGetUniversalAngle2Rate")

    def addTorques(self, torque1, torque2):
        r = _L().real
        _L().f("dJointAddUniversalTorques", None, VP, r, r)(self.jid, torque1, torque2)


class Hinge2Joint(Joint):
    _create, _pfx = "dJointCreateHinge2", "Hinge2"

    def setAnchor(self, pos):
        self._set3("dJointSetHinge2Anchor", pos)

    def getAnchor(self):
        return self._get3("dJointGetHinge2Anchor")

    def getAnchor2(self):
        return self._get3("dJointGetHinge2Anchor2")

    def setAxis1(self, axis):
        self._set3("dJointSetHinge2Axis1", axis)

    def getAxis1(self):
        return self._get3("dJointGetHinge2Axis1")

    def setAxis2(self, axis):
        self._set3("dJointSetHinge2Axis2", axis)

    def getAxis2(self):
        return self._get3("dJointGetHinge2Axis2")

    def getAngle1(self):
        return self._real("dJointGetHinge2Angle1")

    def getAngle1Rate(self):
        return self._real("dJointGetHinge2Angle1Rate")

    def getAngle2Rate(self):
        return self._real("dJointGetHinge2Angle2Rate")

    def addTorques(self, torque1, torque2):
        r = _L().real
        _L().f("dJointAddHinge2Torques", None, VP, r, r)(self.jid, torque1, torque2)


class FixedJoint(Joint):
    _create, _pfx = "dJointCreateFixed", "Fixed"

    def setFixed(self):
        _L().f("dJointSetFixed", None, VP)(self.jid)


class ContactJoint(Joint):
    """ode.pyx:2638-2658: dJointCreateContact(world, group, &contact._contact)"""

    def __init__(self, world, jointgroup, contact):
        self.world, self.jointgroup = world, jointgroup
        self.body1 = self.body2 = None
        self.feedback = None
        self._lib = _L()
        self.jid = self._lib.f("dJointCreateContact", VP, VP, VP, VP)(world.wid, jointgroup.gid if jointgroup is not None else None, C.byref(contact._contact))
        if jointgroup is not None:
            jointgroup._addjoint(self)


class AMotor(Joint):
    _create, _pfx = "dJointCreateAMotor", "AMotor"

    def setMode(self, mode):
        _L().f("dJointSetAMotorMode", None, VP, I_)(self.jid, mode)

    def getMode(self):
        return _L().f("dJointGetAMotorMode", I_, VP)(self.jid)

    def setNumAxes(self, num):
        _L().f("dJointSetAMotorNumAxes", None, VP, I_)(self.jid, num)

    def getNumAxes(self):
        return _L().f("dJointGetAMotorNumAxes", I_, VP)(self.jid)

    def setAxis(self, anum, rel, axis):
        r = _L().real
        _L().f("dJointSetAMotorAxis", None, VP, I_, I_, r, r, r)(self.jid, anum, rel, axis[0], axis[1], axis[2])

    def getAxis(self, anum):
        out = _L().vec3()
        _L().f("dJointGetAMotorAxis", None, VP, I_, VP)(self.jid, anum, out)
        return _v3(out)

    def getAxisRel(self, anum):
        return _L().f("dJointGetAMotorAxisRel", I_, VP, I_)(self.jid, anum)

    def setAngle(self, anum, angle):
        _L().f("dJointSetAMotorAngle", None, VP, I_, _L().real)(self.jid, anum, angle)

    def getAngle(self, anum):
        return float(_L().f("dJointGetAMotorAngle", _L().real, VP, I_)(self.jid, anum))

    def getAngleRate(self, anum):
        return float(_L().f("dJointGetAMotorAngleRate", _L().real, VP, I_)(self.jid, anum))

    def addTorques(self, torque0, torque1, torque2):
        r = _L().real
        _L().f("dJointAddAMotorTorques", None, VP, r, r, r)(self.jid, torque0, torque1, torque2)


class LMotor(Joint):
    _create, _pfx = "dJointCreateLMotor", "LMotor"

    def setNumAxes(self, num):
        _L().f("dJointSetLMotorNumAxes", None, VP, I_)(self.jid, num)

    def getNumAxes(self):
        return _L().f("dJointGetLMotorNumAxes", I_, VP)(self.jid)

    def setAxis(self, anum, rel, axis):
        r = _L().real
        _L().f("dJointSetLMotorAxis", None, VP, I_, I_, r, r, r)(self.jid, anum, rel, axis[0], axis[1], axis[2])

    def getAxis(self, anum):
        out = _L().vec3()
        _L().f("dJointGetLMotorAxis", None, VP, I_, VP)(self.jid, anum, out)
        return _v3(out)


# ------------------------------------------------------------------------------------------------ geoms and spaces (ode.pyx:3046-3990)
class GeomObject(object):
    def __init__(self, *a, **kw):
        raise NotImplementedError("GeomObject base class can't be used directly")

    def _register(self, gid, space):
        self.gid, self.space, self.body = gid, space, None
        _geom_c2py_lut[gid] = self
        if space is not None:
            space._geoms.append(self)       # the space keeps its geoms alive, like the reference's `self.space = space` + lut

    def _id(self):
        return self.gid

    def placeable(self):
        return True

    def setBody(self, body):
        _L().f("dGeomSetBody", None, VP, VP)(self.gid, _bid(body))
        self.body = body

    def getBody(self):
        return self.body

    def _set3(self, name, v):
        r = _L().real
        _L().f(name, None, VP, r, r, r)(self.gid, v[0], v[1], v[2])

    def _getp(self, name, n):
        p = _L().f(name, C.POINTER(_L().real), VP)(self.gid)
        return tuple(float(p[i]) for i in range(n))

    def setPosition(self, pos):
        self._set3("dGeomSetPosition", pos)

    def getPosition(self):
        return self._getp("dGeomGetPosition", 3)

    def setRotation(self, R):
        _L().f("dGeomSetRotation", None, VP, VP)(self.gid, _L().mat3(R[0], R[1], R[2], 0, R[3], R[4], R[5], 0, R[6], R[7], R[8], 0))

    def getRotation(self):
        m = self._getp("dGeomGetRotation", 12)
        return (m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10])

    def setQuaternion(self, q):
        _L().f("dGeomSetQuaternion", None, VP, VP)(self.gid, _L().vec4(q[0], q[1], q[2], q[3]))

    def setOffsetPosition(self, pos):
        self._set3("dGeomSetOffsetPosition", pos)

    def getOffsetPosition(self):
        return self._getp("dGeomGetOffsetPosition", 3)

    def setOffsetRotation(self, R):
        _L().f("dGeomSetOffsetRotation", None, VP, VP)(self.gid, _L().mat3(R[0], R[1], R[2], 0, R[3], R[4], R[5], 0, R[6], R[7], R[8], 0))

    def getOffsetRotation(self):
        m = self._getp("dGeomGetOffsetRotation", 12)
        return (m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10])

    def clearOffset(self):
        _L().f("dGeomClearOffset", None, VP)(self.gid)

    def getAABB(self):
        a = (_L().real * 6)()
        _L().f("dGeomGetAABB", None, VP, VP)(self.gid, a)
        return tuple(float(x) for x in a)

    def isSpace(self):
        return False

    def getSpace(self):
        return self.space

    def setCollideBits(self, bits):
        _L().f("dGeomSetCollideBits", None, VP, C.c_ulong)(self.gid, bits)

    def setCategoryBits(self, bits):
        _L().f("dGeomSetCategoryBits", None, VP, C.c_ulong)(self.gid, bits)

    def getCollideBits(self):
        return _L().f("dGeomGetCollideBits", C.c_ulong, VP)(self.gid)

    def getCategoryBits(self):
        return _L().f("dGeomGetCategoryBits", C.c_ulong, VP)(self.gid)

    def enable(self):
        _L().f("dGeomEnable", None, VP)(self.gid)

    def disable(self):
        _L().f("dGeomDisable", None, VP)(self.gid)

    def isEnabled(self):
        return _L().f("dGeomIsEnabled", I_, VP)(self.gid)


def _sid(space):
    return space.sid if space is not None else None


class GeomSphere(GeomObject):
    def __init__(self, space=None, radius=1.0):
        self._register(_L().f("dCreateSphere", VP, VP, _L().real)(_sid(space), radius), space)

    def getRadius(self):
        return float(_L().f("dGeomSphereGetRadius", _L().real, VP)(self.gid))

    def setRadius(self, radius):
        _L().f("dGeomSphereSetRadius", None, VP, _L().real)(self.gid, radius)


class GeomBox(GeomObject):
    def __init__(self, space=None, lengths=(1.0, 1.0, 1.0)):
        r = _L().real
        self._register(_L().f("dCreateBox", VP, VP, r, r, r)(_sid(space), lengths[0], lengths[1], lengths[2]), space)

    def getLengths(self):
        out = _L().vec3()
        _L().f("dGeomBoxGetLengths", None, VP, VP)(self.gid, out)
        return _v3(out)

    def setLengths(self, lengths):
        self._set3("dGeomBoxSetLengths", lengths)


class GeomPlane(GeomObject):
    def __init__(self, space=None, normal=(0, 0, 1), dist=0):
        r = _L().real
        self._register(_L().f("dCreatePlane", VP, VP, r, r, r, r)(_sid(space), normal[0], normal[1], normal[2], dist), space)

    def placeable(self):
        return False

    def getParams(self):
        out = _L().vec4()
        _L().f("dGeomPlaneGetParams", None, VP, VP)(self.gid, out)
        return (_v3(out), float(out[3]))

    def setParams(self, normal, dist):
        r = _L().real
        _L().f("dGeomPlaneSetParams", None, VP, r, r, r, r)(self.gid, normal[0], normal[1], normal[2], dist)


class GeomCapsule(GeomObject):
    def __init__(self, space=None, radius=0.5, length=1.0):
        r = _L().real
        self._register(_L().f("dCreateCapsule", VP, VP, r, r)(_sid(space), radius, length), space)

    def getParams(self):
        r = _L().real
        a, b = r(), r()
        _L().f("dGeomCapsuleGetParams", None, VP, VP, VP)(self.gid, C.byref(a), C.byref(b))
        return (a.value, b.value)

    def setParams(self, radius, length):
        r = _L().real
        _L().f("dGeomCapsuleSetParams", None, VP, r, r)(self.gid, radius, length)


GeomCCylinder = GeomCapsule      # ode.pyx keeps the old name


class GeomCylinder(GeomObject):
    """ode.pyx:3994-4040"""

    def __init__(self, space=None, radius=0.5, length=1.0):
        r = _L().real
        self._register(_L().f("dCreateCylinder", VP, VP, r, r)(_sid(space), radius, length), space)

    def getParams(self):
        r = _L().real
        a, b = r(), r()
        _L().f("dGeomCylinderGetParams", None, VP, VP, VP)(self.gid, C.byref(a), C.byref(b))
        return (a.value, b.value)

    def setParams(self, radius, length):
        r = _L().real
        _L().f("dGeomCylinderSetParams", None, VP, r, r)(self.gid, radius, length)


class GeomRay(GeomObject):
    """ode.pyx:4043-4122: a ray of length rlen along its local z axis; collide(ray, geom) returns at most one Contact whose depth is
    the distance from the ray's start to the hit"""

    def __init__(self, space=None, rlen=1.0):
        self._register(_L().f("dCreateRay", VP, VP, _L().real)(_sid(space), rlen), space)

    def setLength(self, rlen):
        _L().f("dGeomRaySetLength", None, VP, _L().real)(self.gid, rlen)

    def getLength(self):
        return float(_L().f("dGeomRayGetLength", _L().real, VP)(self.gid))

    def set(self, p, u):
        r = _L().real
        _L().f("dGeomRaySet", None, VP, r, r, r, r, r, r)(self.gid, p[0], p[1], p[2], u[0], u[1], u[2])

    def get(self):
        p, u = _L().vec3(), _L().vec3()
        _L().f("dGeomRayGet", None, VP, VP, VP)(self.gid, p, u)
        return (_v3(p), _v3(u))


class SpaceBase(GeomObject):
    _create = None

    def __init__(self, space=None):
        if space is not None:
            raise NotImplementedError("nested spaces are outside the supported subset")
        self._geoms = []
        self._lib = _L()
        self.sid = self.gid = self._make()
        self.space = self.body = None

    def _make(self):
        return _L().f(self._create, VP, VP)(None)

    def __del__(self):
        if getattr(self, "sid", None):
            self._lib.f("dSpaceDestroy", None, VP)(self.sid)       # cleanup mode 1: the space destroys its geoms (collision_space.cpp:95-113)
            self.sid = None

    def _id(self):
        return self.sid

    def isSpace(self):
        return True

    def placeable(self):
        return False

    def __len__(self):
        return self.getNumGeoms()

    def __iter__(self):
        return iter([self.getGeom(i) for i in range(self.getNumGeoms())])

    def add(self, geom):
        _L().f("dSpaceAdd", None, VP, VP)(self.sid, geom._id())
        self._geoms.append(geom)
        geom.space = self

    def remove(self, geom):
        _L().f("dSpaceRemove", None, VP, VP)(self.sid, geom._id())
        self._geoms = [g for g in self._geoms if g is not geom]
        geom.space = None

    def query(self, geom):
        return any(g is geom for g in self._geoms)

    def getNumGeoms(self):
        return _L().f("dSpaceGetNumGeoms", I_, VP)(self.sid)

    def getGeom(self, idx):
        gid = _L().f("dSpaceGetGeom", VP, VP, I_)(self.sid, idx)
        if gid is None:
            raise IndexError("geom index out of range")
        return _geom_c2py_lut[gid]

    def collide(self, arg, callback):
        """ode.pyx:3536-3559: callback(arg, geom1, geom2) for every potentially intersecting pair (dSpaceCollide)"""
        err = []

        def tramp(data, o1, o2):
            if err:
                return
            try:
                callback(arg, _geom_c2py_lut[o1], _geom_c2py_lut[o2])
            except BaseException as e:      # a ctypes callback cannot propagate: re-raised after dSpaceCollide returns
                err.append(e)
        cb = _L().NearCallback(tramp)
        _L().f("dSpaceCollide", None, VP, VP, _L().NearCallback)(self.sid, None, cb)
        if err:
            raise err[0]


class SimpleSpace(SpaceBase):
    _create = "dSimpleSpaceCreate"


class HashSpace(SpaceBase):
    _create = "dHashSpaceCreate"

    def setLevels(self, minlevel, maxlevel):
        if minlevel > maxlevel:
            raise ValueError("minlevel (%d) must be less than or equal to maxlevel (%d)" % (minlevel, maxlevel))
        _L().f("dHashSpaceSetLevels", None, VP, I_, I_)(self.sid, minlevel, maxlevel)

    def getLevels(self):
        a, b = I_(), I_()
        _L().f("dHashSpaceGetLevels", None, VP, VP, VP)(self.sid, C.byref(a), C.byref(b))
        return (a.value, b.value)


class SweepAndPruneSpace(SpaceBase):
    """dSweepAndPruneSpaceCreate(0, dSAP_AXES_XYZ) (include/ode/collision_space.h:59-64); not in ode.pyx, which predates it"""

    def _make(self):
        return _L().f("dSweepAndPruneSpaceCreate", VP, VP, I_)(None, (0) | (1 << 2) | (2 << 4))


def Space(space_type=0):
    """ode.pyx:3722-3740: 0 = SimpleSpace, 1 = HashSpace"""
    if space_type == 0:
        return SimpleSpace()
    if space_type == 1:
        return HashSpace()
    raise ValueError("Unknown space type (%d)" % space_type)


# ------------------------------------------------------------------------------------------------ module functions (ode.pyx:4392-4508)
def collide(geom1, geom2):
    """collide(geom1, geom2) -> list of Contact (dCollide with room for 150 contact geoms, ode.pyx:4392-4437)"""
    L = _L()
    c = (L.dContactGeom * 150)()
    n = L.f("dCollide", I_, VP, VP, I_, VP, I_)(geom1._id(), geom2._id(), 150, c, C.sizeof(L.dContactGeom))
    res = []
    for i in range(n):
        cont = Contact()
        C.memmove(C.byref(cont._contact.geom), C.byref(c[i]), C.sizeof(L.dContactGeom))
        res.append(cont)
    return res


def areConnected(body1, body2):
    if body1 is environment or body2 is environment:
        return False
    return bool(_L().f("dAreConnected", I_, VP, VP)(body1.bid, body2.bid))


def randSetSeed(seed):
    """dRandSetSeed (include/ode/misc.h; ode/src/misc.cpp:52-61).  Not part of ode.pyx: added because QuickStep's constraint reordering
    draws from this process-global generator, so a reproducible run has to seed it."""
    _L().f("dRandSetSeed", None, C.c_ulong)(seed)


def randGetSeed():
    return int(_L().f("dRandGetSeed", C.c_ulong)())


def CloseODE():
    _L().f("dCloseODE", None)()


def InitODE():
    _lib.f("dInitODE", None)()
