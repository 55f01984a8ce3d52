"""ctypes mirror of include/ode_b200.h.

`SceneLib` binds one shared library exporting the scene-description C interface under a symbol
prefix.  The product library uses the prefix ``odeb_`` (ode_b200/csrc, CUDA).  The same class is
reused by tests/ and bench.py's cpu_baseline leg to drive the checkers under oracle/ (prefixes
``ref_`` and ``orc_``) -- the package itself never opens anything under oracle/.
"""
import ctypes as C
import numpy as np

c_double3 = C.c_double * 3


class OdebWorldParams(C.Structure):
    _fields_ = [
        ("gravity", C.c_double * 3), ("erp", C.c_double), ("cfm", C.c_double),
        ("num_iterations", C.c_int), ("sor_w", C.c_double),
        ("premature_exit_delta", C.c_double), ("max_extra_factor", C.c_double), ("extra_iter_delta", C.c_double),
        ("contact_max_vel", C.c_double), ("contact_surface_layer", C.c_double),
        ("auto_disable", C.c_int), ("adis_linear_thr", C.c_double), ("adis_angular_thr", C.c_double),
        ("adis_steps", C.c_int), ("adis_time", C.c_double), ("adis_samples", C.c_int),
        ("linear_damping", C.c_double), ("angular_damping", C.c_double),
        ("linear_damping_thr", C.c_double), ("angular_damping_thr", C.c_double),
        ("max_angular_speed", C.c_double),
        ("space_type", C.c_int), ("max_contacts", C.c_int), ("skip_connected", C.c_int), ("surf_mode", C.c_int),
        ("mu", C.c_double), ("mu2", C.c_double), ("bounce", C.c_double), ("bounce_vel", C.c_double),
        ("soft_erp", C.c_double), ("soft_cfm", C.c_double),
        ("motion1", C.c_double), ("motion2", C.c_double), ("motionN", C.c_double),
        ("slip1", C.c_double), ("slip2", C.c_double),
        ("rho", C.c_double), ("rho2", C.c_double), ("rhoN", C.c_double),
        ("hash_levels_set", C.c_int), ("hash_minlevel", C.c_int), ("hash_maxlevel", C.c_int),
        ("max_pairs", C.c_int), ("max_contacts_per_world", C.c_int),
    ]


class OdebBodyDesc(C.Structure):
    _fields_ = [("mass", C.c_double), ("inertia", C.c_double * 9), ("flags", C.c_int)]


class OdebGeomDesc(C.Structure):
    _fields_ = [("type", C.c_int), ("body", C.c_int), ("p", C.c_double * 4),
                ("category_bits", C.c_uint32), ("collide_bits", C.c_uint32),
                ("has_offset", C.c_int), ("offset_pos", C.c_double * 3), ("offset_quat", C.c_double * 4)]


class OdebJointDesc(C.Structure):
    _fields_ = [("type", C.c_int), ("body1", C.c_int), ("body2", C.c_int),
                ("anchor", C.c_double * 3), ("axis1", C.c_double * 3), ("axis2", C.c_double * 3),
                ("lo_stop", C.c_double * 3), ("hi_stop", C.c_double * 3),
                ("vel", C.c_double * 3), ("fmax", C.c_double * 3),
                ("fudge_factor", C.c_double * 3), ("bounce", C.c_double * 3),
                ("stop_erp", C.c_double * 3), ("stop_cfm", C.c_double * 3),
                ("susp_erp", C.c_double), ("susp_cfm", C.c_double),
                ("motor_num", C.c_int), ("motor_mode", C.c_int), ("motor_rel", C.c_int * 3),
                ("motor_axis", (C.c_double * 3) * 3), ("motor_angle", C.c_double * 3)]


class OdebStats(C.Structure):
    _fields_ = [("v", C.c_uint32 * 4)]


SPHERE, BOX, CAPSULE, PLANE = 0, 1, 2, 4
CYLINDER, RAY = 3, 5
JOINT_BALL, JOINT_HINGE, JOINT_SLIDER, JOINT_CONTACT, JOINT_UNIVERSAL, JOINT_HINGE2, JOINT_FIXED = 1, 2, 3, 4, 5, 6, 7
JOINT_AMOTOR, JOINT_LMOTOR = 9, 10
AMOTOR_USER, AMOTOR_EULER = 0, 1
SPACE_HASH, SPACE_SAP, SPACE_SIMPLE = 0, 1, 2
CONTACT_MU2, CONTACT_BOUNCE, CONTACT_SOFT_ERP, CONTACT_SOFT_CFM = 0x001, 0x004, 0x008, 0x010
CONTACT_MOTION1, CONTACT_MOTION2, CONTACT_MOTIONN = 0x020, 0x040, 0x080
CONTACT_SLIP1, CONTACT_SLIP2, CONTACT_ROLLING, CONTACT_APPROX1 = 0x100, 0x200, 0x400, 0x7000
BODY_NO_GRAVITY, BODY_NO_GYRO, BODY_DISABLED, BODY_FINITE_ROTATION, BODY_KINEMATIC = 1, 2, 4, 8, 16
INF = float("inf")


def default_world_params(**kw):
    """World defaults of the reference (ode/src/objects.cpp:37-121, include/ode/objects.h:451-475)."""
    p = OdebWorldParams()
    p.gravity[:] = (0.0, 0.0, 0.0)
    p.erp, p.cfm = 0.2, -1.0
    p.num_iterations, p.sor_w = 20, 1.3
    p.premature_exit_delta, p.max_extra_factor, p.extra_iter_delta = 1e-8, 1.0, 1e-2
    p.contact_max_vel, p.contact_surface_layer = INF, 0.0
    p.auto_disable, p.adis_linear_thr, p.adis_angular_thr = 0, 0.01, 0.01
    p.adis_steps, p.adis_time, p.adis_samples = 10, 0.0, 1
    p.linear_damping = p.angular_damping = 0.0
    p.linear_damping_thr = p.angular_damping_thr = 0.01
    p.max_angular_speed = INF
    p.space_type, p.max_contacts, p.skip_connected, p.surf_mode = SPACE_HASH, 4, 1, 0
    p.mu, p.mu2 = INF, 0.0
    for k, v in kw.items():
        if k == "gravity":
            p.gravity[:] = v
        else:
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
    return p


class Scene:
    """A template world + per-world initial state; consumed identically by every implementation."""

    def __init__(self, wp, nworlds=1):
        self.wp = wp
        self.nworlds = nworlds
        self.bodies, self.body_pos, self.body_quat = [], [], []
        self.geoms, self.joints = [], []
        self.state = None  # optional dict(pos, quat, lvel, avel) of [W][NB][k] float64
        self.seeds = None

    def add_body(self, mass, inertia, pos, quat=(1, 0, 0, 0), flags=0):
        d = OdebBodyDesc()
        d.mass = mass
        d.inertia[:] = np.asarray(inertia, dtype=np.float64).reshape(9)
        d.flags = flags
        self.bodies.append(d)
        self.body_pos.append(tuple(float(x) for x in pos))
        self.body_quat.append(tuple(float(x) for x in quat))
        return len(self.bodies) - 1

    def add_geom(self, gtype, params, body=-1, category=0xFFFFFFFF, collide=0xFFFFFFFF, offset_pos=None, offset_quat=None):
        g = OdebGeomDesc()
        g.type, g.body = gtype, body
        if offset_pos is not None or offset_quat is not None:
            g.has_offset = 1
            g.offset_pos[:] = offset_pos if offset_pos is not None else (0, 0, 0)
            g.offset_quat[:] = offset_quat if offset_quat is not None else (1, 0, 0, 0)
        pp = list(params) + [0.0] * (4 - len(params))
        g.p[:] = pp
        g.category_bits, g.collide_bits = category, collide
        self.geoms.append(g)
        return len(self.geoms) - 1

    def add_joint(self, jtype, body1, body2, anchor=(0, 0, 0), axis1=(1, 0, 0), axis2=(0, 1, 0),
                  lo_stop=(-INF, -INF), hi_stop=(INF, INF), vel=(0, 0), fmax=(0, 0),
                  fudge_factor=(-1, -1), bounce=(-1, -1), stop_erp=(-1, -1), stop_cfm=(-1, -1), susp_erp=-1.0, susp_cfm=-1.0,
                  motor_axes=(), motor_mode=0, motor_angle=(0, 0, 0)):
        """motor_axes (LMotor / AMotor): up to three (rel, (x, y, z)) entries; the per-axis limit-motor tuples then take up to 3 values."""
        def pad(v, fill):
            v = tuple(v)
            return v + (fill,) * (3 - len(v))
        j = OdebJointDesc()
        j.type, j.body1, j.body2 = jtype, body1, body2
        j.anchor[:] = anchor
        j.axis1[:] = axis1
        j.axis2[:] = axis2
        j.lo_stop[:] = pad(lo_stop, -INF)
        j.hi_stop[:] = pad(hi_stop, INF)
        j.vel[:] = pad(vel, 0)
        j.fmax[:] = pad(fmax, 0)
        j.fudge_factor[:] = pad(fudge_factor, -1)
        j.bounce[:] = pad(bounce, -1)
        j.stop_erp[:] = pad(stop_erp, -1)
        j.stop_cfm[:] = pad(stop_cfm, -1)
        j.susp_erp, j.susp_cfm = susp_erp, susp_cfm
        j.motor_num, j.motor_mode = len(motor_axes), motor_mode
        for k, (rel, ax) in enumerate(motor_axes):
            j.motor_rel[k] = rel
            j.motor_axis[k][:] = ax
        j.motor_angle[:] = motor_angle
        self.joints.append(j)
        return len(self.joints) - 1

    @property
    def nbody(self):
        return len(self.bodies)

    @property
    def ngeom(self):
        return len(self.geoms)


def box_mass(density, lx, ly, lz):
    m = density * lx * ly * lz
    return m, np.diag([m / 12.0 * (ly * ly + lz * lz), m / 12.0 * (lx * lx + lz * lz), m / 12.0 * (lx * lx + ly * ly)])


def sphere_mass(density, r):
    m = 4.0 / 3.0 * np.pi * r ** 3 * density
    return m, np.eye(3) * (0.4 * m * r * r)


def cylinder_mass(density, r, length):
    """dMassSetCylinder along z (ode/src/mass.cpp:173-198), direction = 3."""
    m = np.pi * r * r * length * density
    Ia = m * (0.25 * r * r + length * length / 12.0)
    return m, np.diag([Ia, Ia, m * 0.5 * r * r])


def capsule_mass(density, r, length):
    """dMassSetCapsule along z (ode/src/mass.cpp:138-166), direction = 3."""
    M1 = np.pi * r * r * length * density
    M2 = 4.0 / 3.0 * np.pi * r ** 3 * density
    m = M1 + M2
    Ia = M1 * (0.25 * r * r + length * length / 12.0) + M2 * (0.4 * r * r + 0.375 * r * length + 0.25 * length * length)
    Ib = (M1 * 0.5 + M2 * 0.4) * r * r
    return m, np.diag([Ia, Ia, Ib])


class SceneLib:
    """One loaded library + symbol prefix."""

    def __init__(self, path, prefix, real):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.real = np.dtype(real)
        self.path = path
        f = self._fn("create")
        f.restype = C.c_void_p
        f.argtypes = [C.POINTER(OdebWorldParams), C.c_int, C.POINTER(OdebBodyDesc), C.POINTER(C.c_double),
                      C.POINTER(C.c_double), C.c_int, C.POINTER(OdebGeomDesc), C.c_int, C.POINTER(OdebJointDesc),
                      C.c_int, C.c_int]
        self._fn("destroy").argtypes = [C.c_void_p]
        self._fn("destroy").restype = None
        for name in ("set_state", "get_state"):
            self._fn(name).argtypes = [C.c_void_p] + [C.c_void_p] * 4
        self._fn("add_force").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        for name in ("set_seeds", "get_seeds", "get_enabled"):
            self._fn(name).argtypes = [C.c_void_p, C.c_void_p]
        self._fn("step").argtypes = [C.c_void_p, C.c_double, C.c_int]
        self._fn("get_pairs").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self._fn("get_contacts").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        if hasattr(self.lib, prefix + "get_ray_hits"):
            self._fn("get_ray_hits").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        self._fn("get_islands").argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self._fn("get_stats").argtypes = [C.c_void_p, C.c_int, C.POINTER(OdebStats)]

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def has(self, name):
        return hasattr(self.lib, self.prefix + name)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Batch:
    """W worlds built from a Scene on one implementation."""

    def __init__(self, slib, scene, device=0):
        self.slib, self.scene = slib, scene
        nb, ng, nj = scene.nbody, scene.ngeom, len(scene.joints)
        bodies = (OdebBodyDesc * max(nb, 1))(*scene.bodies)
        geoms = (OdebGeomDesc * max(ng, 1))(*scene.geoms)
        joints = (OdebJointDesc * max(nj, 1))(*scene.joints)
        pos = np.ascontiguousarray(np.asarray(scene.body_pos, dtype=np.float64).reshape(-1))
        quat = np.ascontiguousarray(np.asarray(scene.body_quat, dtype=np.float64).reshape(-1))
        self._keep = (bodies, geoms, joints, pos, quat)
        self.h = slib._fn("create")(C.byref(scene.wp), nb, bodies, pos.ctypes.data_as(C.POINTER(C.c_double)),
                                    quat.ctypes.data_as(C.POINTER(C.c_double)), ng, geoms, nj, joints,
                                    scene.nworlds, device)
        if not self.h:
            msg = ""
            if slib.has("last_error"):
                fn = slib._fn("last_error")
                fn.restype = C.c_char_p
                msg = (fn() or b"").decode()
            raise RuntimeError("%screate failed: %s" % (slib.prefix, msg))
        self.W, self.NB = scene.nworlds, nb
        if scene.state is not None:
            self.set_state(**scene.state)
        if scene.seeds is not None:
            self.set_seeds(scene.seeds)

    def close(self):
        if self.h:
            self.slib._fn("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _arr(self, a, k):
        if a is None:
            return None
        a = np.ascontiguousarray(np.asarray(a, dtype=self.slib.real).reshape(self.W, self.NB, k))
        return a

    def set_state(self, pos=None, quat=None, lvel=None, avel=None):
        arrs = [self._arr(pos, 3), self._arr(quat, 4), self._arr(lvel, 3), self._arr(avel, 3)]
        self.slib._fn("set_state")(self.h, *[_ptr(a) for a in arrs])

    def alloc_host(self, shape):
        """A numpy array of the library's real type in page-locked host memory (odeb_alloc_host): state / force arrays of this kind move
        to and from the device without the staging copy.  Libraries without the entry point (oracle, reference driver) get np.empty."""
        r = self.slib.real
        if not hasattr(self.slib.lib, self.slib.prefix + "alloc_host"):
            return np.empty(shape, r)
        f = self.slib._fn("alloc_host")
        f.restype, f.argtypes = C.c_void_p, [C.c_size_t]
        n = int(np.prod(shape))
        p = f(n * r.itemsize)
        if not p:
            raise RuntimeError("%salloc_host failed" % self.slib.prefix)
        self._pinned = getattr(self, "_pinned", []) + [p]
        ct = C.c_float if r == np.float32 else C.c_double
        return np.ctypeslib.as_array((ct * n).from_address(p)).reshape(shape)

    def alloc_state(self):
        """observation buffers for get_state(out=...) in page-locked memory"""
        return dict(pos=self.alloc_host((self.W, self.NB, 3)), quat=self.alloc_host((self.W, self.NB, 4)),
                    lvel=self.alloc_host((self.W, self.NB, 3)), avel=self.alloc_host((self.W, self.NB, 3)))

    def get_state(self, out=None):
        """Body state as host arrays [W, NB, 3|4]. `out`: a dict returned by an earlier call, filled in place (the
        steady-state loop of an application that owns its observation buffers)."""
        r = self.slib.real
        if out is None:
            out = dict(pos=np.empty((self.W, self.NB, 3), r), quat=np.empty((self.W, self.NB, 4), r),
                       lvel=np.empty((self.W, self.NB, 3), r), avel=np.empty((self.W, self.NB, 3), r))
        f = self.slib._fn("get_state")
        f.restype = C.c_int
        if not f(self.h, _ptr(out["pos"]), _ptr(out["quat"]), _ptr(out["lvel"]), _ptr(out["avel"])):
            msg = ""
            if self.slib.has("last_error"):
                fn = self.slib._fn("last_error")
                fn.restype = C.c_char_p
                msg = (fn() or b"").decode()
            raise RuntimeError("%sget_state failed: %s" % (self.slib.prefix, msg))
        return out

    def add_force(self, force=None, torque=None):
        self.slib._fn("add_force")(self.h, _ptr(self._arr(force, 3)), _ptr(self._arr(torque, 3)))

    def set_solver_mode(self, mode):
        """0 = replay of the reference's order (default), 1 = canonical order (large-world path, include/ode_b200.h)"""
        f = self.slib._fn("set_solver_mode")
        f.argtypes = [C.c_void_p, C.c_int]
        if not f(self.h, int(mode)):
            raise RuntimeError("%sset_solver_mode failed" % self.slib.prefix)

    def enable_feedback(self, on=True):
        f = self.slib._fn("enable_feedback")
        f.argtypes = [C.c_void_p, C.c_int]
        if not f(self.h, int(bool(on))):
            raise RuntimeError("%senable_feedback failed" % self.slib.prefix)

    def get_feedback(self, world, cap=4096):
        """(feedback [n,12] = f1 t1 f2 t2, state [n]) of the last step for one world; joints in id order
        (permanent joints, then the step's contact joints in creation order)"""
        f = self.slib._fn("get_feedback")
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        out = np.zeros((cap, 12), self.slib.real)
        st = np.zeros(cap, np.int32)
        n = f(self.h, int(world), out.ctypes.data, st.ctypes.data, cap)
        if n < 0:
            raise RuntimeError("%sget_feedback failed" % self.slib.prefix)
        n = min(n, cap)
        return out[:n].copy(), st[:n].copy()

    def snapshot(self):
        f = self.slib._fn("snapshot_size")
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p]
        n = f(self.h)
        buf = np.empty(n, np.uint8)
        g = self.slib._fn("snapshot")
        g.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        if not g(self.h, buf.ctypes.data, n):
            raise RuntimeError("%ssnapshot failed" % self.slib.prefix)
        return buf

    def restore(self, buf):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        g = self.slib._fn("restore")
        g.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        if not g(self.h, buf.ctypes.data, buf.size):
            raise RuntimeError("%srestore failed" % self.slib.prefix)

    def set_seeds(self, seeds):
        s = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint32).reshape(self.W))
        self.slib._fn("set_seeds")(self.h, _ptr(s))

    def get_seeds(self):
        s = np.empty(self.W, np.uint32)
        self.slib._fn("get_seeds")(self.h, _ptr(s))
        return s

    def get_enabled(self):
        e = np.empty((self.W, self.NB), np.int32)
        self.slib._fn("get_enabled")(self.h, _ptr(e))
        return e

    def step(self, h, nsteps=1):
        ok = self.slib._fn("step")(self.h, float(h), int(nsteps))
        if not ok:
            msg = ""
            if self.slib.has("last_error"):
                fn = self.slib._fn("last_error")
                fn.restype = C.c_char_p
                msg = ": " + (fn() or b"").decode()
            raise RuntimeError("%sstep failed%s" % (self.slib.prefix, msg))

    def step_async(self, h, nsteps=1):
        """odeb_step_async: the steps are queued on the batch's stream; the next blocking call (get_state, sync, ...) waits for them.
        Libraries without the entry point (oracle, reference driver) step synchronously."""
        if not self.slib.has("step_async"):
            return self.step(h, nsteps)
        f = self.slib._fn("step_async")
        f.restype, f.argtypes = C.c_int, [C.c_void_p, C.c_double, C.c_int]
        if not f(self.h, float(h), int(nsteps)):
            raise RuntimeError("%sstep_async failed" % self.slib.prefix)

    def sync(self):
        """odeb_sync: waits for the queued steps; raises if one of them overflowed a device capacity"""
        if not self.slib.has("sync"):
            return
        f = self.slib._fn("sync")
        f.restype, f.argtypes = C.c_int, [C.c_void_p]
        if not f(self.h):
            fn = self.slib._fn("last_error")
            fn.restype = C.c_char_p
            raise RuntimeError("%ssync failed: %s" % (self.slib.prefix, (fn() or b"").decode()))

    def get_totals(self):
        """[pairs, contacts, rows, islands, sweeps, row-sweeps] of the most recent step, summed over worlds"""
        out = (C.c_uint64 * 6)()
        f = self.slib._fn("get_totals")
        f.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        if not f(self.h, out):
            raise RuntimeError("%sget_totals failed" % self.slib.prefix)
        return [int(v) for v in out]

    def get_pairs(self, world, cap=1 << 16):
        buf = np.empty((cap, 2), np.int32)
        n = self.slib._fn("get_pairs")(self.h, world, _ptr(buf), cap)
        if n > cap:
            return self.get_pairs(world, n)
        return buf[:n].copy()

    def get_contacts(self, world, cap=1 << 14):
        g = np.empty((cap, 7), self.slib.real)
        ids = np.empty((cap, 2), np.int32)
        n = self.slib._fn("get_contacts")(self.h, world, _ptr(g), _ptr(ids), cap)
        if n > cap:
            return self.get_contacts(world, n)
        return g[:n].copy(), ids[:n].copy()

    def get_ray_hits(self, world, cap=1 << 12):
        """hits of the world's ray geoms in the last step's collide pass: ([pos3, normal3, depth = distance along the ray], (g1, g2))"""
        g = np.empty((cap, 7), self.slib.real)
        ids = np.empty((cap, 2), np.int32)
        n = self.slib._fn("get_ray_hits")(self.h, world, _ptr(g), _ptr(ids), cap)
        if n > cap:
            return self.get_ray_hits(world, n)
        return g[:n].copy(), ids[:n].copy()

    def get_ray_ranges(self):
        """(range[W, nray], hit_geom[W, nray]): nearest hit of every ray geom of every world in the last collide pass (inf / -1: none)"""
        f = self.slib._fn("num_rays")
        f.argtypes = [C.c_void_p]
        nray = f(self.h)
        rng = np.empty((self.W, nray), self.slib.real)
        hit = np.empty((self.W, nray), np.int32)
        g = self.slib._fn("get_ray_ranges")
        g.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        g(self.h, _ptr(rng), _ptr(hit))
        return rng, hit

    def get_islands(self, world):
        lab = np.empty(self.NB, np.int32)
        n = self.slib._fn("get_islands")(self.h, world, _ptr(lab))
        return n, lab

    def get_stats(self, world):
        s = OdebStats()
        self.slib._fn("get_stats")(self.h, world, C.byref(s))
        return np.array(list(s.v), dtype=np.uint32)
