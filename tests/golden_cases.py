"""Deterministic inputs of the golden-vector fixtures (shared by oracle/make_golden.py, which records the
reference's outputs, and the tests, which replay them on the oracle / the CUDA path)."""
import ctypes as C
import numpy as np
from ode_b200 import _binding as B
from ode_b200 import scenes

SPHERE, BOX, CAPSULE, PLANE = B.SPHERE, B.BOX, B.CAPSULE, B.PLANE
PAIR_TYPES = [(SPHERE, SPHERE), (SPHERE, BOX), (BOX, SPHERE), (SPHERE, PLANE), (PLANE, SPHERE), (BOX, BOX), (BOX, PLANE),
              (PLANE, BOX), (CAPSULE, SPHERE), (SPHERE, CAPSULE), (CAPSULE, BOX), (BOX, CAPSULE), (CAPSULE, CAPSULE),
              (CAPSULE, PLANE), (PLANE, CAPSULE)]


def _rand_rot(r):
    q = r.randn(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 0],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x), 0],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y), 0]])


def _geom(r, t):
    if t == SPHERE:
        return [0.3 + 0.4 * r.rand(), 0, 0, 0]
    if t == BOX:
        return list(0.4 + 0.8 * r.rand(3)) + [0]
    if t == CAPSULE:
        return [0.15 + 0.2 * r.rand(), 0.3 + 0.8 * r.rand(), 0, 0]
    n = r.randn(3)
    return list(n) + [0.2 * r.randn()]


def collide_cases(real, per_type=24, seed=4242):
    """Random overlapping-ish configurations for every supported ordered pair of geom classes, with
    max-contact flags 1..8, plus axis-aligned / parallel special cases (face-face boxes, parallel capsules)."""
    r = np.random.RandomState(seed)
    cases = []
    for (t1, t2) in PAIR_TYPES:
        for k in range(per_type):
            p1, p2 = _geom(r, t1), _geom(r, t2)
            aligned = (k % 6 == 5)
            R1 = np.eye(3, 4) if aligned else _rand_rot(r)
            R2 = np.eye(3, 4) if aligned else _rand_rot(r)
            pos1 = 0.3 * r.randn(3)
            pos2 = pos1 + (0.9 if k % 3 else 0.45) * r.randn(3) * np.array([1, 1, 0.6])
            if PLANE in (t1, t2):
                pos1 = 0.4 * r.randn(3)
                pos2 = 0.4 * r.randn(3)
            flags = int(1 + (k % 8))
            cases.append(dict(t1=t1, p1=np.asarray(p1, real), pos1=np.asarray(pos1, real), R1=np.asarray(R1, real).reshape(12),
                              t2=t2, p2=np.asarray(p2, real), pos2=np.asarray(pos2, real), R2=np.asarray(R2, real).reshape(12),
                              flags=flags))
    return cases


def run_collide(slib, prefix, cases):
    fn = getattr(slib.lib, prefix + "collide_pair")
    fa = getattr(slib.lib, prefix + "geom_aabb")
    fa.restype = None
    real = slib.real
    ns, g7s, aabbs = [], [], []
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    fa.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    for c in cases:
        out = np.zeros((8, 7), real)
        n = fn(c["t1"], P(c["p1"]), P(c["pos1"]), P(c["R1"]), c["t2"], P(c["p2"]), P(c["pos2"]), P(c["R2"]), c["flags"], P(out), 8)
        ns.append(n)
        g7s.append(out)
        ab = np.zeros((2, 6), real)
        fa(c["t1"], P(c["p1"]), P(c["pos1"]), P(c["R1"]), P(ab[0]))
        fa(c["t2"], P(c["p2"]), P(c["pos2"]), P(c["R2"]), P(ab[1]))
        aabbs.append(ab)
    return np.array(ns, np.int32), np.array(g7s), np.array(aabbs)


def contact_row_cases(real, n=48, seed=99):
    """Contact row-builder inputs: the two cases of the reference's tests/friction.cpp (mode Mu2|FDir1|Approx1,
    mu/mu2 = 0/1 and 1/0, bodies at (-1,0,0) and (1,0,0), normal (-1,0,0), fps 100, erp 0), then random ones."""
    MU2, FDIR1, APPROX1 = 0x001, 0x002, 0x7000
    cases = []
    base = dict(cpos=np.zeros(3, real), cnormal=np.array([-1, 0, 0], real), depth=0.0, fdir1=np.array([0, 1, 0], real),
                pos1=np.array([-1, 0, 0], real), pos2=np.array([1, 0, 0], real), fps=100.0, erp=0.0)
    cases.append(dict(base, mode=MU2 | FDIR1 | APPROX1, mu=0.0, mu2=1.0))
    cases.append(dict(base, mode=MU2 | FDIR1 | APPROX1, mu=1.0, mu2=0.0))
    r = np.random.RandomState(seed)
    modes = [0, APPROX1, MU2, MU2 | APPROX1, 0x010, 0x008 | 0x010, 0x100 | 0x200 | APPROX1, 0x020 | 0x040 | 0x080]
    for k in range(n):
        nrm = r.randn(3)
        nrm /= np.linalg.norm(nrm)
        mu = [0.0, 0.5, 1.0, np.inf][k % 4]
        cases.append(dict(cpos=np.asarray(0.3 * r.randn(3), real), cnormal=np.asarray(nrm, real), depth=float(abs(0.05 * r.randn())),
                          fdir1=np.array([0, 1, 0], real), pos1=np.asarray(r.randn(3), real), pos2=np.asarray(r.randn(3), real),
                          fps=float(50 + 100 * r.rand()), erp=float(r.rand()), mode=modes[k % len(modes)], mu=mu, mu2=float(r.rand())))
    return cases


def run_contact_rows(slib, prefix, cases):
    fn = getattr(slib.lib, prefix + "contact_rows")
    real = slib.real
    cr = C.c_float if real == np.float32 else C.c_double
    fn.argtypes = [C.c_int, cr, cr, C.c_void_p, C.c_void_p, cr, C.c_void_p, C.c_void_p, C.c_void_p, cr, cr, C.c_void_p, C.c_void_p]
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    ms, rows, fis = [], [], []
    for c in cases:
        out = np.zeros(48, real)
        fi = np.zeros(3, np.int32)
        m = fn(c["mode"], c["mu"], c["mu2"], P(c["cpos"]), P(c["cnormal"]), c["depth"], P(c["fdir1"]), P(c["pos1"]), P(c["pos2"]),
               c["fps"], c["erp"], P(out), P(fi))
        ms.append(m)
        rows.append(out.reshape(3, 16))
        fis.append(fi)
    return np.array(ms, np.int32), np.array(rows), np.array(fis)


# name -> (scene factory, step size, steps, checkpoint stride)
TRAJ_SCENES = {
    "stack": (lambda: scenes.box_stack(nworlds=2, nboxes=6), 0.02, 80, 8),
    "stack_plain": (lambda: scenes.box_stack(nworlds=2, nboxes=5, demo_world_options=False), 0.02, 60, 6),
    "pile": (lambda: scenes.pile(nbodies=27), 0.01, 90, 9),
    "pile_sap": (lambda: scenes.pile(nbodies=27, space_type=B.SPACE_SAP), 0.01, 60, 6),
    "chain": (lambda: scenes.chain(2, nlinks=6), 0.05, 80, 8),
    "free": (lambda: scenes.free_boxes(2, 9, grid=3), 0.01, 50, 5),
    "ragdoll": (lambda: scenes.ragdoll(2), 0.01, 160, 16),
}


def record_traj(batch, h, nsteps, every):
    rec = dict(steps=[], pos=[], quat=[], lvel=[], avel=[], seeds=[], stats=[], npairs=[], pairs=[], ncontacts=[], contacts=[],
               contact_g=[], nislands=[], islands=[])
    W = batch.W
    for s in range(1, nsteps + 1):
        batch.step(h)
        if s % every:
            continue
        st = batch.get_state()
        rec["steps"].append(s)
        for k in ("pos", "quat", "lvel", "avel"):
            rec[k].append(st[k])
        rec["seeds"].append(batch.get_seeds())
        rec["stats"].append(np.stack([batch.get_stats(w) for w in range(W)]))
        pr = [batch.get_pairs(w) for w in range(W)]
        ct = [batch.get_contacts(w) for w in range(W)]
        isl = [batch.get_islands(w) for w in range(W)]
        rec["npairs"].append([len(p) for p in pr])
        rec["pairs"].append(np.concatenate(pr) if sum(len(p) for p in pr) else np.zeros((0, 2), np.int32))
        rec["ncontacts"].append([len(c[1]) for c in ct])
        rec["contacts"].append(np.concatenate([c[0] for c in ct]))
        rec["contact_g"].append(np.concatenate([c[1] for c in ct]))
        rec["nislands"].append([i[0] for i in isl])
        rec["islands"].append(np.stack([i[1] for i in isl]))
    out = {}
    for k in ("steps", "pos", "quat", "lvel", "avel", "seeds", "stats", "npairs", "ncontacts", "nislands", "islands"):
        out[k] = np.array(rec[k])
    for k in ("pairs", "contacts", "contact_g"):
        out[k + "_cat"] = np.concatenate(rec[k]) if len(rec[k]) else np.zeros((0,))
        out[k + "_len"] = np.array([len(x) for x in rec[k]])
    return out


def compare_traj(batch, gold, h, exact=True, tol=None, resync=False, meas=None):
    """Replays the recorded scene on `batch` and compares every checkpoint. Returns list of mismatches.
    resync=True re-uploads the recorded reference state (and dRand seeds) after each checkpoint, so that every
    segment between checkpoints starts from the reference's state (teacher forcing) and last-ulp libm
    differences are not amplified over the whole run.  meas: optional dict that receives the largest contact / state deviation
    seen (tools/measure_tolerances.py: the stated tolerances are 4x these maxima)."""
    bad = []
    W = batch.W
    steps = list(gold["steps"])
    ofs = dict(pairs=0, contacts=0, contact_g=0)
    last = 0
    for ci, s in enumerate(steps):
        batch.step(h, int(s - last))
        last = int(s)
        st = batch.get_state()

        def seg(name):
            n = int(gold[name + "_len"][ci])
            a = gold[name + "_cat"][ofs[name]:ofs[name] + n]
            ofs[name] += n
            return a
        gp, gc, gg = seg("pairs"), seg("contacts"), seg("contact_g")
        pr = [batch.get_pairs(w) for w in range(W)]
        ct = [batch.get_contacts(w) for w in range(W)]
        mp = np.concatenate(pr) if sum(len(p) for p in pr) else np.zeros((0, 2), np.int32)
        if [len(p) for p in pr] != list(gold["npairs"][ci]) or not np.array_equal(mp, gp):
            bad.append("step %d: pair set differs" % s)
        mg = np.concatenate([c[1] for c in ct])
        if [len(c[1]) for c in ct] != list(gold["ncontacts"][ci]) or not np.array_equal(mg, gg):
            bad.append("step %d: contact counts / geoms differ" % s)
        else:
            mc = np.concatenate([c[0] for c in ct])
            if meas is not None and len(mc):
                meas["contact"] = max(meas.get("contact", 0.0), float(np.abs(mc.astype(np.float64) - gc).max()))
            if exact and not np.array_equal(mc, gc):
                bad.append("step %d: contact geometry differs by %.3g" % (s, np.abs(mc - gc).max()))
            elif not exact and len(mc) and np.abs(mc.astype(np.float64) - gc).max() > tol["contact"]:
                bad.append("step %d: contact geometry differs by %.3g" % (s, np.abs(mc.astype(np.float64) - gc).max()))
        isl = [batch.get_islands(w) for w in range(W)]
        if [i[0] for i in isl] != list(gold["nislands"][ci]) or not np.array_equal(np.stack([i[1] for i in isl]), gold["islands"][ci]):
            bad.append("step %d: island labels differ" % s)
        if not np.array_equal(np.stack([batch.get_stats(w) for w in range(W)]), gold["stats"][ci]):
            bad.append("step %d: iteration statistics differ" % s)
        if not np.array_equal(batch.get_seeds(), gold["seeds"][ci]):
            bad.append("step %d: dRand seeds differ" % s)
        for k in ("pos", "quat", "lvel", "avel"):
            if meas is not None:
                meas["state"] = max(meas.get("state", 0.0), float(np.abs(st[k].astype(np.float64) - gold[k][ci]).max()))
            if exact:
                if not np.array_equal(st[k], gold[k][ci]):
                    bad.append("step %d: %s differs by %.3g" % (s, k, np.abs(st[k] - gold[k][ci]).max()))
            else:
                d = np.abs(st[k].astype(np.float64) - gold[k][ci]).max()
                if not d <= tol["state"]:
                    bad.append("step %d: %s differs by %.3g > %.3g" % (s, k, d, tol["state"]))
        if bad:
            break
        if resync:
            batch.set_state(pos=gold["pos"][ci], quat=gold["quat"][ci], lvel=gold["lvel"][ci], avel=gold["avel"][ci])
            batch.set_seeds(gold["seeds"][ci])
    return bad
