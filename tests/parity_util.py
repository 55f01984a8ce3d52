"""Shared helpers for the parity tests: library discovery and observable-by-observable comparison."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from ode_b200 import _binding as B  # noqa: E402

REAL = {"single": np.float32, "double": np.float64}


def ref_lib(prec):
    p = os.path.join(ROOT, "oracle", "_ref", "libode_ref_%s.so" % prec)
    return B.SceneLib(p, "ref_", REAL[prec]) if os.path.exists(p) else None


def orc_lib(prec):
    p = os.path.join(ROOT, "oracle", "liborc_%s.so" % prec)
    return B.SceneLib(p, "orc_", REAL[prec]) if os.path.exists(p) else None


def gpu_lib(prec):
    p = os.path.join(os.environ.get("ODEB_LIB_DIR", os.path.join(ROOT, "ode_b200")), "libode_b200_%s.so" % prec)   # ODEB_LIB_DIR: kernel-variant experiments (tools/)
    if not os.path.exists(p):
        raise RuntimeError("CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % p)
    return B.SceneLib(p, "odeb_", REAL[prec])


def ulp_diff(a, b):
    """max |a-b| measured in units of the last place of max(|a|,|b|,tiny)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.size == 0:
        return 0.0
    eps = np.finfo(a.dtype).eps
    scale = np.maximum(np.maximum(np.abs(a), np.abs(b)), np.finfo(a.dtype).tiny).astype(np.float64)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)) / (scale * eps)))


def compare_step(a, b, nworlds, exact_float=True, tol=None, what=("pairs", "contacts", "islands", "stats", "seeds", "state")):
    """Compare the observables of the most recent step of two batches.

    Integer observables (pair set, per-pair contact counts / contact->geom map, island labels, the four
    dynamic-iteration counters, dRand seed) must be identical. Floating observables are compared
    bit-exactly (exact_float) or to `tol` = dict(contact=abs, state=abs).
    Returns a list of human-readable mismatches (empty = parity).
    """
    bad = []
    for w in range(nworlds):
        if "pairs" in what:
            pa, pb = a.get_pairs(w), b.get_pairs(w)
            if pa.shape != pb.shape or not np.array_equal(pa, pb):
                bad.append("world %d: pair set differs (%d vs %d)" % (w, len(pa), len(pb)))
        if "contacts" in what:
            (ga, ia), (gb, ib) = a.get_contacts(w), b.get_contacts(w)
            if ia.shape != ib.shape or not np.array_equal(ia, ib):
                bad.append("world %d: contact count / geoms differ (%d vs %d)" % (w, len(ia), len(ib)))
            elif exact_float:
                if not np.array_equal(ga, gb):
                    bad.append("world %d: contact geometry differs, max %.3g (%.1f ulp)" % (w, np.abs(ga - gb).max(), ulp_diff(ga, gb)))
            elif len(ga) and np.abs(ga.astype(np.float64) - gb).max() > tol["contact"]:
                bad.append("world %d: contact geometry differs by %.3g" % (w, np.abs(ga.astype(np.float64) - gb).max()))
        if "islands" in what:
            (na, la), (nb_, lb) = a.get_islands(w), b.get_islands(w)
            if na != nb_ or not np.array_equal(la, lb):
                bad.append("world %d: island labels differ (%d vs %d islands)" % (w, na, nb_))
        if "stats" in what:
            sa, sb = a.get_stats(w), b.get_stats(w)
            if not np.array_equal(sa, sb):
                bad.append("world %d: iteration statistics differ %s vs %s" % (w, sa, sb))
    if "seeds" in what and not np.array_equal(a.get_seeds(), b.get_seeds()):
        bad.append("dRand seeds differ")
    if "state" in what:
        sa, sb = a.get_state(), b.get_state()
        for k in ("pos", "quat", "lvel", "avel"):
            if exact_float:
                if not np.array_equal(sa[k], sb[k]):
                    bad.append("%s differs, max %.3g (%.1f ulp)" % (k, np.abs(sa[k] - sb[k]).max(), ulp_diff(sa[k], sb[k])))
            else:
                d = np.abs(sa[k].astype(np.float64) - sb[k]).max()
                if not d <= tol["state"]:
                    bad.append("%s differs by %.3g > %.3g" % (k, d, tol["state"]))
    return bad
