"""Largest deviation of the CUDA path from the recorded reference trajectories / the oracle, per scene and precision (GPU box).
The tolerances stated in tests/test_gpu_parity.py are 4x these maxima; scenes that measure 0 are compared bit-exactly."""
import os, sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from parity_util import *
import golden_cases as G
from ode_b200 import scenes
GOLD = os.path.join(ROOT, "tests", "golden")
HUGE = dict(contact=1e30, state=1e30)
print("golden trajectories (segments between checkpoints re-synchronised to the recorded reference state):")
for name in sorted(G.TRAJ_SCENES):
    for prec in ("single", "double"):
        mk, h, nsteps, every = G.TRAJ_SCENES[name]
        gold = np.load(os.path.join(GOLD, "traj_%s_%s.npz" % (name, prec)))
        m = {}
        bad = G.compare_traj(B.Batch(gpu_lib(prec), mk()), gold, h, exact=False, tol=HUGE, resync=True, meas=m)
        m2 = {}
        bad2 = G.compare_traj(B.Batch(gpu_lib(prec), mk()), gold, h, exact=False, tol=HUGE, resync=False, meas=m2)
        print("  %-12s %-6s resync: contact %.3g state %.3g %s | free-running: contact %.3g state %.3g %s" % (
            name, prec, m.get("contact", 0), m.get("state", 0), "(integer mismatch: %s)" % bad[0] if bad else "", m2.get("contact", 0), m2.get("state", 0), "(integer mismatch: %s)" % bad2[0] if bad2 else ""), flush=True)
print("teacher-forced single steps vs the oracle:")
for name, mk, h, targets in (("pile125", lambda: scenes.pile(nbodies=125), 0.01, (0, 10, 100, 300)), ("ragdoll3", lambda: scenes.ragdoll(3), 0.01, (0, 10, 100, 300)),
                             ("pile1000", lambda: scenes.pile(nbodies=1000), 0.01, (0, 45))):
    for prec in ("single", "double"):
        sc = mk()
        a = B.Batch(orc_lib(prec), sc); b = B.Batch(gpu_lib(prec), sc)
        done = 0; ms = 0.0; mc = 0.0
        for target in targets:
            a.step(h, target - done); done = target
            st = a.get_state(); b.set_state(**st); b.set_seeds(a.get_seeds()); a.set_state(**st)
            a.step(h); b.step(h); done += 1
            sa, sb = a.get_state(), b.get_state()
            for k in ("pos", "quat", "lvel", "avel"):
                ms = max(ms, float(np.abs(sa[k].astype(np.float64) - sb[k]).max()))
            for w in range(sc.nworlds):
                (ga, ia), (gb, ib) = a.get_contacts(w), b.get_contacts(w)
                if ia.shape == ib.shape and len(ga): mc = max(mc, float(np.abs(ga.astype(np.float64) - gb).max()))
        print("  %-12s %-6s contact %.3g state %.3g" % (name, prec, mc, ms), flush=True)
