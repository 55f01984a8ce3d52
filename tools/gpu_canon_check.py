"""Canonical-mode (large-world path) GPU vs oracle, scene by scene, reporting every mismatch category (debug aid)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from parity_util import B, gpu_lib, orc_lib, compare_step
from ode_b200 import scenes
TOL = {"single": dict(contact=2e-5, state=2e-5), "double": dict(contact=1e-12, state=1e-12)}
cases = [("wall12x8_sap_mc8", lambda: scenes.wall(12, 8, max_contacts=8), 0.05, 40, True),
         ("wall9x5_hash", lambda: scenes.wall(9, 5, max_contacts=8, space_type=B.SPACE_HASH, ball=False), 0.05, 25, True),
         ("free100", lambda: scenes.free_boxes(1, 100, grid=10), 0.01, 40, True),
         ("stack16_adis", lambda: scenes.box_stack(nworlds=1, nboxes=16), 0.02, 120, True),
         ("chain", lambda: scenes.chain(1), 0.05, 60, True),
         ("pile216", lambda: scenes.pile(nbodies=216), 0.01, 60, False),
         ("ragdoll", lambda: scenes.ragdoll(1), 0.01, 60, False)]
only = sys.argv[1:] 
for prec in ("single", "double"):
    for name, mk, h, n, exact in cases:
        if only and name not in only:
            continue
        sc = mk()
        a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
        a.set_solver_mode(1); b.set_solver_mode(1)
        ok = True
        t0 = time.time()
        for s in range(n):
            a.step(h); b.step(h)
            bad = compare_step(a, b, 1, exact_float=exact, tol=TOL[prec])
            if bad:
                print(prec, name, "step", s, "MISMATCH", bad[:8], flush=True)
                ok = False
                break
        if ok:
            print(prec, name, "ok (%d steps, %.1fs)" % (n, time.time() - t0), "totals", b.get_totals(), flush=True)
