"""Two CPU-only yardsticks for the order conventions of this library:
 (a) callback order: the UNMODIFIED reference stepped with its natural callback order (contacts created inside the near callback as
     the space reports the pairs: ref_step_plain) against the same reference with the pairs sorted by geom index before the contacts
     are created (ref_step: what tests/classic_app.py and this library's dSpaceCollide do) -- reference vs reference;
 (b) canonical row order (ODEB_MODE_CANONICAL, the large-world path): the oracle in canonical mode (bit-identical to the CUDA path,
     tests/test_gpu_parity.py::test_canonical_mode_bit_exact) against the unmodified reference in its own order, free-running.
Both are order effects of a Gauss-Seidel sweep, not errors; the columns say how fast trajectories that differ only in row order
drift apart, next to what a 1-ulp perturbation of one body's height does to the reference against itself.
usage: order_yardsticks.py [nsteps]"""
import sys, os, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from parity_util import *
from ode_b200 import scenes
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
MARKS = [1, 10, 30, 100, 300, 1000]


def dmax(a, b):
    return (float(np.abs(a["pos"].astype(np.float64) - b["pos"]).max()), float(max(np.abs(a["lvel"].astype(np.float64) - b["lvel"]).max(), np.abs(a["avel"].astype(np.float64) - b["avel"]).max())))


def run(name, mk, h, prec):
    ref = ref_lib(prec)
    fn = ref.lib.ref_step_plain
    fn.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
    sc = mk()
    srt, nat, pert, can = B.Batch(ref, sc), B.Batch(ref, sc), B.Batch(ref, sc), B.Batch(orc_lib(prec), sc)
    can.set_solver_mode(1)
    st = srt.get_state()
    p = st["pos"].copy()
    p[0, 0, 2] = np.nextafter(p[0, 0, 2], np.inf)
    pert.set_state(pos=p)
    print("%s (%s, h=%g, %d bodies)" % (name, prec, h, sc.nbody))
    print("   step | natural vs sorted callback order | canonical row order vs reference | reference vs 1-ulp-perturbed reference   (max|dpos|, max|dv|)")
    for s in range(1, N + 1):
        srt.step(h); pert.step(h); can.step(h)
        fn(nat.h, h, 1, 0, sc.nworlds)
        if s in MARKS or s == N:
            a = srt.get_state()
            d1, d2, d3 = dmax(a, nat.get_state()), dmax(a, can.get_state()), dmax(a, pert.get_state())
            same = all(np.array_equal(srt.get_pairs(0), x.get_pairs(0)) for x in (nat, can))
            print("  %5d | %10.3e %10.3e            | %10.3e %10.3e            | %10.3e %10.3e      pair sets %s" % ((s,) + d1 + d2 + d3 + ("equal" if same else "differ (trajectories have separated)",)), flush=True)


for prec in ("single", "double"):
    run("wall 30 x 20 + cannon ball, sweep-and-prune space (configs[4] shape)", lambda: scenes.wall(30, 20), 0.05, prec)
    run("pile of 343 boxes + spheres, hash space (configs[0] shape)", lambda: scenes.pile(nbodies=343), 0.01, prec)
    run("16-box stack, hash space (configs[1] shape)", lambda: scenes.box_stack(nworlds=1, demo_world_options=False), 0.02, prec)
