"""SURVEY 8(d) protocol (ii): free-running divergence of the CUDA path from the oracle over N steps, for the scenes whose path
calls libm (CUDA libm vs glibc: <= 2 ulp per call, amplified by contact dynamics), beside the yardstick the reference has against
itself: the same oracle started from a state perturbed by one ulp in the height of one body.  Scenes without libm calls are bit-exact
(tests) and are listed with zeros as a check.  usage: divergence_report.py [nsteps]"""
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from parity_util import *
from ode_b200 import scenes

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
MARKS = [1, 10, 100, 300, 1000]


def qangle(qa, qb):
    """largest rotation angle between corresponding unit quaternions (q and -q are the same rotation)"""
    a, b = qa.astype(np.float64), qb.astype(np.float64)
    d = np.minimum(np.linalg.norm(a - b, axis=-1), np.linalg.norm(a + b, axis=-1))
    return float(np.max(4 * np.arcsin(np.clip(d / 2, 0, 1))))


def diff(sa, sb):
    return (float(np.abs(sa["pos"].astype(np.float64) - sb["pos"]).max()), qangle(sa["quat"], sb["quat"]),
            float(max(np.abs(sa["lvel"].astype(np.float64) - sb["lvel"]).max(), np.abs(sa["avel"].astype(np.float64) - sb["avel"]).max())))


def run(name, mk, h, prec):
    sc = mk()
    orc, gpu, pert = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc), B.Batch(orc_lib(prec), sc)
    st = orc.get_state()
    p = st["pos"].copy()
    p[0, 0, 2] = np.nextafter(p[0, 0, 2], np.inf)          # one ulp in the height of one body
    pert.set_state(pos=p)
    print("%s (%s, h=%g, %d worlds x %d bodies)" % (name, prec, h, sc.nworlds, sc.nbody))
    print("   step   | CUDA vs oracle: max|dpos|  max angle   max|dv|   | oracle vs 1-ulp-perturbed oracle: max|dpos|  max angle   max|dv|")
    for s in range(1, N + 1):
        orc.step(h); gpu.step(h); pert.step(h)
        if s in MARKS or s == N:
            a, g, q = orc.get_state(), gpu.get_state(), pert.get_state()
            d1, d2 = diff(a, g), diff(a, q)
            print("   %5d  |            %10.3e %10.3e %10.3e   |                          %10.3e %10.3e %10.3e" % ((s,) + d1 + d2), flush=True)
    gpu.close()


for prec in ("single", "double"):
    run("pile of 125 boxes+spheres (cullPoints atan2)", lambda: scenes.pile(nbodies=125), 0.01, prec)
    run("ragdoll (hinge / universal angles: atan2)", lambda: scenes.ragdoll(4), 0.01, prec)
    run("16-box stack (no libm on the path)", lambda: scenes.box_stack(nworlds=4, demo_world_options=False), 0.02, prec)
