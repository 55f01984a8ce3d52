"""Parity (bit-exact vs oracle) + timing of the body-lane solver on the GPU box."""
import sys, time, os, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from parity_util import *
from ode_b200 import scenes
def run(name, sc, h, nsteps, prec):
    orc = B.Batch(orc_lib(prec), sc); gpu = B.Batch(gpu_lib(prec), sc)
    for s in range(nsteps):
        orc.step(h); gpu.step(h)
        bad = compare_step(orc, gpu, sc.nworlds)
        if bad:
            print(name, prec, "step", s, "MISMATCH:", bad[:6]); return False
    print(name, prec, "bit-exact for", nsteps, "steps")
    return True
if "--notest" not in sys.argv:
    for prec in ("single", "double"):
        run("stack8", scenes.box_stack(nworlds=5, nboxes=8), 0.02, 100, prec)
        run("stack16", scenes.box_stack(nworlds=3), 0.02, 120, prec)
        run("stack24", scenes.box_stack(nworlds=3, nboxes=24, demo_world_options=False), 0.02, 100, prec)
        run("chain", scenes.chain(3), 0.05, 120, prec)
        run("free", scenes.free_boxes(2, 16, grid=4), 0.01, 80, prec)
        run("pile32", scenes.pile(nbodies=32), 0.01, 120, prec)
for prec in ("single", "double"):
    sc = scenes.box_stack(nworlds=4096, demo_world_options=False)
    gpu = B.Batch(gpu_lib(prec), sc)
    gpu.step(0.02, 160)
    t = time.time(); gpu.step(0.02, 50); dt = time.time() - t
    print(prec, "4096 stacks: ms/step %.3f" % (dt / 50 * 1e3), "body-steps/s %.3e" % (4096 * 16 * 50 / dt), flush=True)
    gpu.close()
