"""Cycle accounting of k_solve6 (variant build with -DODEB6_PROF). usage: gpu_prof6.py scene nworlds solver"""
import sys, os, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
scene, nw, solver = sys.argv[1], int(sys.argv[2]), sys.argv[3]
os.environ["ODEB_SOLVER"] = solver
os.environ["ODEB_LIB_DIR"] = "/root/repo/ode_b200/variants/prof6"
from parity_util import *
from ode_b200 import scenes
if scene == "pile64": sc, h, settle = scenes.pile(nworlds=nw, nbodies=64), 0.01, 150
elif scene == "pile64s": sc, h, settle = scenes.pile(nworlds=nw, nbodies=64), 0.01, 400
else: sc, h, settle = scenes.box_stack(nworlds=nw, demo_world_options=False), 0.02, 160
lib = gpu_lib("single"); L = lib.lib
b = B.Batch(lib, sc)
b.step(h, settle); b.step(h, 5)
out = (C.c_uint64 * 16)()
L.odeb_debug_prof(out, 1)
n = 10
b.step(h, n)
L.odeb_debug_prof(out, 0)
v = [int(x) for x in out]
print("%s %d worlds %s: slowest warp %.0f kcyc; per warp-step: total %.0f kcyc, transition wall %.0f kcyc (%.0f %%), trips %.0f, with transition %.0f" % (scene, nw, solver, v[14] / 1e3, v[0] / v[11] / 1e3, v[1] / v[11] / 1e3, 100.0 * v[1] / max(1, v[0]), v[2] / v[11], v[3] / v[11]))
print("   per world-step (leader): transitions %.0f; kcyc control+commit %.1f, fetch %.1f, schedule %.1f, prime %.1f; FY catch-up steps %.0f; serial fallbacks %.2f; slots built %.0f for %.0f rows (%.2f rows/slot)" % (
    v[9] / nw / n, v[4] / nw / n / 1e3, v[6] / nw / n / 1e3, v[7] / nw / n / 1e3, v[8] / nw / n / 1e3, v[5] / nw / n, v[10] / nw / n, v[12] / nw / n, v[13] / nw / n, v[13] / max(1, v[12])))
