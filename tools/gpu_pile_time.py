"""64-body worlds (north-star shape): timing of a batch of piles. usage: gpu_pile_time.py nworlds nbodies [solver]"""
import sys, time, os, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
nw, nb = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 3 and sys.argv[3] != "auto": os.environ["ODEB_SOLVER"] = sys.argv[3]
from parity_util import *
from ode_b200 import scenes
sc = scenes.pile(nworlds=nw, nbodies=nb)
lib = gpu_lib("single"); L = lib.lib
b = B.Batch(lib, sc)
t = time.time(); b.step(0.01, 150); print("settle 150 steps: %.2f s" % (time.time() - t), flush=True)
b.step(0.01, 5)
L.odeb_solver_kernel.restype = C.c_char_p; L.odeb_solver_kernel.argtypes = [C.c_void_p]
t = time.time(); b.step(0.01, 20); dt = time.time() - t
tot = b.get_totals()
print("%d worlds x %d-body pile, solver %s: ms/step %.3f  body-steps/s %.3e  rows/world %.1f islands/world %.2f sweeps/island %.1f" % (
    nw, nb, L.odeb_solver_kernel(b.h).decode(), dt / 20 * 1e3, nw * nb * 20 / dt, tot[2] / nw, tot[3] / nw, tot[4] / max(1, tot[3])), flush=True)
