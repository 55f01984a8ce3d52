"""Feasibility: K independent sub-batches (own streams/graphs) stepping concurrently vs one batch of the same total size."""
import sys, time, os, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from parity_util import *
from ode_b200 import scenes
K = int(sys.argv[1]); total = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
if len(sys.argv) > 3: os.environ["ODEB_SOLVER"] = sys.argv[3]
lib = gpu_lib("single"); L = lib.lib
L.odeb_step_async.argtypes = [C.c_void_p, C.c_double, C.c_int]; L.odeb_sync.argtypes = [C.c_void_p]
bs = [B.Batch(lib, scenes.box_stack(nworlds=total // K, demo_world_options=False, seed0=1000 + i * (total // K))) for i in range(K)]
for b in bs: b.step(0.02, 160)
for b in bs: b.step(0.02, 5)
def run(n):
    for b in bs: L.odeb_step_async(b.h, 0.02, n)
    for b in bs: L.odeb_sync(b.h)
run(5)
t = time.time(); run(50); dt = time.time() - t
print("K=%d total=%d solver=%s: ms/step %.3f body-steps/s %.3e" % (K, total, os.environ.get("ODEB_SOLVER", "auto"), dt / 50 * 1e3, total * 16 * 50 / dt), flush=True)
