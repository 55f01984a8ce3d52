"""Timing only: 4096 x 16-box stacks, given solver (argv[1]) and precision (argv[2]); prints ms/step and solver ms."""
import sys, time, os, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
if len(sys.argv) > 1 and sys.argv[1] != "auto": os.environ["ODEB_SOLVER"] = sys.argv[1]
prec = sys.argv[2] if len(sys.argv) > 2 else "single"
nw = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
from parity_util import *
from ode_b200 import scenes
sc = scenes.box_stack(nworlds=nw, demo_world_options=False)
lib = gpu_lib(prec); L = lib.lib
gpu = B.Batch(lib, sc)
gpu.step(0.02, 160); gpu.step(0.02, 5)
t = time.time(); gpu.step(0.02, 50); dt = time.time() - t
L.odeb_enable_timing.argtypes = [C.c_void_p, C.c_int]
L.odeb_solver_ms.restype = C.c_double; L.odeb_solver_ms.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
L.odeb_enable_timing(gpu.h, 1); gpu.step(0.02, 10); n = C.c_int(0); ms = L.odeb_solver_ms(gpu.h, C.byref(n)); L.odeb_enable_timing(gpu.h, 0)
print(os.environ.get("ODEB_LIB_DIR", "default").split("/")[-1], sys.argv[1:], "ms/step %.3f  solver ms %.3f" % (dt / 50 * 1e3, ms / max(1, n.value)), flush=True)
