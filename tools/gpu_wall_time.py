import sys, time, os, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from parity_util import *
from ode_b200 import scenes
sc = scenes.wall(500, 200)
lib = gpu_lib("single"); L = lib.lib
b = B.Batch(lib, sc); b.set_solver_mode(1)
b.step(0.05, 6)
L.odeb_timed_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_size_t, C.POINTER(C.c_double)]
ms = C.c_double(0); L.odeb_timed_steps(b.h, 0.05, 6, 0, C.byref(ms))
print(os.environ.get("ODEB_LIB_DIR", "default").split("/")[-1], "wall 100k: ms/step %.2f" % (ms.value / 6), b.get_totals(), flush=True)
