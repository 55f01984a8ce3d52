import sys, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from parity_util import *
from ode_b200 import scenes
for prec in ("double",):
    lib = gpu_lib(prec); L = lib.lib
    L.odeb_timed_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_size_t, C.POINTER(C.c_double)]
    sc = scenes.wall(500, 200); b = B.Batch(lib, sc); b.set_solver_mode(1)
    b.step(0.05, 6)
    ms = C.c_double(0); L.odeb_timed_steps(b.h, 0.05, 6, 0, C.byref(ms))
    st = b.get_state()
    import numpy as np
    print(prec, "wall ms/step %.3f" % (ms.value / 6), b.get_totals(), "finite", bool(np.isfinite(st["pos"]).all()), flush=True)
