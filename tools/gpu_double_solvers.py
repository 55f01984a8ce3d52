"""The headline workload in double precision under each solver family. usage: gpu_double_solvers.py"""
import sys, os, ctypes as C, subprocess
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
if len(sys.argv) > 1:
    os.environ["ODEB_SOLVER"] = sys.argv[1]
    from parity_util import *
    from ode_b200 import scenes
    lib = gpu_lib("double"); L = lib.lib
    L.odeb_timed_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_size_t, C.POINTER(C.c_double)]
    L.odeb_solver_kernel.restype = C.c_char_p; L.odeb_solver_kernel.argtypes = [C.c_void_p]
    b = B.Batch(lib, scenes.box_stack(nworlds=4096, demo_world_options=False))
    b.step(0.02, 160)
    ms = C.c_double(0); L.odeb_timed_steps(b.h, 0.02, 20, 256 << 20, C.byref(ms))
    print("double, solver request %-4s -> %-34s %.4f ms/step" % (sys.argv[1], L.odeb_solver_kernel(b.h).decode(), ms.value / 20), flush=True)
else:
    for s in ("v4", "hy", "p4", "d4"):
        subprocess.run([sys.executable, __file__, s])
