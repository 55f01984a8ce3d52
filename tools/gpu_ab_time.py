"""A/B step timing of one scene: usage gpu_ab_time.py scene nworlds flush_mb [libdir]   (device time of 20 steps via odeb_timed_steps)"""
import sys, os, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
scene, nw, flush = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
if len(sys.argv) > 4: os.environ["ODEB_LIB_DIR"] = sys.argv[4]
from parity_util import *
from ode_b200 import scenes
sc, h, settle = {"ragdoll": (lambda: scenes.ragdoll(nw), 0.01, 60), "stack": (lambda: scenes.box_stack(nworlds=nw, demo_world_options=False), 0.02, 160),
                 "chain": (lambda: scenes.chain(nw), 0.05, 40)}[scene]
sc = sc()
lib = gpu_lib("single"); L = lib.lib
b = B.Batch(lib, sc)
b.step(h, settle)
L.odeb_timed_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_size_t, C.POINTER(C.c_double)]
for rep in range(3):
    ms = C.c_double(0)
    L.odeb_timed_steps(b.h, h, 20, flush << 20, C.byref(ms))
    print("%s %d worlds flush %d MB lib %s: %.4f ms/step" % (scene, nw, flush, os.environ.get("ODEB_LIB_DIR", "current"), ms.value / 20), flush=True)
