#!/bin/bash
cd /root/repo
for v in "" ""; do
python - "$v" <<'PY'
import sys, os, ctypes as C
if sys.argv[1]: os.environ["ODEB_LIB_DIR"] = "/root/repo/ode_b200/variants/" + sys.argv[1]
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from parity_util import *
from ode_b200 import scenes
lib = gpu_lib("single"); L = lib.lib
L.odeb_timed_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_size_t, C.POINTER(C.c_double)]
sc = scenes.wall(500, 200); b = B.Batch(lib, sc); b.set_solver_mode(1)
b.step(0.05, 6)
ms = C.c_double(0); L.odeb_timed_steps(b.h, 0.05, 6, 0, C.byref(ms))
print("variant", sys.argv[1] or "shipped", "wall ms/step %.3f" % (ms.value / 6), b.get_totals(), flush=True)
PY
done
