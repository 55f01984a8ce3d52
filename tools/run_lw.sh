#!/bin/bash
# large-world path: canonical-mode parity tests (the full-size wall excluded: it has its own slot in the suite), timing of the 100k-box wall and the
# 1000-body pile, launch list of a wall step
cd /root/repo; mkdir -p gpurun_out
{
for v in 1; do timeout 600 python -m pytest tests -m gpu -x -q -k "canonical" --deselect tests/test_gpu_parity.py::test_canonical_mode_full_size_wall --deselect "tests/test_gpu_parity.py::test_canonical_mode_vs_compiled_reference[single]" 2>&1 | tail -3
python - <<'PY'
import sys, time, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from parity_util import *
from ode_b200 import scenes
lib = gpu_lib("single"); L = lib.lib
L.odeb_timed_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_size_t, C.POINTER(C.c_double)]
for name, mk, h in (("wall 500x200", lambda: scenes.wall(500, 200), 0.05), ("pile 1000", lambda: scenes.pile(nbodies=1000), 0.01)):
    sc = mk(); b = B.Batch(lib, sc); b.set_solver_mode(1)
    b.step(h, 6 if "wall" in name else 60)
    ms = C.c_double(0); L.odeb_timed_steps(b.h, h, 6, 0, C.byref(ms))
    print(name, "ms/step %.3f" % (ms.value / 6), b.get_totals(), flush=True)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_wall.csv python tools/profile_scene.py wall 1 8 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_wall.csv 24
} > gpurun_out/lw_b.log 2>&1
tail -60 gpurun_out/lw_b.log
