"""Large-world path on the GPU: timing of the wall scene (config 5) at a given size. usage: gpu_large.py W H steps [prec]"""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from parity_util import B, gpu_lib
from ode_b200 import scenes
W, H, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
prec = sys.argv[4] if len(sys.argv) > 4 else "single"
t0 = time.time(); sc = scenes.wall(W, H); t1 = time.time()
b = B.Batch(gpu_lib(prec), sc); b.set_solver_mode(1); t2 = time.time()
print("scene %.1fs create %.1fs bodies %d" % (t1 - t0, t2 - t1, sc.nbody), flush=True)
for k in range(steps):
    t = time.time(); b.step(0.05); dt = time.time() - t
    tot = b.get_totals()
    print("step %d: %.2f ms  pairs %d contacts %d rows %d islands %d sweeps %d rowsweeps %d" % (k, dt * 1e3, *tot), flush=True)
st = b.get_state()
print("finite", all(np.isfinite(v).all() for v in st.values()), "zmax", st["pos"][0, :, 2].max())
