"""Break-down of the host-buffer (e2e) path: add_force / step / get_state per step, 4096 x 16-box stacks."""
import sys, time, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from parity_util import *
from ode_b200 import scenes
W = 4096
b = B.Batch(gpu_lib("single"), scenes.box_stack(nworlds=W, demo_world_options=False))
b.step(0.02, 160); b.step(0.02, 10)
force = np.zeros((W, 16, 3), np.float32); force[:, -1, 0] = 0.01
st = b.get_state()
ta = ts = tg = 0.0
n = 30
for i in range(n):
    t0 = time.time(); b.add_force(force=force); t1 = time.time(); b.step(0.02); t2 = time.time(); st = b.get_state(out=st); t3 = time.time()
    ta += t1 - t0; ts += t2 - t1; tg += t3 - t2
print("per step: add_force %.3f ms, step %.3f ms, get_state %.3f ms, total %.3f ms -> %.3e body-steps/s" % (ta / n * 1e3, ts / n * 1e3, tg / n * 1e3, (ta + ts + tg) / n * 1e3, W * 16 * n / (ta + ts + tg)))
