"""Island structure of a scene on the oracle (CPU): per island bodies / rows / sweeps. usage: island_stats.py scene nworlds nsettle nsample"""
import sys, ctypes as C, numpy as np, collections
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from parity_util import *
from ode_b200 import scenes
name, nw, nsettle, nsample = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
sc = {"pile64": lambda: scenes.pile(nworlds=nw, nbodies=64), "stack16": lambda: scenes.box_stack(nworlds=nw, demo_world_options=False),
      "ragdoll": lambda: scenes.ragdoll(nworlds=nw), "chain": lambda: scenes.chain(nw)}[name]()
lib = orc_lib("single")
b = B.Batch(lib, sc)
h = {"pile64": 0.01, "stack16": 0.02, "ragdoll": 0.01, "chain": 0.05}[name]
b.step(h, nsettle)
f = lib.lib.orc_get_island_log; f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
prev = {}
hit = tot = 0
for s in range(nsample):
    b.step(h)
    for w in range(nw):
        buf = np.zeros((256, 3), np.int32)
        n = f(b.h, w, buf.ctypes.data, 256)
        L = buf[:n]
        if s == nsample - 1 and w < 3:
            print("world", w, "islands (nb, m, sweeps):", [tuple(x) for x in L if x[1] > 0], "+ %d empty" % int((L[:, 1] == 0).sum()))
            ser = sum(int(x[1]) * int(x[2]) for x in L)
            mx = max(int(x[1]) * int(x[2]) for x in L)
            print("   row-sweeps total %d, largest island %d" % (ser, mx))
        # prediction of the number of shuffles from the previous step, keyed by island index
        k = [(max(int(x[2]) - 1, 0)) // 8 for x in L]
        p = prev.get(w)
        if p is not None:
            for i, (x, kk) in enumerate(zip(L, k)):
                if x[1] > 1:
                    tot += 1
                    if i < len(p[0]) and p[0][i] == kk and p[1][i] == x[1]: hit += 1
        prev[w] = (k, [int(x[1]) for x in L])
print("shuffle-count prediction from the previous step (same island index, same m): %d / %d" % (hit, tot))
