import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
os.environ["ODEB_SOLVER"] = sys.argv[1] if len(sys.argv) > 1 else "p4"
from parity_util import *
from ode_b200 import scenes
prec = sys.argv[2] if len(sys.argv) > 2 else "double"
which = sys.argv[3] if len(sys.argv) > 3 else "chain"
sc = scenes.chain(3) if which == "chain" else scenes.box_stack(nworlds=5, nboxes=8)
h = 0.05 if which == "chain" else 0.02
orc = B.Batch(orc_lib(prec), sc); gpu = B.Batch(gpu_lib(prec), sc)
for s in range(12):
    orc.step(h); gpu.step(h)
    bad = compare_step(orc, gpu, sc.nworlds)
    print(s, [gpu.get_islands(w)[0] for w in range(sc.nworlds)], gpu.get_totals(), [list(orc.get_stats(w)) for w in range(sc.nworlds)], bad[:3], flush=True)
    if bad: break
