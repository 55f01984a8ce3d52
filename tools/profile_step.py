"""Runs a few steps of the bench workload for ncu (launch list / full capture)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from parity_util import B, gpu_lib
from ode_b200 import scenes
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 30
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
prec = sys.argv[4] if len(sys.argv) > 4 else "single"
sc = scenes.box_stack(nworlds=nw, demo_world_options=False)
gpu = B.Batch(gpu_lib(prec), sc)
gpu.step(0.02, warm)
gpu.step(0.02, steps)
print("done")
