"""Runs a few steps of one scene family for ncu. usage: profile_scene.py stack|chain|ragdoll|wall|pile64 nworlds warm steps [prec]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from parity_util import B, gpu_lib
from ode_b200 import scenes
name = sys.argv[1]; nw = int(sys.argv[2]); warm = int(sys.argv[3]); steps = int(sys.argv[4])
prec = sys.argv[5] if len(sys.argv) > 5 else "single"
mk = {"stack": lambda: (scenes.box_stack(nworlds=nw, demo_world_options=False), 0.02),
      "chain": lambda: (scenes.chain(nw), 0.05), "ragdoll": lambda: (scenes.ragdoll(nw), 0.01),
      "wall": lambda: (scenes.wall(500, 200), 0.05), "pile64": lambda: (scenes.pile(nworlds=nw, nbodies=64), 0.01)}[name]
sc, h = mk()
gpu = B.Batch(gpu_lib(prec), sc)
if name == "wall":
    gpu.set_solver_mode(1)
gpu.step(h, warm)
# ncu --profile-from-start off: only the last `steps` steps are captured
import ctypes
rt = None
for name_ in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
    try:
        rt = ctypes.CDLL(name_); break
    except OSError:
        pass
if rt: rt.cudaProfilerStart()
gpu.step(h, steps)
if rt: rt.cudaProfilerStop()
print("done", gpu.get_totals())
