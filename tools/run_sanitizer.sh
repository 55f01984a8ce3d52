#!/bin/bash
# compute-sanitizer memcheck over the large-world path (small scenes, both precisions) and the multi-island walker solver
cd /root/repo; mkdir -p gpurun_out
{
echo "== memcheck: canonical mode (coloured tile sweeps, persistent phases, feedback)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests -m gpu -x -q -k "canonical_mode_joint_feedback or canonical_mode_needs" 2>&1 | tail -6
echo "== memcheck: one canonical step sequence on wall(12,8) + pile(125)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 python - <<'PY' 2>&1 | tail -8
import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from parity_util import *
from ode_b200 import scenes
for prec in ("single", "double"):
    for mk, h, n in ((lambda: scenes.wall(12, 8, max_contacts=8), 0.05, 12), (lambda: scenes.pile(nbodies=125), 0.01, 30), (lambda: scenes.chain(1), 0.05, 10)):
        sc = mk(); b = B.Batch(gpu_lib(prec), sc); b.set_solver_mode(1)
        b.step(h, n)
        print(prec, sc.nbody, "bodies", n, "steps:", list(b.get_totals()), flush=True)
PY
echo "== racecheck: the same on wall(12,8), single"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 77 python - <<'PY' 2>&1 | tail -12
import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from parity_util import *
from ode_b200 import scenes
sc = scenes.wall(12, 8, max_contacts=8); b = B.Batch(gpu_lib("single"), sc); b.set_solver_mode(1)
b.step(0.05, 6)
print("racecheck run done", b.get_totals(), flush=True)
PY
} > gpurun_out/sanitizer_r2.log 2>&1
tail -40 gpurun_out/sanitizer_r2.log
