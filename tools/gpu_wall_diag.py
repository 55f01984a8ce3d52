"""Diagnostic: which contacts of step 1 of the full 100k-box wall differ between the oracle and the CUDA path (canonical mode)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from parity_util import *
from ode_b200 import scenes
nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (500, 200)
sc = scenes.wall(nx, ny)
a, b = B.Batch(orc_lib("single"), sc), B.Batch(gpu_lib("single"), sc)
a.set_solver_mode(1); b.set_solver_mode(1)
a.step(0.05); b.step(0.05)
print("step 0:", compare_step(a, b, 1) or "identical", flush=True)
a.step(0.05); b.step(0.05)
(ga, ia), (gb, ib) = a.get_contacts(0), b.get_contacts(0)
print("contacts", ga.shape, ia.shape, "index arrays equal:", np.array_equal(ia, ib))
d = np.abs(ga.astype(np.float64) - gb)
rows = np.unique(np.nonzero(d.reshape(len(ga), -1) > 0)[0])
print("contacts with different geometry:", len(rows), "of", len(ga))
for r in rows[:12]:
    print(" contact", r, "index entry", ia[r], "\n   oracle", ga[r], "\n   cuda  ", gb[r])
sa, sb = a.get_state(), b.get_state()
for k in ("pos", "quat", "lvel", "avel"):
    dd = np.abs(sa[k].astype(np.float64) - sb[k]).reshape(sc.nbody, -1).max(axis=1)
    nz = np.nonzero(dd > 0)[0]
    print(k, "bodies that differ:", len(nz), "max", dd.max(), "worst body", int(dd.argmax()))
