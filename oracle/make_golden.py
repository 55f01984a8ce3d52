"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile from
/root/reference). TEST INFRASTRUCTURE: run here (the reference cannot travel), commit the small fixtures.

    python oracle/make_golden.py

Fixtures:
  rand.npz                 dRand / dRandInt streams (ode/src/misc.cpp:35-139) incl. the dTestRand known answers
  collide_<prec>.npz       dCollide on random poses for every supported geom pair (+ AABBs)
  contact_rows_<prec>.npz  dxJointContact::getInfo1/getInfo2 rows, incl. the two cases of tests/friction.cpp:69-173
  traj_<scene>_<prec>.npz  per-step observables of small scenes: state, pair set, contacts, islands, stats, seeds
"""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import B, ref_lib, REAL  # noqa: E402
from golden_cases import collide_cases, contact_row_cases, TRAJ_SCENES, run_collide, run_contact_rows, record_traj  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    lib = ref_lib("single")
    # --- RNG
    fn = lib.lib.ref_rand_next
    fn.restype = C.c_ulong
    seed = C.c_uint32(0)
    stream = np.array([fn(C.byref(seed)) for _ in range(64)], dtype=np.uint64)
    assert list(stream[:5]) == [0x3c6ef35f, 0x47502932, 0xd1ccf6e9, 0xaaf95334, 0x6252e503]
    assert lib.lib.ref_test_rand() == 1
    ns = np.array([1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 100, 255, 256, 257, 1000, 65535, 65536, 65537, 1000000, 2 ** 31 - 1], dtype=np.int64)
    seed = C.c_uint32(12345)
    ri = []
    for rep in range(8):
        for n in ns:
            ri.append(lib.lib.ref_rand_int(C.byref(seed), int(n)))
    np.savez(os.path.join(OUT, "rand.npz"), stream=stream, ns=ns, randint=np.array(ri, dtype=np.int64), final_seed=np.uint32(seed.value))
    for prec in ("single", "double"):
        lib = ref_lib(prec)
        cases = collide_cases(REAL[prec])
        n, g7, aabb = run_collide(lib, "ref_", cases)
        np.savez_compressed(os.path.join(OUT, "collide_%s.npz" % prec), n=n, geom7=g7, aabb=aabb)
        rc = contact_row_cases(REAL[prec])
        m, rows, fi = run_contact_rows(lib, "ref_", rc)
        np.savez_compressed(os.path.join(OUT, "contact_rows_%s.npz" % prec), m=m, rows=rows, findex=fi)
        for name, (mk, h, nsteps, every) in TRAJ_SCENES.items():
            rec = record_traj(B.Batch(lib, mk()), h, nsteps, every)
            np.savez_compressed(os.path.join(OUT, "traj_%s_%s.npz" % (name, prec)), **rec)
            print(prec, name, "recorded", len(rec["steps"]), "checkpoints")


if __name__ == "__main__":
    main()
