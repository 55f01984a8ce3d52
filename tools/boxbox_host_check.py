"""Box-box collider (ode_b200/csrc/odeb_boxbox.cuh, compiled for the host) against the compiled reference (oracle/_ref, ref_collide_pair)
on random, axis-aligned, parallel-edge, touching and deeply penetrating pairs, max-contact flags 1..8: counts and contacts bit-identical.
usage: boxbox_host_check.py [ncases]   (CPU only; builds /tmp/libbbh_{single,double}.so with nvcc)"""
import ctypes as C, os, subprocess, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import ref_lib, orc_lib, REAL
import golden_cases as G
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
bad_total = 0
for prec in ("single", "double"):
    so = "/tmp/libbbh_%s.so" % prec
    cmd = ["/usr/local/cuda/bin/nvcc", "-x", "cu", "-O2", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
           "-o", so, os.path.join(ROOT, "tools", "boxbox_host.cu")] + (["-DODEB_DOUBLE"] if prec == "double" else [])
    subprocess.check_call(cmd)
    L = C.CDLL(so)
    ref = ref_lib(prec) or orc_lib(prec)
    fn = getattr(ref.lib, ref.prefix + "collide_pair")
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.bbh_box_box.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_void_p]
    real = REAL[prec]
    r = np.random.RandomState(77)
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    bad = multi = hits = 0
    for k in range(N):
        s1 = np.asarray(0.2 + r.rand(3) * 1.3, real); s2 = np.asarray(0.2 + r.rand(3) * 1.3, real)
        kind = k % 8
        R1 = np.asarray(np.eye(3, 4) if kind in (1, 2) else G._rand_rot(r), real).reshape(12).copy()
        R2 = np.asarray(np.eye(3, 4) if kind in (1, 3) else G._rand_rot(r), real).reshape(12).copy()
        if kind == 4:      # one shared axis: parallel edges
            R2 = R1.copy().reshape(3, 4); a = r.rand() * 6.28; c, s = np.cos(a), np.sin(a)
            R2[:, :3] = R1.reshape(3, 4)[:, :3] @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]); R2 = np.asarray(R2, real).reshape(12).copy()
        pos1 = np.asarray(0.3 * r.randn(3), real)
        scale = [0.9, 0.45, 0.2, 1.4][k % 4]
        pos2 = np.asarray(pos1 + scale * r.randn(3) * np.array([1, 1, 0.6]), real)
        if kind == 5:      # exactly touching faces along z (zero depth)
            pos2 = np.asarray([pos1[0] + 0.1, pos1[1] - 0.05, pos1[2] + 0.5 * (s1[2] + s2[2])], real); R1 = np.asarray(np.eye(3, 4), real).reshape(12).copy(); R2 = R1.copy()
        flags = 1 + (k % 8)
        if k % 97 == 0: flags |= 0x80000000
        want = np.zeros((8, 7), real); got = np.zeros((8, 7), real); code = C.c_int(0)
        nw = fn(1, P(np.concatenate([s1, [0]]).astype(real)), P(pos1), P(R1), 1, P(np.concatenate([s2, [0]]).astype(real)), P(pos2), P(R2), C.c_int(flags & 0x7fffffff | (flags & 0x80000000)).value, P(want), 8)
        ng = L.bbh_box_box(P(s1), P(pos1), P(R1), P(s2), P(pos2), P(R2), C.c_int(flags).value, P(got), C.byref(code))
        hits += nw > 0; multi += nw > 1
        if nw != ng or not np.array_equal(want[:nw], got[:ng]):
            bad += 1
            if bad <= 5: print("  MISMATCH", prec, k, "kind", kind, "flags", hex(flags), "ref", nw, "ours", ng, "code", code.value, "\n", want[:nw], "\n", got[:ng])
    print("%s: %d pairs, %d colliding, %d with several contacts, %d mismatches" % (prec, N, hits, multi, bad))
    bad_total += bad
sys.exit(1 if bad_total else 0)
