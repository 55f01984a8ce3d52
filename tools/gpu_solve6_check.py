"""k_solve6 (independent world walkers): parity vs the oracle and timings on the GPU box.
usage: gpu_solve6_check.py parity <solver> <prec> | time <scene> <nworlds> <solver> [nsteps]"""
import sys, time, os, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
mode = sys.argv[1]
if mode == "parity":
    solver, prec = sys.argv[2], sys.argv[3]
    if solver != "auto": os.environ["ODEB_SOLVER"] = solver
    from parity_util import *
    from ode_b200 import scenes
    ok = True
    for name, mk, h, n in (("stack8", lambda: scenes.box_stack(nworlds=5, nboxes=8), 0.02, 40),
                           ("stack16", lambda: scenes.box_stack(nworlds=7), 0.02, 60),
                           ("chain", lambda: scenes.chain(3), 0.05, 60),
                           ("free16", lambda: scenes.free_boxes(2, 16, grid=4), 0.01, 40),
                           ("free100", lambda: scenes.free_boxes(3, 100, grid=10), 0.01, 20),
                           ("pile64", lambda: scenes.pile(nworlds=9, nbodies=64, vary=0.05), 0.01, 200),
                           ("pile27", lambda: scenes.pile(nworlds=21, nbodies=27, vary=0.1), 0.01, 120),
                           ("stack70", lambda: scenes.box_stack(nworlds=3, nboxes=70, demo_world_options=False), 0.02, 25)):
        sc = mk()
        a, b = B.Batch(orc_lib(prec), sc), B.Batch(gpu_lib(prec), sc)
        bad = None
        for s in range(n):
            a.step(h); b.step(h)
            bad = compare_step(a, b, sc.nworlds)
            if bad: break
        L = gpu_lib(prec).lib; L.odeb_solver_kernel.restype = C.c_char_p; L.odeb_solver_kernel.argtypes = [C.c_void_p]
        print("  %-8s %-6s %-7s kernel %-12s %s" % (name, solver, prec, L.odeb_solver_kernel(b.h).decode(), ("MISMATCH step %d: %s" % (s, bad[:3])) if bad else "bit-exact %d steps" % n), flush=True)
        ok = ok and not bad
        b.close()
    sys.exit(0 if ok else 1)
else:
    scene, nw, solver = sys.argv[2], int(sys.argv[3]), sys.argv[4]
    nsteps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
    if solver != "auto": os.environ["ODEB_SOLVER"] = solver
    from parity_util import *
    from ode_b200 import scenes
    if scene == "pile64": sc, h, settle, nb = scenes.pile(nworlds=nw, nbodies=64), 0.01, 150, 64
    elif scene == "pile64s": sc, h, settle, nb = scenes.pile(nworlds=nw, nbodies=64), 0.01, 400, 64
    elif scene == "pile64v": sc, h, settle, nb = scenes.pile(nworlds=nw, nbodies=64, vary=0.05), 0.01, 150, 64
    elif scene == "stack16": sc, h, settle, nb = scenes.box_stack(nworlds=nw, demo_world_options=False), 0.02, 160, 16
    elif scene == "ragdoll": sc, h, settle, nb = scenes.ragdoll(nworlds=nw), 0.01, 100, None
    elif scene == "chain": sc, h, settle, nb = scenes.chain(nw), 0.05, 100, 10
    nb = nb or sc.nbody
    lib = gpu_lib("single"); L = lib.lib
    b = B.Batch(lib, sc)
    b.step(h, settle); b.step(h, 5)
    L.odeb_solver_kernel.restype = C.c_char_p; L.odeb_solver_kernel.argtypes = [C.c_void_p]
    L.odeb_timed_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_size_t, C.POINTER(C.c_double)]
    ms = C.c_double(0)
    L.odeb_timed_steps(b.h, h, nsteps, 0, C.byref(ms))
    tot = b.get_totals()
    L.odeb_enable_timing.argtypes = [C.c_void_p, C.c_int]; L.odeb_solver_ms.restype = C.c_double; L.odeb_solver_ms.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    nl = C.c_int(0); L.odeb_solver_ms(b.h, C.byref(nl))
    L.odeb_enable_timing(b.h, 1); b.step(h, 5); L.odeb_enable_timing(b.h, 0)
    ms1 = L.odeb_solver_ms(b.h, C.byref(nl))
    print("   solver kernels: %.3f ms per step" % (ms1 / max(1, nl.value)), flush=True)
    print("%-8s %6d worlds solver %-6s (%s): ms/step %.3f  body-steps/s %.3e  rows/world %.1f islands/world %.2f sweeps/island %.1f row-sweeps %.3e" % (
        scene, nw, solver, L.odeb_solver_kernel(b.h).decode(), ms.value / nsteps, nw * nb * nsteps / (ms.value * 1e-3), tot[2] / nw, tot[3] / nw, tot[4] / max(1, tot[3]), tot[5]), flush=True)
