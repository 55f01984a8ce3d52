#!/usr/bin/env python
"""bench.py -- body-steps/s of the per-step world update (collide -> QuickStep) on batched worlds.

Workload (BASELINE.json configs[1]): 4096 independent worlds per GPU of a 16-box stack on a plane
(demo_boxstack-style surface, quickstep 20 iterations + the reference's default dynamic iteration
adjustment, dt = 0.02, auto-disable OFF so the work per step does not vanish when the stack settles).
A "step" = one pass of the whole hot path over all worlds of the rank.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched through torch.distributed.run)
    python bench.py --impl reference ...                   (the reference's CPU path on the host cores)

Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from ode_b200 import _binding as B  # noqa: E402
from ode_b200 import scenes  # noqa: E402

WORLDS_PER_GPU = 4096
NBOX = 16
H = 0.02
PREC = "single"
METRIC = "body_steps_per_sec"
FLUSH_BYTES = 256 << 20
WORKLOAD = "%d worlds/GPU x %d-box stack on a plane (BASELINE configs[1]), dt=%g, 20 iters + default dynamic adjustment, auto-disable off" % (WORLDS_PER_GPU, NBOX, H)
T_START = time.time()


def make_scene(nworlds, seed0=1000):
    return scenes.box_stack(nworlds=nworlds, nboxes=NBOX, seed0=seed0, demo_world_options=False)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            v = j.get("hbm_gbs") or j.get("hbm_gb_s")
            if v:
                return float(v), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons}


# ---------------------------------------------------------------------------------------------- CPU arm

def _cpu_worker(args):
    path, prefix, nworlds, seed0, steps, warm = args
    real = np.float32 if PREC == "single" else np.float64
    lib = B.SceneLib(path, prefix, real)
    b = B.Batch(lib, make_scene(nworlds, seed0))
    if prefix == "ref_":
        fn = lib.lib.ref_step_plain
        fn.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
        fn(b.h, H, warm, 0, nworlds)     # warm = settle + warm-up steps: same phase of the scene as the GPU arm
        t = time.time()
        fn(b.h, H, steps, 0, nworlds)
        return time.time() - t
    b.step(H, warm)
    t = time.time()
    b.step(H, steps)
    return time.time() - t


def cpu_reference(steps, warm, worlds_per_proc=16, procs=None):
    """The reference's own CPU implementation of the path (oracle/_ref: unmodified ODE built from its sources,
    plain dSpaceCollide / dWorldQuickStep / dJointGroupEmpty loop, self-threaded stepper), one independent
    process per host core each stepping its own worlds (dRand's seed is process-global)."""
    import multiprocessing as mp
    ref = os.path.join(ROOT, "oracle", "_ref", "libode_ref_%s.so" % PREC)
    if os.path.exists(ref):
        path, prefix, kind = ref, "ref_", "reference"
    else:
        path, prefix, kind = os.path.join(ROOT, "oracle", "liborc_%s.so" % PREC), "orc_", "port"
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        t0 = time.time()
        times = pool.map(_cpu_worker, [(path, prefix, worlds_per_proc, 1000 + 97 * i, steps, warm) for i in range(procs)])
        wall = time.time() - t0
    slowest = max(times)
    value = procs * worlds_per_proc * NBOX * steps / slowest
    sample = "%d processes x %d worlds x %d steps of the %d-box stack after %d settle steps (max process time %.2f s, pool wall %.2f s)" % (
        procs, worlds_per_proc, steps, NBOX, warm, slowest, wall)
    return {"value": value, "unit": "body-steps/s", "cores": procs, "kind": kind, "sample": sample}


# ---------------------------------------------------------------------------------------------- other BASELINE configs

CONFIG_SCENES = {
    # name: (scene factory(nworlds, seed), dt, settle steps before timing)
    "chain": (lambda nw, seed: scenes.chain(nw, seed0=seed), 0.05, 40),
    "ragdoll": (lambda nw, seed: scenes.ragdoll(nw, seed0=seed), 0.01, 60),
    "pile64": (lambda nw, seed: scenes.pile(nworlds=nw, nbodies=64, seed=seed), 0.01, 155),
    "pile1000": (lambda nw, seed: scenes.pile(nbodies=1000, seed=seed), 0.01, 60),
    "wall_100k": (lambda nw, seed: scenes.wall(500, 200), 0.05, 0),
}


def _ref_path():
    ref = os.path.join(ROOT, "oracle", "_ref", "libode_ref_%s.so" % PREC)
    if os.path.exists(ref):
        return ref, "ref_", "reference"
    return os.path.join(ROOT, "oracle", "liborc_%s.so" % PREC), "orc_", "port"


def _cpu_scene_worker(args):
    """one process: nworlds worlds of a named scene on the CPU checker library, optional threaded stepper (reference only)"""
    path, prefix, name, nworlds, seed, steps, threads = args
    real = np.float32 if PREC == "single" else np.float64
    lib = B.SceneLib(path, prefix, real)
    mk, h, settle = CONFIG_SCENES[name]
    sc = mk(nworlds, seed)
    b = B.Batch(lib, sc)
    if prefix == "ref_":
        fn = lib.lib.ref_step_plain
        fn.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
        if threads:
            st = lib.lib.ref_set_threads
            st.argtypes = [C.c_void_p, C.c_int]
            st(b.h, threads)
        step = lambda n: fn(b.h, h, n, 0, sc.nworlds)   # noqa: E731
    else:
        step = lambda n: b.step(h, n)                   # noqa: E731
    if settle:
        step(settle)
    t = time.time()
    step(steps)
    dt = time.time() - t
    if prefix == "ref_" and threads:
        lib.lib.ref_set_threads(b.h, 0)
    return dt, sc.nworlds * sc.nbody


def cpu_config_baseline(name, nworlds, steps, procs=None, threads=0):
    """The reference's CPU path on one of the other configs: `procs` independent processes (batched configs: one per host core,
    each with its own worlds and its own dRand seed) or one process (single large worlds), bounded sample."""
    import multiprocessing as mp
    path, prefix, kind = _ref_path()
    procs = procs or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_scene_worker, [(path, prefix, name, nworlds, 1000 + 97 * i, steps, threads) for i in range(procs)])
    slowest = max(r[0] for r in res)
    bodies = sum(r[1] for r in res)
    return {"value": bodies * steps / slowest, "unit": "body-steps/s", "ms_per_step": slowest / steps * 1e3, "cores": procs if not threads else threads,
            "kind": kind, "threads_per_world": threads or 1,
            "sample": "%d process(es) x %d world(s) x %d steps of '%s' after %d settle steps%s" % (
                procs, nworlds, steps, name, CONFIG_SCENES[name][2], (", threaded stepper with a pool of %d (demo_crash.cpp:275-279)" % threads) if threads else "")}


def solver_roofline(L, batch, h, nbodies, step_ms, extra_steps=3):
    """Algorithmic bytes of the solver (per row per sweep 30 s + 16, per row once 46 s + 16, per body 26 s) over the
    CUDA-event duration of the solver launches of a step, against the measured HBM peak."""
    s = 4 if PREC == "single" else 8
    L.odeb_enable_timing(batch.h, 1)
    batch.step(h, extra_steps)
    n = C.c_int(0)
    sol_ms = L.odeb_solver_ms(batch.h, C.byref(n))
    L.odeb_enable_timing(batch.h, 0)
    tot = (C.c_uint64 * 6)()
    L.odeb_get_totals(batch.h, tot)
    pairs, contacts, rows, islands, sweeps, rowsweeps = [int(x) for x in tot]
    alg_bytes = rowsweeps * (30 * s + 16) + rows * (46 * s + 16) + nbodies * 26 * s
    sol_avg_ms = sol_ms / max(1, n.value)
    if sol_avg_ms <= 0:
        return None, (pairs, contacts, rows, islands, sweeps, rowsweeps)
    achieved = alg_bytes / (sol_avg_ms * 1e-3) / 1e9
    peak, which = peaks()
    L.odeb_solver_kernel.restype = C.c_char_p
    L.odeb_solver_kernel.argtypes = [C.c_void_p]
    r = {"bound": "hbm", "kernel": (L.odeb_solver_kernel(batch.h) or b"k_solve").decode(), "achieved": round(achieved, 1), "peak": peak,
         "peak_source": which, "unit": "GB/s", "frac": round(achieved / peak, 4), "algorithmic_bytes_per_launch": int(alg_bytes),
         "launch_ms": round(sol_avg_ms, 4), "share_of_step": round(sol_avg_ms / step_ms, 3) if step_ms else None,
         "row_sweeps_per_step": rowsweeps, "rows_per_step": rows}
    return r, (pairs, contacts, rows, islands, sweeps, rowsweeps)


def other_configs(slib, device, with_cpu=True, t_budget_end=None):
    """Device-timed throughput of the other BASELINE configs on this GPU (bounded: a few seconds each), each with the roofline of
    its solver launches and the reference's CPU number for the same scene beside it.  Not the headline line."""
    L = slib.lib
    out = {}
    ncores = os.cpu_count() or 1

    def timed(batch, h, steps, nbodies_total, roof=True):
        ms = C.c_double(0)
        if not L.odeb_timed_steps(batch.h, h, steps, FLUSH_BYTES, C.byref(ms)):
            raise RuntimeError("timed steps failed")
        tot = (C.c_uint64 * 6)()
        L.odeb_get_totals(batch.h, tot)
        r = {"ms_per_step": ms.value / steps, "body_steps_per_sec": nbodies_total * steps / (ms.value * 1e-3),
             "rows_last_step": int(tot[2]), "islands_last_step": int(tot[3]), "steps": steps}
        if roof:
            rf, _ = solver_roofline(L, batch, h, nbodies_total, ms.value / steps)
            if rf:
                r["roofline"] = rf
        return r

    deferred = []     # CPU legs run AFTER all GPU legs: the GPU drops to idle clocks while the host works, and a short GPU leg that
                      # starts right after a CPU leg is timed during the clock ramp (ragdoll: 0.84 instead of 0.57 ms per step)

    def cpu(name, nworlds, steps, procs, **kw):
        if not with_cpu:
            return None
        slot = {"pending": True}
        deferred.append((slot, name, nworlds, steps, procs, kw))
        return slot

    def run_deferred():
        for slot, name, nworlds, steps, procs, kw in deferred:
            slot.clear()
            if name == "wall_100k" and t_budget_end is not None and time.time() + 70 > t_budget_end:
                slot.update({"skipped": "time budget; measured once in this round: profiles/r2_cpu_baselines.txt"})
                continue
            try:
                slot.update(cpu_config_baseline(name, nworlds, steps, procs, **kw))
            except Exception as e:  # noqa: BLE001
                slot.update({"error": str(e)[:200]})

    try:   # the headline workload in DOUBLE precision (north_star: results match the reference built as dSINGLE and as dDOUBLE; the line's dtype is f32)
        from ode_b200 import load
        dl = load("double")
        Ld = dl.lib
        Ld.odeb_timed_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_size_t, C.POINTER(C.c_double)]
        Ld.odeb_solver_kernel.restype = C.c_char_p
        Ld.odeb_solver_kernel.argtypes = [C.c_void_p]
        b = B.Batch(dl, make_scene(WORLDS_PER_GPU, seed0=1000), device=device)
        b.step(H, 155)
        msd = C.c_double(0)
        if not Ld.odeb_timed_steps(b.h, H, 20, FLUSH_BYTES, C.byref(msd)):
            raise RuntimeError("odeb_timed_steps (double) failed")
        out["stack_double"] = {"ms_per_step": msd.value / 20, "body_steps_per_sec": WORLDS_PER_GPU * NBOX * 20 / (msd.value * 1e-3), "steps": 20, "dtype": "f64",
                               "kernel": (Ld.odeb_solver_kernel(b.h) or b"k_solve").decode(),
                               "workload": "%d worlds x %d-box stack (the headline workload) in double precision, all on this GPU" % (WORLDS_PER_GPU, NBOX)}
        b.close()
    except Exception as e:  # noqa: BLE001
        out["stack_double"] = {"error": str(e)[:200]}
    try:   # configs[2]: 65536 worlds of a 10-link ball-joint chain plus contacts (demo_chain2-style)
        nw = 65536
        b = B.Batch(slib, scenes.chain(nw), device=device)
        b.step(0.05, 40)
        r = timed(b, 0.05, 20, nw * 10)
        r["workload"] = "%d worlds x 10-link chain, dt=0.05 (BASELINE configs[2], all on this GPU)" % nw
        r["cpu_baseline"] = cpu("chain", 64, 100, ncores)
        out["chain"] = r
        b.close()
    except Exception as e:  # noqa: BLE001
        out["chain"] = {"error": str(e)[:200]}
    try:   # configs[3]: capsule ragdolls; 16384 worlds are quoted on 8 GPUs -> 2048 per GPU
        nw = 2048
        b = B.Batch(slib, scenes.ragdoll(nw), device=device)
        b.step(0.01, 60)
        r = timed(b, 0.01, 20, nw * 15)
        r["workload"] = "%d worlds x 15-capsule ragdoll (ball/hinge/universal joints with stops, friction contacts), dt=0.01 (BASELINE configs[3]: 16384 worlds / 8 GPUs)" % nw
        r["cpu_baseline"] = cpu("ragdoll", 16, 100, ncores)
        out["ragdoll"] = r
        b.close()
    except Exception as e:  # noqa: BLE001
        out["ragdoll"] = {"error": str(e)[:200]}
    for key, nw in (("pile64", 4096), ("pile64_16k", 16384)):
        try:   # north_star shape: batched 64-body worlds (piles of boxes and spheres, several islands per world)
            b = B.Batch(slib, scenes.pile(nworlds=nw, nbodies=64), device=device)
            b.step(0.01, 150)
            b.step(0.01, 5)
            r = timed(b, 0.01, 10, nw * 64)
            r["workload"] = "%d worlds x 64-body pile (boxes + spheres dropped on a plane, Approx1 friction), dt=0.01, steps 156-165 (north_star: batched 64-body worlds)" % nw
            if key == "pile64":
                r["cpu_baseline"] = cpu("pile64", 4, 20, ncores)
            out[key] = r
            b.close()
        except Exception as e:  # noqa: BLE001
            out[key] = {"error": str(e)[:200]}
    try:   # configs[0]: the single 1000-body pile (the reference's own CPU-runnable case), large-world path on the GPU
        sc = scenes.pile(nbodies=1000)
        b = B.Batch(slib, sc, device=device)
        b.set_solver_mode(1)
        b.step(0.01, 60)
        r = timed(b, 0.01, 20, sc.nbody, roof=False)
        r["workload"] = "1 world x 1000 boxes + spheres on a plane, dHashSpace semantics, dt=0.01, steps 61-80 (BASELINE configs[0]); GPU: ODEB_MODE_CANONICAL"
        r["cpu_baseline"] = cpu("pile1000", 1, 40, 1)
        if with_cpu:   # the reference's threaded stepper, pool of k threads (timing only: it changes the sweep order)
            r["cpu_threaded_stepper"] = {str(k): cpu("pile1000", 1, 10, 1, threads=k) for k in sorted({2, 4, 8, ncores}) if k <= ncores}
        out["pile1000"] = r
        b.close()
    except Exception as e:  # noqa: BLE001
        out["pile1000"] = {"error": str(e)[:200]}
    try:   # configs[4]: one 100k-box wall, sweep-and-prune space, large-island path (ODEB_MODE_CANONICAL)
        sc = scenes.wall(500, 200)
        b = B.Batch(slib, sc, device=device)
        b.set_solver_mode(1)
        b.step(0.05, 6)
        r = timed(b, 0.05, 6, sc.nbody)
        if "roofline" in r:
            r["roofline"]["bound"] = "latency (grid barrier between colours + the 12-row chain of a contact group), then hbm"
            # static: dram__bytes_read + write of one k_lwt_phase launch (ncu --set full on this workload, profiles/r2_ncu_summary.txt) x the launches of a step
            r["roofline"]["traffic"] = 5 * (2208631000 + 101905000)
            r["roofline"]["traffic_source"] = "static: ncu --set full capture of one phase launch on this workload (profiles/r2_ncu_summary.txt) x 5 phases; not measured in this run"
            if r["roofline"].get("launch_ms"):
                r["roofline"]["dram_frac"] = round(r["roofline"]["traffic"] / (r["roofline"]["launch_ms"] * 1e-3) / 1e9 / r["roofline"]["peak"], 4)
            r["roofline"]["limiter"] = ("one persistent launch per 8 sweeps, ~10 colours per sweep separated by grid barriers: ncu (profiles/r2_ncu_summary.txt) "
                                        "shows 48 % of the warp samples at the barrier and DRAM at ~30 % of peak; real DRAM traffic per row-sweep is the compact "
                                        "84-byte tile record, the algorithmic figure counts SURVEY's 136 bytes")
        r["workload"] = "1 world x %d bodies (500 x 200 brick wall + cannon ball), dSweepAndPruneSpace semantics, dt=0.05 (BASELINE configs[4])" % sc.nbody
        b.close()
        # the reference needs 20-45 s for ONE step of this world on one core: only when the time budget of the run allows it
        r["cpu_baseline"] = cpu("wall_100k", 1, 1, 1)
        out["wall_100k"] = r
    except Exception as e:  # noqa: BLE001
        out["wall_100k"] = {"error": str(e)[:200]}
    run_deferred()
    return out


# ---------------------------------------------------------------------------------------------- GPU arm

def gpu_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ngpu = world
    from ode_b200 import load
    slib = load(PREC)                      # raises if the CUDA extension is missing: no fallback
    L = slib.lib
    L.odeb_timed_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_size_t, C.POINTER(C.c_double)]
    L.odeb_launch_count.restype = C.c_uint64
    L.odeb_launch_count.argtypes = [C.c_void_p]
    L.odeb_get_totals.argtypes = [C.c_void_p, C.c_void_p]
    L.odeb_enable_timing.argtypes = [C.c_void_p, C.c_int]
    L.odeb_solver_ms.restype = C.c_double
    L.odeb_solver_ms.argtypes = [C.c_void_p, C.POINTER(C.c_int)]

    W = WORLDS_PER_GPU
    sc = make_scene(W, seed0=1000 + rank * W)
    batch = B.Batch(slib, sc, device=local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the stack needs ~150 steps to come to rest on its contacts; settle before measuring steady state
    batch.step(H, args.settle)
    batch.step(H, args.warmup)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    l0 = L.odeb_launch_count(batch.h)
    ms = C.c_double(0)
    ok = L.odeb_timed_steps(batch.h, H, args.steps, FLUSH_BYTES, C.byref(ms))
    if not ok:
        raise RuntimeError("timed steps failed")
    launches = int(L.odeb_launch_count(batch.h) - l0)
    barrier()
    sampler.stop_flag = True
    t = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = ngpu * W * NBOX * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the C-ABI with host buffers: per step H2D of per-body forces (the "actions"),
    #      the step, D2H of the full body state (the "observations")
    real = slib.real
    force = batch.alloc_host((W, NBOX, 3))            # the application's action / observation arrays live in page-locked host memory
    force[:] = 0
    force[:, -1, 0] = 0.01
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    st = batch.get_state(out=batch.alloc_state())     # observation buffers, reused every step

    def e2e_loop(gather=None):
        barrier()
        t0 = time.time()
        got = None
        for s in range(e2e_steps):
            batch.add_force(force=force)
            batch.step_async(H)               # queued behind the force upload; get_state below is the blocking call of the step
            if gather is not None:
                got = gather.launch()         # pack + NCCL all_gather of stats and observations from device buffers, side stream
            batch.get_state(out=st)
        if gather is not None:
            gather.wait()
        torch.cuda.synchronize()
        tt = torch.tensor([time.time() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return ngpu * W * NBOX * e2e_steps / float(tt.item()), got

    e2e_value, _ = e2e_loop()
    h2d = int(force.nbytes)
    d2h = int(sum(v.nbytes for v in st.values()))
    e2e_gather = None
    if world > 1 and not args.no_gather:
        # every rank receives the statistics and observations of ALL worlds each step (RL-style global observation),
        # gathered by NCCL straight from the pack kernel's device buffer on a side stream, overlapping the next step
        from ode_b200.shard import DeviceGather
        g = DeviceGather(batch, dist)
        e2e_loop(g)                                   # warm-up (NCCL communicator set-up)
        v, got = e2e_loop(g)
        e2e_gather = {"value": v, "unit": "body-steps/s", "gathered_bytes_per_step_per_rank": int(g.nbytes) * world,
                      "what": "odeb_pack_state_device + the 4 iteration counters of every world, all_gather_into_tensor (NCCL) on a side stream, double-buffered"}

    # ---- other BASELINE configs the way BASELINE states them for several GPUs: a fixed global batch split over the ranks
    #      (world w -> rank floor(w N / W), strong scaling), timed with a barrier on both sides, max over ranks
    strong = None
    if world > 1 and not args.no_extras:
        strong = {}
        for key, mkscene, gw, hh, settle, nb in (("chain", lambda n, o: scenes.chain(n, seed0=7 + o), 65536, 0.05, 40, 10),
                                                 ("ragdoll", lambda n, o: scenes.ragdoll(n, seed0=11 + o), 16384, 0.01, 60, 15)):
            lo, hi = (rank * gw) // world, ((rank + 1) * gw) // world
            bb = B.Batch(slib, mkscene(hi - lo, lo), device=local_rank)
            bb.step(hh, settle)
            barrier()
            msx = C.c_double(0)
            if not L.odeb_timed_steps(bb.h, hh, 20, FLUSH_BYTES, C.byref(msx)):
                raise RuntimeError("timed steps failed")
            barrier()
            tx = torch.tensor([msx.value], dtype=torch.float64, device="cuda")
            dist.all_reduce(tx, op=dist.ReduceOp.MAX)
            rf, _ = solver_roofline(L, bb, hh, (hi - lo) * nb, float(tx.item()) / 20)
            strong[key] = {"global_worlds": gw, "worlds_per_rank": hi - lo, "ms_per_step": float(tx.item()) / 20,
                           "body_steps_per_sec": gw * nb * 20 / (float(tx.item()) * 1e-3), "scaling": "strong",
                           "roofline_rank0": rf,
                           "workload": "BASELINE configs[%d]: %d worlds of the %s split over %d GPUs" % (2 if key == "chain" else 3, gw, key, world)}
            bb.close()

    # ---- roofline of the dominant kernel (k_solve): algorithmic bytes / measured launch duration
    out = None
    if rank == 0:
        roofline, (pairs, contacts, rows, islands, sweeps, rowsweeps) = solver_roofline(L, batch, H, W * NBOX, total_ms / args.steps, extra_steps=5)
        traffic, traffic_src = None, None
        pj = os.path.join(ROOT, "profiles", "solver_traffic.json")
        if os.path.exists(pj):
            try:
                tj = json.load(open(pj))
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = "static: ncu --set full capture of the same kernels on this workload (%s), not measured in this run" % tj.get("source", "profiles/solver_traffic.json")
            except Exception:
                traffic = None
        roofline["traffic"] = traffic
        roofline["traffic_source"] = traffic_src
        if traffic:
            roofline["dram_frac"] = round(traffic / (roofline["launch_ms"] * 1e-3) / 1e9 / roofline["peak"], 4)
        roofline["limiter"] = ("lsu/latency: the row stream is served from L2 and shared memory, ncu shows the SM's LSU / shared-memory pipe and the "
                               "per-world dependency chain as the limit, not DRAM (profiles/r2_*ncu*); `frac` is ALGORITHMIC bytes over the HBM peak, "
                               "`dram_frac` the real DRAM traffic over the same peak")
        extras = None
        if ngpu == 1 and not args.no_extras:
            batch.close()
            extras = other_configs(slib, local_rank, with_cpu=not args.no_cpu, t_budget_end=T_START + args.budget)
        cpu = None
        if ngpu == 1 and not args.no_cpu:
            cpu = cpu_reference(args.cpu_steps, args.settle + args.warmup)
        out = {
            "metric": METRIC, "value": value, "unit": "body-steps/s", "n_gpus": ngpu, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if PREC == "single" else "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "worlds_per_gpu": W, "bodies_per_world": NBOX, "precision": PREC, "settle_steps": args.settle,
                       "l2": "flushed before every timed step (256 MiB memset outside the event pairs)",
                       "per_step": {"pairs": pairs / W, "contacts": contacts / W, "rows": rows / W, "islands": islands / W,
                                    "sweeps_per_island": sweeps / max(1, islands)}},
            "e2e": {"value": e2e_value, "unit": "body-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "odeb_add_force + odeb_step_async + odeb_get_state (page-locked host buffers from odeb_alloc_host)"},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": roofline,
        }
        if e2e_gather is not None:
            out["e2e_with_gather"] = e2e_gather
        if cpu is not None:
            out["cpu_baseline"] = cpu
        if extras is not None:
            out["other_configs"] = extras
        if strong is not None:
            out["other_configs"] = strong
    batch.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    cpu = cpu_reference(min(4000, max(500, steps * 20)), args.settle + max(3, args.warmup))
    out = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "body-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if PREC == "single" else "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sampled_as": cpu["sample"]},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--settle", type=int, default=150)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the bounded runs of BASELINE configs[2..4]")
    ap.add_argument("--cpu-steps", type=int, default=1500)
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the e2e loop with the NCCL gather of stats + observations")
    ap.add_argument("--budget", type=float, default=330.0, help="seconds after which optional slow legs (reference on the 100k wall) are skipped")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
