"""Multi-GPU sharding of batched worlds: one process per GPU, contiguous blocks of worlds per rank
(world w -> rank floor(w * G / W)), no collective on the step path (worlds share nothing, each has its own
dRand seed). torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is used only to
gather per-world statistics / observations after a step."""
import numpy as np
from . import _binding as B


def shard_range(nworlds, rank, world_size):
    """[begin, end) of the worlds owned by `rank`: world w belongs to rank floor(w * G / W)."""
    begin = (rank * nworlds + world_size - 1) // world_size
    end = ((rank + 1) * nworlds + world_size - 1) // world_size
    return begin, end


def slice_scene(scene, begin, end):
    """The sub-batch [begin, end) of a Scene (template shared, per-world state and seeds sliced)."""
    sub = B.Scene(scene.wp, end - begin)
    sub.bodies, sub.body_pos, sub.body_quat = scene.bodies, scene.body_pos, scene.body_quat
    sub.geoms, sub.joints = scene.geoms, scene.joints
    if scene.state is not None:
        sub.state = {k: np.ascontiguousarray(v[begin:end]) for k, v in scene.state.items()}
    if scene.seeds is not None:
        sub.seeds = np.ascontiguousarray(scene.seeds[begin:end])
    return sub


class ShardedBatch:
    """The rank-local part of a global batch + gathers over the process group."""

    def __init__(self, slib, scene, rank, world_size, device=0):
        self.rank, self.world_size = rank, world_size
        self.global_worlds = scene.nworlds
        self.begin, self.end = shard_range(scene.nworlds, rank, world_size)
        self.local = B.Batch(slib, slice_scene(scene, self.begin, self.end), device=device)

    def step(self, h, nsteps=1):
        self.local.step(h, nsteps)

    def local_stats(self):
        return np.stack([self.local.get_stats(w) for w in range(self.local.W)]).astype(np.int64)

    def gather_stats(self, dist, device="cpu"):
        """all_gather of the four per-world dynamic-iteration counters -> [W_global, 4] on every rank."""
        import torch
        mine = torch.from_numpy(self.local_stats()).to(device)
        sizes = [shard_range(self.global_worlds, r, self.world_size) for r in range(self.world_size)]
        mx = max(e - b for b, e in sizes)
        pad = torch.zeros((mx, 4), dtype=torch.int64, device=device)
        pad[: mine.shape[0]] = mine
        out = [torch.zeros_like(pad) for _ in range(self.world_size)]
        dist.all_gather(out, pad)
        return np.concatenate([o[: e - b].cpu().numpy() for o, (b, e) in zip(out, sizes)])

    def gather_observations(self, dist, device="cpu"):
        """all_gather of (pos, quat, lvel, avel) = 13 reals per body -> [W_global, NB, 13] on every rank."""
        import torch
        st = self.local.get_state()
        obs = np.concatenate([st["pos"], st["quat"], st["lvel"], st["avel"]], axis=-1)
        mine = torch.from_numpy(obs).to(device)
        sizes = [shard_range(self.global_worlds, r, self.world_size) for r in range(self.world_size)]
        mx = max(e - b for b, e in sizes)
        pad = torch.zeros((mx,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=device)
        pad[: mine.shape[0]] = mine
        out = [torch.zeros_like(pad) for _ in range(self.world_size)]
        dist.all_gather(out, pad)
        return np.concatenate([o[: e - b].cpu().numpy() for o, (b, e) in zip(out, sizes)])


class _DevBuf:
    """a raw device address as a __cuda_array_interface__ object (torch.as_tensor wraps it without a copy)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class DeviceGather:
    """the per-step gather of statistics and observations WITHOUT a host round trip.

    The batch is put on a torch stream (odeb_set_stream); after a step `launch()` packs the body state on that stream
    (odeb_pack_state_device: one kernel, Real4 SoA -> [pos | quat | lvel | avel] per body), copies it together with the four
    per-world dynamic-iteration counters into one of two send buffers, and issues `all_gather_into_tensor` (NCCL over NVLink) on
    a side stream that only waits for that copy -- the next step's kernels run on the main stream meanwhile.  `launch()` returns
    the receive buffer of this step ([world_size, bytes] uint8); `wait()` makes the host wait for the outstanding gathers.
    Worlds never exchange data on the step path; this is the only collective."""

    def __init__(self, batch, dist, main_stream=None):
        import ctypes as C
        import torch
        self.torch, self.dist, self.batch = torch, dist, batch
        L = batch.slib.lib
        L.odeb_pack_state_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.odeb_device_counters.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.odeb_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        self.L = L
        self.main = main_stream or torch.cuda.Stream()
        self.side = torch.cuda.Stream()
        if not L.odeb_set_stream(batch.h, C.c_void_p(self.main.cuda_stream)):
            raise RuntimeError("odeb_set_stream failed")
        ptr, nb = C.c_void_p(), C.c_size_t()
        if not L.odeb_pack_state_device(batch.h, C.byref(ptr), C.byref(nb)):
            raise RuntimeError("odeb_pack_state_device failed")
        st, sd = C.c_void_p(), C.c_void_p()
        L.odeb_device_counters(batch.h, C.byref(st), C.byref(sd))
        self.obs_bytes, self.stat_bytes = int(nb.value), batch.W * 16
        self.obs_view = torch.as_tensor(_DevBuf(ptr.value, self.obs_bytes), device="cuda")
        self.stat_view = torch.as_tensor(_DevBuf(st.value, self.stat_bytes), device="cuda")
        n = self.obs_bytes + self.stat_bytes
        ws = dist.get_world_size() if dist.is_initialized() else 1
        self.world_size, self.nbytes = ws, n
        self.send = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
        self.recv = [torch.empty(ws * n, dtype=torch.uint8, device="cuda") for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.k = 0
        self.main.synchronize()

    def launch(self):
        import ctypes as C
        torch, k = self.torch, self.k & 1
        with torch.cuda.stream(self.main):
            if self.k >= 2:
                self.main.wait_event(self.done[k])            # the gather that last read this send buffer
            if not self.L.odeb_pack_state_device(self.batch.h, None, None):
                raise RuntimeError("odeb_pack_state_device failed")
            self.send[k][: self.obs_bytes].copy_(self.obs_view, non_blocking=True)
            self.send[k][self.obs_bytes:].copy_(self.stat_view, non_blocking=True)
            self.ready[k].record(self.main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ready[k])
            if self.world_size > 1 or self.dist.is_initialized():
                self.dist.all_gather_into_tensor(self.recv[k], self.send[k])
            else:
                self.recv[k].copy_(self.send[k], non_blocking=True)
            self.done[k].record(self.side)
        self.k += 1
        return self.recv[k].view(self.world_size, self.nbytes)

    def wait(self):
        self.side.synchronize()

    def unpack(self, row):
        """(pos, quat, lvel, avel, stats) of one rank's part of a received buffer, as host numpy arrays"""
        real = self.batch.slib.real
        W, NB = self.batch.W, self.batch.NB
        raw = row.cpu().numpy()
        obs = raw[: self.obs_bytes].view(real)
        n = W * NB
        return (obs[: 3 * n].reshape(W, NB, 3), obs[3 * n: 7 * n].reshape(W, NB, 4), obs[7 * n: 10 * n].reshape(W, NB, 3),
                obs[10 * n: 13 * n].reshape(W, NB, 3), raw[self.obs_bytes:].view(np.uint32).reshape(W, 4))
