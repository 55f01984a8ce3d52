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
