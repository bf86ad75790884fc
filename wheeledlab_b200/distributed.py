"""Multi-GPU sharding of the env set -- one process per GPU, torch.distributed plumbing.

Envs never interact (env_spacing = 0 with inter-env collisions filtered, mushr_drift_env_cfg.py:373), so rank r
owns global env ids [r*N_local, (r+1)*N_local); RNG and domain randomisation are keyed by the GLOBAL id, which makes
the concatenation of the shards bit-identical to one big run.  There is no data-path collective in the step.  The
only exchange is the learner-facing ALL-GATHER of the rollout slab once per PPO iteration (rsl_rl RolloutStorage
consumes [T, N, ...]); the step kernel writes observations / rewards / dones straight into the slab rows, so the
slab IS the send buffer (no staging copy).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_offset(rank: int, envs_per_rank: int) -> int:
    return rank * envs_per_rank


class RolloutSlab:
    """Per-rank rollout buffer in the layout the learner consumes: obs [T,N,D] f32, actions [T,N,A] f32,
    rewards [T,N] f32, dones [T,N] u8 (terminated | truncated << 1).  One contiguous allocation per field so a single
    all_gather_into_tensor per field moves it; fields are views of ONE flat byte buffer => one collective total."""

    def __init__(self, T: int, n_local: int, obs_dim: int, act_dim: int, device):
        self.T, self.n, self.obs_dim, self.act_dim = T, n_local, obs_dim, act_dim
        self.device = torch.device(device)
        f = T * n_local
        self._sizes = {"obs": f * obs_dim * 4, "actions": f * act_dim * 4, "rewards": f * 4, "terminated": f, "truncated": f}
        total = sum((v + 255) // 256 * 256 for v in self._sizes.values())
        self.flat = torch.zeros(total, dtype=torch.uint8, device=self.device)
        off, v = 0, {}
        for k, nbytes in self._sizes.items():
            v[k] = self.flat[off: off + nbytes]
            off += (nbytes + 255) // 256 * 256
        self.obs = v["obs"].view(torch.float32).view(T, n_local, obs_dim)
        self.actions = v["actions"].view(torch.float32).view(T, n_local, act_dim)
        self.rewards = v["rewards"].view(torch.float32).view(T, n_local)
        self.terminated = v["terminated"].view(T, n_local)
        self.truncated = v["truncated"].view(T, n_local)

    @property
    def nbytes(self) -> int:
        return self.flat.numel()

    def step_outputs(self, t: int):
        """Row t as the (obs, rew, terminated, truncated) tuple WheeledSim.step(out=...) writes into."""
        return self.obs[t], self.rewards[t], self.terminated[t], self.truncated[t]

    def all_gather(self, group=None) -> "GatheredRollout":
        """ONE collective: every rank receives every rank's slab (NCCL over NVLink on GPUs, gloo in CPU tests).
        The receive buffer is allocated once and reused: the returned views are overwritten by the next gather (a fresh
        cudaMalloc of world x slab bytes per iteration costs tens of ms)."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if getattr(self, "_recv", None) is None or self._recv.numel() != world * self.flat.numel():
            self._recv = torch.empty(world * self.flat.numel(), dtype=torch.uint8, device=self.device)
        out = self._recv
        if world == 1:
            out.copy_(self.flat)
        else:
            dist.all_gather_into_tensor(out, self.flat, group=group)
        return GatheredRollout(out.view(world, -1), self)


class GatheredRollout:
    """Views [world, T, N_local, ...] of the gathered bytes; .cat(name) gives the learner's [T, world*N_local, ...]
    with global env id = rank*N_local + local id."""

    def __init__(self, buf: torch.Tensor, proto: RolloutSlab):
        self.buf, self.p = buf, proto

    def field(self, name: str) -> torch.Tensor:
        p = self.p
        off = 0
        for k, nbytes in p._sizes.items():
            if k == name:
                raw = self.buf[:, off: off + nbytes]
                break
            off += (nbytes + 255) // 256 * 256
        else:
            raise KeyError(name)
        w = self.buf.shape[0]
        if name == "obs":
            return raw.view(torch.float32).view(w, p.T, p.n, p.obs_dim)
        if name == "actions":
            return raw.view(torch.float32).view(w, p.T, p.n, p.act_dim)
        if name == "rewards":
            return raw.view(torch.float32).view(w, p.T, p.n)
        return raw.view(w, p.T, p.n)

    def cat(self, name: str) -> torch.Tensor:
        f = self.field(name)                       # [W, T, N, ...] -> [T, W*N, ...]
        return f.transpose(0, 1).reshape(f.shape[1], f.shape[0] * f.shape[2], *f.shape[3:])
