"""The learner-facing step right after the rollout: returns / advantages over the [T, N] slab.

``compute_returns`` is rsl_rl's ``RolloutStorage.compute_returns`` (called at
``wheeledlab_rl/utils/modified_rsl_rl_runner.py:116``) [UPSTREAM-RECALL], run as one hand-written kernel directly on the
rollout-slab layout (``distributed.RolloutSlab``); advantage normalisation is left to the caller.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib


def compute_returns(rewards: torch.Tensor, values: torch.Tensor, last_values: torch.Tensor, dones: torch.Tensor,
                    gamma: float, lam: float, time_outs: torch.Tensor | None = None):
    """rewards/values [T,N] f32, last_values [N] f32, dones/time_outs [T,N] u8|bool (CUDA, contiguous) -> (returns, advantages)."""
    T, N = rewards.shape
    for x in (rewards, values, last_values, dones):
        if not (x.is_cuda and x.is_contiguous()):
            raise ValueError("compute_returns expects contiguous CUDA tensors")
    d8 = dones.view(torch.uint8) if dones.dtype == torch.bool else dones
    t8 = None if time_outs is None else (time_outs.view(torch.uint8) if time_outs.dtype == torch.bool else time_outs)
    ret, adv = torch.empty_like(rewards), torch.empty_like(rewards)
    stream = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
    check(lib.wl_gae(C.c_void_p(rewards.data_ptr()), C.c_void_p(values.data_ptr()), C.c_void_p(last_values.data_ptr()),
                     C.c_void_p(d8.data_ptr()), C.c_void_p(t8.data_ptr()) if t8 is not None else None, gamma, lam,
                     C.c_void_p(ret.data_ptr()), C.c_void_p(adv.data_ptr()), T, N, stream), "wl_gae")
    return ret, adv


class DataParallelAdam:
    """Adam for a (small) set of parameters replicated over the ranks of a process group, with the gradient all-reduce
    fused into the update kernel (``wl_dp_adam_step``): parameters and gradients live in flat buffers (``p.data`` / ``p.grad``
    are views), the flat gradient sits in torch symmetric memory so every rank reads its peers' gradients over NVLink inside
    the update kernel -- one launch per optimiser step, no NCCL call.  world_size 1 (or no process group) is plain fused Adam.
    Matches ``torch.optim.Adam`` (amsgrad off, decoupled weight decay off) on the rank-averaged gradient."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, group=None):
        import torch.distributed as dist
        self.params = [p for p in params]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.n, self.lr, self.betas, self.eps, self.wd, self.t = n, lr, betas, eps, weight_decay, 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.handle = None
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm
            group = group if group is not None else dist.group.WORLD
            self.flat_g = symm.empty(n, dtype=torch.float32, device=dev)
            self.flat_g.zero_()
            self.handle = symm.rendezvous(self.flat_g, group.group_name)
            ptrs = [int(x) for x in self.handle.buffer_ptrs]
        else:
            self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
            ptrs = [self.flat_g.data_ptr()]
        self._grad_ptrs = (C.c_void_p * len(ptrs))(*ptrs)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:                          # re-seat parameters and gradients as views of the flat buffers
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + k].view_as(p)
                p.grad = self.flat_g[off:off + k].view_as(p)
                off += k

    def zero_grad(self):
        self.flat_g.zero_()

    def step(self):
        self.t += 1
        stream = C.c_void_p(torch.cuda.current_stream(self.flat_p.device).cuda_stream)
        if self.handle is not None:
            self.handle.barrier()                          # every rank's gradient is complete
        check(lib.wl_dp_adam_step(C.c_void_p(self.flat_p.data_ptr()), C.c_void_p(self.m.data_ptr()), C.c_void_p(self.v.data_ptr()),
                                  self.world, self._grad_ptrs, self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.t, self.n,
                                  stream), "wl_dp_adam_step")
        if self.handle is not None:
            self.handle.barrier()                          # nobody overwrites a gradient a peer is still reading
