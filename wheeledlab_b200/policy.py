"""The policy side of the rollout loop, fused in front of the env step.

``alg.act(obs)`` (modified_rsl_rl_runner.py:72) evaluates rsl_rl's ActorCritic [UPSTREAM-RECALL]: two 64x64 ELU MLPs
(drifting/config/agents/mushr/rsl_rl_ppo_cfg.py:12-17), a Gaussian head with a learned per-action std, and returns the
sampled action while the storage keeps value / mean / std / log-prob.  ``wl_act_step`` does all of that AND the env step in
one launch; this module packs torch ``nn.Linear`` weights into the kernel's blob layout and provides the graph-captured
rollout built on it.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import torch

from ._lib import WlPolicyOut, check, lib
from .distributed import RolloutSlab
from .sim import WheeledSim, _stream_ptr

HIDDEN = 64


def blob_layout(obs_dim: int):
    off = (C.c_int32 * 13)()
    total = lib.wl_policy_blob_floats(obs_dim, off)
    return list(off), total


def _linears(net):
    ls = [m for m in net.modules() if isinstance(m, torch.nn.Linear)]
    if len(ls) != 3 or ls[0].out_features != HIDDEN or ls[1].in_features != HIDDEN or ls[1].out_features != HIDDEN:
        raise NotImplementedError("wl_act_step implements the reference's [64, 64] ELU actor / critic only")
    return ls


def pack_actor_critic(actor, critic, std: torch.Tensor, obs_dim: int, device, out: torch.Tensor | None = None) -> torch.Tensor:
    """actor / critic: torch modules holding three nn.Linear (obs->64->64->2 / ->1, ELU in between); std: [2].
    Returns (or refills, when `out` is given -- e.g. after an optimiser step) the fp32 device blob wl_act_step reads."""
    off, total = blob_layout(obs_dim)
    blob = out if out is not None else torch.zeros(total, dtype=torch.float32, device=device)
    with torch.no_grad():
        for net_i, (net, nout) in enumerate(((actor, 2), (critic, 1))):
            ls = _linears(net)
            if ls[0].in_features != obs_dim or ls[2].out_features != nout:
                raise ValueError("network shape does not match the task")
            for k, lin in enumerate(ls):
                wt = lin.weight.detach().to(device=device, dtype=torch.float32).t().contiguous().reshape(-1)   # input-major
                o = off[net_i * 6 + 2 * k]
                blob[o:o + wt.numel()].copy_(wt)
                o = off[net_i * 6 + 2 * k + 1]
                blob[o:o + lin.bias.numel()].copy_(lin.bias.detach().to(device=device, dtype=torch.float32))
        blob[off[12]:off[12] + 2].copy_(std.detach().to(device=device, dtype=torch.float32))
    return blob


class PolicyBuffers:
    """[T, N] storage of what alg.act() records per step (rsl_rl RolloutStorage fields)."""

    def __init__(self, T: int, N: int, device):
        self.mean = torch.zeros((T, N, 2), dtype=torch.float32, device=device)
        self.log_prob = torch.zeros((T, N), dtype=torch.float32, device=device)
        self.values = torch.zeros((T, N), dtype=torch.float32, device=device)


def act_step(sim: WheeledSim, obs_in: torch.Tensor, blob: torch.Tensor, actions: torch.Tensor, mean: torch.Tensor,
             log_prob: torch.Tensor, value: torch.Tensor, out, log: torch.Tensor | None, step_counter: int):
    """One launch: value/mean/sample/log-prob of `obs_in` under `blob`, then env.step(sampled action) into `out`."""
    obs, rew, term, trunc = out
    po = WlPolicyOut(actions.data_ptr(), mean.data_ptr(), log_prob.data_ptr(), value.data_ptr())
    check(lib.wl_act_step(sim._h, C.c_void_p(obs_in.data_ptr()), C.c_void_p(blob.data_ptr()), po, C.c_void_p(obs.data_ptr()),
                          C.c_void_p(rew.data_ptr()), C.c_void_p(term.data_ptr()), C.c_void_p(trunc.data_ptr()),
                          C.c_void_p(log.data_ptr()) if log is not None else None, step_counter, _stream_ptr(sim.device)),
          "wl_act_step")


class FusedPolicyRollout:
    """T x wl_act_step captured as ONE CUDA graph: the reference's whole collection loop (modified_rsl_rl_runner.py:70-109)
    with nothing but T kernel launches inside.  The blob is read at replay time, so refilling it in place
    (pack_actor_critic(..., out=blob)) after each PPO update needs no re-capture."""

    def __init__(self, sim: WheeledSim, blob: torch.Tensor, T: int = 128, slab: RolloutSlab | None = None):
        self.sim, self.blob, self.T = sim, blob, T
        self.slab = slab or RolloutSlab(T, sim.num_envs, sim.obs_dim, 2, sim.device, policy_fields=True)
        # values / log-prob / mean go straight into the slab (= the all-gather send buffer) when it has those fields
        self.pol = (SimpleNamespace(values=self.slab.values, log_prob=self.slab.log_prob, mean=self.slab.mean)
                    if hasattr(self.slab, "values") else PolicyBuffers(T, sim.num_envs, sim.device))
        self.logs = torch.zeros((T, 16), dtype=torch.float32, device=sim.device)
        # slab.obs_in[k] = the observation actions[k] / log_prob[k] / values[k] / mean[k] were computed from (rsl_rl
        # RolloutStorage.observations[k]); slab.obs[k] = obs_in[k+1] = what step k returned
        self.obs0 = torch.empty((sim.num_envs, sim.obs_dim), dtype=torch.float32, device=sim.device)    # carried to the next replay
        self.graph = None
        self._base = 0
        self._stream = torch.cuda.Stream(device=sim.device)

    def _body(self):
        self.slab.obs_in[0].copy_(self.obs0)
        for k in range(self.T):
            act_step(self.sim, self.slab.obs_in[k], self.blob, self.slab.actions[k], self.pol.mean[k], self.pol.log_prob[k],
                     self.pol.values[k], self.slab.step_outputs(k), self.logs[k], WheeledSim.device_counter_plus(k))
        self.sim.advance_counter(self.T)
        self.obs0.copy_(self.slab.obs[self.T - 1])

    def capture(self, step_counter: int):
        sim = self.sim
        sim.observe(step_counter, 0, out=self.obs0)
        sim.set_step_counter(step_counter)
        self.graph = torch.cuda.CUDAGraph()
        s = self._stream
        s.wait_stream(torch.cuda.current_stream(sim.device))
        with torch.cuda.stream(s):
            with torch.cuda.graph(self.graph, stream=s):
                self._body()
        torch.cuda.current_stream(sim.device).wait_stream(s)
        sim.set_step_counter(step_counter)       # (the captured advance_counter moved the HOST mirror; put both back)
        self._base = step_counter
        return self

    def run(self) -> RolloutSlab:
        self.graph.replay()
        self._base += self.T
        self.sim.note_device_counter(self._base)
        return self.slab
