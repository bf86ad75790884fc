"""Policy-in-the-loop rollouts as ONE CUDA graph (the caller side of the hot path).

The reference's rollout loop (modified_rsl_rl_runner.py:70-109) alternates ``alg.act(obs)`` and ``env.step(actions)``
T = 128 times with host syncs in between.  Here the step carries no per-step host value (step k of the graph runs on the
device-resident counter base + k, the base advances by T in the graph's last node; on-device curriculum, in-kernel episode
log), so T x (policy -> step -> slab row) is captured once and replayed: one graph launch per PPO iteration, results
bit-identical to the eager loop.  Slab alignment: ``slab.obs_in[k]`` is the observation ``actions[k]`` was computed from
(rsl_rl RolloutStorage.observations[k]); ``slab.obs[k]`` is what step k returned.
"""
from __future__ import annotations

import torch

from .distributed import RolloutSlab
from .sim import WheeledSim


class GraphedRollout:
    def __init__(self, sim: WheeledSim, policy, T: int = 128, slab: RolloutSlab | None = None):
        """policy: callable obs[N,D] -> actions[N,A]; must be CUDA-graph capturable (no host sync, static shapes)."""
        self.sim, self.policy, self.T = sim, policy, T
        self.slab = slab or RolloutSlab(T, sim.num_envs, sim.obs_dim, 2, sim.device)
        self.logs = torch.zeros((T, 16), dtype=torch.float32, device=sim.device)
        # the observation the next replay starts from; copied into slab.obs_in[0] first thing in the graph, so that after a
        # replay the slab is self-consistent (obs_in[k] is what actions[k] were computed from, for every k)
        self.obs0 = torch.empty((sim.num_envs, sim.obs_dim), dtype=torch.float32, device=sim.device)
        self.graph = None
        self._base = 0
        self._stream = torch.cuda.Stream(device=sim.device)

    def _body(self):
        self.slab.obs_in[0].copy_(self.obs0)
        for k in range(self.T):
            act = self.policy(self.slab.obs_in[k])
            self.slab.actions[k].copy_(act)
            self.sim.step(self.slab.actions[k], WheeledSim.device_counter_plus(k), out=self.slab.step_outputs(k), log=self.logs[k])
        self.sim.advance_counter(self.T)
        self.obs0.copy_(self.slab.obs[self.T - 1])   # next iteration continues from the last observation

    def capture(self, step_counter: int):
        """Prime obs0 with get_observations(), align the device counter, warm up the policy, capture the graph."""
        sim = self.sim
        sim.observe(step_counter, 0, out=self.obs0)
        snapshot = sim.state_snapshot()
        s = self._stream
        s.wait_stream(torch.cuda.current_stream(sim.device))
        with torch.cuda.stream(s):
            for _ in range(2):                   # warm-up (allocator, cuBLAS handles) on the side stream, then roll back
                sim.set_step_counter(step_counter)
                self._body()
        torch.cuda.current_stream(sim.device).wait_stream(s)
        sim.load_state(snapshot)
        sim.observe(step_counter, 0, out=self.obs0)
        sim.set_step_counter(step_counter)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            with torch.cuda.graph(self.graph, stream=s):
                self._body()
        torch.cuda.current_stream(sim.device).wait_stream(s)
        # capture does not execute: state, counter and obs0 are still those of `step_counter`
        sim.set_step_counter(step_counter)       # (the captured advance_counter moved the HOST mirror; put both back)
        self._base = step_counter
        return self

    def run(self) -> RolloutSlab:
        """Replay T steps; returns the slab (views are overwritten by the next run)."""
        self.graph.replay()
        self._base += self.T
        self.sim.note_device_counter(self._base)     # the replay advanced the device base behind the library's back
        return self.slab
