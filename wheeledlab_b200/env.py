"""ManagerBasedRLEnv-compatible environment backed by the fused CUDA step.

Mirrors the surface of ``isaaclab.envs:ManagerBasedRLEnv`` that the reference's callers touch
(SURVEY.md 8b): ``train_rl.py:70-116``, ``modified_rsl_rl_runner.py:46-109``,
``create_and_step_env.py:26-44`` and the term code in ``wheeledlab/envs/mdp`` /
``wheeledlab_tasks``.  Same names, same argument meaning, same error behaviour (exceptions; no
silent CPU path).  Step ordering follows SURVEY.md 3.3 (A..I).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch

from .sim import WheeledSim
from .tasks import TaskSpec, make_task


class _Box:
    """Minimal gymnasium.spaces.Box stand-in (callers assign .low/.high, train_rl.py:73-74)."""

    def __init__(self, low, high, shape, dtype=torch.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def sample(self, generator=None, device="cpu"):
        return torch.randn(self.shape, generator=generator, device=device)


class _LazyLog(dict):
    """extras["log"]: keys are known up front, values are 0-dim views of the device row of that step; the views are only
    created when somebody reads them (the train loop appends the dict every step and reads it once per iteration,
    modified_rsl_rl_runner.py:95-98).  The row of step t is written by the next launch on the handle (no grid-wide sync in
    the step): reading the log of the most recent step publishes it first (wl_log_flush, a 32-thread kernel)."""

    def __init__(self, row: torch.Tensor, index: dict, extra: dict | None = None, env=None):
        super().__init__((k, None) for k in index)
        self._row, self._index = row, index
        self._env, self._step = env, (env.common_step_counter if env is not None else None)
        self._extra = extra or {}                # entries of host-side (Python) terms: already 0-dim device tensors
        for k in self._extra:
            dict.__setitem__(self, k, None)

    def _ready(self):
        # the row of step t is written by the next launch on the handle: if this is still the most recent step, flush now
        env = self._env
        if env is not None and env.common_step_counter == self._step:
            env.sim.flush_log()
        self._env = None

    def __getitem__(self, k):
        if k in self._extra:
            return self._extra[k]
        if self._env is not None:
            self._ready()
        return self._row[self._index[k]]

    def get(self, k, default=None):
        if k in self._extra:
            return self._extra[k]
        if k not in self._index:
            return default
        return self[k]

    def items(self):
        return [(k, self[k]) for k in list(self._index) + list(self._extra)]

    def values(self):
        return [self[k] for k in list(self._index) + list(self._extra)]


class RewardTermCfgView:
    """What reward_manager.get_term_cfg returns: an object with a mutable ``weight`` (curriculums.py:33-35)."""

    def __init__(self, name: str, weight: float):
        self.name, self.weight = name, weight
        self.params = {}


class RewardManager:
    def __init__(self, env):
        self._env = env
        self.active_terms = list(env.spec.reward_names)

    def _slot(self, name):
        try:
            return self.active_terms.index(name)
        except ValueError:
            raise ValueError(f"Reward term '{name}' not found.") from None

    def get_term_cfg(self, name) -> RewardTermCfgView:
        return RewardTermCfgView(name, float(self._env.sim.rew_weight[self._slot(name)].item()))

    def set_term_cfg(self, name, cfg) -> None:
        self._env.sim.rew_weight[self._slot(name)] = float(cfg.weight)

    @property
    def weights(self) -> torch.Tensor:
        return self._env.sim.rew_weight[: len(self.active_terms)]


class CommandManager:
    """command_manager.get_command("goal_pose") -> [N,4] (pos_b xyz, heading_b), UniformPose2dCommand layout."""

    def __init__(self, env):
        self._env = env

    def get_command(self, name):
        from .sim import G_CMDB
        g = self._env.sim.groups[G_CMDB]
        out = torch.zeros((self._env.num_envs, 4), device=self._env.device)
        out[:, 0:2] = g[:, 0:2]
        out[:, 3] = g[:, 2]
        return out


class TerminationManager:
    def __init__(self, env):
        self._env = env
        self.active_terms = [n for n, _ in env.spec.termination_names]
        self.terminated = torch.zeros(env.num_envs, dtype=torch.bool, device=env.device)
        self.time_outs = torch.zeros(env.num_envs, dtype=torch.bool, device=env.device)

    @property
    def dones(self):
        return self.terminated | self.time_outs

    def get_term(self, name):
        """The term's OWN mask for the last step (IsaacLab TerminationManager.get_term), not the union: from the per-term bits
        the step kernel writes when the task has more than one non-time-out term (Elevation: cart_out_of_bounds, stuck,
        rollover, at_goal); host-side (Python) terms keep their own masks."""
        env = self._env
        if name in getattr(env, "_py_term_masks", {}):
            return env._py_term_masks[name]
        names = env.spec.termination_names
        for j, (n, is_to) in enumerate(names):
            if n == name:
                bits = getattr(env, "_term_bits_staged", None) if (env._py_rewards or env._py_terms) else getattr(env, "_term_bits", None)
                if bits is not None:
                    return ((bits >> j) & 1).to(torch.bool)
                others = [m for m, to in names if to == is_to]
                if len(others) == 1:                       # the only term of its kind: its mask is the union
                    return self.time_outs if is_to else self.terminated
                raise RuntimeError(f"termination term {name!r}: per-term masks are not being recorded for this env")
        raise ValueError(f"Termination term '{name}' not found.")


class ActionManager:
    def __init__(self, env):
        self._env = env
        self.total_action_dim = env.spec.action_dim

    @property
    def action(self):
        return self._env.sim.last_action

    @property
    def prev_action(self):
        return self._env.sim.prev_action


class ObservationManager:
    def __init__(self, env):
        self._env = env
        self.group_obs_dim = {"policy": (env.spec.obs_dim,)}
        self._calls = 0

    def compute(self):
        """Re-samples the observation noise, like the reference."""
        env = self._env
        obs = env.sim.observe(env.common_step_counter, self._calls)
        self._calls += 1
        return {"policy": env._append_py_obs(obs)}


class _RobotData:
    def __init__(self, sim: WheeledSim):
        self._sim = sim

    root_pos_w = property(lambda s: s._sim.root_pos_w)
    root_quat_w = property(lambda s: s._sim.root_quat_w)
    root_lin_vel_w = property(lambda s: s._sim.root_lin_vel_w)
    root_ang_vel_w = property(lambda s: s._sim.root_ang_vel_w)
    root_link_ang_vel_w = property(lambda s: s._sim.root_ang_vel_w)

    @staticmethod
    def _rot_inv(q, v):
        w, xyz = q[:, 0:1], q[:, 1:4]
        t = 2.0 * torch.cross(xyz, v, dim=-1)
        return v - w * t + torch.cross(xyz, t, dim=-1)

    @property
    def root_lin_vel_b(self):
        return self._rot_inv(self.root_quat_w, self.root_lin_vel_w)

    @property
    def root_ang_vel_b(self):
        return self._rot_inv(self.root_quat_w, self.root_ang_vel_w)

    @property
    def joint_vel(self):
        n = self._sim.num_envs
        out = torch.zeros((n, 10), device=self._sim.device)
        out[:, 0:4] = self._sim.wheel_vel
        out[:, 4:6] = self._sim.steer_vel
        out[:, 6:10] = self._sim.suspension_state()[1]
        return out

    @property
    def joint_pos(self):
        n = self._sim.num_envs
        out = torch.zeros((n, 10), device=self._sim.device)     # wheel angles are not tracked (not needed by any term)
        out[:, 4:6] = self._sim.steer_pos
        out[:, 6:10] = self._sim.suspension_state()[0]
        return out


class _Robot:
    """Articulation stand-in for env.scene["robot"] (write_root_* used by events.py:132-133)."""

    def __init__(self, env):
        self._env = env
        self.data = _RobotData(env.sim)
        self.joint_names = list(env.spec.joint_names)

    def find_joints(self, name_keys, joint_subset=None, preserve_order=False):
        import re
        keys = [name_keys] if isinstance(name_keys, str) else list(name_keys)
        ids, names = [], []
        for i, jn in enumerate(self.joint_names):
            if any(re.fullmatch(k, jn) for k in keys):
                ids.append(i)
                names.append(jn)
        if not ids:
            raise ValueError(f"Not all regular expressions are matched! {keys} vs {self.joint_names}")
        return ids, names

    def write_root_pose_to_sim(self, root_pose, env_ids=None):
        sim = self._env.sim
        ids = slice(None) if env_ids is None else env_ids
        sim.root_pos_w[ids] = root_pose[:, 0:3]
        sim.root_quat_w[ids] = root_pose[:, 3:7]

    def write_root_velocity_to_sim(self, root_velocity, env_ids=None):
        sim = self._env.sim
        ids = slice(None) if env_ids is None else env_ids
        sim.root_lin_vel_w[ids] = root_velocity[:, 0:3]
        sim.root_ang_vel_w[ids] = root_velocity[:, 3:6]


class _Scene:
    def __init__(self, env):
        self._assets = {"robot": _Robot(env)}
        self.env_origins = torch.zeros((env.num_envs, 3), device=env.device)   # env_spacing = 0 (:373)
        self.num_envs = env.num_envs
        self.sensors = {}

    def __getitem__(self, key):
        return self._assets[key]


class ManagerBasedRLEnv:
    """Drop-in for ``isaaclab.envs.ManagerBasedRLEnv`` on the Drift/Elevation/Visual gym ids."""

    _RING = 512      # step() / step_host() output buffers are reused after this many steps (fewer when rows are wide, see _ring_len)
    _HOST_OBS_RING = 4   # pinned host observation buffers of step_host(host_obs=True)

    metadata = {"render_modes": [None, "human", "rgb_array"], "isaac_sim_version": "b200-native"}

    def __init__(self, cfg: TaskSpec | str = "Isaac-MushrDriftRL-v0", render_mode=None, device="cuda:0", **task_kw):
        self._task_spec: TaskSpec = make_task(cfg, **task_kw) if isinstance(cfg, str) else cfg
        self.cfg = SimpleNamespace(is_finite_horizon=False, seed=int(self.spec.cfg.seed), spec=self.spec)
        self.render_mode = render_mode
        self.device = str(torch.device(device))
        self.sim = WheeledSim(self.spec, device)     # (the elevation height-field travels in spec.heightfield)
        self.num_envs = self.sim.num_envs
        self.step_dt = self.spec.step_dt
        self.physics_dt = float(self.spec.cfg.sim_dt)
        self.max_episode_length_s = self.spec.episode_length_s
        self.max_episode_length = int(self.spec.cfg.max_episode_length)   # ceil(episode_length_s/(dt*decimation)) in doubles
        self.common_step_counter = 0
        self.extras = {}
        self.scene = _Scene(self)
        self.action_manager = ActionManager(self)
        self.observation_manager = ObservationManager(self)
        self.reward_manager = RewardManager(self)
        self.termination_manager = TerminationManager(self)
        self.command_manager = CommandManager(self)
        self.single_action_space = _Box(-math.inf, math.inf, (self.spec.action_dim,))
        self.action_space = _Box(-math.inf, math.inf, (self.num_envs, self.spec.action_dim))
        self.single_observation_space = {"policy": _Box(-math.inf, math.inf, (self.spec.obs_dim,))}
        self.observation_space = {"policy": _Box(-math.inf, math.inf, (self.num_envs, self.spec.obs_dim))}
        self.log_episode_info = True
        self._log_index = {"Episode_Reward/" + n: k for k, n in enumerate(self.spec.reward_names)}
        self._log_index.update({"Episode_Termination/" + n: 9 + j for j, (n, _) in enumerate(self.spec.termination_names)})
        self._ring_len = max(8, min(self._RING, (2 << 30) // max(1, 4 * self.num_envs * self.spec.obs_dim)))
        self._step_ring = None
        self._host_io = None
        self._pinned_ptrs = {}                   # data_ptr -> c_void_p of caller buffers known to be pinned [N,2] f32
        self._ring = None
        self.host_transport = "zero_copy"        # or "copy": staged H2D / D2H copies (wl_step_host)
        self.host_obs = False                    # step_host: True -> observations are returned in pinned host memory too
        self._host_obs_ring = None
        # host-side (Python) MDP terms: evaluated between the two halves of the staged step (see add_reward_term)
        self._py_rewards, self._py_terms, self._py_obs = [], [], []
        self._py_sums = {}
        for name, func, weight, params in self.spec.python_reward_terms:
            self.add_reward_term(name, func, weight, params)
        for name, func, time_out, params in self.spec.python_termination_terms:
            self.add_termination_term(name, func, time_out, params)
        # per-term termination masks (TerminationManager.get_term): recorded by the step kernel when a union is not enough
        self._term_bits = None
        if sum(1 for _, to in self.spec.termination_names if not to) > 1:
            import ctypes as C
            from ._lib import check, lib
            self._term_bits = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
            check(lib.wl_set_term_bits(self.sim._h, C.c_void_p(self._term_bits.data_ptr())), "wl_set_term_bits")
        # event_manager.apply(mode="startup")
        self.sim.startup()
        self._needs_reset = True

    # -- gym surface ---------------------------------------------------------------------------
    @property
    def spec(self) -> TaskSpec:
        """The task description (TaskSpec).  gymnasium's make() assigns ITS EnvSpec to `env.spec`: that one is kept as
        `gym_spec` so that the assignment cannot clobber the task description the env works from."""
        return self._task_spec

    @spec.setter
    def spec(self, value):
        if isinstance(value, TaskSpec):
            self._task_spec = value
        else:
            self.gym_spec = value

    @property
    def unwrapped(self):
        return self

    @property
    def episode_length_buf(self):
        return self.sim.episode_length_buf

    @episode_length_buf.setter
    def episode_length_buf(self, value):
        # modified_rsl_rl_runner.py:46-49 assigns a randint tensor
        self.sim.episode_length_buf.copy_(value.to(torch.int32))

    def seed(self, seed: int = -1) -> int:
        """ManagerBasedEnv.seed: re-key the counter-based generator for every draw from now on (resets, pushes, observation
        noise, commands).  What startup already sampled (DR scatter, the reference-pose table) is not re-drawn -- the reference
        seeds before construction (hydra.py:30 overrides cfg.seed with the agent seed; pass seed= to make()/the task for that)."""
        if seed is not None and seed >= 0:
            self.sim.set_seed(int(seed))
        return int(self.spec.cfg.seed)

    def close(self):
        self.sim.close()

    def render(self, recompute=False):
        """No viewport: the reference renders through Omniverse RTX (video recording only, train_rl.py --video); returns None."""
        return None

    def reset(self, seed=None, options=None):
        self.sim.reset(None, self.common_step_counter)
        self._needs_reset = False
        obs = self.observation_manager.compute()
        return obs, self.extras

    def get_observations(self):
        """RslRlVecEnvWrapper.get_observations."""
        obs = self.observation_manager.compute()
        return obs["policy"], {"observations": obs}

    def _curriculum_fire_mask(self) -> int:
        """Counter conditions of increase_reward_weight_over_time (curriculums.py:23-35) -- the host-side restatement the
        golden-vector test checks; at run time the same conditions are evaluated on the device (log_finalize)."""
        c, L = self.common_step_counter, self.max_episode_length
        if c % L != 0:
            return 0
        mask = 0
        num_episodes = c // L
        for k, t in enumerate(self.spec.curriculum):
            num_increases = num_episodes // t.episodes_per_increase
            if num_increases > t.max_increases:
                continue
            if (num_episodes + 1) % t.episodes_per_increase == 0:
                mask |= 1 << k
        return mask

    # -- host-side (Python) MDP terms: the reference's extension mechanism -----------------------------------------
    def add_reward_term(self, name: str, func, weight: float, params: dict | None = None):
        """Register ``func(env, **params) -> [N]`` as a reward term (IsaacLab RewardTermCfg semantics: contributes
        func * weight * step_dt, accumulated per episode and logged as Episode_Reward/<name>).  It runs on the
        post-physics, PRE-reset state, exactly where ManagerBasedRLEnv.step calls the reward manager: the env switches
        to the staged step (wl_step_stage_a -> Python terms -> wl_step_stage_b)."""
        self._py_rewards.append(SimpleNamespace(name=name, func=func, weight=float(weight), params=dict(params or {})))
        self._py_sums[name] = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)
        self.reward_manager.python_terms = [t.name for t in self._py_rewards]

    def add_termination_term(self, name: str, func, time_out: bool = False, params: dict | None = None):
        """Register ``func(env, **params) -> [N] bool`` as a termination term (TerminationTermCfg semantics; evaluated
        before the rewards, OR-ed into terminated / time_outs, resets the env in the same step)."""
        self._py_terms.append(SimpleNamespace(name=name, func=func, time_out=bool(time_out), params=dict(params or {})))
        self.termination_manager.python_terms = [t.name for t in self._py_terms]

    def add_observation_term(self, name: str, func, params: dict | None = None):
        """Register ``func(env, **params) -> [N, k]``; appended to the policy observation (after the built-in terms)."""
        term = SimpleNamespace(name=name, func=func, params=dict(params or {}))
        k = int(func(self, **term.params).reshape(self.num_envs, -1).shape[1])
        self._py_obs.append(term)
        d = self.observation_manager.group_obs_dim["policy"][0] + k
        self.observation_manager.group_obs_dim["policy"] = (d,)
        self.single_observation_space = {"policy": _Box(-math.inf, math.inf, (d,))}
        self.observation_space = {"policy": _Box(-math.inf, math.inf, (self.num_envs, d))}

    def _append_py_obs(self, obs: torch.Tensor) -> torch.Tensor:
        if not self._py_obs:
            return obs
        return torch.cat([obs] + [t.func(self, **t.params).reshape(self.num_envs, -1).to(torch.float32) for t in self._py_obs], dim=1)

    def _step_staged(self, action: torch.Tensor):
        """env.step with Python terms in the loop: stage a (A-E) -> terminations -> rewards -> stage b (F-I)."""
        t = self.common_step_counter
        sim, tm = self.sim, self.termination_manager
        rew, bits = sim.step_stage_a(action, t)
        built_in_to = (bits & 1).bool()
        built_in_term = (bits & 0xFE).bool()
        self._term_bits_staged = bits            # per-term masks of the built-in terms (get_term)
        extra_term = extra_to = None
        fired = {}
        self._py_term_masks = fired
        for term in self._py_terms:
            v = term.func(self, **term.params).to(torch.bool)
            fired[term.name] = v
            if term.time_out:
                extra_to = v if extra_to is None else (extra_to | v)
            else:
                extra_term = v if extra_term is None else (extra_term | v)
        # what reward terms such as is_terminated_term see (TerminationManager.compute runs before RewardManager.compute)
        tm.terminated = built_in_term if extra_term is None else (built_in_term | extra_term)
        tm.time_outs = built_in_to if extra_to is None else (built_in_to | extra_to)
        for term in self._py_rewards:
            if term.weight == 0.0:
                continue
            val = term.func(self, **term.params).to(torch.float32) * (term.weight * self.step_dt)
            rew += val
            self._py_sums[term.name] += val
        log = torch.empty(16, dtype=torch.float32, device=self.device) if self.log_episode_info else None
        obs, term_u8, trunc_u8 = sim.step_stage_b(bits, t, extra_term, extra_to, log=log)
        self.common_step_counter = t + 1
        terminated, truncated = term_u8.view(torch.bool), trunc_u8.view(torch.bool)
        tm.terminated, tm.time_outs = terminated, truncated
        if log is not None:
            done = terminated | truncated
            cnt = done.sum().clamp(min=1).to(torch.float32)
            extra = {}
            for term in self._py_rewards:                        # RewardManager.reset: mean episodic sum / max_episode_length_s
                sums = self._py_sums[term.name]
                extra["Episode_Reward/" + term.name] = (sums * done).sum() / cnt / self.max_episode_length_s
                sums.masked_fill_(done, 0.0)
            for term in self._py_terms:
                extra["Episode_Termination/" + term.name] = (fired[term.name] & done).sum()
            self.extras["log"] = _LazyLog(log, self._log_index, extra, env=self)
        return {"policy": self._append_py_obs(obs)}, rew, terminated, truncated, self.extras

    def step(self, action: torch.Tensor):
        if self._needs_reset:
            self.reset()
        if action.dtype != torch.float32 or not action.is_contiguous() or str(action.device) != self.device:
            action = action.to(self.device, torch.float32).contiguous()
        if self._py_rewards or self._py_terms:
            return self._step_staged(action)
        t = self.common_step_counter
        # outputs live in a ring of preallocated buffers (no allocator call per step; a returned tensor stays valid for
        # _ring_len steps -- rsl_rl copies into its storage at once, play_policy.py keeps at most one episode of references)
        if self._step_ring is None:
            n, dev = self.num_envs, self.device
            self._step_ring = [((torch.empty((n, self.spec.obs_dim), dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.float32, device=dev),
                                 torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)),
                                torch.empty(16, dtype=torch.float32, device=dev)) for _ in range(self._ring_len)]
        out, log = self._step_ring[t % self._ring_len]
        if not self.log_episode_info:
            log = None
        obs, rew, term_u8, trunc_u8 = self.sim.step(action, t, out=out, log=log)
        self.common_step_counter = t + 1
        # curriculum: applied by the step kernel's last CTA (uses the incremented counter and fires only if >= 1 env
        # reset, like the reference's call from _reset_idx) -- no host logic, no sync
        terminated, truncated = term_u8.view(torch.bool), trunc_u8.view(torch.bool)
        tm = self.termination_manager
        tm.terminated, tm.time_outs = terminated, truncated
        if log is not None:
            self.extras["log"] = self._episode_log(log)
        return {"policy": self._append_py_obs(obs)}, rew, terminated, truncated, self.extras

    def _episode_log(self, log: torch.Tensor):
        """extras["log"] (RewardManager/TerminationManager.reset): lazy views of the row the step
        kernel's last CTA wrote -- device tensors, no extra launches, no host sync."""
        return _LazyLog(log, self._log_index, env=self)

    def step_host(self, action_host: torch.Tensor):
        """env.step() for a HOST-side caller: `action_host` is a CPU tensor [N,2] (pinned for best speed).  One C call
        does H2D(actions) -> fused step -> D2H(reward, terminated, truncated) -> sync.  Returns
        (obs_dict [device], rew [pinned host], terminated [pinned host], truncated [pinned host], extras); the host
        result views are overwritten by the next step_host call.  With ``env.host_obs = True`` the observations are returned in
        pinned HOST memory as well (a ring of _RING buffers): D2H copy after the step (transport "copy") or written by the
        kernel over PCIe (transport "zero_copy")."""
        if self._py_rewards or self._py_terms:
            raise NotImplementedError("step_host: host-side Python terms need the staged step (use env.step)")
        if self._needs_reset:
            self.reset()
        if self._host_io is None:
            self._host_io = self.sim.make_host_io()
        io = self._host_io
        # a pinned, contiguous f32 block of the caller is read in place (DMA / PCIe reads straight from it); anything else is
        # staged through the env's own pinned buffer
        ptr = action_host.data_ptr()
        if ptr == io["h_action"].data_ptr():
            p_action = None
        elif (action_host.dtype == torch.float32 and action_host.is_contiguous() and action_host.shape == (self.num_envs, 2)
              and action_host.is_pinned()):                  # (queried every call: an address may be reused by pageable memory)
            p_action = self._pinned_ptrs.get(ptr)
            if p_action is None:
                import ctypes as C
                if len(self._pinned_ptrs) > 65536:
                    self._pinned_ptrs.clear()
                p_action = self._pinned_ptrs[ptr] = C.c_void_p(ptr)
        else:
            p_action = None
            io["h_action"].copy_(action_host)
        t = self.common_step_counter
        # outputs come from a ring of preallocated device buffers (valid for _RING steps, like IsaacLab's own reuse)
        k = t % self._ring_len
        if self._ring is None:
            import ctypes as C
            self._ring = []
            for _ in range(self._ring_len):
                o = torch.empty((self.num_envs, self.spec.obs_dim), dtype=torch.float32, device=self.device)
                lg = torch.empty(16, dtype=torch.float32, device=self.device)
                self._ring.append((o, lg, C.c_void_p(o.data_ptr()), C.c_void_p(lg.data_ptr()), _LazyLog(lg, self._log_index)))
        obs, log, p_obs, p_log, lazy = self._ring[k]
        if not self.log_episode_info:
            log, p_log = None, None
        h_obs = None
        if self.host_obs:                                # the caller lives on the host: observations come back too
            if self._host_obs_ring is None:
                self._host_obs_ring = [torch.empty((self.num_envs, self.spec.obs_dim), dtype=torch.float32).pin_memory()
                                       for _ in range(self._HOST_OBS_RING)]
            h_obs = self._host_obs_ring[t % self._HOST_OBS_RING]
        if self.host_transport == "zero_copy":
            if h_obs is not None:                        # the kernel writes the observation rows straight into pinned host memory
                import ctypes as C
                self.sim.step_host_zero_copy(io, t, h_obs, log, C.c_void_p(h_obs.data_ptr()), p_log, p_action)
            else:
                self.sim.step_host_zero_copy(io, t, obs, log, p_obs, p_log, p_action)
        else:
            self.sim.step_host(io, t, obs, log, h_obs=h_obs, p_action=p_action)
        if h_obs is not None:
            obs = h_obs
        self.common_step_counter = t + 1
        tm = self.termination_manager
        tm.terminated, tm.time_outs = io["terminated"], io["truncated"]
        if log is not None:
            lazy._env, lazy._step = self, t + 1          # one lazy view per ring slot (its row tensor never changes)
            self.extras["log"] = lazy
        return {"policy": obs}, io["rew"], io["terminated"], io["truncated"], self.extras

    @property
    def host_action_buffer(self) -> torch.Tensor:
        """Pinned [N,2] buffer a host-side policy can write actions into (zero-copy input of step_host)."""
        if self._host_io is None:
            self._host_io = self.sim.make_host_io()
        return self._host_io["h_action"]


def make(task_id: str, cfg=None, render_mode=None, device="cuda:0", **kw) -> ManagerBasedRLEnv:
    """gym.make(<id>, cfg=...) equivalent for the registered WheeledLab ids."""
    return ManagerBasedRLEnv(cfg if cfg is not None else task_id, render_mode=render_mode, device=device, **kw)
