"""Lowering of the REFERENCE's own env-cfg objects onto the fused kernels' POD config.

``gym.make("Isaac-MushrDriftRL-v0", cfg=env_cfg)`` hands the unmodified ``MushrDriftRLEnvCfg`` /
``MushrElevationRLEnvCfg`` / ``F1TenthDriftRLEnvCfg`` instance (built against IsaacLab, or against the stand-in in
``shims/isaaclab``) to the entry point.  This module reads that object -- terms are recognised by the NAME of the
function they reference, values are taken from the cfg (weights, params, dt, decimation, scales, noise, DR ranges) --
and fills a ``TaskSpec``.  A term the kernels do not implement raises ``NotImplementedError`` (no silent fallback); the
observation group is validated term by term (function, order, clip; BlindObs noise stds are read from the cfg).

Reference files: wheeledlab_tasks/drifting/mushr_drift_env_cfg.py, f1tenth_drift_env_cfg.py,
elevation/mushr_elevation_env_cfg.py, common/{actions,observations}.py.
"""
from __future__ import annotations

import math

from . import tasks as T

_DRIFT_REWARDS = {"side_slip": 0, "vel_dist": 1, "track_progress_rate": 2, "turn_left_go_right": 3, "energy_through_turn": 4,
                  "cross_track_dist": 5, "is_terminated_term": 6}
_ALIASES = {"turn_left_go_right_f1": "turn_left_go_right"}    # f1tenth_drift_env_cfg.py:90-104: same math, other joint names
_ELEV_REWARDS = {"goal_progress_rate": 0, "higher_elevation": 1, "is_falling_penalty": 2, "is_terminated_term": 3}
_VIS_REWARDS = {"traversable_reward": 0, "forward_vel": 1}
_ACTION_KINDS = {"RCCarRWDAction": T.ACT_RWD, "RCCar4WDAction": T.ACT_4WD, "AckermannAction": T.ACT_ACKERMANN}
# observation groups the kernels implement: term functions in order (common/observations.py:19-56,
# elevation/mushr_elevation_env_cfg.py:57-88, visual/mushr_visual_env_cfg.py:37-57)
_OBS_BLIND = ["root_pos_w", "root_euler_xyz", "base_lin_vel", "base_ang_vel", "last_action"]
_OBS_ELEV = ["goal_relative_xyz", "root_euler_xyz", "base_lin_vel", "base_ang_vel", "last_action", "world_height_map"]
_OBS_VIS_TAIL = ["base_lin_vel", "base_ang_vel", "last_action"]
_CAMERA_FUNCS = {"camera_data_rgb_flattened_aug": "aug", "camera_data_rgb_flattened": "raw"}
_BOUND = {None: T.BOUND_NONE, "clip": T.BOUND_CLIP, "tanh": T.BOUND_TANH}


def _fields(obj):
    if obj is None:
        return []
    names = getattr(obj, "_cfg_fields", None) or [k for k in vars(obj) if not k.startswith("_")]
    return [(k, getattr(obj, k)) for k in names if getattr(obj, k) is not None]


def _fname(term):
    f = term.func
    return getattr(f, "__name__", type(f).__name__)


def _obs_funcs(cfg):
    return [_fname(t) for _, t in _fields(cfg.observations.policy) if hasattr(t, "func")]


def _task_kind(cfg) -> str:
    """Which kernel family a reference env cfg maps to, decided by its observation group (present in train AND play cfgs)."""
    funcs = _obs_funcs(cfg)
    if "world_height_map" in funcs:
        return "elevation"
    if any(f in _CAMERA_FUNCS for f in funcs):
        return "visual"
    if "root_pos_w" in funcs:
        return "drift"
    raise NotImplementedError(f"observation group {funcs} is not one the fused step implements")


def _validate_obs(cfg, expected, c, blind=False):
    """The observation group must be exactly what the kernel assembles: same term functions in the same order; noise stds
    (BlindObs) and clips are read from the cfg."""
    terms = [(k, t) for k, t in _fields(cfg.observations.policy) if hasattr(t, "func")]
    funcs = [_fname(t) for _, t in terms]
    if funcs != expected:
        raise NotImplementedError(f"observation terms {funcs} differ from the kernel's {expected} (order matters)")
    if getattr(cfg.observations.policy, "concatenate_terms", True) is not True:
        raise NotImplementedError("observation group must concatenate its terms")
    for k, (name, t) in enumerate(terms):
        fn = funcs[k]
        clip = getattr(t, "clip", None)
        if fn == "last_action":
            if clip is not None and tuple(float(x) for x in clip) != (-1.0, 1.0):
                raise NotImplementedError(f"{name}: clip {clip} (the kernel clips the last action to +-1)")
        elif clip is not None:
            lo, hi = (float(x) for x in clip)
            if blind or lo != -hi:
                raise NotImplementedError(f"{name}: clip {clip} is not implemented for this observation group")
            c.obs_clip = hi
        if getattr(t, "scale", None) not in (None, 1, 1.0):
            raise NotImplementedError(f"{name}: observation scale")
        if blind and k < 4:
            n = getattr(t, "noise", None)
            std = 0.0
            if n is not None:
                if type(n).__name__ not in ("AdditiveGaussianNoiseCfg", "GaussianNoiseCfg") or float(getattr(n, "mean", 0.0)) != 0.0:
                    raise NotImplementedError(f"{name}: only zero-mean additive Gaussian noise is implemented")
                std = float(n.std)
            c.noise_std[k] = std


def _lower_common(cfg, spec: T.TaskSpec):
    c = spec.cfg
    sim_dt, dec = float(cfg.sim.dt), int(cfg.decimation)
    c.sim_dt, c.decimation = sim_dt, dec
    c.substeps = max(1, math.ceil(sim_dt / 0.005 - 1e-9))
    spec.episode_length_s = float(cfg.episode_length_s)
    c.episode_length_s = spec.episode_length_s
    c.max_episode_length = math.ceil(spec.episode_length_s / (sim_dt * dec))
    # action term (exactly one in every registered task: actions.throttle_steer)
    terms = _fields(cfg.actions)
    if len(terms) != 1:
        raise NotImplementedError(f"expected one action term, got {[k for k, _ in terms]}")
    a = terms[0][1]
    kind = a.class_type.__name__
    if kind not in _ACTION_KINDS:
        raise NotImplementedError(f"action term {kind} is not implemented by the fused step")
    c.action_kind = _ACTION_KINDS[kind]
    if a.bounding_strategy not in _BOUND:
        raise NotImplementedError(f"bounding_strategy {a.bounding_strategy!r}")
    c.bounding = _BOUND[a.bounding_strategy]
    c.no_reverse = 1 if a.no_reverse else 0
    c.act_scale[0], c.act_scale[1] = a.scale
    c.act_offset[0], c.act_offset[1] = a.offset
    c.base_length, c.base_width, c.wheel_radius_cfg = a.base_length, a.base_width, a.wheel_radius
    pol = cfg.observations.policy
    c.enable_corruption = 1 if getattr(pol, "enable_corruption", False) else 0
    if getattr(cfg, "terminations", None) is None:         # play cfgs ("no terminations"): nothing ever ends an episode
        c.term_enable = 0
        spec.termination_names = []
    # curriculum
    spec.curriculum = []
    for name, term in _fields(getattr(cfg, "curriculum", None)):
        if _fname(term) != "increase_reward_weight_over_time":
            raise NotImplementedError(f"curriculum term {name}: {_fname(term)}")
        p = term.params
        spec.curriculum.append(T.CurriculumTerm(name, p["reward_term_name"], float(p["increase"]),
                                                int(p.get("episodes_per_increase", 1)), p.get("max_increases", math.inf)))
    spec._curriculum_dirty = True


def _lower_rewards(cfg, spec, table, handlers, allow_python_terms=False):
    c = spec.cfg
    names = [None] * len(table)
    for k in range(T_MAX := 8):
        c.rew_weight[k] = 0.0
    for name, term in _fields(getattr(cfg, "rewards", None)):
        fn = _ALIASES.get(_fname(term), _fname(term))
        if fn not in table:
            if allow_python_terms:      # the reference's own Python function, evaluated between the two halves of the staged step
                spec.python_reward_terms.append((name, term.func, float(term.weight), dict(term.params or {})))
                continue
            raise NotImplementedError(f"reward term {name} -> {fn} is not implemented by the fused step "
                                      f"(pass allow_python_terms=True to run it as a host-side term)")
        slot = table[fn]
        names[slot] = name
        c.rew_weight[slot] = float(term.weight)
        handlers.get(fn, lambda p: None)(term.params or {})
    if any(n is None for n in names):
        missing = [f for f, s in table.items() if names[s] is None]
        # absent terms keep weight 0 (skipped by the reward manager); give them placeholder names
        for f in missing:
            names[table[f]] = f
    spec.reward_names = names
    c.num_rew_terms = len(names)
    T.set_curriculum(c, names, spec.curriculum)     # reward slots are known now


def _lower_material_and_mass_events(events, c, seed):
    """Startup DR events shared by the tasks (randomize_rigid_body_material / _mass, randomize_actuator_gains)."""
    wheel_mass = None
    for name, term in events.items():
        fn, p = _fname(term), (term.params or {})
        if fn == "randomize_rigid_body_material":
            c.dr_num_buckets = int(p["num_buckets"])
            D, Cs = T.material_buckets(c.dr_num_buckets, p["static_friction_range"], p["dynamic_friction_range"],
                                       p.get("make_consistent", False), c.ground_mu_s, c.ground_mu_d, seed)
            for k in range(c.dr_num_buckets):
                c.dr_bucket_D[k], c.dr_bucket_C[k] = D[k], Cs[k]
        elif fn == "randomize_actuator_gains":
            c.dr_kd_range[0], c.dr_kd_range[1] = p["damping_distribution_params"]
        elif fn == "randomize_rigid_body_mass":
            bodies = getattr(p["asset_cfg"], "body_names", None)
            op = p.get("operation", "add")
            if bodies is not None and "wheel" in str(bodies):
                if op != "abs":
                    raise NotImplementedError(f"{name}: wheel mass randomisation with operation {op!r}")
                wheel_mass = p["mass_distribution_params"]
            else:
                c.dr_mass_add[0], c.dr_mass_add[1] = p["mass_distribution_params"]
                c.dr_mass_mode = {"add": 0, "abs": 1}[op]
    c.dr_wheel_mass_enable = 1 if wheel_mass is not None else 0
    if wheel_mass is not None:
        c.dr_wheel_mass[0], c.dr_wheel_mass[1] = wheel_mass


def spec_from_reference_cfg(cfg, env_id_offset: int = 0, allow_python_terms: bool = False) -> T.TaskSpec:
    """Build the TaskSpec for a reference env-cfg instance: the train cfgs of the four registered ids (Drift, F1Tenth drift,
    Elevation, Visual, incl. MushrVisualRLRandomEnvCfg) and their play cfgs (rewards / terminations / curriculum = None).
    With ``allow_python_terms`` reward / termination terms the kernels do not implement are kept as host-side terms that
    call the cfg's own function (staged step, env.add_reward_term); otherwise they raise."""
    num_envs = int(cfg.scene.num_envs)
    seed = int(cfg.seed) if getattr(cfg, "seed", None) is not None else 42
    kind = _task_kind(cfg)
    if kind == "visual":
        return _lower_visual(cfg, num_envs, seed, env_id_offset, allow_python_terms)
    if kind == "drift":
        events = dict(_fields(cfg.events))
        randomize = "change_wheel_friction" in events
        drive = "4wd" if cfg.actions.throttle_steer.class_type.__name__ != "RCCarRWDAction" else "2wd"
        usd = str(getattr(getattr(cfg.scene.robot, "spawn", None), "usd_path", "")).lower()
        spec = T.drift_task(num_envs=num_envs, seed=seed, env_id_offset=env_id_offset, randomize=randomize, drive=drive,
                            vehicle="f1tenth" if "f1tenth" in usd else "mushr")
        c = spec.cfg
        _lower_common(cfg, spec)
        _validate_obs(cfg, _OBS_BLIND, c, blind=True)

        def h_slip(p):
            c.slip_min_thresh, c.slip_max_thresh = p["min_thresh"], p["max_thresh"]
            c.slip_min_vel_x = p.get("min_vel_x", 0.5)

        def h_vel(p):
            c.vel_speed_target = p.get("speed_target", T.MAX_SPEED)
            c.vel_offset = p.get("offset", -T.MAX_SPEED ** 2)

        def h_ctd(p):
            c.trk_straight = p["straight"]
            c.ctd_track_radius = p.get("track_radius", (T.CORNER_IN_RADIUS + T.CORNER_OUT_RADIUS) / 2)
            c.ctd_offset = p.get("offset", -1.0)
            if p.get("p", 1.0) != 1:
                raise NotImplementedError("cross_track_dist with p != 1")

        _lower_rewards(cfg, spec, _DRIFT_REWARDS, {
            "side_slip": h_slip, "vel_dist": h_vel, "cross_track_dist": h_ctd,
            "turn_left_go_right": lambda p: setattr(c, "tlgr_ang_vel_thresh", p.get("ang_vel_thresh", math.pi / 4)),
            "energy_through_turn": lambda p: setattr(c, "energy_straight", p["straight"]),
        }, allow_python_terms)
        for name, term in _fields(getattr(cfg, "terminations", None)):
            fn = _fname(term)
            if fn == "cart_off_track":
                p = term.params
                c.trk_straight, c.trk_corner_in, c.trk_corner_out = p["straight"], p["corner_in_radius"], p["corner_out_radius"]
            elif fn != "time_out":
                if not allow_python_terms:
                    raise NotImplementedError(f"termination term {name} -> {fn} (pass allow_python_terms=True to run it host-side)")
                spec.python_termination_terms.append((name, term.func, bool(getattr(term, "time_out", False)), dict(term.params or {})))
        for name, term in events.items():
            fn, p = _fname(term), term.params
            if fn == "reset_root_state_along_track":
                c.num_ref_poses = int(p.get("num_points", 20))
                c.reset_pos_noise, c.reset_yaw_noise = p.get("pos_noise", 0.0), p.get("yaw_noise", 0.0)
                poses = T.generate_reference_poses(c.num_ref_poses, p.get("track_radius", 0.8), p.get("track_straight_dist", 0.8), seed)
                for k, v in enumerate(poses.reshape(-1)):
                    c.ref_poses[k] = v
            elif fn in ("randomize_rigid_body_material", "randomize_actuator_gains", "randomize_rigid_body_mass"):
                pass                                            # lowered together below
            elif fn == "push_by_setting_velocity":
                vr = p["velocity_range"]
                if "x" in vr:                                   # high-frequency push
                    c.push_hf_interval[0], c.push_hf_interval[1] = term.interval_range_s
                    c.push_hf_range[0], c.push_hf_range[1], c.push_hf_range[2] = vr["x"][1], vr.get("y", (0, 0))[1], vr["yaw"][1]
                else:
                    c.push_lf_interval[0], c.push_lf_interval[1] = term.interval_range_s
                    c.push_lf_yaw = vr["yaw"][1]
            elif fn == "disable_all_lidars":
                pass                                            # no-op without omni.isaac.sensor (quirk Q15)
            else:
                raise NotImplementedError(f"event term {name} -> {fn}")
        _lower_material_and_mass_events(events, c, seed)
        return spec
    # elevation (elevation/mushr_elevation_env_cfg.py)
    spec = T.elevation_task(num_envs=num_envs, seed=seed, env_id_offset=env_id_offset)
    c = spec.cfg
    _lower_common(cfg, spec)
    _validate_obs(cfg, _OBS_ELEV, c)
    for _, t in _fields(cfg.observations.policy):
        if hasattr(t, "func") and _fname(t) == "world_height_map":
            c.scan_offset, c.scan_plane_init = float(t.params["offset"]), float(t.params["plane_init_value"])
    _lower_rewards(cfg, spec, _ELEV_REWARDS, {}, allow_python_terms)
    term_bits = {"time_out": 0, "root_height_below_minimum": 1, "stuck": 2, "upright_bool": 3, "close_to_goal": 4}
    if getattr(cfg, "terminations", None) is not None:
        mask = 0
        for name, term in _fields(cfg.terminations):
            fn, p = _fname(term), (term.params or {})
            if fn not in term_bits:
                if not allow_python_terms:
                    raise NotImplementedError(f"termination term {name} -> {fn} (pass allow_python_terms=True to run it host-side)")
                spec.python_termination_terms.append((name, term.func, bool(getattr(term, "time_out", False)), dict(p)))
                continue
            mask |= 1 << term_bits[fn]
            if fn == "root_height_below_minimum":
                c.elev_min_height = float(p["minimum_height"])
            elif fn == "stuck":
                c.elev_stuck_min_vel, c.elev_stuck_spin = float(p.get("min_vel", c.elev_stuck_min_vel)), float(p.get("spin_thresh", c.elev_stuck_spin))
            elif fn == "close_to_goal":
                c.elev_goal_dist = float(p.get("threshold", p.get("dist", c.elev_goal_dist)))
        c.term_enable = mask
    events = dict(_fields(cfg.events))
    for name, term in events.items():
        fn, p = _fname(term), (term.params or {})
        if fn == "reset_root_state_uniform":
            pr, vr = p["pose_range"], p.get("velocity_range", {})
            c.elev_reset_xy[0], c.elev_reset_xy[1] = pr["x"]
            c.elev_reset_yaw = float(pr["yaw"][1])
            if "x" in vr:
                c.elev_reset_vel[0], c.elev_reset_vel[1] = vr["x"]
        elif fn not in ("randomize_rigid_body_material", "randomize_actuator_gains", "randomize_rigid_body_mass"):
            raise NotImplementedError(f"event term {name} -> {fn}")
    _lower_material_and_mass_events(events, c, seed)
    cmds = dict(_fields(getattr(cfg, "commands", None)))
    for name, term in cmds.items():
        rng = getattr(term, "ranges", None)
        if rng is not None and hasattr(rng, "pos_x"):
            c.cmd_pos_range[0], c.cmd_pos_range[1] = rng.pos_x
        rt = getattr(term, "resampling_time_range", None)
        if rt is not None:
            c.cmd_resample_s = float(rt[0])
    return spec


def _lower_visual(cfg, num_envs, seed, env_id_offset, allow_python_terms):
    """MushrVisualRLEnvCfg / MushrVisualRLRandomEnvCfg / MushrVisualPlayEnvCfg (visual/mushr_visual_env_cfg.py:412-473)."""
    import numpy as np
    terrain = cfg.scene.terrain
    trav = np.asarray(terrain.traversability_hashmap, dtype=bool)
    funcs = _obs_funcs(cfg)
    cam_mode = _CAMERA_FUNCS[funcs[0]]
    events = dict(_fields(cfg.events))
    randomize = "change_wheel_friction" in events
    spec = T.visual_task(num_envs=num_envs, seed=seed, env_id_offset=env_id_offset, traversability=trav, camera=cam_mode, randomize=randomize)
    c = spec.cfg
    _lower_common(cfg, spec)
    _validate_obs(cfg, [funcs[0]] + _OBS_VIS_TAIL, c)
    # map geometry (VisualTerrainImporterCfg :63-124) and ground material (:126-135)
    c.vis_row_spacing, c.vis_col_spacing = float(terrain.row_spacing), float(terrain.col_spacing)
    c.vis_width, c.vis_height = float(terrain.width), float(terrain.height)
    c.vis_mesh_x0 = -c.vis_width / 2 - c.vis_row_spacing / 2
    c.vis_mesh_y0 = -c.vis_height / 2 - c.vis_col_spacing / 2
    c.vis_mesh_dx, c.vis_mesh_dy = c.vis_width / (c.vis_rows - 1), c.vis_height / (c.vis_cols - 1)
    mat = getattr(terrain, "physics_material", None)
    if mat is not None:
        if getattr(mat, "friction_combine_mode", "multiply") != "multiply":
            raise NotImplementedError("ground friction combine mode other than 'multiply'")
        c.ground_mu_s, c.ground_mu_d = float(mat.static_friction), float(mat.dynamic_friction)
    # camera (TiledCameraCfg :230-246)
    cam = cfg.scene.camera
    c.vis_cam_w, c.vis_cam_h = int(cam.width), int(cam.height)
    c.vis_cam_row0 = c.vis_cam_h // 3
    sp = cam.spawn
    c.vis_cam_fx = float(sp.focal_length) * c.vis_cam_w / float(sp.horizontal_aperture)
    c.vis_cam_fy = float(sp.focal_length) * c.vis_cam_h / float(sp.vertical_aperture)
    c.vis_cam_cx, c.vis_cam_cy = c.vis_cam_w / 2.0, c.vis_cam_h / 2.0
    off = cam.offset.pos
    c.vis_cam_pos[0], c.vis_cam_pos[1], c.vis_cam_pos[2] = 0.022995 + float(off[0]), 0.0175 + float(off[1]), 0.162178 + float(off[2])
    spec.obs_dim = 8 + c.vis_cam_w * (c.vis_cam_h - c.vis_cam_row0)
    _lower_rewards(cfg, spec, _VIS_REWARDS, {}, allow_python_terms)
    if getattr(cfg, "terminations", None) is not None:
        mask = 0
        for name, term in _fields(cfg.terminations):
            fn = _fname(term)
            if fn == "time_out":
                mask |= 1
            elif fn == "out_of_map":
                mask |= 2
            elif allow_python_terms:
                spec.python_termination_terms.append((name, term.func, bool(getattr(term, "time_out", False)), dict(term.params or {})))
            else:
                raise NotImplementedError(f"termination term {name} -> {fn} (pass allow_python_terms=True to run it host-side)")
        c.term_enable = mask
    for name, term in events.items():
        fn = _fname(term)
        if fn not in ("reset_root_state", "randomize_rigid_body_material", "randomize_actuator_gains", "randomize_rigid_body_mass"):
            raise NotImplementedError(f"event term {name} -> {fn}")
    if randomize:
        _lower_material_and_mass_events(events, c, seed)
    else:                                                      # material table follows the (possibly overridden) ground friction
        D, Cs = T.material_buckets(1, (1.0, 1.0), (1.0, 1.0), True, c.ground_mu_s, c.ground_mu_d, seed)
        c.dr_bucket_D[0], c.dr_bucket_C[0] = D[0], Cs[0]
    return spec


def env_from_reference_cfg(cfg, render_mode=None, device=None, allow_python_terms: bool = False, **kwargs):
    """Entry point used by ``shims/isaaclab/envs:ManagerBasedRLEnv`` (gym kwargs: cfg=<env cfg>)."""
    from .env import ManagerBasedRLEnv
    kwargs.pop("env_cfg_entry_point", None); kwargs.pop("rsl_rl_cfg_entry_point", None); kwargs.pop("play_env_cfg_entry_point", None)
    if isinstance(cfg, (str, T.TaskSpec)):           # a gym id / an already lowered TaskSpec instead of a reference cfg object
        return ManagerBasedRLEnv(cfg, render_mode=render_mode, device=device or "cuda:0")
    dev = device or getattr(getattr(cfg, "sim", None), "device", None) or "cuda:0"
    env = ManagerBasedRLEnv(spec_from_reference_cfg(cfg, allow_python_terms=allow_python_terms), render_mode=render_mode, device=dev)
    env.cfg = cfg                                   # writer.log_config(self.env.cfg, ...) (modified_rsl_rl_runner.py:42-44)
    if not hasattr(cfg, "is_finite_horizon"):
        cfg.is_finite_horizon = False
    return env
