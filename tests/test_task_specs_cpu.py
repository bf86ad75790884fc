"""Host-side task definitions (wheeledlab_b200/tasks.py): every registered gym id lowers to a config whose fields are
populated (a field added to wl_config and forgotten in a task stays 0 and silently breaks the physics), with the values
the reference cfgs state."""
import math

import numpy as np
import pytest


def _specs():
    import wheeledlab_b200 as wl
    small = np.ones((8, 8), dtype=bool)
    return {
        "Isaac-MushrDriftRL-v0": wl.make_task("Isaac-MushrDriftRL-v0", num_envs=8),
        "Isaac-F1TenthDriftRL-v0": wl.make_task("Isaac-F1TenthDriftRL-v0", num_envs=8),
        "Isaac-MushrElevationRL-v0": wl.make_task("Isaac-MushrElevationRL-v0", num_envs=8, terrain="procedural"),
        "Isaac-MushrVisualRL-v0": wl.make_task("Isaac-MushrVisualRL-v0", num_envs=8, traversability=small),
    }


def test_every_gym_id_has_a_spec_with_reference_timing_and_dims():
    s = _specs()
    exp = {   # (sim_dt, decimation, episode_length_s, max_episode_length, obs_dim)
        "Isaac-MushrDriftRL-v0": (0.005, 4, 5.0, 250, 14), "Isaac-F1TenthDriftRL-v0": (0.005, 4, 5.0, 250, 14),
        "Isaac-MushrElevationRL-v0": (0.01, 10, 20.0, 200, 689), "Isaac-MushrVisualRL-v0": (0.02, 10, 10.0, 50, 3208),
    }
    for gid, (dt, dec, ep_s, L, od) in exp.items():
        c = s[gid].cfg
        assert abs(c.sim_dt - dt) < 1e-9 and c.decimation == dec and abs(s[gid].episode_length_s - ep_s) < 1e-9, gid
        assert c.max_episode_length == L == math.ceil(ep_s / (dt * dec)) and s[gid].obs_dim == od, gid
        assert c.substeps >= 1 and c.sim_dt / c.substeps <= 0.005 + 1e-9, gid          # integrator sub-step <= 5 ms


@pytest.mark.parametrize("gid", ["Isaac-MushrDriftRL-v0", "Isaac-F1TenthDriftRL-v0", "Isaac-MushrElevationRL-v0", "Isaac-MushrVisualRL-v0"])
def test_physical_parameters_are_populated(gid):
    c = _specs()[gid].cfg
    positive = ["mass_nominal", "wheel_radius", "wheel_inertia", "susp_k", "susp_c", "comp_max", "steer_kp", "steer_kd", "steer_inertia",
                "steer_pos_limit", "steer_vel_limit", "dc_saturation", "dc_vel_limit", "tire_mx", "tire_my", "gravity", "base_length",
                "base_width", "wheel_radius_cfg", "ground_mu_s", "ground_mu_d"]
    names = {f[0] for f in type(c)._fields_}
    for name in positive:
        if name in names:
            assert float(getattr(c, name)) > 0.0, (gid, name)
    assert all(float(x) > 0 for x in c.inertia_nominal) and c.num_rew_terms >= 2 and c.dr_num_buckets >= 1
    assert all(float(x) > 0 for x in list(c.dr_bucket_D)[: c.dr_num_buckets])
    # action map constants of the reference (common/actions.py:17-23,41-47,64-70)
    assert tuple(round(float(x), 3) for x in c.act_scale) == (3.0, 0.488) and c.no_reverse == 1
    if "F1Tenth" in gid:
        assert abs(c.base_length - 0.365) < 1e-6 and abs(c.base_width - 0.284) < 1e-6
    else:
        assert abs(c.base_length - 0.325) < 1e-6 and abs(c.base_width - 0.2) < 1e-6 and abs(c.wheel_radius_cfg - 0.05) < 1e-6


def test_reward_tables_match_the_reference_cfgs():
    s = _specs()
    d = s["Isaac-MushrDriftRL-v0"]
    assert d.reward_names == ["side_slip", "vel", "progress", "tlgr", "turn_energy", "cross_track", "term_pens"]
    assert [round(float(x), 3) for x in list(d.cfg.rew_weight)[:7]] == [10.0, -5.0, 40.0, 0.0, 20.0, -50.0, -5000.0]     # :243-299
    assert sorted((t.reward_term_name, t.increase) for t in d.curriculum) == sorted([("side_slip", 20.0), ("tlgr", 10.0), ("term_pens", -1000.0)])
    e = s["Isaac-MushrElevationRL-v0"]
    assert [round(float(x), 3) for x in list(e.cfg.rew_weight)[:4]] == [200.0, 5000.0, 0.0, -200.0]
    v = s["Isaac-MushrVisualRL-v0"]
    assert v.reward_names == ["traversablility", "vel_rew"] and [float(x) for x in list(v.cfg.rew_weight)[:2]] == [5.0, 7.0]
    assert v.cfg.vis_cam == 2 and (v.cfg.vis_cam_w, v.cfg.vis_cam_h, v.cfg.vis_cam_row0) == (80, 60, 20)


def test_config_finalize_fills_every_derived_field():
    import wheeledlab_b200 as wl
    from wheeledlab_b200._lib import check, lib
    import ctypes as C
    for spec in _specs().values():
        c = type(spec.cfg).from_buffer_copy(spec.cfg)
        check(lib.wl_config_finalize(C.byref(c)), "wl_config_finalize")
        for name, *_ in type(c)._fields_:
            if name.startswith("d_") and not name.startswith("d_vis_") and not (name == "d_inv_hf_cell" and c.hf_cell == 0):
                v = getattr(c, name)
                vals = list(v) if hasattr(v, "__len__") else [v]
                assert all(float(x) > 0.0 and math.isfinite(float(x)) for x in vals), name
        assert abs(c.d_step_dt - c.sim_dt * c.decimation) < 1e-7
