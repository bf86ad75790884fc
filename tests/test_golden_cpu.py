"""The oracle against GOLDEN VECTORS produced by executing the reference's own Python
(tests/golden/make_golden.py, run where /root/reference exists; fixtures committed as .npz).
fp32 tolerance: the reference evaluates with torch/libm, the oracle with its deterministic polynomials
(|err| <= ~3e-7 relative, test_oracle_cpu.py) -> rtol 2e-6 / atol 2e-6; masks must match exactly away
from float ties."""
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O

G = Path(__file__).resolve().parent / "golden"
RTOL, ATOL = 2e-6, 2e-6


def _task(**kw):
    import wheeledlab_b200 as wl
    return wl.drift_task(num_envs=4, **kw)


def test_action_terms_match_reference():
    g = np.load(G / "actions.npz")
    a = g["actions"]
    rwd = _task().cfg
    wheel, steer = O.action_map(rwd, a)
    # reference RWD writes [bl, br] only (rc_car_actions.py:27)
    assert np.allclose(wheel[:, :2], g["rwd_wheel"], rtol=RTOL, atol=ATOL)
    assert np.allclose(steer, g["rwd_steer"], rtol=RTOL, atol=ATOL)
    fwd = _task(drive="4wd").cfg
    wheel, steer = O.action_map(fwd, a)
    assert np.allclose(wheel, g["fwd_wheel"], rtol=2e-5, atol=ATOL)       # |R/(R r)| near tan->0 amplifies 1 ulp of tan
    assert np.allclose(steer, g["fwd_steer"], rtol=RTOL, atol=ATOL)
    ack = _task(drive="4wd").cfg
    ack.action_kind = 0
    wheel, steer = O.action_map(ack, a)
    assert np.allclose(wheel, g["ack_wheel"], rtol=2e-5, atol=ATOL)
    assert np.allclose(steer, g["ack_steer"], rtol=2e-5, atol=ATOL)
    f1 = _task(drive="4wd").cfg
    f1.base_length, f1.base_width = 0.365, 0.284                            # F1Tenth4WDActionCfg, common/actions.py:64-66
    wheel, steer = O.action_map(f1, a)
    assert np.allclose(wheel, g["f1_wheel"], rtol=2e-5, atol=ATOL)
    # processed actions: clip*scale, no_reverse
    assert np.allclose(g["rwd_processed"][:, 0], np.maximum(np.clip(a[:, 0], -1, 1) * 3.0, 0.0))
    assert list(g["fwd_wheel_ids"]) == [0, 1, 2, 3] and list(g["rwd_wheel_ids"]) == [0, 1]


def test_drift_reward_and_termination_terms_match_reference():
    g = np.load(G / "drift_terms.npz")
    cfg = _task().cfg
    n = g["pos"].shape[0]
    root = np.zeros((n, 13), np.float32)
    root[:, 0:3] = g["pos"]; root[:, 3] = 1.0; root[:, 7:10] = g["vel_b"]
    ep = np.where(g["time_outs"] > 0, 250, 3).astype(np.int32)
    root[:, 10:13] = g["ang_w"]                       # identity orientation: world == body
    f_w, oob = O.drift_terms(cfg, root, g["steer"], ep)
    root[:, 10:13] = g["ang_b"]
    f_b, _ = O.drift_terms(cfg, root, g["steer"], ep)
    assert np.array_equal(oob, g["out_of_bounds"])
    # side_slip thresholds: compare away from the three float thresholds
    slip = np.abs(np.arctan2(g["vel_b"][:, 1].astype(np.float64), g["vel_b"][:, 0].astype(np.float64)))
    safe = (np.abs(slip - 0.55) > 1e-5) & (np.abs(slip - 0.25) > 1e-5) & (np.abs(np.abs(g["vel_b"][:, 0]) - 1.0) > 1e-5)
    assert np.allclose(f_w[safe, 0], g["f_side_slip"][safe], rtol=RTOL, atol=ATOL) and safe.mean() > 0.99
    assert np.allclose(f_w[:, 1], g["f_vel"], rtol=RTOL, atol=1e-5)
    assert np.array_equal(f_w[:, 2], g["f_progress"])
    assert np.allclose(f_b[:, 3], g["f_tlgr"], rtol=RTOL, atol=ATOL)
    assert np.allclose(f_w[:, 4], g["f_turn_energy"], rtol=RTOL, atol=1e-5)
    assert np.allclose(f_w[:, 5], g["f_cross_track"], rtol=RTOL, atol=ATOL)
    assert np.array_equal(f_w[:, 6], g["f_term_pens"])
    import wheeledlab_b200 as wl
    w = np.array(list(cfg.rew_weight))[:7]
    assert np.allclose(w, [g["w_side_slip"], g["w_vel"], g["w_progress"], g["w_tlgr"], g["w_turn_energy"],
                           g["w_cross_track"], g["w_term_pens"]])


def test_reference_poses_and_reset_match_reference():
    g = np.load(G / "reset_along_track.npz")
    from wheeledlab_b200.tasks import reference_poses_from_dists
    dists = g["u"] * np.float32(2 * np.pi * 0.8 + 4 * 0.8)
    mine = reference_poses_from_dists(dists.astype(np.float32), 0.8, 0.8)
    ref = g["reference_poses"]                         # [20,2,3]
    assert np.allclose(mine[:, 0:2], ref[:, 0, 0:2], atol=2e-6)
    assert np.allclose(mine[:, 2], ref[:, 1, 2], atol=1e-4)
    assert (ref[:, 0, 2] == 0).all() and (ref[:, 1, 0:2] == 0).all()
    cfg = _task().cfg
    for k in range(20):
        cfg.ref_poses[3 * k + 0], cfg.ref_poses[3 * k + 1], cfg.ref_poses[3 * k + 2] = ref[k, 0, 0], ref[k, 0, 1], ref[k, 1, 2]
    pose = O.drift_reset_pose(cfg, g["idx"], g["u_xy"], g["u_yaw"])
    assert np.allclose(pose, g["pose"], rtol=RTOL, atol=ATOL)
    assert (g["velocity"] == 0).all()


def test_curriculum_matches_reference():
    """weights after running the reference's increase_reward_weight_over_time vs the host fire-mask logic."""
    g = np.load(G / "curriculum.npz")["rows"]
    import wheeledlab_b200 as wl
    env = wl.ManagerBasedRLEnv.__new__(wl.ManagerBasedRLEnv)
    env.spec = wl.drift_task(num_envs=4)
    env.max_episode_length = 250
    w = {"side_slip": 10.0, "tlgr": 0.0, "term_pens": -5000.0}
    k = 0
    for c in range(1, int(g[-1, 0]) + 1):
        env.common_step_counter = c
        m = env._curriculum_fire_mask()
        for j, t in enumerate(env.spec.curriculum):
            if (m >> j) & 1:
                w[t.reward_term_name] += t.increase
        if c == int(g[k, 0]):
            assert (w["side_slip"], w["tlgr"], w["term_pens"]) == tuple(g[k, 1:4]), c
            k += 1
    assert k == len(g)


def test_root_euler_xyz_matches_reference():
    g = np.load(G / "euler.npz")
    e = O.euler_xyz(g["quat"])
    d = np.abs(e - g["euler"])
    d = np.minimum(d, 2 * np.pi - d)                 # 0 == 2pi on the wrap
    # asin near +-1 is ill-conditioned: 1 ulp of sin_pitch moves the angle by ~sqrt(ulp)
    sp = 2 * (g["quat"][:, 0] * g["quat"][:, 2] - g["quat"][:, 3] * g["quat"][:, 1])
    ok = np.abs(sp) < 0.999
    assert d[ok].max() < 5e-6 and d.max() < 2e-3
