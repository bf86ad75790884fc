"""Long-run robustness check: many thousands of env.steps per task at full size under random actions; every output must
stay finite and bounded, resets must keep happening, episode statistics must be sane.

    python tools/soak.py [--envs 4096] [--steps 20000]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import wheeledlab_b200 as wl  # noqa: E402


def soak(name, spec, steps, dist):
    sim = wl.WheeledSim(spec, "cuda:0")
    sim.startup(); sim.reset(None, 0)
    n = sim.num_envs
    out = None
    stats = {"task": name, "envs": n, "steps": steps, "action_dist": dist}
    n_term = torch.zeros((), device="cuda"); n_trunc = torch.zeros((), device="cuda")
    rew_sum = torch.zeros((), device="cuda", dtype=torch.float64)
    bad = torch.zeros((), device="cuda")
    max_abs_obs = torch.zeros((), device="cuda"); max_speed = torch.zeros((), device="cuda"); max_z = torch.zeros((), device="cuda")
    for t in range(steps):
        act = sim.synth_actions(t, dist)
        out = sim.step(act, t, out=out)
        obs, rew, term, trunc = out
        if t % 16 == 0 or t == steps - 1:                       # checks every 16 steps (they cost more than the step)
            bad += (~torch.isfinite(obs)).sum() + (~torch.isfinite(rew)).sum() + (~torch.isfinite(sim.groups[:9])).sum()
            max_abs_obs = torch.maximum(max_abs_obs, obs[:, :13].abs().max())
            max_speed = torch.maximum(max_speed, sim.root_lin_vel_w.norm(dim=1).max())
            max_z = torch.maximum(max_z, sim.root_pos_w[:, 2].abs().max())
        n_term += term.sum(); n_trunc += trunc.sum(); rew_sum += rew.sum(dtype=torch.float64)
    torch.cuda.synchronize()
    stats.update(non_finite=int(bad), terminated=int(n_term), truncated=int(n_trunc), mean_reward=float(rew_sum) / (n * steps),
                 max_abs_proprio_obs=float(max_abs_obs), max_speed=float(max_speed), max_abs_z=float(max_z),
                 episodes_per_env=float(n_term + n_trunc) / n)
    return stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20000)
    a = ap.parse_args()
    rows = []
    for name, mk, steps in (("drift", lambda: wl.drift_task(num_envs=a.envs, seed=11), a.steps),
                            ("drift_4wd_gauss", lambda: wl.drift_task(num_envs=a.envs, seed=12, drive="4wd"), a.steps // 2),
                            ("elevation", lambda: wl.elevation_task(num_envs=a.envs, seed=13), a.steps // 4),
                            ("visual", lambda: wl.visual_task(num_envs=a.envs, seed=14), a.steps // 4)):
        r = soak(name, mk(), steps, 1 if "gauss" in name else 0)
        print(json.dumps(r), flush=True)
        rows.append(r)
    ok = all(r["non_finite"] == 0 and r["max_speed"] < 50.0 and r["max_abs_z"] < 10.0 and r["episodes_per_env"] > 1 for r in rows)
    print("SOAK", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
