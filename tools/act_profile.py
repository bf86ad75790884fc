"""Time wl_act_step (policy fused in the step kernel) against wl_step on the same state: the difference is the MLP cost.
Usage: python tools/act_profile.py [--envs 4096] [--iters 400]; run under ncu with -k regex:wl_act_step for the kernel profile."""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import wheeledlab_b200 as wl                                        # noqa: E402
from wheeledlab_b200.policy import act_step, pack_actor_critic      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=400)
    a = ap.parse_args()
    dev, n = "cuda:0", a.envs
    torch.manual_seed(0)
    mk = lambda out: torch.nn.Sequential(torch.nn.Linear(14, 64), torch.nn.ELU(), torch.nn.Linear(64, 64), torch.nn.ELU(), torch.nn.Linear(64, out)).to(dev)
    blob = pack_actor_critic(mk(2), mk(1), torch.ones(2), 14, dev)
    res = {}
    for mode in ("step", "act_step"):
        sim = wl.WheeledSim(wl.drift_task(num_envs=n, seed=1), dev)
        sim.startup(); sim.reset(None, 0)
        obs = sim.observe(0, 0)
        out = (torch.empty_like(obs), torch.empty(n, device=dev), torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev))
        act = torch.zeros((n, 2), device=dev); mean = torch.empty((n, 2), device=dev); lp = torch.empty(n, device=dev); val = torch.empty(n, device=dev)
        log = torch.empty(16, device=dev)
        sim.set_step_counter(0)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.cuda.graph(g, stream=s):
            for k in range(50):
                if mode == "step":
                    sim.step(act, wl.WheeledSim.device_counter_plus(k), out=out, log=log)
                else:
                    act_step(sim, out[0], blob, act, mean, lp, val, out, log, wl.WheeledSim.device_counter_plus(k))
            sim.advance_counter(50)
        torch.cuda.current_stream().wait_stream(s)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(1, a.iters // 50)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        res[mode + "_us"] = e0.elapsed_time(e1) * 1e3 / (reps * 50)
    res["mlp_us"] = res["act_step_us"] - res["step_us"]
    res["envs"] = n
    print(json.dumps(res))


if __name__ == "__main__":
    main()
