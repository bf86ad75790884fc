#!/usr/bin/env python
"""A/B of the end-to-end host path (env.step_host) over kernel variants and transports; one JSON line."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import wheeledlab_b200 as wl

E, K = 4096, 300
res = {}
for variant in (8, 4):
    env = wl.ManagerBasedRLEnv(wl.drift_task(num_envs=E, seed=42), device="cuda:0")
    env.sim.set_kernel_variant(variant)
    env.reset()
    acts = torch.stack([env.sim.synth_actions(t) for t in range(64)]).cpu().pin_memory()
    rows = [acts[k] for k in range(64)]
    for transport, host_obs in (("zero_copy", True), ("copy", True), ("zero_copy", False), ("copy", False)):
        env.host_transport, env.host_obs = transport, host_obs
        for k in range(20):
            env.step_host(rows[k % 64])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(K):
            env.step_host(rows[k % 64])
        torch.cuda.synchronize()
        res[f"v{variant}_{transport}_{'hostobs' if host_obs else 'devobs'}_us"] = round((time.perf_counter() - t0) * 1e6 / K, 2)
print(json.dumps(res))
