"""Drive every kernel of the library once or twice at tiny sizes: the workload for compute-sanitizer
(`compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/exercise_all.py`)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import wheeledlab_b200 as wl                                              # noqa: E402
from wheeledlab_b200.distributed import RolloutSlab                       # noqa: E402
from wheeledlab_b200.learner import compute_returns                       # noqa: E402
from wheeledlab_b200.policy import act_step, pack_actor_critic            # noqa: E402

dev = "cuda:0"


def drive(spec, steps=3, variants=(1, 4)):
    sim = wl.WheeledSim(spec, dev)
    n = sim.num_envs
    sim.startup(); sim.reset(None, 0)
    obs = sim.observe(0, 0)
    log = torch.empty(16, device=dev)
    t = 0
    for v in variants:
        sim.set_kernel_variant(v)
        for _ in range(steps):
            a = sim.synth_actions(t)
            obs, rew, term, trunc = sim.step(a, t, log=log)
            t += 1
    sim.reset(torch.arange(0, n, 3, device=dev, dtype=torch.int32), t)
    sim.suspension_state()
    sim.curriculum([0], [0.5], 1)
    torch.cuda.synchronize()
    return sim, t


def main():
    torch.manual_seed(0)
    # Drift: thread + quad kernels, fused rollout, host transports, fused policy
    sim, t = drive(wl.drift_task(num_envs=301, seed=3))
    n = sim.num_envs
    slab = RolloutSlab(8, n, sim.obs_dim, 2, dev); logs = torch.empty((8, 16), device=dev)
    sim.rollout(8, t, slab, logs); t += 8
    io = sim.make_host_io()
    o = torch.empty((n, sim.obs_dim), device=dev)
    sim.step_host(io, t, o); sim.step_host_zero_copy(io, t + 1, o); t += 2
    mk = lambda out: torch.nn.Sequential(torch.nn.Linear(14, 64), torch.nn.ELU(), torch.nn.Linear(64, 64), torch.nn.ELU(), torch.nn.Linear(64, out)).to(dev)
    blob = pack_actor_critic(mk(2), mk(1), torch.ones(2), 14, dev)
    act = torch.empty((n, 2), device=dev); mean = torch.empty((n, 2), device=dev); lp = torch.empty(n, device=dev); val = torch.empty(n, device=dev)
    out = (torch.empty_like(o), torch.empty(n, device=dev), torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev))
    for k in range(3):
        act_step(sim, o, blob, act, mean, lp, val, out, None, t + k)
        o = out[0].clone()
    compute_returns(torch.randn(16, n, device=dev), torch.randn(16, n, device=dev), torch.randn(n, device=dev),
                    (torch.rand(16, n, device=dev) < 0.1).to(torch.uint8), 0.99, 0.95, (torch.rand(16, n, device=dev) < 0.05).to(torch.uint8))
    # 4WD + F1Tenth parameter sets
    drive(wl.drift_task(num_envs=65, seed=4, drive="4wd", vehicle="f1tenth"), steps=2)
    # Elevation: both scan paths, both step variants; Visual
    es, _ = drive(wl.elevation_task(num_envs=70, seed=5), steps=2)
    es.set_scan_tma(False); es.step(es.synth_actions(99), 99)
    drive(wl.visual_task(num_envs=70, seed=6), steps=2)
    vs, _ = drive(wl.visual_task(num_envs=33, seed=6, camera="aug"), steps=2)          # camera kernel (drawn parameters)
    drive(wl.visual_task(num_envs=9, seed=6, camera="raw"), steps=1)
    buf = torch.zeros((33, vs.obs_dim), device=dev)
    vs.camera(3, buf, torch.tensor([1.1, 0.9, 1.2, 0.1, 1.5, 2, 0, 3, 1], dtype=torch.float32, device=dev))
    torch.cuda.synchronize()
    print("exercise_all ok")


if __name__ == "__main__":
    main()
