import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wheeledlab_b200 as wl
spec = wl.elevation_task(num_envs=64, seed=1)
for tma in (False, True):
    sim = wl.WheeledSim(spec, "cuda:0")
    sim.set_scan_tma(tma)
    sim.startup(); torch.cuda.synchronize(); print("startup ok", flush=True)
    sim.reset(None, 0); torch.cuda.synchronize(); print("reset ok", flush=True)
    try:
        o = sim.observe(0); torch.cuda.synchronize(); print("observe ok tma=", tma, float(o[:, 13:].mean()), flush=True)
        for v in (1, 4):
            sim.set_kernel_variant(v)
            out = sim.step(sim.synth_actions(0), 0); torch.cuda.synchronize(); print("step ok variant", v, "tma", tma, flush=True)
    except Exception as e:
        print("FAILED tma=", tma, repr(e)[:200], flush=True)
        break
