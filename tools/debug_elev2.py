import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wheeledlab_b200 as wl
x, y = float(sys.argv[1]), float(sys.argv[2])
spec = wl.elevation_task(num_envs=1, seed=1)
sim = wl.WheeledSim(spec, "cuda:0")
sim.startup(); sim.reset(None, 0)
sim.root_pos_w[0, 0] = x; sim.root_pos_w[0, 1] = y
sim.set_scan_tma(False); ref = sim.observe(0).clone(); torch.cuda.synchronize()
sim.set_scan_tma(True)
try:
    o = sim.observe(0); torch.cuda.synchronize()
    print("TMA ok at", x, y, "equal to plain:", bool(torch.equal(o, ref)), flush=True)
except Exception as e:
    print("TMA FAILED at", x, y, repr(e)[:120], flush=True)
