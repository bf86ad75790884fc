"""A/B of the height-scan kernels (Elevation, 4096 envs): 0 plain loads, 1 one TMA tile per CTA, 2 TMA producer/consumer pipeline.
20 observe() calls in one CUDA graph, replayed; prints microseconds per call.  Outputs of the three are compared bit for bit."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wheeledlab_b200 as wl

n = int(os.environ.get("ENVS", 4096))
spec = wl.make_task("Isaac-MushrElevationRL-v0", num_envs=n, terrain="procedural")
sim = wl.WheeledSim(spec, "cuda:0"); sim.startup(); sim.reset(None, 0)
for t in range(5):
    sim.step(sim.synth_actions(t), t)
obs = torch.empty((n, sim.obs_dim), device="cuda")
res, ref = {}, None
for mode in (0, 1, 2):
    sim.set_scan_tma(mode)
    sim.observe(5, out=obs); torch.cuda.synchronize()
    if ref is None:
        ref = obs.clone()
    assert torch.equal(ref.view(torch.int32), obs.view(torch.int32)), mode
    g = torch.cuda.CUDAGraph(); cs = torch.cuda.Stream()
    cs.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cs), torch.cuda.graph(g, stream=cs):
        for _ in range(20):
            sim.observe(5, out=obs)
    torch.cuda.current_stream().wait_stream(cs)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200_000); a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / 20)
    res[f"mode{mode}_us"] = round(min(ts), 3)
res["envs"] = n; res["ctas_per_sm"] = os.environ.get("WL_SCAN_CTAS_PER_SM", "6")
print(json.dumps(res))
