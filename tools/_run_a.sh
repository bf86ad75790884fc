mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "torch_ops" 2>&1 | tail -15
timeout 100 python - <<'PY' 2>&1 | tail -3
import time, torch, wheeledlab_b200 as wl
from wheeledlab_b200 import torch_ops
ops = torch_ops.load()
sim = wl.WheeledSim(wl.drift_task(num_envs=4096, seed=1), "cuda:0"); sim.startup(); sim.reset(None, 0)
act = sim.synth_actions(0); outs = sim.step(act, 0); t = 1
f = sim.bind_step(act, outs)
for name, fn in (("ctypes sim.step", lambda t: sim.step(act, t, out=outs)), ("ctypes bound", lambda t: f(t)),
                 ("torch.ops step_out", lambda t: ops.step_out(sim.handle, act, outs[0], outs[1], outs[2], outs[3], None, t))):
    for _ in range(200): fn(t); t += 1
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3000): fn(t); t += 1
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print(name, "host us per call:", round((t1 - t0) / 3000 * 1e6, 2))
PY
