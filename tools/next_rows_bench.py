"""Timings of the SURVEY 8f rows (the components either side of the hot path): GAE kernel, staged step vs fused step,
env.step with a Python reward term.  CUDA events, warm L2 (these run back to back with the rollout in practice).

    python tools/next_rows_bench.py            # one JSON object on stdout
"""
import json
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import wheeledlab_b200 as wl                              # noqa: E402
from bench import _peaks                                 # noqa: E402
from wheeledlab_b200.learner import compute_returns      # noqa: E402


def timed(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev) * 1e3      # us


def timed_graph(fn, reps=20):
    """Device time per call with the host out of the picture: `reps` calls captured in ONE CUDA graph, replayed."""
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def main():
    dev = "cuda:0"
    peak, src = _peaks()
    out = {"peak_GBps": peak, "peak_source": src}
    # ---- f-2: GAE(lambda) over the [T, N] slab
    rows = []
    for T, N in ((128, 4096), (128, 65536), (128, 1048576)):
        r, v = torch.randn(T, N, device=dev), torch.randn(T, N, device=dev)
        lv = torch.randn(N, device=dev)
        d = (torch.rand(T, N, device=dev) < 0.02).to(torch.uint8); to = (torch.rand(T, N, device=dev) < 0.01).to(torch.uint8)
        import ctypes as C
        ret, adv = torch.empty_like(r), torch.empty_like(r)
        P = lambda x: C.c_void_p(x.data_ptr())

        def gae():
            wl.lib.wl_gae(P(r), P(v), P(lv), P(d), P(to), 0.99, 0.95, P(ret), P(adv), T, N, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        us = timed_graph(gae)
        us_eager = timed(lambda: compute_returns(r, v, lv, d, 0.99, 0.95, to), reps=20)
        nbytes = T * N * (4 + 4 + 1 + 1 + 4 + 4) + 4 * N
        rows.append({"T": T, "N": N, "us": us, "us_eager_incl_python_and_alloc": us_eager, "bytes": nbytes, "GBps": nbytes / us / 1e3,
                     "frac_of_peak": nbytes / us / 1e3 / peak, "kernel": "wl_gae_seg_kernel" if N <= 65536 else "wl_gae_kernel"})
    out["wl_gae"] = rows
    # ---- f-3: staged step (a + b) vs the fused step, and env.step with one Python reward term
    n = 4096
    a = wl.WheeledSim(wl.drift_task(num_envs=n, seed=1), dev); a.startup(); a.reset(None, 0)
    act = a.synth_actions(0)
    outs = tuple(torch.empty_like(x) for x in a.step(act, 0))
    rew, bits = a.step_stage_a(act, 1)
    sb = tuple(torch.empty_like(x) for x in a.step_stage_b(bits, 1))
    t = [2]

    def fused():
        a.step(act, t[0], out=outs); t[0] += 1

    def staged():
        a.step_stage_a(act, t[0], rew, bits); a.step_stage_b(bits, t[0], out=sb); t[0] += 1

    a.set_kernel_variant(1); f1 = timed(fused)
    a.set_kernel_variant(0); f4 = timed(fused)
    st = timed(staged)
    env = wl.ManagerBasedRLEnv(wl.drift_task(num_envs=n, seed=1), device=dev)
    env.add_reward_term("forward_speed", lambda e: e.scene["robot"].data.root_lin_vel_b[:, 0], weight=1.0)
    env.reset()
    py = timed(lambda: env.step(act))
    plain = wl.ManagerBasedRLEnv(wl.drift_task(num_envs=n, seed=1), device=dev); plain.reset()
    pl = timed(lambda: plain.step(act))
    out["staged_step_4096"] = {"fused_quad_us": f4, "fused_thread_per_env_us": f1, "stage_a_plus_b_us": st,
                               "env_step_plain_us": pl, "env_step_with_one_python_reward_term_us": py,
                               "note": "per call, launch to completion event, warm L2; the staged kernels use the thread-per-env code path"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
