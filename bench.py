#!/usr/bin/env python
"""bench.py -- env-steps/s of the fused WheeledLab step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one env.step() of the RSS_DRIFT_CONFIG workload (MushrDriftRL: dt 5 ms x 4, DR + pushes +
obs noise on, auto-reset) over E envs per GPU (default 4096 = BASELINE configs[1]) with synthetic
U[-1,1]^2 actions from the counter-based generator.  Rank 0 prints ONE JSON line.

  value     whole-job env-steps/s, inputs resident in HBM, L2 flushed between timed steps, per-step CUDA
            events on the launch stream, max over ranks.
  e2e       same metric through the public ManagerBasedRLEnv.step() with HOST (pinned) action buffers:
            H2D of the actions and D2H of reward + done masks inside the timed region, every step.
  roofline  algorithmic bytes of the step kernel / its measured duration vs the measured HBM peak.
  cpu_baseline  the CPU oracle (oracle/wl_oracle.c, "port") timed on this box's host cores.
--impl reference times that same CPU implementation (all host threads) as the reference arm: the
reference's own PhysX pipeline is a closed binary that is not in /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
# algorithmic bytes per env-step of the Drift step kernel (DESIGN.md "Kernels"):
# reads 13 state/param groups x16 B + action 8 B; writes 9 state groups x16 B + obs 56 + rew 4 + 2 masks
BYTES_PER_ENV_STEP = (13 * 16 + 8) + (9 * 16 + 56 + 4 + 2)


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        self.gpu = ",".join(str(g) for g in (gpu_indices if isinstance(gpu_indices, (list, tuple, range)) else [gpu_indices]))
        self.rows, self.proc = [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", self.gpu], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons, per = [], [], set(), {}
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                per.setdefault(int(r[0]), []).append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if len(per) > 1:                                      # one median per GPU of the job
            out["per_gpu_sm_mhz"] = [statistics.median(per[g]) for g in sorted(per)]
        return out


def _oracle(spec, native=True, threads=1):
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_lib import Oracle
    return Oracle(spec.cfg, kind="native" if native else "f32", threads=threads)


def _best_threads(spec, probe_s: float = 0.4) -> int:
    """The box's cores may be shared/limited: probe a few OpenMP widths and keep the fastest."""
    ncpu = os.cpu_count() or 1
    cands = sorted({1, 2, 4, 8, 16, 32, 64, ncpu} & set(range(1, ncpu + 1)) | {ncpu})
    best, best_rate = 1, 0.0
    for th in cands:
        orc = _oracle(spec, native=True, threads=th)
        orc.startup(); orc.reset(None, 0)
        a = orc.synth_actions(0)
        orc.step(a, 0)
        t0 = time.perf_counter(); k = 0
        while time.perf_counter() - t0 < probe_s:
            orc.step(a, 1 + k); k += 1
        rate = k / (time.perf_counter() - t0)
        if rate > best_rate:
            best, best_rate = th, rate
    return best


def cpu_baseline(envs: int, seed: int, budget_s: float = 12.0, threads: int | None = None):
    """Time the CPU oracle on a bounded sample: `envs` envs, as many steps as fit in ~budget_s."""
    import wheeledlab_b200 as wl
    spec = wl.drift_task(num_envs=envs, seed=seed)
    threads = threads or _best_threads(spec)
    orc = _oracle(spec, native=True, threads=threads)
    orc.startup(); orc.reset(None, 0)
    acts = [orc.synth_actions(t) for t in range(8)]
    for t in range(3):
        orc.step(acts[t % 8], t)
    t0 = time.perf_counter(); steps = 0
    while True:
        orc.step(acts[steps % 8], 3 + steps); steps += 1
        el = time.perf_counter() - t0
        if el > budget_s or steps >= 2000:
            break
    return {"value": envs * steps / el, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{envs} envs x {steps} env-steps of RSS_DRIFT (oracle/wl_oracle.c -O3 -march=native, OpenMP {threads} threads), {el:.1f} s"}


def run_reference(args):
    """Reference arm: the CPU implementation of the path on this box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import wheeledlab_b200 as wl
    envs = args.envs
    spec = wl.drift_task(num_envs=envs, seed=args.seed)
    threads = _best_threads(spec)        # "all the host threads it can use": widest setting that actually scales
    orc = _oracle(spec, native=True, threads=threads)
    orc.startup(); orc.reset(None, 0)
    acts = [orc.synth_actions(t) for t in range(8)]
    for t in range(args.warmup):
        orc.step(acts[t % 8], t)
    t0 = time.perf_counter()
    for k in range(args.steps):
        orc.step(acts[k % 8], args.warmup + k)
    el = time.perf_counter() - t0
    val = envs * args.steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"RSS_DRIFT_CONFIG {envs} envs (CPU, one box)", "envs_per_step": envs,
                   "note": "PhysX is not runnable here; this is the CPU restatement of the same step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{envs} envs x {args.steps} env-steps, OpenMP {threads} of {os.cpu_count()} threads (best of a width probe)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class _StdoutGuard:
    """Everything any library prints to fd 1 (e.g. NCCL's version banner) goes to stderr; emit() writes the ONE JSON line
    to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self._real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text: str):
        os.write(self._real, (text + "\n").encode())


def run_ours(args):
    guard = _StdoutGuard()
    import torch
    import torch.distributed as dist
    import wheeledlab_b200 as wl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's version/warn lines must not pollute the ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device(dev))

    E, K, W = args.envs, args.steps, args.warmup
    spec = wl.drift_task(num_envs=E, seed=args.seed, env_id_offset=rank * E)
    sim = wl.WheeledSim(spec, dev)
    sim.startup(); sim.reset(None, 0)
    from wheeledlab_b200.distributed import RolloutSlab
    T_ROLL = 128                                                                  # rsl_rl num_steps_per_env (rsl_rl_ppo_cfg.py:6)
    acts = torch.stack([sim.synth_actions(t) for t in range(W + K)])           # resident in HBM
    slabs = [RolloutSlab(T_ROLL, E, sim.obs_dim, 2, dev) for _ in range(2 if world > 1 else 1)]   # the step writes straight into
    slab = slabs[0]                                                               # the send buffer; two of them so that the
    outs = slab.step_outputs(0)                                                   # gather of iteration i overlaps rollout i+1
    sim.step(acts[0], 0, out=outs)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2
    peak, peak_src = _peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- device-resident timing: per-step events around the step launch, L2 flushed between steps ----
    t = 1
    for _ in range(W):
        sim.step(acts[t % (W + K)], t, out=outs); flush.fill_(0.0); t += 1
    if world > 1:                                       # untimed: NCCL communicator / channel set-up, receive buffers
        for _ in range(2):
            for sl in slabs:
                sl.all_gather()
    sampler = ClockSampler(list(range(world)) if world > 1 else local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = sim.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    gev, wev, gathered = [], [], None
    gstream = torch.cuda.Stream(device=dev) if world > 1 else None
    gdone = [None] * len(slabs)                         # per slab: event after which its rows may be overwritten again
    main = torch.cuda.current_stream()
    # pointers resolved once per (action row, slab row): the timed loop is event / one ctypes call / event / flush
    n_it = (K + T_ROLL - 1) // T_ROLL
    bound = [sim.bind_step(acts[(W + k) % (W + K)], slabs[(k // T_ROLL) % len(slabs)].step_outputs(k % T_ROLL)) for k in range(K)]
    for k in range(K):
        row, cur = k % T_ROLL, (k // T_ROLL) % len(slabs)
        if world > 1 and row == 0 and gdone[cur] is not None:
            # the slab about to be refilled must have been gathered: any time the step stream has to WAIT for that is the
            # exposed (non-overlapped) cost of the collective and is added to the step times
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record(); main.wait_event(gdone[cur]); w1.record(); wev.append((w0, w1))
        ev[k][0].record()
        bound[k](t); t += 1
        ev[k][1].record()
        flush.fill_(0.0)
        if world > 1 and row == T_ROLL - 1:            # one all-gather of the rollout slab per PPO iteration, on its own
            filled = torch.cuda.Event(); filled.record()       # stream: it overlaps the next iteration's steps
            with torch.cuda.stream(gstream):
                gstream.wait_event(filled)
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(); gathered = slabs[cur].all_gather(); g1.record(); gev.append((g0, g1))
                gdone[cur] = g1
    if world > 1:                                       # the last gathers must be finished before the clock stops
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for g in gdone:
            if g is not None:
                main.wait_event(g)
        w1.record(); wev.append((w0, w1))
    barrier()
    launches = sim.launch_count - l0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    gather_each = [a.elapsed_time(b) for a, b in gev]
    gather_ms = sum(gather_each)                        # duration of the collectives on their own stream (mostly hidden)
    exposed_ms = sum(a.elapsed_time(b) for a, b in wev) # what the step stream actually waited for them
    if gev and rank == 0:
        print(f"[bench] all-gather ms per call: {[round(x, 3) for x in gather_each]}, exposed {exposed_ms:.3f} ms", file=sys.stderr)
    tot_ms = sum(step_ms) + exposed_ms
    # ---- warm-L2, CUDA-graph replay of K steps (supplementary: how the loop is meant to be driven) ----
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for k in range(K):
                sim.step(acts[(W + k) % (W + K)], t + k, out=outs)
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); barrier()
    graph_ms = e0.elapsed_time(e1)
    # ---- end-to-end through the public API with host buffers ----
    env = wl.ManagerBasedRLEnv(wl.drift_task(num_envs=E, seed=args.seed, env_id_offset=rank * E), device=dev)
    env.reset()
    h_act = acts.cpu().pin_memory()
    h_rows = [h_act[k] for k in range(W + K)]           # row views of the pinned block (what a host-side policy hands over)

    def e2e_step(k):
        # the call a host-side user makes: host actions in, host reward / done masks out (one C-ABI call inside:
        # H2D actions -> fused step -> D2H results -> stream sync); observations stay on the device for the policy
        # h_rows[k]: this step's actions in PINNED host memory (a different block every step), read in place
        obs, rew, term, trunc, extras = env.step_host(h_rows[k % (W + K)])
        return rew, term, trunc

    def time_e2e(transport):
        env.host_transport = transport
        for k in range(W):
            e2e_step(k)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(K):
            e2e_step(W + k)
        e1.record(); barrier()
        return e0.elapsed_time(e1)

    e2e_copy_ms = time_e2e("copy")
    e2e_ms = time_e2e("zero_copy")
    # ---- policy in the loop: 128 x (64x64 ELU MLP -> step -> slab row) captured as ONE CUDA graph (supplementary) ----
    pil = None
    try:
        if args.no_extras:
            raise RuntimeError('skipped (--no-extras)')
        from wheeledlab_b200.rollout import GraphedRollout
        torch.manual_seed(0)
        mlp = torch.nn.Sequential(torch.nn.Linear(sim.obs_dim, 64), torch.nn.ELU(), torch.nn.Linear(64, 64), torch.nn.ELU(),
                                  torch.nn.Linear(64, 2)).to(dev)
        sim_p = wl.WheeledSim(wl.drift_task(num_envs=E, seed=args.seed, env_id_offset=rank * E), dev)
        sim_p.startup(); sim_p.reset(None, 0)
        with torch.no_grad():
            roll = GraphedRollout(sim_p, lambda o: mlp(o), T_ROLL).capture(0)
            roll.run(); barrier()
            R = 4
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(R):
                roll.run()
            p1.record(); barrier()
        pil_ms = max_over_ranks(p0.elapsed_time(p1)) if world > 1 else p0.elapsed_time(p1)
        pil = {"value": E * world * T_ROLL * R / (pil_ms * 1e-3), "unit": UNIT, "ms_per_step": pil_ms / (T_ROLL * R),
               "note": "rsl_rl-sized actor (14-64-64-2 ELU, torch/cuBLAS) + fused env step, 128 steps per CUDA-graph launch"}
    except Exception as ex:                                      # supplementary figure only
        pil = {"error": repr(ex)[:200]}
    # ---- actor + critic + Gaussian sampling FUSED into the step kernel (wl_act_step), 128 launches per graph (supplementary) ----
    pfu = None
    try:
        if args.no_extras:
            raise RuntimeError('skipped (--no-extras)')
        from wheeledlab_b200.policy import FusedPolicyRollout, pack_actor_critic
        torch.manual_seed(0)
        mk = lambda out: torch.nn.Sequential(torch.nn.Linear(sim.obs_dim, 64), torch.nn.ELU(), torch.nn.Linear(64, 64), torch.nn.ELU(),
                                             torch.nn.Linear(64, out)).to(dev)
        blob = pack_actor_critic(mk(2), mk(1), torch.ones(2), sim.obs_dim, dev)
        sim_q = wl.WheeledSim(wl.drift_task(num_envs=E, seed=args.seed, env_id_offset=rank * E), dev)
        sim_q.startup(); sim_q.reset(None, 0)
        froll = FusedPolicyRollout(sim_q, blob, T_ROLL).capture(0)
        froll.run(); barrier()
        R = 4
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(R):
            froll.run()
        q1.record(); barrier()
        pfu_ms = max_over_ranks(q0.elapsed_time(q1)) if world > 1 else q0.elapsed_time(q1)
        pfu = {"value": E * world * T_ROLL * R / (pfu_ms * 1e-3), "unit": UNIT, "ms_per_step": pfu_ms / (T_ROLL * R),
               "note": "rsl_rl actor AND critic (14-64-64-2/1 ELU), Gaussian sample + log-prob, and the env step in ONE kernel; "
                       "128 launches per CUDA graph"}
    except Exception as ex:                                      # supplementary figure only
        pfu = {"error": repr(ex)[:200]}
    # ---- fused K-step synthetic rollout: K env.steps per launch, state in registers, in-kernel actions (supplementary) ----
    fused = None
    try:
        if args.no_extras:
            raise RuntimeError('skipped (--no-extras)')
        KF = 125                                                  # divides the 250-step episode: windows end on curriculum boundaries
        sim_f = wl.WheeledSim(wl.drift_task(num_envs=E, seed=args.seed, env_id_offset=rank * E), dev)
        sim_f.startup(); sim_f.reset(None, 0)
        slab_f = RolloutSlab(KF, E, sim_f.obs_dim, 2, dev)
        logs_f = torch.empty((KF, 16), dtype=torch.float32, device=dev)
        sim_f.rollout(KF, 0, slab_f, logs_f); barrier()
        RF = 8
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for r in range(RF):
            sim_f.rollout(KF, KF * (1 + r), slab_f, logs_f)
        f1.record(); barrier()
        f_ms = max_over_ranks(f0.elapsed_time(f1))
        fused = {"value": E * world * KF * RF / (f_ms * 1e-3), "unit": UNIT, "ms_per_step": f_ms / (KF * RF), "K": KF,
                 "note": "wl_rollout: K env.steps per launch, state in registers, in-kernel U[-1,1]^2 actions; every step still "
                         "writes its obs/action/reward/done slab rows and episode-log row; bit-identical to K wl_step calls"}
    except Exception as ex:
        fused = {"error": repr(ex)[:200]}
    clocks = sampler.stop() if rank == 0 else None

    tot_ms, graph_ms, e2e_ms = max_over_ranks(tot_ms), max_over_ranks(graph_ms), max_over_ranks(e2e_ms)
    e2e_copy_ms = max_over_ranks(e2e_copy_ms)
    gather_ms = max_over_ranks(gather_ms)
    exposed_ms = max_over_ranks(exposed_ms)
    per_rank = None
    if world > 1:                                       # diagnostics: which rank sets the max
        st = torch.tensor([statistics.mean(step_ms), statistics.median(step_ms), max(step_ms)], dtype=torch.float64, device=dev)
        allst = [torch.zeros_like(st) for _ in range(world)]
        dist.all_gather(allst, st)
        per_rank = {"step_us_mean": [round(float(x[0]) * 1e3, 3) for x in allst],
                    "step_us_median": [round(float(x[1]) * 1e3, 3) for x in allst],
                    "step_us_max": [round(float(x[2]) * 1e3, 3) for x in allst]}
    if rank == 0:
        total_envs = E * world
        value = total_envs * K / (tot_ms * 1e-3)
        kern_s = statistics.mean(step_ms) * 1e-3
        achieved = BYTES_PER_ENV_STEP * E / kern_s / 1e9
        cpu = cpu_baseline(E, args.seed, budget_s=args.cpu_budget) if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": tot_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"RSS_DRIFT_CONFIG MushrDriftRL {E} envs/GPU x {world} GPU, dt 5ms x4, DR+push+noise on",
                       "envs_per_gpu": E, "global_envs": total_envs, "actions": "U[-1,1]^2 philox(seed,env,step)",
                       "l2": "flushed between timed steps (256 MiB fill)", "parallelism": f"env-shard x{world}"},
            "clocks": clocks,
            "e2e": {"value": total_envs * K / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": E * 8,
                    "d2h_bytes_per_step": E * 6, "ms_per_step": e2e_ms / K,
                    "staged_copy_transport_ms_per_step": e2e_copy_ms / K,
                    "api": "ManagerBasedRLEnv.step_host(pinned actions) -> wl_step_host_zero_copy: the kernel reads the "
                           "actions from / writes reward+dones to pinned host memory over PCIe, then stream sync, every step"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of
                         # this kernel at this size (profiles/r01_ncu_v3_step_4096.txt: 931 072 B read, 0 B written -- the
                         # 1.7 MB working set is written back from L2 later, so DRAM traffic < algorithmic bytes)
                         "traffic": 931072 if E == 4096 else None, "traffic_unit": "bytes/launch", "peak_source": peak_src, "kernel": "wl_step_kernel<DRIFT>",
                         "bytes_per_env_step": BYTES_PER_ENV_STEP, "avg_kernel_us": kern_s * 1e6,
                         "kernel_variant": "quad (4 lanes/env)" if E <= 148 * 4 * 32 * 2 else "thread-per-env",
                         "note": "N=4096 moves 1.7 MB/launch: launch-latency bound, see profiles/ for the N sweep"},
            "cpu_baseline": cpu,
            "collective": {"kind": "all_gather_into_tensor(rollout slab)", "per_iteration_steps": T_ROLL,
                           "bytes_per_rank": slab.nbytes, "count": len(gev), "ms_total": gather_ms,
                           "overlapped_with_next_iteration": True, "exposed_ms_total": exposed_ms,
                           "note": "double-buffered slabs; the gather of iteration i runs on its own stream under the steps of "
                                   "iteration i+1; `value` charges the time the step stream waited for it (exposed_ms_total)"}
            if world > 1 else None,
            "per_rank": per_rank,
            "policy_in_loop_graph": pil,
            "policy_fused_in_step": pfu,
            "rollout_fused": fused,
            "warm_l2_graph": {"value": total_envs * K / (graph_ms * 1e-3), "unit": UNIT, "ms_per_step": graph_ms / K,
                              "note": "K steps captured in one CUDA graph, state L2-resident (supplementary)"},
        }
        guard.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-extras", action="store_true", help="skip the supplementary figures (policy-in-loop, fused rollout)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
