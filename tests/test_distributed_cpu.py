"""world_size-2 gloo test of the multi-GPU host logic: shard offsets, slab layout, the single
all-gather, and shard invariance (2 x N/2 shards == one N run) using the CPU oracle as the stepper."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
N_LOCAL, T = 24, 12


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, outdir):
    sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import wheeledlab_b200 as wl
    from wheeledlab_b200.distributed import RolloutSlab, shard_offset
    from oracle_lib import Oracle
    spec = wl.drift_task(num_envs=N_LOCAL, seed=5, env_id_offset=shard_offset(rank, N_LOCAL))
    orc = Oracle(spec.cfg); orc.startup(); orc.reset(None, 0)
    slab = RolloutSlab(T, N_LOCAL, 14, 2, "cpu", policy_fields=True)      # + values / log_prob / mean (written by wl_act_step on GPUs)
    gid = torch.arange(N_LOCAL, dtype=torch.float32) + rank * N_LOCAL
    for t in range(T):
        slab.values[t] = gid + 1000.0 * t; slab.log_prob[t] = -gid; slab.mean[t] = torch.stack([gid, gid + 0.5], dim=1)
        a = orc.synth_actions(t)
        obs, rew, term, trunc = orc.step(a, t)
        slab.actions[t] = torch.from_numpy(a); slab.obs[t] = torch.from_numpy(obs); slab.rewards[t] = torch.from_numpy(rew)
        slab.terminated[t] = torch.from_numpy(term); slab.truncated[t] = torch.from_numpy(trunc)
    g = slab.all_gather()
    if rank == 0:
        np.savez(Path(outdir) / "gathered.npz", obs=g.cat("obs").numpy(), rewards=g.cat("rewards").numpy(),
                 actions=g.cat("actions").numpy(), terminated=g.cat("terminated").numpy(), truncated=g.cat("truncated").numpy(),
                 values=g.cat("values").numpy(), log_prob=g.cat("log_prob").numpy(), mean=g.cat("mean").numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_run(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npz")
    sys.path[:0] = [str(ROOT / "tests")]
    import wheeledlab_b200 as wl
    from oracle_lib import Oracle
    spec = wl.drift_task(num_envs=world * N_LOCAL, seed=5)
    orc = Oracle(spec.cfg); orc.startup(); orc.reset(None, 0)
    for t in range(T):
        a = orc.synth_actions(t)
        obs, rew, term, trunc = orc.step(a, t)
        assert np.array_equal(got["actions"][t].view(np.uint32), a.view(np.uint32))
        assert np.array_equal(got["obs"][t].view(np.uint32), obs.view(np.uint32)), f"shard/gather mismatch at t={t}"
        assert np.array_equal(got["rewards"][t].view(np.uint32), rew.view(np.uint32))
        assert np.array_equal(got["terminated"][t], term) and np.array_equal(got["truncated"][t], trunc)
    gid = np.arange(world * N_LOCAL, dtype=np.float32)          # policy fields: global env id = rank * N_local + local id
    assert np.array_equal(got["values"], gid[None, :] + 1000.0 * np.arange(T, dtype=np.float32)[:, None])
    assert np.array_equal(got["log_prob"][3], -gid) and np.array_equal(got["mean"][5], np.stack([gid, gid + 0.5], 1))


def test_slab_layout_single_process():
    from wheeledlab_b200.distributed import RolloutSlab
    s = RolloutSlab(4, 8, 14, 2, "cpu")
    s.obs[2, 3, 5] = 7.0; s.rewards[1, 2] = -1.5; s.terminated[3, 7] = 1
    g = s.all_gather()
    assert g.cat("obs").shape == (4, 8, 14) and g.cat("obs")[2, 3, 5] == 7.0
    assert g.cat("rewards")[1, 2] == -1.5 and g.cat("terminated")[3, 7] == 1
    assert s.nbytes % 256 == 0
    obs, rew, term, trunc = s.step_outputs(1)
    assert obs.is_contiguous() and obs.data_ptr() == s.obs[1].data_ptr() and rew.shape == (8,)


def test_rollout_recording_uses_the_reference_playback_format(tmp_path):
    """play_policy.py:131-165: {'observations': [steps, N, D], 'actions': [steps, N, A]} torch-saved as <name>-rollouts.pt."""
    import torch
    from wheeledlab_b200.distributed import RolloutSlab
    from wheeledlab_b200.recording import load_rollouts, save_slab
    slab = RolloutSlab(6, 5, 14, 2, "cpu")
    slab.obs.copy_(torch.arange(6 * 5 * 14, dtype=torch.float32).view(6, 5, 14)); slab.actions.fill_(0.25)
    p = save_slab(str(tmp_path / "playback"), "demo", slab, steps=4)
    assert p.endswith("demo-rollouts.pt")
    d = load_rollouts(p)
    assert d["observations"].shape == (4, 5, 14) and d["actions"].shape == (4, 5, 2)
    assert torch.equal(d["observations"], slab.obs[:4]) and float(d["actions"].mean()) == 0.25


def test_bench_reference_arm_under_torchrun_prints_one_line(tmp_path):
    """`bench.py --impl reference` launched the way the driver launches it for N > 1 (torchrun, one process per GPU): rank 0
    alone times the CPU implementation and prints ONE JSON line with the bench contract's keys; the other ranks exit 0."""
    import json
    import subprocess
    import sys
    root = Path(__file__).resolve().parent.parent
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29741", str(root / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["n_gpus"] == 2 and d["steps"] == 3
    assert d["value"] > 0 and d["higher_is_better"] is True and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
