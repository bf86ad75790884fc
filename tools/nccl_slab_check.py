#!/usr/bin/env python
"""torchrun helper of tests/test_gpu_parity.py::test_two_process_nccl_gather_equals_single_rank.

Each rank owns `--envs` envs (global ids rank*envs ...), fills a rollout slab for `--steps` steps with the synthetic
actions of the counter-based generator, all-gathers it over NCCL and rank 0 saves the learner-facing [T, W*N, ...] fields."""
import argparse
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--envs", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--mode", default="nccl", choices=["nccl", "fanout", "mcast", "ce"])
    a = ap.parse_args()
    import wheeledlab_b200 as wl
    from wheeledlab_b200.distributed import RolloutSlab
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device(dev))
    sim = wl.WheeledSim(wl.drift_task(num_envs=a.envs, seed=a.seed, env_id_offset=rank * a.envs), dev)
    sim.startup(); sim.reset(None, 0)
    fields = ("obs", "actions", "rewards", "terminated", "truncated")
    if a.mode in ("fanout", "mcast", "ce"):
        from wheeledlab_b200.distributed import SymmetricRolloutSlab
        sym = SymmetricRolloutSlab(a.steps, a.envs, sim.obs_dim, 2, dev)
        slab = sym.slab
        if a.mode in ("fanout", "mcast"):        # fused: the step kernel stores its rows into every peer's symmetric buffer
            if a.mode == "mcast" and sym.mc_delta == 0:              # (mcast: one multimem.st replicated by the NVSwitch)
                print("NO_MULTICAST", flush=True)
                dist.barrier(); dist.destroy_process_group(); sys.exit(77)
            sym.attach(sim, multicast=(a.mode == "mcast"))
            fields = ("obs", "rewards", "terminated", "truncated")   # (actions are written by the caller, not by env.step)
    else:
        slab = RolloutSlab(a.steps, a.envs, sim.obs_dim, 2, dev)
    for t in range(a.steps):
        act = sim.synth_actions(t)
        slab.actions[t].copy_(act)
        sim.step(act, t, out=slab.step_outputs(t))
    if a.mode in ("fanout", "mcast"):
        sym.barrier(); torch.cuda.synchronize(); dist.barrier()
        g = sym.gathered()
    elif a.mode == "ce":                         # copy-engine pull of the peers' slabs
        g = sym.gather_ce()
    else:
        g = slab.all_gather()
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({k: g.cat(k).cpu() for k in fields}, a.out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
