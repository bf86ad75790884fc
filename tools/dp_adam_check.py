#!/usr/bin/env python
"""torchrun helper: DataParallelAdam (gradient all-reduce fused into the Adam kernel over symmetric memory) vs torch.optim.Adam
on the rank-averaged gradient.  Every rank prints/asserts; rank 0 writes 'ok' to --out."""
import argparse, os, sys
from pathlib import Path
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--out", required=True); a = ap.parse_args()
    from wheeledlab_b200.learner import DataParallelAdam
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local); dev = f"cuda:{local}"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device(dev))
    torch.manual_seed(0)                                     # identical initial replicas
    mk = lambda: torch.nn.Sequential(torch.nn.Linear(14, 64), torch.nn.ELU(), torch.nn.Linear(64, 64), torch.nn.ELU(), torch.nn.Linear(64, 2)).to(dev)
    net, ref = mk(), mk()
    ref.load_state_dict(net.state_dict())
    opt = DataParallelAdam(net.parameters(), lr=1e-3)
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    g = torch.Generator(device=dev).manual_seed(100 + rank)  # different data per rank
    for it in range(5):
        x = torch.randn(256, 14, generator=g, device=dev); y = torch.randn(256, 2, generator=g, device=dev)
        opt.zero_grad(); ((net(x) - y) ** 2).mean().backward(); opt.step()
        ropt.zero_grad(); ((ref(x) - y) ** 2).mean().backward()
        for p in ref.parameters():
            dist.all_reduce(p.grad); p.grad /= world
        ropt.step()
    torch.cuda.synchronize()
    err = max(float((p - q).abs().max()) for p, q in zip(net.parameters(), ref.parameters()))
    assert err < 2e-6, err
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    allp = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(allp, flat)
    assert all(torch.equal(allp[0], q) for q in allp), "replicas diverged"
    if rank == 0:
        Path(a.out).write_text(f"ok {err:.3e}")
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
