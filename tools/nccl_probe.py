import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
x = torch.ones(36_700_160, dtype=torch.uint8, device="cuda")
out = torch.empty(world * x.numel(), dtype=torch.uint8, device="cuda")
for _ in range(5):
    dist.all_gather_into_tensor(out, x)
torch.cuda.synchronize(); dist.barrier()
for nb in (1 << 20, 36_700_160):
    xs = x[:nb]; os_ = out[: world * nb]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_gather_into_tensor(os_, xs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0:
        print(f"PROBE world={world} bytes/rank={nb} all_gather {ms:.3f} ms  busbw={(world-1)*nb/ms/1e6:.1f} GB/s", flush=True)
dist.destroy_process_group()
