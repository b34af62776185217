"""torch.profiler view of the DDP training step on N GPUs (run under torchrun): which NCCL kernels run, how long, and how
much of the all-reduce is exposed (step time with DDP minus step time of the same rank without gradient sync).
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/profile_step_ddp.py [bf16hook]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_longformer_b200 import build_vil  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
net = build_vil("vil_small", img_size=224).to(dev).train()
ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], gradient_as_bucket_view=True, static_graph=True)
if "bf16hook" in sys.argv:
    from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
    ddp.register_comm_hook(None, default_hooks.bf16_compress_hook)
opt = torch.optim.AdamW(ddp.parameters(), lr=5e-4, weight_decay=0.05, fused=True)
x = torch.randn(256, 3, 224, 224, device=dev)
y = torch.randint(0, 1000, (256,), device=dev)


def step(sync=True):
    ctx = ddp.no_sync() if not sync else torch.autograd.profiler.record_function("ddp_step")
    with ctx:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(ddp(x), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
    opt.step()


def timed(sync, n=10):
    for _ in range(3):
        step(sync)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step(sync)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_sync, t_nosync = timed(True), timed(False)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step(True)
    torch.cuda.synchronize()
if rank == 0:
    print(f"world {world}  hook {'bf16_compress' if 'bf16hook' in sys.argv else 'none (fp32 all-reduce)'}")
    print(f"ms/step with gradient all-reduce {t_sync:.2f}   without (no_sync) {t_nosync:.2f}   exposed DDP cost {t_sync - t_nosync:.2f}")
    rows = [e for e in prof.key_averages() if "nccl" in e.key.lower()]
    for e in sorted(rows, key=lambda e: -e.device_time_total):
        print(f"  {e.key[:90]:90s} calls/step {e.count / 3:5.1f}  us/step {e.device_time_total / 3:9.1f}")
dist.barrier()
dist.destroy_process_group()
