"""torch.profiler breakdown of one ViL-Small training step (where does the non-attention time go?)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_longformer_b200 import build_vil  # noqa: E402

dev = torch.device("cuda")
B = next((int(a) for a in sys.argv[1:] if a.isdigit()), 256)
if "cudnnbench" in sys.argv:
    torch.backends.cudnn.benchmark = True          # let cuDNN time its algorithms for the four patch-embedding convolutions
FUSED = "stock" not in sys.argv          # "stock": the plain PyTorch residual / bias composition (fused_residual=False)
net = build_vil("vil_small", img_size=224, fused_residual=FUSED).to(dev).train()
opt = torch.optim.AdamW(net.parameters(), lr=5e-4, weight_decay=0.05, fused=True)
x = torch.randn(B, 3, 224, 224, device=dev)
y = torch.randint(0, 1000, (B,), device=dev)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.functional.cross_entropy(net(x), y)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("fused_residual", FUSED, "batch", B, "ms/step median", sorted(ts)[5], "img/s", B / sorted(ts)[5] * 1e3, flush=True)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
