"""Launch only the hot-path kernels (BASELINE config 2 shapes) - the target of `ncu --set full` captures.
usage: python tools/kernel_only.py [S1|S2] [reps] [rpe]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_longformer_b200 import vil_attention_raw_backward, vil_attention_raw_forward  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "S1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rpe = len(sys.argv) > 3 and sys.argv[3] == "rpe"
H, M, nx, ny = {"S1": (3, 32, 56, 56), "S2": (3, 64, 28, 28)}[tag]
B, w, g = 256, 7, 1
N = g + nx * ny
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(300)
mk = lambda *s: torch.randn(*s, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
q, k, v, qg, go, gog = mk(B, H, nx * ny, M), mk(B, H, N, M), mk(B, H, N, M), mk(B, H, g, M), mk(B, H, nx * ny, M), mk(B, H, g, M)
o, og = torch.empty_like(q), torch.empty_like(qg)
dq, dk, dv, dqg = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v), torch.empty_like(qg)
table = g2l = g2g = dt = dgl = dgg = None
if rpe:
    table = 0.02 * torch.randn((4 * w - 1) ** 2, H, device=dev)
    g2l, g2g = 0.02 * torch.randn(2, H, g, device=dev), 0.02 * torch.randn(H, g, g, device=dev)
    dt, dgl, dgg = torch.zeros_like(table), torch.zeros_like(g2l), torch.zeros_like(g2g)
kw = dict(nx=nx, ny=ny, w=w, exact=0, mode=0, scale=M ** -0.5)
for _ in range(reps):
    lse, lse_g = vil_attention_raw_forward(q, k, v, qg, k, v, table, g2l, g2g, o, og, **kw)
    vil_attention_raw_backward(q, k, v, qg, k, v, table, g2l, g2g, o, og, lse, lse_g, go, gog, dq, dk, dv, dqg, dk, dv,
                               dt, dgl, dgg, **kw)
torch.cuda.synchronize()
print("done", tag, reps)
