"""Launch only the residual / LayerNorm / bias epilogue kernels at the ViL-Small stage-1 / stage-2
token streams - the target of `ncu --set full` captures.   usage: python tools/epilogue_only.py [S1|S2] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_longformer_b200 import B200LayerNorm, epilogue  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "S1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N, C = {"S1": (1 + 56 * 56, 96), "S2": (1 + 28 * 28, 192)}[tag]
B = 256
dev = torch.device("cuda")
ln = B200LayerNorm(C, eps=1e-6).to(dev)
bias = torch.zeros(C, device=dev, requires_grad=True)
b1 = torch.zeros(4 * C, device=dev, requires_grad=True)
lin = torch.nn.Linear(C, 2 * C).to(dev)
scale = torch.ones(B, device=dev)
for _ in range(reps):
    x = torch.randn(B, N, C, device=dev, requires_grad=True)
    br = torch.randn(B, N, C, device=dev).to(torch.bfloat16).requires_grad_(True)
    z = torch.randn(B, N, 4 * C, device=dev).to(torch.bfloat16).requires_grad_(True)
    xo, y = epilogue.add_norm(x, br, bias, scale, ln, out_dtype=torch.bfloat16)
    (xo.sum() + y.float().sum()).backward()
    epilogue.bias_gelu(z, b1).float().sum().backward()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        epilogue.linear_colsum_bias(y.detach(), lin.weight, lin.bias).float().sum().backward()
torch.cuda.synchronize()
print("done", tag, reps)
