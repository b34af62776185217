"""Dense attention of the ViL-Small `s0` stages (14x14+1 tokens, H=6, D=64 and 7x7, H=12, D=64), B=256, bf16:
the operator's single-chunk kernels (impl='vil') against stock SDPA / cuDNN (impl='sdpa'); module fwd+bwd incl. the
qkv / proj Linears, CUDA events, median of 20."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_longformer_b200 import _lib  # noqa: E402
from vision_longformer_b200.msvit import DenseAttention  # noqa: E402

dev = torch.device("cuda")
for dim, H, w, g in ((384, 6, 14, 1), (768, 12, 7, 0)):
    x = torch.randn(256, g + w * w, dim, device=dev, requires_grad=True)
    gy = torch.randn(256, g + w * w, dim, device=dev)
    for impl in ("sdpa", "vil"):
        mod = DenseAttention(dim, num_heads=H, qkv_bias=True, rpe=False, wx=w, wy=w, nglo=g, impl=impl).to(dev)

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = mod(x)
            y.backward(gy.to(y.dtype))

        def fwd():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                mod(x)
        res = {}
        for name, fn in (("fwd_bwd_ms", step), ("fwd_ms", fwd)):
            for _ in range(5):
                fn()
            ts = []
            for _ in range(20):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            res[name] = round(ts[10], 4)
        print(json.dumps(dict(dim=dim, heads=H, tokens=g + w * w, impl=impl, family=_lib.last_impl() if impl == "vil" else "cudnn", **res)), flush=True)
