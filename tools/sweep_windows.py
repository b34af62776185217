"""BASELINE configs 4 / 5 at kernel level: window sweep w in {7, 8, 12, 15, 31} x g in {1, 8} at the Base-Deep 512^2 and
Medium-Deep 384^2 hot-layer shapes, bf16, local kernels only (CUDA events, median of 20).  Prints ms, the family that ran
and the tensor-core / HBM fractions computed from SURVEY.md section 8(d)'s algorithmic work.
usage: python tools/sweep_windows.py [bwd]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import algorithmic_work, peaks  # noqa: E402
from vision_longformer_b200 import _lib, vil_attention_raw_backward, vil_attention_raw_forward  # noqa: E402

do_bwd = "bwd" in sys.argv
dev = torch.device("cuda")
hbm, tflops, _ = peaks()
CASES = [  # (name, B, H, M, nx, ny, w, g)
    ("base-deep-512 S1", 8, 3, 32, 128, 128, 7, 1), ("base-deep-512 S1", 8, 3, 32, 128, 128, 15, 1),
    ("base-deep-512 S1", 8, 3, 32, 128, 128, 31, 1), ("base-deep-512 S1", 8, 3, 32, 128, 128, 15, 8),
    ("base-deep-512 S2", 8, 3, 64, 64, 64, 7, 1), ("base-deep-512 S2", 8, 3, 64, 64, 64, 15, 1),
    ("base-deep-512 S2", 8, 3, 64, 64, 64, 31, 1), ("base-deep-512 S2", 8, 3, 64, 64, 64, 31, 8),
    ("medium-deep-384 S1", 32, 3, 32, 96, 96, 8, 1), ("medium-deep-384 S2", 32, 3, 64, 48, 48, 12, 1),
]


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for name, B, H, M, nx, ny, w, g in CASES:
    N = g + nx * ny
    gen = torch.Generator(device=dev).manual_seed(300)
    mk = lambda *s: torch.randn(*s, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
    q, k, v, qg, go, gog = mk(B, H, nx * ny, M), mk(B, H, N, M), mk(B, H, N, M), mk(B, H, g, M), mk(B, H, nx * ny, M), mk(B, H, g, M)
    o, og = torch.empty_like(q), torch.empty_like(qg)
    dq, dk, dv, dqg = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v), torch.empty_like(qg)
    kw = dict(nx=nx, ny=ny, w=w, exact=0, mode=0, scale=M ** -0.5)
    fwd = lambda sk: vil_attention_raw_forward(q, k, v, qg, k, v, None, None, None, o, og, skip_mask=sk, **kw)
    lse, lse_g = fwd(0)
    fam = _lib.last_impl()
    flops, byts = algorithmic_work(nx, ny, w, g, H, M)
    ms = timeit(lambda: fwd(1))
    rec = {"case": name, "B": B, "w": w, "g": g, "family": fam, "fwd_local_ms": round(ms, 4),
           "fwd_tensor_frac": round(B * flops / (ms * 1e-3) / 1e12 / tflops, 4),
           "fwd_hbm_frac": round(B * byts / (ms * 1e-3) / 1e9 / hbm, 4)}
    if do_bwd:
        bwd = lambda sk: vil_attention_raw_backward(q, k, v, qg, k, v, None, None, None, o, og, lse, lse_g, go, gog, dq, dk, dv,
                                                    dqg, dk, dv, None, None, None, skip_mask=sk, **kw)
        bwd(0)
        rec["family_bwd"] = _lib.last_impl()
        mb = timeit(lambda: bwd(1 | 8))                   # both local passes, no prologue / global kernels
        rec["bwd_local_ms"] = round(mb, 4)
        rec["bwd_tensor_frac"] = round(B * 2 * flops / (mb * 1e-3) / 1e12 / tflops, 4)
    print(json.dumps(rec), flush=True)
