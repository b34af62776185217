"""A/B timing of the attention operator at BASELINE config-2 shapes (S1 / S2, B = 256, bf16): the fused round-2 kernels
against the round-1 multi-kernel pipeline (VIL_FLAG_UNFUSED), through the C ABI, CUDA events on the launching stream,
median of `reps` after 5 warm-ups; the working set of one call (> 300 MB) exceeds the 126 MB L2.
Layout "linear" = the production strided views of the q / kv Linear outputs (ops._heads).
usage: python tools/ab_kernels.py [linear|contig] [exact1] [reps=N]      env: VIL_FWD2_POLY, VIL_FWD2_P16"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_longformer_b200 import _lib, vil_attention_raw_backward, vil_attention_raw_forward  # noqa: E402

layout = "contig" if "contig" in sys.argv else "linear"
exact = 1 if "exact1" in sys.argv else 0
reps = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("reps=")), 30)
dev = torch.device("cuda")


def heads(t, H, which=0, parts=1):
    B, T, C = t.shape
    return t.view(B, T, parts, H, C // (parts * H))[:, :, which].permute(0, 2, 1, 3)


def timeit(fn):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return round(ts[len(ts) // 2], 4)


for tag, (H, D, nx, ny) in {"S1": (3, 32, 56, 56), "S2": (3, 64, 28, 28)}.items():
    B, w, g = 256, 7, 1
    N, C = g + nx * ny, H * D
    gen = torch.Generator(device=dev).manual_seed(300)
    mk = lambda *s: torch.randn(*s, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
    if layout == "linear":
        q_all, kv, d_out = mk(B, N, C), mk(B, N, 2 * C), mk(B, N, C)
        out, dq_all, dkv = torch.empty_like(q_all), torch.empty_like(q_all), torch.empty_like(kv)
        q, qg = heads(q_all, H)[:, :, g:], heads(q_all, H)[:, :, :g]
        k, v = heads(kv, H, 0, 2), heads(kv, H, 1, 2)
        o, og = heads(out, H)[:, :, g:], heads(out, H)[:, :, :g]
        go, gog = heads(d_out, H)[:, :, g:], heads(d_out, H)[:, :, :g]
        dq, dqg = heads(dq_all, H)[:, :, g:], heads(dq_all, H)[:, :, :g]
        dk, dv = heads(dkv, H, 0, 2), heads(dkv, H, 1, 2)
    else:
        q, k, v, qg, go, gog = mk(B, H, nx * ny, D), mk(B, H, N, D), mk(B, H, N, D), mk(B, H, g, D), mk(B, H, nx * ny, D), mk(B, H, g, D)
        o, og = torch.empty_like(q), torch.empty_like(qg)
        dq, dk, dv, dqg = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v), torch.empty_like(qg)
    kw = dict(nx=nx, ny=ny, w=w, exact=exact, mode=0, scale=D ** -0.5)
    for name, flags in (("fused", 0), ("unfused", _lib.VIL_FLAG_UNFUSED)):
        fwd = lambda sk: vil_attention_raw_forward(q, k, v, qg, k, v, None, None, None, o, og, skip_mask=sk, flags=flags, **kw)
        n0 = _lib.launch_count()
        lse, lse_g = fwd(0)
        n_f = _lib.launch_count() - n0
        bwd = lambda sk: vil_attention_raw_backward(q, k, v, qg, k, v, None, None, None, o, og, lse, lse_g, go, gog, dq, dk, dv,
                                                    dqg, dk, dv, None, None, None, skip_mask=sk, flags=flags, **kw)
        n0 = _lib.launch_count()
        bwd(0)
        n_b = _lib.launch_count() - n0
        res = dict(shape=tag, layout=layout, exact=exact, pipeline=name, launches_fwd=n_f, launches_bwd=n_b,
                   poly=os.environ.get("VIL_FWD2_POLY", ""), p16=os.environ.get("VIL_FWD2_P16", ""),
                   fwd_ms=timeit(lambda: fwd(0)), fwd_main_kernel_ms=timeit(lambda: fwd(1)),
                   bwd_ms=timeit(lambda: bwd(0)), bwd_dq_ms=timeit(lambda: bwd(1 | 4 | 8)), bwd_dkv_ms=timeit(lambda: bwd(1 | 2 | 8)))
        print(json.dumps(res), flush=True)
