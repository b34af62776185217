"""Time the forward (and optionally backward) hot-path kernels alone at BASELINE config-2 shapes (CUDA events, L2 not flushed:
the working set of one call, > 300 MB, exceeds the 126 MB L2).   usage: python tools/time_fwd.py [bwd] [rpe]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_longformer_b200 import vil_attention_raw_backward, vil_attention_raw_forward  # noqa: E402

do_bwd = "bwd" in sys.argv
rpe = "rpe" in sys.argv
dev = torch.device("cuda")
for tag, (H, M, nx, ny) in {"S1": (3, 32, 56, 56), "S2": (3, 64, 28, 28)}.items():
    B, w, g = 256, 7, 1
    N = g + nx * ny
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

    def timeit(fn, reps=30):
        for _ in range(5):
            fn()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    res = {}
    fwd = lambda sk: vil_attention_raw_forward(q, k, v, qg, k, v, table, g2l, g2g, o, og, skip_mask=sk, **kw)
    res["fwd_local_ms"] = timeit(lambda: fwd(1))
    lse, lse_g = fwd(0)
    if do_bwd:
        bwd = lambda sk: vil_attention_raw_backward(q, k, v, qg, k, v, table, g2l, g2g, o, og, lse, lse_g, go, gog, dq, dk, dv,
                                                    dqg, dk, dv, dt, dgl, dgg, skip_mask=sk, **kw)
        res["bwd_dq_ms"] = timeit(lambda: bwd(1 | 4 | 8))
        res["bwd_dkv_ms"] = timeit(lambda: bwd(1 | 2 | 8))
        res["bwd_all_ms"] = timeit(lambda: bwd(0))
        res["bwd_global_ms"] = timeit(lambda: bwd(2 | 4 | 8))          # global-token row / column kernels only
        res["bwd_prep_ms"] = timeit(lambda: bwd(1 | 2 | 4))            # delta + chunk-ordered lse/delta only
        res["fwd_global_ms"] = timeit(lambda: fwd(2))
    print(tag, "rpe" if rpe else "norpe", {k2: round(v2, 4) for k2, v2 in res.items()}, flush=True)
