"""Debug aid: tcgen05 vs SIMT family on the same inputs (forward + backward), per-output error / NaN report."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_longformer_b200 import _lib, vil_attention_raw_backward, vil_attention_raw_forward  # noqa: E402

dev = torch.device("cuda")
RPE = len(sys.argv) > 1 and sys.argv[1] == "rpe"
CASES = [(1, 1, 32, 14, 14, 0, 7, 0), (1, 1, 32, 14, 14, 1, 7, 0), (2, 3, 32, 56, 56, 1, 7, 0), (2, 3, 64, 28, 28, 1, 7, 0),
         (1, 2, 32, 21, 35, 1, 7, 0), (1, 2, 32, 24, 40, 1, 8, 0), (1, 2, 32, 22, 20, 1, 7, 5), (1, 2, 64, 18, 15, 1, 6, 0)]
for (B, H, D, nx, ny, g, w, mode) in CASES:
    gen = torch.Generator(device=dev).manual_seed(1)
    N = g + nx * ny
    mk = lambda *s: torch.randn(*s, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
    q, k, v, qg, go, gog = mk(B, H, nx * ny, D), mk(B, H, N, D), mk(B, H, N, D), mk(B, H, max(g, 1), D)[:, :, :g], mk(B, H, nx * ny, D), mk(B, H, max(g, 1), D)[:, :, :g]
    kw = dict(nx=nx, ny=ny, w=w, exact=0, mode=mode, scale=D ** -0.5)
    table = g2l = g2g = None
    if RPE:
        table = 0.5 * torch.randn((4 * w - 1) ** 2, H, generator=gen, device=dev)
        if g:
            g2l, g2g = 0.5 * torch.randn(2, H, g, generator=gen, device=dev), 0.5 * torch.randn(H, g, g, generator=gen, device=dev)
    res = {}
    for impl in ("simt", "tcgen05"):
        o, og = torch.empty_like(q), (torch.empty_like(qg) if g else None)
        G = lambda t: t if g else None
        lse, lse_g = vil_attention_raw_forward(q, k, v, G(qg), G(k), G(v), table, g2l, g2g, o, og, impl=impl, **kw)
        dq, dk, dv = torch.full_like(q, 7.0), torch.full_like(k, 7.0), torch.full_like(v, 7.0)
        dqg = torch.empty_like(qg) if g else None
        zl = lambda t: None if t is None else torch.zeros_like(t)
        dt, dgl, dgg = zl(table), zl(g2l), zl(g2g)
        vil_attention_raw_backward(q, k, v, G(qg), G(k), G(v), table, g2l, g2g, o, og, lse, lse_g, go, G(gog), dq, dk, dv,
                                   dqg, G(dk), G(dv), dt, dgl, dgg, impl=impl, **kw)
        torch.cuda.synchronize()
        res[impl] = dict(o=o.float(), lse=lse, dq=dq.float(), dk=dk.float(), dv=dv.float())
        if RPE:
            res[impl]["dtab"] = dt
    line = f"B{B} H{H} D{D} {nx}x{ny} g{g} w{w} m{mode}: "
    for n in ("o", "lse", "dq", "dk", "dv") + (("dtab",) if RPE else ()):
        a, b = res["tcgen05"][n], res["simt"][n]
        nan = int(torch.isnan(a).sum())
        err = ((a - b).norm() / b.norm()).item() if nan == 0 else float("nan")
        line += f"{n}: err={err:.2e} nan={nan}  "
        if nan:
            idx = torch.isnan(a).nonzero()
            line += f"first_nan={idx[0].tolist()} last_nan={idx[-1].tolist()} "
    print(line, flush=True)
