"""In-kernel timeline of CTA 0 (debug library built by tools/build_trace_lib.sh, -DVIL_TRACE).
Prints, per traced thread, the mean cycles between consecutive trace tags over the steady-state blocks.
usage: VIL_ATTN_LIB=vision_longformer_b200/libvil_attn_sm100_trace.so python tools/trace_timeline.py [fwd|dq|dkv] [S1|S2]"""
import collections
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_longformer_b200 import _lib, vil_attention_raw_backward, vil_attention_raw_forward  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "dq"
tag = sys.argv[2] if len(sys.argv) > 2 else "S1"
H, M, nx, ny = {"S1": (3, 32, 56, 56), "S2": (3, 64, 28, 28)}[tag]
B, w, g = 256, 7, 1
N = g + nx * ny
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(300)
mk = lambda *s: torch.randn(*s, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
q, k, v, qg, go, gog = mk(B, H, nx * ny, M), mk(B, H, N, M), mk(B, H, N, M), mk(B, H, g, M), mk(B, H, nx * ny, M), mk(B, H, g, M)
o, og = torch.empty_like(q), torch.empty_like(qg)
dq, dk, dv, dqg = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v), torch.empty_like(qg)
kw = dict(nx=nx, ny=ny, w=w, exact=0, mode=0, scale=M ** -0.5)
lib = _lib.load()
buf = torch.zeros(8 * 2048, dtype=torch.int64, device=dev)
lse, lse_g = vil_attention_raw_forward(q, k, v, qg, k, v, None, None, None, o, og, **kw)
torch.cuda.synchronize()


def run():
    if which == "fwd":
        vil_attention_raw_forward(q, k, v, qg, k, v, None, None, None, o, og, skip_mask=1, **kw)
    else:
        vil_attention_raw_backward(q, k, v, qg, k, v, None, None, None, o, og, lse, lse_g, go, gog, dq, dk, dv, dqg, dk, dv,
                                   None, None, None, skip_mask=(1 | 4 | 8) if which == "dq" else (1 | 2 | 8), **kw)


run(); torch.cuda.synchronize()                      # warm
lib.vil_attn_debug_set_trace.argtypes = [ctypes.c_void_p]
assert lib.vil_attn_debug_set_trace(ctypes.c_void_p(buf.data_ptr())) == 0
run(); torch.cuda.synchronize()
lib.vil_attn_debug_set_trace(ctypes.c_void_p(0))
t = buf.cpu().view(8, 1024, 2)
names = {0: "compute/softmax thread 0", 1: "compute/softmax second traced thread", 2: "MMA issuer"}
for slot in range(3):
    ev = [(int(a), int(c)) for a, c in t[slot] if int(c) != 0]
    if not ev:
        continue
    print(f"--- slot {slot} ({names[slot]}): {len(ev)} events, span {ev[-1][1] - ev[0][1]} cycles")
    # skip the first 100 events (pipeline fill), then average the gap for every (prev tag -> tag) transition
    gaps = collections.defaultdict(list)
    for (a0, c0), (a1, c1) in zip(ev[100:-1], ev[101:]):
        gaps[(a0, a1)].append(c1 - c0)
    tot = sum(sum(vs) for vs in gaps.values())
    for (a0, a1), vs in sorted(gaps.items()):
        print(f"   {a0:3d} -> {a1:3d}: n={len(vs):4d} mean={sum(vs) / len(vs):8.1f} max={max(vs):7d}  share={100 * sum(vs) / tot:5.1f}%")
    print("   first 40 events (tag:delta):", " ".join(f"{a}:{c - ev[0][1]}" for a, c in ev[:40]))
