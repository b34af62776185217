"""Bring-up check of the key-row-block forward (vil_tc_fwd5) against the fp64 oracle: lean w = 7 / D <= 32 geometries, printing the
errors and the kernel variant that ran.  usage: python tools/check_fwd5.py [f32out] [fp16]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_parity import kernel_run, make_inputs, oracle_run  # noqa: E402
from tests.util import relerr  # noqa: E402
from vision_longformer_b200 import _lib  # noqa: E402

CASES = [(2, 3, 32, 28, 28, 1, 7, 0, 0, False), (1, 2, 32, 21, 35, 2, 7, 0, 0, False), (1, 2, 32, 7, 7, 1, 7, 0, 0, False),
         (1, 2, 32, 7, 21, 0, 7, 0, 0, False), (1, 2, 16, 14, 14, 1, 7, 0, 0, False), (1, 1, 32, 56, 56, 8, 7, 0, 0, False),
         (2, 3, 32, 56, 56, 1, 7, 0, 0, False)]
f32out = "f32out" in sys.argv
dtype = torch.float16 if "fp16" in sys.argv else torch.bfloat16
for case in CASES:
    B, H, D, nx, ny, g, w, exact, mode, rpe = case
    t = make_inputs(B, H, D, nx, ny, g, w, rpe, seed=301)
    ref = oracle_run(t, nx, ny, w, exact, mode, D ** -0.5, dtype)
    try:
        out, ff, fb = kernel_run(t, nx, ny, w, exact, mode, D ** -0.5, dtype, "auto", layout="linear", f32out=f32out)
    except Exception as e:  # noqa: BLE001
        print(case, "ERROR", repr(e)[:300], flush=True)
        continue
    names = ["o", "lse", "dq", "dk", "dv"] + (["og", "lse_g", "dqg"] if g else [])
    print(case, ff, fb, _lib.last_kernel(), {n: "%.2e" % relerr(out[n], ref[n]) for n in names}, flush=True)
