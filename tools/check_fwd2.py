"""Bring-up check of the fused kernels against the fp64 oracle (small cases, fwd + bwd), printing the errors instead of
asserting - run with VIL_FWD2_P16=0/1, VIL_FWD2_POLY=0/2/4 to compare operand formats / polynomial exp2 fractions."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_parity import kernel_run, make_inputs, oracle_run  # noqa: E402
from tests.util import relerr  # noqa: E402

CASES = [(2, 3, 32, 28, 28, 1, 7, 0, 0, False), (1, 3, 64, 28, 28, 1, 7, 0, 0, False), (1, 2, 32, 21, 35, 2, 7, 0, 0, False),
         (1, 2, 32, 20, 22, 1, 7, 1, 0, False), (1, 2, 32, 24, 40, 1, 8, 0, 0, False), (1, 2, 32, 23, 33, 1, 7, 0, 3, False)]
f32out = "f32out" in sys.argv
if "quick" in sys.argv:
    CASES = CASES[:2]
for case in CASES:
    B, H, D, nx, ny, g, w, exact, mode, rpe = case
    t = make_inputs(B, H, D, nx, ny, g, w, rpe, seed=301)
    for dtype in (torch.bfloat16,):
        ref = oracle_run(t, nx, ny, w, exact, mode, D ** -0.5, dtype)
        try:
            out, ff, fb = kernel_run(t, nx, ny, w, exact, mode, D ** -0.5, dtype, "auto", layout="linear", f32out=f32out)
        except Exception as e:  # noqa: BLE001
            print(case, "ERROR", repr(e)[:300], flush=True)
            continue
        names = ["o", "lse", "dq", "dk", "dv"] + (["og", "lse_g", "dqg"] if g else [])
        print(case, ff, fb, {n: "%.2e" % relerr(out[n], ref[n]) for n in names}, flush=True)
