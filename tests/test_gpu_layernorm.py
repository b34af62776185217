"""GPU: the fused LayerNorm kernels against torch's fp32 layer_norm, through the C ABI."""
import pytest
import torch

from tests.util import relerr
from vision_longformer_b200 import B200LayerNorm

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("C", [48, 96, 192, 384, 768, 1024, 100])
@pytest.mark.parametrize("mode", ["fp32", "bf16", "autocast"])
def test_layernorm_matches_torch(C, mode):
    torch.manual_seed(C)
    rows = (3, 517)
    ln = B200LayerNorm(C, eps=1e-6).to(DEV)
    with torch.no_grad():
        ln.weight.normal_(1.0, 0.3)
        ln.bias.normal_(0.0, 0.3)
    x64 = torch.randn(*rows, C, dtype=torch.float64) * 2 + 0.5
    gy64 = torch.randn(*rows, C, dtype=torch.float64)
    xdt = torch.bfloat16 if mode == "bf16" else torch.float32
    x = x64.to(DEV, xdt).requires_grad_(True)
    if mode == "bf16":
        ln = ln.to(torch.bfloat16)
    xr = x.detach().double().cpu().requires_grad_(True)
    w, b = ln.weight.detach().double().cpu().requires_grad_(True), ln.bias.detach().double().cpu().requires_grad_(True)
    y_ref = torch.nn.functional.layer_norm(xr, (C,), w, b, 1e-6)
    if mode == "autocast":
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = ln(x)
        assert y.dtype == torch.bfloat16
    else:
        y = ln(x)
        assert y.dtype == xdt
    gy = gy64.to(DEV, y.dtype)
    (y * gy).sum().backward()
    (y_ref * gy.double().cpu()).sum().backward()
    tol = 1e-6 if mode == "fp32" else 4e-3
    assert relerr(y, y_ref) < tol
    assert relerr(x.grad, xr.grad) < (2e-6 if mode == "fp32" else 6e-3)
    assert relerr(ln.weight.grad, w.grad) < (1e-5 if mode == "fp32" else 1e-2)
    assert relerr(ln.bias.grad, b.grad) < (1e-5 if mode == "fp32" else 1e-2)


def test_layernorm_state_dict_is_nn_layernorm():
    a, b = B200LayerNorm(96, eps=1e-6), torch.nn.LayerNorm(96, eps=1e-6)
    assert set(a.state_dict().keys()) == set(b.state_dict().keys())
    x = torch.randn(4, 96)
    assert torch.equal(a(x), b(x))          # CPU tensors fall through to nn.LayerNorm


@pytest.mark.parametrize("C", [96, 192, 768])
def test_patch_embed_norm_widens_to_fp32_under_autocast(C):
    """keep_dtype norm (the patch-embedding norm): a bf16 Conv2d output under autocast comes out as fp32, like
    nn.LayerNorm under autocast, through the low-precision-in -> fp32-out kernel variant (one pass)."""
    torch.manual_seed(C)
    ln = B200LayerNorm(C, eps=1e-6, keep_dtype=True).to(DEV)
    ref = torch.nn.LayerNorm(C, eps=1e-6).to(DEV)
    with torch.no_grad():
        ln.weight.normal_(1.0, 0.3); ln.bias.normal_(0.0, 0.3)
    ref.load_state_dict(ln.state_dict())
    x = (torch.randn(5, 331, C, device=DEV) * 2).bfloat16()
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    gy = torch.randn(5, 331, C, device=DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ya, yb = ln(xa), ref(xb)
    assert ya.dtype == torch.float32 and yb.dtype == torch.float32
    (ya * gy).sum().backward()
    (yb * gy).sum().backward()
    assert relerr(ya, yb) < 1e-6
    assert xa.grad.dtype == torch.bfloat16 and relerr(xa.grad, xb.grad) < 6e-3
    assert relerr(ln.weight.grad, ref.weight.grad) < 1e-4 and relerr(ln.bias.grad, ref.bias.grad) < 1e-4
    y32 = ln(x.float())                      # fp32 input stays fp32 (no autocast)
    assert y32.dtype == torch.float32
