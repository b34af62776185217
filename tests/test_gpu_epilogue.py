"""GPU: the residual / LayerNorm / bias epilogue kernels (vil_addnorm_*, vil_bias_act_*) against fp64
restatements of the element-wise chain of AttnBlock / MlpBlock (src/models/msvit.py:313-316, 337-339), through the C ABI,
and the harness with `fused_residual=True` against the same network with the stock PyTorch composition."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import relerr
from vision_longformer_b200 import B200LayerNorm, _lib, epilogue
from vision_longformer_b200.msvit import Mlp, build_vil

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("C", [48, 96, 192, 384, 768, 1024, 100])
@pytest.mark.parametrize("mode", ["fp32", "bf16", "fp16", "bf16_nobias_noscale"])
def test_add_norm_matches_fp64(C, mode):
    torch.manual_seed(C)
    B, N = 3, 517
    low = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16, "bf16_nobias_noscale": torch.bfloat16}[mode]
    full = mode != "bf16_nobias_noscale"
    ln = B200LayerNorm(C, eps=1e-6).to(DEV)
    with torch.no_grad():
        ln.weight.normal_(1.0, 0.3)
        ln.bias.normal_(0.0, 0.3)
    x = (torch.randn(B, N, C, device=DEV) * 2 + 0.5).requires_grad_(True)
    br = torch.randn(B, N, C, device=DEV).to(low).requires_grad_(True)
    bias = (0.3 * torch.randn(C, device=DEV)).requires_grad_(True) if full else None
    scale = torch.tensor([0.0, 1.0 / 0.9, 1.0 / 0.9], device=DEV) if full else None
    g_xo = torch.randn(B, N, C, device=DEV)
    g_y = torch.randn(B, N, C, device=DEV).to(low)
    assert epilogue.addnorm_applies(x, br, C)
    n0 = _lib.launch_count()
    xo, y = epilogue.add_norm(x, br, bias, scale, ln, out_dtype=low)
    assert _lib.launch_count() == n0 + 1                      # ONE kernel: add + scale + bias + norm
    assert xo.dtype == torch.float32 and y.dtype == low
    n0 = _lib.launch_count()
    ((xo * g_xo).sum() + (y.float() * g_y.float()).sum()).backward()
    assert _lib.launch_count() == n0 + 2                      # backward + the column-sum reduce
    # fp64 restatement on the same (rounded) inputs
    xr, brr = x.detach().double().requires_grad_(True), br.detach().double().requires_grad_(True)
    w, b = ln.weight.detach().double().requires_grad_(True), ln.bias.detach().double().requires_grad_(True)
    biasr = bias.detach().double().requires_grad_(True) if full else None
    t = brr + biasr if full else brr
    if full:
        t = t * scale.double().view(B, 1, 1)
    xo_r = xr + t
    y_r = F.layer_norm(xo_r, (C,), w, b, 1e-6)
    ((xo_r * g_xo.double()).sum() + (y_r * g_y.double()).sum()).backward()
    lo = mode != "fp32"
    assert relerr(xo, xo_r) < 1e-6
    assert relerr(y, y_r) < (4e-3 if low == torch.bfloat16 else 6e-4 if lo else 1e-6)
    assert relerr(x.grad, xr.grad) < 2e-6
    assert relerr(br.grad, brr.grad) < (4e-3 if low == torch.bfloat16 else 6e-4 if lo else 2e-6)
    assert relerr(ln.weight.grad, w.grad) < 1e-5
    assert relerr(ln.bias.grad, b.grad) < 1e-5
    if full:
        assert relerr(bias.grad, biasr.grad) < 1e-5
        assert torch.all(br.grad[0] == 0)                     # the dropped sample gets no branch gradient


def test_add_norm_without_residual_gradient_and_empty():
    """The final norm of the network: xo is not used downstream (its gradient is None)."""
    C = 192
    ln = B200LayerNorm(C, eps=1e-6).to(DEV)
    x = torch.randn(2, 50, C, device=DEV, requires_grad=True)
    br = torch.randn(2, 50, C, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    gy = torch.randn(2, 50, C, device=DEV)
    _, y = epilogue.add_norm(x, br, None, None, ln, out_dtype=torch.bfloat16)
    (y.float() * gy.to(torch.bfloat16).float()).sum().backward()
    xr, brr = x.detach().double().requires_grad_(True), br.detach().double().requires_grad_(True)
    y_r = F.layer_norm(xr + brr, (C,), ln.weight.detach().double(), ln.bias.detach().double(), 1e-6)
    (y_r * gy.to(torch.bfloat16).double()).sum().backward()
    assert relerr(x.grad, xr.grad) < 1e-5 and relerr(br.grad, brr.grad) < 4e-3
    # empty stream: nothing is launched, parameter gradients are zeros
    x0 = torch.zeros(0, 7, C, device=DEV, requires_grad=True)
    br0 = torch.zeros(0, 7, C, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    xo0, y0 = epilogue.add_norm(x0, br0, None, None, ln, out_dtype=torch.bfloat16)
    ln.weight.grad = None
    (xo0.sum() + y0.float().sum()).backward()
    assert xo0.shape == x0.shape and torch.all(ln.weight.grad == 0)


@pytest.mark.parametrize("C", [384, 768, 1536, 3072, 96, 200])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_bias_gelu_matches_fp64(C, dtype):
    torch.manual_seed(C)
    rows = (2, 1031)
    z = (2 * torch.randn(*rows, C, device=DEV)).to(dtype).requires_grad_(True)
    bias = (0.5 * torch.randn(C, device=DEV)).requires_grad_(True)
    da = torch.randn(*rows, C, device=DEV).to(dtype)
    assert epilogue.bias_act_applies(z)
    n0 = _lib.launch_count()
    a = epilogue.bias_gelu(z, bias)
    assert _lib.launch_count() == n0 + 1 and a.dtype == dtype
    (a.float() * da.float()).sum().backward()
    zr, br = z.detach().double().requires_grad_(True), bias.detach().double().requires_grad_(True)
    ar = F.gelu(zr + br)
    (ar * da.double()).sum().backward()
    tol = {torch.bfloat16: 4e-3, torch.float16: 6e-4, torch.float32: 2e-6}[dtype]
    assert relerr(a, ar) < tol
    assert relerr(z.grad, zr.grad) < tol
    assert relerr(bias.grad, br.grad) < max(tol, 1e-5)


@pytest.mark.parametrize("cin,cout", [(96, 96), (96, 192), (384, 1152), (192, 384), (64, 200)])
def test_linear_colsum_bias_matches_linear(cin, cout):
    torch.manual_seed(cout)
    lin = torch.nn.Linear(cin, cout).to(DEV)
    x = torch.randn(3, 411, cin, device=DEV, requires_grad=True)
    gy = torch.randn(3, 411, cout, device=DEV)
    for autocast in (False, True):
        for p in (*lin.parameters(), x):
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            n0 = _lib.launch_count()
            y = epilogue.linear_colsum_bias(x, lin.weight, lin.bias)
            y_ref = lin(x)
        assert y.dtype == y_ref.dtype and torch.equal(y, y_ref)
        (y.float() * gy).sum().backward()
        assert _lib.launch_count() == n0 + 2                  # column sums + their reduce
        got = [t.grad.clone() for t in (x, lin.weight, lin.bias)]
        for p in (*lin.parameters(), x):
            p.grad = None
        (y_ref.float() * gy).sum().backward()
        tol = 1e-2 if autocast else 1e-5
        for a, b in zip(got, (x.grad, lin.weight.grad, lin.bias.grad)):
            assert a.dtype == b.dtype and relerr(a, b) < tol
        # the bias gradient against the exact column sum of what the GEMMs saw
        exact = gy.to(y.dtype).double().sum(dim=(0, 1))
        assert relerr(got[2], exact) < (4e-3 if autocast else 1e-6)


def test_mlp_deferred_matches_stock():
    torch.manual_seed(3)
    mlp = Mlp(192, 768).to(DEV)
    x = torch.randn(2, 300, 192, device=DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ref = mlp(x)
        out, bias = mlp.forward_deferred(x.to(torch.bfloat16))
    assert bias is mlp.fc2.bias
    assert relerr(out.float() + bias, ref) < 1e-2


@pytest.mark.parametrize("arch,img", [("vil_tiny", 224),
                                      ("l1,h2,d64,n2,s1,g1,p4,f7_l2,h2,d128,n2,s0,g1,p2,f7_l3,h4,d256,n1,s0,g0,p2,f7", 112)])
@pytest.mark.parametrize("train", [False, True])
def test_harness_fused_residual_matches_stock_composition(arch, img, train):
    """Same weights, same DropPath draws: `fused_residual=True` (add + DropPath scale + deferred bias + LayerNorm in one kernel per
    block boundary, bias + GELU, column-sum bias gradients) against the stock PyTorch composition with the same attention
    kernels, under bf16 autocast: logits, input gradient and EVERY parameter gradient."""
    torch.manual_seed(0)
    kw = dict(img_size=img, num_classes=50, drop_path_rate=0.2 if train else 0.0)
    a = build_vil(arch, fused_residual=True, **kw).to(DEV)
    b = build_vil(arch, fused_residual=False, **kw).to(DEV)
    b.load_state_dict(a.state_dict())
    assert set(a.state_dict().keys()) == set(b.state_dict().keys())
    a.train(train), b.train(train)
    x = torch.randn(4, 3, img, img, device=DEV)
    gy = torch.randn(4, 50, device=DEV)
    outs = []
    for net in (a, b):
        xg = x.clone().requires_grad_(True)
        torch.manual_seed(123)                                # same per-sample DropPath masks in both nets (same draw order)
        torch.cuda.manual_seed(123)
        n0 = _lib.launch_count()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = net(xg)
        (y.float() * gy).sum().backward()
        outs.append((y.float(), xg.grad, {k: p.grad for k, p in net.named_parameters()}, _lib.launch_count() - n0))
    (ya, dxa, ga, la), (yb, dxb, gb, lb) = outs
    assert la > lb                                            # the epilogue kernels really ran
    assert relerr(ya, yb) < 3e-2 and relerr(dxa, dxb) < 6e-2, (relerr(ya, yb), relerr(dxa, dxb))
    worst = max((relerr(ga[k], gb[k]), k) for k in ga if gb[k] is not None and gb[k].numel() > 0 and gb[k].abs().max() > 0)
    assert all((ga[k] is None) == (gb[k] is None) for k in ga)
    assert worst[0] < 8e-2, worst
