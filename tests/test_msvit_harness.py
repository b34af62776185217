"""CPU: the stock-PyTorch MsViT harness (vision_longformer_b200/msvit.py) against golden vectors from the
reference MsViT, with the ORACLE attention module plugged in (the B200 module has no CPU path)."""
import pytest
import torch

from oracle.vil_oracle import OracleLong2DSCSelfAttention
from tests.util import load_golden, load_state, relerr
from vision_longformer_b200 import ARCHS, MsViT, build_vil, parse_arch


@pytest.mark.parametrize("name", ["tiny_rpe", "tiny_ape"])
def test_harness_matches_reference_msvit(name):
    gold = load_golden(f"msvit_{name}.pt")
    net = MsViT(attn_cls=OracleLong2DSCSelfAttention, **gold["kwargs"]).double().eval()
    assert sum(p.numel() for p in net.parameters()) == gold["n_params"]
    assert set(net.state_dict().keys()) == set(gold["state_dict"].keys())
    load_state(net, gold["state_dict"])
    x = gold["x"].double().requires_grad_(True)
    y = net(x)
    assert relerr(y, gold["y"]) < 1e-10
    (y * gold["gy"]).sum().backward()
    assert relerr(x.grad, gold["dx"]) < 1e-9
    grads = dict(net.named_parameters())
    for n, gref in gold["param_grads"].items():
        assert relerr(grads[n].grad, gref) < 1e-5, n


def test_published_archs_parameter_counts():
    # README.md:77-95 of the reference: 6.7 / 24.6 / 39.7 / 55.7 M parameters
    want = {"vil_tiny": 6.71e6, "vil_small": 24.64e6, "vil_medium_deep": 39.74e6, "vil_base_deep": 55.72e6}
    for name, n in want.items():
        img = 384 if name == "vil_base_deep" else 224
        net = build_vil(name, img_size=img, attn_cls=OracleLong2DSCSelfAttention)
        got = sum(p.numel() for p in net.parameters())
        assert abs(got - n) / n < 2e-3, (name, got)


def test_arch_parser_defaults_and_rpe_switch():
    cfg = parse_arch(ARCHS["vil_small"])
    assert [c["f"] for c in cfg] == [7, 7, 7, 7] and [c["s"] for c in cfg] == [1, 1, 0, 0]
    assert all(c["a"] == 1 for c in cfg)       # published strings omit `a` -> absolute pos-embed, rpe off
    net = build_vil("l1,h2,d16,n1,s1,g1,p4,f4,a0_l2,h2,d32,n1,s0,g1,p2,f4,a0_l3,h2,d32,n1,s1,g0,p2,f4,a0",
                    img_size=32, attn_cls=OracleLong2DSCSelfAttention)
    assert net.layer1[1].attn.rpe and type(net.layer1[1].attn).__name__ == "OracleLong2DSCSelfAttention"
    # sticky 'full' quirk: stage 3 is s1 but comes after an s0 stage
    assert type(net.layer3[1].attn).__name__ == "DenseAttention"


def test_pending_branch_plumbing_matches_the_inline_composition():
    """The fused residual path of the harness hands a block's branch to the next block as (output without bias, bias, DropPath
    scale); on CPU it is joined by `_flush` with stock ops.  x + drop_path(branch + bias) must come out either way, and
    DropPath.forward must use the very draw `sample_scale` hands to the residual kernel."""
    import torch
    from vision_longformer_b200.msvit import DropPath, _flush, _unpack
    torch.manual_seed(0)
    x, br, bias = torch.randn(3, 5, 8), torch.randn(3, 5, 8), torch.randn(8)
    dp = DropPath(0.5).train()
    torch.manual_seed(7)
    ref = x + dp(br + bias)
    torch.manual_seed(7)
    scale = dp.sample_scale(3, x.device)
    assert scale.shape == (3,) and set(scale.tolist()) <= {0.0, 2.0}
    assert torch.allclose(_flush(x, (br, bias, scale)), ref)
    assert torch.equal(_flush(x, None), x) and torch.allclose(_flush(x, (br, None, None)), x + br)
    assert dp.eval().sample_scale(3, x.device) is None and DropPath(0.0).train().sample_scale(3, x.device) is None
    assert _unpack((x, 2, 3)) == (x, 2, 3, None) and _unpack((x, 2, 3, "p"))[3] == "p"


def test_epilogue_entry_points_keep_the_stock_ops_on_cpu():
    """The epilogue kernels are CUDA ops; on CPU tensors the callers must take the stock PyTorch composition (the harness runs on
    CPU with the oracle attention for the reference arm) - no silent half-fused path."""
    import torch
    import torch.nn.functional as F
    from vision_longformer_b200 import epilogue
    from vision_longformer_b200.msvit import Mlp
    x, br = torch.randn(2, 5, 8), torch.randn(2, 5, 8)
    assert not epilogue.addnorm_applies(x, br, 8) and not epilogue.bias_act_applies(x)
    lin = torch.nn.Linear(8, 16)
    assert torch.equal(epilogue.linear_colsum_bias(x, lin.weight, lin.bias), F.linear(x, lin.weight, lin.bias))
    mlp = Mlp(8, 32)
    out, bias = mlp.forward_deferred(x)
    assert bias is None and torch.equal(out, mlp(x))          # nothing deferred: the complete stock result
