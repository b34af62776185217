"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI
(libvil_attn_sm100.so via ctypes) and is compared with
  * the golden vectors generated from the unmodified reference (tests/golden/attn_*.pt), and
  * the CPU oracle (oracle/vil_oracle.py) on seeded inputs,
plus size-independent properties at the BASELINE shapes.

Tolerances (norm-relative, ||x - ref||_F / ||ref||_F, documented in DESIGN.md):
  fp32 I/O : 1e-5   (BASELINE north_star)
  fp16 I/O : 1e-3   (north_star)
  bf16 I/O : 4e-3 forward / 8e-3 backward -- the bf16 OUTPUT rounding alone is 1.65e-3
             and the reference module itself sits at 3.3e-3 / 6.8e-3 in bf16; the kernel-internal error is
             isolated by the fp16 and fp32 runs.
"""
import pytest
import torch

from oracle import vil_oracle as vo
from tests.util import attn_cases, load_attn, load_golden, load_state, relerr
from vision_longformer_b200 import (B200Long2DSCSelfAttention, MsViT, _lib, vil_attention,
                                    vil_attention_raw_backward, vil_attention_raw_forward)

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {torch.float32: (1e-5, 2e-5), torch.float16: (1e-3, 2e-3), torch.bfloat16: (4e-3, 8e-3)}


# --------------------------------------------------------------------------- golden module parity
@pytest.mark.parametrize("name", attn_cases())
@pytest.mark.parametrize("impl", ["simt", "auto"])
def test_module_matches_reference_golden_fp32(name, impl):
    gold = load_attn(name)
    mod = B200Long2DSCSelfAttention(**gold["kwargs"]).to(DEV)
    load_state(mod, gold["state_dict"])
    mod.impl = impl
    pick = gold["picked_mode"]
    mod.train(pick is not None)
    if pick is not None:
        mod._pick_mode = lambda: pick
    x = gold["x"].to(DEV).requires_grad_(True)
    y = mod(x, gold["nx"], gold["ny"])
    (y * gold["gy"].to(DEV)).sum().backward()
    assert relerr(y, gold["y"]) < 1e-5
    assert relerr(x.grad, gold["dx"]) < 2e-5
    grads = {n: p.grad for n, p in mod.named_parameters()}
    for n, gref in gold["param_grads"].items():
        assert grads[n] is not None, n
        assert relerr(grads[n], gref) < 5e-5, n


@pytest.mark.parametrize("name", ["w7_g1_exact0_rpe", "w7_g1_exact1_rpe", "w8_g1_exact0_d32", "w7_g1_exact0_d64_28",
                                  "w4_g2_exact0_norpe_nosharew"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_module_matches_reference_golden_lowp(name, dtype):
    gold = load_attn(name)
    mod = B200Long2DSCSelfAttention(**gold["kwargs"]).to(DEV)
    load_state(mod, gold["state_dict"])
    mod = mod.to(dtype).eval()
    x = gold["x"].to(DEV, dtype).requires_grad_(True)
    y = mod(x, gold["nx"], gold["ny"])
    (y * gold["gy"].to(DEV, dtype)).sum().backward()
    # the Linears run in low precision here as well, so this is a loose end-to-end check
    assert relerr(y, gold["y"]) < 3e-2
    assert relerr(x.grad, gold["dx"]) < 6e-2


# --------------------------------------------------------------------------- op-level parity vs oracle
def make_inputs(B, H, D, nx, ny, g, w, rpe, seed=300, dtype=torch.float64):
    gen = torch.Generator().manual_seed(seed)       # the reference tests' seed
    N = g + nx * ny
    r = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float64)
    t = dict(q=r(B, H, nx * ny, D), k=r(B, H, N, D), v=r(B, H, N, D), qg=r(B, H, max(g, 1), D)[:, :, :g],
             table=0.5 * r((4 * w - 1) ** 2, H) if rpe else None,
             g2l=0.5 * r(2, H, g) if (rpe and g) else None, g2g=0.5 * r(H, g, g) if (rpe and g) else None,
             go=r(B, H, nx * ny, D), gog=r(B, H, max(g, 1), D)[:, :, :g])
    return t


_ORACLE_CACHE = {}


def oracle_run(t, nx, ny, w, exact, mode, scale, dtype, key=None):
    """fp64 oracle on the values the kernel actually sees (inputs rounded to `dtype`)."""
    if key is not None and (key, dtype) in _ORACLE_CACHE:
        return _ORACLE_CACHE[(key, dtype)]
    out = _oracle_run(t, nx, ny, w, exact, mode, scale, dtype)
    if key is not None:
        _ORACLE_CACHE[(key, dtype)] = out
    return out


def _oracle_run(t, nx, ny, w, exact, mode, scale, dtype):
    rd = lambda x: None if x is None else x.to(dtype).double().requires_grad_(True)
    q, k, v, qg = rd(t["q"]), rd(t["k"]), rd(t["v"]), rd(t["qg"])
    table, g2l, g2g = [None if t[n] is None else t[n].float().double().requires_grad_(True) for n in ("table", "g2l", "g2g")]
    g = k.shape[2] - q.shape[2]
    o, og, lse, lse_g = vo.dense_attention(q, k, v, qg if g else None, k, v, table, g2l, g2g, nx=nx, ny=ny, w=w,
                                           exact=exact, mode=mode, scale=scale)
    go, gog = t["go"].to(dtype).double(), t["gog"].to(dtype).double()
    loss = (o * go).sum() + ((og * gog).sum() if g else 0)
    ins = [x for x in (q, k, v, qg if g else None, table, g2l, g2g) if x is not None]
    grads = torch.autograd.grad(loss, ins)
    names = [n for n, x in zip(("q", "k", "v", "qg", "table", "g2l", "g2g"), (q, k, v, qg if g else None, table, g2l, g2g)) if x is not None]
    return dict(o=o, og=og, lse=lse, lse_g=lse_g, **{"d" + n: gr for n, gr in zip(names, grads)})


def kernel_run(t, nx, ny, w, exact, mode, scale, dtype, impl):
    dev = lambda x: None if x is None else x.to(DEV, dtype).contiguous()
    f32 = lambda x: None if x is None else x.to(DEV, torch.float32).contiguous()
    q, k, v, qg = dev(t["q"]), dev(t["k"]), dev(t["v"]), dev(t["qg"])
    g = k.shape[2] - q.shape[2]
    table, g2l, g2g = f32(t["table"]), f32(t["g2l"]), f32(t["g2g"])
    o, og = torch.empty_like(q), (torch.empty_like(qg) if g else None)
    kw = dict(nx=nx, ny=ny, w=w, exact=exact, mode=mode, scale=scale, impl=impl)
    lse, lse_g = vil_attention_raw_forward(q, k, v, qg if g else None, k if g else None, v if g else None, table, g2l,
                                           g2g, o, og, **kw)
    fam_f = _lib.last_impl()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    dqg = torch.empty_like(qg) if g else None
    zl = lambda x: None if x is None else torch.zeros_like(x)
    dt, dgl, dgg = zl(table), zl(g2l), zl(g2g)
    vil_attention_raw_backward(q, k, v, qg if g else None, k if g else None, v if g else None, table, g2l, g2g, o, og,
                               lse, lse_g, dev(t["go"]), dev(t["gog"]) if g else None, dq, dk, dv, dqg,
                               dk if g else None, dv if g else None, dt, dgl, dgg, **kw)
    torch.cuda.synchronize()
    out = dict(o=o, og=og, lse=lse, lse_g=lse_g, dq=dq, dk=dk, dv=dv, dqg=dqg, dtable=dt, dg2l=dgl, dg2g=dgg)
    return out, fam_f, _lib.last_impl()


OP_CASES = [
    # B, H, D, nx, ny, g, w, exact, mode, rpe
    (2, 3, 32, 14, 14, 1, 7, 0, 0, True),
    (2, 3, 32, 14, 14, 1, 7, 0, 0, False),
    (1, 2, 64, 21, 14, 1, 7, 0, 0, True),
    (2, 2, 32, 16, 24, 1, 8, 0, 0, False),
    (1, 3, 32, 19, 17, 1, 7, 0, 0, True),      # padding in both directions
    (1, 2, 32, 14, 14, 1, 7, 1, 0, True),      # exact window
    (1, 2, 64, 15, 13, 2, 7, 1, 0, False),
    (1, 2, 32, 15, 13, 2, 7, 0, 3, True),      # random-shift modes
    (1, 2, 32, 15, 13, 1, 7, 0, 8, False),
    (1, 2, 32, 15, 13, 1, 7, 0, -1, True),
    (1, 2, 16, 10, 9, 3, 4, -1, 0, True),      # cyclic chunks + padding quirk
    (1, 2, 16, 8, 5, 1, 4, -1, 0, False),      # mx, my <= 2: chunks visited twice
    (1, 1, 48, 14, 14, 1, 7, 0, 0, True),      # ViL-Tiny stage-1 head dim
    (1, 2, 32, 24, 24, 1, 12, 0, 0, True),     # w^2 > 64: multi-piece chunks
    (1, 2, 32, 30, 17, 1, 15, 1, 0, False),
    (1, 2, 32, 12, 12, 0, 6, 0, 0, True),      # no global tokens
    (1, 2, 32, 12, 12, 8, 6, 0, 0, True),      # g = 8
]


@pytest.mark.parametrize("case", OP_CASES, ids=lambda c: "B%d_H%d_D%d_%dx%d_g%d_w%d_e%d_m%d_%s" % (c[:9] + ("rpe" if c[9] else "nob",)))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("impl", ["simt", "auto"])
def test_op_matches_oracle(case, dtype, impl):
    B, H, D, nx, ny, g, w, exact, mode, rpe = case
    t = make_inputs(B, H, D, nx, ny, g, w, rpe)
    scale = D ** -0.5
    ref = oracle_run(t, nx, ny, w, exact, mode, scale, dtype, key=case)
    out, fam_f, fam_b = kernel_run(t, nx, ny, w, exact, mode, scale, dtype, impl)
    tf, tb = TOL[dtype]
    assert relerr(out["o"], ref["o"]) < tf
    assert relerr(out["lse"], ref["lse"]) < 1e-5 if dtype == torch.float32 else relerr(out["lse"], ref["lse"]) < 1e-3
    for n in ("dq", "dk", "dv"):
        assert relerr(out[n], ref[n]) < tb, n
    if g:
        assert relerr(out["og"], ref["og"]) < tf
        assert relerr(out["dqg"], ref["dqg"]) < tb
    if rpe:
        # bias gradients are sums of dS over thousands of (query, key) pairs with heavy cancellation; in low
        # precision they inherit the rounding of the STORED o (delta = dO.o uses the bf16/fp16 output, exactly as
        # the reference's autograd does), hence the looser bound there.
        tbias = {torch.float32: 1e-4, torch.float16: 1e-2, torch.bfloat16: 5e-2}[dtype]
        assert relerr(out["dtable"], ref["dtable"]) < tbias
        if g:
            assert relerr(out["dg2l"], ref["dg2l"]) < tbias
            assert relerr(out["dg2g"], ref["dg2g"]) < tbias


TC_CASES = [
    # B, H, D, nx, ny, g, w, exact, mode, rpe  -- all must be served by the tcgen05 family in the forward
    (2, 3, 32, 56, 56, 1, 7, 0, 0, False),     # ViL-Small stage 1 (rpe off, published arch)
    (2, 3, 64, 28, 28, 1, 7, 0, 0, False),     # ViL-Small stage 2
    (2, 3, 32, 28, 28, 1, 7, 0, 0, True),
    (1, 2, 64, 21, 35, 1, 7, 0, 0, True),      # odd number of chunk columns (slot B missing in the last pair)
    (1, 1, 48, 19, 17, 2, 7, 0, 0, True),      # D=48 (padded to 64 by TMA), padding rows/cols, 2 global tokens
    (1, 2, 32, 24, 40, 1, 8, 0, 0, True),      # w=8: full 64-row slots
    (1, 2, 64, 18, 15, 1, 6, 1, 0, True),      # w=6, exact window, padding
    (1, 2, 32, 20, 22, 1, 7, 1, 0, False),     # exact window without rpe (mask-only table)
    (1, 2, 32, 22, 20, 1, 7, 0, 5, True),      # random-shift mode
    (1, 2, 32, 22, 20, 0, 7, 0, -1, False),    # own chunk only, no global tokens
    (1, 2, 16, 15, 29, 16, 7, 0, 0, True),     # D=16 (padded to 32), 16 global tokens
]


@pytest.mark.parametrize("case", TC_CASES, ids=lambda c: "B%d_H%d_D%d_%dx%d_g%d_w%d_e%d_m%d_%s" % (c[:9] + ("rpe" if c[9] else "nob",)))
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_tcgen05_forward_matches_oracle(case, dtype):
    B, H, D, nx, ny, g, w, exact, mode, rpe = case
    t = make_inputs(B, H, D, nx, ny, g, w, rpe, seed=301)
    scale = D ** -0.5
    ref = oracle_run(t, nx, ny, w, exact, mode, scale, dtype, key=("tc",) + case)
    out, fam_f, fam_b = kernel_run(t, nx, ny, w, exact, mode, scale, dtype, "auto")
    assert fam_f == "tcgen05", fam_f            # no silent fallback
    assert fam_b == "tcgen05", fam_b            # incl. the bias-table-gradient variant of the dQ pass
    tf, tb = TOL[dtype]
    assert relerr(out["o"], ref["o"]) < tf
    assert relerr(out["lse"], ref["lse"]) < 1e-4
    for n in ("dq", "dk", "dv"):                # backward consumes the tcgen05 forward's o / lse
        assert relerr(out[n], ref[n]) < tb, n
    if g:
        assert relerr(out["og"], ref["og"]) < tf
    if rpe:
        tbias = {torch.float16: 1e-2, torch.bfloat16: 5e-2}[dtype]
        assert relerr(out["dtable"], ref["dtable"]) < tbias
        if g:
            assert relerr(out["dg2l"], ref["dg2l"]) < tbias


TC_BIG_CASES = [
    # chunk sizes > 8: tcgen05 kernels tiled by chunk pieces (vil_tc_big.cuh); with a bias table the backward is SIMT
    (1, 2, 32, 24, 24, 1, 12, 0, 0, True),     # Medium-Deep-384 stage-2 window
    (1, 2, 32, 26, 37, 2, 12, 0, 3, True),     # padding, 2 global tokens, random-shift mode
    (1, 2, 64, 30, 17, 1, 15, 1, 0, False),    # exact window (mask-only table), short last piece
    (1, 1, 48, 15, 15, 1, 15, 0, 0, True),     # single chunk
    (1, 1, 32, 62, 40, 1, 31, 0, 0, False),    # w = 31: 16 pieces per chunk, padding
    (2, 2, 32, 36, 25, 1, 12, 0, 0, False),    # 3 x 3 chunks, padding, tcgen05 backward
    (1, 2, 64, 31, 45, 2, 15, 0, 6, False),    # random-shift mode, D = 64, tcgen05 backward
]


@pytest.mark.parametrize("case", TC_BIG_CASES, ids=lambda c: "B%d_H%d_D%d_%dx%d_g%d_w%d_e%d_m%d_%s" % (c[:9] + ("rpe" if c[9] else "nob",)))
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_tcgen05_big_window_forward_matches_oracle(case, dtype):
    B, H, D, nx, ny, g, w, exact, mode, rpe = case
    t = make_inputs(B, H, D, nx, ny, g, w, rpe, seed=302)
    scale = D ** -0.5
    ref = oracle_run(t, nx, ny, w, exact, mode, scale, dtype, key=("tcbig",) + case)
    out, fam_f, fam_b = kernel_run(t, nx, ny, w, exact, mode, scale, dtype, "auto")
    assert fam_f == "tcgen05" and fam_b == ("simt" if rpe else "tcgen05"), (fam_f, fam_b)
    tf, tb = TOL[dtype]
    assert relerr(out["o"], ref["o"]) < tf
    assert relerr(out["lse"], ref["lse"]) < 1e-4
    for n in ("dq", "dk", "dv"):
        assert relerr(out[n], ref[n]) < tb, n


def test_autograd_function_on_strided_linear_outputs():
    """q / kv consumed in place from the Linear layouts, output produced head-merged; separate global weights."""
    torch.manual_seed(5)
    B, H, D, nx, ny, g, w = 2, 3, 32, 15, 14, 2, 7
    C, N = H * D, g + nx * ny
    mk = lambda *s: torch.randn(*s, dtype=torch.float64)
    q_all, qg_all, kv, kvg = mk(B, nx * ny, C), mk(B, g, C), mk(B, N, 2 * C), mk(B, N, 2 * C)
    table, g2l, g2g = 0.3 * mk((4 * w - 1) ** 2, H), 0.3 * mk(2, H, g), 0.3 * mk(H, g, g)
    gy = mk(B, N, C)
    ins64 = [x.clone().requires_grad_(True) for x in (q_all, kv, qg_all, kvg, table, g2l, g2g)]
    hd = lambda t, i=0, p=1: t.view(B, t.shape[1], p, H, D)[:, :, i].permute(0, 2, 1, 3)
    o, og, _, _ = vo.dense_attention(hd(ins64[0]), hd(ins64[1], 0, 2), hd(ins64[1], 1, 2), hd(ins64[2]),
                                     hd(ins64[3], 0, 2), hd(ins64[3], 1, 2), ins64[4], ins64[5], ins64[6],
                                     nx=nx, ny=ny, w=w, exact=0, mode=0, scale=D ** -0.5)
    y_ref = torch.cat([og.transpose(1, 2).reshape(B, g, C), o.transpose(1, 2).reshape(B, nx * ny, C)], dim=1)
    g_ref = torch.autograd.grad((y_ref * gy).sum(), ins64)
    ins = [x.to(DEV, torch.float32).requires_grad_(True) for x in (q_all, kv, qg_all, kvg, table, g2l, g2g)]
    y = vil_attention(ins[0], ins[1], ins[2], ins[3], ins[4], ins[5], ins[6], num_heads=H, nx=nx, ny=ny, w=w, nglo=g,
                      exact=0, mode=0, scale=D ** -0.5)
    grads = torch.autograd.grad((y * gy.to(DEV, torch.float32)).sum(), ins)
    assert relerr(y, y_ref) < 1e-5
    for a, b in zip(grads, g_ref):
        assert relerr(a, b) < 3e-5


# --------------------------------------------------------------------------- error behaviour
def test_errors_mirror_reference():
    q = torch.randn(1, 2, 49, 32, device=DEV)
    k = torch.randn(1, 2, 50, 32, device=DEV)
    o, og = torch.empty_like(q), torch.empty(1, 2, 1, 32, device=DEV)
    with pytest.raises(ValueError, match="exact"):
        vil_attention_raw_forward(q, k, k, k[:, :, :1], k, k, None, None, None, o, og, nx=7, ny=7, w=7, exact=2)
    with pytest.raises(ValueError):
        vil_attention_raw_forward(q, k, k, k[:, :, :1], k, k, None, None, None, o, og, nx=7, ny=7, w=7, exact=1, mode=2)
    with pytest.raises(AssertionError, match="Global dimension"):
        vil_attention_raw_forward(q, k, k, k[:, :, :1], k, k, None, None, None, o, og, nx=6, ny=7, w=7)
    with pytest.raises(RuntimeError, match="no CPU"):
        vil_attention_raw_forward(q.cpu(), k.cpu(), k.cpu(), None, None, None, None, None, None, o.cpu(), None,
                                  nx=7, ny=7, w=7)


# --------------------------------------------------------------------------- properties at BASELINE shapes
@pytest.mark.parametrize("shape", [(8, 3, 32, 56, 56), (8, 3, 64, 28, 28)], ids=["S1", "S2"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
def test_full_size_properties(shape, dtype):
    """ViL-Small stage-1 / stage-2 shapes (BASELINE config 2; batch reduced to 8, the kernel is batch-independent):
    (1) rows of P sum to one: V = 1 -> O = 1;  (2) linearity in V;  (3) batch independence;
    (4) dV column sums equal dO column sums (sum_j dV_j = sum_i dO_i because P rows sum to 1);
    (5) the two kernel families agree where both apply."""
    B, H, D, nx, ny = shape
    w, g = 7, 1
    N = g + nx * ny
    gen = torch.Generator(device=DEV).manual_seed(300)
    r = lambda *s: torch.randn(*s, generator=gen, device=DEV, dtype=torch.float32).to(dtype)
    q, k, v1, v2 = r(B, H, nx * ny, D), r(B, H, N, D), r(B, H, N, D), r(B, H, N, D)
    qg = r(B, H, g, D)
    kw = dict(nx=nx, ny=ny, w=w, exact=0, mode=0, scale=D ** -0.5)

    def fwd(qq, kk, vv, qgg, impl="auto"):
        o, og = torch.empty_like(qq), torch.empty_like(qgg)
        lse, lse_g = vil_attention_raw_forward(qq, kk, vv, qgg, kk, vv, None, None, None, o, og, impl=impl, **kw)
        return o, og, lse, lse_g

    ones = torch.ones_like(v1)
    o1, og1, _, _ = fwd(q, k, ones, qg)
    tol = 1e-5 if dtype == torch.float32 else 8e-3
    assert (o1.float() - 1).abs().max() < tol and (og1.float() - 1).abs().max() < tol
    oa, _, lse_a, _ = fwd(q, k, v1, qg)
    ob, _, _, _ = fwd(q, k, v2, qg)
    oc, _, _, _ = fwd(q, k, (v1.float() + v2.float()).to(dtype), qg)
    assert relerr(oc, oa.float() + ob.float()) < (2e-5 if dtype == torch.float32 else 1e-2)
    o_half, _, lse_h, _ = fwd(q[:2], k[:2], v1[:2], qg[:2])
    assert torch.equal(o_half, oa[:2]) and torch.equal(lse_h, lse_a[:2])
    o_s, _, lse_s, _ = fwd(q, k, v1, qg, impl="simt")
    assert relerr(oa, o_s) < (1e-5 if dtype == torch.float32 else 6e-3)
    assert relerr(lse_a, lse_s) < 1e-3

    go, gog = r(B, H, nx * ny, D), r(B, H, g, D)
    o, og, lse, lse_g = fwd(q, k, v1, qg)
    dq, dk, dv, dqg = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v1), torch.empty_like(qg)
    vil_attention_raw_backward(q, k, v1, qg, k, v1, None, None, None, o, og, lse, lse_g, go, gog, dq, dk, dv, dqg,
                               dk, dv, None, None, None, **kw)
    lhs = dv.float().sum(dim=2)
    rhs = go.float().sum(dim=2) + gog.float().sum(dim=2)
    assert relerr(lhs, rhs) < (2e-5 if dtype == torch.float32 else 2e-2)
    # dK columns: sum_j dk_j . anything is not conserved, but sum over keys of dS is zero per query row:
    # check via dq . q  +  ... skipped; gradient parity is covered by test_op_matches_oracle.


# --------------------------------------------------------------------------- MsViT end to end
@pytest.mark.parametrize("name", ["tiny_rpe", "tiny_ape"])
def test_msvit_end_to_end_fp32(name):
    gold = load_golden(f"msvit_{name}.pt")
    net = MsViT(**gold["kwargs"]).to(DEV).eval()
    load_state(net, gold["state_dict"])
    x = gold["x"].to(DEV).requires_grad_(True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        y = net(x)
        (y * gold["gy"].to(DEV).float()).sum().backward()
    assert relerr(y, gold["y"]) < 2e-5
    assert relerr(x.grad, gold["dx"]) < 1e-4


def test_gpu_launch_counter_and_family():
    before = _lib.launch_count()
    t = make_inputs(1, 2, 32, 14, 14, 1, 7, False)
    kernel_run(t, 14, 14, 7, 0, 0, 32 ** -0.5, torch.bfloat16, "auto")
    assert _lib.launch_count() - before >= 6
    assert _lib.last_impl() in ("simt", "tcgen05")
