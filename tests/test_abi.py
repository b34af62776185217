"""CPU: the C-ABI shared library loads, exports every symbol include/vil_attn.h declares, and its
host-side validation mirrors the reference's error behaviour.  No kernel is launched here."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as ge
from vision_longformer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    ge.build()
    return _lib.load()


def test_header_and_binding_list_the_same_symbols():
    hdr = open(os.path.join(ROOT, "include", "vil_attn.h")).read()
    declared = set(re.findall(r"\b(vil_(?:attn|layernorm|addnorm|bias_act)_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for sym in _lib.EXPORTS:
        assert hasattr(raw, sym), sym
    assert lib.vil_attn_abi_version() == _lib.ABI_VERSION
    assert isinstance(_lib.launch_count(), int)


def _params(**kw):
    p = _lib.VilAttnParams()
    p.struct_bytes = ctypes.sizeof(_lib.VilAttnParams)
    p.dtype, p.impl = _lib.VIL_BF16, _lib.VIL_IMPL_AUTO
    p.B, p.H, p.D, p.nx, p.ny, p.w, p.nglo, p.exact, p.mode = 2, 3, 32, 56, 56, 7, 1, 0, 0
    p.scale = 32 ** -0.5
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def test_workspace_query_and_validation(lib):
    p = _params()
    fwd = lib.vil_attn_workspace_bytes(ctypes.byref(p), 0)
    bwd = lib.vil_attn_workspace_bytes(ctypes.byref(p), 1)
    assert fwd >= 0 and bwd >= 2 * 3 * 56 * 56 * 4          # backward: at least one fp32 per (b, h, query) row
    # mask_invalid_locations: ValueError("longsc exact should be in [0,1,-1]!")  (slidingchunk_2d.py:343)
    bad = _params(exact=2)
    assert lib.vil_attn_workspace_bytes(ctypes.byref(bad), 0) == _lib.VIL_E_BADARG
    assert "exact" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.raise_for(_lib.VIL_E_BADARG)
    # exact=1 with a random-shift mode raises in the reference (slidingchunk_2d.py:331-343)
    assert lib.vil_attn_workspace_bytes(ctypes.byref(_params(exact=1, mode=3)), 0) == _lib.VIL_E_BADARG
    # rpe parameters come together (longformer2d.py:68-100): g2l / g2g without a table are rejected, not half-applied
    assert lib.vil_attn_workspace_bytes(ctypes.byref(_params(g2l=256, g2g=256)), 0) == _lib.VIL_E_BADARG
    assert "bias_table" in _lib.last_error()
    # ABI drift guard
    assert lib.vil_attn_workspace_bytes(ctypes.byref(_params(struct_bytes=8)), 0) == _lib.VIL_E_BADARG
    # NULL tensors are rejected before any launch
    assert lib.vil_attn_fwd_sm100(ctypes.byref(_params()), None) == _lib.VIL_E_BADARG


def test_epilogue_validation(lib):
    """addnorm / bias_act entry points validate before any launch (SURVEY.md section 8 (f) row 4 kernels)."""
    a = _lib.VilAddNormParams()
    a.struct_bytes = ctypes.sizeof(_lib.VilAddNormParams)
    a.b_dtype, a.y_dtype, a.C, a.rows, a.eps = _lib.VIL_BF16, _lib.VIL_BF16, 96, 1000, 1e-6
    assert lib.vil_addnorm_workspace_bytes(ctypes.byref(a)) >= 148 * 3 * 96 * 4
    assert lib.vil_addnorm_fwd_sm100(ctypes.byref(a), None) == _lib.VIL_E_BADARG          # NULL tensors
    a.C = 98
    assert lib.vil_addnorm_fwd_sm100(ctypes.byref(a), None) == _lib.VIL_E_UNSUPPORTED     # C % 4
    a.struct_bytes = 8
    assert lib.vil_addnorm_fwd_sm100(ctypes.byref(a), None) == _lib.VIL_E_BADARG
    b = _lib.VilBiasActParams()
    b.struct_bytes = ctypes.sizeof(_lib.VilBiasActParams)
    b.dtype, b.C, b.act, b.rows = _lib.VIL_BF16, 384, _lib.VIL_ACT_GELU, 1000
    assert lib.vil_bias_act_workspace_bytes(ctypes.byref(b)) >= 384 * 4
    assert lib.vil_bias_act_bwd_sm100(ctypes.byref(b), None) == _lib.VIL_E_BADARG
    b.C = 100                                                                             # 200-byte rows: not 16-byte vectors
    assert lib.vil_bias_act_fwd_sm100(ctypes.byref(b), None) == _lib.VIL_E_UNSUPPORTED
    b.C, b.act = 384, 7
    assert lib.vil_bias_act_fwd_sm100(ctypes.byref(b), None) == _lib.VIL_E_BADARG


def test_module_refuses_cpu_tensors():
    import torch
    from vision_longformer_b200 import B200Long2DSCSelfAttention
    mod = B200Long2DSCSelfAttention(32, num_heads=2, w=4, nglo=1)
    with pytest.raises(RuntimeError, match="no CPU"):
        mod(torch.randn(1, 17, 32), 4, 4)


def test_module_state_dict_keys_match_reference_golden():
    import torch
    from tests.util import attn_cases, load_attn
    from vision_longformer_b200 import B200Long2DSCSelfAttention
    for name in attn_cases():
        gold = load_attn(name)
        mod = B200Long2DSCSelfAttention(**gold["kwargs"])
        sd = mod.state_dict()
        assert set(sd.keys()) == set(gold["state_dict"].keys()), name
        for k, v in gold["state_dict"].items():
            assert tuple(sd[k].shape) == tuple(v.shape), (name, k)
        if "relative_position_index" in sd:
            assert torch.equal(sd["relative_position_index"].int(), gold["state_dict"]["relative_position_index"])
