"""CPU-side checks of the drop-in boundary: C-ABI symbols, state_dict keys, yaml parsing,
reference-compatible helpers, error behaviour without a GPU.  No kernel is launched."""
import ctypes
import json
import os
import re

import pytest
import torch

import sodt_b200
from oracle import attention_ref as A
from sodt_b200 import _capi, ops
from sodt_b200.basics.models import backbone_vit as bv
from sodt_b200.basics.models.model import Detect, Model, parse_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
YAML = os.path.join(ROOT, "small-object-detection-transformers_b200", "models", "model.yaml")


def header_functions():
    text = open(os.path.join(ROOT, "include", "sodt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sodt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 10
    assert os.path.exists(_capi.LIB_PATH), "build the library first: __graft_entry__.build()"
    handle = ctypes.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/sodt_b200.h but not exported"
    assert sorted(_capi.SIGNATURES) == names, "ctypes signature table out of sync with the header"


def test_library_identity_without_gpu():
    L = _capi.lib()
    assert L.sodt_version() == 100
    assert L.sodt_built_for_sm() == 100
    assert L.sodt_status_string(0) == b"ok"
    assert L.sodt_status_string(-2) == b"unsupported shape"
    assert L.sodt_nms_workspace_bytes(2, 4096, 8, 1) > L.sodt_nms_workspace_bytes(2, 4096, 8, 0) > 0
    # argument validation happens before any CUDA call
    assert L.sodt_window_attn_fwd(None, None, None, None, 1, 8, 8, 16, 1, 8, 0, 0, 1.0, -100.0, None, 0, None) == -1
    assert L.sodt_nms(None, None, 0, None, None, None, None, 0, 1, 1, 1, 0.25, 0.45, 0, 0, 1, 1, 300, 30000, 4096.0, None) == -1


def test_ops_refuse_cpu_tensors():
    qkv = torch.zeros(1, 8, 8, 48)
    with pytest.raises(_capi.SodtError):
        ops.window_attention(qkv, torch.zeros(225, 1), 1, 8)
    with pytest.raises(_capi.SodtError):
        ops.nms(torch.zeros(1, 16, 13))
    with pytest.raises(_capi.SodtError):
        ops.detect_decode(torch.zeros(1, 39, 4, 4), torch.ones(3, 2), 4.0)


def test_state_dict_matches_reference_keys():
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_512.json")))
    m = Model(YAML, input_mode="RGB+IR", ch_steam=3, ch=128, nc=8)
    sd = m.state_dict()
    mine = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()}
    assert mine == meta
    assert sum(p.numel() for p in m.parameters()) == 22007851
    det = m.detect[-1]
    assert isinstance(det, Detect) and det.stride.tolist() == [4.0]
    assert torch.allclose(det.anchors.flatten(), torch.tensor([2.5, 3.25, 4.0, 7.5, 8.25, 5.75]))
    assert abs(det.m[0].bias.view(3, -1)[:, 4].mean().item() - (-8.07)) < 0.2   # objectness prior, model.py:305


def test_parse_model_schema():
    import yaml
    d = yaml.safe_load(open(YAML))
    head, save = parse_model(d, "head", ch=[128])
    assert [type(m).__name__ for m in head] == ["Conv", "Upsample", "Concat", "C3", "Conv", "Upsample", "Concat", "C3", "Detect"]
    assert head[3].cv3.conv.out_channels == 256 and head[7].cv3.conv.out_channels == 128   # width multiple 0.5
    assert len(head[3].m) == 1                                                                # depth multiple 0.33
    assert head[-1].f == [10] and head[2].f == [-1, 1]
    bb, _ = parse_model(d, "backbone", ch=[128])
    assert isinstance(bb, bv.ImageEncoderViT)
    assert [b.shift_size for b in bb.stage1] == [0, 2, 0, 2, 0, 2]
    assert [b.mlp.linear for b in bb.stage1] == [True, False, True, False, True, False]
    assert bb.stage3[0].window_size == 32 and bb.stage3[0].shift_size == 0 and bb.stage3[0].mlp.linear
    assert bb.channel_embed_r.proj.padding == (1, 1) and bb.channel_embed_g.proj.padding == (0, 0)


@pytest.mark.parametrize("H,W,ws,top", [(16, 16, 8, False), (10, 12, 8, False), (14, 14, 7, False), (9, 13, 4, True)])
def test_window_helpers_match_oracle(H, W, ws, top):
    x = torch.randn(2, H, W, 5)
    w1, p1 = bv.window_partition(x, ws, top)
    w2, p2 = A.window_partition(x, ws, top)
    assert p1 == p2 and torch.equal(w1, w2)
    assert torch.equal(bv.window_unpartition(w1, ws, p1, (H, W), top), x)


def test_buffers_match_closed_forms():
    blk = bv.SwinTransformerBlock(48, (12, 16), 3, window_size=8, shift_size=2)
    assert torch.equal(blk.attn_mask, A.shift_attn_mask(12, 16, 8, 2))
    assert torch.equal(blk.attn.relative_position_index, A.relative_position_index(8, 8))
    g = bv.SwinTransformerBlock(64, (8, 8), 2, window_size=32, shift_size=2)
    assert g.window_size == 8 and g.shift_size == 0 and g.attn_mask is None


def test_forward_without_cuda_fails_loudly():
    blk = bv.SwinTransformerBlock(48, (8, 8), 3, window_size=8).eval()
    with pytest.raises(_capi.SodtError):
        blk(torch.randn(1, 64, 48))


def test_cattention_module_forward_matches_reference_golden(golden):
    """CAttention.forward(q, k, v, dimensions, mask) (reference backbone_vit.py:589-616) as a standalone module: host-side
    library math with the reference's order of operations, checked against the reference's own output (masked case)."""
    from oracle import fixtures as fx
    q, k, v = (fx.det_input(f"cattn:masked:{i}", (2 * 4, 16, 48)) for i in range(3))
    mask = A.shift_attn_mask(8, 8, 4, 2)
    y = bv.CAttention(48, 12)(q, k, v, (8, 8), mask)
    ref = torch.as_tensor(golden("cattn")["masked/y"])
    assert y.shape == ref.shape and ((y.double() - ref.double()).norm() / ref.double().norm()).item() < 2e-6
    assert torch.equal(bv.CAttention(48, 12)(q[:, :1], k[:, :1], v[:, :1]), v[:, :1])      # one token per window: returns v


def test_window_attention_module_with_explicit_mask_has_no_cpu_path():
    """WindowAttention.forward(x, mask) with a dense mask tensor (reference backbone_vit.py:961-990) runs on the exact CUDA
    kernel (tests/test_gpu_parity_r2.py checks it against the oracle); like every op of the package it has no CPU fallback."""
    m = bv.WindowAttention(48, (4, 4), 12)
    x = torch.randn(2 * 4, 16, 48)
    mask = A.shift_attn_mask(8, 8, 4, 2)
    with pytest.raises(NotImplementedError):
        m(x, mask)
