"""Pins the CPU oracle (oracle/) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import attention_ref as A
from oracle import fixtures as fx
from oracle import model_ref, nms_ref
from oracle.detect_ref import detect_decode
from tests.golden_cases import (CATTN_CASES, DETECT_ANCHORS, DETECT_FEATS, DETECT_STRIDES, MF_SHAPES, NMS_CASES, SAM_CASES,
                                SWIN_BIG_CASES, SWIN_CASES, V2ATTN_CASES, nms_kwargs, sam_state_shapes, swin_state_shapes,
                                v2attn_state_shapes)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def swin_params(name):
    dim, res, heads, ws, shift, lin, B = SWIN_CASES[name]
    return {k: fx.deterministic_tensor(k, s, seed=1) for k, s in swin_state_shapes(dim, ws, lin, heads, res).items()}


@pytest.mark.parametrize("name", list(SWIN_CASES))
def test_swin_block_matches_reference(golden, name):
    dim, res, heads, ws, shift, lin, B = SWIN_CASES[name]
    g = golden("swin_blocks")
    x = fx.det_input("swin:" + name, (B, res[0] * res[1], dim))
    y = A.swin_block(x, swin_params(name), "", res[0], res[1], heads, ws, shift, lin)
    assert rel_err(y, g[name + "/y"]) < 2e-6
    assert (y.float() - torch.from_numpy(g[name + "/y"])).abs().max() < 2e-5


@pytest.mark.parametrize("name", list(SWIN_CASES))
def test_closed_form_index_and_mask(golden, name):
    dim, res, heads, ws, shift, lin, B = SWIN_CASES[name]
    g = golden("swin_blocks")
    ws_e, shift_e = A.effective_window(res[0], res[1], ws, shift)
    assert np.array_equal(A.relative_position_index(ws_e, ws_e).numpy(), g[name + "/rel_index"])
    if shift_e > 0:
        assert np.array_equal(A.shift_attn_mask(res[0], res[1], ws_e, shift_e).numpy(), g[name + "/attn_mask"])
    else:
        assert name + "/attn_mask" not in g


@pytest.mark.parametrize("name", list(SWIN_CASES))
def test_qkv_image_contract_equals_block_path(name):
    """The op contract (attention on the un-rolled qkv image) reproduces the window path."""
    dim, res, heads, ws, shift, lin, B = SWIN_CASES[name]
    H, W = res
    p = {k: v.double() for k, v in swin_params(name).items()}
    x = fx.det_input("swin:" + name, (B, H * W, dim)).double()
    ws_e, shift_e = A.effective_window(H, W, ws, shift)
    ln = torch.nn.functional.layer_norm(x, (dim,), p["norm1.weight"], p["norm1.bias"], 1e-5)
    qkv = torch.nn.functional.linear(ln, p["attn.qkv.weight"], p["attn.qkv.bias"]).reshape(B, H, W, 3 * dim)
    o = A.attention_on_qkv_image(qkv, p["attn.relative_position_bias_table"], heads, ws_e, shift_e,
                                 pad_qkv=p["attn.qkv.bias"])
    y1 = x + torch.nn.functional.linear(o.reshape(B, H * W, dim), p["attn.proj.weight"], p["attn.proj.bias"])
    z = torch.nn.functional.layer_norm(y1, (dim,), p["norm2.weight"], p["norm2.bias"], 1e-5)
    y = y1 + A.mlp(z, p, "mlp.", H, W, lin, torch.float64)
    ref = A.swin_block(x, p, "", H, W, heads, ws, shift, lin)
    assert rel_err(y, ref) < 1e-12


@pytest.mark.parametrize("name", list(CATTN_CASES))
def test_cattention_block_matches_reference(golden, name):
    variant, C, heads, hw, ws, B = CATTN_CASES[name]
    g = golden("cattn")
    streams = [fx.det_input(f"cattn:{name}:{i}", (B, hw[0], hw[1], C)) for i in range(4)]
    ln_w = [fx.deterministic_tensor(f"norm{i}.weight", (C,), seed=2) for i in range(1, 5)]
    ln_b = [fx.deterministic_tensor(f"norm{i}.bias", (C,), seed=2) for i in range(1, 5)]
    shift = 1 if variant == "vit_shift1" else 0
    ys = A.cattention_block(streams, ln_w, ln_b, heads, ws=ws, shift=shift)
    for i, y in enumerate(ys):
        ref = g[f"{name}/y{i}"]
        if hw[0] > 64:
            y = y[:, ::5, ::3]
        assert rel_err(y, ref) < 2e-6, (name, i)


def test_cattention_n1_is_identity_on_v():
    """with one token per window CAttention returns v exactly."""
    q, k, v = (fx.det_input(f"n1:{i}", (37, 1, 48)) for i in range(3))
    assert torch.equal(A.cattention(q, k, v, 12), v)


def test_cattention_masked_matches_reference(golden):
    q, k, v = (fx.det_input(f"cattn:masked:{i}", (2 * 4, 16, 48)) for i in range(3))
    mask = A.shift_attn_mask(8, 8, 4, 2)
    y = A.cattention(q.double(), k.double(), v.double(), 12, mask.double())
    assert rel_err(y, golden("cattn")["masked/y"]) < 2e-6


def test_detect_decode_matches_reference(golden):
    g = golden("detect")
    zs = []
    for lvl, shape in enumerate(DETECT_FEATS):
        na = 3
        w = fx.deterministic_tensor(f"m.{lvl}.weight", (na * 13, shape[1], 1, 1), seed=3)
        b = fx.deterministic_tensor(f"m.{lvl}.bias", (na * 13,), seed=3)
        raw = torch.nn.functional.conv2d(fx.det_input(f"detect:{lvl}", shape), w, b)
        anchors = torch.tensor(DETECT_ANCHORS[lvl], dtype=torch.float32).reshape(-1, 2)
        z, xp = detect_decode(raw, anchors, DETECT_STRIDES[lvl])
        zs.append(z)
        assert rel_err(xp, g[f"x{lvl}"]) < 1e-6
    z = torch.cat(zs, 1)
    ref = torch.from_numpy(g["z"]).double()
    assert ((z[..., :4] - ref[..., :4]).abs() / ref[..., :4].abs().clamp_min(1.0)).max() < 1e-5
    assert (z[..., 4:] - ref[..., 4:]).abs().max() < 1e-6


@pytest.mark.parametrize("name", list(NMS_CASES))
def test_nms_matches_reference(golden, name):
    B, R, img, active, seed, kw = NMS_CASES[name]
    g = golden("nms")
    pred = fx.synthetic_predictions(B, R, 8, img, active, seed)
    outs = nms_ref.non_max_suppression(pred, early_stop=(name == "cap_30000"), **nms_kwargs(kw, B, img))
    for i, d in enumerate(outs):
        n = int(g[name + "/count"][i])
        assert d.shape[0] == n, (name, i)
        ref = g[name + "/det"][i, :n]
        # conf and class are bit-exact; merged box coordinates carry only summation-order noise
        assert np.array_equal(d[:, 4:], ref[:, 4:])
        assert np.abs(d[:, :4] - ref[:, :4]).max(initial=0.0) < 1e-3


def test_greedy_nms_matches_torchvision():
    tv = pytest.importorskip("torchvision")
    r = np.random.RandomState(0)
    for trial in range(6):
        n = 700
        xy = r.uniform(0, 100, size=(n, 2)).astype(np.float32)
        wh = r.uniform(0, 30, size=(n, 2)).astype(np.float32)
        boxes = np.concatenate([xy, xy + wh], 1)
        scores = r.uniform(size=n).astype(np.float32)
        if trial >= 2:  # ties, duplicates and zero-area boxes
            scores = np.round(scores * 20) / 20
            boxes[::7] = boxes[1::7][: boxes[::7].shape[0]]
            boxes[::11, 2:] = boxes[::11, :2]
        thr = [0.45, 0.6, 0.3, 0.5, 0.45, 0.7][trial]
        ref = tv.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
        got = nms_ref.greedy_nms(boxes, scores, thr)
        assert np.array_equal(ref, got), trial
        assert np.array_equal(ref[:50], nms_ref.greedy_nms(boxes, scores, thr, max_keep=50))


def test_model_forward_matches_reference(golden):
    g = golden("model_512")
    with open(os.path.join(GOLDEN, "state_dict_512.json")) as f:
        meta = json.load(f)
    p = {}
    for k, (shape, dtype) in meta.items():
        if dtype.startswith("float") and not any(s in k for s in ("attn_mask", "anchors", "anchor_grid")):
            p[k] = fx.deterministic_tensor(k, shape, seed=0)
    p["detect.8.anchor_grid"] = torch.tensor([10, 13, 16, 30, 33, 23], dtype=torch.float32).reshape(1, 1, 3, 1, 1, 2)
    rgb = fx.det_input("model:rgb", (1, 3, 512, 512), kind="uniform")
    ir = fx.det_input("model:ir", (1, 3, 512, 512), kind="uniform")
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    x4 = torch.cat((rgb, ir[:, 0:1]), 1)
    feats = model_ref.backbone_forward(x4, p, dtype=torch.float32)
    for i, t in enumerate(feats):
        assert rel_err(t[0, ::7, ::5, ::3], g[f"feat{i}_sub"]) < 1e-4, i
    pred, raw = model_ref.head_forward(feats, p, dtype=torch.float32)
    assert rel_err(raw[0][0].reshape(-1, 13)[::61], g["raw_rows"]) < 1e-4
    ref = torch.from_numpy(g["pred_rows"])
    got = pred[0, ::61]
    assert ((got[:, :4] - ref[:, :4]).abs() / ref[:, :4].abs().clamp_min(1.0)).max() < 1e-3
    assert (got[:, 4:] - ref[:, 4:]).abs().max() < 1e-4


# ------------------------------------------------------------------------------ round-2 fixtures
@pytest.mark.parametrize("name", list(SWIN_BIG_CASES))
def test_swin_block_at_detector_widths_matches_reference(golden, name):
    dim, res, heads, ws, shift, lin, B = SWIN_BIG_CASES[name]
    p = {k: fx.deterministic_tensor(k, s, seed=1) for k, s in swin_state_shapes(dim, ws, lin, heads, res).items()}
    x = fx.det_input("swin:" + name, (B, res[0] * res[1], dim))
    y = A.swin_block(x, p, "", res[0], res[1], heads, ws, shift, lin)
    assert rel_err(y[:, ::3], golden("swin_blocks_big")[name + "/y"]) < 2e-6


@pytest.mark.parametrize("name", list(V2ATTN_CASES))
def test_cosine_window_attention_matches_reference(golden, name):
    dim, ws, heads, B_, masked = V2ATTN_CASES[name]
    p = {k: fx.deterministic_tensor(k, s, seed=5) for k, s in v2attn_state_shapes(dim, ws, heads).items()}
    x = fx.det_input("v2attn:" + name, (B_, ws * ws, dim))
    mask = A.shift_attn_mask(2 * ws, 2 * ws, ws, ws // 2) if masked else None
    assert rel_err(A.cosine_window_attention(x, p, "", heads, ws, mask), golden("variants")[name + "/y"]) < 2e-6


@pytest.mark.parametrize("name", list(SAM_CASES))
def test_sam_attention_matches_reference(golden, name):
    dim, heads, S, B, rel = SAM_CASES[name]
    p = {k: fx.deterministic_tensor(k, s, seed=6) for k, s in sam_state_shapes(dim, heads, S, rel).items()}
    x = fx.det_input("sam:" + name, (B, S, S, dim))
    assert rel_err(A.sam_attention(x, p, "", heads, rel), golden("variants")[name + "/y"]) < 2e-6


def test_mf_block_matches_reference(golden):
    p = {k: fx.deterministic_tensor(k, s, seed=7) for k, s in MF_SHAPES.items()}
    rgb = fx.det_input("mf:rgb", (2, 3, 24, 32), kind="uniform")
    ir = fx.det_input("mf:ir", (2, 1, 24, 32), kind="uniform")
    assert rel_err(A.mf_block(rgb, ir, p), golden("variants")["mf/y"]) < 2e-6
