"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.

Tolerances (BASELINE.json north_star): attention / fusion outputs rel. error <= 1e-5 in fp32 mode,
<= 2e-2 in bf16 mode (bf16 kernel vs the fp32/fp64 oracle); box coordinates <= 1e-3 px-relative;
NMS kept-index sets bit-exact.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import attention_ref as A
from oracle import fixtures as fx
from oracle import model_ref, nms_ref
from oracle.detect_ref import detect_decode as detect_oracle
from tests.golden_cases import (CATTN_CASES, DETECT_ANCHORS, DETECT_FEATS, DETECT_STRIDES, NMS_CASES, SWIN_CASES, nms_kwargs,
                                swin_state_shapes)

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
YAML = os.path.join(ROOT, "small-object-detection-transformers_b200", "models", "model.yaml")
TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2}


@pytest.fixture(scope="module", autouse=True)
def exact_fp32_libraries():
    """fp32 parity mode: cuBLAS / cuDNN must not silently use TF32."""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def ops():
    from sodt_b200 import ops as o
    return o


# --------------------------------------------------------------------------- window attention op
# (B, H, W, C, heads, ws, shift, use_pad_bias)
WINDOW_CASES = [
    (2, 16, 16, 192, 12, 8, 0, False),    # stage-1 geometry, hd 16
    (2, 16, 16, 192, 12, 8, 2, False),
    (1, 32, 24, 384, 12, 8, 2, False),    # stage-2 geometry, hd 32, rectangular
    (1, 32, 32, 128, 2, 32, 0, False),    # stage-3 geometry, hd 64, N = 1024
    (2, 12, 12, 48, 3, 8, 2, True),       # padding + shift, pad tokens carry the qkv bias
    (1, 10, 12, 48, 3, 8, 0, True),
    (1, 14, 14, 48, 3, 7, 3, False),      # N = 49
    (2, 8, 12, 32, 4, 4, 1, False),       # hd 8, N = 16
    (1, 16, 16, 128, 2, 16, 5, False),    # N = 256 with a shift
    (1, 9, 9, 40, 2, 3, 1, False),        # hd 20 (not a power of two), N = 9
    (3, 8, 8, 24, 4, 8, 0, False),        # hd 6
    (1, 64, 64, 192, 12, 8, 4, False),    # shift = ws/2 (swinv2-style)
    (1, 64, 64, 192, 3, 32, 0, False),    # 4 global windows of 1024 tokens, hd 64 (1024^2-input stage 3)
    (2, 32, 64, 128, 2, 32, 0, False),    # rectangular grid of 1024-token windows
    (1, 8, 24, 64, 4, 8, 2, False),       # odd number of windows (pair kernel tail), hd 16
    (3, 8, 8, 64, 2, 8, 0, False),        # odd number of windows, hd 32
    (2, 40, 24, 192, 12, 8, 7, False),    # largest shift
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", WINDOW_CASES)
def test_window_attention_vs_oracle(case, dtype):
    B, H, W, C, heads, ws, shift, pad = case
    qkv = fx.det_input(f"wa:{case}", (B, H, W, 3 * C))
    table = 0.5 * fx.det_input(f"wa_table:{case}", ((2 * ws - 1) ** 2, heads))
    pad_qkv = 0.3 * fx.det_input(f"wa_pad:{case}", (3 * C,)) if pad else None
    q_dev = qkv.to("cuda", dtype)
    pad_dev = None if pad_qkv is None else pad_qkv.to("cuda", dtype)
    out = ops().window_attention(q_dev, table.cuda(), heads, ws, shift, pad_qkv=pad_dev)
    assert out.shape == (B, H, W, C) and out.dtype == dtype
    # the oracle sees exactly the values the kernel sees (bf16-rounded inputs in bf16 mode)
    ref = A.attention_on_qkv_image(q_dev.float().cpu(), table, heads, ws, shift,
                                   pad_qkv=None if pad_dev is None else pad_dev.float().cpu())
    assert rel_err(out, ref) < TOL[dtype]
    if dtype == torch.float32:
        assert (out.double().cpu() - ref).abs().max() < 5e-5


def test_window_attention_rejects_bad_arguments():
    from sodt_b200._capi import SodtError
    q = torch.zeros(1, 8, 8, 96, device="cuda")
    with pytest.raises(ValueError):
        ops().window_attention(q, torch.zeros(10, 2, device="cuda"), 2, 8)
    with pytest.raises(SodtError):   # shift >= ws
        ops().window_attention(q, torch.zeros(225, 2, device="cuda"), 2, 8, shift=8)
    with pytest.raises(SodtError):   # head_dim 96 > 64 -> unsupported
        ops().window_attention(torch.zeros(1, 8, 8, 288, device="cuda"), torch.zeros(225, 1, device="cuda"), 1, 8)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_window_attention_full_size_properties(dtype):
    """BASELINE config-2 stage-1 geometry (256x256 tokens, C=192, 12 heads, ws 8, shift 2), batch 2:
    size-independent properties instead of an oracle run."""
    B, H, W, C, heads, ws, shift = 2, 256, 256, 192, 12, 8, 2
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B, H, W, 3 * C, device="cuda", generator=g).to(dtype)
    table = (0.5 * torch.randn((2 * ws - 1) ** 2, heads, device="cuda", generator=g))
    # (1) softmax rows sum to one: constant v gives constant output
    q1 = qkv.clone()
    q1[..., 2 * C:] = 0.75
    o1 = ops().window_attention(q1, table, heads, ws, shift)
    assert (o1.float() - 0.75).abs().max() < (1e-5 if dtype == torch.float32 else 8e-3)
    # (2) linear in v
    qa, qb = qkv.clone(), qkv.clone()
    qb[..., 2 * C:] = torch.randn(B, H, W, C, device="cuda", generator=g).to(dtype)
    qs = qkv.clone()
    qs[..., 2 * C:] = (qa[..., 2 * C:].float() + qb[..., 2 * C:].float()).to(dtype)
    oa, ob, os_ = (ops().window_attention(t, table, heads, ws, shift).float() for t in (qa, qb, qs))
    assert ((oa + ob - os_).norm() / os_.norm()).item() < (1e-5 if dtype == torch.float32 else 2e-2)
    # (3) translating the image by whole windows (cyclically) translates the output when shift == 0
    o0 = ops().window_attention(qkv, table, heads, ws, 0)
    o0r = ops().window_attention(torch.roll(qkv, (ws, 2 * ws), (1, 2)), table, heads, ws, 0)
    assert torch.equal(torch.roll(o0, (ws, 2 * ws), (1, 2)), o0r)
    # (4) a sampled window against the oracle: with shift == 0 windows are independent crops
    crop = qkv[:1, 40:48, 80:88].float().cpu()
    ref = A.attention_on_qkv_image(crop, table.cpu(), heads, ws, 0)
    assert rel_err(o0[:1, 40:48, 80:88], ref) < TOL[dtype]


@pytest.mark.parametrize("H,C,ws,shift", [(512, 192, 8, 0), (512, 192, 8, 2), (512, 192, 8, 4), (128, 384, 8, 2), (64, 768, 32, 0)])
def test_window_attention_microbench_geometries(H, C, ws, shift):
    """The window-attention microbenchmark geometries of SURVEY.md section 8(d) C3 (2048^2 px at stride 4: 512^2 tokens, C=192;
    128^2 tokens with C=384; the stage-3 geometry N=1024, head_dim 64), bf16, batch 1: interior windows are independent crops
    of the image shifted by `shift`, so sampled windows are checked against the oracle; constant-v gives the row-sum check."""
    heads, B = 12, 1
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = torch.randn(B, H, H, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
    table = 0.5 * torch.randn((2 * ws - 1) ** 2, heads, device="cuda", generator=g)
    out = ops().window_attention(qkv, table, heads, ws, shift)
    for wy, wx in ((0, 0), (3 % (H // ws - 1), 5 % (H // ws - 1)), (H // ws - 2, H // ws - 2)):      # never the last (masked) row / column
        y0, x0 = wy * ws + shift, wx * ws + shift
        crop = qkv[:1, y0:y0 + ws, x0:x0 + ws].float().cpu()
        ref = A.attention_on_qkv_image(crop, table.cpu(), heads, ws, 0)
        assert rel_err(out[:1, y0:y0 + ws, x0:x0 + ws], ref) < TOL[torch.bfloat16], (wy, wx)
    q1 = qkv.clone()
    q1[..., 2 * C:] = -1.25
    assert (ops().window_attention(q1, table, heads, ws, shift).float() + 1.25).abs().max() < 1.6e-2


# ------------------------------------------------------------------------------------ fused Linear
@pytest.mark.parametrize("M,N,K,act,res", [(1000, 576, 192, None, False), (777, 768, 192, "gelu", False), (4096, 192, 768, None, True),
                                           (300, 192, 192, None, True), (129, 1152, 384, None, False), (2048, 1536, 384, "gelu", False),
                                           (513, 384, 1536, None, True), (640, 3072, 768, "gelu", False), (100, 768, 3072, None, True),
                                           (50000, 768, 192, "gelu", False),
                                           # K = 384 with enough rows for the resident-W CTA pairs (ragged last tile, odd tile count)
                                           (13000, 1152, 384, None, False), (20077, 384, 384, None, True), (20000, 1536, 384, "gelu", False),
                                           (16500, 256, 384, None, False)])
def test_linear_tc_vs_torch(M, N, K, act, res):
    x = fx.det_input(f"lin_x:{M}:{K}", (M, K)).to("cuda", torch.bfloat16)
    w = (fx.det_input(f"lin_w:{N}:{K}", (N, K)) / K ** 0.5).to("cuda", torch.bfloat16)
    b = 0.2 * fx.det_input(f"lin_b:{N}", (N,))
    r = fx.det_input(f"lin_r:{M}:{N}", (M, N)).to("cuda", torch.bfloat16) if res else None
    assert ops()._capi.lib().sodt_linear_supported(M, N, K, 1) == 1
    out = ops().linear(x, w, b.cuda(), act=act, residual=r)
    ref = x.double().cpu() @ w.double().cpu().t() + b.double()
    if act == "gelu":
        ref = torch.nn.functional.gelu(ref)
    if res:
        ref = ref + r.double().cpu()
    assert out.shape == (M, N) and out.dtype == torch.bfloat16
    assert rel_err(out, ref) < 4e-3
    assert (out.double().cpu() - ref).abs().max() < 0.06 * max(1.0, ref.abs().max().item() / 8)


@pytest.mark.parametrize("M,N,K,act", [(1000, 64, 64, "silu"), (777, 128, 384, "silu"), (4096, 128, 512, None), (130, 64, 128, "gelu"),
                                       (3000, 320, 64, "silu")])
def test_linear_tc_small_n_and_silu(M, N, K, act):
    x = fx.det_input(f"lin2_x:{M}:{K}", (M, K)).to("cuda", torch.bfloat16)
    w = (fx.det_input(f"lin2_w:{N}:{K}", (N, K)) / K ** 0.5).to("cuda", torch.bfloat16)
    b = 0.2 * fx.det_input(f"lin2_b:{N}", (N,))
    out = ops().linear(x, w, b.cuda(), act=act)
    ref = x.double().cpu() @ w.double().cpu().t() + b.double()
    ref = torch.nn.functional.silu(ref) if act == "silu" else torch.nn.functional.gelu(ref) if act == "gelu" else ref
    assert out.shape == (M, N) and rel_err(out, ref) < 4e-3


def test_linear_tc_strided_operands_and_split_k_sources():
    """Column slices as x / residual / out (row strides) and an A operand continued by a second tensor (x2)."""
    M, K1, K2, N = 1500, 192, 192, 256
    wide = fx.det_input("lin3_wide", (M, 512)).to("cuda", torch.bfloat16)
    x1, x2 = wide[:, 64:64 + K1], fx.det_input("lin3_x2", (M, K2)).to("cuda", torch.bfloat16)
    w = (fx.det_input("lin3_w", (N, K1 + K2)) / 20).to("cuda", torch.bfloat16)
    rwide = fx.det_input("lin3_r", (M, 2 * N)).to("cuda", torch.bfloat16)
    obuf = torch.full((M, 3 * N), 7.0, dtype=torch.bfloat16, device="cuda")
    out = ops().linear(x1, w, None, residual=rwide[:, N:], x2=x2, out=obuf[:, N:2 * N])
    assert out.data_ptr() == obuf[:, N:2 * N].data_ptr()
    ref = torch.cat((x1, x2), 1).double().cpu() @ w.double().cpu().t() + rwide[:, N:].double().cpu()
    assert rel_err(obuf[:, N:2 * N], ref) < 4e-3
    assert (obuf[:, :N] == 7).all() and (obuf[:, 2 * N:] == 7).all()      # neighbours of the slice untouched


@pytest.mark.parametrize("M,K,N,act", [(1000, 192, 576, None), (515, 384, 1536, "gelu"), (300, 768, 768, None), (4096, 192, 192, None),
                                       (16384 + 300, 384, 1152, None)])      # resident-W CTA pairs on both GEMMs
def test_linear_tc_layernorm_fold_vs_torch(M, K, N, act):
    """LayerNorm folded into the GEMM: statistics from row_stats and from the producing GEMM's epilogue (want_stats)."""
    o = ops()
    # producer GEMM writes x (with a residual, like proj / fc2) and its row statistics
    a = fx.det_input(f"lnf_a:{M}:{K}", (M, K)).to("cuda", torch.bfloat16)
    wp = (fx.det_input(f"lnf_wp:{K}", (K, K)) / K ** 0.5).to("cuda", torch.bfloat16)
    r = (3.0 * fx.det_input(f"lnf_r:{M}:{K}", (M, K)) + 1.5).to("cuda", torch.bfloat16)         # non-zero mean rows
    x, st = o.linear(a, wp, None, residual=r, want_stats=True)
    assert st.shape == (K // 64, M, 2)
    xd = x.double().cpu()
    # emitted from the fp32 values before the bf16 rounding of x
    assert rel_err(st.sum(0)[:, 0], xd.sum(1)) < 2e-3 and rel_err(st.sum(0)[:, 1], (xd * xd).sum(1)) < 2e-3
    mean, var = xd.mean(1), xd.var(1, unbiased=False)
    mr0, mr1 = o.finalize_stats(st, K, 1e-5), o.row_stats(x, 1e-5)
    assert rel_err(mr1[:, 0], mean) < 1e-5 and rel_err(mr1[:, 1], (var + 1e-5).rsqrt()) < 1e-4
    assert rel_err(mr0[:, 0], mean) < 2e-3 and rel_err(mr0[:, 1], (var + 1e-5).rsqrt()) < 2e-3
    # consumer: Linear(LayerNorm(x))
    w = (fx.det_input(f"lnf_w:{N}:{K}", (N, K)) / K ** 0.5).to("cuda", torch.bfloat16)
    b = (0.2 * fx.det_input(f"lnf_b:{N}", (N,))).to("cuda", torch.bfloat16)
    g = (1.0 + 0.2 * fx.det_input(f"lnf_g:{K}", (K,))).to("cuda", torch.bfloat16)
    be = (0.1 * fx.det_input(f"lnf_be:{K}", (K,))).to("cuda", torch.bfloat16)
    ref = torch.nn.functional.layer_norm(xd, (K,), g.double().cpu(), be.double().cpu(), 1e-5) @ w.double().cpu().t() + b.double().cpu()
    ref = torch.nn.functional.gelu(ref) if act == "gelu" else ref
    for mr in (mr0, mr1):
        out = o.linear(x, w, b, act=act, ln=(mr, g, be))
        assert out.shape == (M, N) and rel_err(out, ref) < 6e-3


def test_linear_tc_batch_broadcast_residual():
    """A residual with fewer rows than the output repeats over the batch (pos_embed in the patch-embedding GEMM)."""
    B, L, K, N = 3, 256, 192, 192
    x = fx.det_input("lin4_x", (B, L, K)).to("cuda", torch.bfloat16)
    w = (fx.det_input("lin4_w", (N, K)) / 14).to("cuda", torch.bfloat16)
    b = 0.2 * fx.det_input("lin4_b", (N,))
    pos = fx.det_input("lin4_pos", (1, L, N)).to("cuda", torch.bfloat16)
    out = ops().linear(x, w, b.cuda(), residual=pos)
    ref = x.double().cpu() @ w.double().cpu().t() + b.double() + pos.double().cpu()
    assert out.shape == (B, L, N) and rel_err(out, ref) < 4e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,pad,act", [
    (2, 16, 256, 192, 192, (2, 2), (0, 0), "gelu"),      # conv-MLP geometry, stage 1 row width
    (3, 8, 128, 384, 384, (2, 2), (0, 0), "gelu"),       # stage 2 row width
    (2, 64, 64, 128, 128, (3, 3), (1, 1), "silu"),       # head 3x3, two image rows per tile
    (1, 128, 128, 64, 64, (3, 3), (1, 1), "silu"),
    (2, 32, 32, 64, 128, (1, 1), (0, 0), "silu"),        # four rows per tile
    (2, 4, 256, 128, 64, (1, 1), (0, 0), None),
    (1, 16, 16, 64, 64, (3, 3), (1, 1), None),           # eight rows per tile
    (4, 32, 256, 192, 192, (2, 2), (0, 0), "gelu"),      # enough rows for CTA pairs: halo tiles on cta_group::2
    (2, 128, 128, 128, 128, (3, 3), (1, 1), "silu"),     # 3x3 halo tiles on CTA pairs (K = 1152)
    (2, 8, 256, 64, 64, (2, 2), (1, 1), None),           # 2x2 with the padding above / left
    (1, 8, 128, 64, 128, (3, 2), (2, 0), "silu"),        # kh != kw, two padding rows above
    (1, 4, 384, 64, 64, (1, 3), (0, 2), None),           # three tiles per image row
])
def test_conv2d_nhwc_tap_gemm_vs_torch(B, H, W, Cin, Cout, k, pad, act):
    x = fx.det_input(f"cv_x:{B}:{H}:{W}:{Cin}", (B, H, W, Cin)).to("cuda", torch.bfloat16)
    w = (fx.det_input(f"cv_w:{Cout}:{Cin}:{k}", (Cout, Cin, *k)) / (Cin * k[0] * k[1]) ** 0.5).to("cuda", torch.bfloat16)
    b = 0.2 * fx.det_input(f"cv_b:{Cout}", (Cout,))
    assert ops().conv2d_nhwc_supported(x, Cout, *k)
    out = ops().conv2d_nhwc(x, ops().conv_weight_taps(w), b.cuda(), k, pad, act)
    xin = x.double().cpu().permute(0, 3, 1, 2)
    # zero rows / columns: pad[0] above, pad[1] left, and as many below / right as the kernel needs for a same-size output
    xin = torch.nn.functional.pad(xin, (pad[1], k[1] - 1 - pad[1], pad[0], k[0] - 1 - pad[0]))
    ref = torch.nn.functional.conv2d(xin, w.double().cpu(), b.double())
    ref = torch.nn.functional.silu(ref) if act == "silu" else torch.nn.functional.gelu(ref) if act == "gelu" else ref
    assert out.shape == (B, H, W, Cout)
    assert rel_err(out, ref.permute(0, 2, 3, 1)) < 4e-3


@pytest.mark.parametrize("B,H,W,C1,C2,Cout,act", [(2, 8, 256, 128, 256, 128, "silu"), (2, 16, 128, 256, 256, 256, "silu"),
                                                     (1, 32, 64, 64, 128, 64, None), (3, 16, 32, 128, 64, 128, "silu")])
def test_conv2d_nhwc_upsample_concat_as_operand_addressing(B, H, W, C1, C2, Cout, act):
    """1x1 conv over cat(up2x(low), skip) through zero-stride TMA dimensions == the materialised upsample + concat + conv."""
    low = fx.det_input(f"uc_low:{B}:{H}:{W}:{C1}", (B, H // 2, W // 2, C1)).to("cuda", torch.bfloat16)
    skip = fx.det_input(f"uc_skip:{B}:{H}:{W}:{C2}", (B, H, W, C2)).to("cuda", torch.bfloat16)
    w = (fx.det_input(f"uc_w:{Cout}:{C1 + C2}", (Cout, C1 + C2)) / (C1 + C2) ** 0.5).to("cuda", torch.bfloat16)
    b = 0.2 * fx.det_input(f"uc_b:{Cout}", (Cout,))
    uc = ops().UpCat(low.permute(0, 3, 1, 2), skip.permute(0, 3, 1, 2))
    assert ops().upcat_conv1x1_supported(uc, Cout)
    out = ops().upcat_conv1x1(uc, w, b.cuda(), act)
    up = low.double().cpu().repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    ref = torch.cat((up, skip.double().cpu()), -1) @ w.double().cpu().t() + b.double()
    ref = torch.nn.functional.silu(ref) if act == "silu" else ref
    assert out.shape == (B, H, W, Cout) and rel_err(out, ref) < 4e-3
    assert torch.equal(uc.materialize().permute(0, 2, 3, 1), torch.cat((up, skip.double().cpu()), -1).to(torch.bfloat16).cuda())


def test_conv2d_nhwc_channel_slices():
    """Input and output as channel slices of wider NHWC buffers (the head's C3 without torch.cat)."""
    B, H, W, C = 2, 32, 128, 64
    wide = fx.det_input("cvs_x", (B, H, W, 2 * C)).to("cuda", torch.bfloat16)
    w = (fx.det_input("cvs_w", (C, C, 3, 3)) / 24).to("cuda", torch.bfloat16)
    b = 0.1 * fx.det_input("cvs_b", (C,))
    obuf = torch.zeros((B, H, W, 2 * C), dtype=torch.bfloat16, device="cuda")
    ops().conv2d_nhwc(wide[..., C:], ops().conv_weight_taps(w), b.cuda(), (3, 3), (1, 1), "silu", out=obuf[..., :C])
    ref = torch.nn.functional.silu(torch.nn.functional.conv2d(wide[..., C:].double().cpu().permute(0, 3, 1, 2), w.double().cpu(),
                                                              b.double(), padding=1)).permute(0, 2, 3, 1)
    assert rel_err(obuf[..., :C], ref) < 4e-3 and (obuf[..., C:] == 0).all()


@pytest.mark.parametrize("B,H,W,C,N", [(2, 16, 256, 192, 384), (2, 32, 128, 384, 768), (3, 8, 64, 64, 128), (1, 64, 16, 128, 64)])
def test_patch_merge_linear_vs_torch(B, H, W, C, N):
    x = fx.det_input(f"pm_x:{B}:{H}:{W}:{C}", (B, H, W, C)).to("cuda", torch.bfloat16)
    w = (fx.det_input(f"pm_w:{N}:{C}", (N, 4 * C)) / (4 * C) ** 0.5).to("cuda", torch.bfloat16)
    assert ops()._capi.lib().sodt_patch_merge_linear_supported(B, H, W, C, N, 1) == 1
    out = ops().patch_merge_linear(x, w)
    xd = x.double().cpu()
    g = torch.cat([xd[:, 0::2, 0::2], xd[:, 1::2, 0::2], xd[:, 0::2, 1::2], xd[:, 1::2, 1::2]], -1)   # reference order
    ref = g.reshape(B, -1, 4 * C) @ w.double().cpu().t()
    assert out.shape == (B, (H // 2) * (W // 2), N) and rel_err(out, ref) < 4e-3


# ------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,C,with_r,with_e", [(1000, 192, False, True), (257, 384, True, True), (64, 768, True, False),
                                                  (33, 48, False, False), (7, 1024, True, True), (5, 250, False, True),
                                                  (4099, 192, True, True), (1001, 384, False, False), (515, 768, False, True)])
def test_add_layernorm_vs_torch(rows, C, with_r, with_e, dtype):
    a = fx.det_input(f"ln_a:{rows}:{C}", (rows, C)).to("cuda", dtype)
    r = fx.det_input(f"ln_r:{rows}:{C}", (rows, C)).to("cuda", dtype) if with_r else None
    w = 1.0 + 0.1 * fx.det_input("ln_w", (C,))
    b = 0.1 * fx.det_input("ln_b", (C,))
    e = 0.2 * fx.det_input("ln_e", (C,)) if with_e else None
    y, s = ops().add_layernorm(a, r, w.cuda(), b.cuda(), 1e-5, extra_bias=None if e is None else e.cuda(), want_sum=True)
    sd = a.double().cpu() + (r.double().cpu() if with_r else 0.0)
    ref = torch.nn.functional.layer_norm(sd, (C,), w.double(), b.double(), 1e-5)
    assert rel_err(y, ref) < (2e-6 if dtype == torch.float32 else 5e-3)
    assert rel_err(s, sd + (e.double() if with_e else 0.0)) < (1e-6 if dtype == torch.float32 else 5e-3)


# ------------------------------------------------------------------------------------ swin block
def swin_params(name):
    dim, res, heads, ws, shift, lin, B = SWIN_CASES[name]
    return {k: fx.deterministic_tensor(k, s, seed=1) for k, s in swin_state_shapes(dim, ws, lin, heads, res).items()}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", list(SWIN_CASES))
def test_swin_block_module_vs_reference_golden(golden, name, dtype):
    from sodt_b200.basics.models.backbone_vit import SwinTransformerBlock
    dim, res, heads, ws, shift, lin, B = SWIN_CASES[name]
    blk = SwinTransformerBlock(dim, res, heads, window_size=ws, shift_size=shift, linear_mlp=lin).eval()
    missing = blk.load_state_dict(swin_params(name), strict=False)
    assert set(missing.missing_keys) <= {"attn_mask", "attn.relative_position_index"} and not missing.unexpected_keys
    blk = blk.to("cuda", dtype)
    x = fx.det_input("swin:" + name, (B, res[0] * res[1], dim)).to("cuda", dtype)
    with torch.no_grad():
        y = blk(x)
    assert rel_err(y, golden("swin_blocks")[name + "/y"]) < (2e-5 if dtype == torch.float32 else 2e-2)


# ------------------------------------------------------------------------- cross-channel block
@pytest.mark.parametrize("layout", ["nhwc", "nchw_view"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", list(CATTN_CASES))
def test_cattn_block_vs_reference_golden(golden, name, dtype, layout):
    variant, C, heads, hw, ws, B = CATTN_CASES[name]
    g = golden("cattn")
    streams = [fx.det_input(f"cattn:{name}:{i}", (B, hw[0], hw[1], C)).to("cuda", dtype) for i in range(4)]
    if layout == "nchw_view":   # what PatchEmbed.forward hands over: a permuted view of NCHW memory
        streams = [s.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1) for s in streams]
    ln_w = torch.stack([fx.deterministic_tensor(f"norm{i}.weight", (C,), seed=2) for i in range(1, 5)]).cuda()
    ln_b = torch.stack([fx.deterministic_tensor(f"norm{i}.bias", (C,), seed=2) for i in range(1, 5)]).cuda()
    shift = 1 if variant == "vit_shift1" else 0
    out = ops().cattn_block(*streams, ln_w, ln_b, heads, ws=ws, shift=shift)
    assert out.shape == (B, hw[0], hw[1], 4 * C)
    for i, y in enumerate(torch.split(out, C, dim=-1)):
        ref = g[f"{name}/y{i}"]
        if hw[0] > 64:
            y = y[:, ::5, ::3]
        assert rel_err(y, ref) < TOL[dtype], (name, i)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,heads,ws,shift", [(48, 12, 4, 2), (24, 3, 2, 1), (96, 6, 8, 3), (48, 4, 3, 1), (40, 5, 1, 0)])
def test_cattn_block_shifted_vs_oracle(C, heads, ws, shift, dtype):
    """Masked / shifted general-window path (never exercised by the shipped model, part of the contract)."""
    B, h, w = 2, 16, 12
    streams = [fx.det_input(f"cattn_s:{C}:{ws}:{i}", (B, h, w, C)).to("cuda", dtype) for i in range(4)]
    ln_w = 1.0 + 0.1 * fx.det_input("cattn_s:w", (4, C))
    ln_b = 0.1 * fx.det_input("cattn_s:b", (4, C))
    out = ops().cattn_block(*streams, ln_w.cuda(), ln_b.cuda(), heads, ws=ws, shift=shift)
    ref = torch.cat(A.cattention_block([s.float().cpu() for s in streams], list(ln_w), list(ln_b), heads, ws, shift), -1)
    assert rel_err(out, ref) < TOL[dtype]


# ------------------------------------------------------------------------ head glue
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_upsample2x_concat_vs_torch(dtype):
    low = fx.det_input("up:low", (2, 16, 5, 7)).to("cuda", dtype).contiguous(memory_format=torch.channels_last)
    skip = fx.det_input("up:skip", (2, 24, 10, 14)).to("cuda", dtype)
    out = ops().upsample2x_concat(low, skip)
    ref = torch.cat((torch.nn.functional.interpolate(low, scale_factor=2, mode="nearest"), skip), 1)
    assert out.shape == ref.shape and torch.equal(out, ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", ["gelu", "silu", None])
def test_bias_act_crop_vs_torch(act, dtype):
    x = fx.det_input(f"bac:{act}", (2, 48, 9, 11)).to("cuda", dtype).contiguous(memory_format=torch.channels_last)
    b = 0.3 * fx.det_input("bac:b", (48,))
    out = ops().bias_act_crop(x, b.cuda(), act, (8, 10), (1, 1))
    ref = x.double().cpu()[:, :, 1:, 1:] + b.double().view(1, -1, 1, 1)
    if act == "gelu":
        ref = torch.nn.functional.gelu(ref)
    elif act == "silu":
        ref = torch.nn.functional.silu(ref)
    assert out.shape == (2, 8, 10, 48)
    assert rel_err(out, ref.permute(0, 2, 3, 1)) < (2e-6 if dtype == torch.float32 else 5e-3)


def test_fused_model_matches_unfused():
    """Model.fuse() (Conv+BN folding, reference model.py:317-325) must not change the predictions."""
    from sodt_b200.basics.models.model import Model
    m = load_det_weights(Model(YAML, input_mode="RGB+IR", ch_steam=3, ch=128, nc=8)).eval().cuda()
    rgb = fx.det_input("model:rgb", (1, 3, 512, 512), kind="uniform").cuda()
    ir = fx.det_input("model:ir", (1, 3, 512, 512), kind="uniform").cuda()
    with torch.no_grad():
        p0 = m(rgb, ir, "RGB+IR")[0]
        p1 = m.fuse()(rgb, ir, "RGB+IR")[0]
    assert rel_err(p1, p0) < 1e-5


# ------------------------------------------------------------------------------------- front end
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("pad_r", [1, 0])
def test_frontend_vs_oracle(pad_r, dtype):
    """Fused channel embeddings + cross-channel block against conv2d + the cross-channel oracle."""
    B, H, W, E = 2, 64, 48, 48
    x = fx.det_input("fe:x", (B, 4, H, W), kind="uniform").to("cuda", dtype)
    cw = 0.3 * fx.det_input("fe:cw", (4, E, 16))
    cb = 0.1 * fx.det_input("fe:cb", (4, E))
    lw = 1.0 + 0.1 * fx.det_input("fe:lw", (4, E))
    lb = 0.1 * fx.det_input("fe:lb", (4, E))
    out = ops().frontend(x, cw.cuda(), cb.cuda(), lw.cuda(), lb.cuda(), pad_r=pad_r)
    xc = x.double().cpu()
    streams = []
    for s in range(4):
        pad = pad_r if s == 0 else 0
        e = torch.nn.functional.conv2d(xc[:, s:s + 1], cw[s].double().reshape(E, 1, 4, 4), cb[s].double(), 4, pad)
        streams.append(e.permute(0, 2, 3, 1))
    ref = torch.cat(A.cattention_block(streams, list(lw), list(lb), 12, ws=1), -1)
    assert out.shape == ref.shape
    assert rel_err(out, ref) < (1e-5 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_frontend_u8_matches_converted_input(dtype):
    """uint8 images straight into the front-end kernel == the separate /255 conversion + channel pick + concat."""
    B, H, W, E = 2, 64, 96, 48
    g = torch.Generator().manual_seed(11)
    rgb = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g).cuda()
    ir = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g).cuda()
    cw = (0.3 * fx.det_input("fe8:cw", (4, E, 16))).cuda()
    cb = (0.1 * fx.det_input("fe8:cb", (4, E))).cuda()
    lw = (1.0 + 0.1 * fx.det_input("fe8:lw", (4, E))).cuda()
    lb = (0.1 * fx.det_input("fe8:lb", (4, E))).cuda()
    x = torch.cat((rgb.to(dtype).div_(255.0), ir[:, 0:1].to(dtype).div_(255.0)), 1)
    want = ops().frontend(x, cw, cb, lw, lb, pad_r=1, eps=1e-5)
    got = ops().frontend_u8(rgb, ir, cw, cb, lw, lb, dtype, pad_r=1, eps=1e-5)
    assert got.dtype == dtype and torch.equal(got, want)


# ---------------------------------------------------------------------------------------- Detect
@pytest.mark.parametrize("memory_format", ["nchw", "channels_last"])
def test_detect_module_vs_reference_golden(golden, memory_format):
    from sodt_b200.basics.models.model import Detect
    g = golden("detect")
    det = Detect(nc=8, anchors=DETECT_ANCHORS, ch=(16, 24)).eval()
    det.stride = torch.tensor(DETECT_STRIDES)
    sd = {k: fx.deterministic_tensor(k, v.shape, seed=3) for k, v in det.state_dict().items() if k.startswith("m.")}
    det.load_state_dict(sd, strict=False)
    det = det.cuda()
    feats = [fx.det_input(f"detect:{i}", s).cuda() for i, s in enumerate(DETECT_FEATS)]
    if memory_format == "channels_last":
        feats = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    with torch.no_grad():
        z, xs = det(feats)
    ref = torch.from_numpy(g["z"]).double()
    zc = z.double().cpu()
    assert ((zc[..., :4] - ref[..., :4]).abs() / ref[..., :4].abs().clamp_min(1.0)).max() < 1e-3 * 1e-2
    assert (zc[..., 4:] - ref[..., 4:]).abs().max() < 1e-5
    for i, x in enumerate(xs):
        assert rel_err(x, g[f"x{i}"]) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_detect_decode_full_size_vs_oracle(dtype):
    """BASELINE config 5 decode geometry: 256x256 cells, 3 anchors, 13 outputs."""
    raw = fx.det_input("detect_big", (2, 39, 256, 256)).to("cuda", dtype)
    anchors = torch.tensor(DETECT_ANCHORS[0], dtype=torch.float32).reshape(-1, 2)
    z, xp = ops().detect_decode(raw, anchors.cuda(), 4.0)
    zr, xr = detect_oracle(raw.float().cpu(), anchors, 4.0)
    zc = z.double().cpu()
    assert ((zc[..., :4] - zr[..., :4]).abs() / zr[..., :4].abs().clamp_min(1.0)).max() < 1e-5
    assert (zc[..., 4:] - zr[..., 4:]).abs().max() < 1e-6
    assert torch.equal(xp.float().cpu(), xr.float())


# ------------------------------------------------------------------------------------------- NMS
def run_cuda_nms(pred_np, kw):
    pred = torch.from_numpy(pred_np).cuda()
    det, counts, keep = ops().nms(pred, want_keep_idx=True, **kw)
    torch.cuda.synchronize()
    return det.cpu().numpy(), counts.cpu().numpy(), keep.cpu().numpy()


@pytest.mark.parametrize("name", list(NMS_CASES))
def test_nms_vs_reference_golden_and_oracle(golden, name):
    B, R, img, active, seed, kw = NMS_CASES[name]
    g = golden("nms")
    pred = fx.synthetic_predictions(B, R, 8, img, active, seed)
    if "labels_seed" in kw:         # apriori labels (general.py:451-458) are a feature of the reference-signature wrapper
        from sodt_b200.basics.utils.general import non_max_suppression
        kwl = nms_kwargs(kw, B, img)
        dets = non_max_suppression(torch.from_numpy(pred).cuda(), **{**kwl, "labels": [torch.from_numpy(l) for l in kwl["labels"]]})
        outs = nms_ref.non_max_suppression(pred, early_stop=True, **kwl)
        assert [d.shape[0] for d in dets] == list(g[name + "/count"])
        for i, d in enumerate(dets):
            d, ref = d.cpu().numpy(), g[name + "/det"][i, :d.shape[0]]
            assert np.array_equal(d[:, 4:], ref[:, 4:]) and np.array_equal(d[:, 4:], outs[i][:, 4:])
            assert np.abs(d[:, :4] - ref[:, :4]).max(initial=0.0) < 1e-3
        return
    det, counts, keep = run_cuda_nms(pred, kw)
    outs, idxs = nms_ref.non_max_suppression(pred, return_indices=True, early_stop=True, **kw)
    assert np.array_equal(counts, g[name + "/count"])
    for i in range(B):
        n = int(counts[i])
        assert np.array_equal(keep[i, :n], idxs[i]), "kept-index set must be bit-exact"
        assert np.all(keep[i, n:] == -1)
        ref = g[name + "/det"][i, :n]
        assert np.array_equal(det[i, :n, 4:], ref[:, 4:])           # confidence and class: bit-exact
        assert np.abs(det[i, :n, :4] - ref[:, :4]).max(initial=0.0) < 1e-3
        assert np.all(det[i, n:] == 0)


@pytest.mark.parametrize("seed,R,img,active,kw", [
    (11, 20000, 512, 0.3, dict(conf_thres=0.25, iou_thres=0.45)),
    (12, 20000, 160, 0.5, dict(conf_thres=0.3, iou_thres=0.5)),                       # heavy overlap, many chunks
    (13, 6000, 512, 0.1, dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)),
    (14, 3000, 64, 1.0, dict(conf_thres=0.5, iou_thres=0.2, agnostic=True)),          # extreme overlap
    (15, 1000, 300, 1.0, dict(conf_thres=0.01, iou_thres=0.45, multi_label=True, classes=[0, 3, 7])),
])
def test_nms_random_vs_oracle(seed, R, img, active, kw):
    pred = fx.synthetic_predictions(2, R, 8, img, active, seed)
    det, counts, keep = run_cuda_nms(pred, kw)
    outs, idxs = nms_ref.non_max_suppression(pred, return_indices=True, early_stop=True, **kw)
    for i in range(2):
        n = int(counts[i])
        assert n == len(idxs[i])
        assert np.array_equal(keep[i, :n], idxs[i])
        assert np.array_equal(det[i, :n, 4:], outs[i][:, 4:])
        assert np.abs(det[i, :n, :4] - outs[i][:, :4]).max(initial=0.0) < 1e-3


def test_nms_ties_and_degenerate_boxes_vs_oracle():
    """Equal scores (stable order), duplicate boxes, zero-area boxes (NaN IoU keeps both)."""
    r = np.random.RandomState(5)
    R = 1500
    pred = np.zeros((1, R, 6), dtype=np.float32)   # nc = 1
    pred[0, :, 0:2] = r.uniform(20, 80, size=(R, 2))
    pred[0, :, 2:4] = r.uniform(0, 30, size=(R, 2))
    pred[0, :, 4] = np.round(r.uniform(0.3, 1.0, size=R) * 16) / 16   # many exact ties
    pred[0, :, 5] = 1.0
    pred[0, ::9, :4] = pred[0, 1::9, :4][: pred[0, ::9].shape[0]]      # duplicates
    pred[0, ::13, 2:4] = 0.0                                           # zero area
    kw = dict(conf_thres=0.25, iou_thres=0.5)
    det, counts, keep = run_cuda_nms(pred, kw)
    outs, idxs = nms_ref.non_max_suppression(pred, return_indices=True, **kw)
    n = int(counts[0])
    assert n == len(idxs[0]) and np.array_equal(keep[0, :n], idxs[0])


def test_nms_core_vs_torchvision_cuda():
    """The suppression core against the third-party kernel the reference actually calls on a GPU
    (torchvision.ops.nms, general.py:496), fed the same class-offset boxes and scores."""
    tv = pytest.importorskip("torchvision")
    pred = fx.synthetic_predictions(1, 8000, 8, 200, 0.6, 21)
    kw = dict(conf_thres=0.25, iou_thres=0.45)
    x = nms_ref.candidates(pred[0], 0.25, False)
    boxes = torch.from_numpy(x[:, :4] + x[:, 5:6] * np.float32(4096)).cuda()
    scores = torch.from_numpy(x[:, 4]).cuda()
    ref_keep = tv.ops.nms(boxes, scores, 0.45)[:300].cpu().numpy()
    pred_t = torch.from_numpy(pred).cuda()
    det, counts, keep = ops().nms(pred_t, merge=False, want_keep_idx=True, **kw)
    n = int(counts[0])
    assert np.array_equal(keep[0, :n].cpu().numpy(), ref_keep)


def test_nms_full_size_batch_consistency():
    """BASELINE config 5: 196,608 rows per image.  Every image of a batch must equal the same
    image run alone (batch-independence), at both operating points; first image vs the oracle."""
    o = ops()
    pred = fx.synthetic_predictions(4, 196608, 8, 1024, 0.02, 31)
    pt = torch.from_numpy(pred).cuda()
    for kw in (dict(conf_thres=0.25, iou_thres=0.45), dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)):
        det, counts, keep = o.nms(pt, want_keep_idx=True, **kw)
        for i in range(4):
            d1, c1, k1 = o.nms(pt[i:i + 1], want_keep_idx=True, **kw)
            assert torch.equal(d1[0], det[i]) and torch.equal(c1[0], counts[i]) and torch.equal(k1[0], keep[i])
        outs, idxs = nms_ref.non_max_suppression(pred[:1], return_indices=True, early_stop=True, **kw)
        n = int(counts[0])
        assert n == len(idxs[0]) and np.array_equal(keep[0, :n].cpu().numpy(), idxs[0])


# ------------------------------------------------------------------------------------ whole model
def load_det_weights(model):
    sd = model.state_dict()
    new = {}
    for k, v in sd.items():
        if torch.is_floating_point(v) and not any(s in k for s in ("attn_mask", "anchors", "anchor_grid")):
            new[k] = fx.deterministic_tensor(k, v.shape, seed=0)
    model.load_state_dict(new, strict=False)
    return model


def test_model_512_fp32_vs_reference_golden(golden):
    from sodt_b200.basics.models.model import Model
    g = golden("model_512")
    m = load_det_weights(Model(YAML, input_mode="RGB+IR", ch_steam=3, ch=128, nc=8)).eval().cuda()
    rgb = fx.det_input("model:rgb", (1, 3, 512, 512), kind="uniform").cuda()
    ir = fx.det_input("model:ir", (1, 3, 512, 512), kind="uniform").cuda()
    with torch.no_grad():
        pred, raw, feats = m(rgb, ir, "RGB+IR")
    for i in range(3):
        assert rel_err(feats[i][0, ::7, ::5, ::3], g[f"feat{i}_sub"]) < 1e-4, i
    assert rel_err(raw[0][0].reshape(-1, 13)[::61], g["raw_rows"]) < 1e-4
    ref = torch.from_numpy(g["pred_rows"]).double()
    got = pred[0, ::61].double().cpu()
    assert ((got[:, :4] - ref[:, :4]).abs() / ref[:, :4].abs().clamp_min(1.0)).max() < 1e-3
    assert (got[:, 4:] - ref[:, 4:]).abs().max() < 1e-4


def test_model_512_bf16_vs_reference_golden(golden):
    from sodt_b200.basics.models.model import Model
    g = golden("model_512")
    m = load_det_weights(Model(YAML, input_mode="RGB+IR", ch_steam=3, ch=128, nc=8)).eval().cuda().to(torch.bfloat16)
    rgb = fx.det_input("model:rgb", (1, 3, 512, 512), kind="uniform").cuda().to(torch.bfloat16)
    ir = fx.det_input("model:ir", (1, 3, 512, 512), kind="uniform").cuda().to(torch.bfloat16)
    with torch.no_grad():
        pred, raw, feats = m(rgb, ir, "RGB+IR")
    assert pred.dtype == torch.float32            # decode is fp32 in every mode
    for i in range(3):
        assert rel_err(feats[i][0, ::7, ::5, ::3], g[f"feat{i}_sub"]) < 3e-2, i


def test_model_1024_runs_and_matches_oracle_adapter():
    """1024x1024 is beyond the reference's hard-coded grid; the oracle adapter (reference classes'
    math at a scaled token grid, oracle/model_ref.py) is the checker.  Batch 1 keeps the CPU oracle short."""
    from sodt_b200.basics.models.model import Model
    m = load_det_weights(Model(YAML, input_mode="RGB+IR", ch_steam=3, ch=128, nc=8)).eval()
    p = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda()
    rgb = fx.det_input("model1024:rgb", (1, 3, 1024, 1024), kind="uniform")
    ir = fx.det_input("model1024:ir", (1, 3, 1024, 1024), kind="uniform")
    with torch.no_grad():
        pred, raw, feats = m(rgb.cuda(), ir.cuda(), "RGB+IR")
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        pred_ref, raw_ref = model_ref.model_forward(rgb, ir, p)
    assert pred.shape == (1, 3 * 256 * 256, 13)
    assert rel_err(raw[0], raw_ref[0]) < 1e-4
    got, ref = pred.double().cpu(), pred_ref.double()
    assert ((got[..., :4] - ref[..., :4]).abs() / ref[..., :4].abs().clamp_min(1.0)).max() < 1e-3


def test_model_detector_graph_replay_matches_eager_and_stream():
    """runtime.Detector: the CUDA-graph replay of a step, the eager step and the pipelined host API give the same detections."""
    from sodt_b200.runtime import Detector
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randint(0, 256, (2, 3, 512, 512), dtype=torch.uint8, generator=g).pin_memory(),
                torch.randint(0, 256, (2, 3, 512, 512), dtype=torch.uint8, generator=g).pin_memory()) for _ in range(3)]
    eager = Detector(device="cuda", dtype=torch.bfloat16, seed=3, conf_thres=1e-6, cuda_graph=False)
    graph = Detector(device="cuda", dtype=torch.bfloat16, seed=3, conf_thres=1e-6, cuda_graph=True)
    want = [tuple(t.clone() for t in eager.detect(*b)) for b in batches]
    assert sum(int(c.sum()) for _, c in want) > 0
    for b, (d, c) in zip(batches, want):                      # replayed graph (second and third call reuse the capture)
        d2, c2 = graph.detect(*b)
        assert torch.equal(c2, c) and torch.equal(d2, d)
    assert graph.launches_per_step(batches[0][0].cuda(), batches[0][1].cuda()) >= 70      # 80 sodt kernels per step at this size (86 with SODT_FUSED_ATTN=0)
    got = [(d.clone(), c.clone()) for d, c in graph.detect_stream(iter(batches))]
    assert len(got) == len(want)
    for (d2, c2), (d, c) in zip(got, want):
        assert torch.equal(c2, c) and torch.equal(d2, d)
