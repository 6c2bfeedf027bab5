"""Generates tests/golden/*.npz and state_dict_512.json by running the UNMODIFIED reference.

Run once in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

The reference is imported from /root/reference through tests/golden/_refshim.py (three
stub modules, no source edits).  Inputs and weights come from oracle/fixtures.py
(pure functions of names and seeds), so the fixtures hold only reference OUTPUTS.
Nothing here runs on the GPU box.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.append(ROOT)

import _refshim  # noqa: E402

bv, sw, md, ge = _refshim.import_reference()
from oracle import fixtures as fx  # noqa: E402  (repo root is behind /root/reference on sys.path)

torch.manual_seed(0)
torch.set_grad_enabled(False)

# name -> (dim, (H, W), heads, ws, shift, linear_mlp, B)
SWIN_CASES = {
    "hd16_shift2_convmlp": (96, (16, 16), 6, 8, 2, False, 2),
    "hd16_shift0_linmlp": (96, (16, 16), 6, 8, 0, True, 2),
    "hd32_shift2_rect": (64, (24, 16), 2, 8, 2, False, 1),
    "pad_shift2": (48, (12, 12), 3, 8, 2, True, 2),
    "pad_shift0_rect": (48, (10, 12), 3, 8, 0, False, 1),
    "global_clamped": (128, (8, 8), 2, 32, 0, True, 2),
    "ws7_shift3": (48, (14, 14), 3, 7, 3, True, 1),
    "ws4_shift1_hd8": (32, (8, 12), 4, 4, 1, False, 2),
    "global_n256_hd64": (128, (16, 16), 2, 32, 0, True, 1),
}

# name -> (variant, C, heads, (h, w), ws, B)
CATTN_CASES = {
    "vit_ws1": ("vit", 48, 12, (16, 16), 1, 2),
    "vit_ws1_shift1": ("vit_shift1", 48, 12, (128, 128), 1, 1),  # mask buffer is built for 128x128 only (backbone_vit.py:439)
    "v2_ws2": ("v2", 24, 12, (16, 16), 2, 2),
    "v2_ws3_pad": ("v2", 48, 12, (16, 16), 3, 1),
    "v2_ws7_pad": ("v2", 48, 6, (16, 20), 7, 1),
    "v2_ws8": ("v2", 96, 4, (16, 16), 8, 1),
    "v2_ws4_h1": ("v2", 24, 1, (8, 8), 4, 2),
}


def np32(t):
    return t.detach().to(torch.float32).numpy()


def gen_swin():
    out = {}
    for name, (dim, res, heads, ws, shift, lin, B) in SWIN_CASES.items():
        blk = bv.SwinTransformerBlock(dim, res, heads, window_size=ws, shift_size=shift, linear_mlp=lin).eval()
        sd = fx.fill_state_dict(blk.state_dict(), seed=1)
        blk.load_state_dict(sd)
        x = fx.det_input("swin:" + name, (B, res[0] * res[1], dim))
        out[name + "/y"] = np32(blk(x))
        if blk.attn_mask is not None:
            out[name + "/attn_mask"] = np32(blk.attn_mask)
        out[name + "/rel_index"] = blk.attn.relative_position_index.numpy().astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "swin_blocks.npz"), **out)


def gen_cattn():
    out = {}
    for name, (variant, C, heads, hw, ws, B) in CATTN_CASES.items():
        streams = [fx.det_input(f"cattn:{name}:{i}", (B, hw[0], hw[1], C)) for i in range(4)]
        if variant.startswith("vit"):
            blk = bv.CAttentionBlock(C, heads, shift_size=1 if variant == "vit_shift1" else 0).eval()
        else:
            blk = sw.CAttentionBlock(C, heads).eval()
        blk.load_state_dict(fx.fill_state_dict(blk.state_dict(), seed=2))
        if variant.startswith("vit"):
            ys = blk(*streams)
        else:
            ys = torch.split(blk(*streams, window_size=ws), C, dim=-1)
        for i, y in enumerate(ys):
            out[f"{name}/y{i}"] = np32(y[:, ::5, ::3]) if hw[0] > 64 else np32(y)
    # masked CAttention (mask added before the 1/sqrt(c) scaling, backbone_vit.py:601-608)
    mask_src = bv.SwinTransformerBlock(48, (8, 8), 12, window_size=4, shift_size=2)
    q, k, v = (fx.det_input(f"cattn:masked:{i}", (2 * 4, 16, 48)) for i in range(3))
    out["masked/y"] = np32(bv.CAttention(48, 12)(q, k, v, (8, 8), mask=mask_src.attn_mask))
    np.savez_compressed(os.path.join(HERE, "cattn.npz"), **out)


def gen_detect():
    out = {}
    anchors = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119]]
    det = md.Detect(nc=8, anchors=anchors, ch=(16, 24)).eval()
    det.stride = torch.tensor([4.0, 8.0])
    det.load_state_dict(fx.fill_state_dict(det.state_dict(), seed=3))
    feats = [fx.det_input("detect:0", (2, 16, 6, 10)), fx.det_input("detect:1", (2, 24, 3, 5))]
    z, xs = det([f.clone() for f in feats])
    out["z"] = np32(z)
    for i, x in enumerate(xs):
        out[f"x{i}"] = np32(x)
    np.savez_compressed(os.path.join(HERE, "detect.npz"), **out)


NMS_CASES = {
    # name -> (B, R, img, active, seed, kwargs)
    "single_label": (3, 4096, 256, 0.2, 0, dict(conf_thres=0.25, iou_thres=0.45)),
    "multi_label": (2, 2048, 256, 0.1, 1, dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)),
    "agnostic": (2, 2048, 128, 0.2, 2, dict(conf_thres=0.25, iou_thres=0.45, agnostic=True)),
    "class_filter": (2, 4096, 256, 0.2, 3, dict(conf_thres=0.1, iou_thres=0.45, classes=[1, 5])),
    "dense_no_merge": (1, 16384, 192, 0.9, 4, dict(conf_thres=0.05, iou_thres=0.3)),
    "cap_30000": (1, 49152, 1024, 0.2, 5, dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)),
    "empty": (2, 512, 256, 0.0, 6, dict(conf_thres=0.25, iou_thres=0.45)),
    "apriori_labels": (3, 2048, 256, 0.2, 7, dict(conf_thres=0.25, iou_thres=0.45, labels_seed=9)),
}
import importlib.util  # noqa: E402
_gc = importlib.util.spec_from_file_location("_golden_cases", os.path.join(os.path.dirname(HERE), "golden_cases.py"))
_gcm = importlib.util.module_from_spec(_gc)
_gc.loader.exec_module(_gcm)
nms_kwargs = _gcm.nms_kwargs          # expands labels_seed into the reference's `labels` argument


def gen_nms():
    out = {}
    for name, (B, R, img, active, seed, kw) in NMS_CASES.items():
        pred = torch.from_numpy(fx.synthetic_predictions(B, R, 8, img, active, seed))
        kw = nms_kwargs(kw, B, img)
        if "labels" in kw:
            kw["labels"] = [torch.from_numpy(l) for l in kw["labels"]]
        dets = ge.non_max_suppression(pred.clone(), **kw)
        padded = np.zeros((B, 300, 6), dtype=np.float32)
        counts = np.zeros((B,), dtype=np.int32)
        for i, d in enumerate(dets):
            counts[i] = d.shape[0]
            padded[i, : d.shape[0]] = np32(d)
        out[name + "/det"] = padded
        out[name + "/count"] = counts
        print("nms", name, counts)
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **out)


def gen_model():
    m = md.Model("/root/reference/models/model.yaml", input_mode="RGB+IR", ch_steam=3, ch=128, nc=8).eval()
    sd0 = m.state_dict()
    meta = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd0.items()}
    with open(os.path.join(HERE, "state_dict_512.json"), "w") as f:
        json.dump(meta, f, indent=0, sort_keys=True)
    m.load_state_dict(fx.fill_state_dict(sd0, seed=0))
    rgb = fx.det_input("model:rgb", (1, 3, 512, 512), kind="uniform")
    ir = fx.det_input("model:ir", (1, 3, 512, 512), kind="uniform")
    feats = {}
    m.image_encoder.register_forward_hook(lambda mod, i, o: feats.update(y=[t.clone() for t in o]))
    pred, raw, _ = m(rgb, ir, "RGB+IR")
    out = {"pred_rows": np32(pred[0, ::61]), "raw_rows": np32(raw[0].reshape(-1, 13)[::61])}
    for i, t in enumerate(feats["y"]):
        out[f"feat{i}_sub"] = np32(t[0, ::7, ::5, ::3])
        out[f"feat{i}_stats"] = np.array([t.mean().item(), t.std().item(), t.abs().max().item()], dtype=np.float64)
    out["pred_stats"] = np.array([pred.mean().item(), pred.std().item(), pred.abs().max().item()], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "model_512.npz"), **out)
    print("model pred stats", out["pred_stats"])


# Round 2: block-level cases at the detector's own widths (they reach the tcgen05 window / flash / GEMM kernels in bf16)
# name -> (dim, (H, W), heads, ws, shift, linear_mlp, B)
SWIN_BIG_CASES = {
    "s1_dim192_shift0_lin": (192, (16, 16), 12, 8, 0, True, 2),
    "s1_dim192_shift2_conv": (192, (16, 16), 12, 8, 2, False, 2),
    "s2_dim384_shift2_conv": (384, (16, 16), 12, 8, 2, False, 1),
    "s2_dim384_shift0_lin": (384, (16, 8), 12, 8, 0, True, 2),
    "s3_dim768_global": (768, (32, 32), 12, 32, 0, True, 1),
}


def gen_swin_big():
    out = {}
    for name, (dim, res, heads, ws, shift, lin, B) in SWIN_BIG_CASES.items():
        blk = bv.SwinTransformerBlock(dim, res, heads, window_size=ws, shift_size=shift, linear_mlp=lin).eval()
        blk.load_state_dict(fx.fill_state_dict(blk.state_dict(), seed=1))
        x = fx.det_input("swin:" + name, (B, res[0] * res[1], dim))
        y = blk(x)
        out[name + "/y"] = np32(y[:, ::3])                      # every third token keeps the fixture small
    np.savez_compressed(os.path.join(HERE, "swin_blocks_big.npz"), **out)


def fill_parameters(module, seed):
    sd = module.state_dict()
    for k, v in module.named_parameters():
        sd[k] = fx.deterministic_tensor(k, v.shape, seed)
    module.load_state_dict(sd)


def gen_variants():
    """Attention variants SURVEY.md section 8(f) rank 4 names: the reference defines them (and they run standalone) although its
    detector never instantiates them; and SuperYOLO's MF fusion block."""
    import basics.models.common as cm
    out = {}
    # SwinV2 cosine attention (backbone_swinv2.py:837-949), with and without a shift mask
    for name, (dim, ws, heads, B_, masked) in {"v2attn_ws8": (96, 8, 6, 4, False), "v2attn_ws4_mask": (48, 4, 3, 8, True),
                                                "v2attn_ws7": (64, 7, 2, 3, False)}.items():
        m = sw.WindowAttention(dim, (ws, ws), heads).eval()
        fill_parameters(m, seed=5)                                # buffers (coordinate table, index) stay as constructed
        x = fx.det_input("v2attn:" + name, (B_, ws * ws, dim))
        mask = None
        if masked:
            mask = bv.SwinTransformerBlock(dim, (2 * ws, 2 * ws), heads, window_size=ws, shift_size=ws // 2).attn_mask
        out[name + "/y"] = np32(m(x, mask))
    # SAM-style global attention with decomposed relative position embeddings (backbone_vit.py:347-404,705-740)
    for name, (dim, heads, S, B, rel) in {"sam_rel_16": (64, 4, 16, 2, True), "sam_norel_8": (48, 3, 8, 2, False),
                                           "sam_rel_32": (128, 2, 32, 1, True)}.items():
        m = bv.Attention(dim, heads, qkv_bias=True, use_rel_pos=rel, input_size=(S, S)).eval()
        m.load_state_dict(fx.fill_state_dict(m.state_dict(), seed=6))
        x = fx.det_input("sam:" + name, (B, S, S, dim))
        out[name + "/y"] = np32(m(x))
    # MF fusion block (common.py:183-212)
    mf = cm.MF(3).eval()
    mf.load_state_dict(fx.fill_state_dict(mf.state_dict(), seed=7))
    rgb = fx.det_input("mf:rgb", (2, 3, 24, 32), kind="uniform")
    ir = fx.det_input("mf:ir", (2, 1, 24, 32), kind="uniform")
    out["mf/y"] = np32(mf([rgb, ir]))
    np.savez_compressed(os.path.join(HERE, "variants.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["swin", "cattn", "detect", "nms", "model", "swin_big", "variants"]
    for w in which:
        globals()["gen_" + w]()
        print("generated", w)
