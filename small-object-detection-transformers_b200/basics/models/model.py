"""Model assembly with the reference's surface: Detect, Model, parse_model
(reference basics/models/model.py).  The yaml schema ([from, number, module, args] rows,
depth / width multiples, anchors) and the resulting state_dict keys are unchanged.

Detect's 1x1 conv and decode run in sm_100a kernels (ops.conv2d_nhwc, ops.detect_decode); the backbone and the head
call the kernels described in backbone_vit.py / common.py; nn.Upsample + Concat are folded into the next C3 (ops.UpCat).
"""
import logging
import math
from copy import deepcopy
from pathlib import Path

import torch
import torch.nn as nn

from ... import ops
from ..utils.general import make_divisible
from .backbone_vit import ImageEncoderViT
from .common import C3, MF, SPP, Bottleneck, Concat, Conv, DWConv, Focus, SE_Block

logger = logging.getLogger(__name__)


class Detect(nn.Module):
    """YOLOv5 detection layer (reference model.py:32-70)."""
    stride = None   # set by Model
    export = False
    want_raw = True   # eval mode returns (z, x) like the reference; runtime.Detector only needs z and turns the second output off

    def __init__(self, nc=80, anchors=(), ch=()):
        super().__init__()
        self.nc = nc
        self.no = nc + 5
        self.nl = len(anchors)
        self.na = len(anchors[0]) // 2
        self.grid = [torch.zeros(1)] * self.nl   # kept for attribute compatibility; the kernel needs no grid tensor
        a = torch.tensor(anchors).float().view(self.nl, -1, 2)
        self.register_buffer("anchors", a)
        self.register_buffer("anchor_grid", a.clone().view(self.nl, 1, -1, 1, 1, 2))
        self.m = nn.ModuleList(nn.Conv2d(c, self.no * self.na, 1) for c in ch)

    def _level_conv(self, i, xi):
        """The level's 1x1 conv (reference model.py:53).  Inference on bf16: the tcgen05 GEMM with the na*no = 39 output
        channels zero-padded to 64 (the cuBLAS path falls back to a pre-tensor-core kernel for N = 39); the decode kernel
        reads the first na*no channels through strides."""
        conv = self.m[i]
        ch = self.no * self.na
        if not self.training and xi.is_cuda and xi.dtype == torch.bfloat16 and conv.kernel_size == (1, 1):
            npad = (ch + 63) // 64 * 64
            xh = xi.permute(0, 2, 3, 1)
            if ops.conv2d_nhwc_supported(xh, npad, 1, 1):
                w = ops.cached_derived(conv.weight, "pad64", lambda t: torch.cat(
                    (t.detach().reshape(ch, -1), t.new_zeros(npad - ch, t.shape[1]))).contiguous())
                b = ops.cached_derived(conv.bias, "pad64", lambda t: torch.cat(
                    (t.detach().float(), torch.zeros(npad - ch, dtype=torch.float32, device=t.device))).contiguous())
                y = ops.conv2d_nhwc(xh, w, b, (1, 1), (0, 0), None)
                return y[..., :ch].permute(0, 3, 1, 2)
        return conv(xi)

    def forward(self, x):
        self.training |= self.export
        raws = [self._level_conv(i, x[i]) for i in range(self.nl)]
        if self.training:
            for i, r in enumerate(raws):
                bs, _, ny, nx = r.shape
                x[i] = r.view(bs, self.na, self.no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
            return x
        bs = raws[0].shape[0]
        rows = [self.na * r.shape[2] * r.shape[3] for r in raws]
        z = torch.empty((bs, sum(rows), self.no), dtype=torch.float32, device=raws[0].device)
        off = 0
        for i, r in enumerate(raws):
            anchors = ops.cached_derived(self.anchor_grid, ("level", i), lambda t, i=i: t[i].detach().float().reshape(-1, 2).contiguous())
            _, x[i] = ops.detect_decode(r, anchors, float(self.stride[i]), want_perm=self.want_raw, z=z,
                                        rows_total=z.shape[1], row_offset=off)
            off += rows[i]
        return z, (x if self.want_raw else None)


_MODULES = {
    "Conv": Conv, "DWConv": DWConv, "Bottleneck": Bottleneck, "C3": C3, "SPP": SPP, "Focus": Focus,
    "Concat": Concat, "Detect": Detect, "ImageEncoderViT": ImageEncoderViT, "MF": MF, "SE_Block": SE_Block,
    "nn.Upsample": nn.Upsample, "nn.BatchNorm2d": nn.BatchNorm2d,
}
_WIDTH_SCALED = (Conv, Bottleneck, SPP, DWConv, Focus, C3)


def _resolve(token, names):
    """yaml args are strings like 'nc', 'anchors', 'None', 'nearest'."""
    if not isinstance(token, str):
        return token
    if token in names:
        return names[token]
    if token in ("None", "True", "False"):
        return {"None": None, "True": True, "False": False}[token]
    return token


def parse_model(d, string, ch, config=None):
    """Builds the 'backbone' module or the 'head' nn.Sequential from a model dict
    (reference model.py:350-435).  ``ch`` is the running list of channel counts."""
    anchors, nc, gd, gw = d["anchors"], d["nc"], d["depth_multiple"], d["width_multiple"]
    na = len(anchors[0]) // 2 if isinstance(anchors, list) else anchors
    no = na * (nc + 5)
    names = {"nc": nc, "anchors": anchors}
    parts = string.split("+")
    rows = sum((d[p] for p in parts), []) if len(parts) == 2 else d[parts[-1]]
    # 'backbone+head' = the upstream SuperYOLO / YOLOv5 layer list (SRyolo_MF.yaml, SRyolo_PF.yaml): ch = [input channels, out of
    # layer 0, out of layer 1, ...], so a `from` index j >= 0 reads ch[j + 1].  The reference dropped that + 1 when it re-seeded
    # ch for its ViT head (model.py:410: "changed removed + 1"), which is why these configs fail there.
    legacy = len(parts) == 2
    chan = (lambda j: ch[j if j < 0 else j + 1]) if legacy else (lambda j: ch[j])
    if string == "head":   # the three backbone outputs seed the head's channel list (reference model.py:367-370)
        ch[0] = 256
        ch += [256, 512]
    layers, save, c2 = [], [], ch[-1]
    for i, (f, n, m, args) in enumerate(rows):
        if isinstance(m, str):
            if m not in _MODULES:
                raise KeyError(f"module {m!r} is outside the scope of this package")
            m = _MODULES[m]
        args = [_resolve(a, names) for a in args]
        n = max(round(n * gd), 1) if n > 1 else n
        if string == "backbone":
            mod = m(img_size=args[0], patch_size=4, embed_dim=args[2], in_chans=args[3], out_chans=args[4],
                    window_size=args[5])
        else:
            if m in _WIDTH_SCALED or m is DWConv:
                c1, c2 = ch[f], args[0]
                c2 = make_divisible(c2 * gw, 8) if c2 != no else c2
                args = [c1, c2, *args[1:]]
                if m is C3:
                    args.insert(2, n)
                    n = 1
            elif m is nn.BatchNorm2d:
                args = [ch[f]]
            elif m is Concat:
                c2 = sum(chan(j) for j in f)
            elif m is Detect:
                args.append([chan(j) for j in f])
                if isinstance(args[1], int):
                    args[1] = [list(range(args[1] * 2))] * len(f)
            else:
                c2 = ch[f if f < 0 else f + 1]
            mod = nn.Sequential(*(m(*args) for _ in range(n))) if n > 1 else m(*args)
        mod.i, mod.f = i, f
        mod.type = f"{m.__module__}.{m.__name__}" if hasattr(m, "__name__") else str(m)
        mod.np = sum(p.numel() for p in mod.parameters())
        save.extend(j % (i + 0.00001) for j in ([f] if isinstance(f, int) else f) if j != -1)
        layers.append(mod)
        ch.append(c2)
    if string == "backbone":
        return layers[0], sorted(save)
    return nn.Sequential(*layers), sorted(save)


def fuse_conv_and_bn(conv, bn):
    """Folds an eval-mode BatchNorm into the preceding conv (reference utils/torch_utils.py:182)."""
    fused = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding,
                      groups=conv.groups, bias=True).requires_grad_(False).to(conv.weight.device, conv.weight.dtype)
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    fused.weight.copy_(conv.weight * scale.view(-1, 1, 1, 1))
    b = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
    fused.bias.copy_((b - bn.running_mean) * scale + bn.bias)
    return fused


class Model(nn.Module):
    """The RGB+IR detector: ImageEncoderViT backbone + YOLOv5 head (reference model.py:73-348)."""
    export = False

    def __init__(self, cfg="yolov5s.yaml", input_mode="RGB", ch_steam=3, ch=3, nc=None, anchors=None, config=None,
                 sr=False, factor=2):
        super().__init__()
        self.init_params = dict(cfg=cfg, input_mode=input_mode, ch_steam=ch_steam, ch=ch, nc=nc, anchors=anchors,
                                config=config, sr=sr, factor=factor)
        if isinstance(cfg, dict):
            self.yaml = cfg
        else:
            import yaml
            self.yaml_file = Path(cfg).name
            with open(cfg) as f:
                self.yaml = yaml.safe_load(f)
        if sr:
            raise NotImplementedError("the super-resolution training branch is out of scope")
        self.sr = False
        ch = self.yaml["ch"] = self.yaml.get("ch", ch)
        if nc and nc != self.yaml["nc"]:
            self.yaml["nc"] = nc
        if anchors:
            self.yaml["anchors"] = round(anchors)
        # Two families of configs share the schema.  models/model.yaml (the detector of the paper, the only config the reference can
        # run): an ImageEncoderViT backbone row + a head whose `from` indices address [y0, y1, y2, head layers...].
        # SRyolo_MF.yaml / SRyolo_PF.yaml (SuperYOLO leftovers, named by BASELINE.json): one 'backbone+head' layer list with the
        # upstream index semantics; the reference fails on them, so this path has no oracle.
        self.legacy = self.yaml["backbone"][0][2] != "ImageEncoderViT"
        if self.legacy:
            self.model, self.save = parse_model(deepcopy(self.yaml), "backbone+head", ch=[ch], config=config)
            det = self.model[-1]
        else:
            self.image_encoder, self.save1 = parse_model(deepcopy(self.yaml), "backbone", ch=[ch], config=config)
            self.detect, self.save2 = parse_model(deepcopy(self.yaml), "head", ch=[ch], config=config)
            det = self.detect[-1]
        if isinstance(det, Detect):
            if self.legacy:      # upstream: strides from a dummy forward at 256 px (training mode returns the raw level maps)
                s = 256
                with torch.no_grad():
                    raw = self.forward(torch.zeros(1, ch_steam, s, s), torch.zeros(1, ch_steam, s, s), input_mode)[0]
                det.stride = torch.tensor([s / r.shape[-2] for r in raw])
            else:
                det.stride = torch.tensor([4.0])         # hard-coded in the reference (model.py:130)
            det.anchors /= det.stride.view(-1, 1, 1)
            area = det.anchor_grid.prod(-1).view(-1)     # check_anchor_order (utils/autoanchor.py:13)
            if (area[-1] - area[0]).sign() != (det.stride[-1] - det.stride[0]).sign():
                det.anchors[:] = det.anchors.flip(0)
                det.anchor_grid[:] = det.anchor_grid.flip(0)
            self.stride = det.stride
            self._initialize_biases()
        for m in self.modules():                         # initialize_weights (utils/torch_utils.py:145)
            if type(m) is nn.BatchNorm2d:
                m.eps, m.momentum = 1e-3, 0.03

    def _detect_layer(self):
        return self.model[-1] if self.legacy else self.detect[-1]

    def _initialize_biases(self, cf=None):
        det = self._detect_layer()
        with torch.no_grad():
            for conv, s in zip(det.m, det.stride):
                b = conv.bias.view(det.na, -1)        # in-place view: objectness and class priors (arXiv 1708.02002, 3.3)
                b[:, 4] += math.log(8 / (640 / float(s)) ** 2)
                b[:, 5:] += math.log(0.6 / (det.nc - 0.99)) if cf is None else torch.log(cf / cf.sum())

    def _stem(self, x, ir, input_mode):
        if input_mode == "RGB+IR":
            return torch.cat((x, ir[:, 0:1]), 1)
        if input_mode == "RGB":
            return x
        if input_mode == "IR":
            return ir
        if input_mode == "RGB+IR+MF":                    # the MF block takes the pair (reference model.py:190)
            return [x, ir[:, 0:1]]
        raise NotImplementedError(f"input_mode {input_mode!r} is out of scope")

    def forward(self, x, ir=None, input_mode="RGB+IR", augment=False, profile=False):
        if augment:
            raise NotImplementedError("test-time augmentation is out of scope")
        if ir is None:
            ir = x
        if x.dtype == torch.uint8:
            # extension: the evaluation loop's uint8 images (basics/test.py:124-130); /255 happens inside the front-end kernel
            if self.legacy or input_mode != "RGB+IR" or ir.dtype != torch.uint8:
                raise NotImplementedError("uint8 input is supported for input_mode='RGB+IR' with uint8 rgb and ir")
            stem = (x, ir)
        else:
            stem = self._stem(x, ir, input_mode)
        head_out, feats = self.forward_once(stem, "yolo", profile)
        self.training |= self.export
        if self.training:
            return head_out, feats
        return head_out[0], head_out[1], feats

    def _forward_legacy(self, x):
        """Upstream layer loop over the 'backbone+head' list: `from` -1 = previous output, j >= 0 = saved output of layer j."""
        y = []
        for m in self.model:
            if m.f != -1:
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            x = m(x)
            y.append(x if m.i in self.save else None)
        return x, y

    def forward_once(self, x, string="yolo", profile=False):
        if self.legacy:
            return self._forward_legacy(x)
        y = list(self.image_encoder(x[0], ir_u8=x[1]) if isinstance(x, tuple) else self.image_encoder(x))
        x = y[-1]
        mods = list(self.detect)
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if (isinstance(m, nn.Upsample) and m.f == -1 and isinstance(nxt, Concat) and nxt.d == 1 and m.mode == "nearest"
                    and m.scale_factor in (2, 2.0) and isinstance(nxt.f, list) and len(nxt.f) == 2 and nxt.f[0] == -1
                    and x.is_cuda and (x.shape[1] * x.element_size()) % 16 == 0
                    and (y[nxt.f[1]].shape[1] * x.element_size()) % 16 == 0):
                # nn.Upsample(2, nearest) + Concat([-1, k]) are not executed: the following C3 reads the two sources through TMA
                # addressing (ops.UpCat); any other consumer materialises the concatenation in one fused pass
                x = ops.UpCat(x, y[nxt.f[1]])
                y.append(None)
                y.append(x)
                i += 2
                continue
            if isinstance(x, ops.UpCat) and not isinstance(m, C3):
                x = x.materialize()
            if m.f != -1:
                for j in ([m.f] if isinstance(m.f, int) else m.f):          # a later consumer of a lazy concatenation
                    if j != -1 and isinstance(y[j], ops.UpCat):
                        y[j] = y[j].materialize()
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            x = m(x)
            y.append(x)
            i += 1
        return x, y

    @torch.no_grad()
    def fuse(self):
        for m in self.modules():
            if type(m) is Conv and hasattr(m, "bn"):
                m.conv = fuse_conv_and_bn(m.conv, m.bn)
                delattr(m, "bn")
                m.forward = m.fuseforward
        return self

    def info(self, verbose=False, img_size=640):
        n = sum(p.numel() for p in self.parameters())
        logger.info("Model: %d layers, %d parameters", len(list(self.modules())), n)
