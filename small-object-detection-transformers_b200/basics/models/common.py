"""YOLOv5 head modules used by models/model.yaml (reference basics/models/common.py).

Only the conv stack the detector's head needs (SURVEY.md section 8a lists it as a caller of the hot path, 8f as "next").
After Model.fuse() the bf16 convs run as tap GEMMs on the tcgen05 kernel with bias + SiLU in the epilogue
(ops.conv2d_nhwc, ops.upcat_conv1x1); otherwise they are cuDNN calls.  Module / parameter names follow the
reference so that state_dict keys (detect.N.cv1.conv.weight ...) match.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def autopad(k, p=None):
    """'same' padding for odd kernels."""
    if p is None:
        p = k // 2 if isinstance(k, int) else [v // 2 for v in k]
    return p


class Conv(nn.Module):
    """conv (no bias) -> BatchNorm -> SiLU (reference common.py:38)."""

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p), groups=g, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.act = nn.SiLU() if act is True else (act if isinstance(act, nn.Module) else nn.Identity())

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))

    def fuseforward(self, x, out=None):
        """BatchNorm already folded into the conv (Model.fuse).  ``out``: optional [B,H,W,c2] channel-slice view that
        receives the result (C3 writes its two branches into one buffer instead of concatenating them)."""
        conv = self.conv
        if (x.is_cuda and isinstance(self.act, nn.SiLU) and conv.bias is not None and conv.out_channels % 8 == 0
                and x.dtype in (torch.float32, torch.bfloat16)):
            from ... import ops
            kh, kw = conv.kernel_size
            xh = x.permute(0, 2, 3, 1)                            # [B,H,W,C] view of channels-last memory
            if (conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1 and conv.padding == (kh // 2, kw // 2)
                    and kh % 2 == 1 and kw % 2 == 1 and ops.conv2d_nhwc_supported(xh, conv.out_channels, kh, kw)):
                # conv + folded BN + SiLU as one tap GEMM on the tcgen05 kernel (1x1: a plain GEMM)
                y = ops.conv2d_nhwc(xh, ops.conv_weight_taps(conv.weight), conv.bias, (kh, kw), conv.padding, "silu", out=out)
                return y.permute(0, 3, 1, 2)
            y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
            y = ops.bias_act_crop(y, conv.bias, "silu")           # bias + SiLU in one pass
        else:
            y = self.act(conv(x)).permute(0, 2, 3, 1)
        if out is not None:
            out.copy_(y)
            y = out
        return y.permute(0, 3, 1, 2)


def DWConv(c1, c2, k=1, s=1, act=True):
    return Conv(c1, c2, k, s, g=math.gcd(c1, c2), act=act)


class Bottleneck(nn.Module):
    """1x1 -> 3x3 with optional residual (reference common.py:55)."""

    def __init__(self, c1, c2, shortcut=True, g=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_, c2, 3, 1, g=g)
        self.add = shortcut and c1 == c2

    def forward(self, x, out=None):
        fused = out is not None and not hasattr(self.cv2, "bn") and not self.add
        y = self.cv2.fuseforward(self.cv1(x), out=out) if fused else self.cv2(self.cv1(x))
        y = x + y if self.add else y
        if out is not None and not fused:
            out.copy_(y.permute(0, 2, 3, 1))
            y = out.permute(0, 3, 1, 2)
        return y


class C3(nn.Module):
    """CSP bottleneck with three convolutions (reference common.py:114)."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, e=1.0) for _ in range(n)))

    def _branch_weights(self):
        """cv1 and cv2 read the same input: one GEMM with the stacked [2c_, c1] weight (cached per parameter version)."""
        from ... import ops
        w1, w2, b1, b2 = self.cv1.conv.weight, self.cv2.conv.weight, self.cv1.conv.bias, self.cv2.conv.bias
        key = tuple((id(t), t._version, t.data_ptr()) for t in (w1, w2, b1, b2))
        if getattr(self, "_bw_key", None) != key:
            self._bw = (torch.cat((ops.conv_weight_taps(w1), ops.conv_weight_taps(w2))).contiguous(),
                        torch.cat((b1.detach().float(), b2.detach().float())).contiguous())
            self._bw_key = key
        return self._bw

    def forward(self, x):
        """``x``: NCHW-shaped tensor, or an ops.UpCat (upsample + concat not materialised: read through TMA addressing)."""
        from ... import ops
        lazy = x if isinstance(x, ops.UpCat) else None
        convs = (self.cv1, self.cv2, self.cv3)
        if (x.is_cuda and x.dtype == torch.bfloat16 and all(not hasattr(c, "bn") and isinstance(c.act, nn.SiLU)
                                                            and c.conv.kernel_size == (1, 1) and c.conv.bias is not None for c in convs)):
            c_ = self.cv1.conv.out_channels
            buf = None
            if lazy is not None and ops.upcat_conv1x1_supported(lazy, 2 * c_):
                w12, b12 = self._branch_weights()
                buf = ops.upcat_conv1x1(lazy, w12, b12, "silu")
            else:
                if lazy is not None:
                    x, lazy = lazy.materialize(), None
                xh = x.permute(0, 2, 3, 1)
                if ops.conv2d_nhwc_supported(xh, 2 * c_, 1, 1):
                    w12, b12 = self._branch_weights()
                    buf = ops.conv2d_nhwc(xh, w12, b12, (1, 1), (0, 0), "silu")
            if buf is not None:
                # both branches land in the halves of one [B,H,W,2c_] buffer: no torch.cat, the bottleneck chain's last
                # conv writes over the cv1 half it no longer needs
                half = buf[..., :c_]
                t = half.permute(0, 3, 1, 2)
                for i, m in enumerate(self.m):
                    t = m(t, out=half) if i == len(self.m) - 1 else m(t)
                return self.cv3(buf.permute(0, 3, 1, 2))
        if lazy is not None:
            x = lazy.materialize()
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), dim=1))


class SPP(nn.Module):
    """Spatial pyramid pooling (reference common.py:129)."""

    def __init__(self, c1, c2, k=(5, 9, 13)):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * (len(k) + 1), c2, 1, 1)
        self.m = nn.ModuleList(nn.MaxPool2d(kernel_size=v, stride=1, padding=v // 2) for v in k)

    def forward(self, x):
        x = self.cv1(x)
        return self.cv2(torch.cat([x] + [m(x) for m in self.m], 1))


class SE_Block(nn.Module):
    """Squeeze-and-excitation channel gate (reference common.py:165)."""

    def __init__(self, ch_in, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(ch_in, ch_in // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(ch_in // reduction, ch_in, bias=False), nn.Sigmoid())

    def forward(self, x):
        b, c = x.shape[:2]
        return x * self.fc(self.avg_pool(x).view(b, c)).view(b, c, 1, 1)


class MF(nn.Module):
    """SuperYOLO's multimodal fusion block (reference common.py:183-212): SE gates on RGB and IR, 1x1 "mask" convs, 3x3 convs to
    48 + 16 channels, SE gate on their concatenation.  No attention in it; forward([rgb, ir1]) -> 64 ch."""

    def __init__(self, channels):
        super().__init__()
        self.mask_map_r = nn.Conv2d(channels, 1, 1, 1, 0, bias=True)
        self.mask_map_i = nn.Conv2d(1, 1, 1, 1, 0, bias=True)
        self.softmax = nn.Softmax(-1)
        self.bottleneck1 = nn.Conv2d(1, 16, 3, 1, 1, bias=False)
        self.bottleneck2 = nn.Conv2d(channels, 48, 3, 1, 1, bias=False)
        self.se = SE_Block(64, 16)
        self.se_r = SE_Block(3, 3)
        self.se_i = SE_Block(1, 1)

    def forward(self, x):
        rgb0, ir0 = x[0], x[1]
        rgb, ir = self.se_r(rgb0), self.se_i(ir0)
        rgb_m = self.mask_map_r(rgb).repeat(1, 3, 1, 1) * rgb
        ir_m = self.mask_map_i(ir) * ir
        out_ir = self.bottleneck1(ir_m + ir0)
        out_rgb = self.bottleneck2(rgb_m + rgb0)
        return self.se(torch.cat([out_rgb, out_ir], 1))


class Focus(nn.Module):
    """Space-to-depth then conv (reference common.py:68)."""

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, act=True):
        super().__init__()
        self.conv = Conv(c1 * 4, c2, k, s, p, g, act)

    def forward(self, x):
        return self.conv(torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1))


class Concat(nn.Module):
    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, x):
        return torch.cat(x, self.d)


class NMS(nn.Module):
    """NMS as a module (reference common.py:285); runs the CUDA NMS."""
    conf = 0.25
    iou = 0.45
    classes = None

    def forward(self, x):
        from ..utils.general import non_max_suppression
        return non_max_suppression(x[0], conf_thres=self.conf, iou_thres=self.iou, classes=self.classes)
