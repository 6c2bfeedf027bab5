"""SwinV2 attention variant of the reference (basics/models/backbone_swinv2.py): cosine window attention with a learned
per-head logit scale and a continuous relative position bias (log-spaced coordinates -> 2-layer MLP -> 16 sigmoid).

Only the attention module is mirrored: same constructor, parameter / buffer names and shapes as
the reference's ``WindowAttention`` (:837-949), so its state_dicts load.  The score epilogue runs inside the exact attention
kernel (``ops.window_attention_ex``: L2-normalised q / k, per-head scale, optional dense mask); the bias table
16 * sigmoid(cpb_mlp(relative_coords_table)) depends on parameters only and is evaluated once per weight version.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from .backbone_vit import _relative_position_index, to_2tuple


class WindowAttention(nn.Module):
    """Cosine window attention (reference backbone_swinv2.py:837).  forward(x [num_windows*B, N, C], mask [nW, N, N] | None)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, attn_drop=0., proj_drop=0., pretrained_window_size=(0, 0)):
        super().__init__()
        if attn_drop or proj_drop:
            raise NotImplementedError("inference path: dropout must be 0")
        self.dim = dim
        self.window_size = to_2tuple(window_size)
        self.pretrained_window_size = tuple(pretrained_window_size)
        self.num_heads = num_heads
        self.logit_scale = nn.Parameter(torch.log(10 * torch.ones((num_heads, 1, 1))))
        self.cpb_mlp = nn.Sequential(nn.Linear(2, 512, bias=True), nn.ReLU(inplace=True), nn.Linear(512, num_heads, bias=False))
        wh, ww = self.window_size
        ch = torch.arange(-(wh - 1), wh, dtype=torch.float32)
        cw = torch.arange(-(ww - 1), ww, dtype=torch.float32)
        table = torch.stack(torch.meshgrid([ch, cw], indexing="ij")).permute(1, 2, 0).contiguous().unsqueeze(0)   # [1, 2Wh-1, 2Ww-1, 2]
        ph, pw = self.pretrained_window_size
        table[..., 0] /= (ph - 1) if ph > 0 else (wh - 1)
        table[..., 1] /= (pw - 1) if ph > 0 else (ww - 1)
        table *= 8                                                                                       # normalise to [-8, 8]
        table = torch.sign(table) * torch.log2(torch.abs(table) + 1.0) / math.log2(8)
        self.register_buffer("relative_coords_table", table)
        self.register_buffer("relative_position_index", _relative_position_index(wh, ww))
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(dim))
            self.v_bias = nn.Parameter(torch.zeros(dim))
        else:
            self.q_bias = None
            self.v_bias = None
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.softmax = nn.Softmax(dim=-1)

    def bias_table(self):
        """[(2Wh-1)(2Ww-1), heads] fp32 = 16 sigmoid(cpb_mlp(coords)): the table form of the reference's gathered bias (:913-918)."""
        srcs = [p for p in self.cpb_mlp.parameters()]
        key = tuple((id(t), t._version, t.data_ptr()) for t in srcs)
        if getattr(self, "_bt_key", None) != key:
            with torch.no_grad():
                t = self.cpb_mlp(self.relative_coords_table.to(srcs[0].dtype)).view(-1, self.num_heads)
                self._bt = (16 * torch.sigmoid(t.float())).contiguous()
            self._bt_key = key
        return self._bt

    def forward(self, x, mask=None):
        wh, ww = self.window_size
        if wh != ww or not x.is_cuda:
            raise NotImplementedError("square windows on CUDA tensors")
        B_, N, C = x.shape
        bias = None
        if self.q_bias is not None:
            bias = torch.cat((self.q_bias, torch.zeros_like(self.v_bias), self.v_bias))
        qkv = F.linear(x, self.qkv.weight, bias).view(B_, wh, ww, 3 * C)
        head_scale = torch.clamp(self.logit_scale.detach().float(), max=math.log(1.0 / 0.01)).exp().view(-1)
        o = ops.window_attention_ex(qkv, self.bias_table(), self.num_heads, wh, 0, dense_mask=mask, head_scale=head_scale,
                                    normalize_qk=True)
        return self.proj(o.view(B_, N, C))
