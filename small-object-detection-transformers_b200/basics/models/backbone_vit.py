"""Host-side mirror of the reference backbone (basics/models/backbone_vit.py) for the hot path.

Same class names, constructor arguments, parameter / buffer names and shapes as the reference,
so its state_dicts and pickled checkpoints resolve here.  The math does not run in torch: in bf16
every layer of the backbone calls the sm_100a kernels through the C ABI (``ops.window_attention``,
``ops.linear`` with LayerNorm / GELU / residual folded into the GEMM, ``ops.conv2d_nhwc`` for the
conv-MLP's taps, ``ops.patch_merge_linear``, ``ops.frontend_u8``); fp32 (the 1e-5 exactness mode) uses
the exact attention / LayerNorm / front-end kernels with cuBLAS / cuDNN for the GEMMs and convs.

Differences from the reference, all deliberate:
  * the token grid follows the input instead of the hard-coded 128x128
    (reference backbone_vit.py:119,136,153,1087), so 1024^2 and 2048^2 inputs run;
  * no window / rolled / score / bias / mask tensor is ever materialised;
  * activations stay channels-last ([B,H,W,C]) end to end; 1x1 convs run as GEMMs on that layout.
"""
import math
from functools import partial
from typing import Optional, Tuple, Type

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops

MASK_VALUE = -100.0


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


# ------------------------------------------------------------------ reference-compatible helpers
def window_partition(x: torch.Tensor, window_size: int, top_padding: bool = False):
    """[B,H,W,C] -> ([B*nW, ws, ws, C], (Hp, Wp)); API of reference backbone_vit.py:619.
    Kept for callers of the reference API; the fused kernels never call it."""
    B, H, W, C = x.shape
    ph, pw = (-H) % window_size, (-W) % window_size
    if ph or pw:
        x = F.pad(x, (0, 0, pw, 0, ph, 0) if top_padding else (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.reshape(B, Hp // window_size, window_size, Wp // window_size, window_size, C)
    return x.transpose(2, 3).reshape(-1, window_size, window_size, C), (Hp, Wp)


def window_unpartition(windows: torch.Tensor, window_size: int, pad_hw: Tuple[int, int], hw: Tuple[int, int],
                       top_padding: bool = False):
    """Inverse of window_partition + crop; API of reference backbone_vit.py:646."""
    Hp, Wp = pad_hw
    H, W = hw
    nh, nw = Hp // window_size, Wp // window_size
    B = windows.shape[0] // (nh * nw)
    x = windows.reshape(B, nh, nw, window_size, window_size, -1).transpose(2, 3).reshape(B, Hp, Wp, -1)
    if Hp > H or Wp > W:
        x = x[:, Hp - H:, Wp - W:] if top_padding else x[:, :H, :W]
    return x.contiguous()


def get_channels(x):
    """Four single-channel views of a [B,4,H,W] tensor (reference backbone_vit.py:810)."""
    return tuple(x[:, c:c + 1] for c in range(4))


def _relative_position_index(wh: int, ww: int) -> torch.Tensor:
    ys = torch.arange(wh).repeat_interleave(ww)
    xs = torch.arange(ww).repeat(wh)
    return (ys[:, None] - ys[None, :] + wh - 1) * (2 * ww - 1) + (xs[:, None] - xs[None, :] + ww - 1)


def _shift_mask(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """The reference's attn_mask buffer ([nW, N, N], 0 / -100) in closed form.  Registered only so
    that state_dict keys and shapes match; the kernels evaluate the same region ids on the fly."""
    def axis(n):
        a = torch.arange(n)
        return (a >= n - ws).long() + (a >= n - shift).long()
    ids = (3 * axis(H)[:, None] + axis(W)[None, :]).float().reshape(1, H, W, 1)
    win, _ = window_partition(ids, ws)
    flat = win.reshape(-1, ws * ws)
    diff = flat[:, None, :] - flat[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, MASK_VALUE), torch.zeros_like(diff))


# --------------------------------------------------------------------------------- small layers
class PatchEmbed(nn.Module):
    """Conv patch embedding returning channels-last tokens (reference backbone_vit.py:742).
    Note the reference's default padding of (1, 1), which channel_embed_r inherits."""

    def __init__(self, kernel_size=(16, 16), stride=(16, 16), padding=(1, 1), in_chans: int = 3, embed_dim: int = 768):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=padding)

    def forward(self, x):
        return self.proj(x).permute(0, 2, 3, 1)

    def forward_tokens(self, x, pos=None, want_stats=False):
        """1x1 / stride-1 projection applied directly on [B,H,W,Cin] tokens as a GEMM; ``pos`` [1,H,W,C] (the absolute
        position embedding, reference backbone_vit.py:212-214) is added in the GEMM epilogue, broadcast over the batch."""
        w = self.proj.weight
        w = w.reshape(w.shape[0], w.shape[1])
        if x.is_cuda and x.dtype == torch.bfloat16 and (pos is None or (tuple(pos.shape[1:]) == tuple(x.shape[1:3]) + (w.shape[0],)
                                                                        and (pos.shape[1] * pos.shape[2]) % 128 == 0)):
            if want_stats and ops.linear_ln_supported(x, w.shape[0]):
                return ops.linear(x, w, self.proj.bias, residual=pos, want_stats=True)    # row statistics for the first norm1
            y = ops.linear(x, w, self.proj.bias, residual=pos)
            return (y, None) if want_stats else y
        y = F.linear(x, w, self.proj.bias)
        y = y if pos is None else y + pos
        return (y, None) if want_stats else y


class PatchMerging(nn.Module):
    """2x2 neighbourhood gather + Linear(4C -> 2C) + LayerNorm (reference backbone_vit.py:823)."""

    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(2 * dim)

    def forward(self, x, input_resolution):
        H, W = input_resolution
        B, L, C = x.shape
        if L != H * W or H % 2 or W % 2:
            raise ValueError(f"PatchMerging: bad token grid {H}x{W} for {L} tokens")
        # gather in the channel order of the reference concat, (dy,dx) = (0,0), (1,0), (0,1), (1,1), + Linear(4C -> 2C):
        # the gather is the TMA addressing of the GEMM's A operand, the concatenated tensor is never written
        y = ops.patch_merge_linear(x.view(B, H, W, C), self.reduction.weight, self.reduction.bias)
        return ops.add_layernorm(y, None, self.norm.weight, self.norm.bias, self.norm.eps)[0]


class Mlp(nn.Module):
    """Linear MLP, or the conv-enhanced variant fc1 -> 2x2 conv -> GELU -> fc2
    (reference backbone_vit.py:863).  bf16: tcgen05 GEMMs (the 2x2 conv as a tap GEMM with TMA-addressed taps)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, linear_mlp=True, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.linear = linear_mlp
        self.bs = in_features
        if linear_mlp:
            self.fc1 = nn.Linear(in_features, hidden_features)
            self.act = act_layer()
            self.fc2 = nn.Linear(hidden_features, out_features)
        else:
            self.fc1 = nn.Linear(in_features, in_features)
            self.act = act_layer()
            self.conv1 = nn.Conv2d(in_features, in_features, 2)
            self.fc2 = nn.Linear(in_features, out_features)
        self.drop = nn.Dropout(drop)

    def hidden(self, x, H, W, ln=None):
        """Everything before fc2: fc1 -> GELU, or fc1 -> zero-pad -> 2x2 conv -> GELU.  [B, L, hidden].
        ``ln=(norm, mean_rstd)``: ``x`` is the un-normalised residual stream and ``norm`` (the block's norm2) is folded into fc1."""
        exact_gelu = isinstance(self.act, nn.GELU) and self.act.approximate == "none"
        w1, b1, fold = self.fc1.weight, self.fc1.bias, None
        if ln is not None:
            norm, stats = ln
            fold = (stats, norm.weight, norm.bias, norm.eps)
        if self.linear:
            if exact_gelu and x.is_cuda:
                return ops.linear(x, w1, b1, act="gelu", ln=fold)    # GELU fused into the GEMM epilogue (bf16)
            if fold is not None:                                     # any other activation: norm2 still folds into fc1
                return self.act(ops.linear(x, w1, b1, ln=fold))
            return self.act(self.fc1(x))
        B, L, C = x.shape
        h = ops.linear(x, w1, b1, ln=fold) if x.is_cuda else self.fc1(x)
        h4 = h.view(B, H, W, C)
        conv = self.conv1
        if exact_gelu and conv.kernel_size == (2, 2) and ops.conv2d_nhwc_supported(h4, C, 2, 2):
            # zero-pad right / below + valid 2x2 conv + bias + GELU as one tap GEMM: the four taps are TMA boxes of the
            # fc1 output shifted by (ky, kx); rows / columns past the image come from the TMA out-of-bounds fill
            return ops.conv2d_nhwc(h4, ops.conv_weight_taps(conv.weight), conv.bias, (2, 2), (0, 0), "gelu").view(B, L, C)
        h = h4.permute(0, 3, 1, 2)                             # NCHW view of channels-last memory
        if isinstance(self.act, nn.GELU) and self.act.approximate == "none" and C % 8 == 0:
            # pad(0,1,0,1) + valid 2x2 conv == rows/cols 1.. of the same conv with symmetric padding 1; the crop, the
            # conv bias and the GELU run in one sm_100a pass instead of a pad copy, a bias-add pass and a GELU pass
            o = F.conv2d(h, conv.weight, None, 1, 1)
            return ops.bias_act_crop(o, conv.bias, "gelu", (H, W), (1, 1)).view(B, L, C)
        h = F.pad(h, (0, 1, 0, 1))                              # zero column right, zero row below
        h = conv(h).permute(0, 2, 3, 1).reshape(B, L, C)
        return self.act(h)

    def forward(self, x, H, W):
        return self.drop(self.fc2(self.drop(self.hidden(x, H, W))))


# ------------------------------------------------------------------------------- window attention
class WindowAttention(nn.Module):
    """W-MSA / SW-MSA with relative position bias (reference backbone_vit.py:913).

    Parameters and buffers match the reference: qkv, proj, relative_position_bias_table and the
    int64 relative_position_index buffer (kept for state_dict compatibility; the kernel computes
    the index in closed form)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        if attn_drop or proj_drop:
            raise NotImplementedError("inference path: dropout must be 0 (the reference model uses 0)")
        self.dim = dim
        self.window_size = to_2tuple(window_size)
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        wh, ww = self.window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * wh - 1) * (2 * ww - 1), num_heads))
        self.register_buffer("relative_position_index", _relative_position_index(wh, ww))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)

    def forward_image(self, x, ws, shift):
        """x: [B,H,W,C] normalised, un-rolled tokens -> [B,H,W,C] (after proj).
        roll / partition / bias / mask / softmax / AV / reverse all happen inside the kernel."""
        if ws != self.window_size[0] or ws != self.window_size[1]:
            raise ValueError("window size does not match the bias table")
        qkv = self.qkv(x)
        o = ops.window_attention(qkv, self.relative_position_bias_table, self.num_heads, ws, shift,
                                 pad_qkv=self.qkv.bias, scale=self.scale, mask_value=MASK_VALUE)
        return self.proj(o)

    def forward(self, x, mask=None):
        """Reference signature: x [num_windows*B, N, C] already partitioned (backbone_vit.py:961)."""
        if mask is not None:
            # An explicit mask [nW, N, N] (reference backbone_vit.py:979-984): every partitioned window is a one-window image and
            # window b_ reads mask[b_ % nW], evaluated inside the exact attention kernel (sodt_window_attn_ex_fwd).  Inside the
            # detector this never happens: SwinTransformerBlock passes the shift to the kernel, which evaluates the
            # shifted-window mask in closed form.
            wh, ww = self.window_size
            if wh != ww or not x.is_cuda:
                raise NotImplementedError("explicit masks: square windows on CUDA tensors")
            B_, N, C = x.shape
            qkv = self.qkv(x).view(B_, wh, ww, 3 * C)
            o = ops.window_attention_ex(qkv, self.relative_position_bias_table, self.num_heads, wh, 0, scale=self.scale,
                                        mask_value=MASK_VALUE, dense_mask=mask)
            return self.proj(o.view(B_, N, C))
        wh, ww = self.window_size
        if wh != ww:
            raise NotImplementedError("square windows only")
        B_, N, C = x.shape
        return self.forward_image(x.reshape(B_, wh, ww, C), wh, 0).reshape(B_, N, C)

    def extra_repr(self):
        return f"dim={self.dim}, window_size={self.window_size}, num_heads={self.num_heads}"


class Attention(nn.Module):
    """SAM-style global multi-head attention with decomposed relative position embeddings (reference backbone_vit.py:347-404,
    add_decomposed_rel_pos :705-740).  Defined by the reference but never instantiated by its ImageEncoderViT; kept for API
    completeness.  forward(x [B,H,W,C]) -> [B,H,W,C]; the whole map is one window of the exact attention kernel, the
    content-dependent terms q . Rh[yi-yj], q . Rw[xi-xj] are evaluated inside it (sodt_window_attn_ex_fwd)."""

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = True, use_rel_pos: bool = False,
                 rel_pos_zero_init: bool = True, input_size: Optional[Tuple[int, int]] = None):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.use_rel_pos = use_rel_pos
        if use_rel_pos:
            if input_size is None:
                raise ValueError("input_size must be provided if using relative positional encoding")
            self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size[0] - 1, head_dim))
            self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size[1] - 1, head_dim))

    def forward(self, x):
        B, H, W, C = x.shape
        if H != W:
            raise NotImplementedError("the global-attention kernel path covers square token maps")
        rel = None
        if self.use_rel_pos:
            if self.rel_pos_h.shape[0] != 2 * H - 1 or self.rel_pos_w.shape[0] != 2 * W - 1:
                raise NotImplementedError("interpolated relative position tables (get_rel_pos with a size mismatch) are out of scope")
            rel = (self.rel_pos_h, self.rel_pos_w)
        zero_table = ops.cached_derived(self.qkv.weight, ("zero_table", H), lambda t: torch.zeros(
            (2 * H - 1) ** 2, self.num_heads, dtype=torch.float32, device=t.device))
        o = ops.window_attention_ex(self.qkv(x), zero_table, self.num_heads, H, 0, scale=self.scale, rel_pos=rel)
        return self.proj(o)


class SwinTransformerBlock(nn.Module):
    """Swin block (reference backbone_vit.py:1011): x + Attn(LN(x)), then x + Mlp(LN(x))."""

    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 fused_window_process=False, linear_mlp=True):
        super().__init__()
        if drop_path:
            raise NotImplementedError("inference path: drop_path must be 0")
        self.dim = dim
        self.input_resolution = tuple(input_resolution)
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        if min(self.input_resolution) <= self.window_size:   # window covers the map: one global window, no shift
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        if not 0 <= self.shift_size < self.window_size:
            raise ValueError("shift_size must be in [0, window_size)")
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, to_2tuple(self.window_size), num_heads, qkv_bias, qk_scale, attn_drop, drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), act_layer=act_layer, linear_mlp=linear_mlp)
        mask = _shift_mask(*self.input_resolution, self.window_size, self.shift_size) if self.shift_size > 0 else None
        self.register_buffer("attn_mask", mask)
        self.fused_window_process = fused_window_process   # subsumed: windows are never materialised

    def forward(self, x, hw: Optional[Tuple[int, int]] = None, stats=None, want_stats=False):
        """x [B, H*W, C].  ``hw`` overrides the construction-time token grid (extension).
        ``stats`` / ``want_stats`` (extension, bf16): row statistics of ``x`` from the kernel that produced it, and whether to
        return those of the result -> (x, stats); they let norm1 / norm2 run inside the qkv / fc1 GEMMs."""
        H, W = hw if hw is not None else self.input_resolution
        B, L, C = x.shape
        if L != H * W:
            raise ValueError("input feature has wrong size")
        if min(H, W) < self.window_size:
            raise ValueError(f"token grid {H}x{W} is smaller than the window {self.window_size}")
        attn, mlp = self.attn, self.mlp
        if x.dtype == torch.bfloat16 and ops.USE_TC_LINEAR and ops.linear_ln_supported(x, C):
            # bf16, LayerNorms folded: qkv and fc1 read the raw residual stream and normalise in their epilogues from the row
            # statistics that the proj / fc2 GEMM (or the producer of x) emitted with its result
            # row statistics: up to 6 partial pairs per row (C <= 384) go to the GEMM as they are, more are reduced by a small kernel first
            prep = lambda st, eps: st if st.shape[0] <= 6 else ops.finalize_stats(st, C, eps)
            mr = prep(stats, self.norm1.eps) if stats is not None else ops.row_stats(x, self.norm1.eps)
            ln1 = (mr, self.norm1.weight, self.norm1.bias, self.norm1.eps)
            if (ops.USE_FUSED_ATTN and (self.shift_size == 0 or ops.FUSED_ATTN_SHIFTED)
                    and ops.attn_block_supported(x.view(B, H, W, C), attn.num_heads, self.window_size, self.shift_size)):
                # norm1 + qkv + window attention as one kernel: the qkv tensor never exists (bit-identical to the two kernels below)
                o = ops.attn_block(x.view(B, H, W, C), ln1, attn.qkv.weight, attn.qkv.bias, attn.relative_position_bias_table,
                                   attn.num_heads, self.window_size, self.shift_size, scale=attn.scale, mask_value=MASK_VALUE)
            else:
                qkv = ops.linear(x, attn.qkv.weight, attn.qkv.bias, ln=ln1)
                o = ops.window_attention(qkv.view(B, H, W, 3 * C), attn.relative_position_bias_table, attn.num_heads,
                                         self.window_size, self.shift_size, pad_qkv=attn.qkv.bias, scale=attn.scale,
                                         mask_value=MASK_VALUE)
            x, st = ops.linear(o.view(B, L, C), attn.proj.weight, attn.proj.bias, residual=x, want_stats=True)
            if (mlp.linear and isinstance(mlp.act, nn.GELU) and mlp.act.approximate == "none"
                    and ops.mlp_ln_supported(x, mlp.fc1.weight.shape[0]) and mlp.fc2.weight.shape[0] == C):
                # the whole MLP half as one kernel: the [B*L, 4C] hidden activation stays in tensor memory
                r = ops.mlp_ln(x, (prep(st, self.norm2.eps), self.norm2.weight, self.norm2.bias, self.norm2.eps),
                               mlp.fc1.weight, mlp.fc1.bias, mlp.fc2.weight, mlp.fc2.bias, want_stats=want_stats)
                return r
            h = mlp.hidden(x, H, W, ln=(self.norm2, prep(st, self.norm2.eps)))
            if want_stats:
                return ops.linear(h, mlp.fc2.weight, mlp.fc2.bias, residual=x, want_stats=True)
            return ops.linear(h, mlp.fc2.weight, mlp.fc2.bias, residual=x)
        if x.dtype == torch.bfloat16 and ops.USE_TC_LINEAR:
            # bf16: every Linear runs on the tcgen05 GEMM with its bias / GELU / residual fused into the epilogue
            y, _ = ops.add_layernorm(x, None, self.norm1.weight, self.norm1.bias, self.norm1.eps)
            qkv = ops.linear(y, attn.qkv.weight, attn.qkv.bias)
            o = ops.window_attention(qkv.view(B, H, W, 3 * C), attn.relative_position_bias_table, attn.num_heads,
                                     self.window_size, self.shift_size, pad_qkv=attn.qkv.bias, scale=attn.scale,
                                     mask_value=MASK_VALUE)
            x = ops.linear(o.view(B, L, C), attn.proj.weight, attn.proj.bias, residual=x)
            z, _ = ops.add_layernorm(x, None, self.norm2.weight, self.norm2.bias, self.norm2.eps)
            h = mlp.hidden(z, H, W)
            x = ops.linear(h, mlp.fc2.weight, mlp.fc2.bias, residual=x)
            return (x, None) if want_stats else x
        # fp32 (exact) mode: cuBLAS GEMMs; LN1's second output is the residual stream with proj.bias pre-added, so that the
        # proj GEMM adds the residual in its epilogue (addmm) and no elementwise add pass is left.
        y, xb = ops.add_layernorm(x, None, self.norm1.weight, self.norm1.bias, self.norm1.eps,
                                  extra_bias=attn.proj.bias, want_sum=True)
        o = ops.window_attention(attn.qkv(y.view(B, H, W, C)), attn.relative_position_bias_table, attn.num_heads,
                                 self.window_size, self.shift_size, pad_qkv=attn.qkv.bias, scale=attn.scale,
                                 mask_value=MASK_VALUE)
        x = torch.addmm(xb.view(B * L, C), o.view(B * L, C), attn.proj.weight.t()).view(B, L, C)
        z, xb = ops.add_layernorm(x, None, self.norm2.weight, self.norm2.bias, self.norm2.eps,
                                  extra_bias=mlp.fc2.bias, want_sum=True)
        h = mlp.hidden(z, H, W)
        x = torch.addmm(xb.view(B * L, C), h.view(B * L, h.shape[-1]), mlp.fc2.weight.t()).view(B, L, C)
        return (x, None) if want_stats else x

    def extra_repr(self):
        return (f"dim={self.dim}, input_resolution={self.input_resolution}, num_heads={self.num_heads}, "
                f"window_size={self.window_size}, shift_size={self.shift_size}, mlp_ratio={self.mlp_ratio}")


# -------------------------------------------------------------------------- cross-channel attention
class CAttention(nn.Module):
    """Parameter-free multi-head cross attention over windows (reference backbone_vit.py:566).
    Exists for state_dict / pickle compatibility; CAttentionBlock runs all four of them in one kernel."""

    def __init__(self, embedding_dim: int, num_heads: int = 8, shift_size=0):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.num_heads = num_heads

    def forward(self, q, k, v, dimensions=None, mask=None):
        """q, k, v [B_, N, C] (+ mask [nW, N, N]) -> [B_, N, C] with the reference's order of operations (scores, + mask
        BEFORE scaling, / sqrt(C / heads), softmax; backbone_vit.py:589-616).  Standalone use (inside the detector
        CAttentionBlock runs the four cross attentions and their LayerNorms in one kernel).  CUDA tensors whose windows are
        square run on the exact attention kernel (every window a one-window image of cat(q, k, v); (s + mask) / sqrt(d) =
        s / sqrt(d) + mask / sqrt(d), so the mask goes in pre-scaled); host tensors and other shapes: library math in fp32."""
        B_, N, C = q.shape
        ws = math.isqrt(N)
        if (q.is_cuda and ws * ws == N and ws > 1 and C % self.num_heads == 0 and q.dtype in (torch.float32, torch.bfloat16)
                and k.shape == q.shape and v.shape == q.shape):
            hd = C // self.num_heads
            scale = 1.0 / math.sqrt(hd)
            table = torch.zeros((2 * ws - 1) ** 2, self.num_heads, dtype=torch.float32, device=q.device)      # no position bias
            try:
                o = ops.window_attention_ex(torch.cat((q, k, v), dim=-1).view(B_, ws, ws, 3 * C), table, self.num_heads, ws, 0,
                                            scale=scale, dense_mask=None if mask is None else mask.float() * scale)
                return o.view(B_, N, C)
            except ops._capi.SodtError:           # a shape the exact kernel does not cover (head_dim > 64, huge windows)
                pass
        B_, N, C = q.shape
        h = self.num_heads
        c = C // h
        split = lambda t: t.float().reshape(B_, N, h, c).transpose(1, 2)
        s = split(q) @ split(k).transpose(-1, -2)
        if mask is not None:
            nW = mask.shape[0]
            s = (s.reshape(B_ // nW, nW, h, N, N) + mask.float()[None, :, None]).reshape(B_, h, N, N)
        o = torch.softmax(s / math.sqrt(c), dim=-1) @ split(v)
        return o.transpose(1, 2).reshape(B_, N, C).to(q.dtype)


class CAttentionBlock(nn.Module):
    """R<-G, G<-B, B<-IR, IR<-G cross attention + LayerNorm (reference backbone_vit.py:407).

    ``forward`` returns the four streams like the reference; ``forward_fused`` returns their
    concatenation [B,h,w,4C] straight from the kernel (what ImageEncoderViT consumes)."""

    def __init__(self, embedding_dim: int, num_heads: int, out_dim: int = 192, activation: Type[nn.Module] = nn.ReLU,
                 skip_pe: bool = True, shift_size=0):
        super().__init__()
        self.r2g_attn = CAttention(embedding_dim, num_heads, shift_size)
        self.norm1 = nn.LayerNorm(embedding_dim)
        self.rg2b_attn = CAttention(embedding_dim, num_heads, shift_size)
        self.norm2 = nn.LayerNorm(embedding_dim)
        self.rgb2ir_attn = CAttention(embedding_dim, num_heads, shift_size)
        self.norm3 = nn.LayerNorm(embedding_dim)
        self.ir2rgb_attn = CAttention(embedding_dim, num_heads, shift_size)
        self.norm4 = nn.LayerNorm(embedding_dim)
        self.num_heads = num_heads
        self.window_size = 1
        self.input_resolution = (128, 128)
        self.shift_size = shift_size
        mask = _shift_mask(*self.input_resolution, self.window_size, shift_size) if shift_size > 0 else None
        self.register_buffer("attn_mask", mask)

    def forward_fused(self, r, g, b, ir, window_size: Optional[int] = None):
        ws = self.window_size if window_size is None else window_size
        norms = (self.norm1, self.norm2, self.norm3, self.norm4)
        ln_w = torch.stack([n.weight for n in norms])
        ln_b = torch.stack([n.bias for n in norms])
        return ops.cattn_block(r, g, b, ir, ln_w, ln_b, self.num_heads, ws, self.shift_size, self.norm1.eps, MASK_VALUE)

    def forward(self, r, g, b, ir, window_size: int = 1):
        # like the reference, the `window_size` argument is ignored in favour of self.window_size
        C = r.shape[-1]
        return torch.split(self.forward_fused(r, g, b, ir), C, dim=-1)


# ---------------------------------------------------------------------------------------- backbone
class ImageEncoderViT(nn.Module):
    """Channel embedding -> cross-channel block -> 3 Swin-like stages -> 1x1 necks
    (reference backbone_vit.py:11-272).  forward(x [B,4,H,W]) -> [y0, y1, y2] in NCHW shape
    ([B,256,H/4,W/4], [B,256,H/8,W/8], [B,512,H/16,W/16]; channels-last memory)."""

    STAGE_SHIFTS = (0, 2, 0, 2, 0, 2, 0, 2)

    def __init__(self, img_size: int = 512, patch_size: int = 16, in_chans: int = 4, embed_dim: int = 768,
                 depth: int = 11, num_heads: int = 12, mlp_ratio: float = 4.0, out_chans: int = 256,
                 qkv_bias: bool = True, norm_layer: Type[nn.Module] = partial(nn.LayerNorm, eps=1e-6),
                 act_layer: Type[nn.Module] = nn.GELU, use_abs_pos: bool = True, use_rel_pos: bool = True,
                 rel_pos_zero_init: bool = True, window_size: int = 0, global_attn_indexes: Tuple[int, ...] = ()):
        super().__init__()
        self.img_size = img_size
        grid = img_size // 4
        self.patch_embed = PatchEmbed(kernel_size=(1, 1), stride=(1, 1), padding=(0, 0), in_chans=192, embed_dim=embed_dim)
        self.pos_embed: Optional[nn.Parameter] = None
        if use_abs_pos:
            self.pos_embed = nn.Parameter(torch.zeros(1, grid, grid, embed_dim))
        ks = (patch_size, patch_size)
        self.channel_embed_r = PatchEmbed(kernel_size=ks, stride=(4, 4), in_chans=1, embed_dim=48)  # default padding (1,1)
        self.channel_embed_g = PatchEmbed(kernel_size=ks, stride=(4, 4), padding=(0, 0), in_chans=1, embed_dim=48)
        self.channel_embed_b = PatchEmbed(kernel_size=ks, stride=(4, 4), padding=(0, 0), in_chans=1, embed_dim=48)
        self.channel_embed_i = PatchEmbed(kernel_size=ks, stride=(4, 4), padding=(0, 0), in_chans=1, embed_dim=48)
        self.chan_block = CAttentionBlock(embedding_dim=48, num_heads=num_heads)

        def stage(n, dim, res, ws, all_linear=False):
            return nn.ModuleList(
                SwinTransformerBlock(dim=dim, input_resolution=res, num_heads=num_heads, window_size=ws,
                                     shift_size=self.STAGE_SHIFTS[i], mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                                     act_layer=act_layer, linear_mlp=all_linear or self.STAGE_SHIFTS[i] == 0)
                for i in range(n))

        # literal sizes of the reference: dims 192/384/768 -> embed_dim, 2x, 4x at embed_dim=192
        self.stage1 = stage(6, embed_dim, (128, 128), 8)
        self.pmerging1 = PatchMerging((128, 128), embed_dim)
        self.stage2 = stage(4, 384, (64, 64), 8)
        self.pmerging2 = PatchMerging((64, 64), 384)
        self.stage3 = stage(1, 768, (32, 32), 32, all_linear=True)
        self.neck3 = nn.Conv2d(768, 512, kernel_size=1, bias=False)
        self.neck2 = nn.Conv2d(384, 256, kernel_size=1, bias=False)
        self.neck1 = nn.Conv2d(384, 256, kernel_size=1, bias=False)

    @staticmethod
    def _neck(conv, tokens, tokens2=None):
        w = conv.weight
        w = w.reshape(w.shape[0], w.shape[1])
        if tokens.is_cuda:
            return ops.linear(tokens, w, None, x2=tokens2).permute(0, 3, 1, 2)
        if tokens2 is not None:
            tokens = torch.cat((tokens, tokens2), dim=-1)
        return F.linear(tokens, w).permute(0, 3, 1, 2)

    @staticmethod
    def _run_stage(blocks, x, hw):
        st = None                                  # row statistics travel with x from GEMM epilogue to GEMM epilogue
        for blk in blocks:
            x, st = blk(x, hw, stats=st, want_stats=True)
        return x

    def _front_end_params(self):
        """fp32 [4,E,16] conv weights, [4,E] conv biases and LayerNorm parameters of the four streams (cached per
        parameter version; rebuilt after load_state_dict / .to())."""
        embeds = (self.channel_embed_r, self.channel_embed_g, self.channel_embed_b, self.channel_embed_i)
        cb = self.chan_block
        srcs = [e.proj.weight for e in embeds] + [e.proj.bias for e in embeds] + \
               [n.weight for n in (cb.norm1, cb.norm2, cb.norm3, cb.norm4)] + [n.bias for n in (cb.norm1, cb.norm2, cb.norm3, cb.norm4)]
        key = tuple((id(t), t._version, t.data_ptr()) for t in srcs)
        if getattr(self, "_fe_key", None) != key:
            f = lambda ts: torch.stack([t.detach().float().reshape(t.shape[0], -1) for t in ts]).contiguous()
            self._fe_cache = (f(srcs[0:4]), f(srcs[4:8]).squeeze(-1), f(srcs[8:12]).squeeze(-1), f(srcs[12:16]).squeeze(-1))
            self._fe_key = key
        return self._fe_cache

    def _front_end(self, x, ir_u8=None):
        cb = self.chan_block
        e = self.channel_embed_r.proj
        if ir_u8 is not None:
            # uint8 images straight into the fused kernel (x = rgb uint8 [B,3,H,W]): no conversion / scaling / concat passes
            if not (x.is_cuda and cb.window_size == 1 and e.out_channels == 48 and e.kernel_size == (4, 4) and e.stride == (4, 4)
                    and self.channel_embed_g.proj.padding == (0, 0) and e.padding in ((1, 1), (0, 0))
                    and x.shape[2] % 4 == 0 and x.shape[3] % 4 == 0):
                raise ValueError("uint8 input needs the fused front end (CUDA, 4x4/4 channel embeddings, window-1 block)")
            cw, cbias, lw, lb = self._front_end_params()
            return ops.frontend_u8(x, ir_u8, cw, cbias, lw, lb, e.weight.dtype, pad_r=e.padding[0], eps=cb.norm1.eps)
        fused = (x.is_cuda and cb.window_size == 1 and e.out_channels == 48 and e.kernel_size == (4, 4) and e.stride == (4, 4)
                 and self.channel_embed_g.proj.padding == (0, 0) and e.padding in ((1, 1), (0, 0))
                 and x.shape[2] % 4 == 0 and x.shape[3] % 4 == 0)
        if fused:
            cw, cbias, lw, lb = self._front_end_params()
            return ops.frontend(x, cw, cbias, lw, lb, pad_r=e.padding[0], eps=cb.norm1.eps)
        r, g, b, i = get_channels(x)
        r = self.channel_embed_r(r)
        g = self.channel_embed_g(g)
        b = self.channel_embed_b(b)
        i = self.channel_embed_i(i)
        return self.chan_block.forward_fused(r, g, b, i)

    def forward(self, x: torch.Tensor, ir_u8: Optional[torch.Tensor] = None):
        """x [B,4,H,W] in the model dtype, or (extension) x = uint8 RGB [B,3,H,W] with ``ir_u8`` uint8 [B,>=1,H,W]."""
        pos = self.pos_embed                                     # silently skipped on a size mismatch, like the reference
        st = None
        if ir_u8 is not None and self.patch_embed.proj.weight.dtype == torch.bfloat16:
            # uint8 images to patch-embedded tokens in one kernel: the 192-channel concat never reaches HBM
            cb, e, pe = self.chan_block, self.channel_embed_r.proj, self.patch_embed.proj
            pos_f = pos if pos is not None and pos.shape[1] == x.shape[2] // 4 and pos.shape[2] == x.shape[3] // 4 else None
            if (cb.window_size == 1 and e.out_channels == 48 and e.kernel_size == (4, 4) and e.stride == (4, 4)
                    and self.channel_embed_g.proj.padding == (0, 0) and e.padding in ((1, 1), (0, 0)) and pe.kernel_size == (1, 1)
                    and ops.frontend_embed_u8_supported(x, ir_u8, 48, pe.out_channels, pos_f)):
                cw, cbias, lw, lb = self._front_end_params()
                x, st = ops.frontend_embed_u8(x, ir_u8, cw, cbias, lw, lb, pe.weight, pe.bias, pos_f, pad_r=e.padding[0],
                                              eps=cb.norm1.eps, want_stats=True)
        if st is None:
            x = self._front_end(x, ir_u8)                        # [B,h,w,192], the reference's concat
            if pos is not None and x.shape[1] != pos.shape[1]:
                pos = None
            x, st = self.patch_embed.forward_tokens(x, pos, want_stats=True)
        B, h, w, C = x.shape
        x = x.reshape(B, h * w, C)
        kept = []
        for n, blk in enumerate(self.stage1):
            x, st = blk(x, (h, w), stats=st, want_stats=True)
            if n in (4, 5):
                kept.append(x.view(B, h, w, C))
        x = self.pmerging1(x, (h, w))
        h2, w2 = h // 2, w // 2
        x = self._run_stage(self.stage2, x, (h2, w2))
        y1 = x.view(B, h2, w2, -1)
        x = self.pmerging2(x, (h2, w2))
        h3, w3 = h2 // 2, w2 // 2
        blk3 = self.stage3[0]
        if min(h3, w3) <= blk3.window_size and (h3 != w3 or h3 != blk3.window_size):
            raise ValueError(f"stage-3 token grid {h3}x{w3} does not match its {blk3.window_size}-token window; "
                             "inputs must be multiples of 512 px (or exactly 16*window)")
        x = self._run_stage(self.stage3, x, (h3, w3))
        y2 = x.view(B, h3, w3, -1)
        # neck1 reads concat(block 5, block 6) of stage 1: two A operands of one GEMM instead of a concatenated copy
        return [self._neck(self.neck1, kept[0], kept[1]), self._neck(self.neck2, y1), self._neck(self.neck3, y2)]
