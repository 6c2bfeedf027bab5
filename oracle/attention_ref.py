"""Oracle (test infrastructure): window attention, Swin block, cross-channel attention.

CPU restatement in torch; math runs in ``dtype`` (float64 by default so that both the
fp32 reference and the fp32 CUDA kernels sit ~1e-7 away from it).  All ``file:line``
citations are into /root/reference/basics/models/.
"""
import math

import torch
import torch.nn.functional as F

MASK_VALUE = -100.0  # backbone_vit.py:1077 -- finite, NOT -inf


# --------------------------------------------------------------------------- windows
def pad_amounts(H, W, ws):
    """backbone_vit.py:632-633."""
    return (ws - H % ws) % ws, (ws - W % ws) % ws


def window_partition(x, ws, top_padding=False):
    """[B,H,W,C] -> ([B*nW, ws, ws, C], (Hp, Wp)).  backbone_vit.py:619-643.

    Zero padding goes to the bottom/right, or to the top/left with ``top_padding``.
    """
    B, H, W, C = x.shape
    ph, pw = pad_amounts(H, W, ws)
    Hp, Wp = H + ph, W + pw
    canvas = x.new_zeros(B, Hp, Wp, C)
    if top_padding:
        canvas[:, ph:, pw:] = x
    else:
        canvas[:, :H, :W] = x
    tiles = canvas.reshape(B, Hp // ws, ws, Wp // ws, ws, C).transpose(2, 3)
    return tiles.reshape(-1, ws, ws, C).contiguous(), (Hp, Wp)


def window_unpartition(windows, ws, pad_hw, hw, top_padding=False):
    """Inverse of window_partition followed by the crop.  backbone_vit.py:646-672."""
    Hp, Wp = pad_hw
    H, W = hw
    nwh, nww = Hp // ws, Wp // ws
    B = windows.shape[0] // (nwh * nww)
    img = windows.reshape(B, nwh, nww, ws, ws, -1).transpose(2, 3).reshape(B, Hp, Wp, -1)
    if top_padding:
        return img[:, Hp - H:, Wp - W:].contiguous()
    return img[:, :H, :W].contiguous()


def relative_position_index(wh, ww):
    """Closed form of the buffer built at backbone_vit.py:941-951.

    index[i, j] = (yi - yj + wh - 1) * (2*ww - 1) + (xi - xj + ww - 1).
    """
    ys = torch.arange(wh).repeat_interleave(ww)
    xs = torch.arange(ww).repeat(wh)
    dy = ys[:, None] - ys[None, :] + (wh - 1)
    dx = xs[:, None] - xs[None, :] + (ww - 1)
    return dy * (2 * ww - 1) + dx


def region_ids(H, W, ws, shift):
    """Region id image of the shifted-window mask in the ROLLED frame.

    backbone_vit.py:1060-1072: three slices per axis, [0,-ws), [-ws,-shift), [-shift,end).
    """
    def axis(n):
        a = torch.arange(n)
        return (a >= n - ws).long() + (a >= n - shift).long()
    return 3 * axis(H)[:, None] + axis(W)[None, :]


def shift_attn_mask(H, W, ws, shift, dtype=torch.float32):
    """[nW, N, N] additive mask, 0 or -100.  backbone_vit.py:1058-1079.

    The id image is zero padded by window_partition exactly like the activations
    (padded tokens therefore carry region id 0).
    """
    ids = region_ids(H, W, ws, shift).to(dtype).reshape(1, H, W, 1)
    win, _ = window_partition(ids, ws)
    flat = win.reshape(-1, ws * ws)
    diff = flat[:, None, :] - flat[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, MASK_VALUE), torch.zeros_like(diff))


# ------------------------------------------------------------------ window attention
def window_attention_core(q, k, v, bias_table, wh, ww, scale, mask=None):
    """softmax(scale*q k^T + table[index] (+ mask)) v per window and head.

    q,k,v: [B_, heads, N, hd]; bias_table [(2wh-1)(2ww-1), heads]; mask [nW,N,N] or None.
    backbone_vit.py:971-989 (scale is applied to q first, bias then mask are added
    to the scaled scores, softmax over keys).
    """
    B_, nh, N, hd = q.shape
    s = (q * scale) @ k.transpose(-1, -2)
    idx = relative_position_index(wh, ww).reshape(-1)
    bias = bias_table.to(s.dtype)[idx].reshape(N, N, nh).permute(2, 0, 1)
    s = s + bias[None]
    if mask is not None:
        nW = mask.shape[0]
        s = (s.reshape(B_ // nW, nW, nh, N, N) + mask.to(s.dtype)[None, :, None]).reshape(B_, nh, N, N)
    p = torch.softmax(s, dim=-1)
    return p @ v


def window_attention(xw, p, prefix, heads, wh, ww, mask=None, dtype=torch.float64):
    """WindowAttention.forward, backbone_vit.py:961-992.  xw [B_, N, C]."""
    B_, N, C = xw.shape
    hd = C // heads
    w = lambda n: p[prefix + n].to(dtype)
    qkv = F.linear(xw.to(dtype), w("qkv.weight"), w("qkv.bias"))
    qkv = qkv.reshape(B_, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    o = window_attention_core(qkv[0], qkv[1], qkv[2], w("relative_position_bias_table"),
                              wh, ww, hd ** -0.5, mask)
    o = o.transpose(1, 2).reshape(B_, N, C)
    return F.linear(o, w("proj.weight"), w("proj.bias"))


def attention_on_qkv_image(qkv, bias_table, heads, ws, shift, pad_qkv=None, dtype=torch.float64):
    """The exact contract of the CUDA op ``sodt_window_attn_fwd``.

    qkv: [B, H, W, 3C] projection of the UN-rolled, UN-padded token image (channel
    order (3, heads, hd) as produced by backbone_vit.py:968).  Performs roll(-shift),
    pad, partition, masked window attention, un-partition, crop and roll(+shift) and
    returns [B, H, W, C] (the input of ``proj``).  Tokens added by padding carry
    ``pad_qkv`` ([3C], the qkv bias: Linear applied to a zero row, backbone_vit.py:638,968)
    or zeros.  Follows backbone_vit.py:1094-1123 + 961-989.
    """
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    x = qkv.to(dtype)
    ph, pw = pad_amounts(H, W, ws)
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    win, (Hp, Wp) = window_partition(x, ws)
    if (ph or pw) and pad_qkv is not None:
        flag = torch.ones(1, H, W, 1, dtype=dtype)
        fw, _ = window_partition(flag, ws)
        nW = fw.shape[0]
        fw = fw.repeat(B, 1, 1, 1)
        win = win + (1.0 - fw) * pad_qkv.to(dtype).reshape(1, 1, 1, C3)
    N = ws * ws
    t = win.reshape(-1, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    mask = shift_attn_mask(H, W, ws, shift, dtype) if shift > 0 else None
    o = window_attention_core(t[0], t[1], t[2], bias_table.to(dtype), ws, ws, hd ** -0.5, mask)
    o = o.transpose(1, 2).reshape(-1, ws, ws, C)
    img = window_unpartition(o, ws, (Hp, Wp), (H, W))
    if shift > 0:
        img = torch.roll(img, shifts=(shift, shift), dims=(1, 2))
    return img


# ------------------------------------------------------------------------ swin block
def effective_window(H, W, ws, shift):
    """backbone_vit.py:1042-1045: clamp the window to the map, drop the shift."""
    if min(H, W) <= ws:
        return min(H, W), 0
    return ws, shift


def mlp(x, p, prefix, H, W, linear_mlp, dtype):
    """Mlp.forward, backbone_vit.py:884-908.  x [B, L, C]."""
    w = lambda n: p[prefix + n].to(dtype)
    if linear_mlp:
        h = F.gelu(F.linear(x, w("fc1.weight"), w("fc1.bias")))
        return F.linear(h, w("fc2.weight"), w("fc2.bias"))
    B, L, C = x.shape
    h = F.linear(x, w("fc1.weight"), w("fc1.bias"))
    h = h.transpose(1, 2).reshape(B, C, H, W)
    h = F.pad(h, (0, 1, 0, 1))  # one zero column right, one zero row below (:896)
    h = F.conv2d(h, w("conv1.weight"), w("conv1.bias"))  # 2x2, stride 1 -> H x W again
    h = F.gelu(h.permute(0, 2, 3, 1).reshape(B, L, C))
    return F.linear(h, w("fc2.weight"), w("fc2.bias"))


def swin_block(x, p, prefix, H, W, heads, ws, shift, linear_mlp, dtype=torch.float64):
    """SwinTransformerBlock.forward, backbone_vit.py:1084-1130.  x [B, H*W, C].

    ``p`` maps state_dict keys (with ``prefix``) to tensors.  LayerNorm eps is the
    nn.LayerNorm default 1e-5 (backbone_vit.py:1048).
    """
    B, L, C = x.shape
    ws, shift = effective_window(H, W, ws, shift)
    w = lambda n: p[prefix + n].to(dtype)
    x = x.to(dtype)
    y = F.layer_norm(x, (C,), w("norm1.weight"), w("norm1.bias"), 1e-5).reshape(B, H, W, C)
    if shift > 0:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    win, pad_hw = window_partition(y, ws)
    mask = shift_attn_mask(H, W, ws, shift, dtype) if shift > 0 else None
    a = window_attention(win.reshape(-1, ws * ws, C), p, prefix + "attn.", heads, ws, ws, mask, dtype)
    y = window_unpartition(a.reshape(-1, ws, ws, C), ws, pad_hw, (H, W))
    if shift > 0:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    x = x + y.reshape(B, L, C)
    z = F.layer_norm(x, (C,), w("norm2.weight"), w("norm2.bias"), 1e-5)
    return x + mlp(z, p, prefix + "mlp.", H, W, linear_mlp, dtype)


# ------------------------------------------------------------- cross-channel attention
def cattention(q, k, v, heads, mask=None):
    """Parameter-free multi-head cross attention.  backbone_vit.py:589-616
    (general-N twin: backbone_swinv2.py:503-517).  q,k,v [B_, N, C].

    Order of operations: scores, + mask (BEFORE scaling, :601-604), / sqrt(c), softmax.
    """
    B_, N, C = q.shape
    c = C // heads
    split = lambda t: t.reshape(B_, N, heads, c).transpose(1, 2)
    s = split(q) @ split(k).transpose(-1, -2)
    if mask is not None:
        nW = mask.shape[0]
        s = (s.reshape(B_ // nW, nW, heads, N, N) + mask.to(s.dtype)[None, :, None]).reshape(B_, heads, N, N)
    s = s / math.sqrt(c)
    o = torch.softmax(s, dim=-1) @ split(v)
    return o.transpose(1, 2).reshape(B_, N, C)


CATTN_PAIRS = ((0, 1), (1, 2), (2, 3), (3, 1))  # (query stream, key/value stream): R<-G, G<-B, B<-IR, IR<-G


def cattention_block(streams, ln_w, ln_b, heads, ws=1, shift=0, eps=1e-5, dtype=torch.float64):
    """CAttentionBlock.forward, backbone_vit.py:469-561 (ws=1, shipped) and
    backbone_swinv2.py:429-469 (general ws).  ``streams`` = (r, g, b, ir), each
    [B,h,w,C]; ln_w / ln_b: four [C] vectors.  Returns 4 tensors [B,h,w,C].

    x_i = LayerNorm_i(stream_i + attn(q=stream_i, k=v=stream_partner)).
    """
    s = [t.to(dtype) for t in streams]
    B, h, w, C = s[0].shape
    if shift > 0:
        s_att = [torch.roll(t, shifts=(-shift, -shift), dims=(1, 2)) for t in s]
        mask = shift_attn_mask(h, w, ws, shift, dtype)
    else:
        s_att, mask = s, None
    wins, pad_hw = [], None
    for t in s_att:
        wt, pad_hw = window_partition(t, ws)
        wins.append(wt.reshape(-1, ws * ws, C))
    outs = []
    for i, (qi, ki) in enumerate(CATTN_PAIRS):
        o = cattention(wins[qi], wins[ki], wins[ki], heads, mask)
        o = window_unpartition(o.reshape(-1, ws, ws, C), ws, pad_hw, (h, w))
        if shift > 0:
            o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
        outs.append(F.layer_norm(s[qi] + o, (C,), ln_w[i].to(dtype), ln_b[i].to(dtype), eps))
    return outs


# ------------------------------------------------------------------ attention variants
def cosine_window_attention(xw, p, prefix, heads, ws, mask=None, dtype=torch.float64):
    """SwinV2 cosine window attention, backbone_swinv2.py:895-949.  xw [B_, N, C]; p holds qkv.weight, q_bias, v_bias,
    logit_scale [heads,1,1], cpb_mlp.{0.weight,0.bias,2.weight}, proj.{weight,bias}."""
    B_, N, C = xw.shape
    w = lambda n: p[prefix + n].to(dtype)
    bias = torch.cat((w("q_bias"), torch.zeros(C, dtype=dtype), w("v_bias")))
    qkv = F.linear(xw.to(dtype), w("qkv.weight"), bias).reshape(B_, N, 3, heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    s = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)                       # :905
    s = s * torch.clamp(w("logit_scale"), max=math.log(1.0 / 0.01)).exp()                         # :906-907
    # continuous relative position bias: log-spaced coordinate table -> MLP -> 16 sigmoid (:858-874, :909-915)
    c = torch.arange(-(ws - 1), ws, dtype=torch.float32)
    table = torch.stack(torch.meshgrid([c, c], indexing="ij")).permute(1, 2, 0).contiguous()
    table = table / (ws - 1) * 8
    table = (torch.sign(table) * torch.log2(torch.abs(table) + 1.0) / math.log2(8)).to(dtype)
    t = F.linear(F.relu(F.linear(table, w("cpb_mlp.0.weight"), w("cpb_mlp.0.bias"))), w("cpb_mlp.2.weight")).reshape(-1, heads)
    idx = relative_position_index(ws, ws).reshape(-1)
    s = s + (16 * torch.sigmoid(t[idx].reshape(N, N, heads).permute(2, 0, 1)))[None]
    if mask is not None:
        nW = mask.shape[0]
        s = (s.reshape(B_ // nW, nW, heads, N, N) + mask.to(dtype)[None, :, None]).reshape(B_, heads, N, N)
    o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(o, w("proj.weight"), w("proj.bias"))


def sam_attention(x, p, prefix, heads, use_rel_pos, dtype=torch.float64):
    """SAM-style global attention with decomposed relative position embeddings, backbone_vit.py:386-404 and
    add_decomposed_rel_pos :705-740 (square maps whose size matches the tables: no interpolation).  x [B,H,W,C]."""
    B, H, W, C = x.shape
    w = lambda n: p[prefix + n].to(dtype)
    hd = C // heads
    qkv = F.linear(x.to(dtype), w("qkv.weight"), w("qkv.bias")).reshape(B, H * W, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.reshape(3, B * heads, H * W, hd).unbind(0)
    s = (q * hd ** -0.5) @ k.transpose(-2, -1)
    if use_rel_pos:
        ar = torch.arange(H)
        Rh = w("rel_pos_h")[(ar[:, None] - ar[None, :]) + (H - 1)]          # get_rel_pos, equal sizes (:674-703)
        aw = torch.arange(W)
        Rw = w("rel_pos_w")[(aw[:, None] - aw[None, :]) + (W - 1)]
        rq = q.reshape(B * heads, H, W, hd)
        rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
        rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
        s = (s.view(-1, H, W, H, W) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(-1, H * W, H * W)
    o = (torch.softmax(s, dim=-1) @ v).view(B, heads, H, W, hd).permute(0, 2, 3, 1, 4).reshape(B, H, W, C)
    return F.linear(o, w("proj.weight"), w("proj.bias"))


def mf_block(rgb, ir, p, prefix="", dtype=torch.float64):
    """SuperYOLO's MF fusion block, common.py:165-212.  rgb [B,3,H,W], ir [B,1,H,W] -> [B,64,H,W]."""
    w = lambda n: p[prefix + n].to(dtype)

    def se(x, name):
        y = x.mean(dim=(2, 3))
        y = torch.sigmoid(F.linear(F.relu(F.linear(y, w(name + ".fc.0.weight"))), w(name + ".fc.2.weight")))
        return x * y[:, :, None, None]

    rgb, ir = rgb.to(dtype), ir.to(dtype)
    r, i = se(rgb, "se_r"), se(ir, "se_i")
    rm = F.conv2d(r, w("mask_map_r.weight"), w("mask_map_r.bias")).repeat(1, 3, 1, 1) * r
    im = F.conv2d(i, w("mask_map_i.weight"), w("mask_map_i.bias")) * i
    out_ir = F.conv2d(im + ir, w("bottleneck1.weight"), None, 1, 1)
    out_rgb = F.conv2d(rm + rgb, w("bottleneck2.weight"), None, 1, 1)
    return se(torch.cat([out_rgb, out_ir], 1), "se")
