"""Tensor-level wrappers around the C ABI (include/sodt_b200.h).

Each function validates shapes / dtypes / devices, allocates the output with torch (the
kernels never allocate), and enqueues the kernels on ``torch.cuda.current_stream()``.
The same functions are registered as ``torch.ops.sodt.*`` custom ops (with fake-tensor
shape functions) so that the modules stay traceable.  No CPU path exists.
"""
import math
import weakref

import os

import torch

from . import _capi

_DT = {torch.float32: 0, torch.bfloat16: 1}


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _capi.SodtError("sodt_b200 ops run on CUDA tensors only (there is no CPU fallback)")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


# Optional per-call CUDA-event timing on the launching stream (bench.py's roofline leg).
_timing = None


def enable_kernel_timing(on=True):
    """Start (or stop) recording a CUDA-event pair around every sodt kernel sequence."""
    global _timing
    _timing = {} if on else None


def kernel_timing_enabled():
    return _timing is not None


def kernel_timings():
    """{label: [ms, ...]} of everything recorded since enable_kernel_timing(); synchronises."""
    if _timing is None:
        return {}
    torch.cuda.synchronize()
    return {k: [a.elapsed_time(b) for a, b in v] for k, v in _timing.items()}


class _Timed:
    def __init__(self, label):
        self.label = label

    def __enter__(self):
        if _timing is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _timing is not None:
            self.b.record()
            _timing.setdefault(self.label, []).append((self.a, self.b))
        return False


def launch_count():
    return _capi.lib().sodt_launch_count()


def reset_launch_count():
    _capi.lib().sodt_reset_launch_count()


_scratch_bufs = {}


def _scratch(tag, device, nbytes):
    """Per (device, stream) scratch buffer owned by torch's allocator; the kernels never allocate."""
    key = (tag, device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch_bufs.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)
        _scratch_bufs[key] = buf
    return buf


# ------------------------------------------------------------------------------ window attention
CACHE_BIAS_TABLE = True     # prepare the kernels' bias-table image once per weight version (False: once per call)


def window_attention(qkv, bias_table, heads, ws, shift=0, pad_qkv=None, scale=None, mask_value=-100.0):
    """qkv [B,H,W,3C] (f32 / bf16) -> [B,H,W,C].  See sodt_window_attn_fwd."""
    _require_cuda(qkv, bias_table, pad_qkv)
    if qkv.dim() != 4 or qkv.shape[-1] % 3:
        raise ValueError("qkv must be [B, H, W, 3*C]")
    if qkv.dtype not in _DT:
        raise TypeError(f"unsupported dtype {qkv.dtype}")
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    if C % heads:
        raise ValueError("C must be divisible by heads")
    span = (2 * ws - 1) ** 2
    if tuple(bias_table.shape) != (span, heads):
        raise ValueError(f"bias_table must be [{span}, {heads}], got {tuple(bias_table.shape)}")
    qkv = qkv.contiguous()
    table = _as_f32(bias_table)          # cached fp32 copy (the parameter is bf16 in a bf16 model)
    if pad_qkv is not None:
        pad_qkv = pad_qkv.detach().to(qkv.dtype).contiguous()
        if pad_qkv.numel() != C3:
            raise ValueError("pad_qkv must have 3*C elements")
    out = torch.empty((B, H, W, C), dtype=qkv.dtype, device=qkv.device)
    if scale is None:
        scale = (C // heads) ** -0.5
    lib = _capi.lib()
    dt = _DT[qkv.dtype]
    kclass = lib.sodt_window_attn_kernel_class(B, H, W, C, heads, ws, shift, dt)
    if kclass < 0:
        _capi.check(kclass, "sodt_window_attn_kernel_class")
    label = f"window_attn[B={B},H={H},W={W},C={C},heads={heads},ws={ws},shift={shift}]"
    if kclass > 0 and CACHE_BIAS_TABLE:
        # the tensor-core kernels' image of the bias table is prepared once per weight version (and kernel class), not per call
        def prepare(t):
            t32 = _as_f32(t)
            buf = torch.empty(max(int(lib.sodt_window_attn_workspace_bytes(C, heads, ws)), 16), dtype=torch.uint8, device=t.device)
            with torch.cuda.device(t.device):
                _capi.check(lib.sodt_window_attn_prepare(t32.data_ptr(), B, H, W, C, heads, ws, shift, dt, buf.data_ptr(), buf.numel(),
                                                         _stream()), "sodt_window_attn_prepare")
            buf.record_stream(torch.cuda.current_stream(t.device))
            return buf
        wsp = cached_derived(bias_table, ("wattn_image", kclass, ws), prepare)
        with torch.cuda.device(qkv.device), _Timed(label):
            st = lib.sodt_window_attn_fwd_prepared(qkv.data_ptr(), table.data_ptr(), _ptr(pad_qkv), out.data_ptr(),
                                                   B, H, W, C, heads, ws, shift, dt, float(scale),
                                                   float(mask_value), wsp.data_ptr(), wsp.numel(), _stream())
        _capi.check(st, "sodt_window_attn_fwd_prepared")
        return out
    wsp = _scratch("wattn", qkv.device, lib.sodt_window_attn_workspace_bytes(C, heads, ws))
    with torch.cuda.device(qkv.device), _Timed(label):
        st = lib.sodt_window_attn_fwd(qkv.data_ptr(), table.data_ptr(), _ptr(pad_qkv), out.data_ptr(),
                                      B, H, W, C, heads, ws, shift, dt, float(scale),
                                      float(mask_value), wsp.data_ptr(), wsp.numel(), _stream())
    _capi.check(st, "sodt_window_attn_fwd")
    return out


def window_attention_ex(qkv, bias_table, heads, ws, shift=0, pad_qkv=None, scale=None, mask_value=-100.0, dense_mask=None,
                        head_scale=None, normalize_qk=False, rel_pos=None):
    """window_attention with a non-standard score epilogue (see sodt_window_attn_ex_fwd): an explicit dense additive mask
    [mask_windows, N, N], a per-head score multiplier, L2-normalised q / k (SwinV2 cosine attention), decomposed relative
    position embeddings ``rel_pos=(rel_pos_h, rel_pos_w)``, each [2*ws-1, head_dim] (SAM-style Attention).  Exact fp32-math kernel."""
    _require_cuda(qkv, bias_table, pad_qkv, dense_mask, head_scale)
    rph = rpw = None
    if rel_pos is not None:
        rph, rpw = (t.detach().to(torch.float32).contiguous() for t in rel_pos)
        _require_cuda(rph, rpw)
        hd_ = qkv.shape[-1] // 3 // heads
        if tuple(rph.shape) != (2 * ws - 1, hd_) or tuple(rpw.shape) != (2 * ws - 1, hd_):
            raise ValueError(f"rel_pos tables must be [{2 * ws - 1}, {hd_}]")
    if qkv.dim() != 4 or qkv.shape[-1] % 3 or qkv.dtype not in _DT:
        raise ValueError("qkv must be [B, H, W, 3*C] in fp32 / bf16")
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    if C % heads or tuple(bias_table.shape) != ((2 * ws - 1) ** 2, heads):
        raise ValueError("bias_table must be [(2*ws-1)^2, heads] and C divisible by heads")
    qkv = qkv.contiguous()
    table = _as_f32(bias_table)
    if pad_qkv is not None:
        pad_qkv = pad_qkv.detach().to(qkv.dtype).contiguous()
    mask_windows = 0
    if dense_mask is not None:
        N = ws * ws
        if dense_mask.dim() != 3 or tuple(dense_mask.shape[1:]) != (N, N):
            raise ValueError(f"dense_mask must be [mask_windows, {N}, {N}]")
        dense_mask = dense_mask.detach().to(torch.float32).contiguous()
        mask_windows = dense_mask.shape[0]
    if head_scale is not None:
        head_scale = head_scale.detach().to(torch.float32).reshape(-1).contiguous()
        if head_scale.numel() != heads:
            raise ValueError("head_scale must have one entry per head")
    out = torch.empty((B, H, W, C), dtype=qkv.dtype, device=qkv.device)
    if scale is None:
        scale = (C // heads) ** -0.5
    with torch.cuda.device(qkv.device), _Timed(f"window_attn_ex[B={B},H={H},W={W},C={C},heads={heads},ws={ws},shift={shift}]"):
        st = _capi.lib().sodt_window_attn_ex_fwd(qkv.data_ptr(), table.data_ptr(), _ptr(pad_qkv), out.data_ptr(), B, H, W, C, heads, ws,
                                                 shift, _DT[qkv.dtype], float(scale), float(mask_value), _ptr(dense_mask), mask_windows,
                                                 _ptr(head_scale), int(bool(normalize_qk)), _ptr(rph), _ptr(rpw), _stream())
    _capi.check(st, "sodt_window_attn_ex_fwd")
    return out


# ------------------------------------------------------------------------------------ LayerNorm
_f32_cache = {}


def _as_f32(t):
    """fp32 contiguous copy of a small parameter.  Cached per tensor OBJECT (weak reference) and validated
    against its version counter and storage address, so bf16 models do not re-convert their LayerNorm
    weights on every call and a recycled allocation can never alias a stale copy."""
    if t is None:
        return None
    if t.dtype == torch.float32 and t.is_contiguous():
        return t.detach()
    key = id(t)
    hit = _f32_cache.get(key)
    if hit is not None:
        ref, version, ptr, val = hit
        if ref() is t and version == t._version and ptr == t.data_ptr():
            return val
    if len(_f32_cache) > 4096:
        _f32_cache.clear()
    val = t.detach().to(torch.float32).contiguous()
    _f32_cache[key] = (weakref.ref(t), t._version, t.data_ptr(), val)
    return val


_derived_cache = {}


def cached_derived(t, tag, fn):
    """fn(t) cached per tensor object / version / storage address (weights re-laid-out once per load, e.g. conv taps)."""
    key = (id(t), tag)
    hit = _derived_cache.get(key)
    if hit is not None:
        ref, version, ptr, val = hit
        if ref() is t and version == t._version and ptr == t.data_ptr():
            return val
    if len(_derived_cache) > 4096:
        _derived_cache.clear()
    val = fn(t)
    _derived_cache[key] = (weakref.ref(t), t._version, t.data_ptr(), val)
    return val


def add_layernorm(a, r, weight, bias, eps=1e-5, extra_bias=None, want_sum=False):
    """y = LayerNorm(a (+ r)) over the last dim; optionally also returns a (+ r) (+ extra_bias).
    See sodt_add_layernorm_fwd.  Returns (y, sum_or_None)."""
    _require_cuda(a, r, weight, bias, extra_bias)
    if a.dtype not in _DT:
        raise TypeError(f"unsupported dtype {a.dtype}")
    C = a.shape[-1]
    a = a.contiguous()
    if r is not None:
        if r.shape != a.shape or r.dtype != a.dtype:
            raise ValueError("residual must match the input")
        r = r.contiguous()
    rows = a.numel() // C
    y = torch.empty_like(a)
    s = torch.empty_like(a) if want_sum else None
    w32, b32, e32 = _as_f32(weight), _as_f32(bias), _as_f32(extra_bias)
    with torch.cuda.device(a.device), _Timed(f"add_layernorm[rows={rows},C={C}]"):
        st = _capi.lib().sodt_add_layernorm_fwd(a.data_ptr(), _ptr(r), w32.data_ptr(), b32.data_ptr(), _ptr(e32), _ptr(s),
                                                y.data_ptr(), rows, C, float(eps), _DT[a.dtype], _stream())
    _capi.check(st, "sodt_add_layernorm_fwd")
    return y, s


# ------------------------------------------------------------------------- cross-channel block
def cattn_block(r, g, b, ir, ln_w, ln_b, heads, ws=1, shift=0, eps=1e-5, mask_value=-100.0):
    """Four streams [B,h,w,C] (any common strides) -> [B,h,w,4C].  ln_w / ln_b: [4,C].
    See sodt_cattn_block_fwd."""
    _require_cuda(r, g, b, ir, ln_w, ln_b)
    streams = [r, g, b, ir]
    if any(t.shape != r.shape or t.dtype != r.dtype for t in streams) or r.dim() != 4:
        raise ValueError("the four streams must share shape [B,h,w,C] and dtype")
    if r.dtype not in _DT:
        raise TypeError(f"unsupported dtype {r.dtype}")
    if any(t.stride() != r.stride() for t in streams):
        streams = [t.contiguous() for t in streams]
    B, h, w, C = r.shape
    sb, sy, sx, sc = streams[0].stride()
    ln_w = ln_w.detach().to(torch.float32).contiguous()
    ln_b = ln_b.detach().to(torch.float32).contiguous()
    if ln_w.numel() != 4 * C or ln_b.numel() != 4 * C:
        raise ValueError("ln_w / ln_b must be [4, C]")
    out = torch.empty((B, h, w, 4 * C), dtype=r.dtype, device=r.device)
    with torch.cuda.device(r.device), _Timed(f"cattn_block[B={B},h={h},w={w},C={C},ws={ws}]"):
        st = _capi.lib().sodt_cattn_block_fwd(streams[0].data_ptr(), streams[1].data_ptr(), streams[2].data_ptr(),
                                              streams[3].data_ptr(), sb, sy, sx, sc, ln_w.data_ptr(), ln_b.data_ptr(),
                                              out.data_ptr(), B, h, w, C, heads, ws, shift, float(eps),
                                              float(mask_value), _DT[r.dtype], _stream())
    _capi.check(st, "sodt_cattn_block_fwd")
    return out


# ------------------------------------------------------------------------------------- front end
def frontend(x, conv_w, conv_b, ln_w, ln_b, pad_r=1, eps=1e-5):
    """x [B,4,H,W] -> [B,H/4,W/4,4E]: four channel embeddings + window-1 cross-channel block.  conv_w [4,E,16],
    conv_b / ln_w / ln_b [4,E] (fp32).  See sodt_frontend_fwd."""
    _require_cuda(x, conv_w, conv_b, ln_w, ln_b)
    if x.dim() != 4 or x.shape[1] != 4 or x.dtype not in _DT:
        raise ValueError("x must be [B, 4, H, W] in fp32 / bf16")
    B, _, H, W = x.shape
    E = conv_w.shape[1]
    out = torch.empty((B, (H - 4) // 4 + 1, (W - 4) // 4 + 1, 4 * E), dtype=x.dtype, device=x.device)
    sb, sc, sy, sx = x.stride()
    with torch.cuda.device(x.device), _Timed(f"frontend[B={B},H={H},W={W}]"):
        st = _capi.lib().sodt_frontend_fwd(x.data_ptr(), sb, sc, sy, sx, conv_w.data_ptr(), conv_b.data_ptr(), ln_w.data_ptr(),
                                           ln_b.data_ptr(), out.data_ptr(), B, H, W, E, pad_r, float(eps), _DT[x.dtype], _stream())
    _capi.check(st, "sodt_frontend_fwd")
    return out


# --------------------------------------------------------------------------------- Detect decode
def frontend_u8(rgb_u8, ir_u8, conv_w, conv_b, ln_w, ln_b, dtype, pad_r=1, eps=1e-5):
    """frontend() straight from the uint8 images: rgb_u8 [B,3,H,W], ir_u8 [B,>=1,H,W] (channel 0 is read); pixels are scaled
    by 1/255 and rounded to ``dtype`` inside the kernel.  See sodt_frontend_u8_fwd."""
    _require_cuda(rgb_u8, ir_u8, conv_w, conv_b, ln_w, ln_b)
    if rgb_u8.dtype != torch.uint8 or ir_u8.dtype != torch.uint8 or rgb_u8.dim() != 4 or rgb_u8.shape[1] != 3 or dtype not in _DT:
        raise ValueError("rgb_u8 must be uint8 [B,3,H,W], ir_u8 uint8 [B,C,H,W]")
    B, _, H, W = rgb_u8.shape
    if ir_u8.shape[0] != B or tuple(ir_u8.shape[2:]) != (H, W):
        raise ValueError("ir_u8 must match rgb_u8 in batch and image size")
    E = conv_w.shape[1]
    out = torch.empty((B, (H - 4) // 4 + 1, (W - 4) // 4 + 1, 4 * E), dtype=dtype, device=rgb_u8.device)
    rb, rc, ry, rx = rgb_u8.stride()
    ib, _, iy, ix = ir_u8.stride()
    with torch.cuda.device(rgb_u8.device), _Timed(f"frontend_u8[B={B},H={H},W={W}]"):
        st = _capi.lib().sodt_frontend_u8_fwd(rgb_u8.data_ptr(), rb, rc, ry, rx, ir_u8.data_ptr(), ib, iy, ix, conv_w.data_ptr(),
                                              conv_b.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), out.data_ptr(), B, H, W, E,
                                              pad_r, float(eps), _DT[dtype], _stream())
    _capi.check(st, "sodt_frontend_u8_fwd")
    return out


USE_FUSED_FRONTEND = True   # uint8 images -> patch-embedded tokens in one tcgen05 kernel (False: frontend_u8 + the embedding GEMM)


def frontend_embed_u8_supported(rgb_u8, ir_u8, E, embed_dim, pos):
    """True if ``frontend_embed_u8`` can run on these uint8 images (CUDA, unit pixel stride, 4-byte aligned rows)."""
    if not (USE_FUSED_FRONTEND and USE_TC_LINEAR and rgb_u8.is_cuda and rgb_u8.dtype == torch.uint8 and ir_u8.dtype == torch.uint8):
        return False
    B, _, H, W = rgb_u8.shape
    rb, rc, ry, rx = rgb_u8.stride()
    ib, _, iy, ix = ir_u8.stride()
    if rx != 1 or ix != 1 or any(v % 4 for v in (rb, rc, ry, ib, iy, rgb_u8.data_ptr(), ir_u8.data_ptr())):
        return False
    pos_rows = 0 if pos is None else pos.numel() // embed_dim
    return bool(_capi.lib().sodt_frontend_embed_u8_supported(B, H, W, E, embed_dim, pos_rows))


def frontend_embed_u8(rgb_u8, ir_u8, conv_w, conv_b, ln_w, ln_b, pe_weight, pe_bias, pos=None, pad_r=1, eps=1e-5, want_stats=False):
    """uint8 RGB [B,3,H,W] + IR [B,>=1,H,W] -> patch-embedded bf16 tokens [B, H/4, W/4, embed_dim] (+ the partial row
    statistics [embed_dim/64, M, 2] for the first norm1): channel embeddings, window-1 cross-channel block, concat, 1x1 patch
    embedding, bias and position embedding in one kernel (sodt_frontend_embed_u8_fwd).  conv_w [4,E,16] / conv_b / ln_w / ln_b
    [4,E] fp32 as for ``frontend``; pe_weight [embed_dim, 4E] (or the conv's [embed_dim, 4E, 1, 1]); pos [1,h,w,embed_dim]."""
    _require_cuda(rgb_u8, ir_u8, conv_w, conv_b, ln_w, ln_b, pe_weight, pe_bias, pos)
    B, _, H, W = rgb_u8.shape
    E = conv_w.shape[1]
    D = pe_weight.shape[0]
    if not frontend_embed_u8_supported(rgb_u8, ir_u8, E, D, pos):
        raise _capi.SodtError("frontend_embed_u8: unsupported geometry / strides (frontend_embed_u8_supported)")
    cw = cached_derived(conv_w, "bf16", lambda t: t.detach().to(torch.bfloat16).contiguous())
    pw = cached_derived(pe_weight, "pe_bf16", lambda t: t.detach().reshape(t.shape[0], -1).to(torch.bfloat16).contiguous())
    pb = _as_f32(pe_bias) if pe_bias is not None else cached_derived(pe_weight, "zero_bias", lambda w: torch.zeros(
        w.shape[0], dtype=torch.float32, device=w.device))
    posr = None
    if pos is not None:
        posr = cached_derived(pos, "pos_bf16", lambda t: t.detach().reshape(-1, D).to(torch.bfloat16).contiguous())
    h, w = H // 4, W // 4
    M = B * h * w
    out = torch.empty((B, h, w, D), dtype=torch.bfloat16, device=rgb_u8.device)
    stats = torch.empty((D // 64, M, 2), dtype=torch.float32, device=rgb_u8.device) if want_stats else None
    rb, rc, ry, _ = rgb_u8.stride()
    ib, _, iy, _ = ir_u8.stride()
    with torch.cuda.device(rgb_u8.device), _Timed(f"frontend_embed_u8[B={B},H={H},W={W}]"):
        st = _capi.lib().sodt_frontend_embed_u8_fwd(rgb_u8.data_ptr(), rb, rc, ry, ir_u8.data_ptr(), ib, iy, cw.data_ptr(),
                                                    conv_b.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), pw.data_ptr(), pb.data_ptr(),
                                                    _ptr(posr), 0 if posr is None else posr.shape[0], out.data_ptr(), _ptr(stats),
                                                    B, H, W, E, D, pad_r, float(eps), _stream())
    _capi.check(st, "sodt_frontend_embed_u8_fwd")
    return (out, stats) if want_stats else out


def detect_decode(raw, anchors_px, stride, want_perm=True, z=None, rows_total=None, row_offset=0):
    """raw [B, na*no, ny, nx] (any strides) -> (z [B, na*ny*nx, no] fp32, x_perm [B,na,ny,nx,no] or None)."""
    _require_cuda(raw, anchors_px, z)
    if raw.dim() != 4 or raw.dtype not in _DT:
        raise ValueError("raw must be a 4-D f32 / bf16 tensor")
    B, ch, ny, nx = raw.shape
    anchors_px = _as_f32(anchors_px).reshape(-1, 2)
    na = anchors_px.shape[0]
    if ch % na:
        raise ValueError("channels must be divisible by the number of anchors")
    no = ch // na
    rows = na * ny * nx
    if z is None:
        rows_total = rows
        z = torch.empty((B, rows, no), dtype=torch.float32, device=raw.device)
    elif rows_total is None:
        rows_total = z.shape[1]
    xp = torch.empty((B, na, ny, nx, no), dtype=raw.dtype, device=raw.device) if want_perm else None
    sb, sc, sy, sx = raw.stride()
    with torch.cuda.device(raw.device), _Timed(f"detect_decode[B={B},ny={ny},nx={nx}]"):
        st = _capi.lib().sodt_detect_decode(raw.data_ptr(), sb, sc, sy, sx, anchors_px.data_ptr(), z.data_ptr(), _ptr(xp),
                                            B, na, no, ny, nx, float(stride), rows_total, row_offset, _DT[raw.dtype],
                                            _stream())
    _capi.check(st, "sodt_detect_decode")
    return z, xp


# ----------------------------------------------------------------------- Linear with fused epilogue
USE_TC_LINEAR = True    # bf16 Linear layers on the tcgen05 kernel (False: cuBLAS + separate elementwise passes)


_LIN_ACT = {None: 0, "none": 0, "gelu": 1, "silu": 2}


def _rows(t):
    """(t2, ld): ``t`` seen as rows of its last dim with one uniform row stride ``ld`` (elements); copies only if the
    strides do not collapse (column slices of contiguous tensors collapse)."""
    if t.stride(-1) != 1:
        t = t.contiguous()
    ld = t.stride(-2) if t.dim() > 1 else t.shape[-1]
    ok = ld % 8 == 0 and ld >= t.shape[-1] and t.data_ptr() % 16 == 0
    for d in range(t.dim() - 2):
        if t.shape[d] != 1 and t.stride(d) != t.stride(d + 1) * t.shape[d + 1]:
            ok = False
    if not ok:
        t = t.contiguous()
        ld = t.shape[-1]
    return t, ld


def _torch_act(y, act):
    if act == "gelu":
        return torch.nn.functional.gelu(y)
    if act == "silu":
        return torch.nn.functional.silu(y)
    if act not in (None, "none"):
        raise ValueError(f"unsupported activation {act!r}")
    return y


USE_LN_FOLD = True      # LayerNorm -> Linear pairs of the Swin blocks as one GEMM (statistics from the producing GEMM's epilogue)

_fold_cache = {}


def fold_layernorm(weight, bias, ln_weight, ln_bias):
    """(W', colsum, b') for ``Linear(LayerNorm(x))`` as a GEMM on the raw rows (see sodt_linear_ln_fwd):
    W' = W diag(ln_weight) in bf16, colsum = row sums of the bf16 W' in fp32, b' = bias + W ln_bias in fp32.
    Cached per parameter objects / versions / storage."""
    srcs = (weight, bias, ln_weight, ln_bias)
    key = tuple((id(t), t._version, t.data_ptr()) if t is not None else None for t in srcs)
    hit = _fold_cache.get(id(weight))
    if hit is not None and hit[0] == key and all(r() is t for r, t in zip(hit[1], srcs) if t is not None):
        return hit[2]
    with torch.no_grad():
        w32 = weight.detach().float()
        wf = (w32 * ln_weight.detach().float()[None, :]).to(torch.bfloat16).contiguous()
        colsum = wf.float().sum(dim=1).contiguous()
        b2 = w32 @ ln_bias.detach().float()
        if bias is not None:
            b2 = b2 + bias.detach().float()
        val = (wf, colsum, b2.contiguous())
    if len(_fold_cache) > 1024:
        _fold_cache.clear()
    _fold_cache[id(weight)] = (key, tuple(weakref.ref(t) if t is not None else None for t in srcs), val)
    return val


def row_stats(x, eps=1e-5):
    """[M, 2] fp32 (mean, rstd) of every row of bf16 ``x`` [..., C] (input of a LayerNorm folded into a GEMM)."""
    _require_cuda(x)
    if x.dtype != torch.bfloat16:
        raise TypeError("row_stats: bf16 only")
    xa, ld = _rows(x)
    C = x.shape[-1]
    M = x.numel() // C
    st = torch.empty((M, 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device), _Timed(f"row_stats[rows={M},C={C}]"):
        rc = _capi.lib().sodt_row_stats(xa.data_ptr(), ld, st.data_ptr(), M, C, float(eps), 1, _stream())
    _capi.check(rc, "sodt_row_stats")
    return st


def finalize_stats(partials, C, eps=1e-5):
    """[boxes, M, 2] partial (sum, sum of squares) from ``linear(..., want_stats=True)`` -> [M, 2] (mean, rstd)."""
    _require_cuda(partials)
    if partials.dtype != torch.float32 or partials.dim() != 3 or partials.shape[2] != 2 or not partials.is_contiguous():
        raise ValueError("partials must be contiguous fp32 [boxes, M, 2]")
    boxes, M, _ = partials.shape
    st = torch.empty((M, 2), dtype=torch.float32, device=partials.device)
    with torch.cuda.device(partials.device), _Timed(f"stats_finalize[rows={M},boxes={boxes}]"):
        rc = _capi.lib().sodt_stats_finalize(partials.data_ptr(), boxes, st.data_ptr(), M, C, float(eps), _stream())
    _capi.check(rc, "sodt_stats_finalize")
    return st


def linear_ln_supported(x, n_out):
    """True if ``linear(x, ..., ln=...)`` / ``want_stats`` can run (bf16 CUDA rows the tcgen05 GEMM covers)."""
    K = x.shape[-1]
    return bool(USE_TC_LINEAR and USE_LN_FOLD and x.is_cuda and x.dtype == torch.bfloat16
                and _capi.lib().sodt_linear_supported(x.numel() // K, n_out, K, 1))


def linear(x, weight, bias=None, act=None, residual=None, x2=None, out=None, ln=None, want_stats=False):
    """act(cat(x, x2) @ weight.T + bias) (+ residual) over the last dim.
    ``ln=(stats, ln_weight, ln_bias[, eps])``: computes ``Linear(LayerNorm(x))`` from the raw rows ``x`` and their statistics:
    ``stats`` = [M, 2] (mean, rstd) from row_stats / finalize_stats, or the [boxes <= 6, M, 2] partial sums a ``want_stats``
    GEMM emitted (reduced in the epilogue with ``eps``); the LayerNorm is folded into the weights (cached) and the epilogue.
    ``want_stats``: also return the [N/64, M, 2] partial (sum, sum of squares) of the result rows -> (out, partials).
    Both need a shape for which linear_ln_supported() holds.  bf16 shapes the tcgen05 kernel supports run there with the
    activation / residual fused into the epilogue; ``x``, ``x2``, ``residual`` and ``out`` may be column slices of wider
    tensors (row-strided).  fp32 and other shapes use cuBLAS (torch) plus elementwise ops."""
    _require_cuda(x, weight, bias, residual, x2, out)
    if act not in _LIN_ACT:
        raise ValueError(f"unsupported activation {act!r}")
    K1 = x.shape[-1]
    K = K1 + (x2.shape[-1] if x2 is not None else 0)
    N = weight.shape[0]
    M = x.numel() // K1
    if weight.shape[1] != K:
        raise ValueError("weight must be [N, K]")
    if (USE_TC_LINEAR and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and (x2 is None or K1 % 64 == 0)
            and _capi.lib().sodt_linear_supported(M, N, K, 1)):
        xa, ldx = _rows(x)
        xb, ldx2 = _rows(x2) if x2 is not None else (None, 0)
        res, ldr, res_rows = None, 0, 0
        if residual is not None:
            if residual.dtype != torch.bfloat16 or residual.shape[-1] != N:
                raise ValueError("residual must be bf16 with N columns")
            if residual.numel() == M * N:
                res, ldr = _rows(residual.reshape(x.shape[:-1] + (N,)) if residual.dim() != x.dim() else residual)
            else:       # broadcast over the leading (batch) dims: the residual repeats every res_rows rows
                res_rows = residual.numel() // N
                if res_rows % 128 or M % res_rows:
                    raise ValueError("a broadcast residual must have a multiple of 128 rows that divides M")
                res, ldr = residual.reshape(res_rows, N).contiguous(), N
        if out is None:
            out = torch.empty(x.shape[:-1] + (N,), dtype=torch.bfloat16, device=x.device)
            ldo = N
        else:
            o2, ldo = _rows(out)
            if o2 is not out or out.dtype != torch.bfloat16 or out.numel() != M * N:
                raise ValueError("out must be a bf16 row-strided view with M*N elements")
        if (ln is not None or want_stats) and x2 is not None:
            raise ValueError("ln / want_stats cannot be combined with x2")
        mr, ln_w, ln_b, ln_eps = (tuple(ln) + (1e-5,))[:4] if ln is not None else (None, None, None, 0.0)
        ln_boxes = 0
        if mr is not None:
            if mr.dtype != torch.float32 or not mr.is_contiguous() or tuple(mr.shape[-2:]) != (M, 2) or mr.dim() > 3:
                raise ValueError("ln statistics must be contiguous fp32 [M, 2] or [boxes, M, 2]")
            if mr.dim() == 3:
                ln_boxes = mr.shape[0]
                if ln_boxes > 6:
                    raise ValueError("more than 6 partial pairs per row: reduce them with finalize_stats first")
        stats_out = torch.empty((N // 64, M, 2), dtype=torch.float32, device=x.device) if want_stats else None
        label = f"linear[M={M},N={N},K={K},act={act},res={residual is not None},ln={ln is not None},stats={want_stats}]"
        if ln is not None or want_stats:
            w, colsum, b32 = (fold_layernorm(weight, bias, ln_w, ln_b) if ln is not None
                              else (weight.detach().contiguous(), None, _as_f32(bias)))
            with torch.cuda.device(x.device), _Timed(label):
                st = _capi.lib().sodt_linear_ln_fwd(xa.data_ptr(), ldx, _ptr(mr), ln_boxes, float(ln_eps), _ptr(colsum), w.data_ptr(), _ptr(b32),
                                                    _ptr(res), ldr, res_rows, out.data_ptr(), ldo, _ptr(stats_out), M, N, K,
                                                    _LIN_ACT[act], 1, _stream())
            _capi.check(st, "sodt_linear_ln_fwd")
        else:
            w = weight.detach().contiguous()
            b32 = _as_f32(bias)
            with torch.cuda.device(x.device), _Timed(label):
                st = _capi.lib().sodt_linear_strided_fwd(xa.data_ptr(), ldx, _ptr(xb), ldx2, K1 if x2 is not None else 0, w.data_ptr(),
                                                         _ptr(b32), _ptr(res), ldr, res_rows, out.data_ptr(), ldo, M, N, K,
                                                         _LIN_ACT[act], 1, _stream())
            _capi.check(st, "sodt_linear_strided_fwd")
        return (out, stats_out) if want_stats else out
    if ln is not None or want_stats:
        raise _capi.SodtError("linear(ln=..., want_stats=...) needs a shape covered by the tcgen05 GEMM (linear_ln_supported)")
    xin = torch.cat((x, x2), dim=-1) if x2 is not None else x
    y = _torch_act(torch.nn.functional.linear(xin, weight, bias), act)
    if residual is not None:
        y = y + (residual.view_as(y) if residual.numel() == y.numel() else residual)
    if out is not None:
        out.copy_(y)
        return out
    return y


# norm1 + qkv Linear + window attention of the C = 192 Swin blocks as one kernel (sodt_attn_block_fwd: no qkv tensor).  Bit-identical
# to the two kernels.  Alone on an idle GPU: 1.13 ms against 0.75 + 0.75 ms (shift 0), 1.41 against 1.51 ms (shift 2: wrapped, masked
# border windows).  Inside the power-capped step it gains more than that difference -- 4.8 GB less HBM traffic per block lets the SM
# clock rise (same box, alternating runs: 820 -> 834 images/s with the un-shifted blocks fused, another +1.1 % with the shifted ones).
# SODT_FUSED_ATTN=0 / SODT_FUSED_ATTN_SHIFTED=0 select the two kernels for all / for the shifted blocks.
USE_FUSED_ATTN = os.environ.get("SODT_FUSED_ATTN", "1") == "1"
FUSED_ATTN_SHIFTED = os.environ.get("SODT_FUSED_ATTN_SHIFTED", "1") == "1"


def attn_block_supported(x, heads, ws, shift):
    """True if ``attn_block`` covers ``x`` [B, H, W, C] (bf16, C = 192, 8 x 8 windows, head_dim 16 / 32)."""
    if not (USE_LN_FOLD and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4):
        return False
    B, H, W, C = x.shape
    return bool(_capi.lib().sodt_attn_block_supported(B, H, W, C, heads, ws, shift, 1))


def attn_block(x, ln, qkv_weight, qkv_bias, bias_table, heads, ws, shift=0, scale=None, mask_value=-100.0):
    """``window_attention(Linear(LayerNorm(x)))`` -> [B, H, W, C] in one kernel (sodt_attn_block_fwd): the qkv tensor never exists.
    ``ln=(stats, ln_weight, ln_bias[, eps])`` as in ``linear``; bit-identical to ``window_attention(linear(x, ..., ln=ln), ...)``."""
    _require_cuda(x, qkv_weight, qkv_bias, bias_table)
    if not attn_block_supported(x, heads, ws, shift):
        raise _capi.SodtError("attn_block: unsupported shape (see sodt_attn_block_fwd)")
    B, H, W, C = x.shape
    M = B * H * W
    mr, ln_w, ln_b, ln_eps = (tuple(ln) + (1e-5,))[:4]
    if mr.dtype != torch.float32 or not mr.is_contiguous() or tuple(mr.shape[-2:]) != (M, 2) or mr.dim() > 3:
        raise ValueError("ln statistics must be contiguous fp32 [M, 2] or [boxes, M, 2]")
    ln_boxes = mr.shape[0] if mr.dim() == 3 else 0
    if ln_boxes > 6:
        raise ValueError("more than 6 partial pairs per row: reduce them with finalize_stats first")
    span = (2 * ws - 1) ** 2
    if tuple(bias_table.shape) != (span, heads):
        raise ValueError(f"bias_table must be [{span}, {heads}], got {tuple(bias_table.shape)}")
    x = x.contiguous()
    w, colsum, b32 = fold_layernorm(qkv_weight, qkv_bias, ln_w, ln_b)
    lib = _capi.lib()

    def prepare(t):
        t32 = _as_f32(t)
        buf = torch.empty(max(int(lib.sodt_window_attn_workspace_bytes(C, heads, ws)), 16), dtype=torch.uint8, device=t.device)
        with torch.cuda.device(t.device):
            _capi.check(lib.sodt_window_attn_prepare(t32.data_ptr(), B, H, W, C, heads, ws, shift, 1, buf.data_ptr(), buf.numel(), _stream()),
                        "sodt_window_attn_prepare")
        buf.record_stream(torch.cuda.current_stream(t.device))
        return buf
    kclass = lib.sodt_window_attn_kernel_class(B, H, W, C, heads, ws, shift, 1)
    wsp = cached_derived(bias_table, ("wattn_image", kclass, ws), prepare)
    out = torch.empty((B, H, W, C), dtype=torch.bfloat16, device=x.device)
    if scale is None:
        scale = (C // heads) ** -0.5
    with torch.cuda.device(x.device), _Timed(f"attn_block[B={B},H={H},W={W},C={C},heads={heads},ws={ws},shift={shift}]"):
        st = lib.sodt_attn_block_fwd(x.data_ptr(), mr.data_ptr(), ln_boxes, float(ln_eps), colsum.data_ptr(), w.data_ptr(), b32.data_ptr(),
                                     out.data_ptr(), B, H, W, C, heads, ws, shift, 1, float(scale), float(mask_value),
                                     wsp.data_ptr(), wsp.numel(), _stream())
    _capi.check(st, "sodt_attn_block_fwd")
    return out


USE_FUSED_MLP = True    # fc1 + GELU + fc2 + residual of the linear-MLP Swin blocks as one kernel (hidden stays in TMEM)


def mlp_ln_supported(x, hidden):
    """True if ``mlp_ln`` can run: bf16 CUDA rows of width 64 / 128 / 192, hidden a multiple of 128 (>= 256)."""
    C = x.shape[-1]
    return bool(USE_TC_LINEAR and USE_LN_FOLD and USE_FUSED_MLP and x.is_cuda and x.dtype == torch.bfloat16
                and _capi.lib().sodt_mlp_supported(x.numel() // C, C, hidden, 1))


MLP_HIDDEN_FP16 = False  # fused MLP: True = hidden operand as fp16 2*GELU against 0.5*fc2.weight in fp16 (3 more significand bits, measured
                         # 2 % slower); False = bf16 hidden operand, bit-identical to the fc1 / fc2 GEMM pair


def mlp_ln(x, ln, fc1_weight, fc1_bias, fc2_weight, fc2_bias, want_stats=False, hidden_fp16=None):
    """``x + fc2(GELU(fc1(LayerNorm(x))))`` over the last dim in one kernel (sodt_mlp_ln_fwd): the second half of a Swin block
    with a linear MLP (reference backbone_vit.py:885-890,1128).  ``ln = (stats, ln_weight, ln_bias[, eps])`` as in ``linear``;
    ``want_stats``: also return the [C/64, M, 2] partial row statistics of the result -> (out, partials).
    ``hidden_fp16`` (default MLP_HIDDEN_FP16): see sodt_mlp_ln_fwd's ``w2_fp16``."""
    _require_cuda(x, fc1_weight, fc1_bias, fc2_weight, fc2_bias)
    C = x.shape[-1]
    M = x.numel() // C
    hidden = fc1_weight.shape[0]
    if tuple(fc1_weight.shape) != (hidden, C) or tuple(fc2_weight.shape) != (C, hidden):
        raise ValueError("fc1_weight must be [hidden, C] and fc2_weight [C, hidden]")
    if not mlp_ln_supported(x, hidden) or fc2_weight.dtype != torch.bfloat16:
        raise _capi.SodtError("mlp_ln needs a shape covered by the fused MLP kernel (mlp_ln_supported)")
    mr, ln_w, ln_b, ln_eps = (tuple(ln) + (1e-5,))[:4]
    if mr.dtype != torch.float32 or not mr.is_contiguous() or tuple(mr.shape[-2:]) != (M, 2) or mr.dim() > 3:
        raise ValueError("ln statistics must be contiguous fp32 [M, 2] or [boxes, M, 2]")
    ln_boxes = mr.shape[0] if mr.dim() == 3 else 0
    if ln_boxes > 3:
        raise ValueError("more than 3 partial pairs per row: reduce them with finalize_stats first")
    xa, ldx = _rows(x)
    w1, colsum, b1 = fold_layernorm(fc1_weight, fc1_bias, ln_w, ln_b)
    hidden_fp16 = MLP_HIDDEN_FP16 if hidden_fp16 is None else bool(hidden_fp16)
    w2 = (cached_derived(fc2_weight, "half_fp16", lambda w: (0.5 * w.detach().float()).to(torch.float16).contiguous())
          if hidden_fp16 else fc2_weight.detach().contiguous())
    b2 = _as_f32(fc2_bias) if fc2_bias is not None else cached_derived(fc2_weight, "zero_bias", lambda w: torch.zeros(
        w.shape[0], dtype=torch.float32, device=w.device))
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    stats_out = torch.empty((C // 64, M, 2), dtype=torch.float32, device=x.device) if want_stats else None
    with torch.cuda.device(x.device), _Timed(f"mlp_ln[M={M},C={C},hidden={hidden},stats={want_stats}]"):
        st = _capi.lib().sodt_mlp_ln_fwd(xa.data_ptr(), ldx, mr.data_ptr(), ln_boxes, float(ln_eps), colsum.data_ptr(), w1.data_ptr(),
                                         b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), out.data_ptr(), C, _ptr(stats_out), M, C, hidden,
                                         int(hidden_fp16), 1, _stream())
    _capi.check(st, "sodt_mlp_ln_fwd")
    return (out, stats_out) if want_stats else out


def conv2d_nhwc_supported(x, cout, kh, kw):
    """True if ops.conv2d_nhwc runs ``x`` [B,H,W,Cin] on the tcgen05 tap-GEMM kernel."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and USE_TC_LINEAR):
        return False
    B, H, W, Cin = x.shape
    return bool(_capi.lib().sodt_conv2d_nhwc_supported(B, H, W, Cin, cout, kh, kw, 1))


def conv_weight_taps(weight):
    """Conv2d weight [Cout, Cin, kh, kw] -> the [Cout, kh*kw*Cin] tap-major matrix sodt_conv2d_nhwc_fwd reads (cached)."""
    return cached_derived(weight, "taps", lambda w: w.detach().permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous())


def conv2d_nhwc(x, w_taps, bias, kernel, pad=(0, 0), act=None, out=None):
    """Stride-1 convolution of a channels-last tensor as a tap GEMM: x [B,H,W,Cin] bf16 (may be a channel slice of a wider
    tensor), w_taps = conv_weight_taps(weight), ``pad`` = zero rows above / columns left (rows below / right of the image
    read as zero as far as the kernel reaches).  Returns act(conv + bias) as [B,H,W,Cout]; ``out`` may be a channel slice."""
    _require_cuda(x, w_taps, bias, out)
    B, H, W, Cin = x.shape
    kh, kw = kernel
    Cout = w_taps.shape[0]
    if w_taps.shape[1] != kh * kw * Cin or w_taps.dtype != torch.bfloat16 or x.dtype != torch.bfloat16:
        raise ValueError("w_taps must be bf16 [Cout, kh*kw*Cin]")
    xa, ldx = _rows(x)
    if out is None:
        out = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=x.device)
        ldo = Cout
    else:
        o2, ldo = _rows(out)
        if o2 is not out or out.dtype != torch.bfloat16 or tuple(out.shape) != (B, H, W, Cout):
            raise ValueError("out must be a bf16 [B,H,W,Cout] channel-slice view")
    b32 = _as_f32(bias)
    with torch.cuda.device(x.device), _Timed(f"conv2d_nhwc[B={B},H={H},W={W},Cin={Cin},Cout={Cout},k={kh}x{kw},act={act}]"):
        st = _capi.lib().sodt_conv2d_nhwc_fwd(xa.data_ptr(), ldx, w_taps.data_ptr(), _ptr(b32), out.data_ptr(), ldo, B, H, W, Cin, Cout,
                                              kh, kw, pad[0], pad[1], _LIN_ACT[act], 1, _stream())
    _capi.check(st, "sodt_conv2d_nhwc_fwd")
    return out


def patch_merge_linear(x, weight, bias=None):
    """PatchMerging's gather + reduction: x [B,H,W,C] -> [B, H/2 * W/2, N] with the 2x2 neighbourhood concatenated in the
    reference's channel order (dy,dx) = (0,0),(1,0),(0,1),(1,1) and multiplied by weight [N, 4C]."""
    _require_cuda(x, weight, bias)
    B, H, W, C = x.shape
    N = weight.shape[0]
    if (USE_TC_LINEAR and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
            and _capi.lib().sodt_patch_merge_linear_supported(B, H, W, C, N, 1)):
        xa = x.contiguous()
        w = weight.detach().contiguous()
        out = torch.empty((B, (H // 2) * (W // 2), N), dtype=torch.bfloat16, device=x.device)
        b32 = _as_f32(bias)
        with torch.cuda.device(x.device), _Timed(f"patch_merge_linear[B={B},H={H},W={W},C={C},N={N}]"):
            st = _capi.lib().sodt_patch_merge_linear_fwd(xa.data_ptr(), w.data_ptr(), _ptr(b32), out.data_ptr(), B, H, W, C, N, 1,
                                                         _stream())
        _capi.check(st, "sodt_patch_merge_linear_fwd")
        return out
    g = x.view(B, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 4, 2, 5).reshape(B, (H // 2) * (W // 2), 4 * C)
    return torch.nn.functional.linear(g, weight, bias)


# ------------------------------------------------------------------- fused bias + activation (+ crop)
_ACT = {None: 0, "none": 0, "gelu": 1, "silu": 2}


def bias_act_crop(x_nchw, bias, act, out_hw=None, offset=(0, 0)):
    """x: NCHW-shaped tensor in channels-last memory.  Returns act(x[:, :, oy:oy+H, ox:ox+W] + bias) as a contiguous
    [B, H, W, C] tensor.  See sodt_bias_act_crop_nhwc."""
    _require_cuda(x_nchw, bias)
    if x_nchw.dtype not in _DT:
        raise TypeError(f"unsupported dtype {x_nchw.dtype}")
    B, C, inH, inW = x_nchw.shape
    H, W = out_hw if out_hw is not None else (inH - offset[0], inW - offset[1])
    x = x_nchw.permute(0, 2, 3, 1).contiguous()        # no copy when the tensor is channels-last
    out = torch.empty((B, H, W, C), dtype=x.dtype, device=x.device)
    b32 = _as_f32(bias)
    with torch.cuda.device(x.device), _Timed(f"bias_act_crop[B={B},H={H},W={W},C={C},act={act}]"):
        st = _capi.lib().sodt_bias_act_crop_nhwc(x.data_ptr(), b32.data_ptr(), out.data_ptr(), B, H, W, C, inH, inW,
                                                 offset[0], offset[1], _ACT[act], _DT[x.dtype], _stream())
    _capi.check(st, "sodt_bias_act_crop_nhwc")
    return out


# ------------------------------------------------------------------- head glue: upsample + concat
def upsample2x_concat(low, skip):
    """cat([nearest_upsample_2x(low), skip], dim=1) for NCHW-shaped tensors in channels-last memory.
    Returns an NCHW-shaped channels-last tensor.  See sodt_upsample2x_concat_nhwc."""
    _require_cuda(low, skip)
    B, C1, H, W = low.shape
    B2, C2, H2, W2 = skip.shape
    if (B2, H2, W2) != (B, 2 * H, 2 * W) or low.dtype != skip.dtype:
        raise ValueError("skip must be [B, C2, 2H, 2W] with the dtype of low")
    lo = low.permute(0, 2, 3, 1).contiguous()     # no copy when already channels-last
    sk = skip.permute(0, 2, 3, 1).contiguous()
    out = torch.empty((B, H2, W2, C1 + C2), dtype=low.dtype, device=low.device)
    with torch.cuda.device(low.device), _Timed(f"upsample2x_concat[B={B},H={H},W={W},C1={C1},C2={C2}]"):
        st = _capi.lib().sodt_upsample2x_concat_nhwc(lo.data_ptr(), sk.data_ptr(), out.data_ptr(), B, H, W, C1, C2,
                                                     low.element_size(), _stream())
    _capi.check(st, "sodt_upsample2x_concat_nhwc")
    return out.permute(0, 3, 1, 2)


class UpCat:
    """cat([nearest_upsample_2x(low), skip], dim=1) that has not been materialised: NCHW-shaped channels-last tensors
    ``low`` [B,C1,H/2,W/2] and ``skip`` [B,C2,H,W].  The head's C3 reads it through TMA addressing (upcat_conv1x1);
    anything else calls ``materialize()``."""

    def __init__(self, low, skip):
        self.low, self.skip = low, skip
        self.shape = (skip.shape[0], low.shape[1] + skip.shape[1], skip.shape[2], skip.shape[3])
        self.dtype, self.device, self.is_cuda = skip.dtype, skip.device, skip.is_cuda

    def materialize(self):
        return upsample2x_concat(self.low, self.skip)


def upcat_conv1x1_supported(uc, cout):
    B, _, H, W = uc.shape
    return bool(USE_TC_LINEAR and uc.is_cuda and uc.dtype == torch.bfloat16 and uc.low.dtype == torch.bfloat16
                and _capi.lib().sodt_upcat_conv1x1_supported(B, H, W, uc.low.shape[1], uc.skip.shape[1], cout, 1))


def upcat_conv1x1(uc, weight, bias, act=None):
    """act(conv1x1(cat(up2x(low), skip)) + bias) -> [B,H,W,Cout] without the upsampled / concatenated tensors.
    weight [Cout, C1 + C2] bf16 (input channel order: upsampled first)."""
    _require_cuda(uc.low, uc.skip, weight, bias)
    B, _, H, W = uc.shape
    lo = uc.low.permute(0, 2, 3, 1).contiguous()          # no copy when already channels-last
    sk = uc.skip.permute(0, 2, 3, 1).contiguous()
    C1, C2, Cout = lo.shape[-1], sk.shape[-1], weight.shape[0]
    if weight.shape[1] != C1 + C2 or weight.dtype != torch.bfloat16:
        raise ValueError("weight must be bf16 [Cout, C1 + C2]")
    out = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=sk.device)
    b32 = _as_f32(bias)
    with torch.cuda.device(sk.device), _Timed(f"upcat_conv1x1[B={B},H={H},W={W},C1={C1},C2={C2},Cout={Cout},act={act}]"):
        st = _capi.lib().sodt_upcat_conv1x1_fwd(lo.data_ptr(), sk.data_ptr(), weight.data_ptr(), _ptr(b32), out.data_ptr(), Cout,
                                                B, H, W, C1, C2, Cout, _LIN_ACT[act], 1, _stream())
    _capi.check(st, "sodt_upcat_conv1x1_fwd")
    return out


# ------------------------------------------------------------------------------------------- NMS
_workspaces = {}


def _nms_workspace(device, B, R, nc, multi_label):
    need = _capi.lib().sodt_nms_workspace_bytes(B, R, nc, int(multi_label))
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


_class_filters = {}


def class_filter(classes, device):
    """The ``classes`` filter of non_max_suppression as a device int32 tensor.  A list / tuple is uploaded ONCE per
    (values, device) and cached, so repeated calls enqueue no host-to-device copy (a pageable copy blocks the host and is
    illegal inside CUDA-graph capture); a device int32 tensor is used as it is."""
    if classes is None:
        return None
    if isinstance(classes, torch.Tensor):
        if classes.dtype != torch.int32 or classes.device != device or not classes.is_contiguous():
            raise ValueError("a tensor class filter must be a contiguous int32 tensor on the predictions' device")
        return classes
    key = (tuple(int(c) for c in classes), device.index)
    hit = _class_filters.get(key)
    if hit is None:
        if torch.cuda.is_current_stream_capturing():
            raise _capi.SodtError("class filter not uploaded before CUDA-graph capture: call ops.class_filter(classes, device) first")
        hit = torch.tensor(list(key[0]), dtype=torch.int32).to(device)
        _class_filters[key] = hit
    return hit


def nms(pred, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, merge=True,
        redundant=True, max_det=300, max_nms=30000, max_wh=4096.0, out=None, counts=None, want_keep_idx=False):
    """pred [B,R,5+nc] fp32 -> (out [B,max_det,6] fp32, counts [B] int32, keep_idx [B,max_det] int32 | None).
    ``out`` / ``counts`` may be caller-provided (e.g. slices of a communication buffer).  ``classes``: list / tuple of class
    ids (uploaded once and cached, see class_filter) or a device int32 tensor."""
    _require_cuda(pred, out, counts)
    if pred.dim() != 3 or pred.shape[2] < 6 or pred.dtype != torch.float32:
        raise ValueError("pred must be fp32 [B, R, 5+nc] with nc >= 1")
    pred = pred.contiguous()
    B, R, no = pred.shape
    nc = no - 5
    dev = pred.device
    if out is None:
        out = torch.empty((B, max_det, 6), dtype=torch.float32, device=dev)
    if counts is None:
        counts = torch.empty((B,), dtype=torch.int32, device=dev)
    if not (out.is_contiguous() and counts.is_contiguous()) or out.dtype != torch.float32 or counts.dtype != torch.int32:
        raise ValueError("out must be contiguous fp32 [B,max_det,6] and counts contiguous int32 [B]")
    keep = torch.empty((B, max_det), dtype=torch.int32, device=dev) if want_keep_idx else None
    cls_t = class_filter(classes, dev)
    ws = _nms_workspace(dev, B, R, nc, multi_label)
    with torch.cuda.device(dev), _Timed(f"nms[B={B},R={R},nc={nc}]"):
        st = _capi.lib().sodt_nms(pred.data_ptr(), _ptr(cls_t), 0 if cls_t is None else cls_t.numel(), out.data_ptr(),
                                  counts.data_ptr(), _ptr(keep), ws.data_ptr(), ws.numel(), B, R, nc, float(conf_thres),
                                  float(iou_thres), int(bool(multi_label)), int(bool(agnostic)), int(bool(merge)),
                                  int(bool(redundant)), max_det, max_nms, float(max_wh), _stream())
    _capi.check(st, "sodt_nms")
    return out, counts, keep


# ---------------------------------------------------------------------------- torch.ops.sodt.*
def _register_custom_ops():
    lib = torch.library

    @lib.custom_op("sodt::window_attn_fwd", mutates_args=())
    def window_attn_fwd(qkv: torch.Tensor, bias_table: torch.Tensor, pad_qkv: torch.Tensor, heads: int, ws: int,
                        shift: int, scale: float, mask_value: float) -> torch.Tensor:
        return window_attention(qkv, bias_table, heads, ws, shift, pad_qkv, scale, mask_value)

    @window_attn_fwd.register_fake
    def _(qkv, bias_table, pad_qkv, heads, ws, shift, scale, mask_value):
        B, H, W, C3 = qkv.shape
        return qkv.new_empty((B, H, W, C3 // 3))

    @lib.custom_op("sodt::cattn_block_fwd", mutates_args=())
    def cattn_block_fwd(r: torch.Tensor, g: torch.Tensor, b: torch.Tensor, ir: torch.Tensor, ln_w: torch.Tensor,
                        ln_b: torch.Tensor, heads: int, ws: int, shift: int, eps: float) -> torch.Tensor:
        return cattn_block(r, g, b, ir, ln_w, ln_b, heads, ws, shift, eps)

    @cattn_block_fwd.register_fake
    def _(r, g, b, ir, ln_w, ln_b, heads, ws, shift, eps):
        B, h, w, C = r.shape
        return r.new_empty((B, h, w, 4 * C))

    @lib.custom_op("sodt::detect_decode", mutates_args=())
    def detect_decode_op(raw: torch.Tensor, anchors_px: torch.Tensor, stride: float) -> torch.Tensor:
        return detect_decode(raw, anchors_px, stride, want_perm=False)[0]

    @lib.custom_op("sodt::nms", mutates_args=())
    def nms_op(pred: torch.Tensor, conf_thres: float, iou_thres: float, agnostic: bool, multi_label: bool) -> tuple[torch.Tensor, torch.Tensor]:
        out, counts, _ = nms(pred, conf_thres, iou_thres, None, agnostic, multi_label)
        return out, counts

    @nms_op.register_fake
    def _(pred, conf_thres, iou_thres, agnostic, multi_label):
        return pred.new_empty((pred.shape[0], 300, 6), dtype=torch.float32), pred.new_empty((pred.shape[0],), dtype=torch.int32)

    @detect_decode_op.register_fake
    def _(raw, anchors_px, stride):
        B, ch, ny, nx = raw.shape
        na = anchors_px.numel() // 2
        return raw.new_empty((B, na * ny * nx, ch // na), dtype=torch.float32)


try:  # registration is idempotent per process; a second import of this module must not fail
    _register_custom_ops()
except RuntimeError:
    pass
