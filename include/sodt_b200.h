/*
 * sodt_b200 -- C ABI of the B200-native (sm_100a) attention / Detect / NMS hot path of
 * the cross-channel RGB+IR small-object detector.
 *
 * The reference (Bissmella/Small-object-detection-transformers) is pure Python/PyTorch and
 * has no FFI layer; its only native hook is the dangling WindowProcess /
 * WindowProcessReverse pair behind `fused_window_process` (basics/models/backbone_vit.py:1100,1120).
 * Each entry point below therefore replaces a span of reference Python, cited per function.
 * The binding a reference maintainer would add is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - the caller owns all memory (inputs, outputs, workspaces); nothing is allocated or freed here;
 *   - every function returns a sodt_status (0 = ok, negative = error) and never throws;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises the device;
 *   - `dtype` selects the storage type of activations: SODT_F32 or SODT_BF16.  Softmax, LayerNorm
 *     statistics, Detect decode and all NMS arithmetic are always fp32;
 *   - entry points are re-entrant and keep no mutable global state.
 */
#ifndef SODT_B200_H_
#define SODT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    SODT_OK = 0,
    SODT_ERR_INVALID_ARG = -1,   /* null pointer, non-positive extent, C % heads != 0 ... */
    SODT_ERR_UNSUPPORTED = -2,   /* shape outside what the kernels implement (see each function) */
    SODT_ERR_WORKSPACE = -3,     /* workspace pointer null or too small */
    SODT_ERR_CUDA = -4,          /* a CUDA runtime call failed; see sodt_last_cuda_error() */
    SODT_ERR_ALIGNMENT = -5      /* a pointer is not 16-byte aligned */
} sodt_status;

typedef enum { SODT_F32 = 0, SODT_BF16 = 1 } sodt_dtype;

/* Library / build identification. */
int sodt_version(void);                         /* e.g. 100 = 0.1.0 */
const char* sodt_status_string(int status);     /* static string */
const char* sodt_last_cuda_error(void);         /* static string of the last CUDA error seen by this thread */
int sodt_built_for_sm(void);                    /* 100 (sm_100a) */

/*
 * Window multi-head self-attention on the projected token image.
 * Replaces backbone_vit.py:1094-1123 (torch.roll, window_partition, WindowAttention's
 * score/bias/mask/softmax/AV at :971-989, window_unpartition, reverse roll) -- everything
 * between the qkv Linear (:968) and the proj Linear (:990).  No window tensor, score tensor,
 * bias tensor or mask tensor is materialised.
 *
 *   qkv        [B, H, W, 3*C]  channel order (3, heads, C/heads), un-rolled, un-padded
 *   bias_table [(2*ws-1)^2, heads] fp32 (relative_position_bias_table; the int64 index buffer is
 *              not needed: index = (yi-yj+ws-1)*(2*ws-1) + (xi-xj+ws-1))
 *   pad_qkv    [3*C] or NULL: q/k/v of tokens added by padding H, W up to multiples of ws
 *              (the qkv bias; NULL = zeros).  Only read when H % ws or W % ws is non-zero.
 *   out        [B, H, W, C]
 *   shift      cyclic shift (0 <= shift < ws).  shift > 0 applies the shifted-window mask
 *              (region ids of backbone_vit.py:1060-1072 evaluated in closed form) with
 *              `mask_value` (-100.0 in the reference, finite on purpose).
 *   scale      multiplies q (head_dim^-0.5 in the reference).
 *   workspace  sodt_window_attn_workspace_bytes(C, heads, ws) bytes of device scratch (16-byte aligned)
 *              used by the tensor-core kernels for a transposed, log2(e)-scaled copy of the bias table.
 * Supported: C % heads == 0, head_dim <= 64, any ws >= 1 (tokens per window unbounded).
 * Kernel selection (by shape, one code path per shape): bf16, head_dim 64, ws 32, no shift -> tcgen05
 * flash kernel; bf16, ws 8, head_dim 16 / 32, C % 64 == 0, H and W multiples of 8 -> TMA-fed tcgen05 window-pair kernel;
 * everything else (and all of SODT_F32) -> the exact fp32 CUDA-core kernel.
 */
size_t sodt_window_attn_workspace_bytes(int C, int heads, int ws);
int sodt_window_attn_fwd(const void* qkv, const float* bias_table, const void* pad_qkv, void* out,
                         int B, int H, int W, int C, int heads, int ws, int shift,
                         int dtype, float scale, float mask_value,
                         void* workspace, size_t workspace_bytes, void* stream);
/*
 * The same with the bias-table image prepared once per weight version instead of once per call (the reference gathers
 * table[index] on every forward, backbone_vit.py:974-977): sodt_window_attn_prepare fills `workspace` for the kernel that the
 * shape (B, H, W, C, heads, ws, shift, dtype) selects; sodt_window_attn_fwd_prepared then only reads it (same arguments, same
 * shape).  A workspace prepared for one shape class must not be used with another.
 */
/* 0 = exact CUDA-core kernel (no workspace image), 1 = tcgen05 flash kernel, 2 = TMA-fed tcgen05 window-pair kernel; < 0 = error.
 * Workspaces prepared by sodt_window_attn_prepare are interchangeable between shapes of the same class, heads and ws. */
int sodt_window_attn_kernel_class(int B, int H, int W, int C, int heads, int ws, int shift, int dtype);
int sodt_window_attn_prepare(const float* bias_table, int B, int H, int W, int C, int heads, int ws, int shift, int dtype,
                             void* workspace, size_t workspace_bytes, void* stream);
int sodt_window_attn_fwd_prepared(const void* qkv, const float* bias_table, const void* pad_qkv, void* out,
                                  int B, int H, int W, int C, int heads, int ws, int shift,
                                  int dtype, float scale, float mask_value,
                                  const void* prepared_workspace, size_t workspace_bytes, void* stream);

/*
 * Window attention with a non-standard score epilogue, always on the exact fp32-math kernel:
 *   dense_mask   [mask_windows, N, N] fp32 additive mask or NULL, added after the bias; window w of image b uses entry
 *                (b * nW + w) % mask_windows -- WindowAttention.forward(x, mask) of the reference (backbone_vit.py:979-984,
 *                x = [num_windows * B, N, C] is the image stack [B_, ws, ws, C] of one-window images)
 *   head_scale   [heads] fp32 or NULL: per-head multiplier of the scores instead of `scale`
 *   normalize_qk 1: q and k are L2-normalised over head_dim (eps 1e-12) first.  head_scale + normalize_qk = the SwinV2 cosine
 *                attention, backbone_swinv2.py:895-921, with head_scale = exp(min(logit_scale, ln 100)) and bias_table =
 *                16 sigmoid(cpb_mlp(relative_coords_table)) evaluated once per weight version by the host.
 *   rel_pos_h/w  [2*ws-1, head_dim] fp32 or NULL (both or neither): decomposed relative position embeddings of the SAM-style
 *                global Attention (backbone_vit.py:347-404, add_decomposed_rel_pos :705-740), shared by all heads:
 *                score += q_unscaled . (rel_pos_h[yi-yj+ws-1] + rel_pos_w[xi-xj+ws-1]); bias_table may be all zeros.
 */
int sodt_window_attn_ex_fwd(const void* qkv, const float* bias_table, const void* pad_qkv, void* out,
                            int B, int H, int W, int C, int heads, int ws, int shift,
                            int dtype, float scale, float mask_value,
                            const float* dense_mask, int mask_windows, const float* head_scale, int normalize_qk,
                            const float* rel_pos_h, const float* rel_pos_w, void* stream);

/*
 * Fused (residual add +) LayerNorm over the channels of token rows.
 * Replaces the norm1 / norm2 LayerNorms and the residual adds of SwinTransformerBlock.forward
 * (backbone_vit.py:1089-1090,1125,1128) and PatchMerging.norm (backbone_vit.py:858); eps as given, biased variance,
 * statistics in fp32.
 *
 *   a, r       [rows, C]; r may be NULL.  s = a (+ r)
 *   w, b       [C] fp32 LayerNorm weight / bias
 *   y          [rows, C] = LayerNorm(s)
 *   sum_out    [rows, C] or NULL: s (+ extra_bias).  extra_bias [C] fp32 or NULL is the bias of the NEXT projection,
 *              pre-added to the residual stream so that the projection GEMM can add the residual in its epilogue.
 * Supported: C even, C <= 1024.
 */
int sodt_add_layernorm_fwd(const void* a, const void* r, const float* w, const float* b, const float* extra_bias,
                           void* sum_out, void* y, long long rows, int C, float eps, int dtype, void* stream);

/*
 * Cross-channel attention block over the four token streams R, G, B, IR.
 * Replaces CAttentionBlock.forward, backbone_vit.py:469-561 (and its general-window twin
 * backbone_swinv2.py:429-469): four parameter-free multi-head cross attentions
 * (R<-G, G<-B, B<-IR, IR<-G; CAttention.forward backbone_vit.py:589-616) each followed by
 * LayerNorm_i(stream_i + attn_i).
 *
 *   r,g,b,ir   [B, h, w, C] each, addressed through element strides (sb, sy, sx, sc) shared by
 *              the four streams, so both NHWC tensors and the NCHW outputs of the
 *              channel-embedding convs (backbone_vit.py:196-199,772) are read in place
 *   ln_w, ln_b [4, C] fp32 (norm1..norm4 weight / bias)
 *   out        [B, h, w, 4*C] contiguous: the concatenation of backbone_vit.py:210
 *   ws         window size (1 in the shipped model => attention is the identity on v and the
 *              block is 4 fused add+LayerNorms); shift as for window attention, mask added
 *              BEFORE the 1/sqrt(C/heads) scaling (backbone_vit.py:601-608)
 * Supported: C % heads == 0, C <= 128, ws*ws <= 144.
 */
int sodt_cattn_block_fwd(const void* r, const void* g, const void* b, const void* ir,
                         long long sb, long long sy, long long sx, long long sc,
                         const float* ln_w, const float* ln_b, void* out,
                         int B, int h, int w, int C, int heads, int ws, int shift,
                         float eps, float mask_value, int dtype, void* stream);

/*
 * Fused front end: the four single-channel 4x4 / stride-4 patch embeddings (R with padding `pad_r` = 1, the others 0:
 * backbone_vit.py:69-98,751) + the window-1 cross-channel block (backbone_vit.py:469-561) + the concatenation (:210).
 *   x        [B, 4, H, W] through element strides (sb, sc, sy, sx); channels R, G, B, IR
 *   conv_w   [4, E, 16] fp32 (the four Conv2d(1, E, 4, 4) weights), conv_b [4, E], ln_w / ln_b [4, E]
 *   out      [B, H/4, W/4, 4*E]
 * Supported: E == 48, and H, W such that the padded R stream has the same output size as the others (H, W % 4 == 0).
 */
int sodt_frontend_fwd(const void* x, long long sb, long long sc, long long sy, long long sx,
                      const float* conv_w, const float* conv_b, const float* ln_w, const float* ln_b, void* out,
                      int B, int H, int W, int E, int pad_r, float eps, int dtype, void* stream);
/* The same from the uint8 images of the evaluation loop (basics/test.py:124-130: img / 255; model.py:192: IR channel 0):
 * streams R, G, B = channels 0-2 of `rgb` (strides rb, rc, ry, rx), stream IR = the plane at `ir` (strides ib, iy, ix).
 * Pixels are u8 * (1/255) rounded to `dtype` (the type of `out`), i.e. what the separate conversion passes produced. */
int sodt_frontend_u8_fwd(const void* rgb, long long rb, long long rc, long long ry, long long rx,
                         const void* ir, long long ib, long long iy, long long ix,
                         const float* conv_w, const float* conv_b, const float* ln_w, const float* ln_b, void* out,
                         int B, int H, int W, int E, int pad_r, float eps, int dtype, void* stream);

/*
 * The whole front end as one tcgen05 kernel, straight from the uint8 images (bf16 output only):
 *   sodt_frontend_u8_fwd (four 4x4/4 channel embeddings + window-1 cross-channel block + concat, backbone_vit.py:69-98,
 *   469-561,210) followed by the 1x1 patch embedding Linear(4E -> embed_dim) + bias + absolute position embedding
 *   (backbone_vit.py:212-214) -- the convs run as K = 16 GEMMs, the concat tile stays in shared memory as the A operand of
 *   the patch-embedding GEMM.  rgb uint8 [B,3,H,W] / ir uint8 [B,>=1,H,W] with unit pixel stride and 4-byte aligned rows;
 *   conv_w bf16 [4][E][16] (streams R,G,B,IR; taps ky*4+kx), conv_b / ln_w / ln_b fp32 [4][E]; pe_w bf16 [embed_dim, 4E],
 *   pe_b fp32; pos bf16 [pos_rows = H/4*W/4, embed_dim] (broadcast over the batch) or NULL with pos_rows = 0;
 *   out bf16 [B*H/4*W/4, embed_dim]; stats_out (optional) [embed_dim/64][M][2] partial row statistics for the first norm1.
 *   E = 48, embed_dim = 192, H and W multiples of 4, (H/4*W/4) % 128 == 0 when pos is given.
 */
int sodt_frontend_embed_u8_supported(int B, int H, int W, int E, int embed_dim, int pos_rows);
int sodt_frontend_embed_u8_fwd(const void* rgb, long long rb, long long rc, long long ry, const void* ir, long long ib,
                               long long iy, const void* conv_w, const float* conv_b, const float* ln_w, const float* ln_b,
                               const void* pe_w, const float* pe_b, const void* pos, int pos_rows, void* out,
                               float* stats_out, int B, int H, int W, int E, int embed_dim, int pad_r, float eps, void* stream);
/*
 * YOLOv5 Detect decode for one level.  Replaces model.py:55-64 (view/permute/contiguous,
 * sigmoid, grid + anchor decode, view) for the output of the level's 1x1 conv (model.py:53).
 *
 *   raw        [B, na*no, ny, nx] addressed through element strides (sb, sc, sy, sx)
 *              (NCHW or channels_last)
 *   anchors_px [na, 2] fp32 (Detect.anchor_grid, pixels)
 *   z          fp32 rows: z[(b*rows_total + row_offset + (a*ny + y)*nx + x)*no + o]; with one level
 *              rows_total = na*ny*nx and row_offset = 0 (model.py:64-65 concatenates levels on rows)
 *   x_perm     [B, na, ny, nx, no] in `dtype`, or NULL (the second return value of Detect.forward)
 */
int sodt_detect_decode(const void* raw, long long sb, long long sc, long long sy, long long sx,
                       const float* anchors_px, float* z, void* x_perm,
                       int B, int na, int no, int ny, int nx, float stride,
                       long long rows_total, long long row_offset, int dtype, void* stream);

/*
 * Batched non_max_suppression.  Replaces basics/utils/general.py:425-512 including the
 * torchvision.ops.nms call at :496 (class-offset boxes, greedy suppression in stable
 * descending-score order, IoU = inter / (area_a + area_b - inter) in fp32, suppress iff IoU > thr),
 * the max_det cut, merge-NMS and the `redundant` filter.  The 10 s watchdog (:508-510) is not
 * reproduced.
 *
 *   pred       [B, R, 5+nc] fp32 decoded predictions (cx, cy, w, h, obj, cls...)
 *   classes    [n_classes] int32 class filter or NULL
 *   out        [B, max_det, 6] fp32 rows (x1, y1, x2, y2, conf, cls), descending conf; rows beyond
 *              counts[b] are zero.  May point into an NCCL send buffer (no pack step).
 *   counts     [B] int32
 *   keep_idx   [B, max_det] int32 or NULL: kept candidate indices (reference candidate order after
 *              the max_nms cut) -- the quantity that is bit-exact against the reference
 *   workspace  sodt_nms_workspace_bytes(B, R, nc, multi_label) bytes, 16-byte aligned
 */
size_t sodt_nms_workspace_bytes(int B, int R, int nc, int multi_label);
int sodt_nms(const float* pred, const int* classes, int n_classes, float* out, int* counts, int* keep_idx,
             void* workspace, size_t workspace_bytes, int B, int R, int nc,
             float conf_thres, double iou_thres, int multi_label, int agnostic, int merge, int redundant,
             int max_det, int max_nms, float max_wh, void* stream);

/*
 * bf16 GEMM-shaped layers with fused epilogue on tcgen05 tensor cores (TMA-fed, TMEM accumulators):
 *   out[M, N] = act(A[M, K] . w[N, K]^T + bias[N]) (+ residual[M, N]),  act: 0 identity, 1 GELU, 2 SiLU.
 * w bf16 [N, K] row-major, bias fp32 [N] or NULL, out / residual bf16.  GELU is x Phi(x) with Phi evaluated through one
 * MUFU.TANH on a fitted odd polynomial (|error| < 0.5 |x| 2^-11 against the erf form: below a quarter bf16 ulp), SiLU is
 * 0.5 x (1 + tanh(x/2)).  All: SODT_BF16 only, K % 64 == 0, N % 64 == 0; others -> SODT_ERR_UNSUPPORTED (the host wrapper
 * then uses the cuBLAS / cuDNN library path).
 *
 * sodt_linear_fwd            A = x[M, K] contiguous.  Replaces nn.Linear + the separate GELU pass of Mlp
 *                            (backbone_vit.py:885-890) and nn.Linear + the separate residual add of SwinTransformerBlock
 *                            (backbone_vit.py:968,990,1125,1128).
 * sodt_linear_strided_fwd    the same with row strides (elements, multiples of 8) for x, residual and out, so operands /
 *                            results can be column slices of wider tensors (the head's C3 writes both branches into one
 *                            buffer instead of torch.cat, common.py:114-126); if x2 != NULL, columns [0, k_split) of A come
 *                            from x and [k_split, K) from x2 (neck over concat(block5, block6), backbone_vit.py:239,262).
 *                            res_rows > 0: the residual has res_rows rows (multiple of 128 dividing M) and repeats, i.e. it is
 *                            broadcast over the batch (pos_embed added by the 1x1 patch embedding's GEMM, backbone_vit.py:212-214).
 * sodt_conv2d_nhwc_fwd       stride-1 kh x kw convolution as a tap GEMM on an NHWC tensor: A[(b,y,x), (ky,kx,c)] =
 *                            in[b, y+ky-pad_t, x+kx-pad_l, c] addressed by a rank-4 TMA map, zero padding from the TMA
 *                            out-of-bounds fill; w = conv weight permuted to [Cout, kh, kw, Cin].  Replaces F.pad + Conv2d +
 *                            GELU of the conv-MLP (backbone_vit.py:896-902) and Conv2d + BatchNorm (folded) + SiLU of the
 *                            head (common.py:38-52).  Needs W % 128 == 0, or 128 % W == 0 and H % (128 / W) == 0.
 * sodt_patch_merge_linear_fwd  PatchMerging's 2x2 gather + Linear(4C -> N) (backbone_vit.py:840-853): A[(b,i,j), (dx,dy,c)] =
 *                            x[b, 2i+dy, 2j+dx, c] addressed by a rank-5 TMA map; x bf16 [B, H, W, C] contiguous.
 */
int sodt_linear_supported(int M, int N, int K, int dtype);
int sodt_linear_fwd(const void* x, const void* w, const float* bias, const void* residual, void* out,
                    int M, int N, int K, int act, int dtype, void* stream);
int sodt_linear_strided_fwd(const void* x, int ldx, const void* x2, int ldx2, int k_split, const void* w,
                            const float* bias, const void* residual, int ldr, int res_rows, void* out, int ldo,
                            int M, int N, int K, int act, int dtype, void* stream);
/*
 * LayerNorm folded into the following Linear (norm1 -> attn.qkv, norm2 -> mlp.fc1: backbone_vit.py:1089-1093,1128):
 *   out = act(LayerNorm(x) . W^T + b) (+ residual)  computed as  rstd_row * (x . W'^T) - mean_row * rstd_row * colsum + b'
 * with W' = W * diag(ln_weight) (bf16), colsum[n] = sum_k W'[n,k] (fp32, of the bf16 values), b' = b + W . ln_bias, all
 * prepared by the caller once per weight load.  ln_mean_rstd [M][2] fp32 holds (mean, rstd) of every row of x, from
 *   sodt_row_stats       one read pass over x, or
 *   sodt_stats_finalize  from the partial (sum, sum of squares) pairs, one per 64-column box, [N/64][M][2], that the GEMM
 *                        which WROTE x emitted through `stats_out` of this function (from its fp32 values before the bf16
 *                        rounding) - then no pass over the token tensor is left for the LayerNorm at all.
 * ln_boxes in 1..6: ln_mean_rstd points at that many partial pairs per row instead ([ln_boxes][M][2], i.e. the stats_out of
 * a GEMM with N <= 384) and the epilogue reduces them itself with ln_eps: no sodt_stats_finalize launch.
 * ln_mean_rstd == NULL: a plain Linear that only emits stats_out.  res_rows as in sodt_linear_strided_fwd (the patch
 * embedding adds the batch-broadcast pos_embed and emits the statistics for the first block's norm1).
 */
int sodt_linear_ln_fwd(const void* x, int ldx, const float* ln_mean_rstd, int ln_boxes, float ln_eps, const float* ln_colsum,
                       const void* w, const float* bias, const void* residual, int ldr, int res_rows, void* out, int ldo,
                       float* stats_out, int M, int N, int K, int act, int dtype, void* stream);
/*
 * The linear MLP half of a Swin block as one kernel (backbone_vit.py:885-890 Mlp.forward, :1128 the block's second residual):
 *   out = x + fc2(GELU(fc1(LayerNorm(x))))
 * x bf16 [M, C] is the raw residual stream (LayerNorm input and residual); the LayerNorm is folded exactly as in
 * sodt_linear_ln_fwd (w1 = fc1.weight * diag(ln_weight) in bf16, ln_colsum its fp32 row sums, b1 = fc1.bias + fc1.weight .
 * ln_bias; ln_mean_rstd / ln_boxes as there).  The [M, hidden] activation never reaches HBM: it lives in tensor memory
 * between the two GEMMs.  stats_out (optional) receives the [C/64][M][2] partial row statistics of out for the next
 * block's norm1.  w2_fp16 = 0: w2 is fc2.weight in bf16 and the hidden activation is rounded to bf16 (bit-identical to
 * sodt_linear_ln_fwd(GELU) followed by sodt_linear_ln_fwd(residual)).  w2_fp16 = 1: w2 is 0.5 * fc2.weight in IEEE fp16 and the
 * hidden operand is the fp16 value of 2 GELU(.) (11-bit significand instead of 8; the GELU is evaluated in packed fp16
 * either way, so the fp16 form skips two conversions per pair).  C in {64, 128, 192}, hidden a multiple of 128 and >= 256,
 * bf16 activations only; out may alias neither x nor weights.
 */
/*
 * The attention half of a Swin block up to the projection as one kernel (backbone_vit.py:1090-1123: norm1, roll, window
 * partition, WindowAttention's qkv Linear :968 and attention core :969-989, reverse partition / roll):
 *   out = window_attention( LayerNorm(x) w_qkv^T + b_qkv )
 * x bf16 [B, H, W, C] is the raw residual stream; the LayerNorm is folded exactly as in sodt_linear_ln_fwd (w_qkv = qkv.weight *
 * diag(ln_weight) in bf16 [3C, C], ln_colsum its fp32 row sums, b_qkv = qkv.bias + qkv.weight . ln_bias; ln_mean_rstd / ln_boxes
 * as there, indexed by the token's row in [B*H*W]).  The [B, H, W, 3C] qkv tensor never exists; the result is bit-identical to
 * sodt_linear_ln_fwd followed by sodt_window_attn_fwd_prepared.  `workspace` = the bias-table image sodt_window_attn_prepare
 * wrote for the same (C, heads, ws = 8).  8 x 8 windows, C = 192 with head_dim 16 (12 heads), H % 8 == 0, W % 16 == 0, 0 <= shift < 8,
 * bf16 only; proj + residual remain sodt_linear_ln_fwd.
 */
int sodt_attn_block_supported(int B, int H, int W, int C, int heads, int ws, int shift, int dtype);
int sodt_attn_block_fwd(const void* x, const float* ln_mean_rstd, int ln_boxes, float ln_eps, const float* ln_colsum,
                        const void* w_qkv, const float* b_qkv, void* out, int B, int H, int W, int C, int heads, int ws, int shift,
                        int dtype, float scale, float mask_value, const void* workspace, size_t workspace_bytes, void* stream);
int sodt_mlp_supported(int M, int C, int hidden, int dtype);
int sodt_mlp_ln_fwd(const void* x, int ldx, const float* ln_mean_rstd, int ln_boxes, float ln_eps, const float* ln_colsum,
                    const void* w1, const float* b1, const void* w2, const float* b2, void* out, int ldo,
                    float* stats_out, int M, int C, int hidden, int w2_fp16, int dtype, void* stream);
int sodt_row_stats(const void* x, long long ld, float* mean_rstd, long long rows, int C, float eps, int dtype, void* stream);
int sodt_stats_finalize(const float* partials, int boxes, float* mean_rstd, long long rows, int C, float eps, void* stream);
int sodt_conv2d_nhwc_supported(int B, int H, int W, int Cin, int Cout, int kh, int kw, int dtype);
int sodt_conv2d_nhwc_fwd(const void* x, int ldx, const void* w, const float* bias, void* out, int ldo,
                         int B, int H, int W, int Cin, int Cout, int kh, int kw, int pad_t, int pad_l,
                         int act, int dtype, void* stream);
/* 1x1 convolution over cat(nearest_upsample_2x(low), skip) on the channel axis -- the head's "nn.Upsample, Concat, C3.cv1/cv2"
 * rows (models/model.yaml head rows 1-3 and 5-7; common.py:114-126) -- without the upsampled or the concatenated tensor:
 * the first C1 input channels of pixel (y, x) are read from low[b, y/2, x/2, :] through a TMA map with zero-stride
 * duplicate dimensions, the other C2 from skip[b, y, x, :].  low [B, H/2, W/2, C1], skip [B, H, W, C2], w [Cout, C1 + C2],
 * out [B, H, W, Cout] (row stride ldo), all bf16 channels-last; H, W = the output size. */
int sodt_upcat_conv1x1_supported(int B, int H, int W, int C1, int C2, int Cout, int dtype);
int sodt_upcat_conv1x1_fwd(const void* low, const void* skip, const void* w, const float* bias, void* out, int ldo,
                           int B, int H, int W, int C1, int C2, int Cout, int act, int dtype, void* stream);
int sodt_patch_merge_linear_supported(int B, int H, int W, int C, int N, int dtype);
int sodt_patch_merge_linear_fwd(const void* x, const void* w, const float* bias, void* out, int B, int H, int W,
                                int C, int N, int dtype, void* stream);

/*
 * Fused bias + activation (+ crop) on channels-last activations: out[b,y,x,c] = act(in[b, y+off_y, x+off_x, c] + bias[c]),
 * act: 0 identity, 1 exact (erf) GELU, 2 SiLU.  in is [B, in_H, in_W, C], out [B, H, W, C], bias [C] fp32.
 * Replaces, after the 2x2 conv of the conv-enhanced MLP (backbone_vit.py:896-902), the F.pad copy, the conv's separate
 * bias-add pass and the GELU pass (the conv runs with padding 1 and no bias; row / column 0 are cropped here); also the
 * bias + SiLU of the head's fused Conv (common.py:38-50).  C multiple of 8 (bf16) / 4 (fp32).
 */
int sodt_bias_act_crop_nhwc(const void* in, const float* bias, void* out, int B, int H, int W, int C,
                            int in_H, int in_W, int off_y, int off_x, int act, int dtype, void* stream);

/*
 * Head glue: nearest-neighbour 2x upsample of `low` [B,H,W,C1] concatenated with `skip` [B,2H,2W,C2] on the
 * channel axis, channels-last memory, one pass.  Replaces nn.Upsample(None, 2, 'nearest') + Concat(1) of the
 * detector head (models/model.yaml head rows 1-2 and 5-6).  C1*elem_bytes and C2*elem_bytes multiples of 16.
 */
int sodt_upsample2x_concat_nhwc(const void* low, const void* skip, void* out, int B, int H, int W, int C1, int C2,
                                int elem_bytes, void* stream);

/* Number of kernels launched by this library on the calling thread since the last reset
 * (bench.py reports it as gpu_launches). */
long long sodt_launch_count(void);
void sodt_reset_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SODT_B200_H_ */
