// Generic fused window attention (any window size, head_dim <= 64, fp32 math on CUDA cores).
//
// This is the exact-arithmetic path: it serves SODT_F32 (the 1e-5 parity mode, where tensor
// cores are not an option) and every shape the tcgen05 kernels do not specialise.
// One CTA = one (window, head, 64-query block); keys/values stream through shared memory in
// blocks of 64 with an online softmax, so no score matrix ever reaches HBM.  Roll, partition,
// padding, reverse partition and reverse roll are pure address arithmetic (WinGeom).
//
// Reference semantics: basics/models/backbone_vit.py:971-989 (scale on q, + bias, + mask, softmax, AV),
// :1094-1123 (roll / partition / unpartition), :1058-1079 (mask).
// Score variants (struct ScoreExtras): an explicit dense additive mask [mask_windows, N, N] (WindowAttention.forward(x, mask),
// backbone_vit.py:979-984); the SwinV2 cosine attention (backbone_swinv2.py:895-921): q and k L2-normalised per head
// (F.normalize, eps 1e-12) and a per-head logit scale instead of head_dim^-0.5; the decomposed relative position terms of the
// SAM-style global Attention (backbone_vit.py:347-404,705-740), q . (Rh[dy] + Rw[dx]) with the UNSCALED q.
#include "common.cuh"

namespace sodt {

namespace {

constexpr int BQ = 64;
constexpr int BK = 64;
constexpr int THREADS = 256;
constexpr int MAX_TABLE_SMEM = 4096;  // floats; (2*32-1)^2 = 3969 fits

template <typename T>
__device__ __forceinline__ void load_vec4(const T* p, bool vec_ok, int nvalid, float out[4]) {
    if (vec_ok && nvalid == 4) {
        if constexpr (sizeof(T) == 4) {
            float4 v = *reinterpret_cast<const float4*>(p);
            out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
        } else {
            uint2 raw = *reinterpret_cast<const uint2*>(p);
            __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
            __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
            out[0] = __low2float(a); out[1] = __high2float(a);
            out[2] = __low2float(b); out[3] = __high2float(b);
        }
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) out[e] = e < nvalid ? to_f32<T>(p[e]) : 0.f;
    }
}

// Stage 64 tokens x hd of q, k or v (which = 0,1,2) of one head into smem as fp32 [64][HDP+4].
template <typename T, int HDP>
__device__ __forceinline__ void stage_tile(float* dst, const T* __restrict__ qkv, const T* __restrict__ pad_qkv,
                                           const WinGeom& g, int b, int win, int t0, int N, int C, int hd,
                                           int head, int which, bool vec_ok) {
    constexpr int LD = HDP + 4;
    constexpr int CH = HDP / 4;
    for (int idx = threadIdx.x; idx < BK * CH; idx += THREADS) {
        int row = idx / CH, c4 = idx - row * CH;
        int t = t0 + row;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        int d = c4 * 4;
        int nvalid = hd - d; nvalid = nvalid > 4 ? 4 : nvalid;
        if (t < N && nvalid > 0) {
            int yr, xr;
            long long coff = (long long)which * C + head * hd + d;
            if (g.rolled(win, t, yr, xr)) {
                int ys, xs;
                g.source(yr, xr, ys, xs);
                const T* p = qkv + (((long long)b * g.H + ys) * g.W + xs) * (3LL * C) + coff;
                load_vec4<T>(p, vec_ok, nvalid, v);
            } else if (pad_qkv != nullptr) {
                load_vec4<T>(pad_qkv + coff, vec_ok, nvalid, v);
            }
        }
        *reinterpret_cast<float4*>(dst + row * LD + d) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

struct ScoreExtras {
    const float* dense_mask;   // [mask_windows][N][N] additive, or null; window id = (b * nW + win) % mask_windows
    int mask_windows;
    const float* head_scale;   // [heads] multiplier of the (normalised) scores, or null (use `scale`)
    int normalize_qk;          // 1: q, k rows are L2-normalised over head_dim before the dot product
    const float* rel_pos_h;    // [2 ws - 1][hd] or null: decomposed relative position embeddings (SAM / MViTv2 Attention,
    const float* rel_pos_w;    //   backbone_vit.py:705-740): score += q_i . (rel_pos_h[yi - yj + ws - 1] + rel_pos_w[xi - xj + ws - 1])
};

template <typename T, int HDP, bool kExact>
__global__ void __launch_bounds__(THREADS)
window_attn_generic_kernel(const T* __restrict__ qkv, const float* __restrict__ table,
                           const T* __restrict__ pad_qkv, T* __restrict__ out,
                           int H, int W, int C, int heads, int hd, int ws, int shift,
                           float scale, float mask_value, int table_in_smem, int vec_ok, ScoreExtras ex) {
    constexpr int LD = HDP + 4;
    constexpr int DPT = HDP / 4;  // output dims per thread
    extern __shared__ __align__(16) float smem[];
    float* ks = smem;                  // [BK][LD]   (also stages the query tile once)
    float* vs = ks + BK * LD;          // [BK][LD]
    float* ps = vs + BK * LD;          // [BQ][BK+1]
    int* kmeta = reinterpret_cast<int*>(ps + BQ * (BK + 1));  // [BK] ty | tx<<10 | region<<20
    float* knorm = reinterpret_cast<float*>(kmeta + BK);      // [BK] 1 / |k| (cosine attention only)
    float* tab = knorm + BK;                                  // [(2ws-1)^2] this head's bias column

    const WinGeom g(H, W, ws, shift);
    const int N = ws * ws;
    const int nW = g.nwh * g.nww;
    const int head = blockIdx.y;
    const int b = blockIdx.z / nW;
    const int win = blockIdx.z - b * nW;
    const int tid = threadIdx.x;
    const int i = tid >> 2, part = tid & 3;
    const int tq = blockIdx.x * BQ + i;
    const int span = 2 * ws - 1;

    if (table_in_smem) {
        for (int e = tid; e < span * span; e += THREADS) tab[e] = table[(long long)e * heads + head];
    }
    stage_tile<T, HDP>(ks, qkv, pad_qkv, g, b, win, blockIdx.x * BQ, N, C, hd, head, 0, vec_ok != 0);
    __syncthreads();
    float q[HDP];
    float q_unscale = 1.f;         // q_raw = q * q_unscale (decomposed relative position terms)
    {
        float qs = ex.head_scale != nullptr ? ex.head_scale[head] : scale;
        q_unscale = 1.f / qs;
        if (ex.normalize_qk) {
            float n2 = 0.f;
#pragma unroll
            for (int d = 0; d < HDP; ++d) n2 = fmaf(ks[i * LD + d], ks[i * LD + d], n2);
            qs /= fmaxf(sqrtf(n2), 1e-12f);
        }
#pragma unroll
        for (int d = 0; d < HDP; ++d) q[d] = ks[i * LD + d] * qs;  // reference scales q first (:971)
    }
    int yq = 0, xq = 0;
    const bool q_real = tq < N && g.rolled(win, tq, yq, xq);
    const int tq_c = tq < N ? tq : N - 1;   // rows past the window only pad the tile; keep their bias index in range
    const int tyq = tq_c / ws, txq = tq_c - tyq * ws;
    const int rq = shift > 0 ? g.region(yq, xq) : 0;
    __syncthreads();

    float m_run = -INFINITY, l_run = 0.f;
    float o[DPT];
#pragma unroll
    for (int e = 0; e < DPT; ++e) o[e] = 0.f;

    for (int k0 = 0; k0 < N; k0 += BK) {
        stage_tile<T, HDP>(ks, qkv, pad_qkv, g, b, win, k0, N, C, hd, head, 1, vec_ok != 0);
        stage_tile<T, HDP>(vs, qkv, pad_qkv, g, b, win, k0, N, C, hd, head, 2, vec_ok != 0);
        if (ex.normalize_qk) {
            __syncthreads();
            if (tid < BK) {
                float n2 = 0.f;
#pragma unroll
                for (int d = 0; d < HDP; ++d) n2 = fmaf(ks[tid * LD + d], ks[tid * LD + d], n2);
                knorm[tid] = 1.f / fmaxf(sqrtf(n2), 1e-12f);
            }
        }
        if (tid < BK) {
            int t = k0 + tid, yr = 0, xr = 0;
            int ty = t / ws, tx = t - ty * ws;
            int reg = 0;
            if (t < N) { g.rolled(win, t, yr, xr); reg = shift > 0 ? g.region(yr, xr) : 0; }
            kmeta[tid] = ty | (tx << 10) | (reg << 20);
        }
        __syncthreads();

        float s[16];
        float blk_max = -INFINITY;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int j = part + 4 * e;
            const float* kr = ks + j * LD;
            float acc = 0.f;
#pragma unroll
            for (int d = 0; d < HDP; d += 4) {
                float4 kv = *reinterpret_cast<const float4*>(kr + d);
                acc = fmaf(q[d], kv.x, acc);
                acc = fmaf(q[d + 1], kv.y, acc);
                acc = fmaf(q[d + 2], kv.z, acc);
                acc = fmaf(q[d + 3], kv.w, acc);
            }
            if (ex.normalize_qk) acc *= knorm[j];
            if (k0 + j < N) {
                const int meta = kmeta[j];
                const int tyk = meta & 1023, txk = (meta >> 10) & 1023, rk = meta >> 20;
                if (ex.rel_pos_h != nullptr) {        // the reference multiplies the UNSCALED q (add_decomposed_rel_pos gets q, not q * scale)
                    const float* rh = ex.rel_pos_h + (long long)(tyq - tyk + ws - 1) * hd;
                    const float* rw = ex.rel_pos_w + (long long)(txq - txk + ws - 1) * hd;
                    float rel = 0.f;
                    for (int d = 0; d < hd; ++d) rel = fmaf(q[d], __ldg(rh + d) + __ldg(rw + d), rel);
                    acc += rel * q_unscale;
                }
                const int bidx = (tyq - tyk + ws - 1) * span + (txq - txk + ws - 1);
                const float bias = table_in_smem ? tab[bidx] : __ldg(table + (long long)bidx * heads + head);
                acc += bias;
                if (rk != rq) acc += mask_value;
                if (ex.dense_mask != nullptr)
                    acc += __ldg(ex.dense_mask + ((long long)(blockIdx.z % ex.mask_windows) * N + tq_c) * N + (k0 + j));
            } else {
                acc = -INFINITY;
            }
            s[e] = acc;
            blk_max = fmaxf(blk_max, acc);
        }
        blk_max = fmaxf(blk_max, __shfl_xor_sync(0xffffffffu, blk_max, 1));
        blk_max = fmaxf(blk_max, __shfl_xor_sync(0xffffffffu, blk_max, 2));
        const float m_new = fmaxf(m_run, blk_max);
        const float corr = kExact ? expf(m_run - m_new) : __expf(m_run - m_new);
        float blk_sum = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const float p = kExact ? expf(s[e] - m_new) : __expf(s[e] - m_new);
            blk_sum += p;
            ps[i * (BK + 1) + part + 4 * e] = p;
        }
        blk_sum += __shfl_xor_sync(0xffffffffu, blk_sum, 1);
        blk_sum += __shfl_xor_sync(0xffffffffu, blk_sum, 2);
        l_run = l_run * corr + blk_sum;
        m_run = m_new;
#pragma unroll
        for (int e = 0; e < DPT; ++e) o[e] *= corr;
        __syncwarp();
        const float* pr = ps + i * (BK + 1);
        const int d0 = part * DPT;
#pragma unroll 8
        for (int j = 0; j < BK; ++j) {
            const float p = pr[j];
            const float* vr = vs + j * LD + d0;
            if constexpr (DPT >= 4) {
#pragma unroll
                for (int e = 0; e < DPT; e += 4) {
                    float4 vv = *reinterpret_cast<const float4*>(vr + e);
                    o[e] = fmaf(p, vv.x, o[e]);
                    o[e + 1] = fmaf(p, vv.y, o[e + 1]);
                    o[e + 2] = fmaf(p, vv.z, o[e + 2]);
                    o[e + 3] = fmaf(p, vv.w, o[e + 3]);
                }
            } else {
#pragma unroll
                for (int e = 0; e < DPT; ++e) o[e] = fmaf(p, vr[e], o[e]);
            }
        }
        __syncthreads();
    }

    if (q_real) {
        int ys, xs;
        g.source(yq, xq, ys, xs);
        T* dst = out + (((long long)b * H + ys) * W + xs) * C + head * hd;
        const float inv = 1.f / l_run;
        const int d0 = part * DPT;
#pragma unroll
        for (int e = 0; e < DPT; ++e)
            if (d0 + e < hd) dst[d0 + e] = from_f32<T>(o[e] * inv);
    }
}

template <typename T, int HDP, bool kExact>
int launch(const void* qkv, const float* table, const void* pad_qkv, void* out, int B, int H, int W, int C,
           int heads, int hd, int ws, int shift, float scale, float mask_value, const ScoreExtras& ex, cudaStream_t stream) {
    const WinGeom g(H, W, ws, shift);
    const int N = ws * ws;
    const int span = 2 * ws - 1;
    const int table_in_smem = span * span <= MAX_TABLE_SMEM;
    size_t smem = (size_t)(2 * BK * (HDP + 4) + BQ * (BK + 1)) * sizeof(float) + BK * (sizeof(int) + sizeof(float)) +
                  (table_in_smem ? (size_t)span * span * sizeof(float) : 0);
    auto kern = window_attn_generic_kernel<T, HDP, kExact>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    const int vec_ok = (hd % 4 == 0) && (C % 4 == 0);
    long long nwin = (long long)B * g.nwh * g.nww;
    if (nwin > 2147483647LL || heads > 65535) return SODT_ERR_UNSUPPORTED;
    dim3 grid((N + BQ - 1) / BQ, heads, (unsigned)nwin);
    kern<<<grid, THREADS, smem, stream>>>(static_cast<const T*>(qkv), table, static_cast<const T*>(pad_qkv),
                                          static_cast<T*>(out), H, W, C, heads, hd, ws, shift, scale, mask_value,
                                          table_in_smem, vec_ok, ex);
    return check_launch();
}

template <typename T, bool kExact>
int dispatch_hd(const void* qkv, const float* table, const void* pad_qkv, void* out, int B, int H, int W, int C,
                int heads, int hd, int ws, int shift, float scale, float mask_value, const ScoreExtras& ex, cudaStream_t stream) {
    if (hd <= 8) return launch<T, 8, kExact>(qkv, table, pad_qkv, out, B, H, W, C, heads, hd, ws, shift, scale, mask_value, ex, stream);
    if (hd <= 16) return launch<T, 16, kExact>(qkv, table, pad_qkv, out, B, H, W, C, heads, hd, ws, shift, scale, mask_value, ex, stream);
    if (hd <= 32) return launch<T, 32, kExact>(qkv, table, pad_qkv, out, B, H, W, C, heads, hd, ws, shift, scale, mask_value, ex, stream);
    return launch<T, 64, kExact>(qkv, table, pad_qkv, out, B, H, W, C, heads, hd, ws, shift, scale, mask_value, ex, stream);
}

}  // namespace

int window_attn_generic_ex(const void* qkv, const float* table, const void* pad_qkv, void* out, int B, int H, int W,
                           int C, int heads, int ws, int shift, int dtype, float scale, float mask_value,
                           const float* dense_mask, int mask_windows, const float* head_scale, int normalize_qk,
                           const float* rel_pos_h, const float* rel_pos_w, cudaStream_t stream) {
    const int hd = C / heads;
    if (hd > 64 || ws > 1023) return SODT_ERR_UNSUPPORTED;
    if (dense_mask != nullptr && mask_windows <= 0) return SODT_ERR_INVALID_ARG;
    if ((rel_pos_h == nullptr) != (rel_pos_w == nullptr) || (rel_pos_h != nullptr && normalize_qk)) return SODT_ERR_INVALID_ARG;
    const ScoreExtras ex{dense_mask, dense_mask ? mask_windows : 1, head_scale, normalize_qk ? 1 : 0, rel_pos_h, rel_pos_w};
    if (dtype == SODT_F32)
        return dispatch_hd<float, true>(qkv, table, pad_qkv, out, B, H, W, C, heads, hd, ws, shift, scale, mask_value, ex, stream);
    return dispatch_hd<__nv_bfloat16, false>(qkv, table, pad_qkv, out, B, H, W, C, heads, hd, ws, shift, scale, mask_value, ex, stream);
}

int window_attn_generic(const void* qkv, const float* table, const void* pad_qkv, void* out, int B, int H, int W,
                        int C, int heads, int ws, int shift, int dtype, float scale, float mask_value,
                        cudaStream_t stream) {
    return window_attn_generic_ex(qkv, table, pad_qkv, out, B, H, W, C, heads, ws, shift, dtype, scale, mask_value, nullptr, 0, nullptr, 0,
                                  nullptr, nullptr, stream);
}

}  // namespace sodt
