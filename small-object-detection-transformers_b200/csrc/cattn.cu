// Cross-channel attention block: R<-G, G<-B, B<-IR, IR<-G parameter-free multi-head cross
// attention, each followed by LayerNorm(stream + attn).  HBM-bound: 4*C elements in, 4*C out per token.
//
// Reference: CAttentionBlock.forward basics/models/backbone_vit.py:469-561 (window 1, shipped),
// backbone_swinv2.py:429-469 (general window); CAttention.forward backbone_vit.py:589-616.
//
//  * ws == 1 (the shipped model): one token per window, softmax over a single score == 1, attention
//    returns v bit-exactly, so the block is four fused add + LayerNorm passes.  One thread per token,
//    every stream read once (G twice), 16-byte vector I/O.
//  * ws > 1: one CTA per group of windows; streams staged in shared memory as fp32, scores of one
//    (pair, head) at a time in shared memory, fp32 softmax, LayerNorm by one warp per token.
#include "common.cuh"

namespace sodt {

namespace {

__constant__ int kPairQ[4] = {0, 1, 2, 3};
__constant__ int kPairK[4] = {1, 2, 3, 1};

struct Streams {
    const void* p[4];
};

// ------------------------------------------------------------------------------ ws == 1
template <typename T, int C>
__device__ __forceinline__ void load_token(const T* __restrict__ base, long long sc, bool vec, float dst[C]) {
    if (vec) {
        constexpr int PER = 16 / sizeof(T);
#pragma unroll
        for (int c = 0; c < C; c += PER) {
            uint4 raw = *reinterpret_cast<const uint4*>(base + c);
            if constexpr (sizeof(T) == 4) {
                dst[c] = __uint_as_float(raw.x); dst[c + 1] = __uint_as_float(raw.y);
                dst[c + 2] = __uint_as_float(raw.z); dst[c + 3] = __uint_as_float(raw.w);
            } else {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int e = 0; e < 4; ++e) { dst[c + 2 * e] = __low2float(h[e]); dst[c + 2 * e + 1] = __high2float(h[e]); }
            }
        }
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) dst[c] = to_f32<T>(base[c * sc]);
    }
}

template <typename T, int C>
__device__ __forceinline__ void add_ln_store(const float a[C], const float k[C], const float* __restrict__ w,
                                             const float* __restrict__ bia, float eps, T* __restrict__ dst) {
    float x[C];
    float mean = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { x[c] = a[c] + k[c]; mean += x[c]; }
    mean *= (1.f / C);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { float d = x[c] - mean; var = fmaf(d, d, var); }
    const float rstd = rsqrtf(var * (1.f / C) + eps);
    constexpr int PER = 16 / sizeof(T);
#pragma unroll
    for (int c = 0; c < C; c += PER) {
        float y[PER];
#pragma unroll
        for (int e = 0; e < PER; ++e) y[e] = (x[c + e] - mean) * rstd * w[c + e] + bia[c + e];
        uint4 raw;
        if constexpr (sizeof(T) == 4) {
            raw = make_uint4(__float_as_uint(y[0]), __float_as_uint(y[1]), __float_as_uint(y[2]), __float_as_uint(y[3]));
        } else {
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(y[2 * e], y[2 * e + 1]);
        }
        *reinterpret_cast<uint4*>(dst + c) = raw;
    }
}

template <typename T, int C>
__global__ void __launch_bounds__(128)
cattn_n1_kernel(Streams st, long long sb, long long sy, long long sx, long long sc, int vec,
                const float* __restrict__ ln_w, const float* __restrict__ ln_b, T* __restrict__ out,
                long long ntok, int h, int w, float eps) {
    __shared__ float s_w[4 * C], s_b[4 * C];
    for (int e = threadIdx.x; e < 4 * C; e += blockDim.x) { s_w[e] = ln_w[e]; s_b[e] = ln_b[e]; }
    __syncthreads();
    const long long tok = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tok >= ntok) return;
    const int x = (int)(tok % w);
    const long long rest = tok / w;
    const int y = (int)(rest % h);
    const long long b = rest / h;
    const long long off = b * sb + y * sy + x * sx;
    T* dst = out + tok * (4LL * C);
    float a[C], k[C];
    load_token<T, C>(static_cast<const T*>(st.p[0]) + off, sc, vec, a);   // R
    load_token<T, C>(static_cast<const T*>(st.p[1]) + off, sc, vec, k);   // G
    add_ln_store<T, C>(a, k, s_w, s_b, eps, dst);                          // R <- G
    load_token<T, C>(static_cast<const T*>(st.p[2]) + off, sc, vec, a);   // B
    add_ln_store<T, C>(k, a, s_w + C, s_b + C, eps, dst + C);              // G <- B
    load_token<T, C>(static_cast<const T*>(st.p[3]) + off, sc, vec, k);   // IR
    add_ln_store<T, C>(a, k, s_w + 2 * C, s_b + 2 * C, eps, dst + 2 * C);  // B <- IR
    load_token<T, C>(static_cast<const T*>(st.p[1]) + off, sc, vec, a);   // G again (L1/L2 hit)
    add_ln_store<T, C>(k, a, s_w + 3 * C, s_b + 3 * C, eps, dst + 3 * C);  // IR <- G
}

// ------------------------------------------------------------------------------ general ws
template <typename T>
__global__ void __launch_bounds__(256)
cattn_general_kernel(Streams st, long long sb, long long sy, long long sx, long long sc,
                     const float* __restrict__ ln_w, const float* __restrict__ ln_b, T* __restrict__ out,
                     int B, int h, int w, int C, int heads, int ws, int shift, int G, long long total_windows,
                     float eps, float mask_value) {
    extern __shared__ __align__(16) float smem[];
    const WinGeom g(h, w, ws, shift);
    const int N = ws * ws;
    const int nW = g.nwh * g.nww;
    const int LDC = C + 1, LDS = N + 1;
    const int GN = G * N;
    float* xs = smem;                         // [4][G*N][LDC]
    float* os = xs + 4 * GN * LDC;            // [G*N][LDC]
    float* sc_ = os + GN * LDC;               // [G*N][LDS]
    const int nfloats = 5 * GN * LDC + GN * LDS;
    long long* dsto = reinterpret_cast<long long*>(smem + nfloats + (nfloats & 1));  // [G*N] output token index or -1
    int* reg = reinterpret_cast<int*>(dsto + GN);                                      // [G*N] mask region
    const int tid = threadIdx.x, nthr = blockDim.x;
    const long long win0 = (long long)blockIdx.x * G;
    const int c = C / heads;
    const float denom = sqrtf((float)c);

    for (int e = tid; e < GN; e += nthr) {
        const long long wg = win0 + e / N;
        const int t = e % N;
        long long dst = -1;
        int rr = 0;
        if (wg < total_windows) {
            const int b = (int)(wg / nW), win = (int)(wg % nW);
            int yr, xr;
            if (g.rolled(win, t, yr, xr)) {
                int ys, xsrc;
                g.source(yr, xr, ys, xsrc);
                dst = ((long long)b * h + ys) * w + xsrc;
                rr = shift > 0 ? g.region(yr, xr) : 0;
            }
        }
        dsto[e] = dst;
        reg[e] = rr;
    }
    __syncthreads();
    for (int e = tid; e < 4 * GN * C; e += nthr) {
        const int ch = e % C;
        const int tok = (e / C) % GN;
        const int s = e / (C * GN);
        const long long dst = dsto[tok];
        float v = 0.f;  // zero padding, backbone_vit.py:632-639
        if (dst >= 0) {
            const long long xx = dst % w, yy = (dst / w) % h, bb = dst / ((long long)w * h);
            v = to_f32<T>(static_cast<const T*>(st.p[s])[bb * sb + yy * sy + xx * sx + ch * sc]);
        }
        xs[(s * GN + tok) * LDC + ch] = v;
    }
    __syncthreads();

    for (int p = 0; p < 4; ++p) {
        const float* qx = xs + kPairQ[p] * GN * LDC;
        const float* kx = xs + kPairK[p] * GN * LDC;
        for (int hh = 0; hh < heads; ++hh) {
            const int c0 = hh * c;
            for (int e = tid; e < GN * N; e += nthr) {
                const int j = e % N, qi = e / N;  // qi = gw*N + i
                const int kj = (qi / N) * N + j;
                float s = 0.f;
                for (int d = 0; d < c; ++d) s = fmaf(qx[qi * LDC + c0 + d], kx[kj * LDC + c0 + d], s);
                if (shift > 0 && reg[qi] != reg[kj]) s += mask_value;   // mask BEFORE scaling (:601-608)
                sc_[qi * LDS + j] = s / denom;
            }
            __syncthreads();
            for (int row = tid >> 5; row < GN; row += nthr >> 5) {
                float* sr = sc_ + row * LDS;
                const int lane = tid & 31;
                float m = -INFINITY;
                for (int j = lane; j < N; j += 32) m = fmaxf(m, sr[j]);
                m = warp_max(m);
                float sum = 0.f;
                for (int j = lane; j < N; j += 32) { float pe = expf(sr[j] - m); sr[j] = pe; sum += pe; }
                sum = warp_sum(sum);
                const float inv = 1.f / sum;
                for (int j = lane; j < N; j += 32) sr[j] *= inv;
            }
            __syncthreads();
            for (int e = tid; e < GN * c; e += nthr) {
                const int d = e % c, qi = e / c;
                const int kbase = (qi / N) * N;
                float acc = 0.f;
                for (int j = 0; j < N; ++j) acc = fmaf(sc_[qi * LDS + j], kx[(kbase + j) * LDC + c0 + d], acc);
                os[qi * LDC + c0 + d] = acc;
            }
            __syncthreads();
        }
        // LayerNorm_p(stream_q + attn) by one warp per token
        for (int row = tid >> 5; row < GN; row += nthr >> 5) {
            const long long dst = dsto[row];
            if (dst < 0) continue;
            const int lane = tid & 31;
            float sum = 0.f;
            for (int ch = lane; ch < C; ch += 32) sum += qx[row * LDC + ch] + os[row * LDC + ch];
            const float mean = warp_sum(sum) / C;
            float var = 0.f;
            for (int ch = lane; ch < C; ch += 32) { float d = qx[row * LDC + ch] + os[row * LDC + ch] - mean; var = fmaf(d, d, var); }
            const float rstd = rsqrtf(warp_sum(var) / C + eps);
            T* o = out + dst * (4LL * C) + p * C;
            for (int ch = lane; ch < C; ch += 32) {
                float v = qx[row * LDC + ch] + os[row * LDC + ch];
                o[ch] = from_f32<T>((v - mean) * rstd * ln_w[p * C + ch] + ln_b[p * C + ch]);
            }
        }
        __syncthreads();
    }
}

template <typename T, int C>
int launch_n1(const Streams& st, long long sb, long long sy, long long sx, long long sc, const float* ln_w,
              const float* ln_b, void* out, int B, int h, int w, float eps, cudaStream_t stream) {
    const long long ntok = (long long)B * h * w;
    const long long es = sizeof(T);
    bool vec = sc == 1 && (sb * es) % 16 == 0 && (sy * es) % 16 == 0 && (sx * es) % 16 == 0;
    for (int i = 0; i < 4; ++i) vec = vec && aligned16(st.p[i]);
    const int threads = 128;
    const long long blocks = (ntok + threads - 1) / threads;
    if (blocks > 2147483647LL) return SODT_ERR_UNSUPPORTED;
    cattn_n1_kernel<T, C><<<(unsigned)blocks, threads, 0, stream>>>(st, sb, sy, sx, sc, vec ? 1 : 0, ln_w, ln_b,
                                                                     static_cast<T*>(out), ntok, h, w, eps);
    return check_launch();
}

template <typename T>
int launch_general(const Streams& st, long long sb, long long sy, long long sx, long long sc, const float* ln_w,
                   const float* ln_b, void* out, int B, int h, int w, int C, int heads, int ws, int shift,
                   float eps, float mask_value, cudaStream_t stream) {
    const WinGeom g(h, w, ws, shift);
    const int N = ws * ws;
    const long long total = (long long)B * g.nwh * g.nww;
    auto bytes = [&](int G) {
        const size_t GN = (size_t)G * N;
        size_t fl = 5 * GN * (C + 1) + GN * (N + 1);
        fl += fl & 1;
        return fl * sizeof(float) + GN * (sizeof(long long) + sizeof(int)) + 16;
    };
    int G = N >= 64 ? 1 : 64 / N;
    while (G > 1 && bytes(G) > 200 * 1024) G >>= 1;
    const size_t smem = bytes(G);
    if (smem > 227 * 1024) return SODT_ERR_UNSUPPORTED;
    auto kern = cattn_general_kernel<T>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    const long long blocks = (total + G - 1) / G;
    if (blocks > 2147483647LL) return SODT_ERR_UNSUPPORTED;
    kern<<<(unsigned)blocks, 256, smem, stream>>>(st, sb, sy, sx, sc, ln_w, ln_b, static_cast<T*>(out), B, h, w, C,
                                                  heads, ws, shift, G, total, eps, mask_value);
    return check_launch();
}

template <typename T>
int dispatch(const Streams& st, long long sb, long long sy, long long sx, long long sc, const float* ln_w,
             const float* ln_b, void* out, int B, int h, int w, int C, int heads, int ws, int shift, float eps,
             float mask_value, cudaStream_t stream) {
    if (ws == 1 && aligned16(out)) {
        // one token per window: softmax over one score is exactly 1, attention == v;
        // the (all-zero) shift mask and the roll cancel.
        if (C == 24) return launch_n1<T, 24>(st, sb, sy, sx, sc, ln_w, ln_b, out, B, h, w, eps, stream);
        if (C == 48) return launch_n1<T, 48>(st, sb, sy, sx, sc, ln_w, ln_b, out, B, h, w, eps, stream);
        if (C == 96) return launch_n1<T, 96>(st, sb, sy, sx, sc, ln_w, ln_b, out, B, h, w, eps, stream);
    }
    return launch_general<T>(st, sb, sy, sx, sc, ln_w, ln_b, out, B, h, w, C, heads, ws, shift, eps, mask_value, stream);
}

}  // namespace
}  // namespace sodt

extern "C" int sodt_cattn_block_fwd(const void* r, const void* g, const void* b, const void* ir,
                                    long long sb, long long sy, long long sx, long long sc,
                                    const float* ln_w, const float* ln_b, void* out,
                                    int B, int h, int w, int C, int heads, int ws, int shift,
                                    float eps, float mask_value, int dtype, void* stream) {
    using namespace sodt;
    if (!r || !g || !b || !ir || !ln_w || !ln_b || !out) return SODT_ERR_INVALID_ARG;
    if (B <= 0 || h <= 0 || w <= 0 || C <= 0 || heads <= 0 || ws <= 0 || C % heads != 0) return SODT_ERR_INVALID_ARG;
    if (shift < 0 || (shift > 0 && shift >= ws && ws > 1)) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_F32 && dtype != SODT_BF16) return SODT_ERR_INVALID_ARG;
    if (C > 128 || ws * ws > 144) return SODT_ERR_UNSUPPORTED;
    Streams st{{r, g, b, ir}};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == SODT_F32)
        return dispatch<float>(st, sb, sy, sx, sc, ln_w, ln_b, out, B, h, w, C, heads, ws, shift, eps, mask_value, s);
    return dispatch<__nv_bfloat16>(st, sb, sy, sx, sc, ln_w, ln_b, out, B, h, w, C, heads, ws, shift, eps, mask_value, s);
}
