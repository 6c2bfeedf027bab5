// Fused front end of ImageEncoderViT: the four single-channel patch embeddings
// (Conv2d(1, E, 4, stride 4); the R stream with PatchEmbed's default padding (1,1), G/B/IR with padding 0 --
// reference basics/models/backbone_vit.py:69-98,751) followed by the window-1 cross-channel block
// (CAttentionBlock, :469-561: with one token per window attention returns v, so x_i = LayerNorm_i(e_i + e_partner) for
// the pairs R<-G, G<-B, B<-IR, IR<-G) and the channel concatenation of :210.
//
// Work split: 8 lanes per token (lane q owns output channels 6q..6q+5 of every stream), 4 tokens per thread, 64 tokens
// per 128-thread block.  The block first stages the 4 x 16 pixels of its tokens in shared memory (fp32); conv weights sit
// in shared memory transposed to [stream][tap][channel] so that the 8 lanes of a token read 8 distinct banks and the four
// token groups of a warp broadcast.  The add + LayerNorm statistics are reduced over the 8 lanes with shuffles.  The four
// conv outputs, the four bias-add passes and the cross-channel block's inputs never reach HBM.
#include "common.cuh"

namespace sodt {
namespace {

constexpr int KS = 4;            // kernel size = stride
constexpr int E = 48;            // embedding channels per stream
constexpr int CPL = 6;           // channels per lane (8 lanes x 6 = 48)
constexpr int TPT = 4;           // tokens per thread: every weight read from shared memory feeds TPT x 6 FMAs (the kernel is LDS-bound)
constexpr int TOK_PER_BLOCK = 16 * TPT;

template <typename T>
__device__ __forceinline__ void store6(T* dst, const float (&y)[CPL]) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int j = 0; j < CPL; j += 2) *reinterpret_cast<float2*>(dst + j) = make_float2(y[j], y[j + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < CPL; j += 2) *reinterpret_cast<__nv_bfloat162*>(dst + j) = __floats2bfloat162_rn(y[j], y[j + 1]);
    }
}

// pixel -> fp32.  uint8 images are scaled like the reference's evaluation loop (img / 255, basics/test.py:124-130): the
// product u * (1/255) is rounded to the activation type first, exactly what the separate conversion pass produced.
template <typename T, typename TI> __device__ __forceinline__ float load_px(const TI* p) { return to_f32<TI>(*p); }
template <> __device__ __forceinline__ float load_px<float, uint8_t>(const uint8_t* p) { return __fmul_rn((float)*p, 1.f / 255.f); }
template <> __device__ __forceinline__ float load_px<__nv_bfloat16, uint8_t>(const uint8_t* p) {
    return __bfloat162float(__float2bfloat16_rn(__fmul_rn((float)*p, 1.f / 255.f)));
}

template <typename T, typename TI>
__global__ void __launch_bounds__(128)
frontend_kernel(const TI* __restrict__ x, long long sb, long long sc, long long sy, long long sx,
                const TI* __restrict__ x_ir, long long ib, long long iy, long long ix,   // optional separate source of stream 3
                const float* __restrict__ conv_w /*[4][E][16]*/, const float* __restrict__ conv_b /*[4][E]*/,
                const float* __restrict__ ln_w /*[4][E]*/, const float* __restrict__ ln_b, T* __restrict__ out,
                int H, int W, int h, int w, long long ntok, int pad0, float eps) {
    __shared__ __align__(16) float s_wt[4][16][E];                   // transposed conv weights
    __shared__ float s_cb[4 * E], s_lw[4 * E], s_lb[4 * E];
    __shared__ __align__(16) float s_px[TOK_PER_BLOCK][4][16];        // pixel patches of the block's tokens
    const int tid = threadIdx.x;
    for (int e = tid; e < 4 * E * 16; e += blockDim.x) {
        const int s = e / (E * 16), c = (e / 16) % E, k = e % 16;
        s_wt[s][k][c] = conv_w[e];
    }
    for (int e = tid; e < 4 * E; e += blockDim.x) { s_cb[e] = conv_b[e]; s_lw[e] = ln_w[e]; s_lb[e] = ln_b[e]; }
    const int lane = tid & 31, warp = tid >> 5;
    const int q = lane & 7, tg = lane >> 3;
    const int tl0 = (warp * 4 + tg) * TPT;                            // first local token of this thread
    const int c0 = q * CPL;
    const long long ngroups = (ntok + TOK_PER_BLOCK - 1) / TOK_PER_BLOCK;
    // weights are staged once per block; the block then loops over groups of 32 tokens
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const long long tok0 = grp * TOK_PER_BLOCK;
    __syncthreads();                                                  // previous group's pixels are no longer read
    for (int half = 0; half < TOK_PER_BLOCK / 32; ++half) {   // stage: thread -> (token tid/4 + 32*half, channel tid%4)
        const int tl = (tid >> 2) + 32 * half, s = tid & 3;
        const long long tok = tok0 + tl;
        float px[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) px[k] = 0.f;
        if (tok < ntok) {
            const int tx = (int)(tok % w);
            const long long rest = tok / w;
            const int ty = (int)(rest % h);
            const long long b = rest / h;
            const int off = s == 0 ? pad0 : 0;                        // only the R stream is padded
            const bool sep = s == 3 && x_ir != nullptr;
            const TI* img = sep ? x_ir + b * ib : x + b * sb + s * sc;
            const long long psy = sep ? iy : sy, psx = sep ? ix : sx;
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
                const int yy = ty * KS + ky - off;
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    const int xx = tx * KS + kx - off;
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W) px[ky * KS + kx] = load_px<T, TI>(img + yy * psy + xx * psx);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 16; k += 4) *reinterpret_cast<float4*>(&s_px[tl][s][k]) = make_float4(px[k], px[k + 1], px[k + 2], px[k + 3]);
    }
    __syncthreads();

    auto embed = [&](int s, float (&e)[TPT][CPL]) {
#pragma unroll
        for (int t = 0; t < TPT; ++t)
#pragma unroll
            for (int j = 0; j < CPL; ++j) e[t][j] = s_cb[s * E + c0 + j];
#pragma unroll
        for (int k4 = 0; k4 < 16; k4 += 4) {
            float4 p[TPT];
#pragma unroll
            for (int t = 0; t < TPT; ++t) p[t] = *reinterpret_cast<const float4*>(&s_px[tl0 + t][s][k4]);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float wv[CPL];
#pragma unroll
                for (int j = 0; j < CPL; j += 2) {
                    const float2 w2 = *reinterpret_cast<const float2*>(&s_wt[s][k4 + kk][c0 + j]);
                    wv[j] = w2.x; wv[j + 1] = w2.y;
                }
#pragma unroll
                for (int t = 0; t < TPT; ++t) {
                    const float pv = kk == 0 ? p[t].x : kk == 1 ? p[t].y : kk == 2 ? p[t].z : p[t].w;
#pragma unroll
                    for (int j = 0; j < CPL; ++j) e[t][j] = fmaf(wv[j], pv, e[t][j]);
                }
            }
        }
    };
    auto pair_out = [&](int pidx, const float (&a)[TPT][CPL], const float (&k)[TPT][CPL]) {
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            float v[CPL], sum = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) { v[j] = a[t][j] + k[t][j]; sum += v[j]; }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum * (1.f / E);
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) { const float d = v[j] - mean; ss = fmaf(d, d, ss); }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float rstd = rsqrtf(ss * (1.f / E) + eps);
            const long long tok = tok0 + tl0 + t;
            if (tok < ntok) {
                float y[CPL];
#pragma unroll
                for (int j = 0; j < CPL; ++j) y[j] = (v[j] - mean) * rstd * s_lw[pidx * E + c0 + j] + s_lb[pidx * E + c0 + j];
                store6<T>(out + tok * (4LL * E) + pidx * E + c0, y);
            }
        }
    };
    float eg[TPT][CPL], ea[TPT][CPL], eb[TPT][CPL];
    embed(0, ea);                 // R
    embed(1, eg);                 // G
    pair_out(0, ea, eg);          // R <- G
    embed(2, eb);                 // B
    pair_out(1, eg, eb);          // G <- B
    embed(3, ea);                 // IR
    pair_out(2, eb, ea);          // B <- IR
    pair_out(3, ea, eg);          // IR <- G
    }
}

}  // namespace
}  // namespace sodt

namespace sodt {
namespace {
template <typename T, typename TI>
int frontend_launch(const void* x, long long sb, long long sc, long long sy, long long sx, const void* x_ir, long long ib, long long iy,
                    long long ix, const float* conv_w, const float* conv_b, const float* ln_w, const float* ln_b, void* out,
                    int B, int H, int W, int pad_r, float eps, cudaStream_t s) {
    // conv output size with padding p: floor((H + 2p - 4) / 4) + 1; the reference needs all four streams to agree
    const int h = (H - 4) / 4 + 1, w = (W - 4) / 4 + 1;
    if ((H + 2 * pad_r - 4) / 4 + 1 != h || (W + 2 * pad_r - 4) / 4 + 1 != w) return SODT_ERR_UNSUPPORTED;
    const long long ntok = (long long)B * h * w;
    long long blocks = (ntok + TOK_PER_BLOCK - 1) / TOK_PER_BLOCK;
    if (blocks > 148LL * 12) blocks = 148LL * 12;
    frontend_kernel<T, TI><<<(unsigned)blocks, 128, 0, s>>>(static_cast<const TI*>(x), sb, sc, sy, sx, static_cast<const TI*>(x_ir), ib, iy, ix,
                                                            conv_w, conv_b, ln_w, ln_b, static_cast<T*>(out), H, W, h, w, ntok, pad_r, eps);
    return check_launch();
}
}  // namespace
}  // namespace sodt

extern "C" int sodt_frontend_fwd(const void* x, long long sb, long long sc, long long sy, long long sx,
                                 const float* conv_w, const float* conv_b, const float* ln_w, const float* ln_b, void* out,
                                 int B, int H, int W, int E, int pad_r, float eps, int dtype, void* stream) {
    using namespace sodt;
    if (!x || !conv_w || !conv_b || !ln_w || !ln_b || !out || B <= 0 || H < 4 || W < 4) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_F32 && dtype != SODT_BF16) return SODT_ERR_INVALID_ARG;
    if (E != 48 || (pad_r != 0 && pad_r != 1)) return SODT_ERR_UNSUPPORTED;
    if (!aligned16(out)) return SODT_ERR_ALIGNMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == SODT_F32)
        return frontend_launch<float, float>(x, sb, sc, sy, sx, nullptr, 0, 0, 0, conv_w, conv_b, ln_w, ln_b, out, B, H, W, pad_r, eps, s);
    return frontend_launch<__nv_bfloat16, __nv_bfloat16>(x, sb, sc, sy, sx, nullptr, 0, 0, 0, conv_w, conv_b, ln_w, ln_b, out, B, H, W, pad_r, eps, s);
}

extern "C" int sodt_frontend_u8_fwd(const void* rgb, long long rb, long long rc, long long ry, long long rx,
                                    const void* ir, long long ib, long long iy, long long ix,
                                    const float* conv_w, const float* conv_b, const float* ln_w, const float* ln_b, void* out,
                                    int B, int H, int W, int E, int pad_r, float eps, int dtype, void* stream) {
    using namespace sodt;
    if (!rgb || !ir || !conv_w || !conv_b || !ln_w || !ln_b || !out || B <= 0 || H < 4 || W < 4) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_F32 && dtype != SODT_BF16) return SODT_ERR_INVALID_ARG;
    if (E != 48 || (pad_r != 0 && pad_r != 1)) return SODT_ERR_UNSUPPORTED;
    if (!aligned16(out)) return SODT_ERR_ALIGNMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == SODT_F32)
        return frontend_launch<float, uint8_t>(rgb, rb, rc, ry, rx, ir, ib, iy, ix, conv_w, conv_b, ln_w, ln_b, out, B, H, W, pad_r, eps, s);
    return frontend_launch<__nv_bfloat16, uint8_t>(rgb, rb, rc, ry, rx, ir, ib, iy, ix, conv_w, conv_b, ln_w, ln_b, out, B, H, W, pad_r, eps, s);
}
