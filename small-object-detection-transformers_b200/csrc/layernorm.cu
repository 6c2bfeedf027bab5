// Fused (residual add +) LayerNorm over the channel dimension of token rows, HBM-bound.
//
// Reference: the norm1 / norm2 LayerNorms and the residual adds of SwinTransformerBlock.forward
// (basics/models/backbone_vit.py:1089-1090,1125,1128) and PatchMerging.norm (:858); eps 1e-5, biased variance.
//
//   s       = a (+ r)                     fp32
//   y       = (s - mean(s)) * rsqrt(var(s) + eps) * w + b
//   sum_out = s (+ extra_bias)            optional second output: the residual stream with the bias of the NEXT
//                                         projection pre-added, so that the next GEMM can add the residual in its
//                                         epilogue (D = A W^T + C) and no separate elementwise add pass remains
//
// One warp per row, two rows in flight per warp, statistics by warp shuffles in fp32 (two-pass: mean, then centred
// variance), 4-byte (bf16x2) / 8-byte (float2) coalesced accesses.  Algorithmic bytes per row: C * (1 [+1] reads +
// 1 [+1] writes) * sizeof(T).
#include "common.cuh"

namespace sodt {
namespace {

constexpr int WARPS = 8;
constexpr int ROWS_PER_WARP = 2;

template <typename T> struct Pair;
template <> struct Pair<float> {
    using V = float2;
    static __device__ __forceinline__ void load(const float* p, float& a, float& b) { float2 v = *reinterpret_cast<const float2*>(p); a = v.x; b = v.y; }
    static __device__ __forceinline__ void store(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
};
template <> struct Pair<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float& a, float& b) {
        __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(p);
        a = __low2float(v); b = __high2float(v);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, float a, float b) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b); }
};

// NP = pairs per lane (C <= 64 * NP)
template <typename T, int NP>
__global__ void __launch_bounds__(WARPS * 32)
add_layernorm_kernel(const T* __restrict__ a, const T* __restrict__ r, const float* __restrict__ w, const float* __restrict__ b,
                     const float* __restrict__ extra_bias, T* __restrict__ sum_out, T* __restrict__ y,
                     long long rows, int C, float eps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npairs = C >> 1;
    float wv[NP][2], bv[NP][2], ev[NP][2];   // parameters stay in registers across the warp's rows
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const int pi = lane + 32 * k;
        const bool ok = pi < npairs;
        wv[k][0] = ok ? w[2 * pi] : 0.f; wv[k][1] = ok ? w[2 * pi + 1] : 0.f;
        bv[k][0] = ok ? b[2 * pi] : 0.f; bv[k][1] = ok ? b[2 * pi + 1] : 0.f;
        ev[k][0] = (ok && extra_bias) ? extra_bias[2 * pi] : 0.f; ev[k][1] = (ok && extra_bias) ? extra_bias[2 * pi + 1] : 0.f;
    }
    const float invC = 1.f / (float)C;
    const long long stride = (long long)gridDim.x * WARPS * ROWS_PER_WARP;
    for (long long row0 = ((long long)blockIdx.x * WARPS + warp) * ROWS_PER_WARP; row0 < rows; row0 += stride) {
        float v[ROWS_PER_WARP][NP][2];
#pragma unroll
        for (int q = 0; q < ROWS_PER_WARP; ++q) {
            const long long row = row0 + q;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const int pi = lane + 32 * k;
                v[q][k][0] = v[q][k][1] = 0.f;
                if (row < rows && pi < npairs) {
                    Pair<T>::load(a + row * C + 2 * pi, v[q][k][0], v[q][k][1]);
                    if (r != nullptr) {
                        float r0, r1;
                        Pair<T>::load(r + row * C + 2 * pi, r0, r1);
                        v[q][k][0] += r0; v[q][k][1] += r1;
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < ROWS_PER_WARP; ++q) {
            const long long row = row0 + q;
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < NP; ++k) s += v[q][k][0] + v[q][k][1];
            const float mean = warp_sum(s) * invC;
            float ss = 0.f;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const int pi = lane + 32 * k;
                if (pi < npairs) {
                    const float d0 = v[q][k][0] - mean, d1 = v[q][k][1] - mean;
                    ss = fmaf(d0, d0, ss); ss = fmaf(d1, d1, ss);
                }
            }
            const float rstd = rsqrtf(warp_sum(ss) * invC + eps);
            if (row < rows) {
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    const int pi = lane + 32 * k;
                    if (pi < npairs) {
                        Pair<T>::store(y + row * C + 2 * pi, (v[q][k][0] - mean) * rstd * wv[k][0] + bv[k][0],
                                       (v[q][k][1] - mean) * rstd * wv[k][1] + bv[k][1]);
                        if (sum_out != nullptr)
                            Pair<T>::store(sum_out + row * C + 2 * pi, v[q][k][0] + ev[k][0], v[q][k][1] + ev[k][1]);
                    }
                }
            }
        }
    }
}

// ---- bf16 fast path: C = 24 * THR (THR = 8, 16, 32 lanes per row), three 16-byte vectors per lane and row --------
// 32/THR rows per warp, parameters in shared memory, 1.5 KB of loads in flight per warp and operand.
template <int THR>
__global__ void __launch_bounds__(256)
add_layernorm_bf16_vec_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ r,
                              const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ extra_bias,
                              __nv_bfloat16* __restrict__ sum_out, __nv_bfloat16* __restrict__ y, long long rows, float eps) {
    constexpr int C = 24 * THR;
    constexpr int RPW = 32 / THR;                       // rows per warp
    __shared__ __align__(16) float sw[C], sb[C], se[C];
    pdl_trigger();
    pdl_wait();                                         // programmatic dependent launch (common.cuh)
    for (int e = threadIdx.x; e < C; e += blockDim.x) { sw[e] = w[e]; sb[e] = b[e]; se[e] = extra_bias ? extra_bias[e] : 0.f; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % THR, rsub = lane / THR;
    const long long stride = (long long)gridDim.x * 8 * RPW;
    constexpr float invC = 1.f / C;
    for (long long row = ((long long)blockIdx.x * 8 + warp) * RPW + rsub; row < rows; row += stride) {
        float v[24];
        const uint4* pa = reinterpret_cast<const uint4*>(a + row * C);
        uint4 raw[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) raw[k] = pa[sub + THR * k];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[k]);
#pragma unroll
            for (int e = 0; e < 4; ++e) { v[8 * k + 2 * e] = __low2float(h[e]); v[8 * k + 2 * e + 1] = __high2float(h[e]); }
        }
        if (r != nullptr) {
            const uint4* pr = reinterpret_cast<const uint4*>(r + row * C);
#pragma unroll
            for (int k = 0; k < 3; ++k) raw[k] = pr[sub + THR * k];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[k]);
#pragma unroll
                for (int e = 0; e < 4; ++e) { v[8 * k + 2 * e] += __low2float(h[e]); v[8 * k + 2 * e + 1] += __high2float(h[e]); }
            }
        }
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 24; ++e) s += v[e];
#pragma unroll
        for (int o = THR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * invC;
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < 24; ++e) { const float d = v[e] - mean; ss = fmaf(d, d, ss); }
#pragma unroll
        for (int o = THR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float rstd = rsqrtf(ss * invC + eps);
        uint4* py = reinterpret_cast<uint4*>(y + row * C);
        uint4* ps = sum_out ? reinterpret_cast<uint4*>(sum_out + row * C) : nullptr;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int c0 = (sub + THR * k) * 8;
            const float4 w0 = *reinterpret_cast<const float4*>(sw + c0), w1 = *reinterpret_cast<const float4*>(sw + c0 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(sb + c0), b1 = *reinterpret_cast<const float4*>(sb + c0 + 4);
            const float* x = v + 8 * k;
            uint4 o;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
            h[0] = __floats2bfloat162_rn((x[0] - mean) * rstd * w0.x + b0.x, (x[1] - mean) * rstd * w0.y + b0.y);
            h[1] = __floats2bfloat162_rn((x[2] - mean) * rstd * w0.z + b0.z, (x[3] - mean) * rstd * w0.w + b0.w);
            h[2] = __floats2bfloat162_rn((x[4] - mean) * rstd * w1.x + b1.x, (x[5] - mean) * rstd * w1.y + b1.y);
            h[3] = __floats2bfloat162_rn((x[6] - mean) * rstd * w1.z + b1.z, (x[7] - mean) * rstd * w1.w + b1.w);
            py[sub + THR * k] = o;
            if (ps != nullptr) {
                const float4 e0 = *reinterpret_cast<const float4*>(se + c0), e1 = *reinterpret_cast<const float4*>(se + c0 + 4);
                h[0] = __floats2bfloat162_rn(x[0] + e0.x, x[1] + e0.y);
                h[1] = __floats2bfloat162_rn(x[2] + e0.z, x[3] + e0.w);
                h[2] = __floats2bfloat162_rn(x[4] + e1.x, x[5] + e1.y);
                h[3] = __floats2bfloat162_rn(x[6] + e1.z, x[7] + e1.w);
                ps[sub + THR * k] = o;
            }
        }
    }
}

template <int THR>
int launch_vec(const void* a, const void* r, const float* w, const float* b, const float* extra_bias, void* sum_out, void* y,
               long long rows, float eps, cudaStream_t stream) {
    constexpr int RPW = 32 / THR;
    long long blocks = (rows + 8 * RPW - 1) / (8 * RPW);
    const long long resident = 148LL * 8;
    if (blocks > resident) blocks = resident;
    const cudaError_t e = launch_pdl(add_layernorm_bf16_vec_kernel<THR>, dim3((unsigned)blocks), dim3(256), 0, stream, true,
        static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(r), w, b, extra_bias,
        static_cast<__nv_bfloat16*>(sum_out), static_cast<__nv_bfloat16*>(y), rows, eps);
    if (e != cudaSuccess) return cuda_status(e);
    return check_launch();
}

template <typename T, int NP>
int launch(const void* a, const void* r, const float* w, const float* b, const float* extra_bias, void* sum_out, void* y,
           long long rows, int C, float eps, cudaStream_t stream) {
    const long long per_block = WARPS * ROWS_PER_WARP;
    long long blocks = (rows + per_block - 1) / per_block;
    const long long resident = 148LL * 8;     // persistent-style grid: parameters are loaded once per warp
    if (blocks > resident) blocks = resident;
    add_layernorm_kernel<T, NP><<<(unsigned)blocks, WARPS * 32, 0, stream>>>(
        static_cast<const T*>(a), static_cast<const T*>(r), w, b, extra_bias, static_cast<T*>(sum_out), static_cast<T*>(y), rows, C, eps);
    return check_launch();
}

template <typename T>
int dispatch(const void* a, const void* r, const float* w, const float* b, const float* extra_bias, void* sum_out, void* y,
             long long rows, int C, float eps, cudaStream_t stream) {
    if (C <= 64) return launch<T, 1>(a, r, w, b, extra_bias, sum_out, y, rows, C, eps, stream);
    if (C <= 192) return launch<T, 3>(a, r, w, b, extra_bias, sum_out, y, rows, C, eps, stream);
    if (C <= 384) return launch<T, 6>(a, r, w, b, extra_bias, sum_out, y, rows, C, eps, stream);
    if (C <= 768) return launch<T, 12>(a, r, w, b, extra_bias, sum_out, y, rows, C, eps, stream);
    return launch<T, 16>(a, r, w, b, extra_bias, sum_out, y, rows, C, eps, stream);
}

// (mean, rstd) of every bf16 row: the statistics a LayerNorm folded into the next GEMM needs (linear_tc.cu).
// 8 lanes per row, 16-byte loads; one read pass, 8 bytes written per row.
__global__ void __launch_bounds__(256)
row_stats_bf16_kernel(const __nv_bfloat16* __restrict__ x, float2* __restrict__ stats, long long rows, int C, long long ld, float eps) {
    const int sub = threadIdx.x & 7;
    const int chunks = C >> 3;
    const float inv = 1.f / (float)C;
    pdl_trigger();
    pdl_wait();
    for (long long row = (long long)blockIdx.x * 32 + (threadIdx.x >> 3); row < rows; row += (long long)gridDim.x * 32) {
        const uint4* src = reinterpret_cast<const uint4*>(x + row * ld);
        float s1 = 0.f, s2 = 0.f;
        for (int c = sub; c < chunks; c += 8) {
            const uint4 q = __ldg(src + c);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xFFFF0000u);
                s1 += lo + hi;
                s2 = fmaf(lo, lo, fmaf(hi, hi, s2));
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const float mean = s1 * inv;
        if (sub == 0) stats[row] = make_float2(mean, rsqrtf(fmaxf(fmaf(-mean, mean, s2 * inv), 0.f) + eps));
    }
}

// Partial (sum, sum of squares) per 64-column box, [boxes][rows][2] as emitted by the GEMM epilogue, -> (mean, rstd) per row.
__global__ void __launch_bounds__(256)
stats_finalize_kernel(const float2* __restrict__ part, float2* __restrict__ stats, long long rows, int boxes, float inv, float eps) {
    pdl_trigger();
    pdl_wait();
    for (long long row = (long long)blockIdx.x * 256 + threadIdx.x; row < rows; row += (long long)gridDim.x * 256) {
        float s1 = 0.f, s2 = 0.f;
        for (int b = 0; b < boxes; b += 4) {
            float2 p[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = b + i < boxes ? __ldg(part + (size_t)(b + i) * rows + row) : make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) { s1 += p[i].x; s2 += p[i].y; }
        }
        const float mean = s1 * inv;
        stats[row] = make_float2(mean, rsqrtf(fmaxf(fmaf(-mean, mean, s2 * inv), 0.f) + eps));
    }
}

}  // namespace
}  // namespace sodt

extern "C" int sodt_row_stats(const void* x, long long ld, float* mean_rstd, long long rows, int C, float eps, int dtype, void* stream) {
    using namespace sodt;
    if (!x || !mean_rstd || rows <= 0 || C <= 0 || ld < C || eps < 0.f) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_BF16 || C % 8 || ld % 8) return SODT_ERR_UNSUPPORTED;
    if (!aligned16(x) || (reinterpret_cast<uintptr_t>(mean_rstd) & 7)) return SODT_ERR_ALIGNMENT;
    long long blocks = (rows + 31) / 32;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const cudaError_t e = launch_pdl(row_stats_bf16_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), true,
        static_cast<const __nv_bfloat16*>(x), reinterpret_cast<float2*>(mean_rstd), rows, C, (long long)ld, eps);
    if (e != cudaSuccess) return cuda_status(e);
    return check_launch();
}

extern "C" int sodt_stats_finalize(const float* partials, int boxes, float* mean_rstd, long long rows, int C, float eps, void* stream) {
    using namespace sodt;
    if (!partials || !mean_rstd || rows <= 0 || boxes <= 0 || C <= 0 || eps < 0.f) return SODT_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(partials) & 7) || (reinterpret_cast<uintptr_t>(mean_rstd) & 7)) return SODT_ERR_ALIGNMENT;
    long long blocks = (rows + 255) / 256;             // one row per thread: latency-bound with fewer threads
    if (blocks > 148 * 64) blocks = 148 * 64;
    const cudaError_t e = launch_pdl(stats_finalize_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), true,
        reinterpret_cast<const float2*>(partials), reinterpret_cast<float2*>(mean_rstd), rows, boxes, 1.f / (float)C, eps);
    if (e != cudaSuccess) return cuda_status(e);
    return check_launch();
}

extern "C" int sodt_add_layernorm_fwd(const void* a, const void* r, const float* w, const float* b, const float* extra_bias,
                                      void* sum_out, void* y, long long rows, int C, float eps, int dtype, void* stream) {
    using namespace sodt;
    if (!a || !w || !b || !y || rows <= 0 || C <= 0) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_F32 && dtype != SODT_BF16) return SODT_ERR_INVALID_ARG;
    if (C % 2 || C > 1024) return SODT_ERR_UNSUPPORTED;
    const uintptr_t al = dtype == SODT_F32 ? 7 : 3;
    if ((reinterpret_cast<uintptr_t>(a) & al) || (reinterpret_cast<uintptr_t>(y) & al) || (r && (reinterpret_cast<uintptr_t>(r) & al)) ||
        (sum_out && (reinterpret_cast<uintptr_t>(sum_out) & al)))
        return SODT_ERR_ALIGNMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == SODT_F32) return dispatch<float>(a, r, w, b, extra_bias, sum_out, y, rows, C, eps, s);
    const bool al16 = aligned16(a) && aligned16(y) && (!r || aligned16(r)) && (!sum_out || aligned16(sum_out));
    if (al16 && C == 192) return launch_vec<8>(a, r, w, b, extra_bias, sum_out, y, rows, eps, s);
    if (al16 && C == 384) return launch_vec<16>(a, r, w, b, extra_bias, sum_out, y, rows, eps, s);
    if (al16 && C == 768) return launch_vec<32>(a, r, w, b, extra_bias, sum_out, y, rows, eps, s);
    return dispatch<__nv_bfloat16>(a, r, w, b, extra_bias, sum_out, y, rows, C, eps, s);
}
