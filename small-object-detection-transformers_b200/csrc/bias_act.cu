// Fused bias + activation (+ spatial crop) on channels-last activations, one HBM pass.
//
//   out[b, y, x, c] = act(in[b, y + oy, x + ox, c] + bias[c])       act = identity | exact (erf) GELU | SiLU
//
// Used after the 2x2 conv of the conv-enhanced MLP (reference basics/models/backbone_vit.py:896-902: F.pad(0,1,0,1) ->
// conv2d -> GELU): the conv runs with symmetric padding 1 and no bias, and this kernel crops row / column 0, adds the bias
// and applies GELU, replacing the pad copy, the separate bias add and the separate GELU pass.  Also used for the fused
// Conv + BN + SiLU of the head (common.py:38-50 after Model.fuse()).  Math in fp32, erff like torch's exact GELU.
#include "common.cuh"

namespace sodt {
namespace {

template <int ACT>
__device__ __forceinline__ float activate(float v) {
    if constexpr (ACT == 1) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    else if constexpr (ACT == 2) return v / (1.f + __expf(-v));
    else return v;
}

template <typename T, int ACT>
__global__ void __launch_bounds__(256)
bias_act_crop_kernel(const T* __restrict__ in, const float* __restrict__ bias, T* __restrict__ out,
                     int H, int W, int C, int inH, int inW, int oy, int ox, long long total_vec) {
    constexpr int PER = 16 / sizeof(T);
    const int vpp = C / PER;   // vectors per pixel
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total_vec; e += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(e % vpp);
        const long long pix = e / vpp;
        const int x = (int)(pix % W);
        const long long t = pix / W;
        const int y = (int)(t % H);
        const long long b = t / H;
        const uint4 raw = *reinterpret_cast<const uint4*>(in + (((b * inH + y + oy) * inW + x + ox) * (long long)C) + v * PER);
        const float* bp = bias + v * PER;
        uint4 o;
        if constexpr (sizeof(T) == 4) {
            const float4 bb = *reinterpret_cast<const float4*>(bp);
            o.x = __float_as_uint(activate<ACT>(__uint_as_float(raw.x) + bb.x));
            o.y = __float_as_uint(activate<ACT>(__uint_as_float(raw.y) + bb.y));
            o.z = __float_as_uint(activate<ACT>(__uint_as_float(raw.z) + bb.z));
            o.w = __float_as_uint(activate<ACT>(__uint_as_float(raw.w) + bb.w));
        } else {
            const float4 b0 = *reinterpret_cast<const float4*>(bp), b1 = *reinterpret_cast<const float4*>(bp + 4);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
            __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
            ho[0] = __floats2bfloat162_rn(activate<ACT>(__low2float(h[0]) + b0.x), activate<ACT>(__high2float(h[0]) + b0.y));
            ho[1] = __floats2bfloat162_rn(activate<ACT>(__low2float(h[1]) + b0.z), activate<ACT>(__high2float(h[1]) + b0.w));
            ho[2] = __floats2bfloat162_rn(activate<ACT>(__low2float(h[2]) + b1.x), activate<ACT>(__high2float(h[2]) + b1.y));
            ho[3] = __floats2bfloat162_rn(activate<ACT>(__low2float(h[3]) + b1.z), activate<ACT>(__high2float(h[3]) + b1.w));
        }
        *reinterpret_cast<uint4*>(out + pix * C + v * PER) = o;
    }
}

template <typename T>
int launch(const void* in, const float* bias, void* out, int B, int H, int W, int C, int inH, int inW, int oy, int ox,
           int act, cudaStream_t stream) {
    constexpr int PER = 16 / sizeof(T);
    const long long total = (long long)B * H * W * (C / PER);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    const T* i = static_cast<const T*>(in);
    T* o = static_cast<T*>(out);
    if (act == 1) bias_act_crop_kernel<T, 1><<<(unsigned)blocks, 256, 0, stream>>>(i, bias, o, H, W, C, inH, inW, oy, ox, total);
    else if (act == 2) bias_act_crop_kernel<T, 2><<<(unsigned)blocks, 256, 0, stream>>>(i, bias, o, H, W, C, inH, inW, oy, ox, total);
    else bias_act_crop_kernel<T, 0><<<(unsigned)blocks, 256, 0, stream>>>(i, bias, o, H, W, C, inH, inW, oy, ox, total);
    return check_launch();
}

}  // namespace
}  // namespace sodt

extern "C" int sodt_bias_act_crop_nhwc(const void* in, const float* bias, void* out, int B, int H, int W, int C,
                                       int in_H, int in_W, int off_y, int off_x, int act, int dtype, void* stream) {
    using namespace sodt;
    if (!in || !bias || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0) return SODT_ERR_INVALID_ARG;
    if (off_y < 0 || off_x < 0 || off_y + H > in_H || off_x + W > in_W || act < 0 || act > 2) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_F32 && dtype != SODT_BF16) return SODT_ERR_INVALID_ARG;
    const int per = dtype == SODT_F32 ? 4 : 8;
    if (C % per) return SODT_ERR_UNSUPPORTED;
    if (!aligned16(in) || !aligned16(out) || !aligned16(bias)) return SODT_ERR_ALIGNMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == SODT_F32) return launch<float>(in, bias, out, B, H, W, C, in_H, in_W, off_y, off_x, act, s);
    return launch<__nv_bfloat16>(in, bias, out, B, H, W, C, in_H, in_W, off_y, off_x, act, s);
}
