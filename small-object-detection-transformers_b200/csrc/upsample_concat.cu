// Head glue of the YOLOv5 neck (models/model.yaml rows "nn.Upsample(None, 2, 'nearest')" + "Concat"):
// out[b, y, x, :] = [ low[b, y/2, x/2, :C1] | skip[b, y, x, :C2] ] on channels-last memory, one HBM pass
// (reference: basics/models/model.py:268-281 driving nn.Upsample and common.Concat, common.py:275-282).
// Not part of the attention path; it replaces two slow library kernels (nearest-upsample NHWC + cat).
#include "common.cuh"

namespace sodt {
namespace {

__global__ void __launch_bounds__(256)
upsample2x_concat_kernel(const uint4* __restrict__ low, const uint4* __restrict__ skip, uint4* __restrict__ out,
                         int H2, int W2, int v1, int v2, long long total) {
    // one thread per 16-byte vector of the output; v1 / v2 = vectors per pixel of low / skip
    const int vp = v1 + v2;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(e % vp);
        const long long pix = e / vp;
        const int x = (int)(pix % W2);
        const long long t = pix / W2;
        const int y = (int)(t % H2);
        const long long b = t / H2;
        uint4 val;
        if (v < v1) val = low[((b * (H2 >> 1) + (y >> 1)) * (W2 >> 1) + (x >> 1)) * v1 + v];
        else val = skip[pix * v2 + (v - v1)];
        out[e] = val;
    }
}

}  // namespace
}  // namespace sodt

extern "C" int sodt_upsample2x_concat_nhwc(const void* low, const void* skip, void* out, int B, int H, int W, int C1, int C2,
                                           int elem_bytes, void* stream) {
    using namespace sodt;
    if (!low || !skip || !out || B <= 0 || H <= 0 || W <= 0 || C1 <= 0 || C2 <= 0) return SODT_ERR_INVALID_ARG;
    if ((C1 * elem_bytes) % 16 || (C2 * elem_bytes) % 16) return SODT_ERR_UNSUPPORTED;
    if (!aligned16(low) || !aligned16(skip) || !aligned16(out)) return SODT_ERR_ALIGNMENT;
    const int v1 = C1 * elem_bytes / 16, v2 = C2 * elem_bytes / 16;
    const long long total = (long long)B * (2 * H) * (2 * W) * (v1 + v2);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    upsample2x_concat_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(low), static_cast<const uint4*>(skip), static_cast<uint4*>(out), 2 * H, 2 * W, v1, v2, total);
    return check_launch();
}
