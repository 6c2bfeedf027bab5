// Shared helpers for the sodt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <utility>

#include "../../include/sodt_b200.h"

namespace sodt {

extern thread_local long long g_launches;
extern thread_local cudaError_t g_last_cuda_error;

inline int check_launch() {
    ++g_launches;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = e;
        (void)cudaGetLastError();
        return SODT_ERR_CUDA;
    }
    return SODT_OK;
}

inline int cuda_status(cudaError_t e) {
    if (e == cudaSuccess) return SODT_OK;
    g_last_cuda_error = e;
    return SODT_ERR_CUDA;
}

// Programmatic dependent launch.  The kernels of the detector step are launched with programmatic stream serialization: their
// CTAs may become resident while the previous kernel of the stream is still draining (each kernel calls pdl_trigger() at its
// start, so its dependents are released as soon as every CTA of the grid has started), run their prologue (barrier
// initialisation, tensor-memory allocation) and then block in pdl_wait() until the previous grid has completed and its memory
// is visible.  NO global memory is read or written before pdl_wait() -- not even weights or prepared tables: a caller may have
// produced them with the launch just before.  What overlaps is the launch latency, the CTA ramp-up and the prologue.
// SODT_PDL=0 in the environment launches every kernel fully serialised (the griddepcontrol instructions are no-ops then).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("SODT_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

// Launches `kern` with the programmatic-stream-serialization attribute (`pdl` false: a plain, fully serialised launch).
// Only for kernels that call pdl_wait() before their first global memory access.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl && pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Geometry of the (rolled, padded) window grid shared by the attention kernels.
// Reference: window_partition / roll, basics/models/backbone_vit.py:619-643,1096.
struct WinGeom {
    int H, W, ws, shift, nwh, nww;
    __host__ __device__ WinGeom(int H_, int W_, int ws_, int shift_)
        : H(H_), W(W_), ws(ws_), shift(shift_), nwh((H_ + ws_ - 1) / ws_), nww((W_ + ws_ - 1) / ws_) {}
    // token t of window w -> coordinates in the rolled frame; false if it is a padding token
    __device__ __forceinline__ bool rolled(int win, int t, int& yr, int& xr) const {
        int wy = win / nww, wx = win - wy * nww;
        int ty = t / ws, tx = t - ty * ws;
        yr = wy * ws + ty;
        xr = wx * ws + tx;
        return yr < H && xr < W;
    }
    // rolled frame -> source / destination pixel of the un-rolled image
    __device__ __forceinline__ void source(int yr, int xr, int& ys, int& xs) const {
        ys = yr + shift; if (ys >= H) ys -= H;
        xs = xr + shift; if (xs >= W) xs -= W;
    }
    // region id of the shifted-window mask (backbone_vit.py:1060-1072); padding tokens carry 0
    __device__ __forceinline__ int region(int yr, int xr) const {
        if (yr >= H || xr >= W) return 0;
        int ry = (yr >= H - ws) + (yr >= H - shift);
        int rx = (xr >= W - ws) + (xr >= W - shift);
        return 3 * ry + rx;
    }
};

}  // namespace sodt
