// Probe: TMA (SWIZZLE_128B) -> tcgen05.mma with SWIZZLE_128B K-major descriptors -> TMEM -> global.
// C[128 x 128] = A[128 x K] * W[128 x K]^T with K = 192 (3 k-blocks of 64).  Prints max abs error.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include "../../small-object-detection-transformers_b200/csrc/tc05.cuh"

using namespace sodt::tc;
constexpr int M = 128, N = 128, K = 192, BK = 64;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// K-major SWIZZLE_128B descriptor: 8-row groups 1024 B apart, layout type 2
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major), canonical value 1
    d |= (uint64_t)(1024 >> 4) << 32;       // SBO
    d |= (uint64_t)1 << 46;                 // version
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, float* C) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[3], done;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { for (int i = 0; i < 3; ++i) mbar_init(&full[i], 1); mbar_init(&done, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&slot, 128); tmem_relinquish(); }
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tm = slot;
    const uint32_t sA = smem_u32(smem), sB = sA + 3 * 16384;
    if (tid == 0) {
        for (int kb = 0; kb < 3; ++kb) {
            mbar_expect_tx(&full[kb], 2 * 16384);
            tma_load_2d(sA + kb * 16384, &ta, &full[kb], kb * BK, 0);
            tma_load_2d(sB + kb * 16384, &tb, &full[kb], kb * BK, 0);
        }
        const uint32_t idesc = idesc_bf16(128, 128, false, false);
        for (int kb = 0; kb < 3; ++kb) {
            mbar_wait(&full[kb], 0);
            fence_after_sync();
            for (int ks = 0; ks < 4; ++ks)
                mma_ss(tm, desc_sw128(sA + kb * 16384 + ks * 32), desc_sw128(sB + kb * 16384 + ks * 32), idesc, (kb | ks) != 0);
        }
        mma_commit(&done);
    }
    mbar_wait(&done, 0);
    fence_after_sync();
    for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c * 32, r);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) C[tid * N + c * 32 + j] = __uint_as_float(r[j]);
    }
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 128);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K);
    srand(2);
    auto rnd = [] { return (rand() % 2001 - 1000) / 1000.0f; };
    for (int i = 0; i < M * K; ++i) { hA[i] = __float2bfloat16(rnd()); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { hB[i] = __float2bfloat16(rnd()); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dC;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dC, M * N * 4);
    cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&enc), cudaEnableDefault, &q);
    if (e != cudaSuccess || !enc) { printf("no cuTensorMapEncodeTiled: %s\n", cudaGetErrorString(e)); return 1; }
    CUtensorMap ta, tb;
    cuuint64_t dimsA[2] = {K, M}, strA[1] = {K * 2}, dimsB[2] = {K, N};
    cuuint32_t box[2] = {BK, 128}, es[2] = {1, 1};
    CUresult r1 = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, strA, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, strA, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
    size_t smem = 6 * 16384 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(ta, tb, dC);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> C(M * N);
    cudaMemcpy(C.data(), dC, M * N * 4, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            float acc = 0; for (int k = 0; k < K; ++k) acc += fA[i * K + k] * fB[j * K + k];
            err = fmax(err, fabs(acc - C[i * N + j]));
        }
    printf("TMA SW128 + UMMA SW128 GEMM probe: max abs err %.3e (C[0][0..2] = %.4f %.4f %.4f)\n", err, C[0], C[1], C[2]);
    return 0;
}
