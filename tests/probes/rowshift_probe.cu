// Probe (run on a B200): may the A operand of a K-major SWIZZLE_128B tcgen05.mma start at a row that is NOT a multiple of 8,
// i.e. a start address 128 * s bytes into a 1024-byte-aligned tile?  The conv kernels want to read the taps kx = 0, 1, 2 of one
// halo tile [128 + kw - 1 pixels x 64 channels] as row-shifted views instead of loading every tap's box separately.
// For s = 0..3 and the descriptor's base-offset field = 0 or s: D[128 x 64] = A[s .. s + 127, :] B^T against a host reference.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../small-object-detection-transformers_b200/csrc/tma.cuh"

using namespace sodt::tc;
constexpr int ROWS = 144, N = 64, K = 64;

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* out, int shift, int base_off) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sp = smem_raw + (sbase - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5;
    // SWIZZLE_128B image of A (ROWS x 64) and B (64 x 64): row r, 16-byte chunk c at r * 128 + ((c ^ (r & 7)) << 4)
    for (int e = tid; e < ROWS * 8; e += 128) {
        const int r = e >> 3, c = e & 7;
        *reinterpret_cast<uint4*>(sp + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * K + c * 8);
    }
    const uint32_t boff = ROWS * 128;            // 18432 = 18 * 1024
    for (int e = tid; e < N * 8; e += 128) {
        const int r = e >> 3, c = e & 7;
        *reinterpret_cast<uint4*>(sp + boff + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * K + c * 8);
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&slot, 64); tmem_relinquish(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = slot;
    if (tid == 0) {
        const uint32_t idesc = idesc_bf16(128, N, false, false);
        const uint64_t da = sodt::tma::desc_sw128(sbase + shift * 128) | ((uint64_t)(base_off & 7) << 49);
        const uint64_t db = sodt::tma::desc_sw128(sbase + boff);
        for (int ks = 0; ks < K / 16; ++ks) mma_ss(tm, da + 2 * ks, db + 2 * ks, idesc, ks > 0);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    uint32_t r[32];
    for (int c = 0; c < 2; ++c) {
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c * 32, r);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) out[tid * N + c * 32 + j] = __uint_as_float(r[j]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 64);
}

int main() {
    std::vector<__nv_bfloat16> hA(ROWS * K), hB(N * K);
    std::vector<float> fA(ROWS * K), fB(N * K);
    srand(1);
    for (int i = 0; i < ROWS * K; ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB;
    float* dO;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    const int smem = ROWS * 128 + N * 128 + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> hO(128 * N);
    for (int s = 0; s < 10; ++s)
        for (int variant = 0; variant < 2; ++variant) {
            const int bo = variant ? (s & 7) : 0;
            if (variant && bo == 0) continue;
            probe<<<1, 128, smem>>>(dA, dB, dO, s, bo);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("shift %d base_off %d: CUDA error %s\n", s, bo, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
            double err = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < K; ++k) ref += (double)fA[(m + s) * K + k] * fB[n * K + k];
                    err = fmax(err, fabs(ref - hO[m * N + n]));
                }
            printf("shift %d rows, base_offset field %d: max abs err %.3e %s\n", s, bo, err, err < 1e-3 ? "OK" : "WRONG");
        }
    return 0;
}
