// Standalone probe of the tcgen05 building blocks used by the attention kernels (run on a B200):
//   T1  S = Q K^T          SS-MMA, both operands K-major, SWIZZLE_NONE canonical layout
//   T2  O = P V            SS-MMA, A = P (K-major, K = 128 keys), B = V (MN-major)
//   T3  O = P V            TS-MMA, A = P read from TMEM (bf16 pairs packed in 32-bit columns)
// Prints max abs error of each against a host fp32 reference.  Usage: umma_probe <1|2|3>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../small-object-detection-transformers_b200/csrc/tc05.cuh"

using namespace sodt::tc;
constexpr int M = 128, NK = 128, HD = 64;

// canonical tile: [chunk c = col/8][row][8 elems] ; bytes: c*rows*16 + row*16
__device__ void fill_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int rows, int cols) {
    for (int e = threadIdx.x; e < rows * cols / 8; e += blockDim.x) {
        int row = e % rows, c = e / rows;
        *reinterpret_cast<uint4*>(dst + (size_t)(c * rows + row) * 8) = *reinterpret_cast<const uint4*>(src + (size_t)row * cols + c * 8);
    }
}

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* V,
                                             float* S_out, float* O_out, int variant) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);          // 128 x 64
    __nv_bfloat16* sK = sQ + M * HD;                                     // 128 x 64
    __nv_bfloat16* sV = sK + NK * HD;                                    // 128 keys x 64 dims
    __nv_bfloat16* sP = sV + NK * HD;                                    // 128 x 128
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    fill_tile(sQ, Q, M, HD);
    fill_tile(sK, K, NK, HD);
    fill_tile(sV, V, NK, HD);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t phase = 0;
    // ---- T1: S[128x128] = Q K^T, 4 K-steps of 16
    if (tid == 0) {
        const uint32_t idesc = idesc_bf16(128, 128, false, false);
        for (int k = 0; k < HD / 16; ++k) {
            uint64_t a = smem_desc(smem_u32(sQ) + k * 2 * (M * 16), M * 16, 128);
            uint64_t b = smem_desc(smem_u32(sK) + k * 2 * (NK * 16), NK * 16, 128);
            mma_ss(tm, a, b, idesc, k > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, phase); phase ^= 1;
    fence_after_sync();
    float srow[128];
    for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tm + lane_base + c * 32, r);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) srow[c * 32 + j] = __uint_as_float(r[j]);
    }
    for (int j = 0; j < 128; ++j) S_out[tid * 128 + j] = srow[j];
    if (variant == 1) goto done;
    if (variant == 7) {
        // T7: TMEM load / store throughput, all 4 warps (128 lanes) concurrently
        __syncthreads();
        long long t0 = clock64();
        uint32_t acc = 0;
        for (int i = 0; i < 64; ++i) {
            uint32_t r0[32], r1[32];
            tmem_ld32(tm + lane_base + ((i & 1) * 64), r0);
            tmem_ld32(tm + lane_base + ((i & 1) * 64) + 32, r1);
            tmem_wait_ld();
            for (int j = 0; j < 32; ++j) acc ^= r0[j] ^ r1[j];
        }
        long long t1 = clock64();
        if (tid == 0) printf("LDTM: 128 lanes x 64 cols (32 KB) per iteration: %6.1f cycles  -> %5.1f B/cycle/SM (acc %u)\n",
                             (double)(t1 - t0) / 64, 32768.0 * 64 / (double)(t1 - t0), acc & 1);
        __syncthreads();
        t0 = clock64();
        for (int i = 0; i < 64; ++i) {
            uint32_t r0[32];
            for (int j = 0; j < 32; ++j) r0[j] = acc + j + i;
            tmem_st32(tm + lane_base + 256 + ((i & 1) * 32), r0);
            tmem_wait_st();
        }
        t1 = clock64();
        if (tid == 0) printf("STTM: 128 lanes x 32 cols (16 KB) per iteration: %6.1f cycles  -> %5.1f B/cycle/SM\n",
                             (double)(t1 - t0) / 64, 16384.0 * 64 / (double)(t1 - t0));
        // one warp only
        __syncthreads();
        if (warp == 0) {
            t0 = clock64();
            for (int i = 0; i < 64; ++i) {
                uint32_t r0[32], r1[32];
                tmem_ld32(tm + ((i & 1) * 64), r0);
                tmem_ld32(tm + ((i & 1) * 64) + 32, r1);
                tmem_wait_ld();
                for (int j = 0; j < 32; ++j) acc ^= r0[j] ^ r1[j];
            }
            t1 = clock64();
            if (tid == 0) printf("LDTM one warp: 32 lanes x 64 cols (8 KB): %6.1f cycles -> %5.1f B/cycle (acc %u)\n",
                                 (double)(t1 - t0) / 64, 8192.0 * 64 / (double)(t1 - t0), acc & 1);
        }
        __syncthreads();
        goto done;
    }
    if (variant == 6) {
        // T6: issue-rate microbenchmark: cycles per tcgen05.mma for the shapes the window kernel uses
        __syncthreads();
        if (tid == 0) {
            const int REP = 512;
            long long t0, t1;
            const uint32_t ids = idesc_bf16(128, 64, false, false), ido16 = idesc_bf16(128, 16, false, true),
                           ido32 = idesc_bf16(128, 32, false, true), ids128 = idesc_bf16(128, 128, false, false);
            uint64_t a = smem_desc(smem_u32(sQ), M * 16, 128), b = smem_desc(smem_u32(sK), NK * 16, 128);
            uint64_t bv = smem_desc(smem_u32(sV), 128, NK * 16);
            t0 = clock64();
            for (int i = 0; i < REP; ++i) mma_ss(tm, a, b, ids128, i > 0);
            mma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1; t1 = clock64();
            printf("SS M128 N128 K16        : %6.1f cycles per MMA\n", (double)(t1 - t0) / REP);
            t0 = clock64();
            for (int i = 0; i < REP; ++i) mma_ss_masked(tm, a, b, ids, i > 0, 0u, 0u, ~0u, ~0u);
            mma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1; t1 = clock64();
            printf("SS M128 N64  K16 masked : %6.1f cycles per MMA\n", (double)(t1 - t0) / REP);
            t0 = clock64();
            for (int i = 0; i < REP; ++i) mma_ts(tm + 128, tm + 256, bv, ido16, i > 0);
            mma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1; t1 = clock64();
            printf("TS M128 N16  K16        : %6.1f cycles per MMA\n", (double)(t1 - t0) / REP);
            t0 = clock64();
            for (int i = 0; i < REP; ++i) mma_ts_masked(tm + 128, tm + 256, bv, ido16, i > 0, 0u, 0u, ~0u, ~0u);
            mma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1; t1 = clock64();
            printf("TS M128 N16  K16 masked : %6.1f cycles per MMA\n", (double)(t1 - t0) / REP);
            t0 = clock64();
            for (int i = 0; i < REP; ++i) mma_ts(tm + 128, tm + 256, bv, ido32, i > 0);
            mma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1; t1 = clock64();
            printf("TS M128 N32  K16        : %6.1f cycles per MMA\n", (double)(t1 - t0) / REP);
            uint64_t bv64 = smem_desc(smem_u32(sV), 128, NK * 16);
            const uint32_t ido64 = idesc_bf16(128, 64, false, true);
            t0 = clock64();
            for (int i = 0; i < REP; ++i) mma_ts(tm + 128, tm + 256, bv64, ido64, i > 0);
            mma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1; t1 = clock64();
            printf("TS M128 N64  K16        : %6.1f cycles per MMA\n", (double)(t1 - t0) / REP);
            // independent accumulators, round-robin: is ~64 cycles a dependent-chain latency or the throughput floor?
            for (int nacc = 2; nacc <= 8; nacc *= 2) {
                t0 = clock64();
                for (int i = 0; i < REP; ++i) mma_ts(tm + 128 + (i % nacc) * 16, tm + 256, bv, ido16, i >= nacc);
                mma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1; t1 = clock64();
                printf("TS M128 N16  K16, %d independent accumulators : %6.1f cycles per MMA\n", nacc, (double)(t1 - t0) / REP);
            }
            for (int nacc = 2; nacc <= 4; nacc *= 2) {
                t0 = clock64();
                for (int i = 0; i < REP; ++i) mma_ss(tm + (i % nacc) * 64, a, b, ids, i >= nacc);
                mma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1; t1 = clock64();
                printf("SS M128 N64  K16, %d independent accumulators : %6.1f cycles per MMA\n", nacc, (double)(t1 - t0) / REP);
            }
            for (int nacc = 2; nacc <= 4; nacc *= 2) {   // masked halves: same columns, disjoint lanes
                t0 = clock64();
                for (int i = 0; i < REP; ++i) {
                    const uint32_t A = 0xFFFFFFFFu; const bool up = i & 1;
                    mma_ts_masked(tm + 128 + ((i >> 1) % nacc) * 16, tm + 256, bv, ido16, i >= 2 * nacc, up ? A : 0u, up ? A : 0u, up ? 0u : A, up ? 0u : A);
                }
                mma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1; t1 = clock64();
                printf("TS M128 N16 masked halves, %d accumulators x 2 lane halves : %6.1f cycles per MMA\n", nacc, (double)(t1 - t0) / REP);
            }
            {   // the window kernel's per-head issue stream: 2 masked SS (N=64) + commit, 8 masked TS (N=16) + 2 commits
                const uint32_t A = 0xFFFFFFFFu;
                t0 = clock64();
                for (int i = 0; i < 128; ++i) {
                    mma_ss_masked(tm, a, b, ids, false, 0u, 0u, A, A);
                    mma_ss_masked(tm, a, b, ids, false, A, A, 0u, 0u);
                    mma_commit(&bar);
                    for (int h = 0; h < 2; ++h)
                        for (int ks = 0; ks < 4; ++ks)
                            mma_ts_masked(tm + 128, tm + 256 + ks * 8, smem_desc(smem_u32(sV) + h * 64 * 16 + ks * 256, 128, NK * 16), ido16, ks > 0,
                                          h ? A : 0u, h ? A : 0u, h ? 0u : A, h ? 0u : A);
                    mma_commit(&bar);
                    mma_commit(&bar);
                }
                t1 = clock64();
                printf("win8 per-head stream (10 MMA + 3 commits), issue only : %6.1f cycles per head\n", (double)(t1 - t0) / 128);
                // drain: 384 arrivals on a count-1 barrier = 384 phases; just wait long enough
                for (volatile int w = 0; w < 200000; ++w) {}
                mbar_init(&bar, 1); phase = 0;
                // unmasked alternative: 1 SS N=128 + commit, 8 TS N=16 K over 128 keys + 2 commits
                t0 = clock64();
                for (int i = 0; i < 128; ++i) {
                    mma_ss(tm, a, b, ids128, false);
                    mma_commit(&bar);
                    for (int ks = 0; ks < 8; ++ks)
                        mma_ts(tm + 128, tm + 256 + ks * 8, smem_desc(smem_u32(sV) + ks * 256, 128, NK * 16), ido16, ks > 0);
                    mma_commit(&bar);
                    mma_commit(&bar);
                }
                t1 = clock64();
                printf("unmasked per-head stream (9 MMA + 3 commits)          : %6.1f cycles per head\n", (double)(t1 - t0) / 128);
                for (volatile int w = 0; w < 200000; ++w) {}
                mbar_init(&bar, 1); phase = 0;
            }
            // latency of one dependent commit round trip
            t0 = clock64();
            for (int i = 0; i < 64; ++i) { mma_ss(tm, a, b, ids128, false); mma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1; }
            t1 = clock64();
            printf("SS N128 issue+commit+wait round trip: %6.1f cycles\n", (double)(t1 - t0) / 64);
        }
        __syncthreads();
        goto done;
    }
    if (variant >= 4) {
        // T4: S2[128 x 64]: rows 0-63 = Q K[0:64]^T, rows 64-127 = Q K[64:128]^T via two N=64 MMAs with lane masks
        // (variant 4: bit set = lane disabled; variant 5: inverted polarity)
        __syncthreads();
        if (tid == 0) {
            const uint32_t idesc = idesc_bf16(128, 64, false, false);
            const uint32_t lo = variant == 4 ? 0u : 0xFFFFFFFFu, hi = ~lo;
            for (int half = 0; half < 2; ++half)
                for (int k = 0; k < HD / 16; ++k) {
                    uint64_t a = smem_desc(smem_u32(sQ) + k * 2 * (M * 16), M * 16, 128);
                    uint64_t b = smem_desc(smem_u32(sK) + half * 64 * 16 + k * 2 * (NK * 16), NK * 16, 128);
                    // half 0 writes rows 0-63 (disable lanes 64-127), half 1 writes rows 64-127
                    if (half == 0) mma_ss_masked(tm + 320, a, b, idesc, k > 0, lo, lo, hi, hi);
                    else mma_ss_masked(tm + 320, a, b, idesc, k > 0, hi, hi, lo, lo);
                }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1;
        fence_after_sync();
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(tm + lane_base + 320 + c * 32, r);
            tmem_wait_ld();
            for (int j = 0; j < 32; ++j) O_out[tid * 64 + c * 32 + j] = __uint_as_float(r[j]);
        }
        goto done;
    }
    // P = bf16(0.05 * S)
    if (variant == 2) {
        for (int c = 0; c < 16; ++c) {
            uint4 v;
            v.x = pack_bf16(0.05f * srow[c * 8 + 0], 0.05f * srow[c * 8 + 1]);
            v.y = pack_bf16(0.05f * srow[c * 8 + 2], 0.05f * srow[c * 8 + 3]);
            v.z = pack_bf16(0.05f * srow[c * 8 + 4], 0.05f * srow[c * 8 + 5]);
            v.w = pack_bf16(0.05f * srow[c * 8 + 6], 0.05f * srow[c * 8 + 7]);
            *reinterpret_cast<uint4*>(sP + (size_t)(c * M + tid) * 8) = v;
        }
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        if (tid == 0) {
            const uint32_t idesc = idesc_bf16(128, 64, false, true);   // B = V is MN-major
            for (int k = 0; k < NK / 16; ++k) {
                uint64_t a = smem_desc(smem_u32(sP) + k * 2 * (M * 16), M * 16, 128);
                // V tile [d chunk][key][16B]: 8-key groups 128 B apart (lbo), d chunks NK*16 B apart (sbo)
                uint64_t b = smem_desc(smem_u32(sV) + k * 256, 128, NK * 16);
                mma_ss(tm + 128, a, b, idesc, k > 0);
            }
            mma_commit(&bar);
        }
    } else {
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            for (int j = 0; j < 32; ++j) r[j] = pack_bf16(0.05f * srow[c * 64 + 2 * j], 0.05f * srow[c * 64 + 2 * j + 1]);
            tmem_st32(tm + lane_base + 256 + c * 32, r);
        }
        tmem_wait_st();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        if (tid == 0) {
            const uint32_t idesc = idesc_bf16(128, 64, false, true);
            for (int k = 0; k < NK / 16; ++k) {
                uint64_t b = smem_desc(smem_u32(sV) + k * 256, 128, NK * 16);
                mma_ts(tm + 128, tm + 256 + k * 8, b, idesc, k > 0);   // 16 bf16 of K = 8 packed columns
            }
            mma_commit(&bar);
        }
    }
    mbar_wait(&bar, phase); phase ^= 1;
    fence_after_sync();
    for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(tm + lane_base + 128 + c * 32, r);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) O_out[tid * 64 + c * 32 + j] = __uint_as_float(r[j]);
    }
done:
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 1;
    std::vector<__nv_bfloat16> hQ(M * HD), hK(NK * HD), hV(NK * HD);
    std::vector<float> fQ(M * HD), fK(NK * HD), fV(NK * HD);
    srand(1);
    auto rnd = [] { return (rand() % 2001 - 1000) / 1000.0f; };
    for (int i = 0; i < M * HD; ++i) { fQ[i] = bf(rnd()); hQ[i] = __float2bfloat16(fQ[i]); }
    for (int i = 0; i < NK * HD; ++i) { fK[i] = bf(rnd()); hK[i] = __float2bfloat16(fK[i]); fV[i] = bf(rnd()); hV[i] = __float2bfloat16(fV[i]); }
    __nv_bfloat16 *dQ, *dK, *dV; float *dS, *dO;
    cudaMalloc(&dQ, M * HD * 2); cudaMalloc(&dK, NK * HD * 2); cudaMalloc(&dV, NK * HD * 2);
    cudaMalloc(&dS, M * NK * 4); cudaMalloc(&dO, M * HD * 4);
    cudaMemcpy(dQ, hQ.data(), M * HD * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dK, hK.data(), NK * HD * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dV, hV.data(), NK * HD * 2, cudaMemcpyHostToDevice);
    cudaMemset(dS, 0, M * NK * 4); cudaMemset(dO, 0, M * HD * 4);
    size_t smem = (size_t)(M * HD + 2 * NK * HD + M * NK) * 2 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(dQ, dK, dV, dS, dO, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    std::vector<float> S(M * NK), O(M * HD);
    cudaMemcpy(S.data(), dS, M * NK * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(O.data(), dO, M * HD * 4, cudaMemcpyDeviceToHost);
    double es = 0, eo = 0;
    std::vector<float> P(M * NK);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < NK; ++j) {
            float acc = 0; for (int d = 0; d < HD; ++d) acc += fQ[i * HD + d] * fK[j * HD + d];
            es = fmax(es, fabs(acc - S[i * NK + j]));
            P[i * NK + j] = bf(0.05f * S[i * NK + j]);
        }
    for (int i = 0; i < M; ++i)
        for (int d = 0; d < HD; ++d) {
            float acc = 0; for (int j = 0; j < NK; ++j) acc += P[i * NK + j] * fV[j * HD + d];
            eo = fmax(eo, fabs(acc - O[i * HD + d]));
        }
    if (variant == 6 || variant == 7) return 0;
    if (variant >= 4) {
        double e4 = 0;
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < 64; ++j) e4 = fmax(e4, fabs(S[i * NK + (i / 64) * 64 + j] - O[i * HD + j]));
        printf("variant %d: masked two-half S max err vs block-diagonal of S: %.3e\n", variant, e4);
        return 0;
    }
    printf("variant %d: S max err %.3e   O max err %.3e   (S[0][0..3] = %.4f %.4f %.4f %.4f)\n", variant, es, eo, S[0], S[1], S[2], S[3]);
    return 0;
}
