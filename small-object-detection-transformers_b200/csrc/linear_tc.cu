// bf16 Linear layers of the Swin blocks on tcgen05 tensor cores with fused epilogues:
//
//   out[M, N] = act(x[M, K] . W[N, K]^T + bias[N]) (+ residual[M, N])          act = identity | exact GELU
//
// replaces nn.Linear + the separate GELU pass of Mlp (reference basics/models/backbone_vit.py:885-890,968,990) and the
// separate residual add of SwinTransformerBlock (:1125,:1128).  These GEMMs are HBM-bound (K = 192..768 on millions of
// token rows), so removing the extra elementwise passes over the 4C-wide hidden tensor is worth more than MMA efficiency.
//
// Persistent CTAs (one per SM), tile 128 x BN (BN = 256 or 192), k-block 64:
//   warp 16     TMA producer: cp.async.bulk.tensor 2-D boxes of x and W, SWIZZLE_128B, 4-stage mbarrier ring
//   warp 17     MMA issuer:   tcgen05.mma kind::f16, M=128, N=BN, K-major SWIZZLE_128B operands, fp32 accumulators in TMEM,
//                             two accumulator buffers (2 x BN columns) so the epilogue of tile t overlaps the MMAs of t+1
//   warps 0-15  epilogue:     four groups of 4 warps take the tile's 64-column boxes (residual box TMA-loaded into the group's
//                             staging buffer first): tcgen05.ld (warp w: TMEM lanes
//                             32*(w%4)..), bias, GELU (erf by A&S 7.1.26, |error| < 2e-7, far below bf16), residual, bf16
//                             pack into a SWIZZLE_128B staging box in shared memory, one TMA store per box (full 128-byte
//                             lines; per-thread 16-byte row stores capped the kernel at ~1.9 TB/s)
#include <cuda.h>

#include "common.cuh"
#include "tc05.cuh"

namespace sodt {
namespace {

using namespace tc;

constexpr int BM = 128, BK = 64;
constexpr int BOX_BYTES = BM * 64 * 2;      // one 128-row x 64-column bf16 output box
constexpr int A_BYTES = BM * BK * 2;
constexpr int EPI_GROUPS = 4;                  // epilogue groups of 4 warps (one 64-column box each at a time)
constexpr int EPI_WARPS = EPI_GROUPS * 4;
constexpr int PRODUCER_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1;
constexpr int NTHREADS = (EPI_WARPS + 2) * 32;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// K-major SWIZZLE_128B operand tile: rows of 128 B (64 bf16), 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// exact-GELU with erf from Abramowitz & Stegun 7.1.26 (max abs error 1.5e-7): one MUFU.RCP, one MUFU.EX2, ~10 FMA
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = p * t * fast_exp2(-z * z * 1.4426950408889634f);   // 1 - erf(z)
    const float phi = x >= 0.f ? 1.f - 0.5f * e : 0.5f * e;            // Phi(x)
    return x * phi;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_r,
                 const float* __restrict__ bias, int has_residual, int M, int N, int K, int act, int num_n_tiles, int num_tiles) {
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2], res_full[EPI_GROUPS];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int nkb = K / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], EPI_WARPS * 32); }
        for (int g = 0; g < EPI_GROUPS; ++g) mbar_init(&res_full[g], 1);
        fence_barrier_init();
    }
    if (warp == MMA_WARP) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_slot;

    if (warp == PRODUCER_WARP) {
        if (lane == 0) {
            int stage = 0, round = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mt = tile / num_n_tiles, nt = tile - mt * num_n_tiles;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (round > 0) mbar_wait(&empty[stage], (uint32_t)((round - 1) & 1));
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    const uint32_t sa = sbase + stage * STAGE_BYTES;
                    tma_load_2d(sa, &tmap_x, &full[stage], kb * BK, mt * BM);
                    tma_load_2d(sa + A_BYTES, &tmap_w, &full[stage], kb * BK, nt * BN);
                    if (++stage == STAGES) { stage = 0; ++round; }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16(BM, BN, false, false);
            int stage = 0, round = 0, it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int a = it & 1;
                if (it >= 2) mbar_wait(&acc_empty[a], (uint32_t)(((it >> 1) - 1) & 1));
                fence_after_sync();
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[stage], (uint32_t)(round & 1));
                    fence_after_sync();
                    const uint64_t da = desc_sw128(sbase + stage * STAGE_BYTES), db = desc_sw128(sbase + stage * STAGE_BYTES + A_BYTES);
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) mma_ss(tm + a * BN, da + 2 * ks, db + 2 * ks, idesc, (kb | ks) != 0);
                    mma_commit(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; ++round; }
                }
                mma_commit(&acc_full[a]);
            }
        }
    } else {
        // epilogue group eg = warp/4 takes the tile's 64-column boxes eg, eg + EPI_GROUPS, ...; thread = one row of the box
        const int quarter = warp & 3, eg = warp >> 2;
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t stage_box = sbase + STAGES * STAGE_BYTES + eg * BOX_BYTES;       // this group's staging box
        const uint32_t my_row = stage_box + row_in_tile * 128;
        const int sw = row_in_tile & 7;                                                   // SWIZZLE_128B: chunk ^= row % 8
        const bool issuer = quarter == 0 && lane == 0;
        constexpr int NBOX = BN / 64;
        int it = 0;
        uint32_t res_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int a = it & 1;
            const int mt = tile / num_n_tiles, nt = tile - mt * num_n_tiles;
            mbar_wait(&acc_full[a], (uint32_t)((it >> 1) & 1));
            fence_after_sync();
#pragma unroll 1
            for (int bx = eg; bx < NBOX; bx += EPI_GROUPS) {
                const int col0 = nt * BN + bx * 64;
                if (issuer) {
                    tma_store_wait_read();                               // the previous box of this group has left smem
                    if (has_residual) {                                  // residual box lands in the staging buffer (TMA, swizzled)
                        mbar_expect_tx(&res_full[eg], BOX_BYTES);
                        tma_load_2d(stage_box, &tmap_r, &res_full[eg], col0, mt * BM);
                    }
                }
                uint32_t r0[32], r1[32];
                tmem_ld32(tm + lane_addr + a * BN + bx * 64, r0);
                tmem_ld32(tm + lane_addr + a * BN + bx * 64 + 32, r1);
                tmem_wait_ld();
                asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");      // staging box reusable for everyone
                if (has_residual) { mbar_wait(&res_full[eg], res_phase); res_phase ^= 1; }
#pragma unroll
                for (int j = 0; j < 64; j += 8) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(j < 32 ? r0[j + e] : r1[j - 32 + e]);
                    if (bias != nullptr) {
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col0 + j + 4));
                        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                    }
                    if (act == 1) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = gelu_erf(v[e]);
                    }
                    const uint32_t dst = my_row + (((j >> 3) ^ sw) << 4);
                    if (has_residual) {
                        uint32_t q0, q1, q2, q3;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3) : "r"(dst) : "memory");
                        const uint32_t qq[4] = {q0, q1, q2, q3};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&qq[e]);
                            v[2 * e] += __low2float(h); v[2 * e + 1] += __high2float(h);
                        }
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16(v[0], v[1])), "r"(pack_bf16(v[2], v[3])),
                                 "r"(pack_bf16(v[4], v[5])), "r"(pack_bf16(v[6], v[7])) : "memory");
                }
                fence_proxy_async();
                asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
                if (issuer) { tma_store_2d(&tmap_o, stage_box, col0, mt * BM); tma_store_commit(); }
            }
            fence_before_sync();
            mbar_arrive(&acc_empty[a]);
        }
        if (issuer) tma_store_wait_all();
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_slot, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

bool make_map(CUtensorMap* m, const void* base, int rows, int K, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows}, es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_out_map(CUtensorMap* m, void* base, int M, int N) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)N * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BM}, es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int STAGES>
int launch(const void* x, const void* w, const float* bias, const void* residual, void* out, int M, int N, int K, int act,
           int num_sms, cudaStream_t stream) {
    CUtensorMap mx, mw, mo, mr;
    if (!make_map(&mx, x, M, K, BM) || !make_map(&mw, w, N, K, BN) || !make_out_map(&mo, out, M, N)) return SODT_ERR_CUDA;
    if (!make_out_map(&mr, const_cast<void*>(residual ? residual : out), M, N)) return SODT_ERR_CUDA;
    const int num_n_tiles = N / BN, num_m_tiles = (M + BM - 1) / BM;
    const long long tiles = (long long)num_n_tiles * num_m_tiles;
    if (tiles > 2147483647LL) return SODT_ERR_UNSUPPORTED;
    const size_t smem = (size_t)STAGES * (A_BYTES + BN * BK * 2) + EPI_GROUPS * BOX_BYTES + 1024;
    auto kern = linear_tc_kernel<BN, STAGES>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    const int grid = (int)(tiles < num_sms ? tiles : num_sms);
    kern<<<grid, NTHREADS, smem, stream>>>(mx, mw, mo, mr, bias, residual != nullptr ? 1 : 0, M, N, K, act, num_n_tiles, (int)tiles);
    return check_launch();
}

}  // namespace

int linear_tc(const void* x, const void* w, const float* bias, const void* residual, void* out, int M, int N, int K, int act,
              int num_sms, cudaStream_t stream) {
    if (N % 256 == 0) return launch<256, 3>(x, w, bias, residual, out, M, N, K, act, num_sms, stream);
    return launch<192, 4>(x, w, bias, residual, out, M, N, K, act, num_sms, stream);
}

bool linear_tc_supported(int M, int N, int K) { return M > 0 && K % BK == 0 && K >= BK && (N % 256 == 0 || N % 192 == 0); }

}  // namespace sodt
