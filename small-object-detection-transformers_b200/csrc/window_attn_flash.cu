// Global-window attention on tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM).
//
// Serves the "window covers a large map" case of the backbone's last stage: windows of WS x WS
// tokens with N = WS*WS a multiple of 128 and head_dim 64 (stage 3: WS = 32, N = 1024,
// basics/models/backbone_vit.py:151-160), no shift.  Flash-style: one CTA owns 128 query rows of
// one (window, head) and streams the keys / values in tiles of 128 through shared memory.
//
// Operand tiles arrive by TMA: 128 tokens of a window = 4 window rows = the (64 channels, 32, 4) box of the [B*H, W, 3C] qkv
// image (window partition as box coordinates), landing as SWIZZLE_128B tiles of 128 rows x 128 B; the output tile leaves
// through a TMA store of the same box shape.
//
//   S = Q K^T          tcgen05.mma SS, M=128 N=128 K=64, both operands K-major SWIZZLE_128B (the descriptor start address
//                      advances 32 B per 16 channels), accumulator in TMEM columns [0,128)
//   softmax            one thread per query row reads its S row from TMEM (tcgen05.ld 32x32b), adds the
//                      relative-position bias (closed-form index into the shared-memory table,
//                      conflict free: the 32 lanes of a warp are 32 consecutive tokens of one window row),
//                      exp2 with a lazily updated running maximum, writes P as packed bf16 pairs back to
//                      TMEM columns [128,192)
//   O += P V           tcgen05.mma TS (A = P from TMEM, B = V MN-major SWIZZLE_128B, 16 keys = 2048 B per step), accumulator in
//                      TMEM columns [192,256); rescaled in TMEM only when a row maximum grew by > 2^8
//
// Warp roles: warps 0-3 softmax + epilogue (128 threads = 128 rows), warp 4 TMA producer (one lane),
// warp 5 MMA issuer (one elected lane).  Two CTAs per SM (256 TMEM columns, ~97 KB smem each) so that one
// CTA's softmax overlaps the other's MMAs.  No score, probability, bias or window tensor reaches HBM.
#include "common.cuh"
#include "tma.cuh"

namespace sodt {
namespace {

using namespace tc;

constexpr int TM = 128;            // query rows per CTA
constexpr int TN = 128;            // keys per tile
constexpr int HD = 64;             // head dim
#ifndef SODT_FLASH_TPR
#define SODT_FLASH_TPR 1
#endif
// threads per score row: 2 = warps w and w + 4 split the 128 keys of a tile (twice the softmax issue slots, one shared-memory
// exchange of the row maximum per tile).  Measured on B200 (B = 32, tools/prof_flash.py): TPR 1 0.771 ms, TPR 2 0.805 ms, with
// 3 of 8 exponentials on the FMA pipe 0.781 / 0.846 ms: the kernel is bound by the latency of the per-tile chain
// S -> TMEM load -> max -> exp -> P -> PV of the two resident CTAs, not by issue slots or the MUFU (48 % busy).
constexpr int TPR = SODT_FLASH_TPR;
constexpr int KPT = 128 / TPR;          // keys per thread and tile
constexpr int SM_WARPS = 4 * TPR, PRODUCER_WARP = SM_WARPS, MMA_WARP = SM_WARPS + 1;
constexpr int SM_THREADS = SM_WARPS * 32;
constexpr int NTHREADS = (SM_WARPS + 2) * 32;
constexpr int TILE_BYTES = TM * HD * 2;   // 16 KB
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units
#ifndef SODT_FLASH_POLY_EXP
#define SODT_FLASH_POLY_EXP 0
#endif
constexpr bool POLY_EXP = SODT_FLASH_POLY_EXP != 0;

struct SmemLayout {
    static constexpr int Q = 0;
    static constexpr int K = Q + TILE_BYTES;            // 2 stages
    static constexpr int V = K + 2 * TILE_BYTES;        // 2 stages
    static constexpr int XCH = V + 2 * TILE_BYTES;      // row maxima [2 parities][2 halves][128] + row sums [2][128] exchanged between the two threads of a row
    static constexpr int TAB = XCH + 6 * 128 * 4;       // (2*WS-1)^2 floats
};

// Transposes the bias table to [heads][(2ws-1)^2] and scales it by log2(e).
__global__ void prep_table_kernel(const float* __restrict__ table, float* __restrict__ out, int entries, int heads) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < entries * heads) {
        const int h = e / entries, i = e - h * entries;
        out[e] = table[(long long)i * heads + h] * LOG2E;
    }
}

template <int WS>
__global__ void __launch_bounds__(NTHREADS, 2)
window_attn_flash_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                         const float* __restrict__ table_t, int H, int W, int C, int heads, float scale) {
    constexpr int N = WS * WS;
    constexpr int T = N / TN;
    constexpr int SPAN = 2 * WS - 1;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar_k_full[2], bar_v_full[2], bar_kv_empty[2], bar_q_full, bar_s_full, bar_s_free, bar_p_full, bar_pv_done;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qtile = blockIdx.x, head = blockIdx.y;
    const int nww = W / WS, nW = (H / WS) * nww;
    const int b = blockIdx.z / nW, win = blockIdx.z - b * nW;
    const int wy = win / nww, wx = win - wy * nww;
    const int x0 = wx * WS, y0 = b * H + wy * WS;                            // the window's corner in the [B*H, W] token image
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;               // SWIZZLE_128B tiles: 1024-byte aligned
    float* tab = reinterpret_cast<float*>(smem + (sbase - smem_u32(smem)) + SmemLayout::TAB);

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&bar_k_full[s], 1); mbar_init(&bar_v_full[s], 1); mbar_init(&bar_kv_empty[s], 1); }
        mbar_init(&bar_q_full, 1);
        mbar_init(&bar_s_full, 1);
        mbar_init(&bar_s_free, SM_THREADS);
        mbar_init(&bar_p_full, SM_THREADS);
        mbar_init(&bar_pv_done, 1);
        fence_barrier_init();
    }
    if (warp == MMA_WARP) { tmem_alloc(&tmem_slot, 256); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm_S = tmem_slot, tm_P = tmem_slot + 128, tm_O = tmem_slot + 192;
    pdl_trigger();
    pdl_wait();                    // programmatic dependent launch (common.cuh): no global memory is touched above

    if (warp == PRODUCER_WARP) {
        // ===================================================================== producer (one lane, TMA)
        if (lane == 0) {
            constexpr int RPT = TN / WS;                        // window rows per 128-token tile
            tma::expect_tx(&bar_q_full, TILE_BYTES);
            tma::load_3d(sbase + SmemLayout::Q, &in_map, &bar_q_full, head * HD, x0, y0 + qtile * (TM / WS));
            for (int t = 0; t < T; ++t) {
                const int s = t & 1;
                if (t >= 2) mbar_wait(&bar_kv_empty[s], ((t >> 1) - 1) & 1);
                tma::expect_tx(&bar_k_full[s], TILE_BYTES);
                tma::load_3d(sbase + SmemLayout::K + s * TILE_BYTES, &in_map, &bar_k_full[s], C + head * HD, x0, y0 + t * RPT);
                tma::expect_tx(&bar_v_full[s], TILE_BYTES);
                tma::load_3d(sbase + SmemLayout::V + s * TILE_BYTES, &in_map, &bar_v_full[s], 2 * C + head * HD, x0, y0 + t * RPT);
            }
        }
    } else if (warp == MMA_WARP) {
        // =================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_s = idesc_bf16(TM, TN, false, false);
            constexpr uint32_t idesc_o = idesc_bf16(TM, HD, false, true);
            auto issue_s = [&](int t) {
                const int s = t & 1;
                mbar_wait(&bar_k_full[s], (t >> 1) & 1);
                if (t > 0) mbar_wait(&bar_s_free, (t - 1) & 1);
                fence_after_sync();
                const uint64_t a = tma::desc_sw128(sbase + SmemLayout::Q), bd = tma::desc_sw128(sbase + SmemLayout::K + s * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) mma_ss(tm_S, a + 2 * k, bd + 2 * k, idesc_s, k > 0);     // 16 channels = 32 B per step
                mma_commit(&bar_s_full);
            };
            mbar_wait(&bar_q_full, 0);
            issue_s(0);
            for (int t = 0; t < T; ++t) {
                const int s = t & 1;
                if (t + 1 < T) issue_s(t + 1);
                mbar_wait(&bar_v_full[s], (t >> 1) & 1);
                mbar_wait(&bar_p_full, t & 1);
                fence_after_sync();
                const uint64_t vd = tma::desc_sw128(sbase + SmemLayout::V + s * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < TN / 16; ++k)          // V tile rows = keys: 16 keys (2048 B) per step, MN-major
                    mma_ts(tm_O, tm_P + k * 8, vd + k * (2048 >> 4), idesc_o, (t > 0) || (k > 0));
                mma_commit(&bar_pv_done);
                mma_commit(&bar_kv_empty[s]);
            }
        }
    } else {
        // ============================================================ softmax + epilogue
        // TPR threads per query row: warp w handles TMEM lanes 32 (w % 4) .. and the keys [KPT (w / 4), KPT (w / 4 + 1)) of every tile;
        // the partial row maxima (per tile) and row sums (once) are exchanged through shared memory
        const int half = warp >> 2, row = (warp & 3) * 32 + lane;          // TMEM lane == row
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        float* xch = reinterpret_cast<float*>(smem + (sbase - smem_u32(smem)) + SmemLayout::XCH);
        for (int e = tid; e < SPAN * SPAN; e += SM_THREADS) tab[e] = table_t[(long long)head * SPAN * SPAN + e];
        asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");        // table visible to the softmax warps
        const int tq = qtile * TM + row;
        const int yq = tq / WS, xq = tq - yq * WS;
        const float* tab_q = tab + (yq + WS - 1) * SPAN + (xq + WS - 1);
        const float c = scale * LOG2E;
        float m_used = -INFINITY, l_run = 0.f;
        for (int t = 0; t < T; ++t) {
            mbar_wait(&bar_s_full, t & 1);
            fence_after_sync();
            // packed fp32x2 arithmetic: t = s * (scale log2 e) + bias, two scores per FFMA2; key j of this tile sits at window row
            // t*(TN/WS) + j/WS, column j%WS
            const float* tb = tab_q - (t * (TN / WS)) * SPAN;
            const uint64_t c2 = pack2(c, c);
            uint64_t t2[KPT / 2];
#pragma unroll
            for (int q4 = 0; q4 < KPT / 32; ++q4) {
                uint32_t r[32];
                tmem_ld32(tm_S + lane_addr + half * KPT + q4 * 32, r);
                tmem_wait_ld();
                if (q4 == KPT / 32 - 1) { fence_before_sync(); mbar_arrive(&bar_s_free); }
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const int k0 = half * KPT + q4 * 32 + j, kl = q4 * 32 + j;
                    const float b0 = tb[-((k0 / WS) * SPAN + (k0 % WS))], b1 = tb[-(((k0 + 1) / WS) * SPAN + ((k0 + 1) % WS))];
                    t2[kl >> 1] = ffma2(pack2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), c2, pack2(b0, b1));
                }
            }
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < KPT / 2; ++j) {
                float lo, hi;
                unpack2(t2[j], lo, hi);
                m4[j & 3] = fmax3(m4[j & 3], lo, hi);
            }
            float m_tile = fmaxf(fmax3(m4[0], m4[1], m4[2]), m4[3]);
            if (TPR == 2) {                                   // the row maximum of the tile over both halves (both threads take the same decisions)
                float* xm = xch + (t & 1) * 256;
                xm[half * 128 + row] = m_tile;
                asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");
                m_tile = fmaxf(m_tile, xm[(half ^ 1) * 128 + row]);
            }
            float alpha = 1.f;
            if (m_tile > m_used + RESCALE_THRESHOLD) {
                alpha = fast_exp2(m_used - m_tile);   // 0 on the first tile (m_used = -inf)
                m_used = m_tile;
            }
            const uint64_t nm2 = pack2(-m_used, -m_used);
            uint64_t sum2[2] = {0ull, 0ull};
            uint32_t pk[KPT / 2];
#pragma unroll
            for (int j = 0; j < KPT / 2; ++j) {
                const uint64_t x2 = fadd2(t2[j], nm2);
                uint64_t p2;
                if (POLY_EXP && ((j & 7) == 1 || (j & 7) == 4 || (j & 7) == 6)) {
                    p2 = exp2_poly2(x2);
                } else {
                    float lo, hi;
                    unpack2(x2, lo, hi);
                    p2 = pack2(fast_exp2(lo), fast_exp2(hi));
                }
                sum2[j & 1] = fadd2(sum2[j & 1], p2);
                float p0, p1;
                unpack2(p2, p0, p1);
                pk[j] = pack_bf16(p0, p1);
            }
            float sum;
            {
                float a, b;
                unpack2(fadd2(sum2[0], sum2[1]), a, b);
                sum = a + b;
            }
            l_run = l_run * alpha + sum;                       // this thread's keys only: the halves are added once at the end
            if (t > 0) {
                mbar_wait(&bar_pv_done, (t - 1) & 1);          // P and O are free again
                fence_after_sync();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {   // each thread rescales its HD / TPR accumulator columns
#pragma unroll
                    for (int h2 = 0; h2 < HD / TPR / 32; ++h2) {
                        uint32_t o[32];
                        tmem_ld32(tm_O + lane_addr + half * (HD / TPR) + h2 * 32, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
                        tmem_st32(tm_O + lane_addr + half * (HD / TPR) + h2 * 32, o);
                    }
                }
            }
            {
                uint32_t part[32];
#pragma unroll
                for (int h2 = 0; h2 < KPT / 64; ++h2) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) part[j] = pk[h2 * 32 + j];
                    tmem_st32(tm_P + lane_addr + half * (KPT / 2) + h2 * 32, part);
                }
            }
            tmem_wait_st();
            fence_before_sync();
            mbar_arrive(&bar_p_full);
        }
        mbar_wait(&bar_pv_done, (T - 1) & 1);
        fence_after_sync();
        if (TPR == 2) {
            float* xl = xch + 512;
            xl[half * 128 + row] = l_run;
            asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");
            l_run += xl[(half ^ 1) * 128 + row];
        }
        const float inv = 1.f / l_run;
        // O / rowsum -> bf16 -> the (now idle) Q tile as a SWIZZLE_128B staging tile -> one TMA store of the window rows
        const uint32_t stg = sbase + SmemLayout::Q + (uint32_t)row * 128;
#pragma unroll
        for (int h2 = 0; h2 < HD / TPR / 32; ++h2) {
            uint32_t o[32];
            const int col0 = half * (HD / TPR) + h2 * 32;
            tmem_ld32(tm_O + lane_addr + col0, o);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                const uint32_t chunk = (uint32_t)((col0 + j) >> 3) ^ (uint32_t)(row & 7);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (chunk << 4)),
                             "r"(pack_bf16(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv)),
                             "r"(pack_bf16(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv)),
                             "r"(pack_bf16(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv)),
                             "r"(pack_bf16(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv)) : "memory");
            }
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");
        if (tid == 0) {
            tma::store_3d(&out_map, sbase + SmemLayout::Q, head * HD, x0, y0 + qtile * (TM / WS));
            tma::store_commit();
            tma::store_wait_all();
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_slot, 256);
}

}  // namespace

size_t window_attn_flash_workspace(int heads, int ws) { return (size_t)heads * (2 * ws - 1) * (2 * ws - 1) * sizeof(float); }

bool window_attn_flash_supported(int H, int W, int C, int heads, int ws, int shift, int dtype) {
    return dtype == SODT_BF16 && shift == 0 && ws == 32 && C == heads * HD && H % ws == 0 && W % ws == 0 && C % 8 == 0;
}

int window_attn_flash_prepare(const float* table, void* workspace, int heads, int ws, cudaStream_t stream) {
    const int entries = (2 * ws - 1) * (2 * ws - 1);
    prep_table_kernel<<<(entries * heads + 255) / 256, 256, 0, stream>>>(table, static_cast<float*>(workspace), entries, heads);
    return check_launch();
}

int window_attn_flash(const void* qkv, const float* table, void* out, void* workspace, int B, int H, int W, int C,
                      int heads, int ws, float scale, bool prepared, cudaStream_t stream) {
    constexpr int WS = 32;
    const int entries = (2 * ws - 1) * (2 * ws - 1);
    float* table_t = static_cast<float*>(workspace);
    if (!prepared) {
        const int st = window_attn_flash_prepare(table, workspace, heads, ws, stream);
        if (st != SODT_OK) return st;
    }
    const size_t smem = SmemLayout::TAB + (size_t)entries * sizeof(float) + 1024;
    auto kern = window_attn_flash_kernel<WS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    const long long nwin = (long long)B * (H / ws) * (W / ws);
    if (nwin > 65535) return SODT_ERR_UNSUPPORTED;      // grid.z limit; the dispatcher routes larger batches to another kernel
    CUtensorMap in_map, out_map;
    {
        const int box[3] = {HD, WS, TN / WS};
        const long long din[3] = {3LL * C, W, (long long)B * H}, sin[2] = {3LL * C, 3LL * C * W};
        const long long dout[3] = {C, W, (long long)B * H}, sout[2] = {C, (long long)C * W};
        if (!tma::make_map_bf16(&in_map, qkv, 3, din, sin, box) ||
            !tma::make_map_bf16(&out_map, out, 3, dout, sout, box, CU_TENSOR_MAP_L2_PROMOTION_NONE)) return SODT_ERR_CUDA;
    }
    dim3 grid(ws * ws / TM, heads, (unsigned)nwin);
    e = launch_pdl(kern, grid, dim3(NTHREADS), smem, stream, true, in_map, out_map, table_t, H, W, C, heads, scale);
    if (e != cudaSuccess) return cuda_status(e);
    return check_launch();
}

}  // namespace sodt
