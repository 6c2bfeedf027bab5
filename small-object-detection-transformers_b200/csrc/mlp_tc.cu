// The linear MLP half of a Swin block as ONE kernel (second half):
//
//   out[M, C] = x + fc2(GELU(fc1(LayerNorm(x))))          reference backbone_vit.py:885-890 (Mlp.forward), :1128 (the block)
//
// The hidden activation [M, hidden] (3.2 GB per block at the benchmark geometry) never leaves the SM: it is produced into
// TMEM 128 hidden columns at a time, normalised / biased / GELU'd in registers, written back to TMEM as packed bf16 and
// consumed from there as the A operand of the second GEMM (tcgen05.mma with A in tensor memory).
//
// One CTA per SM walks 128-row tiles.  Per tile, for every 128-column chunk c of the hidden layer:
//   G1(c)   H[128 x 128]  = X[128 x C] . W1'[c]^T           SS MMAs, X tile resident in shared memory, W1' chunk streamed
//   E(c)    H <- GELU(rstd H - mean rstd colsum + b1')       LayerNorm folded as in linear_tc.cu; fp32 TMEM columns are
//                                                           overwritten in place by the packed 16-bit values: bf16, or
//                                                           (w2_fp16) fp16 values of 2 GELU against 0.5 W2 in fp16, which
//                                                           saves the conversions of the packed-half GELU arithmetic
//   G2(c)   O[128 x C]   += H[128 x 128] . W2[:, c]^T       TS MMAs (A = H from TMEM), W2 chunk streamed
// and once per tile  out = O + b2 + x  (the residual is the X tile itself: it is updated in place in shared memory and
// leaves through TMA stores), plus the partial row statistics the next block's norm1 needs.
//
//   warp 16     TMA producer: X tiles (two buffers), W1' k-blocks [128 x 64] and W2 k-blocks [C x 64] through two mbarrier rings
//   warp 17     MMA issuer, order G1(0), G1(1), G2(0), G1(2), G2(1), G1(3), ...: the epilogue of chunk c overlaps G2(c-1), G1(c+1)
//   warps 0-15  four epilogue groups, group g takes hidden columns [32 g, 32 g + 32) of every chunk (two H buffers in TMEM);
//               the C / 64 output boxes of a tile are dealt round-robin and handled right after the first hidden chunk of the
//               NEXT tile, so that the tensor pipe already has that tile's first chunks while the output tile drains
// TMEM: O at columns [0, C), H buffers at 256 and 384.  Both weight matrices stream from L2 for every tile (590 KB per
// 48 KB of X at C = 192): the kernel is bound by the SM's operand ingest, not by HBM or the tensor pipe.
#include <cuda.h>

#include "common.cuh"
#include "mlp_tc.h"
#include "tc05.cuh"
#include "tma.cuh"

namespace sodt {
namespace {

using namespace tc;

constexpr int BM = 128, HC = 128;
constexpr int EPI_GROUPS = 4, EPI_WARPS = EPI_GROUPS * 4;
constexpr int PRODUCER_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1;
constexpr int NTHREADS = (EPI_WARPS + 2) * 32;
constexpr int R1 = 3, R2 = 3;                       // ring depths (W1' / W2 k-blocks)
constexpr int BOX_BYTES = BM * 128;                 // 128 rows x 64 bf16

template <int C> struct Layout {
    static constexpr int KB = C / 64;               // k-blocks of G1 = 64-column boxes of the X / output tile
    static constexpr int A_BYTES = KB * BOX_BYTES;
    static constexpr int W1_SLOT = HC * 128;        // [128 hidden rows x 64 k]
    static constexpr int W2_SLOT = C * 128;         // [C rows x 64 hidden k]
    static constexpr int A_OFF = 0, W1_OFF = 2 * A_BYTES, W2_OFF = W1_OFF + R1 * W1_SLOT, TOTAL = W2_OFF + R2 * W2_SLOT;
    static constexpr int VEC_OFF = TOTAL;                 // colsum[hidden] then b1[hidden] (fp32), staged once per CTA
    static constexpr int O_COL = 0, H_COL = 256;
};

struct MlpParams {
    const float* b1;            // fp32 [hidden]: fc1 bias with the LayerNorm shift folded in
    const float* colsum;        // fp32 [hidden]: row sums of the folded bf16 W1'
    const float* b2;            // fp32 [C]
    const float* ln_stats;      // [M][2] (mean, rstd) or [ln_boxes][M][2] partial (sum, sum of squares)
    float* stats_out;           // [C / 64][M][2] or null
    int M, hidden, num_m_tiles, ln_boxes;
    float ln_inv_k, ln_eps;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src), "r"(c0), "r"(c1) : "memory");
}

// MLP_TRACE (experiments only): CTA 0 records clock64 at the hand-offs of its first 64 chunks into stats_out
#ifdef MLP_TRACE
#define TRACE(role, q, k) do { if (blockIdx.x == 0 && (q) < 64) reinterpret_cast<long long*>(p.stats_out)[((role) * 64 + (q)) * 8 + (k)] = clock64(); } while (0)
#else
#define TRACE(role, q, k) do { } while (0)
#endif

template <int C, bool F16>
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
              const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_o, const MlpParams p) {
    using L = Layout<C>;
    constexpr int KB = L::KB;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t a_full[2], a_empty[2], w1_full[R1], w1_empty[R1], w2_full[R2], w2_empty[R2], h_full[2], h_ready[2], o_full, o_empty;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int NCH = p.hidden / HC;
    const int n_iter = (int)blockIdx.x < p.num_m_tiles ? (p.num_m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total = n_iter * NCH;                  // hidden chunks this CTA walks, q = tile * NCH + chunk

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], KB); mbar_init(&h_full[s], 1); mbar_init(&h_ready[s], EPI_WARPS * 32); }
        for (int s = 0; s < R1; ++s) { mbar_init(&w1_full[s], 1); mbar_init(&w1_empty[s], 1); }
        for (int s = 0; s < R2; ++s) { mbar_init(&w2_full[s], 1); mbar_init(&w2_empty[s], 1); }
        mbar_init(&o_full, 1);
        mbar_init(&o_empty, KB * 128);
        fence_barrier_init();
    }
    if (warp == MMA_WARP) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_slot;
    pdl_trigger();
    pdl_wait();                                      // programmatic dependent launch: the prologue above overlaps the previous kernel's tail

    if (warp == PRODUCER_WARP) {
        if (lane == 0 && n_iter > 0) {
            int s1 = 0, r1 = 0, s2 = 0, r2 = 0;
            auto load_a = [&](int t) {
                const int ab = t & 1;
                if (t >= 2) mbar_wait(&a_empty[ab], (uint32_t)(((t >> 1) - 1) & 1));
                tma::expect_tx(&a_full[ab], L::A_BYTES);
                const int row0 = ((int)blockIdx.x + t * (int)gridDim.x) * BM;
                for (int kb = 0; kb < KB; ++kb) tma_load_2d(sbase + L::A_OFF + ab * L::A_BYTES + kb * BOX_BYTES, &tmap_x, &a_full[ab], kb * 64, row0);
            };
            auto load_w1 = [&](int c) {
                for (int kb = 0; kb < KB; ++kb) {
                    if (r1 > 0) mbar_wait(&w1_empty[s1], (uint32_t)((r1 - 1) & 1));
#ifdef MLP_DBG_NOLOAD
                    if (r1 > 0) { mbar_arrive(&w1_full[s1]); } else
#endif
                    { tma::expect_tx(&w1_full[s1], L::W1_SLOT);
                    tma_load_2d(sbase + L::W1_OFF + s1 * L::W1_SLOT, &tmap_w1, &w1_full[s1], kb * 64, c * HC); }
                    if (++s1 == R1) { s1 = 0; ++r1; }
                }
            };
            auto load_w2 = [&](int c) {
                for (int bx = 0; bx < 2; ++bx) {
                    if (r2 > 0) mbar_wait(&w2_empty[s2], (uint32_t)((r2 - 1) & 1));
#ifdef MLP_DBG_NOLOAD
                    if (r2 > 0) { mbar_arrive(&w2_full[s2]); } else
#endif
                    { tma::expect_tx(&w2_full[s2], L::W2_SLOT);
                    tma_load_2d(sbase + L::W2_OFF + s2 * L::W2_SLOT, &tmap_w2, &w2_full[s2], c * HC + bx * 64, 0); }
                    if (++s2 == R2) { s2 = 0; ++r2; }
                }
            };
            const int c_pref = NCH / 2 - 1;                         // chunk at which the next X tile is requested (its buffer: tile t-1's,
            load_a(0);                                           // free once that tile's output left, early in tile t)
            load_w1(0);
            if (total > 1) load_w1(NCH > 1 ? 1 : 0);
            int t = 0, c = 0;
            for (int q = 0; q < total; ++q) {                    // same order as the MMA issuer consumes
                if (c == c_pref && t + 1 < n_iter) load_a(t + 1);
                load_w2(c);
                if (q + 2 < total) load_w1((c + 2) % NCH);
                if (++c == NCH) { c = 0; ++t; }
            }
        }
    } else if (warp == MMA_WARP) {
        if (lane == 0 && n_iter > 0) {
            constexpr uint32_t id1 = idesc_bf16(BM, HC, false, false);
            constexpr uint32_t id2 = F16 ? idesc_f16(BM, C, false, false) : idesc_bf16(BM, C, false, false);
            int s1 = 0, r1 = 0, s2 = 0, r2 = 0;
            auto g1 = [&](int q) {
                const int t = q / NCH, c = q - t * NCH, ab = t & 1, hb = q & 1;
                if (c == 0) mbar_wait(&a_full[ab], (uint32_t)((t >> 1) & 1));
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&w1_full[s1], (uint32_t)(r1 & 1));
                    fence_after_sync();
                    const uint64_t da = tma::desc_sw128(sbase + L::A_OFF + ab * L::A_BYTES + kb * BOX_BYTES);
                    const uint64_t db = tma::desc_sw128(sbase + L::W1_OFF + s1 * L::W1_SLOT);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) mma_ss(tm + L::H_COL + hb * HC, da + 2 * ks, db + 2 * ks, id1, (kb | ks) != 0);
                    mma_commit(&w1_empty[s1]);
                    if (++s1 == R1) { s1 = 0; ++r1; }
                }
                mma_commit(&h_full[hb]);
            };
            auto g2 = [&](int q) {
                const int t = q / NCH, c = q - t * NCH, hb = q & 1;
                TRACE(0, q, 0);
                mbar_wait(&h_ready[hb], (uint32_t)((q >> 1) & 1));
                TRACE(0, q, 1);
                if (c == 0 && t > 0) mbar_wait(&o_empty, (uint32_t)((t - 1) & 1));
                TRACE(0, q, 2);
                for (int bx = 0; bx < 2; ++bx) {
                    mbar_wait(&w2_full[s2], (uint32_t)(r2 & 1));
                    fence_after_sync();
                    const uint64_t db = tma::desc_sw128(sbase + L::W2_OFF + s2 * L::W2_SLOT);
                    // epilogue group e packed hidden columns [32 e, 32 e + 32) of the chunk into TMEM columns [32 e, 32 e + 16)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        mma_ts(tm + L::O_COL, tm + L::H_COL + hb * HC + 32 * (2 * bx + (ks >> 1)) + 8 * (ks & 1), db + 2 * ks, id2, (c | bx | ks) != 0);
                    mma_commit(&w2_empty[s2]);
                    if (++s2 == R2) { s2 = 0; ++r2; }
                }
                if (c == NCH - 1) mma_commit(&o_full);
                TRACE(0, q, 3);
            };
            g1(0);
            if (total > 1) g1(1);
            for (int q = 0; q < total; ++q) {           // G2(q) frees H buffer q % 2 for G1(q + 2): the tensor pipe executes in issue order
                g2(q);
                if (q + 2 < total) g1(q + 2);
                TRACE(0, q, 4);
            }
        }
    } else {
        const int quarter = warp & 3, g = warp >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const int sw = row & 7;
        const bool issuer = quarter == 0 && lane == 0;
        int pending_ab = -1;                                     // X buffer whose TMA store this thread still has to see read
        // per-column LayerNorm / bias terms of fc1 from shared memory: global (L1) loads in the chunk loop queue behind the
        // tensor core's operand reads
        float* vec = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)) + L::VEC_OFF);
        for (int i = tid; i < p.hidden; i += EPI_WARPS * 32) { vec[i] = p.colsum[i]; vec[p.hidden + i] = p.b1[i]; }
        for (int i = tid; i < C; i += EPI_WARPS * 32) vec[2 * p.hidden + i] = p.b2[i];      // no L1 is left beside 227 KB of shared memory: a __ldg is an L2 round trip
        asm volatile("bar.sync 5, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        const uint32_t vec_s = sbase + L::VEC_OFF;
        auto load_mr = [&](int t) {                              // (mean, rstd) of this thread's row of tile t
            float2 r = make_float2(0.f, 1.f);
            if (t < n_iter) {
                const int gr = ((int)blockIdx.x + t * (int)gridDim.x) * BM + row;
                if (gr < p.M) {
                    const float2* s = reinterpret_cast<const float2*>(p.ln_stats) + gr;
                    if (p.ln_boxes == 0) {
                        r = __ldg(s);
                    } else {
                        const float2 p0 = __ldg(s);
                        const float2 p1 = p.ln_boxes > 1 ? __ldg(s + p.M) : make_float2(0.f, 0.f);
                        const float2 p2 = p.ln_boxes > 2 ? __ldg(s + 2 * (size_t)p.M) : make_float2(0.f, 0.f);
                        const float mean = (p0.x + p1.x + p2.x) * p.ln_inv_k;
                        r = make_float2(mean, rsqrtf(fmaxf(fmaf(-mean, mean, (p0.y + p1.y + p2.y) * p.ln_inv_k), 0.f) + p.ln_eps));
                    }
                }
            }
            return r;
        };
        // 32 fp32 columns [32 g, 32 g + 32) of H buffer hb -> LayerNorm terms, bias, GELU -> 16 packed bf16 columns in place
        auto h_box = [&](int hb, int c, float2 mr, int q) {
            const int col0 = c * HC + g * 32;
            const uint32_t tsrc = tm + lane_addr + L::H_COL + hb * HC + g * 32;
            const float rstd = mr.y, nmr = -mr.x * mr.y;
            const uint64_t rstd2 = pack2(rstd, rstd), nmr2 = pack2(nmr, nmr);
            uint32_t ra[16], rb[16];
            tmem_ld16(tsrc, ra);
            tmem_ld16(tsrc + 16, rb);
            tmem_wait_ld();
            if (lane == 0 && (warp == 0 || warp == 15)) TRACE(1 + (warp == 15), q, 6);
#pragma unroll
            for (int j = 0; j < 32; j += 16) {
                uint32_t pk[8];
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                    float4 cs, bb;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cs.x), "=f"(cs.y), "=f"(cs.z), "=f"(cs.w) : "r"(vec_s + 4 * (col0 + j + e)));
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w) : "r"(vec_s + 4 * (p.hidden + col0 + j + e)));
                    const uint32_t* a = j ? rb : ra;
                    const uint64_t x01 = ffma2(rstd2, pack2(__uint_as_float(a[e]), __uint_as_float(a[e + 1])), ffma2(nmr2, pack2(cs.x, cs.y), pack2(bb.x, bb.y)));
                    const uint64_t x23 = ffma2(rstd2, pack2(__uint_as_float(a[e + 2]), __uint_as_float(a[e + 3])), ffma2(nmr2, pack2(cs.z, cs.w), pack2(bb.z, bb.w)));
                    if (F16) {                                    // fp16 hidden operand holding 2 GELU (w2 carries the 0.5)
                        float v0, v1, v2, v3;
                        unpack2(x01, v0, v1);
                        unpack2(x23, v2, v3);
                        pk[e >> 1] = gelu2x_f16x2(v0, v1);
                        pk[(e >> 1) + 1] = gelu2x_f16x2(v2, v3);
                    } else {
#ifdef MLP_GELU_F16
                        float v0, v1, v2, v3;
                        unpack2(x01, v0, v1);
                        unpack2(x23, v2, v3);
                        gelu_fast2(v0, v1);
                        gelu_fast2(v2, v3);
                        pk[e >> 1] = pack_bf16(v0, v1);
                        pk[(e >> 1) + 1] = pack_bf16(v2, v3);
#else
                        pk[e >> 1] = gelu_bf16x2_f32x2(x01);      // packed fp32x2 GELU: fewer instructions than the packed-fp16 form + conversions
                        pk[(e >> 1) + 1] = gelu_bf16x2_f32x2(x23);
#endif
                    }
                }
                tmem_st8(tsrc + (j >> 1), pk);
            }
            if (lane == 0 && (warp == 0 || warp == 15)) TRACE(1 + (warp == 15), q, 7);
            tmem_wait_st();
            if (lane == 0 && (warp == 0 || warp == 15)) TRACE(1 + (warp == 15), q, 3);
            fence_before_sync();
            mbar_arrive(&h_ready[hb]);
        };
        // output box j of tile t: O + b2 + residual (the X tile, updated in place) -> TMA store; partial row statistics
        auto o_box = [&](int t, int j) {
            const int ab = t & 1, mt = (int)blockIdx.x + t * (int)gridDim.x, grow = mt * BM + row;
            const uint32_t box = sbase + L::A_OFF + ab * L::A_BYTES + j * BOX_BYTES, my_row = box + row * 128;
            const uint32_t tsrc = tm + lane_addr + L::O_COL + j * 64;
            uint32_t ra[16], rb[16];
            float so = 0.f, sso = 0.f;
            tmem_ld16(tsrc, ra);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 64; k += 16) {
                if (k + 16 < 64) { if (k & 16) tmem_ld16(tsrc + k + 16, ra); else tmem_ld16(tsrc + k + 16, rb); }
#pragma unroll
                for (int e = 0; e < 16; e += 8) {
                    const uint32_t* a = (k & 16) ? rb : ra;
                    float4 b0, b1;
                    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0.x), "=f"(b0.y), "=f"(b0.z), "=f"(b0.w) : "r"(vec_s + 4 * (2 * p.hidden + j * 64 + k + e)));
                    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b1.x), "=f"(b1.y), "=f"(b1.z), "=f"(b1.w) : "r"(vec_s + 4 * (2 * p.hidden + j * 64 + k + e + 4)));
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    const uint32_t dst = my_row + ((((k + e) >> 3) ^ sw) << 4);
                    uint32_t q0, q1, q2, q3;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3) : "r"(dst) : "memory");
                    const uint32_t qq[4] = {q0, q1, q2, q3};
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&qq[i]);
                        v[2 * i] = __uint_as_float(a[e + 2 * i]) + bb[2 * i] + __low2float(h);
                        v[2 * i + 1] = __uint_as_float(a[e + 2 * i + 1]) + bb[2 * i + 1] + __high2float(h);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { so += v[i]; sso = fmaf(v[i], v[i], sso); }
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16(v[0], v[1])), "r"(pack_bf16(v[2], v[3])),
                                 "r"(pack_bf16(v[4], v[5])), "r"(pack_bf16(v[6], v[7])) : "memory");
                }
                if (k + 16 < 64) tmem_wait_ld();
            }
            fence_before_sync();
            mbar_arrive(&o_empty);                               // this thread's accumulator columns are in registers / stored
#ifndef MLP_TRACE
            if (p.stats_out != nullptr && grow < p.M)
                reinterpret_cast<float2*>(p.stats_out)[(size_t)j * p.M + grow] = make_float2(so, sso);
#endif
            fence_proxy_async();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
            if (issuer) {
                tma_store_2d(&tmap_o, box, j * 64, mt * BM);
                tma::store_commit();
                pending_ab = ab;
            }
        };
        auto o_tile = [&](int t) {
            bool waited = false;
#pragma unroll 1
            for (int j = 0; j < KB; ++j) {
                if (((j + t) & (EPI_GROUPS - 1)) != g) continue;
                if (!waited) { mbar_wait(&o_full, (uint32_t)(t & 1)); fence_after_sync(); waited = true; }
                if (issuer && pending_ab >= 0) { tma::store_wait_read<0>(); mbar_arrive(&a_empty[pending_ab]); pending_ab = -1; }
                o_box(t, j);
            }
        };
        float2 mr = make_float2(0.f, 1.f), mr_next = load_mr(0);
        int t = 0, c = 0;
#pragma unroll 1
        for (int q = 0; q < total; ++q) {
            const int hb = q & 1;
            if (c == 0) { mr = mr_next; mr_next = load_mr(t + 1); }
            if (issuer && pending_ab >= 0) { tma::store_wait_read<0>(); mbar_arrive(&a_empty[pending_ab]); pending_ab = -1; }
            if (lane == 0 && (warp == 0 || warp == 15)) TRACE(1 + (warp == 15), q, 0);
            mbar_wait(&h_full[hb], (uint32_t)((q >> 1) & 1));
            fence_after_sync();
            if (lane == 0 && (warp == 0 || warp == 15)) TRACE(1 + (warp == 15), q, 1);
            h_box(hb, c, mr, q);
            if (lane == 0 && (warp == 0 || warp == 15)) TRACE(1 + (warp == 15), q, 4);
            if (c == 0 && t > 0) o_tile(t - 1);
            if (lane == 0 && (warp == 0 || warp == 15)) TRACE(1 + (warp == 15), q, 5);
            if (++c == NCH) { c = 0; ++t; }
        }
        if (n_iter > 0) o_tile(n_iter - 1);
        if (issuer) tma::store_wait_all();
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_slot, 512);
}

bool map_2d(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows, bool is_output) {
    const long long dims[2] = {cols, rows}, strides[1] = {ld};
    const int box[2] = {64, box_rows};
    return tma::make_map_bf16(m, base, 2, dims, strides, box, is_output ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

template <int C, bool F16>
int launch(const MlpTcArgs& g, int num_sms, cudaStream_t stream) {
    using L = Layout<C>;
    CUtensorMap mx, mw1, mw2, mo;
    if (!map_2d(&mx, g.x, g.M, C, g.ldx, BM, false) || !map_2d(&mw1, g.w1, g.hidden, C, C, HC, false) ||
        !map_2d(&mw2, g.w2, C, g.hidden, g.hidden, C, false) || !map_2d(&mo, g.out, g.M, C, g.ldo, BM, true))
        return SODT_ERR_CUDA;
    MlpParams p{};
    p.b1 = g.b1; p.colsum = g.ln_colsum; p.b2 = g.b2; p.ln_stats = g.ln_stats; p.stats_out = g.stats_out;
    p.M = g.M; p.hidden = g.hidden; p.num_m_tiles = (g.M + BM - 1) / BM; p.ln_boxes = g.ln_boxes;
    p.ln_inv_k = 1.f / (float)C; p.ln_eps = g.ln_eps;
    const size_t smem = (size_t)L::TOTAL + 1024 + (size_t)g.hidden * 8 + C * 4;
    if (smem > 227 * 1024) return SODT_ERR_UNSUPPORTED;
    auto kern = mlp_tc_kernel<C, F16>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    const int grid = p.num_m_tiles < num_sms ? p.num_m_tiles : num_sms;
    e = launch_pdl(kern, dim3(grid), dim3(NTHREADS), smem, stream, true, mx, mw1, mw2, mo, p);
    if (e != cudaSuccess) return cuda_status(e);
    return check_launch();
}

}  // namespace

bool mlp_tc_supported(int M, int C, int hidden) {
    if (!(M > 0 && (C == 64 || C == 128 || C == 192) && hidden >= 2 * HC && hidden % HC == 0 && (long long)M + BM < 2147483647LL)) return false;
    const long long smem = 2LL * (C / 64) * BOX_BYTES + R1 * HC * 128 + R2 * C * 128 + 1024 + 8LL * hidden + 4 * C;      // Layout<C>::TOTAL + staged vectors
    return smem <= 227 * 1024;
}

int mlp_tc(const MlpTcArgs& g, int num_sms, cudaStream_t stream) {
    if (!mlp_tc_supported(g.M, g.C, g.hidden)) return SODT_ERR_UNSUPPORTED;
    if (!g.x || !g.w1 || !g.w2 || !g.out || !g.b1 || !g.b2 || !g.ln_stats || !g.ln_colsum || g.ln_boxes < 0 || g.ln_boxes > 3 ||
        g.ldx % 8 || g.ldo % 8 || g.ldx < g.C || g.ldo < g.C)
        return SODT_ERR_INVALID_ARG;
    switch (g.C) {
        case 64: return g.w2_fp16 ? launch<64, true>(g, num_sms, stream) : launch<64, false>(g, num_sms, stream);
        case 128: return g.w2_fp16 ? launch<128, true>(g, num_sms, stream) : launch<128, false>(g, num_sms, stream);
        default: return g.w2_fp16 ? launch<192, true>(g, num_sms, stream) : launch<192, false>(g, num_sms, stream);
    }
}

}  // namespace sodt
