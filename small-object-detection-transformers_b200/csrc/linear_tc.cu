// bf16 GEMM-shaped layers of the detector on tcgen05 tensor cores with TMA-addressed operands and fused epilogues
// (SURVEY.md section 8f, rank 1-2):
//
//   out[M, N] = act(A[M, K] . W[N, K]^T + bias[N]) (+ residual[M, N])          act = identity | GELU | SiLU
//
// The A operand is never materialised when it is only a re-indexing of an activation tensor; the TMA producer addresses
//   plain   A = x[M, K] (row stride ldx), optionally continued by a second tensor x2 for columns >= k_split
//           (neck over the concatenation of two stage outputs, reference backbone_vit.py:239-262, without the concat)
//   conv    A[(b,y,x), (ky,kx,c)] = in[b, y+ky-pad_t, x+kx-pad_l, c]  -- stride-1 kh x kw convolution on an NHWC tensor as
//           a "tap GEMM"; zero padding comes from the TMA out-of-bounds fill (conv-MLP 2x2 conv, reference
//           backbone_vit.py:881-899; the head's 1x1 / 3x3 Conv blocks, common.py:38-52)
//   merge   A[(b,i,j), (dx,dy,c)] = in[b, 2i+dy, 2j+dx, c]  -- the 2x2 neighbourhood gather of PatchMerging
//           (reference backbone_vit.py:840-851) as a rank-5 tensor map
// and replaces nn.Linear + the separate GELU pass of Mlp (backbone_vit.py:885-890,968,990), the separate residual add of
// SwinTransformerBlock (:1125,:1128), F.pad + Conv2d + GELU of the conv-MLP, and conv + BatchNorm(folded) + SiLU of the head.
// These GEMMs are HBM-bound (K = 64..3072 on millions of rows), so removing the extra passes is worth more than MMA efficiency.
//
// Persistent CTAs (one per SM), tile 128 x BN (BN = 256 / 192 / 128 / 64), k-block 64:
//   warp 16     TMA producer: cp.async.bulk.tensor boxes of A and W, SWIZZLE_128B, mbarrier ring of STAGES stages
//   warp 17     MMA issuer:   tcgen05.mma kind::f16, M=128, N=BN, K-major SWIZZLE_128B operands, fp32 accumulators in TMEM,
//                             NACC accumulator buffers so the epilogue of tile t overlaps the MMAs of the next tiles
//   warps 0-15  epilogue:     four groups of 4 warps take the 64-column boxes of the tiles round-robin (residual box
//                             TMA-loaded into the group's staging buffer first): tcgen05.ld (warp w: TMEM lanes
//                             32*(w%4)..), bias, activation, residual, bf16 pack into a SWIZZLE_128B staging box in shared
//                             memory, one TMA store per box (full 128-byte lines; per-thread 16-byte row stores capped
//                             the kernel at ~1.9 TB/s)
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "linear_tc.h"
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

enum { MODE_PLAIN = 0, MODE_CONV = 1, MODE_MERGE = 2, MODE_UPCAT = 3, MODE_HALO = 4 };
// MODE_HALO: the conv taps of one kernel row read ONE halo tile.  A 128-pixel tile of an image row needs the pixels
// [x0 - pad_l, x0 - pad_l + 128 + kw - 1) of input row y + ky - pad_t for its kw taps (ky, 0..kw-1): they are loaded once as a
// (64 channels, 128 + kw - 1) box and tap kx is the A operand that STARTS kx ROWS (kx * 128 B) into the SWIZZLE_128B tile -- the
// tensor core applies the swizzle to absolute shared-memory address bits, so a start address that is not a multiple of the
// 1024-byte atom is exact with the descriptor's base-offset field left 0 (tests/probes/rowshift_probe.cu, all shifts 0-9 on a
// B200).  A k-stage is one (ky, 64-channel block) with kw k-blocks of W: the conv GEMMs run at the SM's operand ingest rate, and
// this halves (2x2) / thirds (3x3) the A bytes a tile pulls from L2.
constexpr int HALO_BYTES = 17 * 1024;       // (128 + kw - 1 <= 136) rows x 128 B, rounded to the swizzle atom

struct Epilogue {
    const float* bias;
    const float* ln_stats;      // [M][2] (mean, rstd) of the A rows or null
    const float* ln_colsum;
    float* stats_out;
    int M, has_residual, act, res_tiles;
    int ln_boxes;               // > 0: ln_stats holds partial (sum, sum of squares) pairs, [ln_boxes][M][2]
    float ln_inv_k, ln_eps;
};

struct Addressing {
    int mode;
    int k_split;     // plain: k-blocks read from x before switching to x2 (== K/64 when there is no x2)
    int cpb;         // conv / merge: 64-channel blocks per tap / per 2x2 position
    int kw, pad_t, pad_l;
    int HW, W;       // conv: pixels per image and row width;  merge: W = output row width
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// CTA-pair forms: the transaction bytes complete on the LEADER's mbarrier (a shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const void* tmap, uint32_t bar_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(bar_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const void* tmap, uint32_t bar_addr, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const void* tmap, uint32_t bar_addr, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
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

template <int BN> struct Cfg {
    static constexpr int NACC = BN <= 128 ? 4 : 2;                                    // TMEM accumulator buffers (NACC * BN <= 512)
};

// Epilogue variants compiled as separate kernels: with every feature a run-time flag, the unrolled epilogue issued ~1500
// instructions per 64-column box (predicated-off bias / LayerNorm / residual / statistics code still takes issue slots and
// instruction-cache space; ncu: 57 % issue utilisation and 20 % instruction-fetch stalls on a bias-free GEMM).
// bits: 1 bias, 2 LayerNorm fold, 4 residual, 8 row statistics out; act in bits 4-5; EPI_ANY keeps everything run-time.
constexpr int EPI_BIAS = 1, EPI_LN = 2, EPI_RES = 4, EPI_STATS = 8, EPI_ACT_SHIFT = 4, EPI_ANY = 1 << 8;
constexpr int epi_code(bool bias, bool ln, bool res, bool stats, int act) {
    return (bias ? EPI_BIAS : 0) | (ln ? EPI_LN : 0) | (res ? EPI_RES : 0) | (stats ? EPI_STATS : 0) | (act << EPI_ACT_SHIFT);
}

// WRES ("weights resident"): the CTA keeps ONE N-tile for its whole life, loads that tile's W once and streams only A.
// Without it every 128-row tile re-fetches its W tile from L2 (72-96 KB for 48 KB of A at K = 192), and the wide GEMMs
// (qkv, fc1) ran at the L2 -> SM rate (~9 TB/s measured), not at the HBM rate.
constexpr int MAX_STAGES = 8;

// LIN_TRACE (experiments only, tools/trace_linear.py): CTA 0 records clock64 at the hand-offs of its first 64 tiles
#ifdef LIN_TRACE
__device__ long long g_lin_trace[6][64][16];
#define LTRACE(role, tile, ev) do { if (blockIdx.x == 0 && (tile) < 64) g_lin_trace[role][tile][ev] = clock64(); } while (0)
#else
#define LTRACE(role, tile, ev) do { } while (0)
#endif

// CTA2 ("CTA pair", cta_group::2): two CTAs of a cluster share one 256-row tile: each loads its own 128 rows of A and HALF of
// the W tile, the leader issues M = 256 MMAs over both shared memories, each CTA runs the epilogue of its own 128 rows.
// Per SM the operand bytes per output drop by a third at BN = 256 (the K >= 384 GEMMs run at the SM's operand ingest rate).
template <int BN, int EPI, bool WRES, bool CTA2>
__global__ void __launch_bounds__(NTHREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_x2,
                 const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_o,
                 const __grid_constant__ CUtensorMap tmap_r, const Epilogue ep,
                 const Addressing ad, int K, int num_n_tiles, int num_m_tiles, int stages) {
    static_assert(!(WRES && CTA2), "the resident-W walk is a single-CTA schedule");
    constexpr int NACC = Cfg<BN>::NACC;
    constexpr int BNL = CTA2 ? BN / 2 : BN;                      // W rows this CTA loads
    constexpr int B_BYTES = BNL * BK * 2;
    const int crank = CTA2 ? (int)cluster_ctarank() : 0;
    const int cid = CTA2 ? (int)blockIdx.x >> 1 : (int)blockIdx.x, ncl = CTA2 ? (int)gridDim.x >> 1 : (int)gridDim.x;   // tile walker id / count
    constexpr int MT_ROWS = CTA2 ? 2 : 1;                        // 128-row tiles per M-tile of the walk
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t full[MAX_STAGES], empty[MAX_STAGES], acc_full[NACC], acc_empty[NACC], res_full[EPI_GROUPS], w_full;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int nkb = K / BK;
    const bool halo = ad.mode == MODE_HALO;
    const int sub = halo ? ad.kw : 1;                            // k-blocks per ring stage
    const int nst = nkb / sub;                                   // ring stages per tile
    const uint32_t a_stage = halo ? HALO_BYTES : A_BYTES;
    // shared memory: [resident W: nkb k-blocks][ring of A (+ B) stages][4 staging boxes]
    const uint32_t ring_base = sbase + (WRES ? nkb * B_BYTES : 0);
    const uint32_t stage_bytes = a_stage + (WRES ? 0 : sub * B_BYTES);
    const uint32_t staging_base = ring_base + stages * stage_bytes;
    // tile walk: WRES: N-tile nt0 is fixed, M-tiles mt0, mt0 + mstep, ...; otherwise tiles blockIdx.x, + gridDim.x, ... in (mt, nt) order
    const int nt0 = WRES ? cid % num_n_tiles : 0;
    const int mt0 = WRES ? cid / num_n_tiles : 0, mstep = WRES ? ncl / num_n_tiles : 1;
    const int num_tiles = num_n_tiles * num_m_tiles;
    const int n_iter = WRES ? (mt0 < num_m_tiles ? (num_m_tiles - mt0 + mstep - 1) / mstep : 0)
                            : (cid < num_tiles ? (num_tiles - cid + ncl - 1) / ncl : 0);
    auto tile_at = [&](int i, int& mt, int& nt) {
        if (WRES) { mt = mt0 + i * mstep; nt = nt0; }
        else { const int tile = cid + i * ncl; mt = tile / num_n_tiles; nt = tile - mt * num_n_tiles; }
    };

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < NACC; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], EPI_WARPS * MT_ROWS); }   // one arrival per epilogue warp (pair: of both CTAs)
        for (int g = 0; g < EPI_GROUPS; ++g) mbar_init(&res_full[g], 1);
        mbar_init(&w_full, 1);
        fence_barrier_init();
    }
    if (warp == MMA_WARP) {
        if (CTA2) { tmem_alloc2(&tmem_slot, 512); tmem_relinquish2(); }
        else { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    }
    fence_before_sync();
    if (CTA2) cluster_sync_all();                                // the peer's barriers are initialised before anything signals them
    else __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_slot;
    pdl_trigger();
    pdl_wait();                                                  // programmatic dependent launch: no global memory is touched above

    if (warp == PRODUCER_WARP) {
        if (lane == 0 && n_iter > 0) {
            if (WRES) {                                          // this CTA's W tile, once
                mbar_expect_tx(&w_full, nkb * B_BYTES);
                for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sbase + kb * B_BYTES, &tmap_w, &w_full, kb * BK, nt0 * BN);
            }
            int stage = 0, round = 0;
            for (int it = 0; it < n_iter; ++it) {
                int mt, nt;
                tile_at(it, mt, nt);
                const int row0 = (mt * MT_ROWS + crank) * BM;
                int p0 = 0, p1 = 0, p2 = 0;                      // conv: x0, y0, b;  merge: j0, bi0
                if (ad.mode == MODE_CONV || ad.mode == MODE_UPCAT || halo) {
                    p2 = row0 / ad.HW;
                    const int rem = row0 - p2 * ad.HW;
                    p1 = rem / ad.W;
                    p0 = rem - p1 * ad.W;
                } else if (ad.mode == MODE_MERGE) {
                    p1 = row0 / ad.W;
                    p0 = row0 - p1 * ad.W;
                }
                int tap = 0, cc = 0;                             // running (tap, channel block) of the k loop
                if (halo) {                                      // stage = (kernel row ky, channel block cc): one halo box + kw blocks of W
                    const uint32_t tx = (uint32_t)(BM + ad.kw - 1) * 128u + (WRES ? 0u : (uint32_t)(ad.kw * B_BYTES));
                    int ky = 0;
                    for (int ks = 0; ks < nst; ++ks) {
                        if (round > 0) mbar_wait(&empty[stage], (uint32_t)((round - 1) & 1));
                        const uint32_t sa = ring_base + stage * stage_bytes;
                        if (CTA2) {
                            if (crank == 0) mbar_expect_tx(&full[stage], 2 * tx);
                            const uint32_t fb = mapa_u32(smem_u32(&full[stage]), 0);
                            tma_load_4d_pair(sa, &tmap_x, fb, cc * BK, p0 - ad.pad_l, p1 + ky - ad.pad_t, p2);
                            for (int kx = 0; kx < ad.kw; ++kx)
                                tma_load_2d_pair(sa + HALO_BYTES + kx * B_BYTES, &tmap_w, fb, ((ky * ad.kw + kx) * ad.cpb + cc) * BK, nt * BN + crank * BNL);
                        } else {
                            mbar_expect_tx(&full[stage], tx);
                            tma_load_4d(sa, &tmap_x, &full[stage], cc * BK, p0 - ad.pad_l, p1 + ky - ad.pad_t, p2);
                            if (!WRES)
                                for (int kx = 0; kx < ad.kw; ++kx)
                                    tma_load_2d(sa + HALO_BYTES + kx * B_BYTES, &tmap_w, &full[stage], ((ky * ad.kw + kx) * ad.cpb + cc) * BK, nt * BN);
                        }
                        if (++cc == ad.cpb) { cc = 0; ++ky; }
                        if (++stage == stages) { stage = 0; ++round; }
                    }
                    continue;
                }
                for (int kb = 0; kb < nkb; ++kb) {
                    if (round > 0) mbar_wait(&empty[stage], (uint32_t)((round - 1) & 1));
                    if (kb < 16) LTRACE(1, it, kb);
                    const uint32_t sa = ring_base + stage * stage_bytes;
                    if (CTA2) {
                        // both CTAs' boxes complete on the leader's barrier, armed by the leader for the bytes of the pair
                        if (crank == 0) mbar_expect_tx(&full[stage], 2 * stage_bytes);
                        const uint32_t fb = mapa_u32(smem_u32(&full[stage]), 0);
                        if (ad.mode == MODE_PLAIN) {
                            if (kb < ad.k_split) tma_load_2d_pair(sa, &tmap_x, fb, kb * BK, row0);
                            else tma_load_2d_pair(sa, &tmap_x2, fb, (kb - ad.k_split) * BK, row0);
                        } else if (ad.mode == MODE_CONV) {
                            const int ky = tap / ad.kw, kx = tap - ky * ad.kw;
                            tma_load_4d_pair(sa, &tmap_x, fb, cc * BK, p0 + kx - ad.pad_l, p1 + ky - ad.pad_t, p2);
                        } else if (ad.mode == MODE_UPCAT) {
                            if (kb < ad.k_split) tma_load_5d_pair(sa, &tmap_x, fb, kb * BK, 0, p0 >> 1, 0, p2 * (ad.pad_t >> 1) + (p1 >> 1));
                            else tma_load_4d_pair(sa, &tmap_x2, fb, (kb - ad.k_split) * BK, p0, p1, p2);
                        } else {
                            tma_load_5d_pair(sa, &tmap_x, fb, cc * BK, tap >> 1, p0, tap & 1, p1);
                        }
                        if (++cc == ad.cpb) { cc = 0; ++tap; }
                        tma_load_2d_pair(sa + A_BYTES, &tmap_w, fb, kb * BK, nt * BN + crank * BNL);
                        if (++stage == stages) { stage = 0; ++round; }
                        continue;
                    }
                    mbar_expect_tx(&full[stage], stage_bytes);
                    if (ad.mode == MODE_PLAIN) {
                        if (kb < ad.k_split) tma_load_2d(sa, &tmap_x, &full[stage], kb * BK, row0);
                        else tma_load_2d(sa, &tmap_x2, &full[stage], (kb - ad.k_split) * BK, row0);
                    } else if (ad.mode == MODE_CONV) {
                        const int ky = tap / ad.kw, kx = tap - ky * ad.kw;
                        tma_load_4d(sa, &tmap_x, &full[stage], cc * BK, p0 + kx - ad.pad_l, p1 + ky - ad.pad_t, p2);
                    } else if (ad.mode == MODE_UPCAT) {
                        // channels [0, k_split * 64): nearest-neighbour 2x upsample of the low-resolution tensor -- pixel (y, x) reads
                        // (y / 2, x / 2) through two zero-stride "duplicate" dimensions of the map; the rest: the skip tensor
                        if (kb < ad.k_split) tma_load_5d(sa, &tmap_x, &full[stage], kb * BK, 0, p0 >> 1, 0, p2 * (ad.pad_t >> 1) + (p1 >> 1));
                        else tma_load_4d(sa, &tmap_x2, &full[stage], (kb - ad.k_split) * BK, p0, p1, p2);
                    } else {                                     // K order of the reference concat: (dy,dx) = (0,0),(1,0),(0,1),(1,1)
                        tma_load_5d(sa, &tmap_x, &full[stage], cc * BK, tap >> 1, p0, tap & 1, p1);
                    }
                    if (++cc == ad.cpb) { cc = 0; ++tap; }
                    if (!WRES) tma_load_2d(sa + A_BYTES, &tmap_w, &full[stage], kb * BK, nt * BN);
                    if (++stage == stages) { stage = 0; ++round; }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        if (lane == 0 && crank == 0) {                           // pair: the leader issues for both CTAs
            constexpr uint32_t idesc = idesc_bf16(BM * MT_ROWS, BN, false, false);
            int stage = 0, round = 0;
            if (WRES && n_iter > 0) { mbar_wait(&w_full, 0); fence_after_sync(); }
            for (int it = 0; it < n_iter; ++it) {
                const int a = it % NACC;
                LTRACE(0, it, 0);
                if (it >= NACC) mbar_wait(&acc_empty[a], (uint32_t)(((it / NACC) - 1) & 1));
                LTRACE(0, it, 1);
                fence_after_sync();
                int ky = 0, cc = 0;                              // halo: (kernel row, channel block) of the stage
                for (int st = 0; st < nst; ++st) {
                    mbar_wait(&full[stage], (uint32_t)(round & 1));
                    if (st < 13) LTRACE(0, it, 2 + st);
                    fence_after_sync();
                    const uint32_t sa = ring_base + stage * stage_bytes;
                    for (int kx = 0; kx < sub; ++kx) {           // halo: tap kx = the tile's rows kx .. kx + 127
                        const int kb = halo ? (ky * sub + kx) * ad.cpb + cc : st;
                        const uint64_t da = desc_sw128(sa + kx * 128), db = desc_sw128(WRES ? sbase + kb * B_BYTES : sa + a_stage + kx * B_BYTES);
#pragma unroll
                        for (int ks = 0; ks < BK / 16; ++ks) {
                            if (CTA2) mma_ss2(tm + a * BN, da + 2 * ks, db + 2 * ks, idesc, (st | kx | ks) != 0);
                            else mma_ss(tm + a * BN, da + 2 * ks, db + 2 * ks, idesc, (st | kx | ks) != 0);
                        }
                    }
                    if (CTA2) mma_commit2(&empty[stage]); else mma_commit(&empty[stage]);
                    if (halo && ++cc == ad.cpb) { cc = 0; ++ky; }
                    if (++stage == stages) { stage = 0; ++round; }
                }
                if (CTA2) mma_commit2(&acc_full[a]); else mma_commit(&acc_full[a]);
                LTRACE(0, it, 15);
            }
        }
    } else {
        // the 64-column boxes of successive tiles go round-robin to the epilogue groups; thread = one row of the box
        const int quarter = warp & 3, eg = warp >> 2;
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t stage_box = staging_base + eg * BOX_BYTES;                        // this group's staging box
        const uint32_t my_row = stage_box + row_in_tile * 128;
        const int sw = row_in_tile & 7;                                                   // SWIZZLE_128B: chunk ^= row % 8
        const bool issuer = quarter == 0 && lane == 0;
        constexpr int NBOX = BN / 64;
        uint32_t res_phase = 0;
        constexpr bool ANY = EPI == EPI_ANY;
        const float* __restrict__ bias = (ANY || (EPI & EPI_BIAS)) ? ep.bias : nullptr;
        const bool has_residual = ANY ? ep.has_residual != 0 : (EPI & EPI_RES) != 0;
        const bool has_ln = ANY ? ep.ln_stats != nullptr : (EPI & EPI_LN) != 0;
        const bool has_stats = ANY ? ep.stats_out != nullptr : (EPI & EPI_STATS) != 0;
        const int act = ANY ? ep.act : (EPI >> EPI_ACT_SHIFT);
        // (mean, rstd) of this thread's A row for a folded LayerNorm, fetched one tile ahead: issued per box, the L2 latency of
        // this load was exposed in every box (+0.25 ms on the fc1 GEMM)
        auto load_mr = [&](int i) {
            float2 r = make_float2(0.f, 1.f);
            if (has_ln && i < n_iter) {
                int mt_, nt_;
                tile_at(i, mt_, nt_);
                const int gr = (mt_ * MT_ROWS + crank) * BM + row_in_tile;
                if (gr < ep.M) {
                    if (ep.ln_boxes == 0) {
                        r = __ldg(reinterpret_cast<const float2*>(ep.ln_stats) + gr);
                    } else {        // partial (sum, sum of squares) pairs from the GEMM that wrote the row: no finalize pass
                        const float2* p = reinterpret_cast<const float2*>(ep.ln_stats) + gr;
                        float sx = 0.f, sy = 0.f;                 // up to six partial pairs (C <= 384): issued together, one L2 round trip
#pragma unroll
                        for (int b = 0; b < 6; ++b) {
                            const float2 q = b < ep.ln_boxes ? __ldg(p + (size_t)b * ep.M) : make_float2(0.f, 0.f);
                            sx += q.x; sy += q.y;
                        }
                        const float mean = sx * ep.ln_inv_k;
                        r = make_float2(mean, rsqrtf(fmaxf(fmaf(-mean, mean, sy * ep.ln_inv_k), 0.f) + ep.ln_eps));
                    }
                }
            }
            return r;
        };
        float2 mr_next = load_mr(0);
        for (int it = 0; it < n_iter; ++it) {
            const int a = it % NACC;
            int mt, nt;
            tile_at(it, mt, nt);
            const int mrow = mt * MT_ROWS + crank;                       // this CTA's 128-row tile
            const int grow = mrow * BM + row_in_tile;                    // global output row of this thread
            const float2 mr = mr_next;
            mr_next = load_mr(it + 1);
            if (issuer) LTRACE(2 + eg, it, 0);
            mbar_wait(&acc_full[a], (uint32_t)((it / NACC) & 1));
            if (issuer) LTRACE(2 + eg, it, 1);
            fence_after_sync();
            bool released = false;
#pragma unroll 1
            for (int bx = 0; bx < NBOX; ++bx) {
                if (((it * NBOX + bx) & (EPI_GROUPS - 1)) != eg) continue;
                const int col0 = nt * BN + bx * 64;
                if (issuer) {
                    LTRACE(2 + eg, it, 2 + 4 * bx);
                    tma_store_wait_read();                               // the previous box of this group has left smem
                    LTRACE(2 + eg, it, 3 + 4 * bx);
                    if (has_residual) {                                  // residual box lands in the staging buffer (TMA, swizzled)
                        mbar_expect_tx(&res_full[eg], BOX_BYTES);
                        tma_load_2d(stage_box, &tmap_r, &res_full[eg], col0, (mrow % ep.res_tiles) * BM);
                    }
                }
                // accumulator columns in chunks of 16, the next chunk in flight while this one is processed (the kernel
                // runs at the register cap of 576 threads: two 32-column loads spilled once the LayerNorm terms were added)
                uint32_t ra[16], rb[16];
                const uint32_t tsrc = tm + lane_addr + a * BN + bx * 64;
                tmem_ld16(tsrc, ra);
                tmem_wait_ld();
                asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");      // staging box reusable for everyone
                if (has_residual) { mbar_wait(&res_full[eg], res_phase); res_phase ^= 1; }
                const float rstd = mr.y, nmr = -mr.x * mr.y;                      // rstd and -mean * rstd of the A row
                const uint64_t rstd2 = pack2(rstd, rstd), nmr2 = pack2(nmr, nmr);
                float so = 0.f, sso = 0.f;                                        // statistics of the output row
#pragma unroll
                for (int j = 0; j < 64; j += 8) {
                    if ((j & 15) == 0 && j + 16 < 64) {                           // prefetch the next 16 columns
                        if (j & 16) tmem_ld16(tsrc + j + 16, ra); else tmem_ld16(tsrc + j + 16, rb);
                    }
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = __uint_as_float((j & 16) ? rb[(j & 8) + e] : ra[(j & 8) + e]);
                    float bb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (bias != nullptr) {
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col0 + j + 4));
                        bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
                    }
                    if (has_ln) {                                                 // rstd acc + (bias - mean rstd colsum[n]): two FMAs
                        const float4 c0 = __ldg(reinterpret_cast<const float4*>(ep.ln_colsum + col0 + j));
                        const float4 c1 = __ldg(reinterpret_cast<const float4*>(ep.ln_colsum + col0 + j + 4));
                        const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                        for (int e = 0; e < 8; e += 2)          // packed fp32x2: two columns per FFMA2
                            unpack2(ffma2(rstd2, pack2(v[e], v[e + 1]), ffma2(nmr2, pack2(cs[e], cs[e + 1]), pack2(bb[e], bb[e + 1]))), v[e], v[e + 1]);
                    } else if (bias != nullptr) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] += bb[e];
                    }
                    if (act == 1) {                               // two elements per instruction: the GELU epilogue is issue-bound
#pragma unroll
                        for (int e = 0; e < 8; e += 2) unpack2(gelu_f32x2(pack2(v[e], v[e + 1])), v[e], v[e + 1]);
                    } else if (act == 2) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = silu_fast(v[e]);
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
                    if (has_stats) {                      // of the fp32 values: differs from the stored bf16 row by < 2^-9 / sqrt(N) rms
#pragma unroll
                        for (int e = 0; e < 8; ++e) { so += v[e]; sso = fmaf(v[e], v[e], sso); }
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16(v[0], v[1])), "r"(pack_bf16(v[2], v[3])),
                                 "r"(pack_bf16(v[4], v[5])), "r"(pack_bf16(v[6], v[7])) : "memory");
                    if ((j & 15) == 8 && j + 8 < 64) tmem_wait_ld();              // the prefetched chunk has landed
                    if (j == 40) {
                        // The box's last accumulator columns are in registers (a group takes at most one box of a tile): the
                        // warp releases the accumulator NOW, not after the remaining arithmetic and the store -- with two
                        // accumulator buffers (BN >= 192) the MMA issuer waited ~500 cycles per tile for the epilogue otherwise
                        fence_before_sync();
                        __syncwarp();
                        if (lane == 0) {
                            if (CTA2) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[a]), 0));
                            else mbar_arrive(&acc_empty[a]);
                        }
                        released = true;
                    }
                }
                if (has_stats && grow < ep.M)
                    reinterpret_cast<float2*>(ep.stats_out)[(size_t)(col0 >> 6) * ep.M + grow] = make_float2(so, sso);
                fence_proxy_async();
                asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
                if (issuer) { LTRACE(2 + eg, it, 4 + 4 * bx); tma_store_2d(&tmap_o, stage_box, col0, mrow * BM); tma_store_commit(); }
            }
            if (!released) {                                             // no box of this tile for this group
                fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    if (CTA2) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[a]), 0));  // the leader's MMA thread waits for both epilogues
                    else mbar_arrive(&acc_empty[a]);
                }
            }
        }
        if (issuer) tma_store_wait_all();
    }
    fence_before_sync();
    if (CTA2) cluster_sync_all();                                // no CTA leaves while its peer may still signal its barriers / read its smem
    else __syncthreads();
    if (warp == MMA_WARP) { if (CTA2) tmem_dealloc2(tmem_slot, 512); else tmem_dealloc(tmem_slot, 512); }
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

// bf16 tensor map of rank `rank`; dims / box innermost first, strides in ELEMENTS for dims 1..rank-1
bool make_map(CUtensorMap* m, const void* base, int rank, const long long* dims, const long long* strides, const int* box,
              CUtensorMapL2promotion promo) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) { d[i] = (cuuint64_t)dims[i]; b[i] = (cuuint32_t)box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = (cuuint64_t)strides[i] * 2;
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_map_2d(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows, bool is_output) {
    const long long dims[2] = {cols, rows}, strides[1] = {ld};
    const int box[2] = {64, box_rows};
    return make_map(m, base, 2, dims, strides, box, is_output ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

constexpr int SMEM_LIMIT = 227 * 1024;

template <int BN, int EPI, bool WRES, bool CTA2>
int launch(const CUtensorMap& mx, const CUtensorMap& mx2, const Addressing& ad, const LinearTcArgs& g, int num_sms, cudaStream_t stream) {
    constexpr int BNL = CTA2 ? BN / 2 : BN, MT = CTA2 ? 2 * BM : BM;
    CUtensorMap mw, mo, mr;
    if (!make_map_2d(&mw, g.w, g.N, g.K, g.K, BNL, false) || !make_map_2d(&mo, g.out, g.M, g.N, g.ldo, BM, true)) return SODT_ERR_CUDA;
    const int res_rows = g.residual && g.res_rows > 0 ? g.res_rows : g.M;
    if (!make_map_2d(&mr, g.residual ? g.residual : g.out, res_rows, g.N, g.residual ? g.ldr : g.ldo, BM, true)) return SODT_ERR_CUDA;
    const int num_n_tiles = g.N / BN, num_m_tiles = (g.M + MT - 1) / MT;
    const long long tiles = (long long)num_n_tiles * num_m_tiles;
    if (tiles > 2147483647LL) return SODT_ERR_UNSUPPORTED;
    constexpr int B_BYTES = BNL * BK * 2;
    const int fixed = EPI_GROUPS * BOX_BYTES + 1024 + (WRES ? (g.K / BK) * B_BYTES : 0);
    const int stage_bytes = ad.mode == MODE_HALO ? HALO_BYTES + (WRES ? 0 : ad.kw * B_BYTES) : A_BYTES + (WRES ? 0 : B_BYTES);
    int stages = (SMEM_LIMIT - fixed) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    static const int stage_cap = [] { const char* v = getenv("SODT_MAX_STAGES"); return v ? atoi(v) : MAX_STAGES; }();   // experiments
    if (stages > stage_cap && stage_cap >= 2) stages = stage_cap;
    if (stages < 2) return SODT_ERR_UNSUPPORTED;
    const size_t smem = (size_t)fixed + (size_t)stages * stage_bytes;
    auto kern = linear_tc_kernel<BN, EPI, WRES, CTA2>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    const int walkers = CTA2 ? num_sms / 2 : num_sms;            // CTAs, or CTA pairs
    int grid = (int)(tiles < walkers ? tiles : walkers);
    if (WRES) grid = (num_sms / num_n_tiles) * num_n_tiles;      // a multiple of the N-tile count: CTA c keeps N-tile c % num_n_tiles
    Epilogue ep{};
    ep.bias = g.bias; ep.ln_stats = g.ln_stats; ep.ln_colsum = g.ln_colsum; ep.stats_out = g.stats_out;
    ep.M = g.M; ep.has_residual = g.residual != nullptr ? 1 : 0; ep.act = g.act;
    ep.res_tiles = (res_rows + BM - 1) / BM;
    ep.ln_boxes = g.ln_boxes; ep.ln_inv_k = 1.f / (float)g.K; ep.ln_eps = g.ln_eps;
    if (CTA2) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
        const int K = g.K;
        e = cudaLaunchKernelEx(&cfg, kern, mx, mx2, mw, mo, mr, ep, ad, K, num_n_tiles, num_m_tiles, stages);
        if (e != cudaSuccess) return cuda_status(e);
        return check_launch();
    }
    e = launch_pdl(kern, dim3(grid), dim3(NTHREADS), smem, stream, true, mx, mx2, mw, mo, mr, ep, ad, g.K, num_n_tiles, num_m_tiles, stages);
    if (e != cudaSuccess) return cuda_status(e);
    return check_launch();
}

int pair_min_k() {           // experiments: SODT_PAIR_MIN_K overrides the K threshold
    static const int k = [] { const char* v = getenv("SODT_PAIR_MIN_K"); return v ? atoi(v) : 768; }();
    return k;
}

bool use_pairs() {          // SODT_NO_CTA2=1 keeps every GEMM on single-CTA tiles (A/B measurements)
    static const bool on = [] { const char* v = getenv("SODT_NO_CTA2"); return !(v && v[0] == '1'); }();
    return on;
}

// W stays resident when the CTA's W tile plus >= 3 A stages fit (K <= 192 at BN = 256 / 192) and the M-tiles fill the machine;
// otherwise wide tiles with a long K loop run on CTA pairs
template <int BN, int EPI>
int launch_res(const CUtensorMap& mx, const CUtensorMap& mx2, const Addressing& ad, const LinearTcArgs& g, int num_sms, cudaStream_t stream, bool wres) {
    if (wres) return launch<BN, EPI, true, false>(mx, mx2, ad, g, num_sms, stream);
    if constexpr (BN >= 128) {
        // measured on B200 (tools/prof_pairs.py): +6-11 % at K >= 768 (1315 TFLOP/s at K = 3072), -5 % at K = 384 where the six
        // k-blocks of a tile do not amortise the pair's hand-shakes
        if (use_pairs() && g.K >= pair_min_k() && (long long)g.M >= 256LL * (num_sms / 2)) return launch<BN, EPI, false, true>(mx, mx2, ad, g, num_sms, stream);
    }
    return launch<BN, EPI, false, false>(mx, mx2, ad, g, num_sms, stream);
}

template <int EPI>
int dispatch_bn(const CUtensorMap& mx, const CUtensorMap& mx2, const Addressing& ad, const LinearTcArgs& g, int num_sms, cudaStream_t stream) {
    const int nkb = g.K / BK;
    const long long m_tiles = (g.M + BM - 1) / BM;
    auto fits = [&](int bn) {          // resident W tile + 3 A stages + staging within the shared-memory limit, enough M-tiles per CTA
        const int n_nt = g.N / bn;
        return g.N % bn == 0 && nkb * bn * BK * 2 + 3 * (ad.mode == MODE_HALO ? HALO_BYTES : A_BYTES) + EPI_GROUPS * BOX_BYTES + 1024 <= SMEM_LIMIT && n_nt <= num_sms &&
               m_tiles >= 4LL * (num_sms / n_nt);
    };
    // the tile width is the widest that divides N; W stays resident only if it fits at that width (narrower resident
    // tiles re-read A more often than the W re-fetches they save: measured slower at K = 384)
    if (g.N % 256 == 0) return launch_res<256, EPI>(mx, mx2, ad, g, num_sms, stream, fits(256));
    if (g.N % 192 == 0) return launch_res<192, EPI>(mx, mx2, ad, g, num_sms, stream, fits(192));
    if (g.N % 128 == 0) return launch_res<128, EPI>(mx, mx2, ad, g, num_sms, stream, fits(128));
    return launch_res<64, EPI>(mx, mx2, ad, g, num_sms, stream, fits(64));
}

int dispatch(const CUtensorMap& mx, const CUtensorMap& mx2, const Addressing& ad, const LinearTcArgs& g, int num_sms, cudaStream_t stream) {
    // the epilogue variants the detector uses are specialised; any other combination runs the all-run-time kernel
    const int code = epi_code(g.bias != nullptr, g.ln_stats != nullptr, g.residual != nullptr, g.stats_out != nullptr, g.act);
    switch (code) {
#define SODT_EPI_CASE(bias, ln, res, stats, act) \
        case epi_code(bias, ln, res, stats, act): return dispatch_bn<epi_code(bias, ln, res, stats, act)>(mx, mx2, ad, g, num_sms, stream);
        SODT_EPI_CASE(false, false, false, false, 0)     // necks, PatchMerging reduction
        SODT_EPI_CASE(true, false, false, false, 0)      // Detect's 1x1 conv, unfolded qkv
        SODT_EPI_CASE(true, true, false, false, 0)       // norm1 + qkv, norm2 + fc1 of the conv-MLP
        SODT_EPI_CASE(true, true, false, false, 1)       // norm2 + fc1 + GELU
        SODT_EPI_CASE(true, false, true, true, 0)        // proj / fc2 + residual, emitting row statistics
        SODT_EPI_CASE(true, false, true, false, 0)       // proj / fc2 + residual; patch embedding + pos_embed
        SODT_EPI_CASE(true, false, false, false, 1)      // conv-MLP taps + GELU, unfolded fc1
        SODT_EPI_CASE(true, false, false, false, 2)      // head Conv: conv + folded BN + SiLU
#undef SODT_EPI_CASE
        default: return dispatch_bn<EPI_ANY>(mx, mx2, ad, g, num_sms, stream);
    }
}

}  // namespace

#ifdef LIN_TRACE
extern "C" int sodt_debug_trace(void* host, size_t bytes) {
    return cudaMemcpyFromSymbol(host, g_lin_trace, bytes < sizeof(g_lin_trace) ? bytes : sizeof(g_lin_trace)) == cudaSuccess ? 0 : -1;
}
#endif

bool linear_tc_supported(int M, int N, int K) { return M > 0 && K % BK == 0 && K >= BK && N % 64 == 0 && N >= 64; }

int linear_tc(const LinearTcArgs& g, const void* x2, int ldx2, int k_split, int num_sms, cudaStream_t stream) {
    if (!linear_tc_supported(g.M, g.N, g.K)) return SODT_ERR_UNSUPPORTED;
    if (g.ln_stats && (!g.ln_colsum || x2 != nullptr || g.ln_boxes < 0 || g.ln_boxes > 6)) return SODT_ERR_INVALID_ARG;
    if (g.residual && g.res_rows > 0 && g.res_rows != g.M && (g.res_rows % BM || g.M % g.res_rows)) return SODT_ERR_INVALID_ARG;
    if (g.ldx % 8 || g.ldo % 8 || (g.residual && g.ldr % 8) || g.ldx < (x2 ? k_split : g.K) || g.ldo < g.N) return SODT_ERR_INVALID_ARG;
    Addressing ad{};
    ad.mode = MODE_PLAIN;
    ad.k_split = g.K / BK;
    ad.cpb = 1 << 30;
    CUtensorMap mx, mx2;
    if (x2 != nullptr) {
        if (k_split <= 0 || k_split >= g.K || k_split % BK || ldx2 % 8 || ldx2 < g.K - k_split) return SODT_ERR_INVALID_ARG;
        ad.k_split = k_split / BK;
        if (!make_map_2d(&mx, g.x, g.M, k_split, g.ldx, BM, false) || !make_map_2d(&mx2, x2, g.M, g.K - k_split, ldx2, BM, false))
            return SODT_ERR_CUDA;
    } else {
        if (!make_map_2d(&mx, g.x, g.M, g.K, g.ldx, BM, false)) return SODT_ERR_CUDA;
        mx2 = mx;
    }
    return dispatch(mx, mx2, ad, g, num_sms, stream);
}

// one 128-row tile = 128 consecutive pixels of an image row (W % 128 == 0) or 128 / W whole rows (128 % W == 0)
static bool tile_geometry(int H, int W, int* bw, int* bh) {
    if (W % BM == 0) { *bw = BM; *bh = 1; return true; }
    if (W > 0 && BM % W == 0 && H % (BM / W) == 0) { *bw = W; *bh = BM / W; return true; }
    return false;
}

bool conv_tc_supported(int B, int H, int W, int Cin, int Cout, int kh, int kw) {
    int bw, bh;
    return B > 0 && Cin % BK == 0 && Cin >= BK && Cout % 64 == 0 && Cout >= 64 && kh >= 1 && kw >= 1 && kh <= 7 && kw <= 7 &&
           tile_geometry(H, W, &bw, &bh) && (long long)B * H * W < 2147483647LL;
}

int conv_tc(const void* x, int ldx, const void* w, const float* bias, void* out, int ldo, int B, int H, int W, int Cin, int Cout,
            int kh, int kw, int pad_t, int pad_l, int act, int num_sms, cudaStream_t stream) {
    int bw, bh;
    if (!conv_tc_supported(B, H, W, Cin, Cout, kh, kw) || !tile_geometry(H, W, &bw, &bh)) return SODT_ERR_UNSUPPORTED;
    if (ldx % 8 || ldx < Cin || ldo % 8 || ldo < Cout || pad_t < 0 || pad_l < 0 || pad_t >= kh + H || pad_l >= kw + W) return SODT_ERR_INVALID_ARG;
    LinearTcArgs g{};
    g.x = x; g.ldx = ldx; g.w = w; g.bias = bias; g.residual = nullptr; g.ldr = 0; g.res_rows = 0; g.out = out; g.ldo = ldo;
    g.M = B * H * W; g.N = Cout; g.K = kh * kw * Cin; g.act = act;
    Addressing ad{};
    ad.mode = MODE_CONV; ad.k_split = 0; ad.cpb = Cin / BK; ad.kw = kw; ad.pad_t = pad_t; ad.pad_l = pad_l; ad.HW = H * W; ad.W = W;
    const long long dims[4] = {Cin, W, H, B}, strides[3] = {ldx, (long long)W * ldx, (long long)H * W * ldx};
    CUtensorMap mx;
    // tiles inside one image row and kw >= 2: the kw taps of a kernel row share one halo box (MODE_HALO);
    // SODT_CONV_HALO=0 keeps one box per tap (A/B measurements)
    static const bool halo_on = [] { const char* v = getenv("SODT_CONV_HALO"); return !(v && v[0] == '0'); }();
    if (halo_on && bh == 1 && kw >= 2 && BM + kw - 1 <= HALO_BYTES / 128) {
        const int hbox[4] = {64, BM + kw - 1, 1, 1};
        if (make_map(&mx, x, 4, dims, strides, hbox, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) {
            ad.mode = MODE_HALO;
                    return dispatch(mx, mx, ad, g, num_sms, stream);
        }
    }
    const int box[4] = {64, bw, bh, 1};
    if (!make_map(&mx, x, 4, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return SODT_ERR_CUDA;
    return dispatch(mx, mx, ad, g, num_sms, stream);
}

// 1x1 conv over cat(upsample2x_nearest(low), skip) without the concatenated tensor (head rows "Upsample, Concat, C3")
bool upcat_tc_supported(int B, int H, int W, int C1, int C2, int Cout) {
    int bw, bh;
    return B > 0 && H % 2 == 0 && W % 2 == 0 && C1 % BK == 0 && C1 >= BK && C2 % BK == 0 && C2 >= BK && Cout % 64 == 0 && Cout >= 64 &&
           tile_geometry(H, W, &bw, &bh) && bw % 2 == 0 && (bh == 1 || bh % 2 == 0) && (long long)B * H * W < 2147483647LL;
}

int upcat_tc(const void* low, const void* skip, const void* w, const float* bias, void* out, int ldo, int B, int H, int W, int C1, int C2,
             int Cout, int act, int num_sms, cudaStream_t stream) {
    int bw, bh;
    if (!upcat_tc_supported(B, H, W, C1, C2, Cout) || !tile_geometry(H, W, &bw, &bh)) return SODT_ERR_UNSUPPORTED;
    if (ldo % 8 || ldo < Cout) return SODT_ERR_INVALID_ARG;
    LinearTcArgs g{};
    g.x = low; g.ldx = C1; g.w = w; g.bias = bias; g.residual = nullptr; g.ldr = 0; g.res_rows = 0; g.out = out; g.ldo = ldo;
    g.M = B * H * W; g.N = Cout; g.K = C1 + C2; g.act = act;
    Addressing ad{};
    ad.mode = MODE_UPCAT; ad.k_split = C1 / BK; ad.cpb = 1 << 30; ad.kw = 1; ad.pad_t = H; ad.pad_l = 0; ad.HW = H * W; ad.W = W;   // pad_t carries H
    // low [B, H/2, W/2, C1] as (c, dup_x, x/2, dup_y, b * H/2 + y/2) with zero strides on the duplicate dimensions
    const long long dl[5] = {C1, 2, W / 2, 2, (long long)B * (H / 2)};
    const long long sl[4] = {0, C1, 0, (long long)(W / 2) * C1};
    const int bl[5] = {64, 2, bw / 2, bh == 1 ? 1 : 2, bh == 1 ? 1 : bh / 2};
    const long long ds[4] = {C2, W, H, B}, ss[3] = {C2, (long long)W * C2, (long long)H * W * C2};
    const int bs[4] = {64, bw, bh, 1};
    CUtensorMap ml, ms;
    if (!make_map(&ml, low, 5, dl, sl, bl, CU_TENSOR_MAP_L2_PROMOTION_L2_256B) ||
        !make_map(&ms, skip, 4, ds, ss, bs, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return SODT_ERR_CUDA;
    return dispatch(ml, ms, ad, g, num_sms, stream);
}

bool merge_tc_supported(int B, int H, int W, int C, int N) {
    int bw, bh;
    return B > 0 && H % 2 == 0 && W % 2 == 0 && C % BK == 0 && C >= BK && N % 64 == 0 && N >= 64 &&
           tile_geometry(B * (H / 2), W / 2, &bw, &bh) && (long long)B * H * W < 2147483647LL;
}

int merge_tc(const void* x, const void* w, const float* bias, void* out, int B, int H, int W, int C, int N, int num_sms,
             cudaStream_t stream) {
    int bw, bh;
    if (!merge_tc_supported(B, H, W, C, N) || !tile_geometry(B * (H / 2), W / 2, &bw, &bh)) return SODT_ERR_UNSUPPORTED;
    LinearTcArgs g{};
    g.x = x; g.ldx = C; g.w = w; g.bias = bias; g.residual = nullptr; g.ldr = 0; g.res_rows = 0; g.out = out; g.ldo = N;
    g.M = B * (H / 2) * (W / 2); g.N = N; g.K = 4 * C; g.act = 0;
    Addressing ad{};
    ad.mode = MODE_MERGE; ad.k_split = 0; ad.cpb = C / BK; ad.W = W / 2;
    // in[b, 2i+dy, 2j+dx, c] as (c, dx, j, dy, b*H/2+i)
    const long long dims[5] = {C, 2, W / 2, 2, (long long)B * (H / 2)};
    const long long strides[4] = {C, 2LL * C, (long long)W * C, 2LL * W * C};
    const int box[5] = {64, 1, bw, 1, bh};
    CUtensorMap mx;
    if (!make_map(&mx, x, 5, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return SODT_ERR_CUDA;
    return dispatch(mx, mx, ad, g, num_sms, stream);
}

}  // namespace sodt
