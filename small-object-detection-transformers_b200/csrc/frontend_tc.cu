// The detector's whole front end as ONE tcgen05 kernel, straight from the uint8 images:
//
//   four single-channel 4x4 / stride-4 patch embeddings (R with PatchEmbed's default padding 1; reference
//   backbone_vit.py:69-98,751)  ->  the window-1 cross-channel block: x_p = LayerNorm_p(e_a + e_k) for the pairs R<-G, G<-B,
//   B<-IR, IR<-G (:469-561)  ->  channel concat (:210)  ->  the 1x1 patch embedding Linear(192 -> 192) + bias + absolute
//   position embedding (:212-214), emitting the row statistics the first norm1 needs.
//
// Per 128-token tile (persistent CTAs, one per SM):
//   loader warps   thread = token: 16 pixels per stream (one 32-bit load per image row), u8 / 255 rounded to bf16, written as
//                  the K = 16 A operand [128 x 16] of each stream (canonical no-swizzle K-major core matrices)
//   MMA thread     E_s[128 x 48] = A_s . Wc_s^T: ONE tcgen05.mma per stream (the conv IS a K = 16 GEMM), accumulators in TMEM
//   epilogue 1     two groups of 4 warps, two pairs each; thread = token reads its 48-channel embeddings from TMEM, adds the
//                  pair (+ conv biases), LayerNorm in registers (no shuffles), writes bf16 into the concat tile in shared
//                  memory -- which is the SWIZZLE_128B A operand of the next GEMM and never reaches HBM
//   MMA thread     O[128 x 192] = concat[128 x 192] . Wpe^T (12 MMAs, Wpe resident in shared memory)
//   epilogue 2     4 warps: O + bias + pos (TMA-loaded tile, batch-broadcast) -> row statistics, bf16, TMA store
// HBM traffic: the uint8 images in (4 bytes per token and stream), one [tokens, 192] bf16 tensor out.
#include <cuda.h>

#include "common.cuh"
#include "tc05.cuh"
#include "tma.cuh"

namespace sodt {
namespace {

using namespace tc;

constexpr int BM = 128, E = 48, C = 4 * E;
// The scheduler prefers the highest warp id among the eligible warps of a sub-partition: the patch-embedding epilogue (ONE warp
// per sub-partition, on the critical path accumulator -> store -> next GEMM) gets the top ids, the loaders the lowest.
constexpr int LOADER_WARP0 = 0, MMA_WARP = 3,            // warps 0-2: loaders (96 threads walk the 128 tokens of a tile)
              EPI1_WARP0 = 4, EPI2_WARP0 = 12;
constexpr int NTHREADS = 16 * 32;                         // 4 warps per scheduler: 128 registers per thread available
constexpr int LOADER_THREADS = 3 * 32;
constexpr int NSTG = 4;                                 // ring of output / position-embedding staging boxes
constexpr int BOX = BM * 128;                           // 128 rows x 64 bf16, SWIZZLE_128B
constexpr int WPE_BOX = C * 128;                        // [192 rows x 64 k]
constexpr int A_STREAM = BM * 32;                       // [128 x 16] bf16
constexpr int WC_STREAM = E * 32;                       // [48 x 16] bf16
constexpr int OFF_WPE = 0, OFF_CAT = OFF_WPE + 3 * WPE_BOX, OFF_STG = OFF_CAT + 3 * BOX, OFF_A = OFF_STG + NSTG * BOX,
              OFF_WC = OFF_A + 4 * A_STREAM, OFF_PAR = OFF_WC + 4 * WC_STREAM, SMEM_TOTAL = OFF_PAR + 4 * C * 4;
constexpr int CONV_COL = 0, PE_COL = 256;               // TMEM: stream s at CONV_COL + 64 s, patch-embedding accumulator at 256

struct FrontParams {
    const uint8_t* rgb; long long rb, rc, ry;           // uint8 [B, 3, H, W], unit pixel stride
    const uint8_t* ir; long long ib, iy;                // uint8 [B, >= 1, H, W], channel 0
    const __nv_bfloat16* conv_w;                        // [4][48][16]
    const float* conv_b; const float* ln_w; const float* ln_b;      // [4][48]
    const float* pe_b;                                  // [192]
    float* stats_out;                                   // [3][M][2] or null
    int H, W, h, w, pad, has_pos, pos_tiles;
    long long M;
    int num_tiles;
    float eps;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src), "r"(c0), "r"(c1) : "memory");
}
// four uint8 pixels (one little-endian word) -> two packed bf16 pairs of u / 255 (product rounded to fp32, then to bf16:
// exactly what `img.float() / 255` followed by the cast to the model dtype gives, reference basics/test.py:124-130)
__device__ __forceinline__ void px4(uint32_t wd, uint32_t& lo, uint32_t& hi) {
    const float s = 1.f / 255.f;
    lo = pack_bf16(__fmul_rn((float)(wd & 0xffu), s), __fmul_rn((float)((wd >> 8) & 0xffu), s));
    hi = pack_bf16(__fmul_rn((float)((wd >> 16) & 0xffu), s), __fmul_rn((float)(wd >> 24), s));
}

// FE_TRACE (experiments only): CTA 0 records clock64 at the hand-offs of its first 32 tiles into stats_out
#ifdef FE_TRACE
#define FTRACE(role, i, k) do { if (blockIdx.x == 0 && (i) < 32) reinterpret_cast<long long*>(p.stats_out)[((role) * 32 + (i)) * 8 + (k)] = clock64(); } while (0)
#else
#define FTRACE(role, i, k) do { } while (0)
#endif

template <bool HAS_POS>
__global__ void __launch_bounds__(NTHREADS, 1)
frontend_tc_kernel(const __grid_constant__ CUtensorMap tmap_wpe, const __grid_constant__ CUtensorMap tmap_pos,
                   const __grid_constant__ CUtensorMap tmap_o, const FrontParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t wpe_full, a_full, a_empty, conv_full, conv_empty, cat_full, cat_empty, acc_full, acc_empty, pos_full[NSTG];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sptr = smem_raw + (sbase - smem_u32(smem_raw));
    float* par = reinterpret_cast<float*>(sptr + OFF_PAR);        // pair bias [4][48] | ln_w | ln_b | pe_b | (spare)
    const int n_iter = (int)blockIdx.x < p.num_tiles ? (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (tid == 0) {
        mbar_init(&wpe_full, 1);
        mbar_init(&a_full, LOADER_THREADS); mbar_init(&a_empty, 1);
        for (int s = 0; s < NSTG; ++s) mbar_init(&pos_full[s], 1);
        mbar_init(&conv_full, 1); mbar_init(&conv_empty, 256);
        mbar_init(&cat_full, 256); mbar_init(&cat_empty, 1);
        mbar_init(&acc_full, 1); mbar_init(&acc_empty, 128);
        fence_barrier_init();
    }
    if (warp == MMA_WARP) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    pdl_trigger();
    pdl_wait();                                      // programmatic dependent launch (common.cuh): no global memory is touched above
    // parameters: LayerNorm input of pair q is e_a + e_k + (conv_b[a] + conv_b[k]); pairs (a, k) = (R,G), (G,B), (B,IR), (IR,G)
    for (int i = tid; i < C; i += NTHREADS) {
        const int q = i / E, c = i - q * E, k = q == 3 ? 1 : q + 1;
        par[i] = p.conv_b[q * E + c] + p.conv_b[k * E + c];
        par[C + i] = p.ln_w[i];
        par[2 * C + i] = p.ln_b[i];
        par[3 * C + i] = p.pe_b[i];
    }
    // conv weights as the B operand [48 x 16] of each stream: 8-row groups of 256 B = two 8 x 16-byte core matrices (k 0-7 | 8-15)
    for (int i = tid; i < 4 * E * 2; i += NTHREADS) {
        const int s = i / (E * 2), r = (i >> 1) % E, j = i & 1;
        const uint4 v = *reinterpret_cast<const uint4*>(p.conv_w + (s * E + r) * 16 + j * 8);
        *reinterpret_cast<uint4*>(sptr + OFF_WC + s * WC_STREAM + (r >> 3) * 256 + j * 128 + (r & 7) * 16) = v;
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_slot;

    if (warp == MMA_WARP) {
        if (lane == 0 && n_iter > 0) {
            tma::expect_tx(&wpe_full, 3 * WPE_BOX);
            for (int kb = 0; kb < 3; ++kb) tma_load_2d(sbase + OFF_WPE + kb * WPE_BOX, &tmap_wpe, &wpe_full, kb * 64, 0);
            constexpr uint32_t idc = idesc_bf16(BM, E, false, false), idp = idesc_bf16(BM, C, false, false);
            auto conv = [&](int i) {
                mbar_wait(&a_full, (uint32_t)(i & 1));
                FTRACE(0, i, 0);
                if (i > 0) mbar_wait(&conv_empty, (uint32_t)((i - 1) & 1));
                FTRACE(0, i, 1);
                fence_after_sync();
#pragma unroll
                for (int s = 0; s < 4; ++s)
                    mma_ss(tm + CONV_COL + 64 * s, smem_desc(sbase + OFF_A + s * A_STREAM, 128, 256),
                           smem_desc(sbase + OFF_WC + s * WC_STREAM, 128, 256), idc, false);
                mma_commit(&a_empty);
                mma_commit(&conv_full);
            };
            conv(0);
            mbar_wait(&wpe_full, 0);
            for (int i = 0; i < n_iter; ++i) {
                if (i + 1 < n_iter) conv(i + 1);
                mbar_wait(&cat_full, (uint32_t)(i & 1));
                FTRACE(0, i, 2);
                if (i > 0) mbar_wait(&acc_empty, (uint32_t)((i - 1) & 1));
                FTRACE(0, i, 3);
                fence_after_sync();
#pragma unroll
                for (int kb = 0; kb < 3; ++kb) {
                    const uint64_t da = tma::desc_sw128(sbase + OFF_CAT + kb * BOX), db = tma::desc_sw128(sbase + OFF_WPE + kb * WPE_BOX);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) mma_ss(tm + PE_COL, da + 2 * ks, db + 2 * ks, idp, (kb | ks) != 0);
                }
                mma_commit(&cat_empty);
                mma_commit(&acc_full);
            }
        }
    } else if (warp < MMA_WARP) {
        // warps 12-14
        // ------------------------------------------------------------------ pixel patches -> A operands
        // the tile's 128 tokens x 4 streams = 512 patches are dealt to the 96 loader threads (patch = lt + 96 k); ALL row loads of
        // a tile are issued before the first conversion: one memory round trip per tile (per-row dependent loads made the
        // kernel latency-bound at 0.95 ms)
        const int lt = tid - LOADER_WARP0 * 32;
        constexpr int NP = (4 * BM + LOADER_THREADS - 1) / LOADER_THREADS;       // 6 patches per thread (the last one partly)
        const long long hw = (long long)p.h * p.w;
        for (int i = 0; i < n_iter; ++i) {
            const long long tok0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * BM;
            const long long b0 = tok0 / hw;                                       // the only 64-bit division of the tile
            const long long rem0 = tok0 - b0 * hw;
            uint32_t cur[NP][4], prv[NP][4], meta[NP];
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const int idx = lt + LOADER_THREADS * k, s = (idx >> 7) & 3, r = idx & (BM - 1);
                const bool ok = idx < 4 * BM && tok0 + r < p.M;
                long long bi = b0, rl = rem0 + r;                                 // a tile may run over image boundaries (tiny images: several)
                while (rl >= hw) { rl -= hw; ++bi; }
                if (!ok) { bi = 0; rl = 0; }                                      // clamped: every load below has a valid address
                const int rem = (int)rl, ty = rem / p.w, tx = rem - ty * p.w;
                const uint8_t* img = s == 3 ? p.ir + bi * p.ib : p.rgb + bi * p.rb + s * p.rc;
                const long long sy = s == 3 ? p.iy : p.ry;
                const int off = s == 0 ? p.pad : 0;                               // only the R stream is padded
                meta[k] = (uint32_t)r | ((uint32_t)s << 8) | (ok ? 1u << 16 : 0u) | (off ? 1u << 17 : 0u) | (tx > 0 ? 1u << 18 : 0u) |
                          (ty > 0 ? 1u << 19 : 0u);
#pragma unroll
                for (int ky = 0; ky < 4; ++ky) {
                    const int y = 4 * ty + ky - off;
                    const uint32_t* row = reinterpret_cast<const uint32_t*>(img + (long long)(y < 0 ? 0 : y) * sy);
                    cur[k][ky] = __ldg(row + tx);
                    prv[k][ky] = off ? __ldg(row + (tx > 0 ? tx - 1 : 0)) : 0u;
                }
            }
            if (lt == 0) FTRACE(1, i, 0);
            if (i >= 1) mbar_wait(&a_empty, (uint32_t)((i - 1) & 1));          // the pixels wait in registers until the previous tile's MMAs have read A
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const uint32_t m = meta[k];
                if (lt + LOADER_THREADS * k < 4 * BM) {
                    const int r = m & 0xff, s = (m >> 8) & 3;
                    uint32_t q[8];
#pragma unroll
                    for (int ky = 0; ky < 4; ++ky) {
                        uint32_t wd = cur[k][ky];
                        if (m & (1u << 17)) {                                     // padded stream: pixels 4 tx - 1 .. 4 tx + 2, row 4 ty + ky - 1
                            wd = (wd << 8) | ((m & (1u << 18)) ? prv[k][ky] >> 24 : 0u);
                            if (ky == 0 && !(m & (1u << 19))) wd = 0u;
                        }
                        if (!(m & (1u << 16))) wd = 0u;
                        px4(wd, q[2 * ky], q[2 * ky + 1]);
                    }
                    const uint32_t dst = sbase + OFF_A + s * A_STREAM + (r >> 3) * 256 + (r & 7) * 16;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 128), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]) : "memory");
                }
            }
            fence_proxy_async();
            if (lt == 0) FTRACE(1, i, 1);
            mbar_arrive(&a_full);                                                 // one arrival per loader thread
        }
    } else if (warp < EPI2_WARP0) {
        // ------------------------------------------------------------------ pair add + LayerNorm -> concat tile
        const int quarter = warp & 3, grp = (warp - EPI1_WARP0) >> 2;           // grp 0: pairs 0, 1;  grp 1: pairs 2, 3
        const int row = quarter * 32 + lane, sw = row & 7;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        // streams: X is kept for both pairs of the group, Y pairs with it first, Z second
        const int sx = grp == 0 ? 1 : 3, sy = grp == 0 ? 0 : 2, sz = grp == 0 ? 2 : 1;
        // packed fp32x2 arithmetic (FADD2 / FFMA2): two channels per instruction; statistics in one pass (sum, sum of squares)
        // 48 accumulator columns of a stream: an x32 and an x16 load in flight together, ONE wait (every tcgen05.wait::ld costs a
        // TMEM round trip of a few hundred cycles while MMAs run; nine of them per tile were the epilogue's critical path)
        auto ld48_issue = [&](int s, uint32_t (&a)[32], uint32_t (&b)[16]) {
            tmem_ld32(tm + lane_addr + CONV_COL + 64 * s, a);
            tmem_ld16(tm + lane_addr + CONV_COL + 64 * s + 32, b);
        };
        auto ld48_pack = [&](const uint32_t (&a)[32], const uint32_t (&b)[16], uint64_t (&v)[E / 2]) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = pack2(__uint_as_float(a[2 * j]), __uint_as_float(a[2 * j + 1]));
#pragma unroll
            for (int j = 0; j < 8; ++j) v[16 + j] = pack2(__uint_as_float(b[2 * j]), __uint_as_float(b[2 * j + 1]));
        };
        const uint32_t par_s = sbase + OFF_PAR;
        auto lds2x2 = [&](uint32_t addr, uint64_t& a, uint64_t& b) {            // four consecutive floats as two packed pairs
            asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));     // parameters: written once before the CTA barrier
        };
        for (int i = 0; i < n_iter; ++i) {
            mbar_wait(&conv_full, (uint32_t)(i & 1));
            fence_after_sync();
            if (lane == 0 && quarter == 0) FTRACE(2 + grp, i, 0);
            uint64_t X[E / 2], Y[E / 2];
            {
                uint32_t xa[32], xb[16], ya[32], yb[16];
                ld48_issue(sx, xa, xb);
                ld48_issue(sy, ya, yb);
                tmem_wait_ld();
                ld48_pack(xa, xb, X);
                ld48_pack(ya, yb, Y);
            }
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int q = 2 * grp + half;                                     // pair = output channel block = LayerNorm index
                if (half == 1) {
                    uint32_t za[32], zb[16];
                    ld48_issue(sz, za, zb);
                    tmem_wait_ld();
                    ld48_pack(za, zb, Y);
                    fence_before_sync();
                    mbar_arrive(&conv_empty);                                     // this thread's embeddings are in registers
                }
                uint64_t s1 = 0ull, s2 = 0ull;
#pragma unroll
                for (int j = 0; j < E / 2; j += 2) {
                    uint64_t b0, b1;
                    lds2x2(par_s + 4 * (q * E + 2 * j), b0, b1);
                    Y[j] = fadd2(fadd2(X[j], Y[j]), b0);
                    Y[j + 1] = fadd2(fadd2(X[j + 1], Y[j + 1]), b1);
                    s1 = fadd2(fadd2(s1, Y[j]), Y[j + 1]);
                    s2 = ffma2(Y[j], Y[j], s2);
                    s2 = ffma2(Y[j + 1], Y[j + 1], s2);
                }
                float a0, a1, q0, q1;
                unpack2(s1, a0, a1);
                unpack2(s2, q0, q1);
                const float mean = (a0 + a1) * (1.f / E);
                const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, (q0 + q1) * (1.f / E)), 0.f) + p.eps), nm = -mean * rstd;
                const uint64_t rstd2 = pack2(rstd, rstd), nm2 = pack2(nm, nm);
                if (lane == 0 && quarter == 0) FTRACE(2 + grp, i, 1 + 2 * half);
                if (half == 0 && i > 0) mbar_wait(&cat_empty, (uint32_t)((i - 1) & 1));      // the previous tile's GEMM has read the concat tile
                if (lane == 0 && quarter == 0) FTRACE(2 + grp, i, 2 + 2 * half);
#pragma unroll
                for (int c = 0; c < E; c += 8) {
                    uint64_t w0, w1, w2, w3, l0, l1, l2, l3;
                    lds2x2(par_s + 4 * (C + q * E + c), w0, w1);
                    lds2x2(par_s + 4 * (C + q * E + c + 4), w2, w3);
                    lds2x2(par_s + 4 * (2 * C + q * E + c), l0, l1);
                    lds2x2(par_s + 4 * (2 * C + q * E + c + 4), l2, l3);
                    const uint64_t y0 = ffma2(ffma2(Y[c / 2], rstd2, nm2), w0, l0), y1 = ffma2(ffma2(Y[c / 2 + 1], rstd2, nm2), w1, l1);
                    const uint64_t y2 = ffma2(ffma2(Y[c / 2 + 2], rstd2, nm2), w2, l2), y3 = ffma2(ffma2(Y[c / 2 + 3], rstd2, nm2), w3, l3);
                    float f0, f1, f2, f3, f4, f5, f6, f7;
                    unpack2(y0, f0, f1); unpack2(y1, f2, f3); unpack2(y2, f4, f5); unpack2(y3, f6, f7);
                    const int ch = q * E + c;                                     // channel of the concat row; 16-byte chunk ch / 8
                    const uint32_t dst = sbase + OFF_CAT + (ch >> 6) * BOX + row * 128 + ((((ch & 63) >> 3) ^ sw) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16(f0, f1)), "r"(pack_bf16(f2, f3)),
                                 "r"(pack_bf16(f4, f5)), "r"(pack_bf16(f6, f7)));
                }
            }
            fence_proxy_async();
            if (lane == 0 && quarter == 0) FTRACE(2 + grp, i, 5);
            mbar_arrive(&cat_full);
        }
    } else {
        // ------------------------------------------------------------------ patch embedding epilogue: + bias + pos, statistics, store
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane, sw = row & 7;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const bool issuer = tid == EPI2_WARP0 * 32;
        const uint32_t peb_s = sbase + OFF_PAR + 4 * 3 * C;
        // Output boxes (64 columns of a tile) go through a ring of NSTG staging boxes: the position-embedding box lands there by
        // TMA, is updated in place and leaves by TMA.  Box n (n = 3 tile + column box) uses slot n % NSTG; its pos load is issued
        // right after the store of box n - NSTG + 1, two box computations ahead of its use (one whole-tile buffer serialised
        // load -> compute -> store: 0.52 ms).
        const int total_boxes = 3 * n_iter;
        auto load_pos = [&](int n) {
            const int tl = (int)blockIdx.x + (n / 3) * (int)gridDim.x, j = n % 3, slot = n % NSTG;
            tma::expect_tx(&pos_full[slot], BOX);
            tma_load_2d(sbase + OFF_STG + slot * BOX, &tmap_pos, &pos_full[slot], j * 64, (tl % p.pos_tiles) * BM);
        };
        if (issuer && HAS_POS)
            for (int n = 0; n < NSTG - 1 && n < total_boxes; ++n) load_pos(n);
        for (int i = 0; i < n_iter; ++i) {
            const int tile = (int)blockIdx.x + i * (int)gridDim.x;
            const long long grow = (long long)tile * BM + row;
            if (issuer) FTRACE(4, i, 0);
            mbar_wait(&acc_full, (uint32_t)(i & 1));
            fence_after_sync();
            if (issuer) FTRACE(4, i, 1);
            uint32_t ta[32], tb[32];
            tmem_ld32(tm + lane_addr + PE_COL, ta);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                uint64_t so2 = 0ull, sso2 = 0ull;
                const int n = 3 * i + j, slot = n % NSTG;
                if (HAS_POS) {
                    mbar_wait(&pos_full[slot], (uint32_t)((n / NSTG) & 1));
                    if (issuer) FTRACE(4, i, 2 + j);
                } else {                                                          // the slot's previous store has been read
                    if (issuer) tma::store_wait_read<NSTG - 1>();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                const uint32_t my_row = sbase + OFF_STG + slot * BOX + row * 128;
#pragma unroll
                for (int c = 0; c < 64; c += 32) {
                    tmem_wait_ld();                                               // this 32-column chunk has landed
                    if (j == 2 && c == 32) { fence_before_sync(); mbar_arrive(&acc_empty); }      // accumulator fully in registers: the next GEMM may start
                    const int nxt = j * 64 + c + 32;                              // prefetch the next chunk into the other buffer
                    if (nxt < C) { if (c == 0) tmem_ld32(tm + lane_addr + PE_COL + nxt, tb); else tmem_ld32(tm + lane_addr + PE_COL + nxt, ta); }
                    const uint32_t* t = c == 0 ? ta : tb;
#pragma unroll
                    for (int e = 0; e < 32; e += 8) {
                        const uint32_t dst = my_row + ((((c + e) >> 3) ^ sw) << 4);
                        uint64_t b0, b1, b2, b3;
                        asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(b0), "=l"(b1) : "r"(peb_s + 4 * (j * 64 + c + e)));
                        asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(b2), "=l"(b3) : "r"(peb_s + 4 * (j * 64 + c + e + 4)));
                        uint64_t v0 = fadd2(pack2(__uint_as_float(t[e]), __uint_as_float(t[e + 1])), b0);
                        uint64_t v1 = fadd2(pack2(__uint_as_float(t[e + 2]), __uint_as_float(t[e + 3])), b1);
                        uint64_t v2 = fadd2(pack2(__uint_as_float(t[e + 4]), __uint_as_float(t[e + 5])), b2);
                        uint64_t v3 = fadd2(pack2(__uint_as_float(t[e + 6]), __uint_as_float(t[e + 7])), b3);
                        if (HAS_POS) {                                          // bf16 pair -> fp32 pair: (w << 16, w & 0xffff0000)
                            uint32_t q0, q1, q2, q3;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3) : "r"(dst));
                            v0 = fadd2(v0, pack2(__uint_as_float(q0 << 16), __uint_as_float(q0 & 0xffff0000u)));
                            v1 = fadd2(v1, pack2(__uint_as_float(q1 << 16), __uint_as_float(q1 & 0xffff0000u)));
                            v2 = fadd2(v2, pack2(__uint_as_float(q2 << 16), __uint_as_float(q2 & 0xffff0000u)));
                            v3 = fadd2(v3, pack2(__uint_as_float(q3 << 16), __uint_as_float(q3 & 0xffff0000u)));
                        }
                        so2 = fadd2(fadd2(so2, v0), fadd2(v1, fadd2(v2, v3)));
                        sso2 = ffma2(v0, v0, sso2); sso2 = ffma2(v1, v1, sso2); sso2 = ffma2(v2, v2, sso2); sso2 = ffma2(v3, v3, sso2);
                        float f0, f1, f2, f3, f4, f5, f6, f7;
                        unpack2(v0, f0, f1); unpack2(v1, f2, f3); unpack2(v2, f4, f5); unpack2(v3, f6, f7);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16(f0, f1)), "r"(pack_bf16(f2, f3)),
                                     "r"(pack_bf16(f4, f5)), "r"(pack_bf16(f6, f7)));
                    }
                }
                float so, sso;
                { float x0, x1, y0, y1; unpack2(so2, x0, x1); unpack2(sso2, y0, y1); so = x0 + x1; sso = y0 + y1; }
#ifndef FE_TRACE
                if (p.stats_out != nullptr && grow < p.M)
                    reinterpret_cast<float2*>(p.stats_out)[(size_t)j * p.M + grow] = make_float2(so, sso);
#endif
                fence_proxy_async();
                if (issuer) FTRACE(4, i, 5 + (j == 2));
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (issuer) {
                    tma_store_2d(&tmap_o, sbase + OFF_STG + slot * BOX, j * 64, tile * BM);
                    tma::store_commit();
                    if (HAS_POS && n + NSTG - 1 < total_boxes) {
                        tma::store_wait_read<1>();                                // every store but the one just issued has been read
                        load_pos(n + NSTG - 1);
                    }
                }
            }
        }
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

}  // namespace
}  // namespace sodt

extern "C" int sodt_frontend_embed_u8_supported(int B, int H, int W, int E, int embed_dim, int pos_rows) {
    if (B <= 0 || H < 4 || W < 4 || H % 4 || W % 4 || E != 48 || embed_dim != 192) return 0;
    const long long hw = (long long)(H / 4) * (W / 4);
    if (pos_rows != 0 && (pos_rows != hw || hw % 128)) return 0;         // a tile's 128 tokens must be 128 consecutive pos rows
    return (long long)B * hw < 2147483647LL - 128 ? 1 : 0;
}

extern "C" int sodt_frontend_embed_u8_fwd(const void* rgb, long long rb, long long rc, long long ry, const void* ir, long long ib,
                                          long long iy, const void* conv_w, const float* conv_b, const float* ln_w, const float* ln_b,
                                          const void* pe_w, const float* pe_b, const void* pos, int pos_rows, void* out,
                                          float* stats_out, int B, int H, int W, int E, int embed_dim, int pad_r, float eps, void* stream) {
    using namespace sodt;
    if (!rgb || !ir || !conv_w || !conv_b || !ln_w || !ln_b || !pe_w || !pe_b || !out) return SODT_ERR_INVALID_ARG;
    if ((pos == nullptr) != (pos_rows == 0) || (pad_r != 0 && pad_r != 1)) return SODT_ERR_INVALID_ARG;
    if (!sodt_frontend_embed_u8_supported(B, H, W, E, embed_dim, pos_rows)) return SODT_ERR_UNSUPPORTED;
    // one aligned 32-bit load per image row and token: rows must start on 4-byte boundaries
    if ((reinterpret_cast<uintptr_t>(rgb) & 3) || (reinterpret_cast<uintptr_t>(ir) & 3) || (rb & 3) || (rc & 3) || (ry & 3) || (ib & 3) || (iy & 3))
        return SODT_ERR_ALIGNMENT;
    if (!aligned16(conv_w) || !aligned16(pe_w) || !aligned16(out) || (pos && !aligned16(pos)) || (reinterpret_cast<uintptr_t>(stats_out) & 7))
        return SODT_ERR_ALIGNMENT;
    FrontParams p{};
    p.rgb = static_cast<const uint8_t*>(rgb); p.rb = rb; p.rc = rc; p.ry = ry;
    p.ir = static_cast<const uint8_t*>(ir); p.ib = ib; p.iy = iy;
    p.conv_w = static_cast<const __nv_bfloat16*>(conv_w); p.conv_b = conv_b; p.ln_w = ln_w; p.ln_b = ln_b; p.pe_b = pe_b;
    p.stats_out = stats_out;
    p.H = H; p.W = W; p.h = H / 4; p.w = W / 4; p.pad = pad_r; p.has_pos = pos ? 1 : 0;
    p.M = (long long)B * p.h * p.w;
    p.num_tiles = (int)((p.M + BM - 1) / BM);
    p.pos_tiles = pos ? pos_rows / BM : 1;
    p.eps = eps;
    CUtensorMap mw, mp, mo;
    if (!map_2d(&mw, pe_w, C, C, C, C, false) || !map_2d(&mo, out, p.M, C, C, BM, true) ||
        !map_2d(&mp, pos ? pos : out, pos ? pos_rows : p.M, C, C, BM, false))
        return SODT_ERR_CUDA;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = SMEM_TOTAL + 1024;
    auto kern = pos ? frontend_tc_kernel<true> : frontend_tc_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    const int grid = p.num_tiles < sms ? p.num_tiles : sms;
    e = launch_pdl(kern, dim3(grid), dim3(NTHREADS), smem, static_cast<cudaStream_t>(stream), true, mw, mp, mo, p);
    if (e != cudaSuccess) return cuda_status(e);
    return check_launch();
}
