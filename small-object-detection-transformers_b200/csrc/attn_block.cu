// The attention half of a Swin block up to the projection, as ONE kernel (first half):
//
//   o = WindowAttention_core( LayerNorm(x) W_qkv^T + b_qkv )      reference backbone_vit.py:1090-1123 (norm1, roll, partition,
//                                                                   qkv Linear :968, scores + bias + mask + softmax + PV :969-989)
//
// for 8x8 windows at C = 192 (stage 1: 12 heads of 16 or 6 heads of 32 channels).  The qkv tensor -- written once (3 T) and read
// once (3 T) per block by the norm1 + qkv GEMM and window_attn_win8 -- never exists: the kernel reads the residual stream x
// (1 T) and writes the attention output o (1 T); proj + residual stay the GEMM that follows (its W and the staging do not fit
// beside this pipeline).
//
// Tile = two horizontally adjacent windows of the rolled frame = 128 tokens, loaded as (64 channels, 8, 8) TMA boxes of the
// [B*H, W, C] image (roll + partition = box coordinates; wrapped border windows per image row, as in window_attn_win8.cu) into
// three SWIZZLE_128B k-block tiles [128 tokens x 64 channels].  Per 64-channel head group cg (3 per tile) -- a "unit":
//   GEMM      D[128 x 192] = X W_cg'^T, 12 tcgen05.mma (M=128, N=192); W_cg' = the q | k | v rows of the group in the
//             LayerNorm-folded weight (three 64 x 64 boxes per k-block, streamed from L2 through a two-stage ring)
//   convert   the eight softmax warps read D from tensor memory, apply rstd * acc + (b' - mean * rstd * colsum) -- the folded
//             LayerNorm, in the operation order of linear_tc's epilogue, so the values are bit-identical to the qkv tensor --
//             and write bf16 K | Q | V operand tiles per window: exactly the stage layout window_attn_win8 loads by TMA
//   attention per window: lane-masked QK^T into S (two heads per 128 lanes), softmax with the
//             relative position bias and the shifted-window mask, P written IN PLACE over S (packed bf16), O = P V accumulated
//             in the S columns the softmax has already read (tensor memory: D 192 + 2 groups x 2 head pairs x 64 = 448 columns)
//   epilogue  O / rowsum -> bf16 -> the window's (dead) Q tile as staging -> TMA store of the un-rolled image tile
// The two softmax groups own alternate units (their own K | Q | V buffer each) and run half a unit apart, so that one group's
// softmax covers the other's conversion, MMA round trips and epilogue.
// Warps: 0-7 softmax / convert / epilogue (two groups of four), 8-9 attention MMA issuers (one per group), 10 GEMM issuer,
// 11 TMA producer (x tiles, W ring).
#include "common.cuh"
#include "tma.cuh"

namespace sodt {
namespace {

using namespace tc;

constexpr int WS = 8, NTOK = 64, ROWS = 128, NG = 2;
constexpr int SM_WARPS = 4;                                  // softmax warps per group
constexpr int MMA_WARP0 = NG * SM_WARPS, GEMM_WARP = MMA_WARP0 + NG, TMA_WARP = GEMM_WARP + 1;
constexpr int NTHREADS = (TMA_WARP + 1) * 32;               // 384
constexpr int WIN_BYTES = NTOK * 128;                        // one operand tile of one window: 64 tokens x 128 B
constexpr int STAGE_BYTES = 3 * WIN_BYTES;                   // K Q V of one window (K first: the odd heads' A operand starts one tile below Q)
constexpr int OFF_K = 0, OFF_Q = WIN_BYTES, OFF_V = 2 * WIN_BYTES;
constexpr int KB = 3;                                        // 64-channel k-blocks of x (C = 192)
constexpr int XKB_BYTES = ROWS * 128;                        // one k-block of the x tile: 128 tokens x 128 B
constexpr int X_BYTES = KB * XKB_BYTES;                      // 48 KB
constexpr int WST_BYTES = 192 * 128;                         // one k-block of W_cg': 192 rows (q | k | v) x 128 B
constexpr int WSTAGES = 2;
constexpr int OFF_X = 0, OFF_QKV = X_BYTES, OFF_W = OFF_QKV + 2 * NG * STAGE_BYTES, OFF_TAB = OFF_W + WSTAGES * WST_BYTES;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int TAB_ROW = 16, TAB_HEAD = (2 * WS - 1) * TAB_ROW, TAB_COPIES = 2;      // bias-table image of window_attn_win8 (prepared workspace)
__host__ __device__ constexpr int tab_copy_stride(int heads) { return ((heads * TAB_HEAD * 4 + 95) / 128 * 128 + 32) / 4; }
constexpr uint32_t TM_D = 0, TM_S = 192;                     // tensor memory: D [0, 192); group g, head pair pr: S / P / O at 192 + g * 128 + pr * 64
constexpr uint32_t ALL = 0xFFFFFFFFu;

// ATTN_TRACE (experiments only, tools/trace_attn_block.py): CTA 0 records clock64 at the hand-offs of its first 64 units
#ifdef ATTN_TRACE
__device__ long long g_attn_trace[6][64][16];
__device__ long long g_cta_cycles[256][2];
#define ATRACE(role, idx, ev) do { if (blockIdx.x == 0 && (idx) < 64) g_attn_trace[role][idx][ev] = clock64(); } while (0)
#else
#define ATRACE(role, idx, ev) do { } while (0)
#endif

struct Geo {
    int H, W, nww, nwh, nW, shift;
    long long total_tiles;                                   // window pairs
};
struct WinBox {
    int x0, y0, yg_base;
    bool wrap_x, wrap_y, last_row, last_col;
};
__device__ __forceinline__ WinBox win_box(const Geo& g, long long wdx) {
    WinBox r;
    const int b = (int)(wdx / g.nW);
    const int win = (int)(wdx - (long long)b * g.nW);
    const int wy = win / g.nww, wx = win - wy * g.nww;
    r.x0 = wx * WS + g.shift;
    r.y0 = wy * WS + g.shift;
#ifdef ATTN_X_NOWRAP      // timing experiments only (border windows read zeros past the image edge and lose the wrapped part)
    r.wrap_x = false;
    r.wrap_y = false;
#else
    r.wrap_x = r.x0 + WS > g.W;
    r.wrap_y = r.y0 + WS > g.H;
#endif
    r.last_row = wy == g.nwh - 1;
    r.last_col = wx == g.nww - 1;
    r.yg_base = b * g.H;
    return r;
}
struct Maps {
    CUtensorMap full, row8, row_a, row_b;                    // boxes (64, 8, 8), (64, 8, 1), (64, 8 - shift, 1), (64, shift, 1)
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// The KB channel blocks of one window of x, issued by the whole producer warp: tile kb lands at dst + kb * XKB_BYTES
template <bool SHIFTED>
__device__ __forceinline__ void load_window(const Maps& m, const Geo& geo, const WinBox& b, uint32_t dst, uint64_t* bar, int lane) {
    if constexpr (!SHIFTED) {                                // unshifted frame: no window wraps
        if (lane < KB) tma::load_3d(dst + lane * XKB_BYTES, &m.full, bar, lane * 64, b.x0, b.yg_base + b.y0);
        return;
    }
    const int per_box = !b.wrap_x && !b.wrap_y ? 1 : (b.wrap_x ? 2 * WS : WS);
    for (int item = lane; item < KB * per_box; item += 32) {
        const int bi = item / per_box, sub = item - bi * per_box;
        const uint32_t sm = dst + bi * XKB_BYTES;
        const int ch = bi * 64;
        if (per_box == 1) {
            tma::load_3d(sm, &m.full, bar, ch, b.x0, b.yg_base + b.y0);
        } else {
            const int ty = b.wrap_x ? sub >> 1 : sub;
            int ys = b.y0 + ty; if (ys >= geo.H) ys -= geo.H;
            const int yg = b.yg_base + ys;
            if (!b.wrap_x) tma::load_3d(sm + ty * 1024, &m.row8, bar, ch, b.x0, yg);
            else if ((sub & 1) == 0) tma::load_3d(sm + ty * 1024, &m.row_a, bar, ch, b.x0, yg);
            else tma::load_3d(sm + ty * 1024 + (WS - geo.shift) * 128, &m.row_b, bar, ch, 0, yg);
        }
    }
}
template <bool SHIFTED>
__device__ __forceinline__ void store_tile(const Maps& m, const Geo& geo, const WinBox& b, uint32_t sm, int c0, int lane) {
    if (!SHIFTED || (!b.wrap_x && !b.wrap_y)) {
        if (lane == 0) tma::store_3d(&m.full, sm, c0, b.x0, b.yg_base + b.y0);
        return;
    }
    const int n = b.wrap_x ? 2 * WS : WS;
    if (lane < n) {
        const int ty = b.wrap_x ? lane >> 1 : lane;
        int ys = b.y0 + ty; if (ys >= geo.H) ys -= geo.H;
        const int yg = b.yg_base + ys;
        if (!b.wrap_x) tma::store_3d(&m.row8, sm + ty * 1024, c0, b.x0, yg);
        else if ((lane & 1) == 0) tma::store_3d(&m.row_a, sm + ty * 1024, c0, b.x0, yg);
        else tma::store_3d(&m.row_b, sm + ty * 1024 + (WS - geo.shift) * 128, c0, 0, yg);
    }
}

struct Params {
    const float* table;        // prepared bias-table image (window_attn_win8_prepare)
    const float* ln_stats;     // [M][2] (mean, rstd), or [ln_boxes][M][2] partial (sum, sum of squares)
    const float* colsum;       // [3C] row sums of the folded weight
    const float* bias;         // [3C] folded bias
    long long M;
    int ln_boxes;
    float ln_inv_k, ln_eps;
    int C, heads;
    float scale, mask_value;
};

// SHIFTED = false: the un-shifted blocks' instantiation without the wrapped-window and mask code (smaller hot loop)
template <int HD, bool SHIFTED>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_block_kernel(const __grid_constant__ Maps in_maps, const __grid_constant__ Maps out_maps, const __grid_constant__ CUtensorMap w_map,
                  const Params p, const Geo geo) {
    constexpr int G = 64 / HD;                // heads per 64-channel group
    constexpr uint32_t O_OFF = HD == 16 ? 32 : 64;    // O of a pair: the high half of its S columns (2 x 16) or the group's second 64 columns (2 x 32)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t x_full, x_empty, w_full[WSTAGES], w_empty[WSTAGES], d_full[NG], d_free, qk_ready[NG], v_ready[NG], s_full[NG][2], s_free[NG][2], p_full[NG][2], pv_done[NG][2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    float* tab = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)) + OFF_TAB);
    float* lnp = tab + TAB_COPIES * tab_copy_stride(p.heads);      // [3C] folded bias, [3C] column sums: read by every conversion
    const int heads = p.heads, C = p.C;
    int my_tiles = 0;
    if ((long long)blockIdx.x < geo.total_tiles) my_tiles = (int)((geo.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int n_units = my_tiles * KB;        // unit u = (tile u / 3, head group u % 3)

    if (tid == 0) {
        mbar_init(&x_full, 1); mbar_init(&x_empty, 1); mbar_init(&d_full[0], 1); mbar_init(&d_full[1], 1); mbar_init(&d_free, SM_WARPS);
        for (int s = 0; s < WSTAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int g = 0; g < NG; ++g) {
            mbar_init(&qk_ready[g], SM_WARPS); mbar_init(&v_ready[g], SM_WARPS);
            for (int pr = 0; pr < 2; ++pr) { mbar_init(&s_full[g][pr], 1); mbar_init(&p_full[g][pr], ROWS); mbar_init(&s_free[g][pr], ROWS); mbar_init(&pv_done[g][pr], 1); }
        }
        fence_barrier_init();
    }
    if (warp == MMA_WARP0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    pdl_trigger();
    pdl_wait();                    // programmatic dependent launch (common.cuh): no global memory is touched above
    {
        const int n4 = TAB_COPIES * tab_copy_stride(heads) / 4;
        const float4* src = reinterpret_cast<const float4*>(p.table);
        float4* dst = reinterpret_cast<float4*>(tab);
        for (int e = tid; e < n4; e += NTHREADS) dst[e] = src[e];
        for (int e = tid; e < 3 * p.C; e += NTHREADS) { lnp[e] = p.bias[e]; lnp[3 * p.C + e] = p.colsum[e]; }
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_slot;
#ifdef ATTN_TRACE
    const long long cta_t0 = clock64();
#endif

    if (warp == TMA_WARP) {
        // ======================================================= producer: x tile of every window pair, W ring of every unit
        if (lane == 0) {
            tma::prefetch_map(&in_maps.full); tma::prefetch_map(&w_map);
            if (geo.shift > 0) { tma::prefetch_map(&in_maps.row8); tma::prefetch_map(&in_maps.row_a); tma::prefetch_map(&in_maps.row_b); }
        }
        int kbc = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
            const WinBox wb0 = win_box(geo, 2 * tile), wb1 = win_box(geo, 2 * tile + 1);     // (before the wait: the x tile's load is on the critical path of every tile)
            if (lane == 0) {
                if (t > 0) mbar_wait(&x_empty, (uint32_t)((t - 1) & 1));
                ATRACE(5, t, 0);
                tma::expect_tx(&x_full, X_BYTES);
            }
            __syncwarp();
            load_window<SHIFTED>(in_maps, geo, wb0, sbase + OFF_X, &x_full, lane);
            load_window<SHIFTED>(in_maps, geo, wb1, sbase + OFF_X + WIN_BYTES, &x_full, lane);
            if (lane == 0) {
                ATRACE(5, t, 1);
                for (int cg = 0; cg < KB; ++cg) {
                    for (int kb = 0; kb < KB; ++kb, ++kbc) {
                        const int s = kbc % WSTAGES;
                        if (kbc >= WSTAGES) mbar_wait(&w_empty[s], (uint32_t)(((kbc / WSTAGES) - 1) & 1));
                        tma::expect_tx(&w_full[s], WST_BYTES);
                        const uint32_t dst = sbase + OFF_W + s * WST_BYTES;
#pragma unroll
                        for (int part = 0; part < 3; ++part)        // q | k | v rows of the head group
                            tma_load_2d(dst + part * (64 * 128), &w_map, &w_full[s], kb * 64, part * C + cg * 64);
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == GEMM_WARP) {
        // ======================================================= qkv GEMM issuer: D = X W_cg'^T, one unit ahead of the conversion
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16(ROWS, 192, false, false);
            int kbc = 0;
            for (int t = 0; t < my_tiles; ++t) {
                ATRACE(0, t * KB, 6);
                mbar_wait(&x_full, (uint32_t)(t & 1));
                ATRACE(0, t * KB, 7);
                for (int cg = 0; cg < KB; ++cg) {
                    const int u = t * KB + cg;
                    ATRACE(0, u, 0);
                    if (u > 0) mbar_wait(&d_free, (uint32_t)((u - 1) & 1));
                    ATRACE(0, u, 1);
                    fence_after_sync();
                    for (int kb = 0; kb < KB; ++kb, ++kbc) {
                        const int s = kbc % WSTAGES;
                        mbar_wait(&w_full[s], (uint32_t)((kbc / WSTAGES) & 1));
                        ATRACE(0, u, 2 + kb);
                        fence_after_sync();
                        const uint64_t da = tma::desc_sw128(sbase + OFF_X + kb * XKB_BYTES), db = tma::desc_sw128(sbase + OFF_W + s * WST_BYTES);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) mma_ss(tm + TM_D, da + 2 * ks, db + 2 * ks, idesc, (kb | ks) != 0);
                        mma_commit(&w_empty[s]);
                    }
                    ATRACE(0, u, 5);
                    mma_commit(&d_full[u & 1]);                     // one barrier per softmax group: a waiter is never two phases behind
                    if (cg == KB - 1) mma_commit(&x_empty);
                }
            }
        }
    } else if (warp >= MMA_WARP0) {
        // ======================================================= attention MMA issuers: group g takes the units g, g + 2, ...
        if (lane == 0) {
            constexpr uint32_t idesc_s = idesc_bf16(ROWS, NTOK, false, false);
            constexpr uint32_t idesc_o = idesc_bf16(ROWS, 2 * HD, false, true);
            const int g = warp - MMA_WARP0;
            const uint64_t d0 = tma::desc_sw128(sbase + OFF_QKV);
            const uint32_t tS = tm + TM_S + g * 128;
            // Items of a unit: i = 0..3 = (window i / 2, head pair i % 2); item i lives in the 64 tensor-memory columns of pair i % 2
            // (S, then P in place, O in the high half).  The scores of item i + 2 are issued as soon as the softmax warps have read
            // O of item i (in the middle of the softmax of item i + 1), so that they are ready when that softmax ends.
            auto issue_qk = [&](int w, int pr) {
                const uint32_t stage = (uint32_t)((g * 2 + w) * STAGE_BYTES);
                const uint32_t off = (stage + pr * (2 * HD * 2)) >> 4;       // see window_attn_win8.cu: both heads of a pair read the same Q tile
                const uint64_t qd = d0 + off + (OFF_Q >> 4), kd = d0 + off + (OFF_K >> 4);
                const uint64_t qd_odd = qd - (WIN_BYTES >> 4) + ((HD * 2) >> 4), kd_odd = kd + ((HD * 2) >> 4);
                fence_after_sync();
                mma_ss_masked(tS + pr * 64, qd, kd, idesc_s, false, 0u, 0u, ALL, ALL);
                mma_ss_masked(tS + pr * 64, qd_odd, kd_odd, idesc_s, false, ALL, ALL, 0u, 0u);
                mma_commit(&s_full[g][pr]);
            };
            for (int u = g, j = 0; u < n_units; u += NG, ++j) {
                mbar_wait_spin(&qk_ready[g], (uint32_t)(j & 1));                 // K and Q tiles of both windows are written (V follows)
#pragma unroll
                for (int pr = 0; pr < 2; ++pr) {                                // items 0, 1: their columns were freed by the previous unit's items 2, 3
                    if (j > 0) mbar_wait_spin(&s_free[g][pr], (uint32_t)((2 * j - 1) & 1));
                    issue_qk(0, pr);
                }
                ATRACE(1 + g, 2 * j, 1);
                mbar_wait_spin(&v_ready[g], (uint32_t)(j & 1));
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int w = i >> 1, pr = i & 1;
                    const uint32_t par = (uint32_t)((2 * j + w) & 1);           // phase of the pair's barriers: one per item on its columns
                    mbar_wait_spin(&p_full[g][pr], par);
                    if (pr == 0) ATRACE(1 + g, 2 * j + w, 2);
                    fence_after_sync();
                    const uint64_t vd = d0 + (((uint32_t)((g * 2 + w) * STAGE_BYTES) + OFF_V + pr * (2 * HD * 2)) >> 4);
#pragma unroll
                    for (int ks = 0; ks < NTOK / 16; ++ks)      // O[128 x 2hd] = P [V_even | V_odd]: P in the low 32 columns of the pair's S
                        mma_ts(tS + pr * 64 + O_OFF, tS + pr * 64 + ks * 8, vd + ks * (2048 >> 4), idesc_o, ks > 0);
                    mma_commit(&pv_done[g][pr]);
                    if (pr == 1) ATRACE(1 + g, 2 * j + w, 3);
                    if (i < 2) {                                                // scores of item i + 2 (window 1, same pair)
                        mbar_wait_spin(&s_free[g][pr], par);
                        issue_qk(1, pr);
                    }
                }
            }
        }
    } else {
        // ======================================================= softmax groups: group g owns the units g, g + 2, ... : it converts
        // D into the K | Q | V tiles of both windows (buffer g) and then runs the attention of the two windows one after the
        // other.  The groups work on different units, half a unit apart: one group's softmax covers the other's conversion,
        // MMA round trips and epilogue (the D hand-off alternates between them).
        const int g = warp / SM_WARPS, quarter = warp & 3;
        const int row = quarter * 32 + lane;                       // TMEM lane: D row = token of the pair; S row = 64 * head parity + token
        const int hp = row >> 6, ti = row & 63, ty = ti >> 3, tx = ti & 7;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t tS = tm + TM_S + g * 128 + lane_addr;
        const uint32_t tD = tm + TM_D + lane_addr;
        const float c = p.scale * LOG2E, mv2 = p.mask_value * LOG2E;
        const uint64_t c2 = pack2(c, c);
        const int r0 = WS - 1 - tx, cp = r0 & 1;
        const float* tab_row = tab + cp * tab_copy_stride(heads) + (ty + WS - 1) * TAB_ROW + (r0 - cp);
        const int s_ = SHIFTED ? geo.shift : 0;
        const uint32_t srow = (uint32_t)ti * 128, sw = (uint32_t)(ti & 7);
        const bool storer = quarter == 0;                          // warp 0 of the group issues the group's TMA stores

        // (mean, rstd) of D row `row` of tile t: the token's row in the un-rolled [B*H*W] image
        auto load_mr = [&](int t) {
            float2 r = make_float2(0.f, 1.f);
            if (t < my_tiles) {
                const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
                const WinBox wb = win_box(geo, 2 * tile + hp);      // D rows 0-63 = first window, 64-127 = second
                int ys = wb.y0 + ty; if (ys >= geo.H) ys -= geo.H;
                int xs = wb.x0 + tx; if (xs >= geo.W) xs -= geo.W;
                const long long gr = ((long long)(wb.yg_base + ys)) * geo.W + xs;
                if (p.ln_boxes == 0) {
                    r = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + gr);
                } else {
                    const float2* q0 = reinterpret_cast<const float2*>(p.ln_stats) + gr;
                    float sx = 0.f, sy = 0.f;
#pragma unroll
                    for (int b = 0; b < 6; ++b) {
                        const float2 q = b < p.ln_boxes ? __ldg(q0 + (size_t)b * p.M) : make_float2(0.f, 0.f);
                        sx += q.x; sy += q.y;
                    }
                    const float mean = sx * p.ln_inv_k;
                    r = make_float2(mean, rsqrtf(fmaxf(fmaf(-mean, mean, sy * p.ln_inv_k), 0.f) + p.ln_eps));
                }
            }
            return r;
        };
        float2 mr_cur = make_float2(0.f, 1.f), mr_next = load_mr(0);
        int cur_t = -1, next_t = 0;

        // the unit's tile geometry is loop-carried: the next unit's is computed while the last item's P V runs (its 64-bit divisions
        // are not free, and nothing else is ready then)
        int t = g / KB, cg = g - t * KB;
        WinBox box0 = win_box(geo, 2 * ((long long)blockIdx.x + (long long)t * gridDim.x));
        WinBox box1 = win_box(geo, 2 * ((long long)blockIdx.x + (long long)t * gridDim.x) + 1);
        for (int u = g; u < n_units; u += NG) {
            // ---------------------------------------------------------------- D of unit u -> bf16 K | Q | V tiles of both windows
            if (t != cur_t) { mr_cur = mr_next; cur_t = t; }
            {
                const int nt = (u + NG) / KB;                       // the tile of the group's next unit: its statistics one unit ahead
                if (nt != next_t) { mr_next = load_mr(nt); next_t = nt; }
            }
            if (tid == g * 128) ATRACE(3 + g, u >> 1, 0);
            mbar_wait(&d_full[g], (uint32_t)((u >> 1) & 1));
            if (tid == g * 128) ATRACE(3 + g, u >> 1, 1);
            fence_after_sync();
            if (u >= NG) {
                // buffer g was read by the group's previous unit: its MMAs have completed (waited for in its epilogues); window 0's
                // output tile (staged in its Q tile) must have left; window 1's (staged in window 0's V tile): before V is written
                if (storer) tma::store_wait_read<1>();
                asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory");
            }
            {
                const float rstd = mr_cur.y, nmr = -mr_cur.x * mr_cur.y;
                const uint64_t rstd2 = pack2(rstd, rstd), nmr2 = pack2(nmr, nmr);
                const uint32_t wbase = sbase + OFF_QKV + (uint32_t)((g * 2 + hp) * STAGE_BYTES) + srow;
                const float* lb = lnp + cg * 64;                        // bias / colsum of D column c: lb[(c / 64) * C + c % 64] (+ 3C)
                uint32_t ra[32], rb[32];
                tmem_ld32(tD, ra);
#pragma unroll
                for (int ch = 0; ch < 6; ++ch) {                        // 32 columns per step, the next step's load in flight
                    tmem_wait_ld();
                    if (ch + 1 < 6) { if (ch & 1) tmem_ld32(tD + (ch + 1) * 32, ra); else tmem_ld32(tD + (ch + 1) * 32, rb); }
                    if (ch == 5) {                                      // D is in registers: the next unit's GEMM may overwrite it
                        fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&d_free);
                    }
                    const int c0 = ch * 32, part = c0 >> 6, cw = c0 & 63;
                    const float* lc = lb + part * C + cw;
                    const uint32_t tile_off = part == 0 ? OFF_Q : part == 1 ? OFF_K : OFF_V;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const float4 b0 = *reinterpret_cast<const float4*>(lc + j), b1 = *reinterpret_cast<const float4*>(lc + j + 4);
                        const float4 s0 = *reinterpret_cast<const float4*>(lc + 3 * C + j), s1 = *reinterpret_cast<const float4*>(lc + 3 * C + j + 4);
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w}, cs[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {        // rstd acc + (bias - mean rstd colsum): the arithmetic of linear_tc's epilogue
                            const uint32_t a0 = (ch & 1) ? rb[j + e] : ra[j + e], a1 = (ch & 1) ? rb[j + e + 1] : ra[j + e + 1];
                            unpack2(ffma2(rstd2, pack2(__uint_as_float(a0), __uint_as_float(a1)),
                                          ffma2(nmr2, pack2(cs[e], cs[e + 1]), pack2(bb[e], bb[e + 1]))), v[e], v[e + 1]);
                        }
                        const uint32_t chunk = (uint32_t)((cw + j) >> 3) ^ sw;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wbase + tile_off + (chunk << 4)), "r"(pack_bf16(v[0], v[1])),
                                     "r"(pack_bf16(v[2], v[3])), "r"(pack_bf16(v[4], v[5])), "r"(pack_bf16(v[6], v[7])) : "memory");
                    }
                    if (ch == 3) {                                      // Q and K of both windows are written: the scores may be issued
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&qk_ready[g]);
                        if (u >= NG) {
                            if (storer) tma::store_wait_read<0>();
                            asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory");
                        }
                    }
                    if (ch == 5) {
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&v_ready[g]);
                    }
                }
                if (tid == g * 128) ATRACE(3 + g, u >> 1, 2);
            }
            // ---------------------------------------------------------------- attention: items (window, head pair) = (0,0) (0,1) (1,0) (1,1)
            // The epilogue of item i - 1 (O -> staging tile) runs in the middle of the softmax of item i, when its P V has long
            // completed; it frees the pair's columns for the scores of item i + 1.
            const int jj = u >> 1;                                              // own unit index: phases of the per-pair barriers
            float prev_sum = 0.f;
            if (storer && lane == 0) {                          // descriptors of the stores this unit will issue (the row maps are rarely used)
                tma::prefetch_map(&out_maps.full);
                if (SHIFTED && (box1.wrap_x || box1.wrap_y || box0.wrap_y)) {
                    tma::prefetch_map(&out_maps.row8); tma::prefetch_map(&out_maps.row_a); tma::prefetch_map(&out_maps.row_b);
                }
            }
            const uint32_t stage0 = sbase + OFF_QKV + (uint32_t)((g * 2) * STAGE_BYTES);
            // staging tiles: window 0 -> its own Q tile (dead after its scores); window 1 -> window 0's V tile (dead after its P V)
            auto epilogue = [&](int i_prev) {
                const int w = i_prev >> 1, pr = i_prev & 1;
                mbar_wait(&pv_done[g][pr], (uint32_t)((2 * jj + w) & 1));
                fence_after_sync();
                uint32_t o[HD];
                tmem_ld16(tS + pr * 64 + O_OFF + hp * HD, o);
                tmem_wait_ld();
                fence_before_sync();
                mbar_arrive(&s_free[g][pr]);                    // the pair's columns are free: the next scores on them may be issued
                const uint32_t tile_s = stage0 + (w == 0 ? OFF_Q : OFF_V);
                const float inv = fast_rcp(prev_sum);
#pragma unroll
                for (int j2 = 0; j2 < HD; j2 += 8) {
                    const uint32_t chunk = (uint32_t)(((2 * pr + hp) * HD + j2) >> 3) ^ sw;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_s + srow + (chunk << 4)),
                                 "r"(pack_bf16(__uint_as_float(o[j2]) * inv, __uint_as_float(o[j2 + 1]) * inv)),
                                 "r"(pack_bf16(__uint_as_float(o[j2 + 2]) * inv, __uint_as_float(o[j2 + 3]) * inv)),
                                 "r"(pack_bf16(__uint_as_float(o[j2 + 4]) * inv, __uint_as_float(o[j2 + 5]) * inv)),
                                 "r"(pack_bf16(__uint_as_float(o[j2 + 6]) * inv, __uint_as_float(o[j2 + 7]) * inv)) : "memory");
                }
                if (pr == 1) {                                  // both pairs of the window are staged: the tile leaves
                    fence_proxy_async();
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory");
                    if (storer) {
                        store_tile<SHIFTED>(out_maps, geo, w ? box1 : box0, tile_s, cg * 64, lane);
                        tma::store_commit();
                    }
                    if (tid == g * 128) ATRACE(3 + g, u >> 1, 6 + 4 * w);
                }
            };
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {                       // (not unrolled: four copies of the softmax body thrash the instruction cache)
                const int w = i >> 1, pr = i & 1;
                // shifted-window mask of a border window: key (yj, xj) is masked for this query if their row regions differ (last window
                // row) or their column regions differ (last window column) -- one test per key row + four precomputed column pairs
                uint32_t rowm = 0;                                              // bit yj: the whole key row is masked
                uint64_t colp[4] = {0ull, 0ull, 0ull, 0ull};                    // addend pairs of key columns (2q, 2q + 1)
                const bool last_row = w ? box1.last_row : box0.last_row, last_col = w ? box1.last_col : box0.last_col;
                if (SHIFTED && s_ > 0) {
                    if (last_row) rowm = (ty >= WS - s_) ? ~(0xFFu << (WS - s_)) & 0xFFu : (0xFFu << (WS - s_)) & 0xFFu;
                    if (last_col) {
                        const uint32_t cm = (tx >= WS - s_) ? ~(0xFFu << (WS - s_)) & 0xFFu : (0xFFu << (WS - s_)) & 0xFFu;
#pragma unroll
                        for (int q = 0; q < 4; ++q) colp[q] = pack2((cm >> (2 * q)) & 1u ? mv2 : 0.f, (cm >> (2 * q + 1)) & 1u ? mv2 : 0.f);
                    }
                }
#ifdef ATTN_X_NOMASK      // timing experiments only (wrong results on border windows)
                const bool any_mask = false;
#else
                const bool any_mask = SHIFTED && s_ > 0 && (last_row || last_col);
#endif
                mbar_wait(&s_full[g][pr], (uint32_t)((2 * jj + w) & 1));
                if (pr == 0 && tid == g * 128) ATRACE(3 + g, u >> 1, 3 + 4 * w);
                fence_after_sync();
                const int h = cg * G + 2 * pr + hp;
                uint64_t tt[NTOK / 2];
                uint32_t pk[NTOK / 2];
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    uint32_t ra[32];
                    tmem_ld32(tS + pr * 64 + part * 32, ra);
                    tmem_wait_ld();
                    const float* tb = tab_row + h * TAB_HEAD - part * 4 * TAB_ROW;
#pragma unroll
                    for (int yj = 0; yj < 4; ++yj) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 b = *reinterpret_cast<const float2*>(tb - yj * TAB_ROW + 2 * q);
                            tt[part * 16 + yj * 4 + q] = ffma2(pack2(__uint_as_float(ra[yj * 8 + 2 * q]), __uint_as_float(ra[yj * 8 + 2 * q + 1])), c2, pack2(b.x, b.y));
                        }
                    }
                }
                if (any_mask) {                                 // (t + 0 = t: the unmasked keys keep their value)
                    const uint64_t mv22 = pack2(mv2, mv2);
#pragma unroll
                    for (int yj = 0; yj < WS; ++yj) {
                        const bool rm = (rowm >> yj) & 1u;
#pragma unroll
                        for (int q = 0; q < 4; ++q) tt[yj * 4 + q] = fadd2(tt[yj * 4 + q], rm ? mv22 : colp[q]);
                    }
                }
                float m4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float lo, hi;
                    unpack2(tt[q], lo, hi);
                    m4[q] = fmaxf(lo, hi);
                }
#pragma unroll
                for (int j2 = 4; j2 < NTOK / 2; ++j2) {
                    float lo, hi;
                    unpack2(tt[j2], lo, hi);
                    m4[j2 & 3] = fmax3(m4[j2 & 3], lo, hi);
                }
                const float mx = fmaxf(fmax3(m4[0], m4[1], m4[2]), m4[3]);
                if (i > 0) epilogue(i - 1);                     // the previous item's O: its P V ran during the score phase above
                const uint64_t nmx2 = pack2(-mx, -mx);
                uint64_t sum2 = 0ull;
#pragma unroll
                for (int j2 = 0; j2 < NTOK / 2; ++j2) {
                    float lo, hi;
                    unpack2(fadd2(tt[j2], nmx2), lo, hi);
                    const float p0 = fast_exp2(lo), p1 = fast_exp2(hi);
                    sum2 = fadd2(sum2, pack2(p0, p1));
                    pk[j2] = pack_bf16(p0, p1);
                }
                float a, b;
                unpack2(sum2, a, b);
                prev_sum = a + b;
                tmem_st(tS + pr * 64, pk);                      // P in place: this lane's scores of the pair are all in registers
                tmem_wait_st();
                fence_before_sync();
                mbar_arrive(&p_full[g][pr]);                    // the pair's P V starts while the next item's softmax runs
                if (pr == 1 && tid == g * 128) ATRACE(3 + g, u >> 1, 4 + 4 * w);
            }
            int nt = t, ncg = cg;
            WinBox nb0 = box0, nb1 = box1;
            if (u + NG < n_units) {
                nt = (u + NG) / KB; ncg = u + NG - nt * KB;
                if (nt != t) {
                    const long long ntile = (long long)blockIdx.x + (long long)nt * gridDim.x;
                    nb0 = win_box(geo, 2 * ntile); nb1 = win_box(geo, 2 * ntile + 1);
                }
            }
            epilogue(3);                                        // drain: the unit's last item (its P V is exposed here)
            t = nt; cg = ncg; box0 = nb0; box1 = nb1;
        }
        if (storer) tma::store_wait_all();
    }
    fence_before_sync();
    __syncthreads();
#ifdef ATTN_TRACE
    if (tid == 0 && blockIdx.x < 256) { g_cta_cycles[blockIdx.x][0] = clock64() - cta_t0; g_cta_cycles[blockIdx.x][1] = my_tiles; }
#endif
    if (warp == MMA_WARP0) tmem_dealloc(tmem_slot, 512);
}

bool make_maps(Maps* m, const void* base, int B, int H, int W, int Cfull, int shift, bool is_output) {
    const long long dims[3] = {Cfull, W, (long long)B * H}, strides[2] = {Cfull, (long long)W * Cfull};
    const int sa = shift > 0 ? WS - shift : WS, sb = shift > 0 ? shift : WS;
    const int bf[3] = {64, WS, WS}, b8[3] = {64, WS, 1}, ba[3] = {64, sa, 1}, bb[3] = {64, sb, 1};
    const CUtensorMapL2promotion promo = is_output ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    return tma::make_map_bf16(&m->full, base, 3, dims, strides, bf, promo) && tma::make_map_bf16(&m->row8, base, 3, dims, strides, b8, promo) &&
           tma::make_map_bf16(&m->row_a, base, 3, dims, strides, ba, promo) && tma::make_map_bf16(&m->row_b, base, 3, dims, strides, bb, promo);
}

size_t smem_bytes(int heads) { return (size_t)OFF_TAB + (size_t)TAB_COPIES * tab_copy_stride(heads) * sizeof(float) + 2 * 3 * KB * 64 * sizeof(float) + 1024; }

}  // namespace

#ifdef ATTN_TRACE
extern "C" int sodt_attn_cta_cycles(void* host, size_t bytes) {
    return cudaMemcpyFromSymbol(host, g_cta_cycles, bytes < sizeof(g_cta_cycles) ? bytes : sizeof(g_cta_cycles)) == cudaSuccess ? 0 : -1;
}
extern "C" int sodt_attn_trace(void* host, size_t bytes) {
    return cudaMemcpyFromSymbol(host, g_attn_trace, bytes < sizeof(g_attn_trace) ? bytes : sizeof(g_attn_trace)) == cudaSuccess ? 0 : -1;
}
#endif

bool attn_block_supported(int H, int W, int C, int heads, int ws, int shift) {
    if (ws != WS || C != KB * 64 || heads <= 0 || C % heads || H % WS || W % (2 * WS) || shift < 0 || shift >= WS) return false;
    const int hd = C / heads;
    return hd == 16 && smem_bytes(heads) <= 227 * 1024;      // (head pairs of 2 x 16 channels: 64 tensor-memory columns hold S, P and O of a pair)
}

int attn_block(const void* x, const float* ln_stats, int ln_boxes, float ln_eps, const void* w, const float* colsum, const float* bias,
               const void* table_ws, void* out, int B, int H, int W, int C, int heads, int shift, float scale, float mask_value,
               int num_sms, cudaStream_t stream) {
    if (!attn_block_supported(H, W, C, heads, WS, shift)) return SODT_ERR_UNSUPPORTED;
    if ((long long)B * H > 2147483647LL) return SODT_ERR_UNSUPPORTED;
    Geo geo;
    geo.H = H; geo.W = W; geo.nww = W / WS; geo.nwh = H / WS; geo.nW = geo.nwh * geo.nww; geo.shift = shift;
    geo.total_tiles = (long long)B * geo.nW / 2;
    Maps in_maps, out_maps;
    CUtensorMap w_map;
    const long long wd[2] = {C, 3LL * C}, wsd[1] = {C};
    const int wb[2] = {64, 64};
    if (!make_maps(&in_maps, x, B, H, W, C, shift, false) || !make_maps(&out_maps, out, B, H, W, C, shift, true) ||
        !tma::make_map_bf16(&w_map, w, 2, wd, wsd, wb)) return SODT_ERR_CUDA;
    Params p{};
    p.table = static_cast<const float*>(table_ws); p.ln_stats = ln_stats; p.colsum = colsum; p.bias = bias;
    p.M = (long long)B * H * W; p.ln_boxes = ln_boxes; p.ln_inv_k = 1.f / (float)C; p.ln_eps = ln_eps;
    p.C = C; p.heads = heads; p.scale = scale; p.mask_value = mask_value;
    const size_t smem = smem_bytes(heads);
    // as in window_attn_win8: a grid size coprime to the tiles per image spreads the (slower) wrapped border windows over the CTAs
    auto gcd = [](long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; };
    int grid = (int)(geo.total_tiles < num_sms ? geo.total_tiles : num_sms);
    if (shift > 0 && geo.total_tiles > num_sms)
        while (grid > 1 && gcd(grid, geo.nW / 2) != 1) --grid;
    auto kern = shift > 0 ? attn_block_kernel<16, true> : attn_block_kernel<16, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    e = launch_pdl(kern, dim3(grid), dim3(NTHREADS), smem, stream, true, in_maps, out_maps, w_map, p, geo);
    if (e != cudaSuccess) return cuda_status(e);
    return check_launch();
}

}  // namespace sodt
