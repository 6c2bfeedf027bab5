// 8x8-window attention (N = 64 tokens, head_dim 16 or 32) on tcgen05 tensor cores with TMA-fed operand tiles: the kernel of
// the backbone's stage 1 and stage 2 (basics/models/backbone_vit.py:114-145; shift 0 or any 0 < shift < 8).
//
// Window partition, cyclic roll, reverse partition and reverse roll (reference backbone_vit.py:619-672,1096,1118) are TMA
// tile addressing: a window of the rolled frame is the (64 channels, 8, 8) box of the [B*H, W, 3C] qkv image at
// (x, y) = (8 wx + shift, 8 wy + shift), landing in shared memory as a SWIZZLE_128B operand tile (64 token rows x 128 B).  The
// windows of the last window row / column wrap around the image: their boxes are loaded per image row ((64, 8, 1), or the two
// parts (64, 8 - shift, 1) + (64, shift, 1) when the row itself wraps) into the same tile, the 40 / 80 small copies of a stage
// spread over the lanes of the producer warp.  The output tile goes back the same way with TMA stores.  No thread of the kernel
// touches q, k, v or o in global memory.
//
// An MMA instruction of these shapes costs ~100 cycles of the tensor pipe whatever N is (64 with independent accumulators;
// tests/probes/umma_probe.cu), so the kernel is organised around the FEWEST instructions: the 128 TMEM lanes hold TWO HEADS of
// one window (lane = 64 * head parity + query token) and every MMA is unmasked.
// Stage = one window x 64 channels (4 heads of 16 or 2 heads of 32 channels): three boxes K, Q, V = 24 KB, seven stages in
// flight (a TMA load takes ~2700 cycles from issue to data under load: the depth of this ring, not the softmax, bounded the
// kernel until it reached 5+ stages).  Per head pair
//   S[128x64]   two lane-masked tcgen05.mma per 16 channels (SS, M=128, N=64): lanes 0-63 = even head, lanes 64-127 = odd head.
//               Both read the SAME Q tile: the odd head's A operand starts one tile below Q at the odd head's column offset, so its
//               rows 64-127 are Q's tokens (its rows 0-63, the K tile, feed disabled lanes)
//   softmax     one thread per (head, query row): tcgen05.ld of its 64 scores, relative position bias (closed-form index
//               into a shared-memory table laid out for 8-byte loads), shifted-window mask from two 64-bit region masks
//               (border windows only), packed fp32x2 arithmetic, exp2 on the MUFU; P goes back to TMEM as packed bf16
//   O[128x2hd]  ONE unmasked TS chain (A = P from TMEM, B = the two heads' adjacent V channels, MN-major SWIZZLE_128B at the
//               pair's column offset, K = 64 keys); lane half p reads its head's hd columns
//   epilogue    O / rowsum -> bf16 -> SWIZZLE_128B staging tile -> TMA store of the un-rolled image tile
// Persistent CTAs (one per SM), 20 warps: two softmax groups of 8 warps (two threads per score row) taking the units
// (stage, head pair) alternately, with private S / P / O columns in TMEM; one MMA-issuing warp PER GROUP; one TMA-load warp;
// one TMA-store warp.
#include "common.cuh"
#include "tma.cuh"

namespace sodt {
namespace {

using namespace tc;

// Timeline instrumentation for tests/probes/win8_trace.cu (compiled only there): clock64 at the hand-offs of CTA 0.
#ifdef SODT_WIN8_TRACE
__device__ long long g_trace[4][128][12];
#define TRACE(role, unit, ev) do { if (blockIdx.x == 0 && (unit) < 128) g_trace[role][unit][ev] = clock64(); } while (0)
#else
#define TRACE(role, unit, ev) do { } while (0)
#endif

constexpr int WS = 8;
constexpr int NTOK = 64;                 // tokens per window
constexpr int ROWS = 128;                // TMEM lanes = 2 heads x 64 tokens
constexpr int NG = 2;                    // softmax groups
#ifndef SODT_WIN8_TPR
#define SODT_WIN8_TPR 1
#endif
constexpr int TPR = SODT_WIN8_TPR;       // threads per score row (1 or 2): 2 = warps w and w + 4 of a group split the 64 keys
constexpr int KPT = NTOK / TPR;          // keys per thread
constexpr int SM_WARPS = 4 * TPR;        // softmax warps per group
constexpr int NTHREADS = (NG * SM_WARPS + NG + 2) * 32;
constexpr int MMA_WARP0 = NG * SM_WARPS, TMA_WARP = MMA_WARP0 + NG, STORE_WARP = TMA_WARP + 1;   // one MMA-issuing warp per softmax group
constexpr int STAGES = TPR == 1 ? 7 : 6;       // 24 KB each; two threads per row spend 16 KB on the exchange buffer instead
constexpr int WIN_BYTES = NTOK * 128;         // one box: 64 token rows x 128 B
constexpr int STAGE_BYTES = 3 * WIN_BYTES;    // K Q V (Q is NOT first: the odd heads' A operand starts one tile below it, see the MMA issuer)
constexpr int OFF_K = 0, OFF_Q = WIN_BYTES, OFF_V = 2 * WIN_BYTES;
constexpr int OT_BYTES = WIN_BYTES;           // output staging tile of a stage
constexpr int OT_RING = 4;
constexpr float LOG2E = 1.4426950408889634f;
// Bias table in shared memory: 2 copies (copy c is shifted left by c entries) of [head][15 rows dy][16] floats holding the
// x-REVERSED table row, R[dy][r] = table[dy][14 - r], so that the 8 biases of one key row are the ascending entries
// r = 7 - tx + xj and start 8-byte aligned in copy (7 - tx) % 2: four LDS.64 per key row instead of eight LDS.32.
constexpr int TAB_ROW = 16, TAB_HEAD = (2 * WS - 1) * TAB_ROW;          // 240 floats per head and copy
__host__ __device__ constexpr int tab_copy_stride(int heads) {         // floats; bytes = 32 (mod 128): the 16 lanes of a half
    return ((heads * TAB_HEAD * 4 + 95) / 128 * 128 + 32) / 4;          // warp (2 rows x 2 copies x 4 offsets) hit 16 distinct bank pairs
}
constexpr int TAB_COPIES = 2;
// TMEM columns per group g and head pair pr: S = g*128 + pr*64, P = 256 + g*64 + pr*32, O = 384 + g*64 + pr*2hd
constexpr uint32_t TM_S = 0, TM_P = 256, TM_O = 384;
constexpr uint32_t ALL = 0xFFFFFFFFu;
constexpr int XCH_BAR0 = 2;                   // named barriers 2, 3: row-maximum exchange inside a softmax group
constexpr int XCH_FLOATS = TPR == 2 ? 2 * 2 * 2 * NG * 2 * ROWS : 0;    // {max, sum} x unit parity x head pair x group x half x row

__global__ void prep_table_win8_kernel(const float* __restrict__ table, float* __restrict__ out, int heads) {
    const int cs = tab_copy_stride(heads);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= TAB_COPIES * cs) return;
    const int c = e / cs, rem = e - c * cs;
    float v = 0.f;
    if (rem < heads * TAB_HEAD) {
        const int h = rem / TAB_HEAD, r2 = rem - h * TAB_HEAD;
        const int dy = r2 / TAB_ROW, r = r2 - dy * TAB_ROW + c;
        if (r <= 2 * WS - 2) v = table[(long long)(dy * (2 * WS - 1) + (2 * WS - 2 - r)) * heads + h] * LOG2E;
    }
    out[e] = v;
}

struct Geo {
    int H, W, nww, nwh, nW, shift;
    long long total_windows;
};

// Source / destination geometry of one window: first image row / column of its box and whether it wraps around the image
struct WinBox {
    int x0, y0, yg_base;       // x0, y0 in the un-rolled image; yg_base = b * H
    bool wrap_x, wrap_y, last_row, last_col;
};
__device__ __forceinline__ WinBox win_box(const Geo& g, long long wdx) {
    WinBox r;
    const int b = (int)(wdx / g.nW);
    const int win = (int)(wdx - (long long)b * g.nW);
    const int wy = win / g.nww, wx = win - wy * g.nww;
    r.x0 = wx * WS + g.shift;
    r.y0 = wy * WS + g.shift;
    r.wrap_x = r.x0 + WS > g.W;
    r.wrap_y = r.y0 + WS > g.H;
    r.last_row = wy == g.nwh - 1;
    r.last_col = wx == g.nww - 1;
    r.yg_base = b * g.H;
    return r;
}

struct Maps {
    CUtensorMap full, row8, row_a, row_b;     // boxes (64, 8, 8), (64, 8, 1), (64, 8 - shift, 1), (64, shift, 1)
};

// The three boxes of a stage (K, Q, V), issued by the whole producer warp: a window inside the image is one box per
// operand tile (lanes 0-2); a window that wraps around the image is 8 row boxes per tile, or 16 row parts when the rows
// themselves wrap (24 / 48 small copies spread over the 32 lanes: one lane needs ~130 cycles per TMA instruction).
template <int HD>
__device__ __forceinline__ void load_stage(const Maps& m, const Geo& geo, const WinBox& b, uint32_t st, int c0, int C, uint64_t* bar, int lane) {
    const int per_box = !b.wrap_x && !b.wrap_y ? 1 : (b.wrap_x ? 2 * WS : WS);
    for (int item = lane; item < 3 * per_box; item += 32) {
        const int bi = item / per_box, sub = item - bi * per_box;
        const uint32_t sm = st + bi * WIN_BYTES;                                 // tiles in the order K Q V
        const int ch = (bi == 0 ? C : bi == 1 ? 0 : 2 * C) + c0;
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

// Output tile -> the window's pixels of the un-rolled image, issued by the whole store warp: one box, or the 8 rows / 16 row
// parts of a wrapped window spread over the lanes.  (Storing a wrapped tile as full boxes at negative coordinates and letting the
// TMA unit clip them is not an option: a tiled store with a negative coordinate raises an illegal-instruction fault, probed.)
__device__ __forceinline__ void store_tile(const Maps& m, const Geo& geo, const WinBox& b, uint32_t sm, int c0, int lane) {
    if (!b.wrap_x && !b.wrap_y) {
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

template <int HD>
__global__ void __launch_bounds__(NTHREADS, 1)
window_attn_win8_kernel(const __grid_constant__ Maps in_maps, const __grid_constant__ Maps out_maps,
                        const float* __restrict__ table_p, Geo geo, int C, int heads, float scale, float mask_value) {
    constexpr int G = 64 / HD;                // heads per stage
    constexpr int HPB = G / 2;                // head pairs per stage = per unit: 2 (head_dim 16) or 1 (head_dim 32)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t stage_full[STAGES], stage_empty[STAGES], s_full[NG], s_free[NG], p_full[NG], pv_done[NG], ot_full[OT_RING], ot_free[OT_RING];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ot_base = sbase + STAGES * STAGE_BYTES;                            // ring of output staging tiles
    float* xch = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)) + STAGES * STAGE_BYTES + OT_RING * OT_BYTES);
    float* tab = xch + XCH_FLOATS;
    const int groups = C / 64;                                                         // stages per window
    int my_windows = 0;
    if ((long long)blockIdx.x < geo.total_windows) my_windows = (int)((geo.total_windows - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int n_stages = my_windows * groups;         // unit = stage; group g takes this CTA's stages g, g + NG, ...

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&stage_full[s], 1); mbar_init(&stage_empty[s], 1); }
        for (int s = 0; s < OT_RING; ++s) { mbar_init(&ot_full[s], TPR * ROWS); mbar_init(&ot_free[s], 1); }
        for (int g = 0; g < NG; ++g) { mbar_init(&s_full[g], 1); mbar_init(&s_free[g], TPR * ROWS); mbar_init(&p_full[g], TPR * ROWS); mbar_init(&pv_done[g], 1); }
        fence_barrier_init();
    }
    if (warp == MMA_WARP0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    pdl_trigger();
    pdl_wait();                    // programmatic dependent launch (common.cuh): no global memory is touched above
    {
        const int n4 = TAB_COPIES * tab_copy_stride(heads) / 4;
        const float4* src = reinterpret_cast<const float4*>(table_p);
        float4* dst = reinterpret_cast<float4*>(tab);
        for (int e = tid; e < n4; e += NTHREADS) dst[e] = src[e];
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_slot;

    if (warp == TMA_WARP) {
        // ============================================================ producer: the warp issues the boxes of every stage
        if (lane == 0) tma::prefetch_map(&in_maps.full);
        int stage = 0, round = 0;
        [[maybe_unused]] int tr_stage = 0;
        for (long long wdx = blockIdx.x; wdx < geo.total_windows; wdx += gridDim.x) {
            const WinBox b = win_box(geo, wdx);
            for (int gi = 0; gi < groups; ++gi) {
                if (lane == 0) {
                    TRACE(3, tr_stage, 0);
                    if (round > 0) mbar_wait(&stage_empty[stage], (uint32_t)((round - 1) & 1));
                    TRACE(3, tr_stage, 1);
                    tma::expect_tx(&stage_full[stage], STAGE_BYTES);
                }
                __syncwarp();
                load_stage<HD>(in_maps, geo, b, sbase + stage * STAGE_BYTES, gi * 64, C, &stage_full[stage], lane);
                if (lane == 0) TRACE(3, tr_stage, 2);
                ++tr_stage;
                if (++stage == STAGES) { stage = 0; ++round; }
            }
        }
    } else if (warp == STORE_WARP) {
        // ============================================================ output stores: the warp stores every finished staging tile
        if (lane == 0) tma::prefetch_map(&out_maps.full);
        int st = 0;
        for (long long wdx = blockIdx.x; wdx < geo.total_windows; wdx += gridDim.x) {
            const WinBox b = win_box(geo, wdx);
            for (int gi = 0; gi < groups; ++gi, ++st) {
                const int slot = st & (OT_RING - 1);
                if (lane == 0) mbar_wait(&ot_full[slot], (uint32_t)((st / OT_RING) & 1));        // every head's columns are in the tile
                __syncwarp();
                store_tile(out_maps, geo, b, ot_base + slot * OT_BYTES, gi * 64, lane);
                tma::store_commit();                                                     // per lane: each lane tracks its own copies
                if (st > 0) {                                                            // the previous tile has been read: its slot is free
                    tma::store_wait_read<1>();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ot_free[(st - 1) & (OT_RING - 1)]);
                }
            }
        }
        tma::store_wait_all();
    } else if (warp >= MMA_WARP0) {
        // =============================================================== MMA issuers: one thread per softmax group
        // (a single issuer serialised the groups: the scores of one group waited behind the other group's P)
        if (lane == 0) {
            constexpr uint32_t idesc_s = idesc_bf16(ROWS, NTOK, false, false);
            constexpr uint32_t idesc_o = idesc_bf16(ROWS, 2 * HD, false, true);
            const int g = warp - MMA_WARP0;
            const uint64_t d0 = tma::desc_sw128(sbase);
            const uint32_t tS = tm + TM_S + g * 128, tP = tm + TM_P + g * 64, tO = tm + TM_O + g * 64;
            const int nk = n_stages > g ? (n_stages - g + NG - 1) / NG : 0;          // unit k of this group = stage NG k + g
            auto issue_qk = [&](int k) {
                const int stg = NG * k + g, slot = stg % STAGES;
                TRACE(2, stg, 0);
                mbar_wait_spin(&stage_full[slot], (uint32_t)((stg / STAGES) & 1));
                TRACE(2, stg, 1);
                if (k > 0) mbar_wait_spin(&s_free[g], (uint32_t)((k - 1) & 1));
                TRACE(2, stg, 2);
                fence_after_sync();
#pragma unroll
                for (int pr = 0; pr < HPB; ++pr) {
                    // lanes 0-63 (even head): A = the 128 rows from the Q tile on, at the even head's columns (rows 64-127 = the V
                    // tile: their output lanes are disabled, so their content does not matter), B = K at the even head's columns;
                    // lanes 64-127 (odd head): A starts one tile BELOW Q, at the odd head's columns, so that its rows 64-127 are the
                    // Q tile's 64 tokens (rows 0-63 = the K tile, disabled lanes), B = K at the odd head's columns.
                    // No second, channel-shifted copy of Q is needed.
                    const uint32_t off = (uint32_t)(slot * STAGE_BYTES + pr * (2 * HD * 2)) >> 4;
                    const uint64_t qd = d0 + off + (OFF_Q >> 4), kd = d0 + off + (OFF_K >> 4);
                    const uint64_t qd_odd = qd - (WIN_BYTES >> 4) + ((HD * 2) >> 4), kd_odd = kd + ((HD * 2) >> 4);
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks) mma_ss_masked(tS + pr * 64, qd + 2 * ks, kd + 2 * ks, idesc_s, ks > 0, 0u, 0u, ALL, ALL);
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks) mma_ss_masked(tS + pr * 64, qd_odd + 2 * ks, kd_odd + 2 * ks, idesc_s, ks > 0, ALL, ALL, 0u, 0u);
                }
                mma_commit(&s_full[g]);
                TRACE(2, stg, 3);
            };
            if (nk > 0) issue_qk(0);
            for (int k = 0; k < nk; ++k) {
                // S is free as soon as softmax(k) has copied it to registers, so the group's next scores are computed while
                // softmax(k) is still running
                if (k + 1 < nk) issue_qk(k + 1);
                const int stg = NG * k + g, slot = stg % STAGES;
                TRACE(2, stg, 4);
                mbar_wait_spin(&p_full[g], (uint32_t)(k & 1));
                TRACE(2, stg, 5);
                fence_after_sync();
#pragma unroll
                for (int ks = 0; ks < NTOK / 16; ++ks) {        // O[128 x 2hd] = P [V_even | V_odd], 16 keys (2048 B of the V tile) per step;
#pragma unroll
                    for (int pr = 0; pr < HPB; ++pr) {          // the pairs' chains interleaved (independent accumulators)
                        const uint64_t vd = d0 + ((uint32_t)(slot * STAGE_BYTES + OFF_V + pr * (2 * HD * 2)) >> 4);
                        mma_ts(tO + pr * (2 * HD), tP + pr * 32 + ks * 8, vd + ks * (2048 >> 4), idesc_o, ks > 0);
                    }
                }
                mma_commit(&pv_done[g]);
                mma_commit(&stage_empty[slot]);                  // the stage's tiles are no longer read
                TRACE(2, stg, 6);
            }
        }
    } else {
        // ====================================================== softmax + epilogue groups
        // Two threads per score row: warps w and w + 4 of a group own the same 32 TMEM lanes (a warp reaches lane quarter
        // warp % 4 only) and split the 64 keys, so 16 softmax warps (4 per SM sub-partition) cover one another's tensor-memory,
        // shared-memory and MUFU latencies; row maximum and row sum are exchanged through shared memory.
        const int g = warp / SM_WARPS;                 // group g takes stages g, g + NG, ...
        const int half = TPR == 2 ? (warp >> 2) & 1 : 0;    // keys [KPT half, KPT half + KPT) = key rows yj from (KPT / 8) half
        const int row = (warp & 3) * 32 + lane;        // TMEM lane = 64 * (head parity) + query token
        const int hp = row >> 6, ti = row & 63, ty = ti >> 3, tx = ti & 7;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tm + TM_S + g * 128 + half * KPT + lane_addr;
        const uint32_t tP = tm + TM_P + g * 64 + half * (KPT / 2) + lane_addr;
        const uint32_t tO = tm + TM_O + g * 64 + hp * HD + half * (HD / TPR) + lane_addr;
        const float c = scale * LOG2E, mv2 = mask_value * LOG2E;
        const uint64_t c2 = pack2(c, c);
        // bias row of this thread's first key row: copy (7 - tx) % 2 at entry (7 - tx) - copy (even), table row dy = ty + 7 - yj
        const int r0 = WS - 1 - tx, cp = r0 & 1;
        const float* tab_row = tab + cp * tab_copy_stride(heads) + (ty + WS - 1 - (KPT / 8) * half) * TAB_ROW + (r0 - cp);
        const int s_ = geo.shift;
        const uint64_t yhi = s_ > 0 ? (~0ull << (8 * (WS - s_))) : 0ull;                       // keys with ty >= ws - shift
        const uint64_t xhi = s_ > 0 ? 0x0101010101010101ull * (uint64_t)((0xFFu << (WS - s_)) & 0xFFu) : 0ull;  // tx >= ws - shift
        const uint32_t srow = (uint32_t)ti * 128, sw = (uint32_t)(ti & 7);
        // exchange slots [parity][pair][group][half][row]: this thread's and its partner's (the other half of the same row)
        float* my_max = xch + (g * 2 + half) * ROWS + row;
        float* pt_max = xch + (g * 2 + (half ^ 1)) * ROWS + row;
        float* my_sum = my_max + XCH_FLOATS / 2;
        float* pt_sum = pt_max + XCH_FLOATS / 2;
        constexpr int XPAIR = NG * 2 * ROWS, XPAR = 2 * XPAIR;      // strides of the pair / parity dimensions
        uint64_t mbits = 0;
        bool any_mask = false;
        float prev_sum[HPB] = {};
        int prev_stg = 0, prev_par = 0;

        // Epilogue of the group's previous unit: this thread's half of its heads' O columns / rowsum -> staging tile of the stage;
        // the store warp sends the tile off when all 256 threads of the group have delivered.
        auto epilogue = [&]() {
            const int slot = prev_stg & (OT_RING - 1);
            const uint32_t tile_s = ot_base + (uint32_t)slot * OT_BYTES;
            constexpr int OC = HD / TPR;                  // O columns per thread
            uint32_t o[HPB][OC];
            float inv[HPB];
#pragma unroll
            for (int pr = 0; pr < HPB; ++pr) {
                if constexpr (OC == 8) tmem_ld8(tO + pr * (2 * HD), o[pr]);
                else if constexpr (OC == 16) tmem_ld16(tO + pr * (2 * HD), o[pr]);
                else tmem_ld32(tO + pr * (2 * HD), o[pr]);
                float total = prev_sum[pr];
                if constexpr (TPR == 2) total += pt_sum[prev_par * XPAR + pr * XPAIR];      // written before the partner's p_full arrival
                inv[pr] = fast_rcp(total);
            }
            tmem_wait_ld();
            if (prev_stg >= OT_RING) mbar_wait(&ot_free[slot], (uint32_t)(((prev_stg / OT_RING) - 1) & 1));   // the slot's previous tile has left
#pragma unroll
            for (int pr = 0; pr < HPB; ++pr) {
#pragma unroll
                for (int j = 0; j < OC; j += 8) {
                    const uint32_t chunk = (uint32_t)(((2 * pr + hp) * HD + half * OC + j) >> 3) ^ sw;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_s + srow + (chunk << 4)),
                                 "r"(pack_bf16(__uint_as_float(o[pr][j]) * inv[pr], __uint_as_float(o[pr][j + 1]) * inv[pr])),
                                 "r"(pack_bf16(__uint_as_float(o[pr][j + 2]) * inv[pr], __uint_as_float(o[pr][j + 3]) * inv[pr])),
                                 "r"(pack_bf16(__uint_as_float(o[pr][j + 4]) * inv[pr], __uint_as_float(o[pr][j + 5]) * inv[pr])),
                                 "r"(pack_bf16(__uint_as_float(o[pr][j + 6]) * inv[pr], __uint_as_float(o[pr][j + 7]) * inv[pr])) : "memory");
                }
            }
            fence_proxy_async();                                          // staging writes -> visible to the TMA store
            mbar_arrive(&ot_full[slot]);
        };

        int k = 0, gi = g % groups, win_it = g / groups, cur_win_it = -1;     // stage stg = NG k + g = (window iteration, channel group)
        for (int stg = g; stg < n_stages; stg += NG, ++k) {
            if (win_it != cur_win_it) {                // new window: shifted-window mask bits of this thread's 32 keys
                cur_win_it = win_it;
                const WinBox wb = win_box(geo, (long long)blockIdx.x + (long long)win_it * gridDim.x);
                uint64_t mb = 0;
                if (s_ > 0) {
                    if (wb.last_row) mb |= (ty >= WS - s_) ? ~yhi : yhi;
                    if (wb.last_col) mb |= (tx >= WS - s_) ? ~xhi : xhi;
                }
                mbits = TPR == 2 ? (mb >> (32 * half)) & 0xFFFFFFFFull : mb;
                any_mask = s_ > 0 && (wb.last_row || wb.last_col);          // uniform over the group
            }
            const int par = k & 1;
            if (row == 0 && half == 0) TRACE(g, stg, 0);
            mbar_wait(&s_full[g], (uint32_t)par);
            if (row == 0 && half == 0) TRACE(g, stg, 1);
            fence_after_sync();
            float sum_cur[HPB];
            bool epi_done = k == 0;
#pragma unroll
            for (int pr = 0; pr < HPB; ++pr) {
                const int h = gi * G + 2 * pr + hp;
                uint64_t t[KPT / 2];
                uint32_t pk[KPT / 2];
#pragma unroll
                for (int part = 0; part < KPT / 32; ++part) {
                    uint32_t ra[32];
                    tmem_ld32(tS + pr * 64 + part * 32, ra);
                    tmem_wait_ld();
                    if (pr == HPB - 1 && part == KPT / 32 - 1) { fence_before_sync(); mbar_arrive(&s_free[g]); }   // the unit's scores are in registers
                    const float* tb = tab_row + h * TAB_HEAD - part * 4 * TAB_ROW;
#pragma unroll
                    for (int yj = 0; yj < 4; ++yj) {        // t = s * (scale log2 e) + bias, two scores per FFMA2
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
#ifdef SODT_X_NOBIAS
                            const float2 b = make_float2(0.f, 0.f);
#else
                            const float2 b = *reinterpret_cast<const float2*>(tb - yj * TAB_ROW + 2 * q);
#endif
                            t[part * 16 + yj * 4 + q] = ffma2(pack2(__uint_as_float(ra[yj * 8 + 2 * q]), __uint_as_float(ra[yj * 8 + 2 * q + 1])), c2, pack2(b.x, b.y));
                        }
                    }
                }
                if (any_mask) {                            // group-uniform: only windows of the last window row / column
#pragma unroll
                    for (int j = 0; j < KPT / 2; ++j) {
                        float lo, hi;
                        unpack2(t[j], lo, hi);
                        if ((mbits >> (2 * j)) & 1ull) lo += mv2;
                        if ((mbits >> (2 * j + 1)) & 1ull) hi += mv2;
                        t[j] = pack2(lo, hi);
                    }
                }
                float m4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float lo, hi;
                    unpack2(t[q], lo, hi);
                    m4[q] = fmaxf(lo, hi);
                }
#pragma unroll
                for (int j = 4; j < KPT / 2; ++j) {
                    float lo, hi;
                    unpack2(t[j], lo, hi);
                    m4[j & 3] = fmax3(m4[j & 3], lo, hi);
                }
                float mx = fmaxf(fmax3(m4[0], m4[1], m4[2]), m4[3]);
                if constexpr (TPR == 2) {
                    my_max[par * XPAR + pr * XPAIR] = mx;
                    asm volatile("bar.sync %0, 256;" ::"r"(XCH_BAR0 + g) : "memory");      // the two halves of every row have published their maxima
                    mx = fmaxf(mx, pt_max[par * XPAR + pr * XPAIR]);
                }
                const uint64_t nmx2 = pack2(-mx, -mx);
                uint64_t sum2 = 0ull;
#pragma unroll
                for (int j = 0; j < KPT / 2; ++j) {
                    float lo, hi;
                    unpack2(fadd2(t[j], nmx2), lo, hi);
#ifdef SODT_X_NOEXP
                    const float p0 = lo * 0.001f + 1.f, p1 = hi * 0.001f + 1.f;
#else
                    const float p0 = fast_exp2(lo), p1 = fast_exp2(hi);
#endif
                    sum2 = fadd2(sum2, pack2(p0, p1));
                    pk[j] = pack_bf16(p0, p1);
                }
                float a, b;
                unpack2(sum2, a, b);
                sum_cur[pr] = a + b;
                if (!epi_done) {                       // previous unit of this group: its P / O columns are free again
                    epi_done = true;
                    mbar_wait(&pv_done[g], (uint32_t)((k - 1) & 1));
                    fence_after_sync();
                    epilogue();
                }
                tmem_st(tP + pr * 32, pk);
            }
            if (row == 0 && half == 0) TRACE(g, stg, 3);
#pragma unroll
            for (int pr = 0; pr < HPB; ++pr) {
                if constexpr (TPR == 2) my_sum[par * XPAR + pr * XPAIR] = sum_cur[pr];   // the partner reads it in its epilogue, after the next exchange barrier
                prev_sum[pr] = sum_cur[pr];
            }
            tmem_wait_st();
            fence_before_sync();
            mbar_arrive(&p_full[g]);
            if (row == 0 && half == 0) TRACE(g, stg, 6);
            prev_stg = stg; prev_par = par;
            gi += NG;
            while (gi >= groups) { gi -= groups; ++win_it; }
        }
        if (k > 0) {
            mbar_wait(&pv_done[g], (uint32_t)((k - 1) & 1));
            fence_after_sync();
            if constexpr (TPR == 2) asm volatile("bar.sync %0, 256;" ::"r"(XCH_BAR0 + g) : "memory");      // the partner's last row sums are visible
            epilogue();
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP0) tmem_dealloc(tmem_slot, 512);
}

size_t win8_table_bytes(int heads) { return (size_t)TAB_COPIES * tab_copy_stride(heads) * sizeof(float); }

size_t win8_smem_bytes(int heads) { return (size_t)STAGES * STAGE_BYTES + OT_RING * OT_BYTES + XCH_FLOATS * sizeof(float) + win8_table_bytes(heads) + 1024; }

bool make_maps(Maps* m, const void* base, int B, int H, int W, int Cfull, int shift, bool is_output) {
    const long long dims[3] = {Cfull, W, (long long)B * H}, strides[2] = {Cfull, (long long)W * Cfull};
    const int sa = shift > 0 ? WS - shift : WS, sb = shift > 0 ? shift : WS;
    const int bf[3] = {64, WS, WS}, b8[3] = {64, WS, 1}, ba[3] = {64, sa, 1}, bb[3] = {64, sb, 1};
    const CUtensorMapL2promotion promo = is_output ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    return tma::make_map_bf16(&m->full, base, 3, dims, strides, bf, promo) && tma::make_map_bf16(&m->row8, base, 3, dims, strides, b8, promo) &&
           tma::make_map_bf16(&m->row_a, base, 3, dims, strides, ba, promo) && tma::make_map_bf16(&m->row_b, base, 3, dims, strides, bb, promo);
}

}  // namespace

size_t window_attn_win8_workspace(int heads) { return win8_table_bytes(heads); }

bool window_attn_win8_supported(int H, int W, int C, int heads, int ws, int shift, int dtype) {
    if (dtype != SODT_BF16 || ws != WS || H % WS || W % WS || heads <= 0 || C % heads || C % 64) return false;
    const int hd = C / heads;
    if (hd != 16 && hd != 32) return false;
    return win8_smem_bytes(heads) <= 227 * 1024;
}

// Writes the shared-memory image of the bias table (2 shifted, x-reversed, log2(e)-scaled copies) into `workspace`
int window_attn_win8_prepare(const float* table, void* workspace, int heads, cudaStream_t stream) {
    const int n = TAB_COPIES * tab_copy_stride(heads);
    prep_table_win8_kernel<<<(n + 255) / 256, 256, 0, stream>>>(table, static_cast<float*>(workspace), heads);
    return check_launch();
}

int window_attn_win8(const void* qkv, const float* table, void* out, void* workspace, int B, int H, int W, int C,
                     int heads, int shift, float scale, float mask_value, int num_sms, bool prepared, cudaStream_t stream) {
    if (!prepared) {
        const int st = window_attn_win8_prepare(table, workspace, heads, stream);
        if (st != SODT_OK) return st;
    }
    if ((long long)B * H > 2147483647LL) return SODT_ERR_UNSUPPORTED;
    Geo geo;
    geo.H = H; geo.W = W; geo.nww = W / WS; geo.nwh = H / WS; geo.nW = geo.nwh * geo.nww; geo.shift = shift;
    geo.total_windows = (long long)B * geo.nW;
    Maps in_maps, out_maps;
    if (!make_maps(&in_maps, qkv, B, H, W, 3 * C, shift, false) || !make_maps(&out_maps, out, B, H, W, C, shift, true)) return SODT_ERR_CUDA;
    const size_t smem = win8_smem_bytes(heads);
    const int hd = C / heads;
    // CTA c takes windows c, c + grid, ...: with gcd(grid, windows per image) > 1 the (slower) wrapped windows of the last window
    // column pile up on a few CTAs (148 and 32 x 32 share the factor 4: a quarter of the CTAs got all of them, +50 % run time
    // at 16 x 16 windows).  A grid size coprime to the window count spreads them evenly; it costs at most a few idle SMs.
    auto gcd = [](long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; };
    int grid = (int)(geo.total_windows < num_sms ? geo.total_windows : num_sms);
    if (shift > 0 && geo.total_windows > num_sms)
        while (grid > 1 && gcd(grid, geo.nW) != 1) --grid;
    cudaError_t e;
    if (hd == 16) {
        auto kern = window_attn_win8_kernel<16>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
        e = launch_pdl(kern, dim3(grid), dim3(NTHREADS), smem, stream, true, in_maps, out_maps, static_cast<const float*>(workspace), geo, C, heads, scale, mask_value);
    } else {
        auto kern = window_attn_win8_kernel<32>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
        e = launch_pdl(kern, dim3(grid), dim3(NTHREADS), smem, stream, true, in_maps, out_maps, static_cast<const float*>(workspace), geo, C, heads, scale, mask_value);
    }
    if (e != cudaSuccess) return cuda_status(e);
    return check_launch();
}

}  // namespace sodt
