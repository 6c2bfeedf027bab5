// 8x8-window attention (N = 64 tokens, head_dim 16 or 32) on tcgen05 tensor cores with TMA-fed operand tiles: the kernel of
// the backbone's stage 1 and stage 2 (basics/models/backbone_vit.py:114-145; shift 0 or any 0 < shift < 8).
//
// Window partition, cyclic roll, reverse partition and reverse roll (reference backbone_vit.py:619-672,1096,1118) are TMA
// tile addressing: a window of the rolled frame is the (64 channels, 8, 8) box of the [B*H, W, 3C] qkv image at
// (x, y) = (8 wx + shift, 8 wy + shift), landing in shared memory as a SWIZZLE_128B operand tile (64 token rows x 128 B).  The
// windows of the last window row / column wrap around the image: their boxes are issued per image row ((64, 8, 1), or the two
// parts (64, 8 - shift, 1) + (64, shift, 1) when the row itself wraps), into the same tile.  The output tile goes back the
// same way with TMA stores.  No thread of the kernel touches q, k, v or o in global memory.
//
// An MMA instruction of these shapes costs ~100 cycles of the tensor pipe whatever N is (64 with independent accumulators;
// tests/probes/umma_probe.cu), so the kernel is organised around the FEWEST instructions: the 128 TMEM lanes hold TWO HEADS of
// one window (lane = 64 * head parity + query token) and every MMA is unmasked.
// Stage = one window x 64 channels (4 heads of 16 or 2 heads of 32 channels): five boxes, 40 KB
//   QA, KA   q / k channels [c0, c0 + 64) of the window's tokens           QB, KB   the same tokens, channels [c0 + hd, c0 + hd + 64)
//   V        v channels [c0, c0 + 64)
// so that an operand of 128 rows starting in QA (KA) at the even head's column offset continues in QB (KB) with the ODD head's
// channels at the same column offset (the second fetch of the same lines is served by L2).  Per head pair
//   S[128x128]  ONE tcgen05.mma per 16 channels (SS, M=128, N=128): lane (p, i), column (p', j) = q_{head p}(i) . k_{head p'}(j);
//               lane half p reads its own 64 columns [64 p, 64 p + 64); the cross-head half is discarded
//   softmax     one thread per (head, query row): tcgen05.ld of its 64 scores, relative position bias (closed-form index
//               into a shared-memory table laid out for 8-byte loads), shifted-window mask from two 64-bit region masks
//               (border windows only), packed fp32x2 arithmetic, exp2 on the MUFU; P goes back to TMEM as packed bf16
//   O[128x2hd]  ONE unmasked TS chain (A = P from TMEM, B = the two heads' adjacent V channels, MN-major SWIZZLE_128B at the
//               pair's column offset, K = 64 keys); lane half p reads its head's hd columns
//   epilogue    O / rowsum -> bf16 -> SWIZZLE_128B staging tile -> TMA store of the un-rolled image tile
// Persistent CTAs (one per SM), 12 warps: two softmax groups of 4 warps taking the units (stage, head pair) alternately, with
// private S / P / O columns in TMEM; one MMA-issuing warp PER GROUP; one TMA-load warp; one TMA-store warp.
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
constexpr int NTHREADS = (NG * 4 + NG + 2) * 32;
constexpr int MMA_WARP0 = NG * 4, TMA_WARP = NG * 4 + NG, STORE_WARP = NG * 4 + NG + 1;   // one MMA-issuing warp per softmax group
constexpr int STAGES = 4;
constexpr int WIN_BYTES = NTOK * 128;         // one box: 64 token rows x 128 B
constexpr int STAGE_BYTES = 5 * WIN_BYTES;    // QA QB KA KB V
constexpr int OFF_Q = 0, OFF_K = 2 * WIN_BYTES, OFF_V = 4 * WIN_BYTES;
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
// TMEM columns per group g: S = g*128 (128 columns), P = 256 + g*32, O = 320 + g*64 (2*hd <= 64 columns)
constexpr uint32_t TM_S = 0, TM_P = 256, TM_O = 320;
constexpr int EXP_BAR0 = 2;                   // named barriers 2, 3: turn-taking of the two softmax groups' exp phases

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

// One window box (64 channels from c0) <-> the 8 KB tile at shared address `sm`; LOAD = global -> shared (completes on bar)
template <bool LOAD>
__device__ __forceinline__ void window_box(const Maps& m, const Geo& geo, const WinBox& b, uint32_t sm, int c0, uint64_t* bar) {
    if (!b.wrap_x && !b.wrap_y) {
        if (LOAD) tma::load_3d(sm, &m.full, bar, c0, b.x0, b.yg_base + b.y0); else tma::store_3d(&m.full, sm, c0, b.x0, b.yg_base + b.y0);
        return;
    }
    for (int ty = 0; ty < WS; ++ty) {
        int ys = b.y0 + ty; if (ys >= geo.H) ys -= geo.H;
        const int yg = b.yg_base + ys;
        const uint32_t d = sm + ty * 1024;
        if (b.wrap_x) {
            const uint32_t d2 = d + (WS - geo.shift) * 128;
            if (LOAD) { tma::load_3d(d, &m.row_a, bar, c0, b.x0, yg); tma::load_3d(d2, &m.row_b, bar, c0, 0, yg); }
            else { tma::store_3d(&m.row_a, d, c0, b.x0, yg); tma::store_3d(&m.row_b, d2, c0, 0, yg); }
        } else {
            if (LOAD) tma::load_3d(d, &m.row8, bar, c0, b.x0, yg); else tma::store_3d(&m.row8, d, c0, b.x0, yg);
        }
    }
}

__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {     // FFMA2: two fp32 FMAs per issue slot
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {                  // FMNMX3
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

template <int HD>
__global__ void __launch_bounds__(NTHREADS, 1)
window_attn_win8_kernel(const __grid_constant__ Maps in_maps, const __grid_constant__ Maps out_maps,
                        const float* __restrict__ table_p, Geo geo, int C, int heads, float scale, float mask_value) {
    constexpr int G = 64 / HD;                // heads per stage
    constexpr int UPS = G / 2;                // units (head pairs) per stage: 2 (head_dim 16) or 1 (head_dim 32)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t stage_full[STAGES], stage_empty[STAGES], s_full[NG], s_free[NG], p_full[NG], pv_done[NG], ot_full[OT_RING], ot_free[OT_RING];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ot_base = sbase + STAGES * STAGE_BYTES;                            // ring of output staging tiles
    float* tab = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)) + STAGES * STAGE_BYTES + OT_RING * OT_BYTES);
    const int groups = C / 64;                                                         // stages per window
    long long my_windows = 0;
    if ((long long)blockIdx.x < geo.total_windows) my_windows = (geo.total_windows - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const long long n_stages = my_windows * groups;
    const long long n_units = n_stages * UPS;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&stage_full[s], 1); mbar_init(&stage_empty[s], UPS); }     // one commit per head pair of the stage
        for (int s = 0; s < OT_RING; ++s) { mbar_init(&ot_full[s], UPS * ROWS); mbar_init(&ot_free[s], 1); }
        for (int g = 0; g < NG; ++g) { mbar_init(&s_full[g], 1); mbar_init(&s_free[g], ROWS); mbar_init(&p_full[g], ROWS); mbar_init(&pv_done[g], 1); }
        fence_barrier_init();
    }
    if (warp == MMA_WARP0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
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
        // ============================================================ producer: one lane issues every box of every stage
        if (lane == 0) {
            tma::prefetch_map(&in_maps.full);
            int stage = 0, round = 0;
            [[maybe_unused]] int tr_stage = 0;
            for (long long wdx = blockIdx.x; wdx < geo.total_windows; wdx += gridDim.x) {
                const WinBox b = win_box(geo, wdx);
                for (int gi = 0; gi < groups; ++gi) {
                    TRACE(3, tr_stage, 0);
                    if (round > 0) mbar_wait(&stage_empty[stage], (uint32_t)((round - 1) & 1));
                    TRACE(3, tr_stage, 1);
                    tma::expect_tx(&stage_full[stage], STAGE_BYTES);
                    const uint32_t st = sbase + stage * STAGE_BYTES;
                    const int c0 = gi * 64;
                    window_box<true>(in_maps, geo, b, st + OFF_Q, c0, &stage_full[stage]);
                    window_box<true>(in_maps, geo, b, st + OFF_Q + WIN_BYTES, c0 + HD, &stage_full[stage]);
                    window_box<true>(in_maps, geo, b, st + OFF_K, C + c0, &stage_full[stage]);
                    window_box<true>(in_maps, geo, b, st + OFF_K + WIN_BYTES, C + c0 + HD, &stage_full[stage]);
                    window_box<true>(in_maps, geo, b, st + OFF_V, 2 * C + c0, &stage_full[stage]);
                    TRACE(3, tr_stage, 2);
                    ++tr_stage;
                    if (++stage == STAGES) { stage = 0; ++round; }
                }
            }
        }
    } else if (warp == STORE_WARP) {
        // ============================================================ output stores: one lane stores every finished staging tile
        if (lane == 0) {
            tma::prefetch_map(&out_maps.full);
            long long st = 0;
            for (long long wdx = blockIdx.x; wdx < geo.total_windows; wdx += gridDim.x) {
                const WinBox b = win_box(geo, wdx);
                for (int gi = 0; gi < groups; ++gi, ++st) {
                    const int slot = (int)(st & (OT_RING - 1));
                    mbar_wait(&ot_full[slot], (uint32_t)((st / OT_RING) & 1));           // both head pairs' columns are in the tile
                    window_box<false>(out_maps, geo, b, ot_base + slot * OT_BYTES, gi * 64, nullptr);
                    tma::store_commit();
                    if (st > 0) {                                                        // the previous tile has been read: its slot is free
                        tma::store_wait_read<1>();
                        mbar_arrive(&ot_free[(int)((st - 1) & (OT_RING - 1))]);
                    }
                }
            }
            tma::store_wait_all();
        }
    } else if (warp >= MMA_WARP0) {
        // =============================================================== MMA issuers: one thread per softmax group
        // (a single issuer serialised the groups: the scores of one group waited behind the other group's P)
        if (lane == 0) {
            constexpr uint32_t idesc_s = idesc_bf16(ROWS, 2 * NTOK, false, false);
            constexpr uint32_t idesc_o = idesc_bf16(ROWS, 2 * HD, false, true);
            const int g = warp - MMA_WARP0;
            const uint64_t d0 = tma::desc_sw128(sbase);
            const uint32_t tS = tm + TM_S + g * 128, tP = tm + TM_P + g * 32, tO = tm + TM_O + g * 64;
            // unit k of this group = global unit u = NG k + g = (stage u / UPS, head pair u % UPS)
            const long long nk = n_units > g ? (n_units - g + NG - 1) / NG : 0;
            auto issue_qk = [&](long long k) {
                const long long u = NG * k + g, stg = u / UPS;
                const int pr = (int)(u - stg * UPS), slot = (int)(stg % STAGES);
                TRACE(2, (int)u, 0);
                mbar_wait_spin(&stage_full[slot], (uint32_t)((stg / STAGES) & 1));
                TRACE(2, (int)u, 1);
                if (k > 0) mbar_wait_spin(&s_free[g], (uint32_t)((k - 1) & 1));
                TRACE(2, (int)u, 2);
                fence_after_sync();
                const uint32_t off = (uint32_t)(slot * STAGE_BYTES + pr * (2 * HD * 2)) >> 4;
                const uint64_t qd = d0 + off + (OFF_Q >> 4), kd = d0 + off + (OFF_K >> 4);
#pragma unroll
                for (int ks = 0; ks < HD / 16; ++ks)            // [Q_even ; Q_odd] x [K_even ; K_odd]^T, 16 channels per step
                    mma_ss(tS, qd + 2 * ks, kd + 2 * ks, idesc_s, ks > 0);
                mma_commit(&s_full[g]);
                TRACE(2, (int)u, 3);
            };
            if (nk > 0) issue_qk(0);
            for (long long k = 0; k < nk; ++k) {
                // S is free as soon as softmax(k) has copied it to registers, so the group's next scores are computed while
                // softmax(k) is still running
                if (k + 1 < nk) issue_qk(k + 1);
                const long long u = NG * k + g, stg = u / UPS;
                const int pr = (int)(u - stg * UPS), slot = (int)(stg % STAGES);
                TRACE(2, (int)u, 4);
                mbar_wait_spin(&p_full[g], (uint32_t)(k & 1));
                TRACE(2, (int)u, 5);
                fence_after_sync();
                const uint64_t vd = d0 + ((uint32_t)(slot * STAGE_BYTES + OFF_V + pr * (2 * HD * 2)) >> 4);
#pragma unroll
                for (int ks = 0; ks < NTOK / 16; ++ks)          // O[128 x 2hd] = P [V_even | V_odd], 16 keys (2048 B of the V tile) per step
                    mma_ts(tO, tP + ks * 8, vd + ks * (2048 >> 4), idesc_o, ks > 0);
                mma_commit(&pv_done[g]);
                mma_commit(&stage_empty[slot]);                  // this pair's reads of the stage are done
                TRACE(2, (int)u, 6);
            }
        }
    } else {
        // ====================================================== softmax + epilogue groups
        const int g = warp >> 2;                       // group g takes units g, g + NG, ...
        const int row = tid & 127;                     // TMEM lane = 64 * (head parity) + query token
        const int hp = row >> 6, ti = row & 63, ty = ti >> 3, tx = ti & 7;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tm + TM_S + g * 128 + hp * 64 + lane_addr;
        const uint32_t tP = tm + TM_P + g * 32 + lane_addr;
        const uint32_t tO = tm + TM_O + g * 64 + hp * HD + lane_addr;
        const float c = scale * LOG2E, mv2 = mask_value * LOG2E;
        const uint64_t c2 = pack2(c, c);
        // bias row of key row yj = 0 for this thread: copy (7 - tx) % 2 at entry (7 - tx) - copy (even), table row dy = ty + 7
        const int r0 = WS - 1 - tx, cp = r0 & 1;
        const float* tab_row = tab + cp * tab_copy_stride(heads) + (ty + WS - 1) * TAB_ROW + (r0 - cp);
        const int s_ = geo.shift;
        const uint64_t yhi = s_ > 0 ? (~0ull << (8 * (WS - s_))) : 0ull;                       // keys with ty >= ws - shift
        const uint64_t xhi = s_ > 0 ? 0x0101010101010101ull * (uint64_t)((0xFFu << (WS - s_)) & 0xFFu) : 0ull;  // tx >= ws - shift
        const uint32_t srow = (uint32_t)ti * 128, sw = (uint32_t)(ti & 7);
        long long cur_win_it = -1;
        uint64_t mbits = 0;
        WinBox wb = {};                                // geometry of the window of the current unit
        float prev_inv = 0.f;
        int prev_pr = 0;
        long long prev_stage = 0;
        [[maybe_unused]] int tr_u = 0;

        // Epilogue of the group's previous unit: O / rowsum -> this head's columns of the stage's staging tile; the store warp
        // sends the tile off when every head pair of the stage has delivered.
        auto epilogue = [&]() {
            const int slot = (int)(prev_stage & (OT_RING - 1));
            const uint32_t tile_s = ot_base + (uint32_t)slot * OT_BYTES;
            uint32_t o[HD];
            if constexpr (HD == 16) tmem_ld16(tO, o); else tmem_ld32(tO, o);
            tmem_wait_ld();
            if (row == 0) TRACE(g, tr_u, 7);
            if (prev_stage >= OT_RING) mbar_wait(&ot_free[slot], (uint32_t)(((prev_stage / OT_RING) - 1) & 1));   // the slot's previous tile has left
            if (row == 0) TRACE(g, tr_u, 8);
            const float inv = prev_inv;
#pragma unroll
            for (int j = 0; j < HD; j += 8) {
                const uint32_t chunk = (uint32_t)(((2 * prev_pr + hp) * HD + j) >> 3) ^ sw;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_s + srow + (chunk << 4)),
                             "r"(pack_bf16(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv)),
                             "r"(pack_bf16(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv)),
                             "r"(pack_bf16(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv)),
                             "r"(pack_bf16(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv)) : "memory");
            }
            fence_proxy_async();                                          // staging writes -> visible to the TMA store
            mbar_arrive(&ot_full[slot]);
            if (row == 0) TRACE(g, tr_u, 9);
        };

        if (g == 1 && n_units > 0) asm volatile("bar.arrive %0, 256;" ::"r"(EXP_BAR0) : "memory");   // group 0 takes the first turn
        long long k = 0, stg = 0, win_it = 0;          // this CTA's stage / window counters of the current unit
        int pr = 0, gi = 0;
        auto advance = [&]() { if (++pr == UPS) { pr = 0; ++stg; if (++gi == groups) { gi = 0; ++win_it; } } };
        for (int i = 0; i < g; ++i) advance();
        for (long long u = g; u < n_units; u += NG, ++k) {
            if (win_it != cur_win_it) {                // new window: geometry, shifted-window mask bits
                cur_win_it = win_it;
                wb = win_box(geo, (long long)blockIdx.x + win_it * gridDim.x);
                mbits = 0;
                if (s_ > 0) {
                    if (wb.last_row) mbits |= (ty >= WS - s_) ? ~yhi : yhi;
                    if (wb.last_col) mbits |= (tx >= WS - s_) ? ~xhi : xhi;
                }
            }
            const bool any_mask = s_ > 0 && (wb.last_row || wb.last_col);
            tr_u = (int)u;
            if (row == 0) TRACE(g, tr_u, 0);
            mbar_wait(&s_full[g], (uint32_t)(k & 1));
            if (row == 0) TRACE(g, tr_u, 1);
            fence_after_sync();
            const int h = gi * G + 2 * pr + hp;
            uint64_t t[NTOK / 2];
            {
                uint32_t ra[32], rb[32];
                tmem_ld32(tS, ra);
                tmem_ld32(tS + 32, rb);
                tmem_wait_ld();
                fence_before_sync();
                mbar_arrive(&s_free[g]);                // the unit's scores are in registers
                if (row == 0) TRACE(g, tr_u, 2);
                const float* tb = tab_row + h * TAB_HEAD;
#pragma unroll
                for (int yj = 0; yj < WS; ++yj) {       // t = s * (scale log2 e) + bias, two scores per FFMA2
                    const uint32_t* r = yj < 4 ? ra + yj * 8 : rb + (yj - 4) * 8;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 b = *reinterpret_cast<const float2*>(tb - yj * TAB_ROW + 2 * q);
                        t[yj * 4 + q] = ffma2(pack2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1])), c2, pack2(b.x, b.y));
                    }
                }
            }
            if (any_mask) {                            // group-uniform: only windows of the last window row / column
#pragma unroll
                for (int j = 0; j < NTOK / 2; ++j) {
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
            for (int j = 4; j < NTOK / 2; ++j) {
                float lo, hi;
                unpack2(t[j], lo, hi);
                m4[j & 3] = fmax3(m4[j & 3], lo, hi);
            }
            const float mx = fmaxf(fmax3(m4[0], m4[1], m4[2]), m4[3]);
            const uint64_t nmx2 = pack2(-mx, -mx);
            // The exponentials of the two groups take turns on the MUFU (16 results per clock and SM: two groups in lock step
            // both crawl at half speed and nothing else of theirs overlaps); while one group is here, the other does its
            // tensor-memory loads, bias FMAs, row maxima and epilogue.
            asm volatile("bar.sync %0, 256;" ::"r"(EXP_BAR0 + g) : "memory");
            if (row == 0) TRACE(g, tr_u, 10);
            uint64_t sum2[2] = {0ull, 0ull};
            uint32_t pk[32];
#pragma unroll
            for (int j = 0; j < NTOK / 2; ++j) {
                float lo, hi;
                unpack2(fadd2(t[j], nmx2), lo, hi);
                const float p0 = fast_exp2(lo), p1 = fast_exp2(hi);
                sum2[j & 1] = fadd2(sum2[j & 1], pack2(p0, p1));
                pk[j] = pack_bf16(p0, p1);
            }
            asm volatile("bar.arrive %0, 256;" ::"r"(EXP_BAR0 + (g ^ 1)) : "memory");       // the other group's turn
            float inv_cur;
            {
                float a, b;
                unpack2(fadd2(sum2[0], sum2[1]), a, b);
                inv_cur = 1.f / (a + b);
            }
            if (row == 0) TRACE(g, tr_u, 3);
            if (k > 0) {                               // previous unit of this group: its P / O columns are free again
                mbar_wait(&pv_done[g], (uint32_t)((k - 1) & 1));
                if (row == 0) TRACE(g, tr_u, 4);
                fence_after_sync();
                epilogue();
                if (row == 0) TRACE(g, tr_u, 5);
            }
            tmem_st32(tP, pk);
            tmem_wait_st();
            fence_before_sync();
            mbar_arrive(&p_full[g]);
            if (row == 0) TRACE(g, tr_u, 6);
            prev_inv = inv_cur; prev_pr = pr; prev_stage = stg;
#pragma unroll
            for (int i = 0; i < NG; ++i) advance();
        }
        if (k > 0) {
            mbar_wait(&pv_done[g], (uint32_t)((k - 1) & 1));
            fence_after_sync();
            epilogue();
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP0) tmem_dealloc(tmem_slot, 512);
}

size_t win8_table_bytes(int heads) { return (size_t)TAB_COPIES * tab_copy_stride(heads) * sizeof(float); }

size_t win8_smem_bytes(int heads) { return (size_t)STAGES * STAGE_BYTES + OT_RING * OT_BYTES + win8_table_bytes(heads) + 1024; }

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
    const int grid = (int)(geo.total_windows < num_sms ? geo.total_windows : num_sms);
    cudaError_t e;
    if (hd == 16) {
        auto kern = window_attn_win8_kernel<16>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
        kern<<<grid, NTHREADS, smem, stream>>>(in_maps, out_maps, static_cast<const float*>(workspace), geo, C, heads, scale, mask_value);
    } else {
        auto kern = window_attn_win8_kernel<32>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
        kern<<<grid, NTHREADS, smem, stream>>>(in_maps, out_maps, static_cast<const float*>(workspace), geo, C, heads, scale, mask_value);
    }
    return check_launch();
}

}  // namespace sodt
