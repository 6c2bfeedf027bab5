// 8x8-window attention (N = 64 tokens, head_dim 16 or 32) on tcgen05 tensor cores: the kernel of the
// backbone's stage 1 and stage 2 (basics/models/backbone_vit.py:114-145; shift 0 or any 0 < shift < 8).
//
// The 128 TMEM lanes of an MMA hold TWO HEADS of one window: lane = 64 * (head parity) + query token.  Per head pair
//   S[128x64]                      lanes 0-63 = Q_h0 K_h0^T, lanes 64-127 = Q_h1 K_h1^T: two tcgen05.mma (M=128, N=64) whose
//                                  "disable output lane" masks let each write only its head's 64 lanes (64 TMEM columns)
//   softmax                        one thread per (head, query row): tcgen05.ld of its 64 scores, relative position bias
//                                  (closed-form index, bank-conflict-free padded table in shared memory), shifted-window
//                                  mask from two 64-bit region masks (border windows only), exp2, row sum; P written
//                                  back to TMEM as packed bf16 (32 columns)
//   O[128 x 2hd]                   = P [V_h0 | V_h1]: ONE unmasked MMA chain (TS: A = P from TMEM, B = the two heads'
//                                  adjacent V channels, MN-major, N = 2 hd); lane half x reads its head's hd columns.
//                                  (Pairing two WINDOWS per MMA instead needed two lane-masked chains per head: a masked
//                                  MMA of these shapes holds the tensor pipe ~100 cycles whatever N is, measured.)
//   epilogue                       O / rowsum -> bf16 -> staging tile -> full 128-byte lines of the un-rolled output image
// Roll, partition, reverse partition and reverse roll are address arithmetic in the producer / epilogue.
//
// Persistent CTAs (one per SM) walk PAIRS of windows (128 token rows per shared-memory stage), 12 warps: softmax group g
// (4 warps) owns window g of every pair and takes its head pairs in order, with its own S / P / O columns in TMEM, so
// one group's TMEM / shared-memory / MUFU latencies are covered by the other.  Hand-offs between the MMA thread and a
// group happen once per unit (head_dim 16: two head pairs = the 64 channels of a stage; head_dim 32: one pair): the
// mbarrier round trips, not the math, bound a per-head hand-off.  Three producer warps (q, k, v) stream 64 channels per
// stage with coalesced 128-byte-per-token loads; one warp issues the MMAs.
#include "common.cuh"
#include "tc05.cuh"

namespace sodt {
namespace {

using namespace tc;

// Timeline instrumentation for tests/probes/win8_trace.cu (compiled only there): clock64 at the hand-offs of CTA 0.
#ifdef SODT_WIN8_TRACE
__device__ long long g_trace[4][96][12];
#define TRACE(role, unit, ev) do { if (blockIdx.x == 0 && (unit) < 96) g_trace[role][unit][ev] = clock64(); } while (0)
#else
#define TRACE(role, unit, ev) do { } while (0)
#endif

constexpr int WS = 8;
constexpr int NTOK = 64;                 // tokens per window
constexpr int ROWS = 128;                // rows per tile = 2 windows
constexpr int NG = 2;                    // softmax groups (heads in flight)
constexpr int NPROD = 3;                 // producer warps: one each for q, k, v
constexpr int NTHREADS = (NG * 4 + 1 + NPROD) * 32;
constexpr int MMA_WARP = NG * 4, PRODUCER_WARP0 = NG * 4 + 1;
constexpr int STAGES = 3;
// A canonical (SWIZZLE_NONE) operand tile is a set of "planes": plane c holds the 16-byte chunk c (8 channels) of all 128
// rows.  Planes are padded by 16 B so that the 8 chunks of one token row fall into 8 different bank groups when a
// producer warp writes a whole 128-byte row segment (conflict-free), which lets the producers use fully coalesced
// 128-byte-per-token global loads (cp.async fetched a 32-byte sector per 16-byte request: 2.5x L2 read traffic).
constexpr int PLANE = ROWS * 16 + 16;         // V: plane c = channels 8c..8c+7 of the stage, rows = (window, token)
// Q and K: plane (head pair p, chunk pc) holds chunk pc of BOTH heads of the pair, rows ordered (window, head parity, token),
// so the 128 rows of one window's head pair are contiguous: the K-major A / B operand of that pair's score MMAs.
constexpr int PLANE2 = 2 * ROWS * 16 + 32;
constexpr int OPERAND_BYTES = 8 * PLANE;      // q, k or v part of a stage (8 planes x 128 rows = 4 planes2 x 256 rows)
constexpr int STAGE_BYTES = 3 * OPERAND_BYTES;
constexpr int OT_LD = 128 + 16;               // row pitch of the output staging tile (64 channels + pad)
constexpr int OT_BYTES = ROWS * OT_LD;
constexpr int TAB_LD = 40;                    // padded row stride of the bias table (bank-conflict free)
constexpr int TAB_ENTRIES = (2 * WS - 1) * TAB_LD;   // 600 floats per head
constexpr float LOG2E = 1.4426950408889634f;
// HPB = head pairs per barrier hand-off (a "unit" = one window x the stage's 64 channels): 2 for head_dim 16, 1 for 32
// TMEM column bases: S[g][hh] = g*HPB*64 + hh*64, P[g][hh] = 256 + g*HPB*32 + hh*32, O[g][hh] = 384 + g*64 + hh*2*hd
constexpr uint32_t TM_S = 0, TM_P = 256, TM_O = 384;

// [heads][15][40] bias table, scaled by log2(e):  tab[h][(dy+7)*40 + (dx+7)]
__global__ void prep_table_win8_kernel(const float* __restrict__ table, float* __restrict__ out, int heads) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= heads * TAB_ENTRIES) return;
    const int h = e / TAB_ENTRIES, r = e - h * TAB_ENTRIES;
    const int dy = r / TAB_LD, dx = r - dy * TAB_LD;
    out[e] = dx < 2 * WS - 1 ? table[(long long)(dy * (2 * WS - 1) + dx) * heads + h] * LOG2E : 0.f;
}

struct Geo {
    int H, W, nww, nW, shift;
    long long total_windows;
    // token (window index wdx, row ty, col tx) -> token index in the un-rolled image
    __device__ __forceinline__ long long token(long long wdx, int ty, int tx) const {
        const int b = (int)(wdx / nW);
        const int win = (int)(wdx - (long long)b * nW);
        const int wy = win / nww, wx = win - wy * nww;
        int ys = wy * WS + ty + shift; if (ys >= H) ys -= H;
        int xs = wx * WS + tx + shift; if (xs >= W) xs -= W;
        return ((long long)b * H + ys) * W + xs;
    }
};

template <int HD>
__global__ void __launch_bounds__(NTHREADS, 1)
window_attn_win8_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ table_p,
                        __nv_bfloat16* __restrict__ out, Geo geo, int C, int heads, float scale, float mask_value,
                        long long ntiles) {
    constexpr int G = 64 / HD;                // heads per stage
    constexpr int HPB = G / 2;                // head pairs per stage = per hand-off
    constexpr int CPH = HD / 8;               // 16-byte chunks per head row
    constexpr int PAIR_BYTES = CPH * PLANE2;  // q or k of one head pair, both windows
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t stage_full[STAGES], stage_empty[STAGES], s_full[NG], s_free[NG], p_full[NG], pv_done[NG];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = smem_u32(smem);
    unsigned char* ot = smem + STAGES * STAGE_BYTES;                                   // [2][ROWS][OT_LD] output staging
    long long* out_off = reinterpret_cast<long long*>(ot + 2 * OT_BYTES);               // [2][ROWS] output element offset / -1
    long long* in_off = out_off + 2 * ROWS;                                            // [NPROD][ROWS] producers' token offsets
    float* tab = reinterpret_cast<float*>(in_off + NPROD * ROWS);
    const int C3 = 3 * C;
    const int groups = heads / G;
    long long my_tiles = 0;
    if ((long long)blockIdx.x < ntiles) my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const long long n_total = my_tiles * groups * 2;      // units this CTA processes: u = (tile_iter*groups + stage)*2 + window

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&stage_full[s], 32 * NPROD); mbar_init(&stage_empty[s], 1); }
        for (int g = 0; g < NG; ++g) { mbar_init(&s_full[g], 1); mbar_init(&s_free[g], ROWS); mbar_init(&p_full[g], ROWS); mbar_init(&pv_done[g], 1); }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    for (int e = tid; e < heads * TAB_ENTRIES; e += NTHREADS) tab[e] = table_p[e];
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_slot;
    if (warp >= PRODUCER_WARP0) {
        // ===================================== producers: warp `which` streams q (0), k (1) or v (2) of every stage.
        // A warp-level load covers 4 token rows x 128 B (lane = 8*row + chunk): full lines, no sector over-fetch.
        const int which = warp - PRODUCER_WARP0;
        const int r4 = lane >> 3, j = lane & 7;                      // row within a group of 4, 16-byte chunk of the 128-byte segment
        const int pg = j / CPH, pc = j % CPH;                        // head in stage, chunk in head
        long long* my_off = in_off + which * ROWS;
        int stage = 0, round = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = lane * 4 + q;
                long long wdx = 2 * tile + (r >> 6);
                if (wdx >= geo.total_windows) wdx = 2 * tile;        // odd tail: duplicate the first window
                my_off[r] = geo.token(wdx, (r >> 3) & 7, r & 7) * C3 + which * C + j * 0;
            }
            __syncwarp();
            for (int gi = 0; gi < groups; ++gi) {
                if (which == 0 && lane == 0) TRACE(3, (int)((tile - blockIdx.x) / gridDim.x) * groups + gi, 0);
                if (round > 0) mbar_wait(&stage_empty[stage], (uint32_t)((round - 1) & 1));
                if (which == 0 && lane == 0) TRACE(3, (int)((tile - blockIdx.x) / gridDim.x) * groups + gi, 1);
                unsigned char* dst = smem + stage * STAGE_BYTES + which * OPERAND_BYTES + r4 * 16 +
                                     (which < 2 ? (pg >> 1) * PAIR_BYTES + pc * PLANE2 + (pg & 1) * (NTOK * 16) : (pg * CPH + pc) * PLANE);
                const int wskip = which < 2 ? NTOK * 16 : 0;        // q, k: window 1 rows start after both heads of window 0
                const __nv_bfloat16* src = qkv + gi * 64 + j * 8;
#pragma unroll 1
                for (int b0 = 0; b0 < ROWS / 4; b0 += 16) {       // 16 x 512 B in flight per warp; b0 = 16 * window
                    uint4 v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __ldg(reinterpret_cast<const uint4*>(src + my_off[(b0 + i) * 4 + r4]));
#pragma unroll
                    for (int i = 0; i < 16; ++i) *reinterpret_cast<uint4*>(dst + (b0 + i) * 64 + (b0 >> 4) * wskip) = v[i];
                }
                fence_proxy_async();
                if (which == 0 && lane == 0) TRACE(3, (int)((tile - blockIdx.x) / gridDim.x) * groups + gi, 2);
                mbar_arrive(&stage_full[stage]);
                if (++stage == STAGES) { stage = 0; ++round; }
            }
        }
    } else if (warp == MMA_WARP) {
        // =============================================================== MMA issuer
        // One thread; its instruction stream is kept short: all counters are 32-bit and incremental (no divisions),
        // descriptors are a per-tile base plus compile-time constants.  Work is issued per unit = HPB consecutive heads.
        if (lane == 0) {
            constexpr uint32_t idesc_s = idesc_bf16(ROWS, NTOK, false, false);
            constexpr uint32_t idesc_o = idesc_bf16(ROWS, 2 * HD, false, true);
            constexpr uint32_t ALL = 0xFFFFFFFFu;
            constexpr int UPS = 2;                                             // units per stage = windows of the pair
            const uint64_t kdesc0 = smem_desc(sbase, PLANE2, 128);            // K-major tiles (Q, K)
            const uint64_t vdesc0 = smem_desc(sbase, 128, PLANE);             // MN-major tile (V)
            const int nt = (int)n_total;
            // ---- cursor of the next unit whose scores are to be issued
            int qn = 0, q_us = 0, q_stage = 0, q_stage_par = 0, q_g = 0, q_k = 0;
            auto issue_qk = [&]() {
                TRACE(2, qn, 0);
                if (q_us == 0) mbar_wait(&stage_full[q_stage], (uint32_t)q_stage_par);
                TRACE(2, qn, 1);
                if (q_k > 0) mbar_wait(&s_free[q_g], (uint32_t)((q_k - 1) & 1));
                TRACE(2, qn, 2);
                fence_after_sync();
#pragma unroll
                for (int hh = 0; hh < HPB; ++hh) {
                    const uint32_t toff = (uint32_t)(q_stage * STAGE_BYTES + hh * PAIR_BYTES + q_us * (2 * NTOK * 16)) >> 4;
                    const uint64_t qd = kdesc0 + toff, kd = kdesc0 + toff + (OPERAND_BYTES >> 4);
                    const uint32_t d = tm + TM_S + q_g * (HPB * 64) + hh * 64;
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks)       // lanes 0-63: even head of the pair
                        mma_ss_masked(d, qd + ((ks * 2 * PLANE2) >> 4), kd + ((ks * 2 * PLANE2) >> 4), idesc_s, ks > 0, 0u, 0u, ALL, ALL);
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks)       // lanes 64-127: odd head (its K rows follow the even head's)
                        mma_ss_masked(d, qd + ((ks * 2 * PLANE2) >> 4), kd + ((NTOK * 16 + ks * 2 * PLANE2) >> 4), idesc_s, ks > 0,
                                      ALL, ALL, 0u, 0u);
                }
                mma_commit(&s_full[q_g]);
                TRACE(2, qn, 3);
                ++qn;
                if (++q_us == UPS) { q_us = 0; if (++q_stage == STAGES) { q_stage = 0; q_stage_par ^= 1; } }
                if (++q_g == NG) { q_g = 0; ++q_k; }
            };
            for (int i = 0; i < NG && i < nt; ++i) issue_qk();
            int us = 0, stage = 0, g = 0, k = 0;
            for (int n = 0; n < nt; ++n) {
                // S[g] is free as soon as softmax(n) has copied it to registers, so the group's next scores are
                // computed while softmax(n) is still running
                if (qn < nt) issue_qk();
                TRACE(2, n, 4);
                mbar_wait(&p_full[g], (uint32_t)(k & 1));
                TRACE(2, n, 5);
                fence_after_sync();
                // O[128 x 2hd] = P [V_even | V_odd] per head pair; the pairs' chains are interleaved (independent accumulators)
#pragma unroll
                for (int ks = 0; ks < NTOK / 16; ++ks) {
#pragma unroll
                    for (int hh = 0; hh < HPB; ++hh) {
                        const uint64_t vd = vdesc0 + ((uint32_t)(stage * STAGE_BYTES + 2 * OPERAND_BYTES + hh * 2 * CPH * PLANE + us * (NTOK * 16)) >> 4);
                        const uint32_t d = tm + TM_O + g * 64 + hh * (2 * HD), a = tm + TM_P + g * (HPB * 32) + hh * 32;
                        mma_ts(d, a + ks * 8, vd + ((ks * 256) >> 4), idesc_o, ks > 0);
                    }
                }
                mma_commit(&pv_done[g]);
                TRACE(2, n, 6);
                if (++us == UPS) { us = 0; mma_commit(&stage_empty[stage]); if (++stage == STAGES) stage = 0; }
                if (++g == NG) { g = 0; ++k; }
            }
        }
    } else {
        // ====================================================== softmax + epilogue groups
        const int g = warp >> 2;                       // group g takes units n = g, g + NG, ... = window g of every pair
        const int row = tid & 127;                     // TMEM lane = 64 * (head parity) + query token
        const int hp = row >> 6, ti = row & 63, ty = ti >> 3, tx = ti & 7;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tm + TM_S + g * (HPB * 64) + lane_addr;
        const uint32_t tP = tm + TM_P + g * (HPB * 32) + lane_addr;
        const uint32_t tO = tm + TM_O + g * 64 + hp * HD + lane_addr;
        const float c = scale * LOG2E, mv2 = mask_value * LOG2E;
        const float* tab_row = tab + (ty + WS - 1) * TAB_LD + (tx + WS - 1);
        const int s_ = geo.shift;
        const uint64_t yhi = s_ > 0 ? (~0ull << (8 * (WS - s_))) : 0ull;                       // keys with ty >= ws - shift
        const uint64_t xhi = s_ > 0 ? 0x0101010101010101ull * (uint64_t)((0xFFu << (WS - s_)) & 0xFFu) : 0ull;  // tx >= ws - shift
        const int nwh = geo.nW / geo.nww;
        const int units_per_tile = groups * 2;
        long long k = 0, cur_iter = -1;
        __nv_bfloat16* out_tok = nullptr;
        uint64_t mbits = 0;
        bool any_mask = false;
        // Epilogue: O / rowsum of the previous unit goes to the stage's staging tile; when both groups have delivered
        // their unit of that stage, all 256 softmax threads write the 128 rows x 128 B of the tile with full-line stores.
        bool have_prev = false;
        long long prev_stage = 0;                       // global stage index of the previous unit (tile_iter*groups + gi)
        float prev_inv[HPB] = {};
        const int st_tid = tid;                         // 0..255 among the softmax threads
        [[maybe_unused]] int trace_n = 0;
        auto epilogue = [&]() {
            unsigned char* tile = ot + (prev_stage & 1) * OT_BYTES;
#pragma unroll
            for (int hh = 0; hh < HPB; ++hh) {
                uint32_t o[HD];
                if constexpr (HD == 16) tmem_ld16(tO + hh * (2 * HD), o); else tmem_ld32(tO + hh * (2 * HD), o);
                tmem_wait_ld();
                if (row == 0 && hh == 0) TRACE(g, trace_n, 9);
                const float inv = prev_inv[hh];
#pragma unroll
                for (int j = 0; j < HD; j += 8) {
                    uint4 v;
                    v.x = pack_bf16(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv);
                    v.y = pack_bf16(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
                    v.z = pack_bf16(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv);
                    v.w = pack_bf16(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv);
                    *reinterpret_cast<uint4*>(tile + (g * NTOK + ti) * OT_LD + (2 * hh + hp) * HD * 2 + j * 2) = v;
                }
            }
            if (row == 0) TRACE(g, trace_n, 10);
            asm volatile("bar.sync 2, 256;" ::: "memory");           // both groups' halves of the stage tile are in smem
            const long long tile_it = prev_stage / groups;
            const int gcol = (int)(prev_stage - tile_it * groups) * 64;
            const long long* offs = out_off + (tile_it & 1) * ROWS;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int e = st_tid + q * 256;
                const int r = e >> 3, ch = e & 7;
                const long long o = offs[r];
                if (o >= 0) *reinterpret_cast<uint4*>(out + o + gcol + ch * 8) = *reinterpret_cast<const uint4*>(tile + r * OT_LD + ch * 16);
            }
        };
        int ut = g;                                     // unit index inside the current pair; (it, ut) advance without divisions
        long long it = 0;
        while (ut >= units_per_tile) { ut -= units_per_tile; ++it; }
        for (long long n = g; n < n_total; n += NG, ++k) {
            if (it != cur_iter) {                      // new window pair: output row pointer and shifted-window mask bits
                cur_iter = it;
                const long long wdx = 2 * ((long long)blockIdx.x + it * gridDim.x) + g;
                out_tok = nullptr;
                mbits = 0;
                if (g == 0) {                          // output offsets of the pair's 128 token rows (row = 64 * window + token)
                    const long long wr = wdx + hp;
                    out_off[(it & 1) * ROWS + row] = wr < geo.total_windows ? geo.token(wr, ty, tx) * (long long)C : -1;
                }
                if (wdx < geo.total_windows) {
                    out_tok = out;
                    if (s_ > 0) {
                        const int win = (int)(wdx % geo.nW);
                        const int wy = win / geo.nww, wx = win - wy * geo.nww;
                        if (wy == nwh - 1) mbits |= (ty >= WS - s_) ? ~yhi : yhi;
                        if (wx == geo.nww - 1) mbits |= (tx >= WS - s_) ? ~xhi : xhi;
                    }
                }
                any_mask = __any_sync(0xffffffffu, mbits != 0);
            }
            if (row == 0) TRACE(g, (int)n, 0);
            mbar_wait(&s_full[g], (uint32_t)(k & 1));
            if (row == 0) TRACE(g, (int)n, 1);
            fence_after_sync();
            float inv_cur[HPB];
#pragma unroll
            for (int hh = 0; hh < HPB; ++hh) {
                const int h = (ut >> 1) * G + 2 * hh + hp;
                float s2[NTOK];
                {
                    uint32_t r0[32], r1[32];
                    tmem_ld32(tS + hh * 64, r0);
                    tmem_ld32(tS + hh * 64 + 32, r1);
                    tmem_wait_ld();
                    if (row == 0) TRACE(g, (int)n, 2 + hh * 4);
#pragma unroll
                    for (int j = 0; j < 32; ++j) { s2[j] = __uint_as_float(r0[j]); s2[32 + j] = __uint_as_float(r1[j]); }
                }
                if (hh == HPB - 1) { fence_before_sync(); mbar_arrive(&s_free[g]); }    // both heads' scores are in registers
                const float* tb = tab_row + h * TAB_ENTRIES;
#pragma unroll
                for (int j = 0; j < NTOK; ++j) s2[j] = fmaf(s2[j], c, tb[-((j >> 3) * TAB_LD + (j & 7))]);
                if (any_mask) {                        // warp-uniform: only windows of the last window row / column
#pragma unroll
                    for (int j = 0; j < NTOK; ++j)
                        if ((mbits >> j) & 1ull) s2[j] += mv2;
                }
                float mx = s2[0];
#pragma unroll
                for (int j = 1; j < NTOK; ++j) mx = fmaxf(mx, s2[j]);
                float sum = 0.f;
                uint32_t pk[32];
#pragma unroll
                for (int j = 0; j < NTOK; j += 2) {
                    const float p0 = fast_exp2(s2[j] - mx), p1 = fast_exp2(s2[j + 1] - mx);
                    sum += p0 + p1;
                    pk[j >> 1] = pack_bf16(p0, p1);
                }
                inv_cur[hh] = 1.f / sum;
                if (row == 0) TRACE(g, (int)n, 3 + hh * 4);
                if (hh == 0 && k > 0) {                // previous unit of this group: its P / O columns are free again
                    mbar_wait(&pv_done[g], (uint32_t)((k - 1) & 1));
                    if (row == 0) TRACE(g, (int)n, 4);
                    fence_after_sync();
                    trace_n = (int)n;
                    epilogue();
                    if (row == 0) TRACE(g, (int)n, 5);
                }
                tmem_st32(tP + hh * 32, pk);
            }
            tmem_wait_st();
            fence_before_sync();
            mbar_arrive(&p_full[g]);
            if (row == 0) TRACE(g, (int)n, 8);
            have_prev = true;
            prev_stage = it * groups + (ut >> 1);
#pragma unroll
            for (int hh = 0; hh < HPB; ++hh) prev_inv[hh] = inv_cur[hh];
            ut += NG;
            while (ut >= units_per_tile) { ut -= units_per_tile; ++it; }
        }
        if (have_prev) {
            mbar_wait(&pv_done[g], (uint32_t)((k - 1) & 1));
            fence_after_sync();
            epilogue();
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_slot, 512);
}


}  // namespace

static size_t win8_smem_bytes(int heads) {
    return (size_t)STAGES * STAGE_BYTES + 2 * OT_BYTES + (2 + NPROD) * ROWS * sizeof(long long) + (size_t)heads * TAB_ENTRIES * sizeof(float);
}

size_t window_attn_win8_workspace(int heads) { return (size_t)heads * TAB_ENTRIES * sizeof(float); }

bool window_attn_win8_supported(int H, int W, int C, int heads, int ws, int shift, int dtype) {
    if (dtype != SODT_BF16 || ws != WS || H % WS || W % WS || heads <= 0 || C % heads) return false;
    const int hd = C / heads;
    if (hd != 16 && hd != 32) return false;
    const int G = 64 / hd;
    if (heads % G || heads % 2) return false;
    return win8_smem_bytes(heads) <= 227 * 1024 - 2048;
}

int window_attn_win8(const void* qkv, const float* table, void* out, void* workspace, int B, int H, int W, int C,
                     int heads, int shift, float scale, float mask_value, int num_sms, cudaStream_t stream) {
    float* table_p = static_cast<float*>(workspace);
    prep_table_win8_kernel<<<(heads * TAB_ENTRIES + 255) / 256, 256, 0, stream>>>(table, table_p, heads);
    int st = check_launch();
    if (st != SODT_OK) return st;
    Geo geo;
    geo.H = H; geo.W = W; geo.nww = W / WS; geo.nW = (H / WS) * (W / WS); geo.shift = shift;
    geo.total_windows = (long long)B * geo.nW;
    const long long ntiles = (geo.total_windows + 1) / 2;
    const size_t smem = win8_smem_bytes(heads);
    const int hd = C / heads;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    cudaError_t e;
    if (hd == 16) {
        auto kern = window_attn_win8_kernel<16>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
        kern<<<grid, NTHREADS, smem, stream>>>(static_cast<const __nv_bfloat16*>(qkv), table_p, static_cast<__nv_bfloat16*>(out),
                                               geo, C, heads, scale, mask_value, ntiles);
    } else {
        auto kern = window_attn_win8_kernel<32>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
        kern<<<grid, NTHREADS, smem, stream>>>(static_cast<const __nv_bfloat16*>(qkv), table_p, static_cast<__nv_bfloat16*>(out),
                                               geo, C, heads, scale, mask_value, ntiles);
    }
    return check_launch();
}

}  // namespace sodt
