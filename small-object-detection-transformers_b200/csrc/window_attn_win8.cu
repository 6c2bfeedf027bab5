// 8x8-window attention (N = 64 tokens, head_dim 16 or 32) on tcgen05 tensor cores: the kernel of the
// backbone's stage 1 and stage 2 (basics/models/backbone_vit.py:114-145; shift 0 or any 0 < shift < 8).
//
// A tile is a PAIR of windows (128 query rows = 128 TMEM lanes).  Per head
//   S[128x64]                      rows 0-63 = Q_w0 K_w0^T, rows 64-127 = Q_w1 K_w1^T: two tcgen05.mma (M=128, N=64) whose
//                                  "disable output lane" masks let each write only its window's 64 lanes, so the score
//                                  buffer holds no wasted off-diagonal block (64 TMEM columns per head in flight)
//   softmax                        one thread per query row: tcgen05.ld of its 64 scores, relative position bias
//                                  (closed-form index, bank-conflict-free padded table in shared memory), shifted-window
//                                  mask from two 64-bit region masks (border windows only), exp2, row sum; P written
//                                  back to TMEM as packed bf16 (32 columns)
//   O[128xhd]                      = P V, again two lane-masked MMAs (TS: A = P from TMEM, B = V MN-major), K = 64 keys
//   epilogue                       O / rowsum -> bf16 -> straight to the un-rolled, un-partitioned output image
// Roll, partition, reverse partition and reverse roll are address arithmetic in the producer / epilogue.
//
// Persistent CTAs (one per SM), 12 warps: NG = 2 softmax groups of 4 warps take head PAIRS round-robin, each with its
// own S / P / O columns in TMEM (2 x (64 + 32 + 32) columns per group), so one group's TMEM / shared-memory / MUFU
// latencies are covered by the other.  Hand-offs between the MMA thread and a group happen once per two heads: the
// mbarrier round trips (not the math) bound the per-head version of this kernel (measured: more groups did not help); three producer warps (q, k, v) stream 64 channels
// (4 or 2 heads) per stage through a 4-stage cp.async ring; one warp issues the MMAs (a tcgen05.mma of these shapes
// occupies the tensor pipe ~64 cycles whatever N is: ten MMAs per head make the issue stream a first-order cost).
#include "common.cuh"
#include "tc05.cuh"

namespace sodt {
namespace {

using namespace tc;

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
constexpr int PLANE = ROWS * 16 + 16;
constexpr int STAGE_BYTES = 3 * 8 * PLANE;    // q,k,v x 8 planes (= G heads x hd/8 chunks = 64 channels)
constexpr int CHUNK_STRIDE = PLANE;           // bytes between 8-element chunks of a canonical tile
constexpr int OT_LD = 128 + 16;               // row pitch of the output staging tile (64 channels + pad)
constexpr int OT_BYTES = ROWS * OT_LD;
constexpr int TAB_LD = 40;                    // padded row stride of the bias table (bank-conflict free)
constexpr int TAB_ENTRIES = (2 * WS - 1) * TAB_LD;   // 600 floats per head
constexpr float LOG2E = 1.4426950408889634f;
// HPB = heads per barrier hand-off (a "unit"); 2 for head_dim 16 (4 heads per stage), 1 for head_dim 32 (2 per stage)
// TMEM column bases: S[g][hh] = g*128 + hh*64, P[g][hh] = 256 + g*64 + hh*32, O[g][hh] = 384 + g*64 + hh*32
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
    constexpr int HPB = HD == 16 ? 2 : 1;     // heads per hand-off
    constexpr int CPH = HD / 8;               // 16-byte chunks per head row
    constexpr int TILE_BYTES = CPH * PLANE;   // one (q|k|v, head) tile
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
    const long long n_total = my_tiles * (heads / HPB);   // units (head pairs) this CTA processes: u = tile_iter*(heads/2) + h/2

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
                if (round > 0) mbar_wait(&stage_empty[stage], (uint32_t)((round - 1) & 1));
                unsigned char* dst = smem + stage * STAGE_BYTES + (which * G + pg) * TILE_BYTES + pc * PLANE + r4 * 16;
                const __nv_bfloat16* src = qkv + gi * 64 + j * 8;
#pragma unroll 1
                for (int b0 = 0; b0 < ROWS / 4; b0 += 16) {       // 16 x 512 B in flight per warp
                    uint4 v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __ldg(reinterpret_cast<const uint4*>(src + my_off[(b0 + i) * 4 + r4]));
#pragma unroll
                    for (int i = 0; i < 16; ++i) *reinterpret_cast<uint4*>(dst + (b0 + i) * 64) = v[i];
                }
                fence_proxy_async();
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
            constexpr uint32_t idesc_o = idesc_bf16(ROWS, HD, false, true);
            constexpr uint32_t ALL = 0xFFFFFFFFu;
            constexpr int UPS = G / HPB;                                       // units per stage
            const uint64_t kdesc0 = smem_desc(sbase, CHUNK_STRIDE, 128);      // K-major tiles (Q, K)
            const uint64_t vdesc0 = smem_desc(sbase, 128, CHUNK_STRIDE);      // MN-major tile (V)
            const int nt = (int)n_total;
            // ---- cursor of the next unit whose scores are to be issued
            int qn = 0, q_us = 0, q_stage = 0, q_stage_par = 0, q_g = 0, q_k = 0;
            auto issue_qk = [&]() {
                if (q_us == 0) mbar_wait(&stage_full[q_stage], (uint32_t)q_stage_par);
                if (q_k > 0) mbar_wait(&s_free[q_g], (uint32_t)((q_k - 1) & 1));
                fence_after_sync();
#pragma unroll
                for (int hh = 0; hh < HPB; ++hh) {
                    const uint32_t toff = (uint32_t)(q_stage * STAGE_BYTES + (q_us * HPB + hh) * TILE_BYTES) >> 4;
                    const uint64_t qd = kdesc0 + toff, kd = kdesc0 + toff + ((G * TILE_BYTES) >> 4);
                    const uint32_t d = tm + TM_S + q_g * (HPB * 64) + hh * 64;
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks)
                        mma_ss_masked(d, qd + ((ks * 2 * CHUNK_STRIDE) >> 4), kd + ((ks * 2 * CHUNK_STRIDE) >> 4), idesc_s, ks > 0, 0u, 0u, ALL, ALL);
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks)
                        mma_ss_masked(d, qd + ((ks * 2 * CHUNK_STRIDE) >> 4), kd + ((NTOK * 16 + ks * 2 * CHUNK_STRIDE) >> 4), idesc_s, ks > 0,
                                      ALL, ALL, 0u, 0u);
                }
                mma_commit(&s_full[q_g]);
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
                mbar_wait(&p_full[g], (uint32_t)(k & 1));
                fence_after_sync();
#pragma unroll
                for (int hh = 0; hh < HPB; ++hh) {
                    const uint64_t vd = vdesc0 + ((uint32_t)(stage * STAGE_BYTES + (2 * G + us * HPB + hh) * TILE_BYTES) >> 4);
                    const uint32_t d = tm + TM_O + g * (HPB * 32) + hh * 32, a = tm + TM_P + g * (HPB * 32) + hh * 32;
#pragma unroll
                    for (int ks = 0; ks < NTOK / 16; ++ks)
                        mma_ts_masked(d, a + ks * 8, vd + ((ks * 256) >> 4), idesc_o, ks > 0, 0u, 0u, ALL, ALL);
#pragma unroll
                    for (int ks = 0; ks < NTOK / 16; ++ks)
                        mma_ts_masked(d, a + ks * 8, vd + ((NTOK * 16 + ks * 256) >> 4), idesc_o, ks > 0, ALL, ALL, 0u, 0u);
                }
                mma_commit(&pv_done[g]);
                if (++us == UPS) { us = 0; mma_commit(&stage_empty[stage]); if (++stage == STAGES) stage = 0; }
                if (++g == NG) { g = 0; ++k; }
            }
        }
    } else {
        // ====================================================== softmax + epilogue groups
        const int g = warp >> 2;                       // group g takes units n = g, g + NG, ...
        const int row = tid & 127;                     // TMEM lane == query row of the pair
        const int wh = row >> 6, ti = row & 63, ty = ti >> 3, tx = ti & 7;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tm + TM_S + g * (HPB * 64) + lane_addr;
        const uint32_t tP = tm + TM_P + g * (HPB * 32) + lane_addr;
        const uint32_t tO = tm + TM_O + g * (HPB * 32) + lane_addr;
        const float c = scale * LOG2E, mv2 = mask_value * LOG2E;
        const float* tab_row = tab + (ty + WS - 1) * TAB_LD + (tx + WS - 1);
        const int s_ = geo.shift;
        const uint64_t yhi = s_ > 0 ? (~0ull << (8 * (WS - s_))) : 0ull;                       // keys with ty >= ws - shift
        const uint64_t xhi = s_ > 0 ? 0x0101010101010101ull * (uint64_t)((0xFFu << (WS - s_)) & 0xFFu) : 0ull;  // tx >= ws - shift
        const int nwh = geo.nW / geo.nww;
        const int units_per_tile = heads / HPB;
        long long k = 0, cur_iter = -1;
        __nv_bfloat16* out_tok = nullptr;
        uint64_t mbits = 0;
        bool any_mask = false;
        // Epilogue: O / rowsum of the previous unit goes to the stage's staging tile; when both groups have delivered
        // their unit of that stage, all 256 softmax threads write the 128 rows x 128 B of the tile with full-line stores.
        bool have_prev = false;
        int prev_hcol = 0;                              // byte column of the previous unit's first head inside the tile
        long long prev_stage = 0;                       // global stage index of the previous unit (tile_iter*groups + gi)
        float prev_inv[HPB] = {};
        const int st_tid = tid;                         // 0..255 among the softmax threads
        auto epilogue = [&]() {
            unsigned char* tile = ot + (prev_stage & 1) * OT_BYTES;
#pragma unroll
            for (int hh = 0; hh < HPB; ++hh) {
                uint32_t o[HD];
                if constexpr (HD == 16) tmem_ld16(tO + hh * 32, o); else tmem_ld32(tO + hh * 32, o);
                tmem_wait_ld();
                const float inv = prev_inv[hh];
#pragma unroll
                for (int j = 0; j < HD; j += 8) {
                    uint4 v;
                    v.x = pack_bf16(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv);
                    v.y = pack_bf16(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
                    v.z = pack_bf16(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv);
                    v.w = pack_bf16(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv);
                    *reinterpret_cast<uint4*>(tile + row * OT_LD + prev_hcol + hh * HD * 2 + j * 2) = v;
                }
            }
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
                const long long wdx = 2 * ((long long)blockIdx.x + it * gridDim.x) + wh;
                out_tok = nullptr;
                mbits = 0;
                if (g == 0) out_off[(it & 1) * ROWS + row] = wdx < geo.total_windows ? geo.token(wdx, ty, tx) * (long long)C : -1;
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
            mbar_wait(&s_full[g], (uint32_t)(k & 1));
            fence_after_sync();
            float inv_cur[HPB];
#pragma unroll
            for (int hh = 0; hh < HPB; ++hh) {
                const int h = ut * HPB + hh;
                float s2[NTOK];
                {
                    uint32_t r0[32], r1[32];
                    tmem_ld32(tS + hh * 64, r0);
                    tmem_ld32(tS + hh * 64 + 32, r1);
                    tmem_wait_ld();
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
                if (hh == 0 && k > 0) {                // previous unit of this group: its P / O columns are free again
                    mbar_wait(&pv_done[g], (uint32_t)((k - 1) & 1));
                    fence_after_sync();
                    epilogue();
                }
                tmem_st32(tP + hh * 32, pk);
            }
            tmem_wait_st();
            fence_before_sync();
            mbar_arrive(&p_full[g]);
            have_prev = true;
            prev_stage = it * groups + (ut * HPB) / G;
            prev_hcol = ((ut * HPB) % G) * HD * 2;
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
