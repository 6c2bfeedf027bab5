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
// Work unit: a PAIR of windows (128 token rows) x 64 channels (4 heads of 16 or 2 heads of 32 channels) = one "stage" of
// 48 KB (q, k, v tiles).  The 128 TMEM lanes of an MMA are (window of the pair, query token).  Per head
//   S[128x64]   two lane-masked tcgen05.mma (SS, M=128, N=64, K=head_dim): lanes 0-63 = Q_w0 K_w0^T, lanes 64-127 = Q_w1 K_w1^T;
//               A / B are K-major SWIZZLE_128B descriptors advanced to the head's 32 / 64 bytes inside the 128-byte rows
//   softmax     one thread per (window, query row): tcgen05.ld of its 64 scores, relative position bias (closed-form index
//               into a shared-memory table laid out for 16-byte loads), shifted-window mask from two 64-bit region masks
//               (border windows only), packed fp32x2 arithmetic, exp2 on the MUFU; P goes back to TMEM as packed bf16
//   O[128xhd]   two lane-masked TS chains (A = P from TMEM, B = V_w MN-major SWIZZLE_128B at the head's column offset, K = 64 keys)
//   epilogue    O / rowsum -> bf16 -> SWIZZLE_128B staging tile -> TMA store of the un-rolled image tile
// Persistent CTAs (one per SM), 10 warps: two softmax groups of 4 warps (group g takes heads [g*HPB, (g+1)*HPB) of every stage,
// with private S / P / O columns in TMEM), one MMA-issuing warp, one TMA-producer warp.
#include "common.cuh"
#include "tma.cuh"

namespace sodt {
namespace {

using namespace tc;

constexpr int WS = 8;
constexpr int NTOK = 64;                 // tokens per window
constexpr int ROWS = 128;                // rows per tile = 2 windows
constexpr int NG = 2;                    // softmax groups
constexpr int NTHREADS = (NG * 4 + 2) * 32;
constexpr int MMA_WARP = NG * 4, TMA_WARP = NG * 4 + 1;
constexpr int STAGES = 3;
constexpr int WIN_BYTES = NTOK * 128;         // one window of one operand: 64 rows x 128 B
constexpr int OPERAND_BYTES = ROWS * 128;     // q, k or v part of a stage
constexpr int STAGE_BYTES = 3 * OPERAND_BYTES;
constexpr int OT_BYTES = ROWS * 128;          // output staging tile of a stage
constexpr float LOG2E = 1.4426950408889634f;
// Bias table in shared memory: 4 copies (copy c is shifted left by c entries) of [head][15 rows dy][12] floats holding the
// x-REVERSED table row, R[dy][r] = table[dy][14 - r], so that the 8 biases of one key row are the ascending entries
// r = 7 - tx + xj and start 16-byte aligned in copy (7 - tx) % 4: two LDS.128 per key row instead of eight LDS.32.
constexpr int TAB_ROW = 12, TAB_HEAD = (2 * WS - 1) * TAB_ROW;          // 180 floats per head and copy
__host__ __device__ constexpr int tab_copy_stride(int heads) {         // floats; bytes = 32 (mod 128): the 8 lanes of a quarter
    return ((heads * TAB_HEAD * 4 + 95) / 128 * 128 + 32) / 4;          // warp (4 copies x 2 offsets) hit 8 distinct bank groups
}
// TMEM columns: S[g][hh] = g*HPB*64 + hh*64, P[g][hh] = 256 + g*HPB*32 + hh*32, O[g][hh] = 384 + g*32 + hh*hd   (HPB*hd = 32)
constexpr uint32_t TM_S = 0, TM_P = 256, TM_O = 384;
constexpr uint32_t ALL = 0xFFFFFFFFu;

__global__ void prep_table_win8_kernel(const float* __restrict__ table, float* __restrict__ out, int heads) {
    const int cs = tab_copy_stride(heads);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 4 * cs) return;
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
    bool wrap_x, wrap_y, valid;
};
__device__ __forceinline__ WinBox win_box(const Geo& g, long long wdx) {
    WinBox r;
    r.valid = wdx < g.total_windows;
    if (!r.valid) wdx = g.total_windows - 1;
    const int b = (int)(wdx / g.nW);
    const int win = (int)(wdx - (long long)b * g.nW);
    const int wy = win / g.nww, wx = win - wy * g.nww;
    r.x0 = wx * WS + g.shift;
    r.y0 = wy * WS + g.shift;
    r.wrap_x = r.x0 + WS > g.W;
    r.wrap_y = r.y0 + WS > g.H;
    r.yg_base = b * g.H;
    return r;
}

struct Maps {
    CUtensorMap full, row8, row_a, row_b;     // boxes (64, 8, 8), (64, 8, 1), (64, 8 - shift, 1), (64, shift, 1)
};

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
                        const float* __restrict__ table_p, Geo geo, int C, int heads, float scale, float mask_value,
                        long long ntiles) {
    constexpr int G = 64 / HD;                // heads per stage
    constexpr int HPB = G / NG;               // heads per softmax group and stage
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t stage_full[STAGES], stage_empty[STAGES], s_full[NG], s_free[NG], p_full[NG], pv_done[NG];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ot_base = sbase + STAGES * STAGE_BYTES;                            // [2] output staging tiles
    float* tab = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)) + STAGES * STAGE_BYTES + 2 * OT_BYTES);
    const int groups = C / 64;                                                         // stages per tile
    long long my_tiles = 0;
    if ((long long)blockIdx.x < ntiles) my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const long long n_stages = my_tiles * groups;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&stage_full[s], 1); mbar_init(&stage_empty[s], 1); }
        for (int g = 0; g < NG; ++g) { mbar_init(&s_full[g], 1); mbar_init(&s_free[g], ROWS); mbar_init(&p_full[g], ROWS); mbar_init(&pv_done[g], 1); }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    {
        const int n4 = tab_copy_stride(heads);          // 4 * stride floats = stride float4
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
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                WinBox wb[2];
                wb[0] = win_box(geo, 2 * tile);
                wb[1] = win_box(geo, 2 * tile + 1);
                if (!wb[1].valid) wb[1] = wb[0];                            // odd tail: the second half of the tile repeats window 0
                for (int gi = 0; gi < groups; ++gi) {
                    if (round > 0) mbar_wait(&stage_empty[stage], (uint32_t)((round - 1) & 1));
                    tma::expect_tx(&stage_full[stage], STAGE_BYTES);
                    const uint32_t st = sbase + stage * STAGE_BYTES;
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        const WinBox b = wb[w];
                        if (!b.wrap_x && !b.wrap_y) {
#pragma unroll
                            for (int op = 0; op < 3; ++op)
                                tma::load_3d(st + op * OPERAND_BYTES + w * WIN_BYTES, &in_maps.full, &stage_full[stage], op * C + gi * 64, b.x0, b.yg_base + b.y0);
                        } else {
                            for (int ty = 0; ty < WS; ++ty) {
                                int ys = b.y0 + ty; if (ys >= geo.H) ys -= geo.H;
                                const int yg = b.yg_base + ys;
#pragma unroll
                                for (int op = 0; op < 3; ++op) {
                                    const uint32_t d = st + op * OPERAND_BYTES + w * WIN_BYTES + ty * 1024;
                                    if (b.wrap_x) {
                                        tma::load_3d(d, &in_maps.row_a, &stage_full[stage], op * C + gi * 64, b.x0, yg);
                                        tma::load_3d(d + (WS - geo.shift) * 128, &in_maps.row_b, &stage_full[stage], op * C + gi * 64, 0, yg);
                                    } else {
                                        tma::load_3d(d, &in_maps.row8, &stage_full[stage], op * C + gi * 64, b.x0, yg);
                                    }
                                }
                            }
                        }
                    }
                    if (++stage == STAGES) { stage = 0; ++round; }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // =============================================================== MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc_s = idesc_bf16(ROWS, NTOK, false, false);
            constexpr uint32_t idesc_o = idesc_bf16(ROWS, HD, false, true);
            const uint64_t d0 = tma::desc_sw128(sbase);
            const int nt = (int)(n_stages * NG);                              // units = (stage, group)
            // ---- cursor of the next unit whose scores are to be issued
            int qn = 0, q_stage = 0, q_stage_par = 0, q_g = 0, q_k = 0;
            auto issue_qk = [&]() {
                if (q_g == 0) mbar_wait(&stage_full[q_stage], (uint32_t)q_stage_par);
                if (q_k > 0) mbar_wait(&s_free[q_g], (uint32_t)((q_k - 1) & 1));
                fence_after_sync();
#pragma unroll
                for (int hh = 0; hh < HPB; ++hh) {
                    const uint32_t off = (uint32_t)(q_stage * STAGE_BYTES + (q_g * HPB + hh) * (HD * 2)) >> 4;
                    const uint64_t qd = d0 + off, kd = d0 + off + (OPERAND_BYTES >> 4);
                    const uint32_t d = tm + TM_S + q_g * (HPB * 64) + hh * 64;
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks)        // lanes 0-63: window 0 of the pair
                        mma_ss_masked(d, qd + 2 * ks, kd + 2 * ks, idesc_s, ks > 0, 0u, 0u, ALL, ALL);
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks)        // lanes 64-127: window 1 (its K rows follow window 0's)
                        mma_ss_masked(d, qd + 2 * ks, kd + (WIN_BYTES >> 4) + 2 * ks, idesc_s, ks > 0, ALL, ALL, 0u, 0u);
                }
                mma_commit(&s_full[q_g]);
                ++qn;
                if (++q_g == NG) { q_g = 0; ++q_k; if (++q_stage == STAGES) { q_stage = 0; q_stage_par ^= 1; } }
            };
            for (int i = 0; i < NG && i < nt; ++i) issue_qk();
            int stage = 0, g = 0, k = 0;
            for (int n = 0; n < nt; ++n) {
                // S[g] is free as soon as softmax(n) has copied it to registers, so the group's next scores are
                // computed while softmax(n) is still running
                if (qn < nt) issue_qk();
                mbar_wait(&p_full[g], (uint32_t)(k & 1));
                fence_after_sync();
#pragma unroll
                for (int hh = 0; hh < HPB; ++hh) {
                    const uint64_t vd = d0 + ((uint32_t)(stage * STAGE_BYTES + 2 * OPERAND_BYTES + (g * HPB + hh) * (HD * 2)) >> 4);
                    const uint32_t d = tm + TM_O + g * 32 + hh * HD, a = tm + TM_P + g * (HPB * 32) + hh * 32;
#pragma unroll
                    for (int ks = 0; ks < NTOK / 16; ++ks)      // window 0: keys = rows 0-63 of the V tile, 16 keys (2048 B) per step
                        mma_ts_masked(d, a + ks * 8, vd + ks * (2048 >> 4), idesc_o, ks > 0, 0u, 0u, ALL, ALL);
#pragma unroll
                    for (int ks = 0; ks < NTOK / 16; ++ks)
                        mma_ts_masked(d, a + ks * 8, vd + (WIN_BYTES >> 4) + ks * (2048 >> 4), idesc_o, ks > 0, ALL, ALL, 0u, 0u);
                }
                mma_commit(&pv_done[g]);
                if (++g == NG) { g = 0; ++k; mma_commit(&stage_empty[stage]); if (++stage == STAGES) stage = 0; }
            }
        }
    } else {
        // ====================================================== softmax + epilogue groups
        const int g = warp >> 2;                       // group g takes heads [g*HPB, (g+1)*HPB) of every stage
        const int row = tid & 127;                     // TMEM lane = 64 * (window of the pair) + query token
        const int wsel = row >> 6, ti = row & 63, ty = ti >> 3, tx = ti & 7;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tm + TM_S + g * (HPB * 64) + lane_addr;
        const uint32_t tP = tm + TM_P + g * (HPB * 32) + lane_addr;
        const uint32_t tO = tm + TM_O + g * 32 + lane_addr;
        const float c = scale * LOG2E, mv2 = mask_value * LOG2E;
        const uint64_t c2 = pack2(c, c);
        // bias row of key row yj = 0 for this thread: copy (7 - tx) % 4 at entry (7 - tx) - copy (0 or 4), table row dy = ty + 7
        const int r0 = WS - 1 - tx, cp = r0 & 3;
        const float* tab_row = tab + cp * tab_copy_stride(heads) + (ty + WS - 1) * TAB_ROW + (r0 - cp);
        const int s_ = geo.shift;
        const uint64_t yhi = s_ > 0 ? (~0ull << (8 * (WS - s_))) : 0ull;                       // keys with ty >= ws - shift
        const uint64_t xhi = s_ > 0 ? 0x0101010101010101ull * (uint64_t)((0xFFu << (WS - s_)) & 0xFFu) : 0ull;  // tx >= ws - shift
        const uint32_t srow = (uint32_t)row * 128, sw = (uint32_t)(row & 7);
        long long cur_tile_it = -1;
        uint64_t mbits = 0;
        bool any_mask = false;
        float prev_inv[HPB] = {};
        WinBox ob[2];                                  // thread 0: destination boxes of the tile being stored

        // Epilogue of stage `ps` (0-based among this CTA's stages): O / rowsum of this group's heads -> staging tile; when both
        // groups have delivered, one thread stores the tile's two windows with TMA.
        auto epilogue = [&](long long ps) {
            const uint32_t tile_s = ot_base + (uint32_t)(ps & 1) * OT_BYTES;
#pragma unroll
            for (int hh = 0; hh < HPB; ++hh) {
                uint32_t o[HD];
                if constexpr (HD == 16) tmem_ld16(tO + hh * HD, o); else tmem_ld32(tO + hh * HD, o);
                tmem_wait_ld();
                const float inv = prev_inv[hh];
#pragma unroll
                for (int j = 0; j < HD; j += 8) {
                    const uint32_t chunk = (uint32_t)(((g * HPB + hh) * HD + j) >> 3) ^ sw;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_s + srow + (chunk << 4)),
                                 "r"(pack_bf16(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv)),
                                 "r"(pack_bf16(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv)),
                                 "r"(pack_bf16(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv)),
                                 "r"(pack_bf16(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv)) : "memory");
                }
            }
            fence_proxy_async();                                          // staging writes -> visible to the TMA store
            if (tid == 0) tma::store_wait_read<0>();                      // the previous stage's store has left its staging tile
            asm volatile("bar.sync 2, 256;" ::: "memory");               // both groups' columns of the stage tile are in smem
            if (tid == 0) {
                const long long tile_it = ps / groups;
                const int gi = (int)(ps - tile_it * groups);
                const long long tile = (long long)blockIdx.x + tile_it * gridDim.x;
                if (gi == 0) { ob[0] = win_box(geo, 2 * tile); ob[1] = win_box(geo, 2 * tile + 1); }
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const WinBox b = ob[w];
                    if (!b.valid) continue;
                    const uint32_t src = tile_s + w * WIN_BYTES;
                    if (!b.wrap_x && !b.wrap_y) {
                        tma::store_3d(&out_maps.full, src, gi * 64, b.x0, b.yg_base + b.y0);
                    } else {
                        for (int r = 0; r < WS; ++r) {
                            int ys = b.y0 + r; if (ys >= geo.H) ys -= geo.H;
                            const int yg = b.yg_base + ys;
                            if (b.wrap_x) {
                                tma::store_3d(&out_maps.row_a, src + r * 1024, gi * 64, b.x0, yg);
                                tma::store_3d(&out_maps.row_b, src + r * 1024 + (WS - geo.shift) * 128, gi * 64, 0, yg);
                            } else {
                                tma::store_3d(&out_maps.row8, src + r * 1024, gi * 64, b.x0, yg);
                            }
                        }
                    }
                }
                tma::store_commit();
            }
        };

        int gi = 0;                                    // stage within the tile
        long long tile_it = 0;
        for (long long n = 0; n < n_stages; ++n) {
            if (tile_it != cur_tile_it) {              // new window pair: shifted-window mask bits of this thread's window
                cur_tile_it = tile_it;
                mbits = 0;
                const long long wdx = 2 * ((long long)blockIdx.x + tile_it * gridDim.x) + wsel;
                if (s_ > 0 && wdx < geo.total_windows) {
                    const int win = (int)(wdx % geo.nW);
                    const int wy = win / geo.nww, wx = win - wy * geo.nww;
                    if (wy == geo.nwh - 1) mbits |= (ty >= WS - s_) ? ~yhi : yhi;
                    if (wx == geo.nww - 1) mbits |= (tx >= WS - s_) ? ~xhi : xhi;
                }
                any_mask = __any_sync(0xffffffffu, mbits != 0);
            }
            mbar_wait(&s_full[g], (uint32_t)(n & 1));
            fence_after_sync();
            float inv_cur[HPB];
#pragma unroll
            for (int hh = 0; hh < HPB; ++hh) {
                const int h = gi * G + g * HPB + hh;
                uint64_t t[NTOK / 2];
                {
                    uint32_t ra[32], rb[32];
                    tmem_ld32(tS + hh * 64, ra);
                    tmem_ld32(tS + hh * 64 + 32, rb);
                    tmem_wait_ld();
                    if (hh == HPB - 1) { fence_before_sync(); mbar_arrive(&s_free[g]); }    // all of the unit's scores are in registers
                    const float* tb = tab_row + h * TAB_HEAD;
#pragma unroll
                    for (int yj = 0; yj < WS; ++yj) {       // t = s * (scale log2 e) + bias, two scores per FFMA2
                        const float4 b0 = *reinterpret_cast<const float4*>(tb - yj * TAB_ROW);
                        const float4 b1 = *reinterpret_cast<const float4*>(tb - yj * TAB_ROW + 4);
                        const uint32_t* r = yj < 4 ? ra + yj * 8 : rb + (yj - 4) * 8;
                        t[yj * 4 + 0] = ffma2(pack2(__uint_as_float(r[0]), __uint_as_float(r[1])), c2, pack2(b0.x, b0.y));
                        t[yj * 4 + 1] = ffma2(pack2(__uint_as_float(r[2]), __uint_as_float(r[3])), c2, pack2(b0.z, b0.w));
                        t[yj * 4 + 2] = ffma2(pack2(__uint_as_float(r[4]), __uint_as_float(r[5])), c2, pack2(b1.x, b1.y));
                        t[yj * 4 + 3] = ffma2(pack2(__uint_as_float(r[6]), __uint_as_float(r[7])), c2, pack2(b1.z, b1.w));
                    }
                }
                if (any_mask) {                        // warp-uniform: only windows of the last window row / column
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
                {
                    float a, b;
                    unpack2(fadd2(sum2[0], sum2[1]), a, b);
                    inv_cur[hh] = 1.f / (a + b);
                }
                if (hh == 0 && n > 0) {                // previous stage of this group: its P / O columns are free again
                    mbar_wait(&pv_done[g], (uint32_t)((n - 1) & 1));
                    fence_after_sync();
                    epilogue(n - 1);
                }
                tmem_st32(tP + hh * 32, pk);
            }
            tmem_wait_st();
            fence_before_sync();
            mbar_arrive(&p_full[g]);
#pragma unroll
            for (int hh = 0; hh < HPB; ++hh) prev_inv[hh] = inv_cur[hh];
            if (++gi == groups) { gi = 0; ++tile_it; }
        }
        if (n_stages > 0) {
            mbar_wait(&pv_done[g], (uint32_t)((n_stages - 1) & 1));
            fence_after_sync();
            epilogue(n_stages - 1);
            if (tid == 0) tma::store_wait_all();
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_slot, 512);
}

size_t win8_table_bytes(int heads) { return (size_t)4 * tab_copy_stride(heads) * sizeof(float); }

size_t win8_smem_bytes(int heads) { return (size_t)STAGES * STAGE_BYTES + 2 * OT_BYTES + win8_table_bytes(heads) + 1024; }

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

// Writes the shared-memory image of the bias table (4 shifted, x-reversed, log2(e)-scaled copies) into `workspace`
int window_attn_win8_prepare(const float* table, void* workspace, int heads, cudaStream_t stream) {
    const int n = 4 * tab_copy_stride(heads);
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
    const long long ntiles = (geo.total_windows + 1) / 2;
    Maps in_maps, out_maps;
    if (!make_maps(&in_maps, qkv, B, H, W, 3 * C, shift, false) || !make_maps(&out_maps, out, B, H, W, C, shift, true)) return SODT_ERR_CUDA;
    const size_t smem = win8_smem_bytes(heads);
    const int hd = C / heads;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    cudaError_t e;
    if (hd == 16) {
        auto kern = window_attn_win8_kernel<16>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
        kern<<<grid, NTHREADS, smem, stream>>>(in_maps, out_maps, static_cast<const float*>(workspace), geo, C, heads, scale, mask_value, ntiles);
    } else {
        auto kern = window_attn_win8_kernel<32>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
        kern<<<grid, NTHREADS, smem, stream>>>(in_maps, out_maps, static_cast<const float*>(workspace), geo, C, heads, scale, mask_value, ntiles);
    }
    return check_launch();
}

}  // namespace sodt
