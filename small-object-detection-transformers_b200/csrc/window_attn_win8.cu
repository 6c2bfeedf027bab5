// 8x8-window attention (N = 64 tokens, head_dim 16 or 32) on tcgen05 tensor cores: the kernel of the
// backbone's stage 1 and stage 2 (basics/models/backbone_vit.py:114-145; shift 0 or any 0 < shift < 8).
//
// A tile is a PAIR of windows (128 query rows): per head
//   S[128x128] = Q_pair K_pair^T   tcgen05.mma SS (K-major operands) -> TMEM; only the two diagonal 64x64
//                                  blocks are meaningful and only they are read back
//   softmax                        one thread per query row: tcgen05.ld of its 64 useful scores, relative
//                                  position bias (closed-form index, bank-conflict-free padded table in shared
//                                  memory), shifted-window mask evaluated from region bit masks, exp2, row sum;
//                                  P written to TMEM as packed bf16 (off-diagonal blocks stay zero)
//   O[128xhd] = P V_pair           tcgen05.mma TS (A = P from TMEM, B = V MN-major), K = 128 keys
//   epilogue                       O / rowsum -> bf16 -> straight to the un-rolled, un-partitioned output image
// Roll, partition, reverse partition and reverse roll are address arithmetic in the producer / epilogue.
//
// Persistent CTAs (one per SM), 10 warps: warps 0-3 and 4-7 are two softmax groups that take even / odd heads
// (each with its own S, P and O buffers in TMEM, so one group's softmax overlaps the other's MMAs), warp 8
// streams q/k/v of G heads at a time through a 3-stage cp.async ring, warp 9 issues the MMAs.
#include "common.cuh"
#include "tc05.cuh"

namespace sodt {
namespace {

using namespace tc;

constexpr int WS = 8;
constexpr int NTOK = 64;                 // tokens per window
constexpr int ROWS = 128;                // rows per tile = 2 windows
constexpr int NTHREADS = 320;
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 3 * ROWS * 128;   // q,k,v x 128 tokens x 128 B (= G heads x hd x 2 B)
constexpr int CHUNK_STRIDE = ROWS * 16;       // bytes between 8-element chunks of a canonical tile
constexpr int TAB_LD = 40;                    // padded row stride of the bias table (bank-conflict free)
constexpr int TAB_ENTRIES = (2 * WS - 1) * TAB_LD;   // 600 floats per head
constexpr float LOG2E = 1.4426950408889634f;

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
    constexpr int CPH = HD / 8;               // 16-byte chunks per head row
    constexpr int TILE_BYTES = ROWS * HD * 2; // one (q|k|v, head) tile
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t stage_full[STAGES], stage_empty[STAGES], s_full[2], s_free[2], p_full[2], pv_done[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = smem_u32(smem);
    float* tab = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
    const int C3 = 3 * C;
    const int groups = heads / G;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&stage_full[s], 32); mbar_init(&stage_empty[s], 1); }
        for (int p = 0; p < 2; ++p) { mbar_init(&s_full[p], 1); mbar_init(&s_free[p], ROWS); mbar_init(&p_full[p], ROWS); mbar_init(&pv_done[p], 1); }
        fence_barrier_init();
    }
    if (warp == 9) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    for (int e = tid; e < heads * TAB_ENTRIES; e += NTHREADS) tab[e] = table_p[e];
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_slot;
    // TMEM columns: S[p] = p*128 (128 cols), P[p] = 256 + p*64 (64 cols), O[p] = 384 + p*32 (HD cols)

    if (warp == 8) {
        // ================================================================ producer
        const int tsub = lane & 7, csub = lane >> 3;
        long long gseq = 0;
        uint64_t* pending = nullptr;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            long long tok[16];
#pragma unroll
            for (int oct = 0; oct < 16; ++oct) {
                long long wdx = 2 * tile + (oct >> 3);
                if (wdx >= geo.total_windows) wdx = 2 * tile;      // odd tail: duplicate the first window
                tok[oct] = geo.token(wdx, oct & 7, tsub) * C3;
            }
            for (int gi = 0; gi < groups; ++gi, ++gseq) {
                const int s = (int)(gseq % STAGES);
                if (gseq >= STAGES) mbar_wait(&stage_empty[s], (uint32_t)((gseq / STAGES - 1) & 1));
                const uint32_t st = sbase + s * STAGE_BYTES;
#pragma unroll
                for (int which = 0; which < 3; ++which) {
                    const int col = which * C + gi * 64;
#pragma unroll
                    for (int oct = 0; oct < 16; ++oct) {
                        const __nv_bfloat16* src = qkv + tok[oct] + col;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const int j = csub + 4 * half;          // chunk 0..7 of the 128-byte segment
                            const int g = j / CPH, c = j % CPH;     // head in group, chunk in head
                            cp_async16(st + (which * G + g) * TILE_BYTES + c * CHUNK_STRIDE + (oct * 8 + tsub) * 16, src + j * 8);
                        }
                    }
                }
                cp_async_commit();
                if (pending) { cp_async_wait<1>(); fence_proxy_async(); mbar_arrive(pending); }
                pending = &stage_full[s];
            }
        }
        if (pending) { cp_async_wait<0>(); fence_proxy_async(); mbar_arrive(pending); }
    } else if (warp == 9) {
        // =============================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_s = idesc_bf16(ROWS, ROWS, false, false);
            constexpr uint32_t idesc_o = idesc_bf16(ROWS, HD, false, true);
            long long my_tiles = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) ++my_tiles;
            const long long n_total = my_tiles * heads;
            auto issue_qk = [&](long long n) {
                const long long gq = n / G;
                const int hg = (int)(n - gq * G), s = (int)(gq % STAGES);
                if (hg == 0) mbar_wait(&stage_full[s], (uint32_t)((gq / STAGES) & 1));
                const int p = (int)(n & 1);
                const long long k = n >> 1;
                if (k > 0) mbar_wait(&s_free[p], (uint32_t)((k - 1) & 1));
                fence_proxy_async();
                fence_after_sync();
                const uint32_t qt = sbase + s * STAGE_BYTES + (0 * G + hg) * TILE_BYTES;
                const uint32_t kt = sbase + s * STAGE_BYTES + (1 * G + hg) * TILE_BYTES;
#pragma unroll
                for (int ks = 0; ks < HD / 16; ++ks)
                    mma_ss(tm + p * 128, smem_desc(qt + ks * 2 * CHUNK_STRIDE, CHUNK_STRIDE, 128),
                           smem_desc(kt + ks * 2 * CHUNK_STRIDE, CHUNK_STRIDE, 128), idesc_s, ks > 0);
                mma_commit(&s_full[p]);
            };
            if (n_total > 0) issue_qk(0);
            if (n_total > 1) issue_qk(1);
            for (long long n = 0; n < n_total; ++n) {
                const int p = (int)(n & 1);
                const long long k = n >> 1, gq = n / G;
                const int hg = (int)(n - gq * G), s = (int)(gq % STAGES);
                // S[p] is free as soon as softmax(n) has copied it to registers (early in its work), so the
                // scores of the group's next head are computed while softmax(n) is still running
                if (n + 2 < n_total) issue_qk(n + 2);
                mbar_wait(&p_full[p], (uint32_t)(k & 1));
                fence_after_sync();
                const uint32_t vt = sbase + s * STAGE_BYTES + (2 * G + hg) * TILE_BYTES;
#pragma unroll
                for (int ks = 0; ks < ROWS / 16; ++ks)
                    mma_ts(tm + 384 + p * 32, tm + 256 + p * 64 + ks * 8, smem_desc(vt + ks * 256, 128, CHUNK_STRIDE), idesc_o, ks > 0);
                mma_commit(&pv_done[p]);
                if (hg == G - 1) mma_commit(&stage_empty[s]);
            }
        }
    } else {
        // ====================================================== softmax + epilogue groups
        const int p = warp >> 2;                       // group 0: even heads, group 1: odd heads
        const int row = tid & 127;                     // TMEM lane == query row of the pair
        const int wh = row >> 6, ti = row & 63, ty = ti >> 3, tx = ti & 7;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tm + p * 128 + wh * 64 + lane_addr;
        const uint32_t tP = tm + 256 + p * 64 + wh * 32 + lane_addr;
        const uint32_t tO = tm + 384 + p * 32 + lane_addr;
        {   // the off-diagonal half of P stays zero for the whole kernel
            uint32_t z[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) z[j] = 0u;
            tmem_st32(tm + 256 + p * 64 + (wh ^ 1) * 32 + lane_addr, z);
            tmem_wait_st();
        }
        const float c = scale * LOG2E, mv2 = mask_value * LOG2E;
        const float* tab_row = tab + (ty + WS - 1) * TAB_LD + (tx + WS - 1);
        const int s_ = geo.shift;
        const uint64_t yhi = s_ > 0 ? (~0ull << (8 * (WS - s_))) : 0ull;                       // keys with ty >= ws - shift
        const uint64_t xhi = s_ > 0 ? 0x0101010101010101ull * (uint64_t)((0xFFu << (WS - s_)) & 0xFFu) : 0ull;  // tx >= ws - shift
        const int nwh = geo.nW / geo.nww;
        long long k = 0;
        __nv_bfloat16* prev_dst = nullptr;
        float prev_inv = 0.f;
        auto epilogue = [&](__nv_bfloat16* dst, float inv) {
            uint32_t o[HD];
            if constexpr (HD == 16) { uint32_t (&o16)[16] = reinterpret_cast<uint32_t (&)[16]>(o); tmem_ld16(tO, o16); }
            else { uint32_t (&o32)[32] = reinterpret_cast<uint32_t (&)[32]>(o); tmem_ld32(tO, o32); }
            tmem_wait_ld();
            if (dst != nullptr) {
#pragma unroll
                for (int j = 0; j < HD; j += 8) {
                    uint4 v;
                    v.x = pack_bf16(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv);
                    v.y = pack_bf16(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
                    v.z = pack_bf16(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv);
                    v.w = pack_bf16(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv);
                    *reinterpret_cast<uint4*>(dst + j) = v;
                }
            }
        };
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long wdx = 2 * tile + wh;
            const bool valid = wdx < geo.total_windows;
            __nv_bfloat16* out_tok = nullptr;
            uint64_t mbits = 0;
            if (valid) {
                out_tok = out + geo.token(wdx, ty, tx) * C;
                if (s_ > 0) {
                    const int win = (int)(wdx % geo.nW);
                    const int wy = win / geo.nww, wx = win - wy * geo.nww;
                    if (wy == nwh - 1) mbits |= (ty >= WS - s_) ? ~yhi : yhi;
                    if (wx == geo.nww - 1) mbits |= (tx >= WS - s_) ? ~xhi : xhi;
                }
            }
            const bool any_mask = __any_sync(0xffffffffu, mbits != 0);
            for (int h = p; h < heads; h += 2, ++k) {
                mbar_wait(&s_full[p], (uint32_t)(k & 1));
                fence_after_sync();
                float s2[NTOK];
                {
                    uint32_t r0[32], r1[32];
                    tmem_ld32(tS, r0);
                    tmem_ld32(tS + 32, r1);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) { s2[j] = __uint_as_float(r0[j]); s2[32 + j] = __uint_as_float(r1[j]); }
                }
                fence_before_sync();
                mbar_arrive(&s_free[p]);
                const float* tb = tab_row + h * TAB_ENTRIES;
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < NTOK; ++j) {
                    s2[j] = fmaf(s2[j], c, tb[-((j >> 3) * TAB_LD + (j & 7))]);
                    if (any_mask && ((mbits >> j) & 1ull)) s2[j] += mv2;
                    mx = fmaxf(mx, s2[j]);
                }
                float sum = 0.f;
                uint32_t pk[32];
#pragma unroll
                for (int j = 0; j < NTOK; j += 2) {
                    const float p0 = fast_exp2(s2[j] - mx), p1 = fast_exp2(s2[j + 1] - mx);
                    sum += p0 + p1;
                    pk[j >> 1] = pack_bf16(p0, p1);
                }
                if (k > 0) {                                   // previous head of this group: P / O buffers free again
                    mbar_wait(&pv_done[p], (uint32_t)((k - 1) & 1));
                    fence_after_sync();
                    epilogue(prev_dst, prev_inv);
                }
                tmem_st32(tP, pk);
                tmem_wait_st();
                fence_before_sync();
                mbar_arrive(&p_full[p]);
                prev_dst = valid ? out_tok + h * HD : nullptr;
                prev_inv = 1.f / sum;
            }
        }
        if (k > 0) {
            mbar_wait(&pv_done[p], (uint32_t)((k - 1) & 1));
            fence_after_sync();
            epilogue(prev_dst, prev_inv);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_slot, 512);
}

}  // namespace

size_t window_attn_win8_workspace(int heads) { return (size_t)heads * TAB_ENTRIES * sizeof(float); }

bool window_attn_win8_supported(int H, int W, int C, int heads, int ws, int shift, int dtype) {
    if (dtype != SODT_BF16 || ws != WS || H % WS || W % WS || heads <= 0 || C % heads) return false;
    const int hd = C / heads;
    if (hd != 16 && hd != 32) return false;
    const int G = 64 / hd;
    if (heads % G || heads % 2) return false;
    return STAGES * STAGE_BYTES + (size_t)heads * TAB_ENTRIES * sizeof(float) <= 220 * 1024;
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
    const size_t smem = STAGES * STAGE_BYTES + (size_t)heads * TAB_ENTRIES * sizeof(float);
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
