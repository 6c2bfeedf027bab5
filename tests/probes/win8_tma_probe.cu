// Standalone probe (run on a B200) of the building blocks of the TMA-fed 8x8-window attention kernel:
//   L  window boxes of the [B*H, W, 3C] qkv image -> SWIZZLE_128B operand tiles: one (64 ch, 8, 8) box for an interior window,
//      per-row part boxes (64, 8-s, 1) + (64, s, 1) for a window that wraps around the image in x and y (roll as addressing)
//   S  scores of one head: two lane-masked tcgen05.mma (lanes 0-63 = window 0, 64-127 = window 1), K-major SWIZZLE_128B
//      operands addressed at the head's column offset inside the 128-byte rows
//   O  O = P V: A = P from TMEM, B = V MN-major SWIZZLE_128B at the head's column offset (vmode 0: N = head_dim,
//      vmode 1: N = 64 = all channels of the tile), lane-masked per window
//   T  output tile staged in SWIZZLE_128B shared memory -> TMA stores (full box / row parts) into the un-rolled image
// Usage: win8_tma_probe <hd: 16|32> <vmode: 0|1>      prints max abs errors against a host reference.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../small-object-detection-transformers_b200/csrc/tma.cuh"

using namespace sodt::tc;
namespace tma = sodt::tma;

constexpr int H = 16, W = 16, C = 64, C3 = 3 * C, SHIFT = 2, WS = 8;
constexpr int OP_BYTES = 128 * 128;          // one operand: 128 rows (2 windows x 64 tokens) x 64 ch bf16
constexpr uint32_t ALL = 0xFFFFFFFFu;

struct Params { int hd, vmode; };

__global__ void __launch_bounds__(160) probe(const __grid_constant__ CUtensorMap m_full, const __grid_constant__ CUtensorMap m_ra,
                                             const __grid_constant__ CUtensorMap m_rb, const __grid_constant__ CUtensorMap o_full,
                                             const __grid_constant__ CUtensorMap o_ra, const __grid_constant__ CUtensorMap o_rb,
                                             const __grid_constant__ CUtensorMap o_map4, int clip_stores,
                                             __nv_bfloat16* dump, float* S_out, float* O_out, Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar_ld, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t stg = sbase + 3 * OP_BYTES;                       // output staging tile (128 rows x 128 B)
    if (tid == 128) { mbar_init(&bar_ld, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
    if (warp == 4) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_slot;
    // ------------------------------------------------------------------ L: loads
    if (tid == 128) {
        tma::expect_tx(&bar_ld, 3 * OP_BYTES);
        for (int op = 0; op < 3; ++op) {
            const uint32_t d0 = sbase + op * OP_BYTES;
            tma::load_3d(d0, &m_full, &bar_ld, op * C, 0 * WS + SHIFT, 0 * WS + SHIFT);               // window (0,0): interior
            for (int ty = 0; ty < WS; ++ty) {                                                        // window (1,1): wraps in x and y
                const int ys = (1 * WS + ty + SHIFT) % H;
                tma::load_3d(d0 + 8192 + ty * 1024, &m_ra, &bar_ld, op * C, 1 * WS + SHIFT, ys);
                tma::load_3d(d0 + 8192 + ty * 1024 + (WS - SHIFT) * 128, &m_rb, &bar_ld, op * C, 0, ys);
            }
        }
    }
    mbar_wait(&bar_ld, 0);
    for (int e = tid; e < 3 * OP_BYTES / 2; e += blockDim.x) dump[e] = reinterpret_cast<const __nv_bfloat16*>(sm)[e];
    __syncthreads();
    // ------------------------------------------------------------------ S and O per head
    const int hd = p.hd, G = 64 / hd;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t TM_S = 0, TM_P = 64, TM_O = 128;
    uint32_t phase = 0;
    for (int g = 0; g < G; ++g) {
        if (tid == 128) {
            const uint32_t idesc_s = idesc_bf16(128, 64, false, false);
            const uint32_t qa = sbase + g * hd * 2, ka = sbase + OP_BYTES + g * hd * 2;
            for (int ks = 0; ks < hd / 16; ++ks)
                mma_ss_masked(tm + TM_S, tma::desc_sw128(qa + ks * 32), tma::desc_sw128(ka + ks * 32), idesc_s, ks > 0, 0u, 0u, ALL, ALL);
            for (int ks = 0; ks < hd / 16; ++ks)
                mma_ss_masked(tm + TM_S, tma::desc_sw128(qa + ks * 32), tma::desc_sw128(ka + 8192 + ks * 32), idesc_s, ks > 0, ALL, ALL, 0u, 0u);
            mma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, phase); phase ^= 1;
        fence_after_sync();
        if (tid < 128) {
            uint32_t r0[32], r1[32], pk[32];
            tmem_ld32(tm + TM_S + lane_addr, r0);
            tmem_ld32(tm + TM_S + lane_addr + 32, r1);
            tmem_wait_ld();
            for (int j = 0; j < 32; ++j) { S_out[(g * 128 + tid) * 64 + j] = __uint_as_float(r0[j]); S_out[(g * 128 + tid) * 64 + 32 + j] = __uint_as_float(r1[j]); }
            for (int j = 0; j < 32; j += 2) {
                pk[j >> 1] = pack_bf16(__uint_as_float(r0[j]) * 0.125f, __uint_as_float(r0[j + 1]) * 0.125f);
                pk[16 + (j >> 1)] = pack_bf16(__uint_as_float(r1[j]) * 0.125f, __uint_as_float(r1[j + 1]) * 0.125f);
            }
            tmem_st32(tm + TM_P + lane_addr, pk);
            tmem_wait_st();
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        const int ncol = p.vmode == 0 ? hd : 64;
        if (tid == 128) {
            const uint32_t idesc_o = idesc_bf16(128, ncol, false, true);
            const uint32_t va = sbase + 2 * OP_BYTES + (p.vmode == 0 ? g * hd * 2 : 0);
            for (int w = 0; w < 2; ++w)
                for (int ks = 0; ks < 4; ++ks)
                    mma_ts_masked(tm + TM_O, tm + TM_P + ks * 8, tma::desc_sw128(va + w * 8192 + ks * 2048), idesc_o, ks > 0,
                                  w == 0 ? 0u : ALL, w == 0 ? 0u : ALL, w == 0 ? ALL : 0u, w == 0 ? ALL : 0u);
            mma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, phase); phase ^= 1;
        fence_after_sync();
        if (tid < 128) {
            uint32_t o[32];
            tmem_ld32(tm + TM_O + lane_addr + (p.vmode == 0 ? 0 : g * hd), o);
            tmem_wait_ld();
            for (int j = 0; j < hd; ++j) O_out[(g * 128 + tid) * 32 + j] = __uint_as_float(o[j]);
            // T: stage the head's output columns (bf16) in the SWIZZLE_128B tile
            for (int j = 0; j < hd; j += 8) {
                const uint32_t chunk = (uint32_t)((g * hd + j) >> 3) ^ (uint32_t)(tid & 7);
                const uint32_t dst = stg + tid * 128 + chunk * 16;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                             "r"(pack_bf16(__uint_as_float(o[j]), __uint_as_float(o[j + 1]))), "r"(pack_bf16(__uint_as_float(o[j + 2]), __uint_as_float(o[j + 3]))),
                             "r"(pack_bf16(__uint_as_float(o[j + 4]), __uint_as_float(o[j + 5]))), "r"(pack_bf16(__uint_as_float(o[j + 6]), __uint_as_float(o[j + 7]))) : "memory");
            }
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
    }
    // ------------------------------------------------------------------ T: stores
    fence_proxy_async();
    __syncthreads();
    if (tid == 128) {
        tma::store_3d(&o_full, stg, 0, SHIFT, SHIFT);
        if (clip_stores) {        // the wrapped window as four full boxes, the out-of-image parts clipped by the TMA unit
            const int x0 = WS + SHIFT, y0 = WS + SHIFT;
            tma::store_4d(&o_map4, stg + 8192, 0, x0, y0, 0);
            tma::store_4d(&o_map4, stg + 8192, 0, x0 - W, y0, 0);
            tma::store_4d(&o_map4, stg + 8192, 0, x0, y0 - H, 0);
            tma::store_4d(&o_map4, stg + 8192, 0, x0 - W, y0 - H, 0);
        } else
        for (int ty = 0; ty < WS; ++ty) {
            const int ys = (WS + ty + SHIFT) % H;
            tma::store_3d(&o_ra, stg + 8192 + ty * 1024, 0, WS + SHIFT, ys);
            tma::store_3d(&o_rb, stg + 8192 + ty * 1024 + (WS - SHIFT) * 128, 0, 0, ys);
        }
        tma::store_commit();
        tma::store_wait_all();
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_slot, 512);
}

static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

int main(int argc, char** argv) {
    Params p{argc > 1 ? atoi(argv[1]) : 16, argc > 2 ? atoi(argv[2]) : 0};
    const int clip_stores = argc > 3 ? atoi(argv[3]) : 0;
    std::vector<__nv_bfloat16> qkv((size_t)H * W * C3);
    std::vector<float> qf(qkv.size());
    unsigned s = 12345u;
    for (size_t i = 0; i < qkv.size(); ++i) {
        s = s * 1664525u + 1013904223u;
        qf[i] = (float)((int)((s >> 20) % 17) - 8) / 8.f;       // multiples of 1/8 in [-1, 1]: exact in bf16, exact fp32 sums
        qkv[i] = __float2bfloat16_rn(qf[i]);
    }
    __nv_bfloat16 *d_qkv, *d_out, *d_dump;
    float *d_S, *d_O;
    CK(cudaMalloc(&d_qkv, qkv.size() * 2));
    CK(cudaMalloc(&d_out, (size_t)H * W * C * 2));
    CK(cudaMalloc(&d_dump, 3 * OP_BYTES));
    CK(cudaMalloc(&d_S, 4 * 128 * 64 * 4));
    CK(cudaMalloc(&d_O, 4 * 128 * 32 * 4));
    CK(cudaMemcpy(d_qkv, qkv.data(), qkv.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0, (size_t)H * W * C * 2));
    CUtensorMap m_full, m_ra, m_rb, o_full, o_ra, o_rb, o_map4;
    {
        const long long dims[3] = {C3, W, H}, strides[2] = {C3, (long long)W * C3};
        const int bf_[3] = {64, 8, 8}, ba[3] = {64, WS - SHIFT, 1}, bb[3] = {64, SHIFT, 1};
        const long long odims[3] = {C, W, H}, ostr[2] = {C, (long long)W * C};
        if (!tma::make_map_bf16(&m_full, d_qkv, 3, dims, strides, bf_) || !tma::make_map_bf16(&m_ra, d_qkv, 3, dims, strides, ba) ||
            !tma::make_map_bf16(&m_rb, d_qkv, 3, dims, strides, bb) ||
            !tma::make_map_bf16(&o_full, d_out, 3, odims, ostr, bf_, CU_TENSOR_MAP_L2_PROMOTION_NONE) ||
            !tma::make_map_bf16(&o_ra, d_out, 3, odims, ostr, ba, CU_TENSOR_MAP_L2_PROMOTION_NONE) ||
            !tma::make_map_bf16(&o_rb, d_out, 3, odims, ostr, bb, CU_TENSOR_MAP_L2_PROMOTION_NONE)) { printf("tensor map encode failed\n"); return 2; }
    }
    {
        const long long d4[4] = {C, W, H, 1}, s4[3] = {C, (long long)W * C, (long long)H * W * C};
        const int b4[4] = {64, 8, 8, 1};
        if (!tma::make_map_bf16(&o_map4, d_out, 4, d4, s4, b4, CU_TENSOR_MAP_L2_PROMOTION_NONE)) { printf("map4 encode failed\n"); return 2; }
    }
    const size_t smem = 4 * OP_BYTES + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe<<<1, 160, smem>>>(m_full, m_ra, m_rb, o_full, o_ra, o_rb, o_map4, clip_stores, d_dump, d_S, d_O, p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<__nv_bfloat16> dump(3 * OP_BYTES / 2), out((size_t)H * W * C);
    std::vector<float> S(4 * 128 * 64), O(4 * 128 * 32);
    CK(cudaMemcpy(dump.data(), d_dump, 3 * OP_BYTES, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out.data(), d_out, out.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(S.data(), d_S, S.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(O.data(), d_O, O.size() * 4, cudaMemcpyDeviceToHost));
    // token (window w of the pair, ti) -> source pixel of the un-rolled image
    auto src = [&](int w, int ti, int& ys, int& xs) {
        const int wy = w, wx = w;                                  // pair = windows (0,0) and (1,1)
        ys = (wy * WS + ti / 8 + SHIFT) % H;
        xs = (wx * WS + ti % 8 + SHIFT) % W;
    };
    auto val = [&](int op, int w, int ti, int ch) { int ys, xs; src(w, ti, ys, xs); return qf[((size_t)ys * W + xs) * C3 + op * C + ch]; };
    // L
    int bad_l[2] = {0, 0};
    for (int op = 0; op < 3; ++op)
        for (int w = 0; w < 2; ++w)
            for (int ti = 0; ti < 64; ++ti)
                for (int ch = 0; ch < 64; ++ch) {
                    const size_t byte = (size_t)op * OP_BYTES + w * 8192 + ti * 128 + ((((ch >> 3) ^ (ti & 7))) << 4) + (ch & 7) * 2;
                    if (__bfloat162float(dump[byte / 2]) != val(op, w, ti, ch)) ++bad_l[w];
                }
    printf("L  full-box window: %d mismatches; row-part (wrapped) window: %d mismatches (of %d each)\n", bad_l[0], bad_l[1], 3 * 64 * 64);
    // S, O
    const int hd = p.hd, G = 64 / hd;
    double es = 0, eo = 0;
    std::vector<float> Oref((size_t)G * 128 * 32, 0.f);
    for (int g = 0; g < G; ++g)
        for (int w = 0; w < 2; ++w)
            for (int ti = 0; ti < 64; ++ti) {
                float srow[64];
                for (int j = 0; j < 64; ++j) {
                    float a = 0.f;
                    for (int c = 0; c < hd; ++c) a += val(0, w, ti, g * hd + c) * val(1, w, j, g * hd + c);
                    srow[j] = a;
                    es = fmax(es, fabs(a - S[(size_t)(g * 128 + w * 64 + ti) * 64 + j]));
                }
                for (int c = 0; c < hd; ++c) {
                    float a = 0.f;
                    for (int j = 0; j < 64; ++j) a += bf(srow[j] * 0.125f) * val(2, w, j, g * hd + c);
                    Oref[(size_t)(g * 128 + w * 64 + ti) * 32 + c] = a;
                    eo = fmax(eo, fabs(a - O[(size_t)(g * 128 + w * 64 + ti) * 32 + c]));
                }
            }
    printf("S  max abs err %.3e   O  max abs err %.3e   (hd %d, vmode %d)\n", es, eo, hd, p.vmode);
    // T
    int bad_t = 0, untouched_bad = 0;
    std::vector<char> touched((size_t)H * W, 0);
    for (int w = 0; w < 2; ++w)
        for (int ti = 0; ti < 64; ++ti) {
            int ys, xs; src(w, ti, ys, xs);
            touched[ys * W + xs] = 1;
            for (int ch = 0; ch < 64; ++ch) {
                const float want = bf(Oref[(size_t)((ch / hd) * 128 + w * 64 + ti) * 32 + ch % hd]);
                if (__bfloat162float(out[((size_t)ys * W + xs) * C + ch]) != want) ++bad_t;
            }
        }
    for (int px = 0; px < H * W; ++px)
        if (!touched[px]) for (int ch = 0; ch < C; ++ch) if (__bfloat162float(out[(size_t)px * C + ch]) != 0.f) ++untouched_bad;
    printf("T  stored tile: %d mismatches of %d; pixels outside the two windows modified: %d\n", bad_t, 2 * 64 * 64, untouched_bad);
    const bool ok = bad_l[0] == 0 && bad_l[1] == 0 && es < 1e-3 && eo < 1e-2 && bad_t == 0 && untouched_bad == 0;
    printf("%s\n", ok ? "PROBE PASS" : "PROBE FAIL");
    return ok ? 0 : 1;
}
