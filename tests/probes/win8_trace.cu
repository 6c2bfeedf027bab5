// Timeline of the 8x8-window attention kernel's hand-offs (CTA 0): builds window_attn_win8.cu with SODT_WIN8_TRACE and
// prints, per unit, the clock64 deltas between the pipeline events of the softmax groups, the MMA thread and the TMA producer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DSODT_WIN8_TRACE \
//        -o small-object-detection-transformers_b200/build/win8_trace tests/probes/win8_trace.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../small-object-detection-transformers_b200/csrc/window_attn_win8.cu"

namespace sodt {
thread_local long long g_launches = 0;
thread_local cudaError_t g_last_cuda_error = cudaSuccess;
}

int main(int argc, char** argv) {
    using namespace sodt;
    const int shift = argc > 1 ? atoi(argv[1]) : 0, C = argc > 2 ? atoi(argv[2]) : 192;
    const int B = 8, H = 256, W = 256, heads = 12;
    const size_t ntok = (size_t)B * H * W;
    std::vector<__nv_bfloat16> h(ntok * 3 * C);
    unsigned s = 12345u;
    for (auto& v : h) { s = s * 1664525u + 1013904223u; v = __float2bfloat16(((s >> 8) & 0xFFFF) / 32768.f - 1.f); }
    __nv_bfloat16 *qkv, *out;
    float *table, *ws;
    cudaMalloc(&qkv, h.size() * 2); cudaMalloc(&out, ntok * C * 2);
    cudaMalloc(&table, 225 * heads * 4); cudaMalloc(&ws, window_attn_win8_workspace(heads));
    cudaMemcpy(qkv, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(table, 0, 225 * heads * 4);
    for (int rep = 0; rep < 2; ++rep) {
        int st = window_attn_win8(qkv, table, out, ws, B, H, W, C, heads, shift, 0.25f, -100.f, 148, false, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (st != 0 || e != cudaSuccess) { printf("error %d %s\n", st, cudaGetErrorString(e)); return 1; }
    }
    {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        for (int rep = 0; rep < 5; ++rep) window_attn_win8(qkv, table, out, ws, B, H, W, C, heads, shift, 0.25f, -100.f, 148, true, 0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        printf("kernel time (B=%d, %dx%d tokens, C=%d): %.3f ms\n", B, H, W, C, ms / 5);
    }
    static long long t[4][128][12];
    cudaMemcpyFromSymbol(t, g_trace, sizeof(t));
    const long long t0 = t[2][0][0];
    printf("unit | softmax group (unit %% 2): wait_s s_ok ld_done math pv_ok epi_done p_arrive [epi: O_loaded at_bar past_bar] | MMA: qk_start stage_ok sfree_ok qk_done pv_wait p_ok pv_done\n");
    for (int n = 24; n < 60; ++n) {
        const int g = n & 1;
        printf("%3d |", n);
        for (int e = 0; e < 7; ++e) printf(" %7lld", t[g][n][e] ? t[g][n][e] - t0 : -1);
        printf(" [%lld %lld %lld] exp[%lld %lld] |", t[g][n][7] - t0, t[g][n][8] - t0, t[g][n][9] - t0, t[g][n][10] - t0, t[g][n][11] - t0);
        for (int e = 0; e < 7; ++e) printf(" %7lld", t[2][n][e] ? t[2][n][e] - t0 : -1);
        printf("\n");
    }
    printf("producer, per stage: start, slot free, loads issued\n");
    for (int n = 8; n < 30; ++n) printf("%3d | %7lld %7lld %7lld\n", n, t[3][n][0] - t0, t[3][n][1] - t0, t[3][n][2] - t0);
    return 0;
}
