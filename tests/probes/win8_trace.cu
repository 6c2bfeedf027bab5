// Timeline of the 8x8-window attention kernel's hand-offs (CTA 0): builds window_attn_win8.cu with SODT_WIN8_TRACE and
// prints, per unit, the clock64 deltas between the pipeline events of the softmax groups, the MMA thread and a producer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I csrc -I ../include \
//        -DSODT_WIN8_TRACE -o build/win8_trace ../tests/probes/win8_trace.cu
#include <cstdio>
#include <vector>
#include "../../small-object-detection-transformers_b200/csrc/window_attn_win8.cu"

namespace sodt {
thread_local long long g_launches = 0;
thread_local cudaError_t g_last_cuda_error = cudaSuccess;
}

int main(int argc, char** argv) {
    using namespace sodt;
    const int B = 8, H = 256, W = 256, C = 192, heads = 12, shift = argc > 1 ? atoi(argv[1]) : 0;
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
        int st = window_attn_win8(qkv, table, out, ws, B, H, W, C, heads, shift, 0.25f, -100.f, 148, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (st != 0 || e != cudaSuccess) { printf("error %d %s\n", st, cudaGetErrorString(e)); return 1; }
    }
    static long long t[4][96][12];
    cudaMemcpyFromSymbol(t, g_trace, sizeof(t));
    const long long t0 = t[2][0][0];
    printf("unit | softmax group (unit %% 2): wait_s s_ok S0 math0 pv_ok epi S1 math1 p_arrive | MMA: qk_start stage_ok sfree_ok qk_done pv_wait p_ok pv_done\n");
    for (int n = 6; n < 30; ++n) {
        const int g = n & 1;
        printf("%3d |", n);
        for (int e : {0, 1, 2, 3, 4, 5, 6, 7, 8}) printf(" %7lld", t[g][n][e] ? t[g][n][e] - t0 : -1);
        // inside the epilogue (between pv_ok and epi): O loaded from TMEM, arrival at the barrier of the two groups
        printf(" [epi: O loaded %lld, at barrier %lld]", t[g][n][9] - t0, t[g][n][10] - t0);
        printf(" |");
        for (int e = 0; e < 7; ++e) printf(" %7lld", t[2][n][e] ? t[2][n][e] - t0 : -1);
        printf("\n");
    }
    printf("producer q, per stage: start, slot free, loads stored\n");
    for (int n = 3; n < 15; ++n) printf("%3d | %7lld %7lld %7lld\n", n, t[3][n][0] - t0, t[3][n][1] - t0, t[3][n][2] - t0);
    return 0;
}
