// Host-side interface of linear_tc.cu (tcgen05 GEMM with TMA-addressed A operand), used by capi.cu.
#pragma once
#include <cuda_runtime.h>

namespace sodt {

struct LinearTcArgs {
    const void* x;          // bf16 A operand (meaning depends on the entry point)
    int ldx;                // row stride of x in elements
    const void* w;          // bf16 [N, K] row-major
    const float* bias;      // fp32 [N] or null
    const void* residual;   // bf16 [M, N] (row stride ldr) or null
    int ldr;
    int res_rows;           // the residual has res_rows rows and repeats every res_rows rows of out (0: M rows)
    void* out;              // bf16 [M, N] (row stride ldo)
    int ldo;
    int M, N, K;
    int act;                // 0 identity, 1 GELU, 2 SiLU
    // LayerNorm folded into the GEMM (optional): A holds the raw rows, w the weights pre-multiplied by the LayerNorm gain,
    // bias = bias + w . ln_bias; the epilogue applies out = rstd (acc - mean colsum[n]) + bias[n] per row.
    const float* ln_stats;  // [M][2] (mean, rstd) of every A row, or (ln_boxes > 0) [ln_boxes][M][2] partial (sum, sum of squares); or null
    const float* ln_colsum; // fp32 [N]: sum over k of the (bf16) folded weights
    // optional: partial (sum, sum of squares) of the bf16 OUTPUT rows per 64-column box, [N/64][M][2]
    float* stats_out;
    int ln_boxes;           // 0, or 1..3: ln_stats holds that many partial pairs per row, reduced in the epilogue with ln_eps
    float ln_eps;
};

bool linear_tc_supported(int M, int N, int K);
int linear_tc(const LinearTcArgs& g, const void* x2, int ldx2, int k_split, int num_sms, cudaStream_t stream);

bool conv_tc_supported(int B, int H, int W, int Cin, int Cout, int kh, int kw);
int conv_tc(const void* x, int ldx, const void* w, const float* bias, void* out, int ldo, int B, int H, int W, int Cin, int Cout,
            int kh, int kw, int pad_t, int pad_l, int act, int num_sms, cudaStream_t stream);

bool upcat_tc_supported(int B, int H, int W, int C1, int C2, int Cout);
int upcat_tc(const void* low, const void* skip, const void* w, const float* bias, void* out, int ldo, int B, int H, int W, int C1, int C2,
             int Cout, int act, int num_sms, cudaStream_t stream);

bool merge_tc_supported(int B, int H, int W, int C, int N);
int merge_tc(const void* x, const void* w, const float* bias, void* out, int B, int H, int W, int C, int N, int num_sms,
             cudaStream_t stream);

}  // namespace sodt
