// Host-side interface of mlp_tc.cu (the linear MLP of a Swin block as one tcgen05 kernel), used by capi.cu.
#pragma once
#include <cuda_runtime.h>

namespace sodt {

struct MlpTcArgs {
    const void* x;          // bf16 [M, C] residual stream (row stride ldx): LayerNorm input AND the residual
    int ldx;
    const float* ln_stats;  // [M][2] (mean, rstd) of the rows of x, or (ln_boxes > 0) [ln_boxes][M][2] partial (sum, sum of squares)
    int ln_boxes;
    float ln_eps;
    const float* ln_colsum; // fp32 [hidden]: row sums of the folded bf16 fc1 weights
    const void* w1;         // bf16 [hidden, C]: fc1.weight * diag(ln_weight)
    const float* b1;        // fp32 [hidden]: fc1.bias + fc1.weight . ln_bias
    const void* w2;         // [C, hidden]: bf16 fc2.weight, or (w2_fp16) IEEE fp16 0.5 * fc2.weight
    const float* b2;        // fp32 [C]
    void* out;              // bf16 [M, C] (row stride ldo)
    int ldo;
    float* stats_out;       // optional [C / 64][M][2] partial (sum, sum of squares) of the output rows
    int M, C, hidden;
    int w2_fp16;            // 1: the hidden activation is kept as fp16 2 GELU(.) and w2 is fp16, pre-scaled by 0.5
};

bool mlp_tc_supported(int M, int C, int hidden);
int mlp_tc(const MlpTcArgs& g, int num_sms, cudaStream_t stream);

}  // namespace sodt
