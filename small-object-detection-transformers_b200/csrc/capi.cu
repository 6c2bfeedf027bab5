// C-ABI entry points that only validate and dispatch (see include/sodt_b200.h), plus library bookkeeping.
#include "common.cuh"
#include "linear_tc.h"
#include "mlp_tc.h"

namespace sodt {

thread_local long long g_launches = 0;
thread_local cudaError_t g_last_cuda_error = cudaSuccess;

int window_attn_generic(const void* qkv, const float* table, const void* pad_qkv, void* out, int B, int H, int W,
                        int C, int heads, int ws, int shift, int dtype, float scale, float mask_value,
                        cudaStream_t stream);

int window_attn_generic_ex(const void* qkv, const float* table, const void* pad_qkv, void* out, int B, int H, int W,
                           int C, int heads, int ws, int shift, int dtype, float scale, float mask_value,
                           const float* dense_mask, int mask_windows, const float* head_scale, int normalize_qk,
                           const float* rel_pos_h, const float* rel_pos_w, cudaStream_t stream);

size_t window_attn_flash_workspace(int heads, int ws);
bool window_attn_flash_supported(int H, int W, int C, int heads, int ws, int shift, int dtype);
int window_attn_flash_prepare(const float* table, void* workspace, int heads, int ws, cudaStream_t stream);
int window_attn_flash(const void* qkv, const float* table, void* out, void* workspace, int B, int H, int W, int C,
                      int heads, int ws, float scale, bool prepared, cudaStream_t stream);

size_t window_attn_win8_workspace(int heads);
bool window_attn_win8_supported(int H, int W, int C, int heads, int ws, int shift, int dtype);
int window_attn_win8_prepare(const float* table, void* workspace, int heads, cudaStream_t stream);
int window_attn_win8(const void* qkv, const float* table, void* out, void* workspace, int B, int H, int W, int C,
                     int heads, int shift, float scale, float mask_value, int num_sms, bool prepared, cudaStream_t stream);

bool attn_block_supported(int H, int W, int C, int heads, int ws, int shift);
int attn_block(const void* x, const float* ln_stats, int ln_boxes, float ln_eps, const void* w, const float* colsum, const float* bias,
               const void* table_ws, void* out, int B, int H, int W, int C, int heads, int shift, float scale, float mask_value,
               int num_sms, cudaStream_t stream);

static int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}

}  // namespace sodt

extern "C" size_t sodt_window_attn_workspace_bytes(int C, int heads, int ws) {
    if (C <= 0 || heads <= 0 || ws <= 0) return 0;
    const size_t a = sodt::window_attn_flash_workspace(heads, ws), b = sodt::window_attn_win8_workspace(heads);
    return a > b ? a : b;
}

extern "C" int sodt_linear_supported(int M, int N, int K, int dtype) {
    return dtype == SODT_BF16 && sodt::linear_tc_supported(M, N, K) ? 1 : 0;
}

extern "C" int sodt_linear_fwd(const void* x, const void* w, const float* bias, const void* residual, void* out,
                               int M, int N, int K, int act, int dtype, void* stream) {
    return sodt_linear_strided_fwd(x, K, nullptr, 0, 0, w, bias, residual, N, 0, out, N, M, N, K, act, dtype, stream);
}

extern "C" int sodt_linear_strided_fwd(const void* x, int ldx, const void* x2, int ldx2, int k_split, const void* w,
                                       const float* bias, const void* residual, int ldr, int res_rows, void* out, int ldo,
                                       int M, int N, int K, int act, int dtype, void* stream) {
    using namespace sodt;
    if (!x || !w || !out || M <= 0 || N <= 0 || K <= 0 || act < 0 || act > 2) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_BF16 || !linear_tc_supported(M, N, K)) return SODT_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(w) || !aligned16(out) || (bias && !aligned16(bias)) || (residual && !aligned16(residual)) ||
        (x2 && !aligned16(x2)))
        return SODT_ERR_ALIGNMENT;
    if (res_rows < 0) return SODT_ERR_INVALID_ARG;
    LinearTcArgs g{x, ldx, w, bias, residual, ldr, res_rows, out, ldo, M, N, K, act};
    return linear_tc(g, x2, ldx2, k_split, sm_count(), static_cast<cudaStream_t>(stream));
}

extern "C" int sodt_linear_ln_fwd(const void* x, int ldx, const float* ln_mean_rstd, int ln_boxes, float ln_eps, const float* ln_colsum,
                                  const void* w, const float* bias, const void* residual, int ldr, int res_rows, void* out, int ldo,
                                  float* stats_out, int M, int N, int K, int act, int dtype, void* stream) {
    using namespace sodt;
    if (!x || !w || !out || M <= 0 || N <= 0 || K <= 0 || act < 0 || act > 2) return SODT_ERR_INVALID_ARG;
    if ((ln_mean_rstd && !ln_colsum) || res_rows < 0 || ln_boxes < 0 || ln_boxes > 6 || ln_eps < 0.f) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_BF16 || !linear_tc_supported(M, N, K)) return SODT_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(w) || !aligned16(out) || (bias && !aligned16(bias)) || (residual && !aligned16(residual)) ||
        (ln_colsum && !aligned16(ln_colsum)) || (reinterpret_cast<uintptr_t>(ln_mean_rstd) & 7) || (reinterpret_cast<uintptr_t>(stats_out) & 7))
        return SODT_ERR_ALIGNMENT;
    LinearTcArgs g{x, ldx, w, bias, residual, ldr, res_rows, out, ldo, M, N, K, act, ln_mean_rstd, ln_colsum, stats_out, ln_boxes, ln_eps};
    return linear_tc(g, nullptr, 0, 0, sm_count(), static_cast<cudaStream_t>(stream));
}

extern "C" int sodt_attn_block_supported(int B, int H, int W, int C, int heads, int ws, int shift, int dtype) {
    return B > 0 && dtype == SODT_BF16 && sodt::attn_block_supported(H, W, C, heads, ws, shift) ? 1 : 0;
}

extern "C" int sodt_attn_block_fwd(const void* x, const float* ln_mean_rstd, int ln_boxes, float ln_eps, const float* ln_colsum,
                                   const void* w_qkv, const float* b_qkv, void* out, int B, int H, int W, int C, int heads, int ws, int shift,
                                   int dtype, float scale, float mask_value, const void* workspace, size_t workspace_bytes, void* stream) {
    using namespace sodt;
    if (!x || !ln_mean_rstd || !ln_colsum || !w_qkv || !b_qkv || !out || !workspace || B <= 0 || H <= 0 || W <= 0 || C <= 0 || heads <= 0)
        return SODT_ERR_INVALID_ARG;
    if (ln_boxes < 0 || ln_boxes > 6 || ln_eps < 0.f) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_BF16 || !attn_block_supported(H, W, C, heads, ws, shift)) return SODT_ERR_UNSUPPORTED;
    if (workspace_bytes < window_attn_win8_workspace(heads)) return SODT_ERR_WORKSPACE;
    if (!aligned16(x) || !aligned16(w_qkv) || !aligned16(out) || !aligned16(b_qkv) || !aligned16(ln_colsum) || !aligned16(workspace) ||
        (reinterpret_cast<uintptr_t>(ln_mean_rstd) & 7))
        return SODT_ERR_ALIGNMENT;
    return attn_block(x, ln_mean_rstd, ln_boxes, ln_eps, w_qkv, ln_colsum, b_qkv, workspace, out, B, H, W, C, heads, shift, scale, mask_value,
                      sm_count(), static_cast<cudaStream_t>(stream));
}

extern "C" int sodt_mlp_supported(int M, int C, int hidden, int dtype) {
    return dtype == SODT_BF16 && sodt::mlp_tc_supported(M, C, hidden) ? 1 : 0;
}

extern "C" int sodt_mlp_ln_fwd(const void* x, int ldx, const float* ln_mean_rstd, int ln_boxes, float ln_eps, const float* ln_colsum,
                               const void* w1, const float* b1, const void* w2, const float* b2, void* out, int ldo,
                               float* stats_out, int M, int C, int hidden, int w2_fp16, int dtype, void* stream) {
    using namespace sodt;
    if (w2_fp16 < 0 || w2_fp16 > 1) return SODT_ERR_INVALID_ARG;
    if (!x || !ln_mean_rstd || !ln_colsum || !w1 || !b1 || !w2 || !b2 || !out || M <= 0 || C <= 0 || hidden <= 0) return SODT_ERR_INVALID_ARG;
    if (ln_boxes < 0 || ln_boxes > 3 || ln_eps < 0.f) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_BF16 || !mlp_tc_supported(M, C, hidden)) return SODT_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(w1) || !aligned16(w2) || !aligned16(out) || !aligned16(b1) || !aligned16(b2) || !aligned16(ln_colsum) ||
        (reinterpret_cast<uintptr_t>(ln_mean_rstd) & 7) || (reinterpret_cast<uintptr_t>(stats_out) & 7))
        return SODT_ERR_ALIGNMENT;
    MlpTcArgs g{x, ldx, ln_mean_rstd, ln_boxes, ln_eps, ln_colsum, w1, b1, w2, b2, out, ldo, stats_out, M, C, hidden, w2_fp16};
    return mlp_tc(g, sm_count(), static_cast<cudaStream_t>(stream));
}

extern "C" int sodt_conv2d_nhwc_supported(int B, int H, int W, int Cin, int Cout, int kh, int kw, int dtype) {
    return dtype == SODT_BF16 && sodt::conv_tc_supported(B, H, W, Cin, Cout, kh, kw) ? 1 : 0;
}

extern "C" int sodt_conv2d_nhwc_fwd(const void* x, int ldx, const void* w, const float* bias, void* out, int ldo,
                                    int B, int H, int W, int Cin, int Cout, int kh, int kw, int pad_t, int pad_l,
                                    int act, int dtype, void* stream) {
    using namespace sodt;
    if (!x || !w || !out || B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || act < 0 || act > 2) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_BF16 || !conv_tc_supported(B, H, W, Cin, Cout, kh, kw)) return SODT_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(w) || !aligned16(out) || (bias && !aligned16(bias))) return SODT_ERR_ALIGNMENT;
    return conv_tc(x, ldx, w, bias, out, ldo, B, H, W, Cin, Cout, kh, kw, pad_t, pad_l, act, sm_count(), static_cast<cudaStream_t>(stream));
}

extern "C" int sodt_upcat_conv1x1_supported(int B, int H, int W, int C1, int C2, int Cout, int dtype) {
    return dtype == SODT_BF16 && sodt::upcat_tc_supported(B, H, W, C1, C2, Cout) ? 1 : 0;
}

extern "C" int sodt_upcat_conv1x1_fwd(const void* low, const void* skip, const void* w, const float* bias, void* out, int ldo,
                                      int B, int H, int W, int C1, int C2, int Cout, int act, int dtype, void* stream) {
    using namespace sodt;
    if (!low || !skip || !w || !out || B <= 0 || H <= 0 || W <= 0 || act < 0 || act > 2) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_BF16 || !upcat_tc_supported(B, H, W, C1, C2, Cout)) return SODT_ERR_UNSUPPORTED;
    if (!aligned16(low) || !aligned16(skip) || !aligned16(w) || !aligned16(out) || (bias && !aligned16(bias))) return SODT_ERR_ALIGNMENT;
    return upcat_tc(low, skip, w, bias, out, ldo, B, H, W, C1, C2, Cout, act, sm_count(), static_cast<cudaStream_t>(stream));
}

extern "C" int sodt_patch_merge_linear_supported(int B, int H, int W, int C, int N, int dtype) {
    return dtype == SODT_BF16 && sodt::merge_tc_supported(B, H, W, C, N) ? 1 : 0;
}

extern "C" int sodt_patch_merge_linear_fwd(const void* x, const void* w, const float* bias, void* out, int B, int H, int W,
                                           int C, int N, int dtype, void* stream) {
    using namespace sodt;
    if (!x || !w || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0 || N <= 0) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_BF16 || !merge_tc_supported(B, H, W, C, N)) return SODT_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(w) || !aligned16(out) || (bias && !aligned16(bias))) return SODT_ERR_ALIGNMENT;
    return merge_tc(x, w, bias, out, B, H, W, C, N, sm_count(), static_cast<cudaStream_t>(stream));
}

extern "C" int sodt_version(void) { return 100; }
extern "C" int sodt_built_for_sm(void) { return 100; }

extern "C" const char* sodt_status_string(int status) {
    switch (status) {
        case SODT_OK: return "ok";
        case SODT_ERR_INVALID_ARG: return "invalid argument";
        case SODT_ERR_UNSUPPORTED: return "unsupported shape";
        case SODT_ERR_WORKSPACE: return "workspace missing, misaligned or too small";
        case SODT_ERR_CUDA: return "CUDA error";
        case SODT_ERR_ALIGNMENT: return "pointer not 16-byte aligned";
        default: return "unknown status";
    }
}

extern "C" const char* sodt_last_cuda_error(void) { return cudaGetErrorString(sodt::g_last_cuda_error); }
extern "C" long long sodt_launch_count(void) { return sodt::g_launches; }
extern "C" void sodt_reset_launch_count(void) { sodt::g_launches = 0; }

static int window_attn_dispatch(const void* qkv, const float* bias_table, const void* pad_qkv, void* out,
                                int B, int H, int W, int C, int heads, int ws, int shift,
                                int dtype, float scale, float mask_value,
                                void* workspace, size_t workspace_bytes, bool prepared, void* stream) {
    using namespace sodt;
    if (!qkv || !bias_table || !out) return SODT_ERR_INVALID_ARG;
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || heads <= 0 || ws <= 0 || C % heads != 0) return SODT_ERR_INVALID_ARG;
    if (shift < 0 || shift >= ws) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_F32 && dtype != SODT_BF16) return SODT_ERR_INVALID_ARG;
    if (!aligned16(qkv) || !aligned16(out) || (pad_qkv && !aligned16(pad_qkv))) return SODT_ERR_ALIGNMENT;
    if (window_attn_flash_supported(H, W, C, heads, ws, shift, dtype) && (long long)B * (H / ws) * (W / ws) <= 65535) {
        if (!workspace || !aligned16(workspace) || workspace_bytes < window_attn_flash_workspace(heads, ws))
            return SODT_ERR_WORKSPACE;
        return window_attn_flash(qkv, bias_table, out, workspace, B, H, W, C, heads, ws, scale, prepared,
                                 static_cast<cudaStream_t>(stream));
    }
    if (window_attn_win8_supported(H, W, C, heads, ws, shift, dtype)) {
        if (!workspace || !aligned16(workspace) || workspace_bytes < window_attn_win8_workspace(heads))
            return SODT_ERR_WORKSPACE;
        return window_attn_win8(qkv, bias_table, out, workspace, B, H, W, C, heads, shift, scale, mask_value, sm_count(), prepared,
                                static_cast<cudaStream_t>(stream));
    }
    return window_attn_generic(qkv, bias_table, pad_qkv, out, B, H, W, C, heads, ws, shift, dtype, scale, mask_value,
                               static_cast<cudaStream_t>(stream));
}

extern "C" int sodt_window_attn_fwd(const void* qkv, const float* bias_table, const void* pad_qkv, void* out,
                                    int B, int H, int W, int C, int heads, int ws, int shift,
                                    int dtype, float scale, float mask_value,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    return window_attn_dispatch(qkv, bias_table, pad_qkv, out, B, H, W, C, heads, ws, shift, dtype, scale, mask_value, workspace,
                                workspace_bytes, false, stream);
}

extern "C" int sodt_window_attn_ex_fwd(const void* qkv, const float* bias_table, const void* pad_qkv, void* out,
                                       int B, int H, int W, int C, int heads, int ws, int shift,
                                       int dtype, float scale, float mask_value,
                                       const float* dense_mask, int mask_windows, const float* head_scale, int normalize_qk,
                                       const float* rel_pos_h, const float* rel_pos_w, void* stream) {
    using namespace sodt;
    if (!qkv || !bias_table || !out) return SODT_ERR_INVALID_ARG;
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || heads <= 0 || ws <= 0 || C % heads != 0) return SODT_ERR_INVALID_ARG;
    if (shift < 0 || shift >= ws || (dense_mask && mask_windows <= 0)) return SODT_ERR_INVALID_ARG;
    if (dtype != SODT_F32 && dtype != SODT_BF16) return SODT_ERR_INVALID_ARG;
    if (!aligned16(qkv) || !aligned16(out) || (pad_qkv && !aligned16(pad_qkv))) return SODT_ERR_ALIGNMENT;
    return window_attn_generic_ex(qkv, bias_table, pad_qkv, out, B, H, W, C, heads, ws, shift, dtype, scale, mask_value, dense_mask,
                                  mask_windows, head_scale, normalize_qk, rel_pos_h, rel_pos_w, static_cast<cudaStream_t>(stream));
}

extern "C" int sodt_window_attn_kernel_class(int B, int H, int W, int C, int heads, int ws, int shift, int dtype) {
    using namespace sodt;
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || heads <= 0 || ws <= 0 || C % heads != 0 || shift < 0 || shift >= ws) return SODT_ERR_INVALID_ARG;
    if (window_attn_flash_supported(H, W, C, heads, ws, shift, dtype) && (long long)B * (H / ws) * (W / ws) <= 65535) return 1;
    if (window_attn_win8_supported(H, W, C, heads, ws, shift, dtype)) return 2;
    return 0;
}

extern "C" int sodt_window_attn_prepare(const float* bias_table, int B, int H, int W, int C, int heads, int ws, int shift, int dtype,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    using namespace sodt;
    if (!bias_table || B <= 0 || H <= 0 || W <= 0 || C <= 0 || heads <= 0 || ws <= 0 || C % heads != 0 || shift < 0 || shift >= ws)
        return SODT_ERR_INVALID_ARG;
    if (window_attn_flash_supported(H, W, C, heads, ws, shift, dtype) && (long long)B * (H / ws) * (W / ws) <= 65535) {
        if (!workspace || !aligned16(workspace) || workspace_bytes < window_attn_flash_workspace(heads, ws)) return SODT_ERR_WORKSPACE;
        return window_attn_flash_prepare(bias_table, workspace, heads, ws, static_cast<cudaStream_t>(stream));
    }
    if (window_attn_win8_supported(H, W, C, heads, ws, shift, dtype)) {
        if (!workspace || !aligned16(workspace) || workspace_bytes < window_attn_win8_workspace(heads)) return SODT_ERR_WORKSPACE;
        return window_attn_win8_prepare(bias_table, workspace, heads, static_cast<cudaStream_t>(stream));
    }
    return SODT_OK;          // the exact kernel reads the parameter itself
}

extern "C" int sodt_window_attn_fwd_prepared(const void* qkv, const float* bias_table, const void* pad_qkv, void* out,
                                             int B, int H, int W, int C, int heads, int ws, int shift,
                                             int dtype, float scale, float mask_value,
                                             const void* prepared_workspace, size_t workspace_bytes, void* stream) {
    return window_attn_dispatch(qkv, bias_table, pad_qkv, out, B, H, W, C, heads, ws, shift, dtype, scale, mask_value,
                                const_cast<void*>(prepared_workspace), workspace_bytes, true, stream);
}
