// C-ABI entry points of the label-mixed Linear pair; dispatch between the SIMT fp32 kernels
// (gemm_simt.cu, any shape) and the tcgen05/TMEM 3xTF32 kernels (gemm_tc.cu, H % 8 == 0 shapes).
#include <stdlib.h>

#include "common.cuh"

namespace glass {
int dw_splits(int64_t n, int h, int k);
int pair_fwd_simt(const float* a1, int64_t lda1, int k1, const float* a2, int64_t lda2, int k2, const float* w0,
                  const float* b0, const float* w1, const float* b1, const uint8_t* mask, float z, int act, float* out,
                  int64_t ldo, float* acts, int64_t n, int h, cudaStream_t st);
int pair_bwd_simt(const float* dout, int64_t lddo, const float* acts, const float* a1, int64_t lda1, int k1,
                  const float* a2, int64_t lda2, int k2, const float* w0, const float* w1, const uint8_t* mask, float z,
                  int act, float* da1, int64_t ldda1, float* da2, int64_t ldda2, float* dw0, float* db0, float* dw1,
                  float* db1, int64_t n, int h, void* workspace, cudaStream_t st);
bool pair_tc_supported(int k1, int k2, int h, int64_t lda1, int64_t lda2, const void* a1, const void* a2);
bool pair_tc_norm_supported(int k1, int k2, int h);
int pair_fwd_tc(const float* a1, int64_t lda1, int k1, const float* a2, int64_t lda2, int k2, const float* w0,
                const float* b0, const float* w1, const float* b1, const uint8_t* mask, float z, int act, float* out,
                int64_t ldo, float* acts, int64_t n, int h, const glass_norm_operand* n1, const glass_norm_operand* n2,
                cudaStream_t st);
bool pair_dw_tc_supported(int k1, int k2, int h, int64_t lda1, int64_t lda2, const void* a1, const void* a2);
size_t pair_dw_tc_workspace_bytes(int64_t n, int h, int k);
int pair_bwd_dw_tc(const float* dout, int64_t lddo, const float* acts, const float* a1, int64_t lda1, int k1,
                   const float* a2, int64_t lda2, int k2, const uint8_t* mask, float z, int act, float* dw0, float* db0,
                   float* dw1, float* db1, int64_t n, int h, void* workspace, const glass_norm_operand* n1,
                   const glass_norm_operand* n2, cudaStream_t st);
int pair_bwd_dx_tc(const float* dout, int64_t lddo, const float* acts, const float* w0, const float* w1,
                   const uint8_t* mask, float z, int act, float* da1, int64_t ldda1, int k1, float* da2, int64_t ldda2,
                   int k2, int64_t n, int h, int acc1, int acc2, cudaStream_t st);
}  // namespace glass

using namespace glass;

static bool check_pair_common(const float* a1, int k1, const float* a2, int k2, const uint8_t* mask, int64_t n, int h,
                              int act) {
    if (!(a1 && k1 > 0 && k2 >= 0 && (k2 == 0 || a2) && mask && n >= 0 && h > 0)) {
        set_error("pair_linear_mix: bad arguments (k1=%d k2=%d n=%lld h=%d)", k1, k2, (long long)n, h);
        return false;
    }
    if (act != GLASS_ACT_NONE && act != GLASS_ACT_RELU && act != GLASS_ACT_ELU) {
        set_error("pair_linear_mix: unknown activation %d", act);
        return false;
    }
    return true;
}

static bool has_norm(const glass_norm_operand* n) { return n && n->stats; }

extern "C" int glass_pair_norm_operand_supported(int k1, int k2, int h) {
    return pair_tc_norm_supported(k1, k2, h) ? 1 : 0;
}

extern "C" int glass_pair_linear_mix_fwd(const float* a1, int64_t lda1, int k1, const float* a2, int64_t lda2, int k2,
                                         const float* w0, const float* b0, const float* w1, const float* b1,
                                         const uint8_t* mask, float z_ratio, int act, float* out, int64_t ldo,
                                         float* acts, int64_t n, int h, int path, void* stream) {
    return glass_pair_linear_mix_fwd_ex(a1, lda1, k1, a2, lda2, k2, w0, b0, w1, b1, mask, z_ratio, act, out, ldo, acts,
                                        n, h, path, nullptr, nullptr, stream);
}

extern "C" int glass_pair_linear_mix_fwd_ex(const float* a1, int64_t lda1, int k1, const float* a2, int64_t lda2, int k2,
                                            const float* w0, const float* b0, const float* w1, const float* b1,
                                            const uint8_t* mask, float z_ratio, int act, float* out, int64_t ldo,
                                            float* acts, int64_t n, int h, int path, const glass_norm_operand* n1,
                                            const glass_norm_operand* n2, void* stream) {
    if (!check_pair_common(a1, k1, a2, k2, mask, n, h, act)) return GLASS_ERR_BAD_ARG;
    GLASS_CHECK_ARG(w0 && b0 && w1 && b1 && out && ldo >= h && lda1 >= k1 && (k2 == 0 || lda2 >= k2),
                    "pair_linear_mix_fwd: bad arguments");
    if (n == 0) return GLASS_OK;
    cudaStream_t st = as_stream(stream);
    const bool tc_ok = pair_tc_supported(k1, k2, h, lda1, lda2, a1, a2);
    const bool norm = has_norm(n1) || has_norm(n2);
    if ((path == GLASS_GEMM_TCGEN05 && !tc_ok) || (norm && (path == GLASS_GEMM_SIMT || !tc_ok))) {
        set_error("pair_linear_mix_fwd: tcgen05 path%s does not support k1=%d k2=%d h=%d",
                  norm ? " (required by normalised operands)" : "", k1, k2, h);
        return GLASS_ERR_UNSUPPORTED;
    }
    if (path == GLASS_GEMM_TCGEN05 || (path == GLASS_GEMM_AUTO && tc_ok))
        return pair_fwd_tc(a1, lda1, k1, a2, lda2, k2, w0, b0, w1, b1, mask, z_ratio, act, out, ldo, acts, n, h, n1, n2,
                           st);
    return pair_fwd_simt(a1, lda1, k1, a2, lda2, k2, w0, b0, w1, b1, mask, z_ratio, act, out, ldo, acts, n, h, st);
}

extern "C" size_t glass_pair_linear_mix_bwd_workspace_bytes(int64_t n, int h, int k) {
    if (n < 0 || h <= 0 || k <= 0) return 0;
    size_t simt = (size_t)dw_splits(n, h, k) * 2 * (size_t)h * ((size_t)k + 1) * sizeof(float);
    size_t tc = pair_dw_tc_workspace_bytes(n, h, k);
    return align_up(simt > tc ? simt : tc, 256);
}

extern "C" int glass_pair_linear_mix_bwd(const float* dout, int64_t lddo, const float* acts, const float* a1,
                                         int64_t lda1, int k1, const float* a2, int64_t lda2, int k2, const float* w0,
                                         const float* w1, const uint8_t* mask, float z_ratio, int act, float* da1,
                                         int64_t ldda1, float* da2, int64_t ldda2, float* dw0, float* db0, float* dw1,
                                         float* db1, int64_t n, int h, void* workspace, size_t workspace_bytes,
                                         int path, void* stream) {
    return glass_pair_linear_mix_bwd_ex(dout, lddo, acts, a1, lda1, k1, a2, lda2, k2, w0, w1, mask, z_ratio, act, da1,
                                        ldda1, da2, ldda2, dw0, db0, dw1, db1, n, h, workspace, workspace_bytes, path,
                                        nullptr, nullptr, 0, 0, stream);
}

extern "C" int glass_pair_linear_mix_bwd_ex(const float* dout, int64_t lddo, const float* acts, const float* a1,
                                            int64_t lda1, int k1, const float* a2, int64_t lda2, int k2,
                                            const float* w0, const float* w1, const uint8_t* mask, float z_ratio,
                                            int act, float* da1, int64_t ldda1, float* da2, int64_t ldda2, float* dw0,
                                            float* db0, float* dw1, float* db1, int64_t n, int h, void* workspace,
                                            size_t workspace_bytes, int path, const glass_norm_operand* n1,
                                            const glass_norm_operand* n2, int accumulate_da1, int accumulate_da2,
                                            void* stream) {
    if (!check_pair_common(a1, k1, a2, k2, mask, n, h, act)) return GLASS_ERR_BAD_ARG;
    GLASS_CHECK_ARG(dout && w0 && w1 && dw0 && db0 && dw1 && db1 && lddo >= h, "pair_linear_mix_bwd: bad arguments");
    GLASS_CHECK_ARG(act == GLASS_ACT_NONE || acts, "pair_linear_mix_bwd: acts required when act != NONE");
    const size_t need = glass_pair_linear_mix_bwd_workspace_bytes(n, h, k1 + k2);
    if (workspace_bytes < need || !workspace) {
        set_error("pair_linear_mix_bwd: workspace %zu < required %zu", workspace_bytes, need);
        return GLASS_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    if (n == 0) {
        GLASS_CUDA(cudaMemsetAsync(dw0, 0, sizeof(float) * (size_t)h * (k1 + k2), st));
        GLASS_CUDA(cudaMemsetAsync(dw1, 0, sizeof(float) * (size_t)h * (k1 + k2), st));
        GLASS_CUDA(cudaMemsetAsync(db0, 0, sizeof(float) * (size_t)h, st));
        GLASS_CUDA(cudaMemsetAsync(db1, 0, sizeof(float) * (size_t)h, st));
        return GLASS_OK;
    }
    // dX goes through tcgen05 when the shape allows; dW/db (reduction over rows) stay on the SIMT split-N kernel.
    const bool tc_ok = pair_tc_supported(k1, k2, h, lda1, lda2, a1, a2);
    const bool special = has_norm(n1) || has_norm(n2) || accumulate_da1 || accumulate_da2;
    if ((path == GLASS_GEMM_TCGEN05 && !tc_ok) ||
        (special && (path == GLASS_GEMM_SIMT || !tc_ok || !pair_dw_tc_supported(k1, k2, h, lda1, lda2, a1, a2)))) {
        set_error("pair_linear_mix_bwd: tcgen05 path%s does not support k1=%d k2=%d h=%d",
                  special ? " (required by normalised operands / accumulation)" : "", k1, k2, h);
        return GLASS_ERR_UNSUPPORTED;
    }
    float* da1_s = da1;
    float* da2_s = da2;
    if ((da1 || da2) && (path == GLASS_GEMM_TCGEN05 || (path == GLASS_GEMM_AUTO && tc_ok))) {
        int rc = pair_bwd_dx_tc(dout, lddo, act == GLASS_ACT_NONE ? nullptr : acts, w0, w1, mask, z_ratio, act, da1,
                                ldda1, k1, da2, ldda2, k2, n, h, accumulate_da1, accumulate_da2, st);
        if (rc != GLASS_OK) return rc;
        da1_s = nullptr;
        da2_s = nullptr;
    }
    if ((path == GLASS_GEMM_TCGEN05 || path == GLASS_GEMM_AUTO) && tc_ok &&
        pair_dw_tc_supported(k1, k2, h, lda1, lda2, a1, a2)) {
        if (da1_s || da2_s) {   // (not reached: dX already went through tcgen05 above)
            int rc = pair_bwd_simt(dout, lddo, acts, a1, lda1, k1, a2, lda2, k2, w0, w1, mask, z_ratio, act, da1_s,
                                   ldda1, da2_s, ldda2, dw0, db0, dw1, db1, n, h, workspace, st);
            if (rc != GLASS_OK) return rc;
        }
        return pair_bwd_dw_tc(dout, lddo, act == GLASS_ACT_NONE ? nullptr : acts, a1, lda1, k1, a2, lda2, k2, mask,
                              z_ratio, act, dw0, db0, dw1, db1, n, h, workspace, n1, n2, st);
    }
    return pair_bwd_simt(dout, lddo, acts, a1, lda1, k1, a2, lda2, k2, w0, w1, mask, z_ratio, act, da1_s, ldda1, da2_s,
                         ldda2, dw0, db0, dw1, db1, n, h, workspace, st);
}
