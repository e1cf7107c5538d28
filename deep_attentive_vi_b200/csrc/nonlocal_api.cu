// C-ABI entry points of the nonlocal attention core (include/davi.h, R4) and
// the dispatch between the tensor-core kernel (nonlocal_tc.cu) and the
// exact-fp32 CUDA-core kernel (nonlocal_simt.cu).
#include <stdlib.h>
#include <string.h>

#include "davi_common.cuh"

namespace davi {
int nonlocal_fwd_simt_launch(const float*, const float*, const float*, const float*, float*, float*,
                             int, int, int, int, int64_t, int64_t, int64_t, int64_t, int64_t, const float*,
                             const float*, cudaStream_t);
int nonlocal_bwd_simt_launch(const float*, const float*, const float*, const float*, const float*,
                             float*, float*, float*, float*, int, int, int, int, int64_t, int64_t,
                             int64_t, int64_t, const float*, const float*, float*, cudaStream_t);
#ifdef DAVI_HAVE_NONLOCAL_TC
bool nonlocal_tc_supported(int B, int N, int d, int C);
int nonlocal_fwd_tc_launch(const float*, const float*, const float*, const float*, float*, float*, float*,
                           int, int, int, int, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, const float*,
                           const float*, cudaStream_t);
int nonlocal_bwd_tc_launch(const float*, const float*, const float*, const float*, const float*, const float*,
                           float*, float*, float*, float*, int, int, int, int, int64_t, int64_t, int64_t, int64_t,
                           int64_t, const float*, const float*, float*, cudaStream_t);
#endif

// davi_set_option(DAVI_OPT_NONLOCAL_IMPL, 1) forces the exact-fp32 CUDA-core kernels
static bool force_simt() { return get_option(DAVI_OPT_NONLOCAL_IMPL) == 1; }
}  // namespace davi

using namespace davi;

extern "C" int davi_nonlocal_attn_fwd(const float* Q, const float* K, const float* V, const float* x,
                                      float* y, float* lse, float* o_save, int B, int N, int d, int C,
                                      int64_t q_s, int64_t k_s, int64_t v_s, int64_t x_s, int64_t y_s,
                                      int64_t o_s, const float* scale, const float* gamma,
                                      davi_stream_t stream) {
    DAVI_REQUIRE(Q && K && V && x && y && lse && gamma, "null pointer");
    DAVI_REQUIRE(aligned16(Q) && aligned16(K) && aligned16(V) && aligned16(x) && aligned16(y), "alignment");
    DAVI_REQUIRE(mult4(q_s) && mult4(k_s) && mult4(v_s) && mult4(x_s) && mult4(y_s), "strides % 4");
    DAVI_REQUIRE(q_s >= d && k_s >= d && v_s >= C && x_s >= C && y_s >= C, "strides >= width");
    DAVI_REQUIRE(!o_save || (aligned16(o_save) && mult4(o_s) && o_s >= C), "o_save alignment / stride");
#ifdef DAVI_HAVE_NONLOCAL_TC
    if (!force_simt() && nonlocal_tc_supported(B, N, d, C))
        return nonlocal_fwd_tc_launch(Q, K, V, x, y, lse, o_save, B, N, d, C, q_s, k_s, v_s, x_s, y_s, o_s,
                                      scale, gamma, (cudaStream_t)stream);
#endif
    (void)force_simt;
    // the CUDA-core gradient recomputes what it needs: o_save is not written on this path
    return nonlocal_fwd_simt_launch(Q, K, V, x, y, lse, B, N, d, C, q_s, k_s, v_s, x_s, y_s, scale, gamma,
                                    (cudaStream_t)stream);
}

extern "C" size_t davi_nonlocal_attn_bwd_workspace_bytes(int B, int N, int d, int C) {
    (void)d; (void)C;
    return (size_t)B * (size_t)N * sizeof(float);
}

extern "C" int davi_nonlocal_attn_bwd(const float* Q, const float* K, const float* V, const float* lse,
                                      const float* O, const float* dy, float* dQ, float* dK, float* dV,
                                      float* dgamma_dscale, int B, int N, int d, int C, int64_t q_s,
                                      int64_t k_s, int64_t v_s, int64_t o_s, int64_t dy_s, const float* scale,
                                      const float* gamma, void* workspace, size_t workspace_bytes,
                                      davi_stream_t stream) {
    DAVI_REQUIRE(Q && K && V && lse && dy && dQ && dK && dV && dgamma_dscale && gamma, "null pointer");
    DAVI_REQUIRE(aligned16(Q) && aligned16(K) && aligned16(V) && aligned16(dy) && aligned16(dQ) &&
                     aligned16(dK) && aligned16(dV), "alignment");
    DAVI_REQUIRE(mult4(q_s) && mult4(k_s) && mult4(v_s) && mult4(dy_s), "strides % 4");
    DAVI_REQUIRE(q_s >= d && k_s >= d && v_s >= C && dy_s >= C, "strides >= width");
    if (!workspace || workspace_bytes < davi_nonlocal_attn_bwd_workspace_bytes(B, N, d, C)) {
        set_error("%s: workspace too small", __func__);
        return DAVI_EWORKSPACE;
    }
#ifdef DAVI_HAVE_NONLOCAL_TC
    if (O && !force_simt() && nonlocal_tc_supported(B, N, d, C)) {
        DAVI_REQUIRE(aligned16(O) && mult4(o_s) && o_s >= C, "O alignment / stride");
        return nonlocal_bwd_tc_launch(Q, K, V, lse, O, dy, dQ, dK, dV, dgamma_dscale, B, N, d, C, q_s, k_s, v_s,
                                      o_s, dy_s, scale, gamma, (float*)workspace, (cudaStream_t)stream);
    }
#endif
    return nonlocal_bwd_simt_launch(Q, K, V, lse, dy, dQ, dK, dV, dgamma_dscale, B, N, d, C, q_s, k_s, v_s,
                                    dy_s, scale, gamma, (float*)workspace, (cudaStream_t)stream);
}
