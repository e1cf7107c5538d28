// R1: depth-wise (cross-layer, per-pixel) attention, forward + hand-written
// backward.  Replaces DepthWiseAttention.apply (layers/attention.py:135-164) +
// SoftmaxAlignment.call (layers/attention.py:98-116) of the reference.
//
// The op is HBM-bound (0.47 flop/B): each pixel reads n key rows (d floats) and
// n value rows (C floats) exactly once.  Layout of the work:
//   * G lanes (8 or 16) own one pixel; a warp owns 32/G adjacent pixels;
//   * scores: lane l computes the full <q,K_j> of slots j = l, l+G, ... (rows of
//     d floats, 16-byte loads), the softmax is a shuffle reduction over the G
//     lanes -- nothing is staged or re-read;
//   * weighted sum: lanes stride over the float4 columns of the value rows,
//     U slots x ITERS columns of independent 16-byte streaming loads are issued
//     before the FMAs so every lane keeps U*ITERS*16 B in flight;
//   * slots are addressed through a pointer table (stacked tensors are just the
//     special case ptr[j] = base + j*stride), so the tf.stack / Permute copies
//     of the reference never exist.
// Algorithmic bytes: fwd 4*B*P*(n*(d+C)+d+C), bwd
// 4*B*P*(2n(d+C)+2d+C); the kernels touch each of those bytes once.
#include "davi_common.cuh"

namespace davi {

struct SlotPtrs {
    const float* k[DAVI_MAX_SLOTS];
    const float* v[DAVI_MAX_SLOTS];
};
struct SlotGradPtrs {
    float* dk[DAVI_MAX_SLOTS];
    float* dv[DAVI_MAX_SLOTS];
};

struct DwGeom {
    int n, P, total_pix, d4, C4;
    int64_t q_sp, k_sb, k_sp, v_sb, v_sp, o_sp;
    const float* scale_ptr;  // device scalar, nullptr = 1.0
};

constexpr int kDwThreads = 128;
constexpr int kDwUnroll = 4;  // slots whose value rows are in flight together

// softmax over the n slots of one pixel; lane l of the group holds slots l+t*G.
// returns a[t] (weights) and, if R != nullptr, the raw dot products r[t].
template <int G, int SPL>
__device__ __forceinline__ void dw_scores(const float* const* s_k, const DwGeom& g, int l,
                                          int64_t koff, const float* qp, float (&a)[SPL],
                                          float* r_raw, unsigned gm, float scale) {
    float s[SPL];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < SPL; ++t) {
        const int j = l + t * G;
        s[t] = -INFINITY;
        if (j < g.n) {
            const float* kp = s_k[j] + koff;
            float dot = 0.f;
            for (int c = 0; c < g.d4; ++c) dot += dot4(ld4(kp + 4 * c), ld4(qp + 4 * c));
            if (r_raw) r_raw[t] = dot;
            s[t] = dot * scale;
        } else if (r_raw) {
            r_raw[t] = 0.f;
        }
        m = fmaxf(m, s[t]);
    }
    m = group_max<G>(m, gm);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < SPL; ++t) {
        a[t] = (l + t * G < g.n) ? __expf(s[t] - m) : 0.f;
        sum += a[t];
    }
    sum = group_sum<G>(sum, gm);
    const float inv = 1.f / sum;
#pragma unroll
    for (int t = 0; t < SPL; ++t) a[t] *= inv;
}

template <int G, int ITERS>
__global__ void __launch_bounds__(kDwThreads)
dw_attn_fwd_kernel(const SlotPtrs sp, const float* __restrict__ q, float* __restrict__ out,
                   float* __restrict__ attn, const DwGeom g) {
    constexpr int SPL = DAVI_MAX_SLOTS / G;
    __shared__ const float* s_k[DAVI_MAX_SLOTS];
    __shared__ const float* s_v[DAVI_MAX_SLOTS];
    if (threadIdx.x < DAVI_MAX_SLOTS) {
        s_k[threadIdx.x] = sp.k[threadIdx.x];
        s_v[threadIdx.x] = sp.v[threadIdx.x];
    }
    __syncthreads();

    const int l = threadIdx.x & (G - 1);
    const int gbase = (threadIdx.x & 31) & ~(G - 1);  // first lane of my group in the warp
    const int pix = (blockIdx.x * kDwThreads + threadIdx.x) / G;
    if (pix >= g.total_pix) return;  // whole groups exit together (G divides 32)
    const int b = pix / g.P, p = pix - b * g.P;
    const int64_t koff = (int64_t)b * g.k_sb + (int64_t)p * g.k_sp;
    const int64_t voff = (int64_t)b * g.v_sb + (int64_t)p * g.v_sp;
    const float* qp = q + (int64_t)pix * g.q_sp;

    float a[SPL];
    const unsigned gm = group_mask<G>();
    const float scale = g.scale_ptr ? __ldg(g.scale_ptr) : 1.f;
    dw_scores<G, SPL>(s_k, g, l, koff, qp, a, nullptr, gm, scale);
    if (attn) {
#pragma unroll
        for (int t = 0; t < SPL; ++t)
            if (l + t * G < g.n) attn[(int64_t)pix * g.n + l + t * G] = a[t];
    }

    float4 acc[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool last_ok = (l + (ITERS - 1) * G) < g.C4;  // only the last column block is ragged

#pragma unroll
    for (int t = 0; t < SPL; ++t) {
        const int jend = min(g.n, (t + 1) * G);
        int j = t * G;
        for (; j + kDwUnroll <= jend; j += kDwUnroll) {
            float4 v[kDwUnroll][ITERS];
#pragma unroll
            for (int u = 0; u < kDwUnroll; ++u) {
                const float* vp = s_v[j + u] + voff + 4 * l;
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
                    if (it < ITERS - 1 || last_ok) v[u][it] = ld_stream4(vp + 4 * it * G);
                    else v[u][it] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < kDwUnroll; ++u) {
                const float aj = __shfl_sync(gm, a[t], gbase + ((j + u) & (G - 1)));
#pragma unroll
                for (int it = 0; it < ITERS; ++it) fma4(acc[it], aj, v[u][it]);
            }
        }
        for (; j < jend; ++j) {
            const float* vp = s_v[j] + voff + 4 * l;
            float4 v[ITERS];
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                if (it < ITERS - 1 || last_ok) v[it] = ld_stream4(vp + 4 * it * G);
                else v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float aj = __shfl_sync(gm, a[t], gbase + (j & (G - 1)));
#pragma unroll
            for (int it = 0; it < ITERS; ++it) fma4(acc[it], aj, v[it]);
        }
    }

    float* op = out + (int64_t)pix * g.o_sp + 4 * l;
#pragma unroll
    for (int it = 0; it < ITERS; ++it)
        if (it < ITERS - 1 || last_ok) st_stream4(op + 4 * it * G, acc[it]);
}

// Backward.  Per pixel: recompute a; pass 1 streams V once for da_j = <V_j, g>;
// ds_j = a_j (da_j - sum_k a_k da_k); pass 2 only writes: dV_j = a_j g,
// dK_j = scale ds_j q, dq = scale sum_j ds_j K_j (K rows re-read from L1/L2).
template <int G, int ITERS, bool ACC>
__global__ void __launch_bounds__(kDwThreads)
dw_attn_bwd_kernel(const SlotPtrs sp, const SlotGradPtrs gp, const float* __restrict__ q,
                   const float* __restrict__ dout, float* __restrict__ dq,
                   float* __restrict__ dscale, const DwGeom g) {
    constexpr int SPL = DAVI_MAX_SLOTS / G;
    __shared__ const float* s_k[DAVI_MAX_SLOTS];
    __shared__ const float* s_v[DAVI_MAX_SLOTS];
    __shared__ float* s_dk[DAVI_MAX_SLOTS];
    __shared__ float* s_dv[DAVI_MAX_SLOTS];
    __shared__ float red[32];
    if (threadIdx.x < DAVI_MAX_SLOTS) {
        s_k[threadIdx.x] = sp.k[threadIdx.x];
        s_v[threadIdx.x] = sp.v[threadIdx.x];
        s_dk[threadIdx.x] = gp.dk[threadIdx.x];
        s_dv[threadIdx.x] = gp.dv[threadIdx.x];
    }
    __syncthreads();

    const int l = threadIdx.x & (G - 1);
    const int gbase = (threadIdx.x & 31) & ~(G - 1);
    const int pix = (blockIdx.x * kDwThreads + threadIdx.x) / G;
    const bool active = pix < g.total_pix;
    float dsc_local = 0.f;

    if (active) {
        const int b = pix / g.P, p = pix - b * g.P;
        const int64_t koff = (int64_t)b * g.k_sb + (int64_t)p * g.k_sp;
        const int64_t voff = (int64_t)b * g.v_sb + (int64_t)p * g.v_sp;
        const float* qp = q + (int64_t)pix * g.q_sp;

        float a[SPL], r[SPL], da[SPL];
        const unsigned gm = group_mask<G>();
        const float scale = g.scale_ptr ? __ldg(g.scale_ptr) : 1.f;
        dw_scores<G, SPL>(s_k, g, l, koff, qp, a, r, gm, scale);

        const bool last_ok = (l + (ITERS - 1) * G) < g.C4;
        float4 go[ITERS];
        {
            const float* gpx = dout + (int64_t)pix * g.o_sp + 4 * l;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                if (it < ITERS - 1 || last_ok) go[it] = ld_stream4(gpx + 4 * it * G);
                else go[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }

        // pass 1: da_j
#pragma unroll
        for (int t = 0; t < SPL; ++t) {
            da[t] = 0.f;
            const int jend = min(g.n, (t + 1) * G);
            int j = t * G;
            for (; j + kDwUnroll <= jend; j += kDwUnroll) {
                float4 v[kDwUnroll][ITERS];
#pragma unroll
                for (int u = 0; u < kDwUnroll; ++u) {
                    const float* vp = s_v[j + u] + voff + 4 * l;
#pragma unroll
                    for (int it = 0; it < ITERS; ++it) {
                        if (it < ITERS - 1 || last_ok) v[u][it] = ld_stream4(vp + 4 * it * G);
                        else v[u][it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int u = 0; u < kDwUnroll; ++u) {
                    float part = 0.f;
#pragma unroll
                    for (int it = 0; it < ITERS; ++it) part += dot4(v[u][it], go[it]);
                    part = group_sum<G>(part, gm);
                    if (((j + u) & (G - 1)) == l) da[t] = part;
                }
            }
            for (; j < jend; ++j) {
                const float* vp = s_v[j] + voff + 4 * l;
                float part = 0.f;
#pragma unroll
                for (int it = 0; it < ITERS; ++it)
                    if (it < ITERS - 1 || last_ok) part += dot4(ld_stream4(vp + 4 * it * G), go[it]);
                part = group_sum<G>(part, gm);
                if ((j & (G - 1)) == l) da[t] = part;
            }
        }

        float dsum = 0.f;
#pragma unroll
        for (int t = 0; t < SPL; ++t) dsum += a[t] * da[t];
        dsum = group_sum<G>(dsum, gm);
        float ds[SPL];
#pragma unroll
        for (int t = 0; t < SPL; ++t) {
            ds[t] = a[t] * (da[t] - dsum);
            dsc_local += ds[t] * r[t];
        }

        // pass 2: writes.  lanes < d4 own one float4 column of q / dq / dK rows.
        const bool kcol = l < g.d4;
        float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), dqv = qv;
        if (kcol) qv = ld4(qp + 4 * l);
#pragma unroll
        for (int t = 0; t < SPL; ++t) {
            const int jend = min(g.n, (t + 1) * G);
            for (int j = t * G; j < jend; ++j) {
                const int src = gbase + (j & (G - 1));
                const float aj = __shfl_sync(gm, a[t], src);
                const float dsj = __shfl_sync(gm, ds[t], src) * scale;
                float* dvp = s_dv[j];
                if (dvp) {
                    dvp += voff + 4 * l;
#pragma unroll
                    for (int it = 0; it < ITERS; ++it) {
                        if (it < ITERS - 1 || last_ok) {
                            float4 o;
                            if (ACC) {
                                o = *reinterpret_cast<const float4*>(dvp + 4 * it * G);
                                fma4(o, aj, go[it]);
                            } else {
                                o = make_float4(aj * go[it].x, aj * go[it].y, aj * go[it].z,
                                                aj * go[it].w);
                            }
                            st_stream4(dvp + 4 * it * G, o);
                        }
                    }
                }
                if (kcol) {
                    const float4 kv = ld4(s_k[j] + koff + 4 * l);
                    fma4(dqv, dsj, kv);
                    float* dkp = s_dk[j];
                    if (dkp) {
                        dkp += koff + 4 * l;
                        float4 o;
                        if (ACC) {
                            o = *reinterpret_cast<const float4*>(dkp);
                            fma4(o, dsj, qv);
                        } else {
                            o = make_float4(dsj * qv.x, dsj * qv.y, dsj * qv.z, dsj * qv.w);
                        }
                        *reinterpret_cast<float4*>(dkp) = o;
                    }
                }
            }
        }
        if (kcol && dq) *reinterpret_cast<float4*>(dq + (int64_t)pix * g.q_sp + 4 * l) = dqv;
    }

    if (dscale) {  // uniform branch: every thread of the block reaches the barrier
        const float tot = block_sum(dsc_local, red);
        if (threadIdx.x == 0) atomicAdd(dscale, tot);
    }
}

// ---- host-side dispatch ----------------------------------------------------

static int dw_validate(const char* fn, int B, int n, int P, int d, int C, const DwGeom& g) {
    if (B <= 0 || P <= 0 || n < 1 || n > DAVI_MAX_SLOTS) {
        set_error("%s: need B,P > 0 and 1 <= n <= %d (got B=%d n=%d P=%d)", fn, DAVI_MAX_SLOTS, B, n, P);
        return DAVI_EINVAL;
    }
    if (d <= 0 || C <= 0 || (d & 3) || (C & 3) || d > 128 || C > 512) {
        set_error("%s: need d %% 4 == 0 (<=128) and C %% 4 == 0 (<=512) (got d=%d C=%d)", fn, d, C);
        return DAVI_EINVAL;
    }
    if (!mult4(g.q_sp) || !mult4(g.k_sb) || !mult4(g.k_sp) || !mult4(g.v_sb) || !mult4(g.v_sp) ||
        !mult4(g.o_sp) || g.q_sp < d || g.k_sp < d || g.v_sp < C || g.o_sp < C) {
        set_error("%s: strides must be multiples of 4 floats and >= the row width", fn);
        return DAVI_EINVAL;
    }
    if ((int64_t)B * P > 0x7fffffff / 32) {
        set_error("%s: B*P too large", fn);
        return DAVI_EINVAL;
    }
    return DAVI_OK;
}

template <int G, int ITERS>
static void launch_fwd(const SlotPtrs& sp, const float* q, float* out, float* attn, const DwGeom& g,
                       cudaStream_t st) {
    const int64_t threads = (int64_t)g.total_pix * G;
    const int grid = (int)((threads + kDwThreads - 1) / kDwThreads);
    dw_attn_fwd_kernel<G, ITERS><<<grid, kDwThreads, 0, st>>>(sp, q, out, attn, g);
}
template <int G, int ITERS>
static void launch_bwd(const SlotPtrs& sp, const SlotGradPtrs& gp, const float* q, const float* dout,
                       float* dq, float* dscale, const DwGeom& g, bool acc, cudaStream_t st) {
    const int64_t threads = (int64_t)g.total_pix * G;
    const int grid = (int)((threads + kDwThreads - 1) / kDwThreads);
    if (acc)
        dw_attn_bwd_kernel<G, ITERS, true><<<grid, kDwThreads, 0, st>>>(sp, gp, q, dout, dq, dscale, g);
    else
        dw_attn_bwd_kernel<G, ITERS, false><<<grid, kDwThreads, 0, st>>>(sp, gp, q, dout, dq, dscale, g);
}

#define DW_DISPATCH(LAUNCH, ...)                                   \
    do {                                                           \
        if (g.C4 <= 32) {                                          \
            switch ((g.C4 + 7) / 8) {                              \
                case 1: LAUNCH<8, 1>(__VA_ARGS__); break;          \
                case 2: LAUNCH<8, 2>(__VA_ARGS__); break;          \
                case 3: LAUNCH<8, 3>(__VA_ARGS__); break;          \
                default: LAUNCH<8, 4>(__VA_ARGS__); break;         \
            }                                                      \
        } else {                                                   \
            switch ((g.C4 + 15) / 16) {                            \
                case 3: LAUNCH<16, 3>(__VA_ARGS__); break;         \
                case 4: LAUNCH<16, 4>(__VA_ARGS__); break;         \
                case 5: LAUNCH<16, 5>(__VA_ARGS__); break;         \
                case 6: LAUNCH<16, 6>(__VA_ARGS__); break;         \
                case 7: LAUNCH<16, 7>(__VA_ARGS__); break;         \
                default: LAUNCH<16, 8>(__VA_ARGS__); break;        \
            }                                                      \
        }                                                          \
    } while (0)

static int dw_fwd_impl(const char* fn, const SlotPtrs& sp, const float* q, float* out, float* attn,
                       int B, int n, int P, int d, int C, DwGeom g, cudaStream_t st) {
    g.n = n; g.P = P; g.total_pix = B * P; g.d4 = d / 4; g.C4 = C / 4;
    int rc = dw_validate(fn, B, n, P, d, C, g);
    if (rc) return rc;
    if (!q || !out || !aligned16(q) || !aligned16(out) || (attn && !aligned16(attn))) {
        set_error("%s: q/out must be non-null and 16-byte aligned", fn);
        return DAVI_EINVAL;
    }
    for (int j = 0; j < n; ++j)
        if (!sp.k[j] || !sp.v[j] || !aligned16(sp.k[j]) || !aligned16(sp.v[j])) {
            set_error("%s: slot %d key/value pointer null or not 16-byte aligned", fn, j);
            return DAVI_EINVAL;
        }
    DW_DISPATCH(launch_fwd, sp, q, out, attn, g, st);
    return check_launch(fn);
}

static int dw_bwd_impl(const char* fn, const SlotPtrs& sp, const SlotGradPtrs& gp, const float* q,
                       const float* dout, float* dq, float* dscale, int B, int n, int P, int d, int C,
                       DwGeom g, bool acc, cudaStream_t st) {
    g.n = n; g.P = P; g.total_pix = B * P; g.d4 = d / 4; g.C4 = C / 4;
    int rc = dw_validate(fn, B, n, P, d, C, g);
    if (rc) return rc;
    if (!q || !dout || !aligned16(q) || !aligned16(dout) || (dq && !aligned16(dq))) {
        set_error("%s: q/dout must be non-null and 16-byte aligned", fn);
        return DAVI_EINVAL;
    }
    for (int j = 0; j < n; ++j) {
        if (!sp.k[j] || !sp.v[j] || !aligned16(sp.k[j]) || !aligned16(sp.v[j]) ||
            !aligned16(gp.dk[j]) || !aligned16(gp.dv[j])) {
            set_error("%s: slot %d pointer null or not 16-byte aligned", fn, j);
            return DAVI_EINVAL;
        }
    }
    if (dscale) DAVI_CUDA(cudaMemsetAsync(dscale, 0, sizeof(float), st));
    DW_DISPATCH(launch_bwd, sp, gp, q, dout, dq, dscale, g, acc, st);
    return check_launch(fn);
}

}  // namespace davi

using namespace davi;

extern "C" int davi_dw_attn_fwd(const float* q, const float* keys, const float* values, float* out,
                                float* attn, int B, int n, int P, int d, int C, int64_t q_sp,
                                int64_t k_sb, int64_t k_sn, int64_t k_sp, int64_t v_sb, int64_t v_sn,
                                int64_t v_sp, int64_t o_sp, const float* scale, davi_stream_t stream) {
    DAVI_REQUIRE(keys && values, "null keys/values");
    DAVI_REQUIRE(n >= 1 && n <= DAVI_MAX_SLOTS, "n out of range");
    DAVI_REQUIRE(mult4(k_sn) && mult4(v_sn), "layer strides must be multiples of 4");
    SlotPtrs sp{};
    for (int j = 0; j < n; ++j) { sp.k[j] = keys + j * k_sn; sp.v[j] = values + j * v_sn; }
    DwGeom g{}; g.q_sp = q_sp; g.k_sb = k_sb; g.k_sp = k_sp; g.v_sb = v_sb; g.v_sp = v_sp; g.o_sp = o_sp;
    g.scale_ptr = scale;
    return dw_fwd_impl(__func__, sp, q, out, attn, B, n, P, d, C, g, (cudaStream_t)stream);
}

extern "C" int davi_dw_attn_bwd(const float* q, const float* keys, const float* values,
                                const float* dout, float* dq, float* dkeys, float* dvalues,
                                float* dscale, int B, int n, int P, int d, int C, int64_t q_sp,
                                int64_t k_sb, int64_t k_sn, int64_t k_sp, int64_t v_sb, int64_t v_sn,
                                int64_t v_sp, int64_t o_sp, const float* scale, davi_stream_t stream) {
    DAVI_REQUIRE(keys && values, "null keys/values");
    DAVI_REQUIRE(n >= 1 && n <= DAVI_MAX_SLOTS, "n out of range");
    DAVI_REQUIRE(mult4(k_sn) && mult4(v_sn), "layer strides must be multiples of 4");
    SlotPtrs sp{}; SlotGradPtrs gp{};
    for (int j = 0; j < n; ++j) {
        sp.k[j] = keys + j * k_sn; sp.v[j] = values + j * v_sn;
        gp.dk[j] = dkeys ? dkeys + j * k_sn : nullptr;
        gp.dv[j] = dvalues ? dvalues + j * v_sn : nullptr;
    }
    DwGeom g{}; g.q_sp = q_sp; g.k_sb = k_sb; g.k_sp = k_sp; g.v_sb = v_sb; g.v_sp = v_sp; g.o_sp = o_sp;
    g.scale_ptr = scale;
    return dw_bwd_impl(__func__, sp, gp, q, dout, dq, dscale, B, n, P, d, C, g, false,
                       (cudaStream_t)stream);
}

extern "C" int davi_dw_attn_slots_fwd(const float* q, const float* const* k_ptrs,
                                      const float* const* v_ptrs, float* out, float* attn, int B,
                                      int n, int P, int d, int C, int64_t q_sp, int64_t k_sb,
                                      int64_t k_sp, int64_t v_sb, int64_t v_sp, int64_t o_sp,
                                      const float* scale, davi_stream_t stream) {
    DAVI_REQUIRE(k_ptrs && v_ptrs, "null slot tables");
    DAVI_REQUIRE(n >= 1 && n <= DAVI_MAX_SLOTS, "n out of range");
    SlotPtrs sp{};
    for (int j = 0; j < n; ++j) { sp.k[j] = k_ptrs[j]; sp.v[j] = v_ptrs[j]; }
    DwGeom g{}; g.q_sp = q_sp; g.k_sb = k_sb; g.k_sp = k_sp; g.v_sb = v_sb; g.v_sp = v_sp; g.o_sp = o_sp;
    g.scale_ptr = scale;
    return dw_fwd_impl(__func__, sp, q, out, attn, B, n, P, d, C, g, (cudaStream_t)stream);
}

extern "C" int davi_dw_attn_slots_bwd(const float* q, const float* const* k_ptrs,
                                      const float* const* v_ptrs, const float* dout, float* dq,
                                      float* const* dk_ptrs, float* const* dv_ptrs, float* dscale,
                                      int B, int n, int P, int d, int C, int64_t q_sp, int64_t k_sb,
                                      int64_t k_sp, int64_t v_sb, int64_t v_sp, int64_t o_sp,
                                      const float* scale, int accumulate, davi_stream_t stream) {
    DAVI_REQUIRE(k_ptrs && v_ptrs && dk_ptrs && dv_ptrs, "null slot tables");
    DAVI_REQUIRE(n >= 1 && n <= DAVI_MAX_SLOTS, "n out of range");
    SlotPtrs sp{}; SlotGradPtrs gp{};
    for (int j = 0; j < n; ++j) {
        sp.k[j] = k_ptrs[j]; sp.v[j] = v_ptrs[j];
        gp.dk[j] = dk_ptrs[j]; gp.dv[j] = dv_ptrs[j];
    }
    DwGeom g{}; g.q_sp = q_sp; g.k_sb = k_sb; g.k_sp = k_sp; g.v_sb = v_sb; g.v_sp = v_sp; g.o_sp = o_sp;
    g.scale_ptr = scale;
    return dw_bwd_impl(__func__, sp, gp, q, dout, dq, dscale, B, n, P, d, C, g, accumulate != 0,
                       (cudaStream_t)stream);
}
