// R4 (exact-fp32 path): nonlocal spatial self-attention core on the CUDA cores.
// Replaces NonlocalResNetBlock.call lines layers/attention.py:318-389:
//     O = softmax_rows(scale * Q K^T) V ;  y = gamma * O + x
// This is the full-fp32 implementation (bit-for-bit the arithmetic the
// reference does in TF 2.3, which has no TF32).  It is the accuracy anchor of
// the tensor-core kernel in nonlocal_tc.cu and the backward used until that
// kernel's gradient lands; selected through davi_nonlocal_attn_* (see the
// dispatch at the bottom).  Never materialises [B,N,N] in HBM: a CTA owns a
// 16-row block of one side and keeps its [16,N] strip of logits in shared
// memory.
//
// All three kernels are built from two strip routines:
//   strip_dot  : out[r][y] = <M[r,:], Y[y,:]>      (M: 16 x w in smem, Y: N x w in HBM)
//   strip_outer: out[r][c] = sum_y S[r][y] Y[y][c]  (S: 16 x N in smem, Y: N x w in HBM)
#include "davi_common.cuh"

namespace davi {

constexpr int kNlBM = 16;        // rows of the block side
constexpr int kNlThreads = 256;
constexpr int kNlPad = 4;

// out[r][y] (stride NP) = <sM[r][0:4*w4], Y[y][0:4*w4]>
__device__ __forceinline__ void strip_dot(const float* sM, int w4, const float* __restrict__ Y,
                                          int64_t y_s, int N, int NP, float* out) {
    for (int y = threadIdx.x; y < N; y += kNlThreads) {
        const float* yp = Y + (int64_t)y * y_s;
        float acc[kNlBM];
#pragma unroll
        for (int r = 0; r < kNlBM; ++r) acc[r] = 0.f;
        for (int c = 0; c < w4; ++c) {
            const float4 v = ld4(yp + 4 * c);
#pragma unroll
            for (int r = 0; r < kNlBM; ++r)
                acc[r] += dot4(v, *reinterpret_cast<const float4*>(sM + (r * w4 + c) * 4));
        }
#pragma unroll
        for (int r = 0; r < kNlBM; ++r) out[r * NP + y] = acc[r];
    }
}

// For items (half, c): rows half*8..half*8+7, float4 column c < w4:
//   acc[rr] = sum_y S[row][y] * Y[y][c];   then emit(row, c, acc)
template <typename Emit>
__device__ __forceinline__ void strip_outer(const float* S, int NP, int N, int w4,
                                            const float* __restrict__ Y, int64_t y_s, Emit emit) {
    for (int item = threadIdx.x; item < 2 * w4; item += kNlThreads) {
        const int half = item / w4, c = item - half * w4;
        float4 acc[8];
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) acc[rr] = make_float4(0, 0, 0, 0);
        const float* Sp = S + (half * 8) * NP;
        for (int y = 0; y < N; y += 4) {
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = ld4(Y + (int64_t)(y + k) * y_s + 4 * c);
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const float4 s = *reinterpret_cast<const float4*>(Sp + rr * NP + y);
                fma4(acc[rr], s.x, v[0]);
                fma4(acc[rr], s.y, v[1]);
                fma4(acc[rr], s.z, v[2]);
                fma4(acc[rr], s.w, v[3]);
            }
        }
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) emit(half * 8 + rr, c, acc[rr]);
    }
}

__device__ __forceinline__ void load_block(float* dst, const float* __restrict__ src, int64_t s,
                                           int w4) {
    for (int i = threadIdx.x; i < kNlBM * w4; i += kNlThreads) {
        const int r = i / w4, c = i - r * w4;
        *reinterpret_cast<float4*>(dst + (r * w4 + c) * 4) = ld4(src + (int64_t)r * s + 4 * c);
    }
}

__global__ void __launch_bounds__(kNlThreads)
nonlocal_fwd_simt(const float* __restrict__ Q, const float* __restrict__ K,
                  const float* __restrict__ V, const float* __restrict__ x, float* __restrict__ y,
                  float* __restrict__ lse, int N, int d4, int C4, int64_t q_s, int64_t k_s,
                  int64_t v_s, int64_t x_s, int64_t y_s, const float* scale_p, const float* gamma_p) {
    const float scale = scale_p ? __ldg(scale_p) : 1.f, gamma = __ldg(gamma_p);
    extern __shared__ __align__(16) float smem[];
    const int NP = N + kNlPad;
    float* sA = smem;                 // [16][NP]
    float* sX = sA + kNlBM * NP;      // [16][d]
    const int b = blockIdx.y, i0 = blockIdx.x * kNlBM;
    const float* Qb = Q + ((int64_t)b * N + i0) * q_s;
    const float* Kb = K + (int64_t)b * N * k_s;
    const float* Vb = V + (int64_t)b * N * v_s;

    load_block(sX, Qb, q_s, d4);
    __syncthreads();
    strip_dot(sX, d4, Kb, k_s, N, NP, sA);   // att.py:105
    __syncthreads();
    // row softmax (att.py:107-116); 8 warps x 2 rows
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < kNlBM; r += kNlThreads / 32) {
        float* row = sA + r * NP;
        float m = -INFINITY;
        for (int j = lane; j < N; j += 32) { row[j] *= scale; m = fmaxf(m, row[j]); }
        m = group_max<32>(m);
        float s = 0.f;
        for (int j = lane; j < N; j += 32) { const float e = __expf(row[j] - m); row[j] = e; s += e; }
        s = group_sum<32>(s);
        const float inv = 1.f / s;
        for (int j = lane; j < N; j += 32) row[j] *= inv;
        if (lane == 0) lse[(int64_t)b * N + i0 + r] = m + logf(s);
    }
    __syncthreads();
    const float* xb = x + ((int64_t)b * N + i0) * x_s;
    float* yb = y + ((int64_t)b * N + i0) * y_s;
    strip_outer(sA, NP, N, C4, Vb, v_s, [&](int r, int c, const float4& o) {   // att.py:371
        const float4 xi = ld4(xb + (int64_t)r * x_s + 4 * c);
        float4 out;
        out.x = fmaf(gamma, o.x, xi.x); out.y = fmaf(gamma, o.y, xi.y);           // att.py:389
        out.z = fmaf(gamma, o.z, xi.z); out.w = fmaf(gamma, o.w, xi.w);
        *reinterpret_cast<float4*>(yb + (int64_t)r * y_s + 4 * c) = out;
    });
}

// Backward, query-block side: dQ, E_i = sum_j P_ij <dy_i, V_j>, dgamma, dscale.
__global__ void __launch_bounds__(kNlThreads)
nonlocal_bwd_q_simt(const float* __restrict__ Q, const float* __restrict__ K,
                    const float* __restrict__ V, const float* __restrict__ lse,
                    const float* __restrict__ dy, float* __restrict__ dQ, float* __restrict__ E,
                    float* __restrict__ dgs, int N, int d4, int C4, int64_t q_s, int64_t k_s,
                    int64_t v_s, int64_t dy_s, const float* scale_p, const float* gamma_p) {
    const float scale = scale_p ? __ldg(scale_p) : 1.f, gamma = __ldg(gamma_p);
    extern __shared__ __align__(16) float smem[];
    __shared__ float red[32];
    __shared__ float sL[kNlBM];
    const int NP = N + kNlPad;
    float* sA = smem;                   // raw logits q.k
    float* sB = sA + kNlBM * NP;        // <dy_i, V_j>, then dS
    float* sX = sB + kNlBM * NP;        // Q block
    float* sM = sX + kNlBM * d4 * 4;    // dy block
    const int b = blockIdx.y, i0 = blockIdx.x * kNlBM;
    const float* Kb = K + (int64_t)b * N * k_s;
    const float* Vb = V + (int64_t)b * N * v_s;

    load_block(sX, Q + ((int64_t)b * N + i0) * q_s, q_s, d4);
    load_block(sM, dy + ((int64_t)b * N + i0) * dy_s, dy_s, C4);
    if (threadIdx.x < kNlBM) sL[threadIdx.x] = lse[(int64_t)b * N + i0 + threadIdx.x];
    __syncthreads();
    strip_dot(sX, d4, Kb, k_s, N, NP, sA);
    strip_dot(sM, C4, Vb, v_s, N, NP, sB);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float dsc = 0.f, dgm = 0.f;
    for (int r = warp; r < kNlBM; r += kNlThreads / 32) {
        const float* ra = sA + r * NP;
        float* rb = sB + r * NP;
        const float l = sL[r];
        float e = 0.f;
        for (int j = lane; j < N; j += 32) e += __expf(scale * ra[j] - l) * rb[j];
        e = group_sum<32>(e);
        if (lane == 0) { dgm += e; E[(int64_t)b * N + i0 + r] = e; }
        for (int j = lane; j < N; j += 32) {
            const float ds = gamma * __expf(scale * ra[j] - l) * (rb[j] - e);
            rb[j] = ds;
            dsc += ds * ra[j];
        }
    }
    __syncthreads();
    float* dQb = dQ + ((int64_t)b * N + i0) * q_s;
    strip_outer(sB, NP, N, d4, Kb, k_s, [&](int r, int c, const float4& o) {
        *reinterpret_cast<float4*>(dQb + (int64_t)r * q_s + 4 * c) =
            make_float4(scale * o.x, scale * o.y, scale * o.z, scale * o.w);
    });
    dgm = block_sum(dgm, red);
    if (threadIdx.x == 0) atomicAdd(dgs + 0, dgm);
    dsc = block_sum(dsc, red);
    if (threadIdx.x == 0) atomicAdd(dgs + 1, dsc);
}

// Backward, key-block side: dK, dV.
__global__ void __launch_bounds__(kNlThreads)
nonlocal_bwd_kv_simt(const float* __restrict__ Q, const float* __restrict__ K,
                     const float* __restrict__ V, const float* __restrict__ lse,
                     const float* __restrict__ E, const float* __restrict__ dy,
                     float* __restrict__ dK, float* __restrict__ dV, int N, int d4, int C4,
                     int64_t q_s, int64_t k_s, int64_t v_s, int64_t dy_s, const float* scale_p, const float* gamma_p) {
    const float scale = scale_p ? __ldg(scale_p) : 1.f, gamma = __ldg(gamma_p);
    extern __shared__ __align__(16) float smem[];
    const int NP = N + kNlPad;
    float* sA = smem;                   // raw logits k_j.q_i, then P
    float* sB = sA + kNlBM * NP;        // <dy_i, V_j>, then dS
    float* sX = sB + kNlBM * NP;        // K block
    float* sM = sX + kNlBM * d4 * 4;    // V block
    const int b = blockIdx.y, j0 = blockIdx.x * kNlBM;
    const float* Qb = Q + (int64_t)b * N * q_s;
    const float* dyb = dy + (int64_t)b * N * dy_s;
    const float* lb = lse + (int64_t)b * N;
    const float* Eb = E + (int64_t)b * N;

    load_block(sX, K + ((int64_t)b * N + j0) * k_s, k_s, d4);
    load_block(sM, V + ((int64_t)b * N + j0) * v_s, v_s, C4);
    __syncthreads();
    strip_dot(sX, d4, Qb, q_s, N, NP, sA);
    strip_dot(sM, C4, dyb, dy_s, N, NP, sB);
    __syncthreads();
    for (int idx = threadIdx.x; idx < kNlBM * N; idx += kNlThreads) {
        const int r = idx / N, i = idx - r * N;
        const float p = __expf(scale * sA[r * NP + i] - lb[i]);
        sA[r * NP + i] = p;
        sB[r * NP + i] = gamma * p * (sB[r * NP + i] - Eb[i]);
    }
    __syncthreads();
    float* dVb = dV + ((int64_t)b * N + j0) * v_s;
    strip_outer(sA, NP, N, C4, dyb, dy_s, [&](int r, int c, const float4& o) {
        *reinterpret_cast<float4*>(dVb + (int64_t)r * v_s + 4 * c) =
            make_float4(gamma * o.x, gamma * o.y, gamma * o.z, gamma * o.w);
    });
    float* dKb = dK + ((int64_t)b * N + j0) * k_s;
    strip_outer(sB, NP, N, d4, Qb, q_s, [&](int r, int c, const float4& o) {
        *reinterpret_cast<float4*>(dKb + (int64_t)r * k_s + 4 * c) =
            make_float4(scale * o.x, scale * o.y, scale * o.z, scale * o.w);
    });
}

static int nl_validate(const char* fn, int B, int N, int d, int C) {
    if (B <= 0 || N <= 0 || (N % kNlBM) || d <= 0 || (d & 3) || d > 64 || C <= 0 || (C & 3) || C > 512) {
        set_error("%s: need N %% 16 == 0, d %% 4 == 0 (<=64), C %% 4 == 0 (<=512); got B=%d N=%d d=%d C=%d",
                  fn, B, N, d, C);
        return DAVI_EINVAL;
    }
    return DAVI_OK;
}

int nonlocal_fwd_simt_launch(const float* Q, const float* K, const float* V, const float* x, float* y,
                             float* lse, int B, int N, int d, int C, int64_t q_s, int64_t k_s,
                             int64_t v_s, int64_t x_s, int64_t y_s, const float* scale, const float* gamma,
                             cudaStream_t st) {
    int rc = nl_validate("davi_nonlocal_attn_fwd", B, N, d, C);
    if (rc) return rc;
    const size_t smem = ((size_t)kNlBM * (N + kNlPad) + (size_t)kNlBM * d) * sizeof(float);
    DAVI_CUDA(cudaFuncSetAttribute(nonlocal_fwd_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nonlocal_fwd_simt<<<dim3(N / kNlBM, B), kNlThreads, smem, st>>>(Q, K, V, x, y, lse, N, d / 4, C / 4, q_s,
                                                                    k_s, v_s, x_s, y_s, scale, gamma);
    return check_launch("davi_nonlocal_attn_fwd");
}

int nonlocal_bwd_simt_launch(const float* Q, const float* K, const float* V, const float* lse,
                             const float* dy, float* dQ, float* dK, float* dV, float* dgs, int B, int N,
                             int d, int C, int64_t q_s, int64_t k_s, int64_t v_s, int64_t dy_s, const float* scale,
                             const float* gamma, float* E, cudaStream_t st) {
    int rc = nl_validate("davi_nonlocal_attn_bwd", B, N, d, C);
    if (rc) return rc;
    const size_t smem = ((size_t)2 * kNlBM * (N + kNlPad) + (size_t)kNlBM * d + (size_t)kNlBM * C) * sizeof(float);
    if (smem > 227 * 1024) {
        set_error("davi_nonlocal_attn_bwd: N=%d C=%d needs %zu B of shared memory", N, C, smem);
        return DAVI_EINVAL;
    }
    DAVI_CUDA(cudaFuncSetAttribute(nonlocal_bwd_q_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DAVI_CUDA(cudaFuncSetAttribute(nonlocal_bwd_kv_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DAVI_CUDA(cudaMemsetAsync(dgs, 0, 2 * sizeof(float), st));
    nonlocal_bwd_q_simt<<<dim3(N / kNlBM, B), kNlThreads, smem, st>>>(Q, K, V, lse, dy, dQ, E, dgs, N, d / 4,
                                                                      C / 4, q_s, k_s, v_s, dy_s, scale, gamma);
    rc = check_launch("davi_nonlocal_attn_bwd");
    if (rc) return rc;
    nonlocal_bwd_kv_simt<<<dim3(N / kNlBM, B), kNlThreads, smem, st>>>(Q, K, V, lse, E, dy, dK, dV, N, d / 4,
                                                                       C / 4, q_s, k_s, v_s, dy_s, scale, gamma);
    return check_launch("davi_nonlocal_attn_bwd");
}

}  // namespace davi
