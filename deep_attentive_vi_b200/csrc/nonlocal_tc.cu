// R4 (tensor-core path): nonlocal spatial self-attention core on the 5th-gen
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
// Replaces NonlocalResNetBlock.call lines layers/attention.py:318-389:
//     O = softmax_rows(scale * Q K^T) V ;  y = gamma * O + x
//
// The reference computes in full fp32 (TF 2.3 has no TF32), so every product is
// error-compensated: each fp32 operand is split into tf32 (hi, lo) halves and
// hi*hi' + lo*hi' + hi*lo' is accumulated in fp32 ("3xTF32", ~2^-21 relative per
// product).  The [B,N,N] score matrix never exists in HBM or shared memory.
//
// One CTA owns 128 query rows of one image.  Warp roles (288 threads):
//   warps 0-3  softmax: thread r owns query row r = TMEM lane r; row statistics,
//              P = softmax tile -> (hi, lo) -> shared memory; final epilogue
//   warps 4-7  loaders: global fp32 -> (hi, lo) -> UMMA canonical smem layouts
//   warp  8    one elected thread issues every tcgen05.mma / tcgen05.commit
// Two passes over the keys in blocks of 128 (S block = 128 TMEM columns):
//   pass A: S = Q K_blk^T, running row max / sum (softmax statistics only)
//   pass B: S recomputed (K-dim d <= 32: ~10 % of the PV flops), P normalised with
//           the final statistics, O += P V in 16-key chunks through a ring of
//           shared-memory stages; O lives in TMEM columns [128, 128 + Cp).
// The recompute keeps TMEM within 512 columns (128 + 320) for every N, so N = 256
// and N = 1024 run the same code and O never needs rescaling.
#include "davi_common.cuh"
#include "tc_common.cuh"

namespace davi {

using namespace tc;

constexpr int kTcM = 128;          // query rows per CTA
constexpr int kTcKB = 128;         // keys per S block
constexpr int kTcKC = 16;          // keys per PV chunk
constexpr int kTcChunks = kTcKB / kTcKC;
constexpr int kTcThreads = 288;
constexpr int kTcMaxStages = 4;
constexpr int kTcTmemCols = 512;
constexpr int kTcOCol = kTcKB;     // first TMEM column of O

struct TcGeom {
    int N, d, C, Cp;               // Cp = C rounded up to 16 (MMA N granularity)
    int n1, n2;                    // PV instruction widths (n1 + n2 = Cp, n2 may be 0)
    int stages;
    int64_t q_s, k_s, v_s, x_s, y_s;
};

struct TcBars {
    uint64_t k_full, k_empty, s_full, s_empty, o_full;
    uint64_t pv_full[kTcMaxStages], pv_empty[kTcMaxStages];
    uint32_t tmem_base;
};

// rows [row0, row0+128) x d of a K-major operand: global fp32 -> hi/lo tiles
__device__ __forceinline__ void load_kmajor_tile(const float* __restrict__ src, int64_t stride, int d4,
                                                 float* hi, float* lo, int lt) {
    // lane -> (row%8, chunk%4): 8 rows x 64 contiguous bytes per warp instruction,
    // and each quarter-warp stores 128 contiguous bytes (no bank conflicts)
    const int lane = lt & 31, w = lt >> 5;
    const int r8 = lane & 7, c4 = lane >> 3;
    const int citems = (d4 + 3) >> 2;
    const int items = (kTcM / 8) * citems;
    for (int it = w; it < items; it += 4) {
        const int rg = it / citems, cg = it - rg * citems;
        const int r = rg * 8 + r8, kc = cg * 4 + c4;
        if (kc < d4) {
            const float4 v = ld4(src + (int64_t)r * stride + 4 * kc);
            float4 h, l;
            split4(v, h, l);
            const int off = (kc * kTcM + r) * 4;
            *reinterpret_cast<float4*>(hi + off) = h;
            *reinterpret_cast<float4*>(lo + off) = l;
        }
    }
}

// 16 keys x Cp channels of V as the B operand [N = channels, K = keys] of O += P V.
// The tensor core wants B K-major (4 consecutive KEYS per 16 bytes; MN-major tf32 is
// not available without swizzling), V rows are channel-contiguous: each group of 4
// lanes loads a 4-key x 4-channel block (float4 per key), transposes it with three
// shuffles, and stores one float4 per channel:
//     byte(c, j) = (j/4)*(Cp*16) + c*16 + (j%4)*4        (LBO = Cp*16, SBO = 128)
// A warp item is 4 keys x 32 channels: 4 rows x 128 contiguous bytes from global,
// 512 contiguous bytes to shared memory (no bank conflicts).
__device__ __forceinline__ float4 transpose4(const float4& v, int jj, int lane) {
    // round r: read from lane (jj + r) % 4 of my 4-lane group the component that
    // lane selects for me: provider p serves receiver (p - r) % 4 with component (p - r) % 4
    float out[4];
    const float in[4] = {v.x, v.y, v.z, v.w};
    const int base = lane & ~3;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int e = (jj - r) & 3;                       // component I provide this round
        const float give = e == 0 ? in[0] : e == 1 ? in[1] : e == 2 ? in[2] : in[3];
        const float got = r == 0 ? give : __shfl_sync(0xffffffffu, give, base + ((jj + r) & 3));
        const int k = (jj + r) & 3;                        // key index of what I received
        if (k == 0) out[0] = got; else if (k == 1) out[1] = got; else if (k == 2) out[2] = got; else out[3] = got;
    }
    return make_float4(out[0], out[1], out[2], out[3]);
}

__device__ __forceinline__ void load_v_chunk(const float* __restrict__ src, int64_t stride, int C4, int Cp,
                                             float* hi, float* lo, int lt) {
    const int lane = lt & 31, w = lt >> 5;
    const int jj = lane & 3, c8 = lane >> 2;
    const int cblocks = (Cp + 31) >> 5;
    const int items = (kTcKC / 4) * cblocks;
    constexpr int kBatch = 5;
    for (int it0 = w; it0 < items; it0 += 4 * kBatch) {
        float4 v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int it = it0 + 4 * u;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (it < items) {
                const int j4 = it / cblocks, cb = it - j4 * cblocks;
                const int c4 = cb * 8 + c8;
                if (c4 < C4) v[u] = ld_stream4(src + (int64_t)(j4 * 4 + jj) * stride + 4 * c4);
            }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int it = it0 + 4 * u;
            if (it < items) {                               // warp-uniform
                const int j4 = it / cblocks, cb = it - j4 * cblocks;
                const float4 t = transpose4(v[u], jj, lane);   // channel cb*32 + lane, keys j4*4 .. +3
                const int c = cb * 32 + lane;
                if (c < Cp) {
                    float4 h, l;
                    split4(t, h, l);
                    const int off = (j4 * Cp + c) * 4;
                    *reinterpret_cast<float4*>(hi + off) = h;
                    *reinterpret_cast<float4*>(lo + off) = l;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kTcThreads, 1)
nonlocal_fwd_tc(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                const float* __restrict__ x, float* __restrict__ y, float* __restrict__ lse,
                const float* __restrict__ scale_p, const float* __restrict__ gamma_p, const TcGeom g) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ TcBars bars;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, i0 = blockIdx.x * kTcM;
    const int d4 = g.d >> 2, C4 = g.C >> 2;
    const int nkb = g.N / kTcKB;

    // ---- shared memory carve-up ---------------------------------------------
    const int qk_bytes = kTcM * g.d * 4;                 // one hi or lo tile of Q / K block
    float* sQh = reinterpret_cast<float*>(smem_raw);
    float* sQl = reinterpret_cast<float*>(smem_raw + qk_bytes);
    float* sKh = reinterpret_cast<float*>(smem_raw + 2 * qk_bytes);
    float* sKl = reinterpret_cast<float*>(smem_raw + 3 * qk_bytes);
    const int p_bytes = kTcM * kTcKC * 4;                // 8 KB
    const int v_bytes = kTcKC * g.Cp * 4;
    const int stage_bytes = 2 * p_bytes + 2 * v_bytes;
    uint8_t* stage0 = smem_raw + 4 * qk_bytes;
    auto sPh = [&](int s) { return reinterpret_cast<float*>(stage0 + s * stage_bytes); };
    auto sPl = [&](int s) { return reinterpret_cast<float*>(stage0 + s * stage_bytes + p_bytes); };
    auto sVh = [&](int s) { return reinterpret_cast<float*>(stage0 + s * stage_bytes + 2 * p_bytes); };
    auto sVl = [&](int s) { return reinterpret_cast<float*>(stage0 + s * stage_bytes + 2 * p_bytes + v_bytes); };

    // ---- one-time setup -------------------------------------------------------
    if (tid == 0) {
        mbar_init(&bars.k_full, 128);
        mbar_init(&bars.k_empty, 1);
        mbar_init(&bars.s_full, 1);
        mbar_init(&bars.s_empty, 128);
        mbar_init(&bars.o_full, 1);
        for (int s = 0; s < kTcMaxStages; ++s) {
            mbar_init(&bars.pv_full[s], 256);
            mbar_init(&bars.pv_empty[s], 1);
        }
        mbar_init_fence();
    }
    if (warp == 8) tmem_alloc(&bars.tmem_base, kTcTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars.tmem_base;

    const float* Qb = Q + ((int64_t)b * g.N + i0) * g.q_s;
    const float* Kb = K + (int64_t)b * g.N * g.k_s;
    const float* Vb = V + (int64_t)b * g.N * g.v_s;

    if (warp < 4) {
        // =====================================================================
        // softmax warps: thread r <-> query row i0 + r <-> TMEM lane r
        // =====================================================================
        const int r = tid;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        const float scale = (scale_p ? __ldg(scale_p) : 1.f) * 1.4426950408889634f;   // logits in log2 units
        float m = -INFINITY, l = 0.f;
        uint32_t n_s = 0;                       // S blocks consumed so far
        float t[kTcKB];

        // ---- pass A: statistics ----
        for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&bars.s_full, n_s & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < kTcKB / 32; ++c) tmem_ld32(trow + c * 32, t + c * 32);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars.s_empty);
            ++n_s;
            float bm = -INFINITY;
#pragma unroll
            for (int j = 0; j < kTcKB; ++j) { t[j] *= scale; bm = fmaxf(bm, t[j]); }
            const float mn = fmaxf(m, bm);
            float bs = 0.f;
#pragma unroll
            for (int j = 0; j < kTcKB; ++j) bs += exp2f(t[j] - mn);
            l = l * exp2f(m - mn) + bs;
            m = mn;
        }
        const float inv_l = 1.f / l;
        lse[(int64_t)b * g.N + i0 + r] = (m + log2f(l)) * 0.6931471805599453f;

        // ---- pass B: P tiles ----
        uint32_t gchunk = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&bars.s_full, n_s & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < kTcKB / 32; ++c) tmem_ld32(trow + c * 32, t + c * 32);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars.s_empty);
            ++n_s;
#pragma unroll
            for (int c = 0; c < kTcChunks; ++c, ++gchunk) {
                const int s = gchunk % g.stages;
                const uint32_t use = gchunk / g.stages;
                float4 ph[kTcKC / 4], pl[kTcKC / 4];
#pragma unroll
                for (int q4 = 0; q4 < kTcKC / 4; ++q4) {
                    float4 p;
                    p.x = exp2f(t[c * kTcKC + 4 * q4 + 0] * scale - m) * inv_l;
                    p.y = exp2f(t[c * kTcKC + 4 * q4 + 1] * scale - m) * inv_l;
                    p.z = exp2f(t[c * kTcKC + 4 * q4 + 2] * scale - m) * inv_l;
                    p.w = exp2f(t[c * kTcKC + 4 * q4 + 3] * scale - m) * inv_l;
                    split4(p, ph[q4], pl[q4]);
                }
                mbar_wait(&bars.pv_empty[s], (use & 1) ^ 1);
                float* ph_s = sPh(s);
                float* pl_s = sPl(s);
#pragma unroll
                for (int q4 = 0; q4 < kTcKC / 4; ++q4) {
                    *reinterpret_cast<float4*>(ph_s + (q4 * kTcM + r) * 4) = ph[q4];
                    *reinterpret_cast<float4*>(pl_s + (q4 * kTcM + r) * 4) = pl[q4];
                }
                fence_async_smem();
                mbar_arrive(&bars.pv_full[s]);
            }
        }

        // ---- epilogue: y = gamma * O + x ----
        const float gamma = __ldg(gamma_p);
        mbar_wait(&bars.o_full, 0);
        tc_fence_after();
        const float* xr = x + ((int64_t)b * g.N + i0 + r) * g.x_s;
        float* yr = y + ((int64_t)b * g.N + i0 + r) * g.y_s;
        for (int cc = 0; cc < g.Cp; cc += 16) {
            float o[16];
            tmem_ld16(trow + kTcOCol + cc, o);
            tmem_ld_wait();
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const int c = cc + 4 * q4;
                if (c < g.C) {
                    const float4 xi = ld4(xr + c);
                    float4 out;
                    out.x = fmaf(gamma, o[4 * q4 + 0], xi.x);
                    out.y = fmaf(gamma, o[4 * q4 + 1], xi.y);
                    out.z = fmaf(gamma, o[4 * q4 + 2], xi.z);
                    out.w = fmaf(gamma, o[4 * q4 + 3], xi.w);
                    *reinterpret_cast<float4*>(yr + c) = out;
                }
            }
        }
        tc_fence_before();
    } else if (warp < 8) {
        // =====================================================================
        // loader warps
        // =====================================================================
        const int lt = tid - 128;
        uint32_t n_k = 0;                       // K blocks loaded so far
        auto load_k = [&](int kb) {
            if (n_k > 0) mbar_wait(&bars.k_empty, (n_k - 1) & 1);
            load_kmajor_tile(Kb + (int64_t)kb * kTcKB * g.k_s, g.k_s, d4, sKh, sKl, lt);
            fence_async_smem();
            mbar_arrive(&bars.k_full);
            ++n_k;
        };
        load_kmajor_tile(Qb, g.q_s, d4, sQh, sQl, lt);
        for (int kb = 0; kb < nkb; ++kb) load_k(kb);          // pass A
        load_k(0);                                             // pass B, first block
        uint32_t gchunk = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            for (int c = 0; c < kTcChunks; ++c, ++gchunk) {
                const int s = gchunk % g.stages;
                const uint32_t use = gchunk / g.stages;
                mbar_wait(&bars.pv_empty[s], (use & 1) ^ 1);
                load_v_chunk(Vb + (int64_t)(kb * kTcKB + c * kTcKC) * g.v_s, g.v_s, C4, g.Cp, sVh(s), sVl(s), lt);
                fence_async_smem();
                mbar_arrive(&bars.pv_full[s]);
                if (c == 1 && kb + 1 < nkb) load_k(kb + 1);    // prefetch the next key block
            }
        }
    } else {
      if (lane == 0) {
        // =====================================================================
        // MMA issuer (one thread)
        // =====================================================================
        const uint32_t idesc_s = idesc_tf32(kTcM, kTcKB, false, false);
        const uint32_t idesc_o1 = idesc_tf32(kTcM, g.n1, false, false);
        const uint32_t idesc_o2 = idesc_tf32(kTcM, g.n2 > 0 ? g.n2 : 16, false, false);
        const uint32_t lbo_k = kTcM * 16, sbo = 128;            // K-major tiles with 128 rows
        const uint32_t aQh = smem_u32(sQh), aQl = smem_u32(sQl), aKh = smem_u32(sKh), aKl = smem_u32(sKl);
        const int ksteps = g.d >> 3;
        uint32_t n_k = 0, n_s = 0;
        auto issue_s = [&]() {
            mbar_wait(&bars.k_full, n_k & 1);
            ++n_k;
            if (n_s > 0) mbar_wait(&bars.s_empty, (n_s - 1) & 1);
            ++n_s;
            tc_fence_after();
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint32_t o = ks * 2 * lbo_k;
                const uint64_t qh = smem_desc(aQh + o, lbo_k, sbo), ql = smem_desc(aQl + o, lbo_k, sbo);
                const uint64_t kh = smem_desc(aKh + o, lbo_k, sbo), kl = smem_desc(aKl + o, lbo_k, sbo);
                mma_tf32(tmem, ql, kh, idesc_s, ks > 0);
                mma_tf32(tmem, qh, kl, idesc_s, true);
                mma_tf32(tmem, qh, kh, idesc_s, true);
            }
            tc_commit(&bars.k_empty);
            tc_commit(&bars.s_full);
        };
        for (int kb = 0; kb < nkb; ++kb) issue_s();            // pass A
        const uint32_t lbo_v = g.Cp * 16;                     // V tile: K-major with Cp rows
        uint32_t gchunk = 0;
        for (int kb = 0; kb < nkb; ++kb) {                      // pass B
            issue_s();
            for (int c = 0; c < kTcChunks; ++c, ++gchunk) {
                const int s = gchunk % g.stages;
                const uint32_t use = gchunk / g.stages;
                mbar_wait(&bars.pv_full[s], use & 1);
                tc_fence_after();
                const uint32_t aPh = smem_u32(sPh(s)), aPl = smem_u32(sPl(s));
                const uint32_t aVh = smem_u32(sVh(s)), aVl = smem_u32(sVl(s));
#pragma unroll
                for (int ks = 0; ks < kTcKC / 8; ++ks) {
                    const uint32_t po = ks * 2 * lbo_k, vo = ks * 2 * lbo_v;
                    const uint64_t ph = smem_desc(aPh + po, lbo_k, sbo), pl = smem_desc(aPl + po, lbo_k, sbo);
                    const bool acc = (gchunk | ks) != 0;
                    {
                        const uint64_t vh = smem_desc(aVh + vo, lbo_v, sbo), vl = smem_desc(aVl + vo, lbo_v, sbo);
                        mma_tf32(tmem + kTcOCol, pl, vh, idesc_o1, acc);
                        mma_tf32(tmem + kTcOCol, ph, vl, idesc_o1, true);
                        mma_tf32(tmem + kTcOCol, ph, vh, idesc_o1, true);
                    }
                    if (g.n2 > 0) {
                        const uint32_t ho = g.n1 * 16;               // row n1 of the V tile
                        const uint64_t vh = smem_desc(aVh + vo + ho, lbo_v, sbo);
                        const uint64_t vl = smem_desc(aVl + vo + ho, lbo_v, sbo);
                        mma_tf32(tmem + kTcOCol + g.n1, pl, vh, idesc_o2, acc);
                        mma_tf32(tmem + kTcOCol + g.n1, ph, vl, idesc_o2, true);
                        mma_tf32(tmem + kTcOCol + g.n1, ph, vh, idesc_o2, true);
                    }
                }
                tc_commit(&bars.pv_empty[s]);
            }
        }
        tc_commit(&bars.o_full);
        tc_fence_before();
      }
      __syncwarp();
    }

    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem, kTcTmemCols);
    }
}

// ---- host side ----------------------------------------------------------------

static bool tc_geom(int N, int d, int C, TcGeom& g) {
    if (N % kTcKB || N < kTcKB || d % 8 || d < 8 || d > 32 || C % 4 || C < 16 || C > 320) return false;
    g.N = N; g.d = d; g.C = C;
    g.Cp = (C + 15) / 16 * 16;
    if (g.Cp <= 256) { g.n1 = g.Cp; g.n2 = 0; }
    else { g.n1 = ((g.Cp / 2) + 15) / 16 * 16; g.n2 = g.Cp - g.n1; }
    const int fixed = 4 * kTcM * d * 4;
    const int stage = 2 * kTcM * kTcKC * 4 + 2 * kTcKC * g.Cp * 4;
    int stages = (227 * 1024 - 1024 - fixed) / stage;
    if (stages > kTcMaxStages) stages = kTcMaxStages;
    if (stages < 2) return false;
    g.stages = stages;
    return true;
}

bool nonlocal_tc_supported(int B, int N, int d, int C) {
    TcGeom g;
    return B > 0 && tc_geom(N, d, C, g);
}

int nonlocal_fwd_tc_launch(const float* Q, const float* K, const float* V, const float* x, float* y,
                           float* lse, int B, int N, int d, int C, int64_t q_s, int64_t k_s, int64_t v_s,
                           int64_t x_s, int64_t y_s, const float* scale, const float* gamma, cudaStream_t st) {
    TcGeom g;
    if (!tc_geom(N, d, C, g)) {
        set_error("davi_nonlocal_attn_fwd: shape not supported by the tensor-core kernel");
        return DAVI_EINVAL;
    }
    g.q_s = q_s; g.k_s = k_s; g.v_s = v_s; g.x_s = x_s; g.y_s = y_s;
    const size_t smem = (size_t)4 * kTcM * d * 4 +
                        (size_t)g.stages * (2 * kTcM * kTcKC * 4 + 2 * kTcKC * g.Cp * 4);
    DAVI_CUDA(cudaFuncSetAttribute(nonlocal_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nonlocal_fwd_tc<<<dim3(N / kTcM, B), kTcThreads, smem, st>>>(Q, K, V, x, y, lse, scale, gamma, g);
    return check_launch("davi_nonlocal_attn_fwd");
}

}  // namespace davi
