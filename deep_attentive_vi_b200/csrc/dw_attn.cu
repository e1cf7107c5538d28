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
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "davi_common.cuh"

namespace davi {

// Optional phase trace for profiling sessions (tools/dw_trace.py builds a SECOND library with -DDAVI_DW_TRACE; the
// product build compiles none of it): clock64 stamps of the first and the last CTA of a persistent launch.
#ifdef DAVI_DW_TRACE
__device__ long long g_dw_trace[2][64];
#define DW_STAMP(i)                                                                          \
    do {                                                                                     \
        if (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) g_dw_trace[blockIdx.x != 0][i] = clock64(); \
    } while (0)
// DW_WAIT(i, stmt): runs stmt and adds its duration to trace slot i (a role's time blocked on a barrier)
#define DW_WAIT(i, stmt)                                                                      \
    do {                                                                                      \
        const long long t0_ = clock64();                                                      \
        stmt;                                                                                 \
        if ((threadIdx.x & 31) == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))      \
            g_dw_trace[blockIdx.x != 0][i] += clock64() - t0_;                                \
    } while (0)
// every lane runs stmt; the lanes for which cond holds (one per role) record its duration
#define DW_WAIT_IF(cond, i, stmt)                                                             \
    do {                                                                                      \
        const long long t0_ = clock64();                                                      \
        stmt;                                                                                 \
        if ((cond) && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))                       \
            g_dw_trace[blockIdx.x != 0][i] += clock64() - t0_;                                \
    } while (0)
#define DW_TRACE_CLEAR()                                                                      \
    do {                                                                                      \
        if (threadIdx.x < 64 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))             \
            g_dw_trace[blockIdx.x != 0][threadIdx.x] = 0;                                     \
        __syncthreads();                                                                      \
    } while (0)
#else
#define DW_STAMP(i) \
    do {            \
    } while (0)
#define DW_WAIT(i, stmt) \
    do {                 \
        stmt;            \
    } while (0)
#define DW_WAIT_IF(cond, i, stmt) \
    do {                          \
        stmt;                     \
    } while (0)
#define DW_TRACE_CLEAR() \
    do {                 \
    } while (0)
#endif

// every slot carries its own (batch, pixel) strides: a slot may be a channel slice of a wider
// per-layer tensor, a dense fresh tensor, or layer j of a stacked tensor
struct SlotPtrs {
    const float* k[DAVI_MAX_SLOTS];
    const float* v[DAVI_MAX_SLOTS];
    int k_sb[DAVI_MAX_SLOTS], k_sp[DAVI_MAX_SLOTS];
    int v_sb[DAVI_MAX_SLOTS], v_sp[DAVI_MAX_SLOTS];
};
struct SlotGradPtrs {
    float* dk[DAVI_MAX_SLOTS];
    float* dv[DAVI_MAX_SLOTS];
    int k_sb[DAVI_MAX_SLOTS], k_sp[DAVI_MAX_SLOTS];
    int v_sb[DAVI_MAX_SLOTS], v_sp[DAVI_MAX_SLOTS];
};

struct DwGeom {
    int n, P, total_pix, d4, C4;
    int64_t q_sp, o_sp;
    uint32_t acc_mask;       // backward: bit j set = slot j's dK/dV are accumulated (+=), else overwritten
    const float* scale_ptr;  // device scalar, nullptr = 1.0
    float* attn_out;         // backward, optional [total_pix, n]: the recomputed softmax weights a_j
    float* ds_out;           // backward, optional [total_pix, n]: scale * ds_j (the weight of q in dK_j)
};

// per-CTA copy of the slot table with the pixel offsets of THIS thread's pixel resolved lazily
struct SlotTable {
    const float* k[DAVI_MAX_SLOTS];
    const float* v[DAVI_MAX_SLOTS];
    int k_sb[DAVI_MAX_SLOTS], k_sp[DAVI_MAX_SLOTS];
    int v_sb[DAVI_MAX_SLOTS], v_sp[DAVI_MAX_SLOTS];
};

constexpr int kDwThreads = 128;
constexpr int kDwUnroll = 4;  // slots whose value rows are in flight together

// softmax over the n slots of one pixel; lane l of the group holds slots l+t*G.
// returns a[t] (weights) and, if R != nullptr, the raw dot products r[t].
template <int G, int SPL>
__device__ __forceinline__ void dw_scores(const SlotTable& tb, const DwGeom& g, int l,
                                          int b, int p, const float* qp, float (&a)[SPL],
                                          float* r_raw, unsigned gm, float scale, float* k_stash = nullptr) {
    float s[SPL];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < SPL; ++t) {
        const int j = l + t * G;
        s[t] = -INFINITY;
        if (j < g.n) {
            const float* kp = tb.k[j] + (int64_t)b * tb.k_sb[j] + (int64_t)p * tb.k_sp[j];
            float dot = 0.f;
            if (k_stash) {      // keep this pixel's key rows on chip for the dq / dK phase: [slot][d]
                // all loads of a batch are issued before the first shared-memory store (the compiler cannot
                // move a generic load above a generic store: one DRAM round trip per float4 otherwise)
                for (int c0 = 0; c0 < g.d4; c0 += 8) {
                    float4 kv[8], qv[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c0 + c < g.d4) { kv[c] = ld4(kp + 4 * (c0 + c)); qv[c] = ld4(qp + 4 * (c0 + c)); }
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c0 + c < g.d4) {
                            *reinterpret_cast<float4*>(k_stash + (j * g.d4 + c0 + c) * 4) = kv[c];
                            dot += dot4(kv[c], qv[c]);
                        }
                }
            } else
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

// UNROLL = slots whose value rows are in flight together.  Calls with few slots (n <= 4: the first
// layers, 24-60 MB) are bound by launch + wave count, not by bytes in flight: UNROLL 1 / 2 halves the
// registers (128 -> ~85 / ~100) so that the whole grid is resident in one or two waves instead of three.
template <int G, int ITERS, int UNROLL>
__device__ __forceinline__ void dw_attn_fwd_body(const SlotPtrs& sp, const float* __restrict__ q,
                                                 float* __restrict__ out, float* __restrict__ attn,
                                                 const DwGeom& g) {
    constexpr int SPL = DAVI_MAX_SLOTS / G;
    __shared__ SlotTable tb;
    if (threadIdx.x < DAVI_MAX_SLOTS) {
        const int j = threadIdx.x;
        tb.k[j] = sp.k[j]; tb.v[j] = sp.v[j];
        tb.k_sb[j] = sp.k_sb[j]; tb.k_sp[j] = sp.k_sp[j];
        tb.v_sb[j] = sp.v_sb[j]; tb.v_sp[j] = sp.v_sp[j];
    }
    __syncthreads();

    const int l = threadIdx.x & (G - 1);
    const int gbase = (threadIdx.x & 31) & ~(G - 1);  // first lane of my group in the warp
    const int pix = (blockIdx.x * kDwThreads + threadIdx.x) / G;
    if (pix >= g.total_pix) return;  // whole groups exit together (G divides 32)
    const int b = pix / g.P, p = pix - b * g.P;
    auto vrow = [&](int j) { return tb.v[j] + (int64_t)b * tb.v_sb[j] + (int64_t)p * tb.v_sp[j] + 4 * l; };
    const float* qp = q + (int64_t)pix * g.q_sp;

    float a[SPL];
    const unsigned gm = group_mask<G>();
    const float scale = g.scale_ptr ? __ldg(g.scale_ptr) : 1.f;
    dw_scores<G, SPL>(tb, g, l, b, p, qp, a, nullptr, gm, scale);
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
        for (; j + UNROLL <= jend; j += UNROLL) {
            float4 v[UNROLL][ITERS];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const float* vp = vrow(j + u);
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
                    if (it < ITERS - 1 || last_ok) v[u][it] = ld_stream4(vp + 4 * it * G);
                    else v[u][it] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const float aj = __shfl_sync(gm, a[t], gbase + ((j + u) & (G - 1)));
#pragma unroll
                for (int it = 0; it < ITERS; ++it) fma4(acc[it], aj, v[u][it]);
            }
        }
        for (; j < jend; ++j) {
            const float* vp = vrow(j);
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

template <int G, int ITERS>
__global__ void __launch_bounds__(kDwThreads)
dw_attn_fwd_kernel(const SlotPtrs sp, const float* __restrict__ q, float* __restrict__ out,
                   float* __restrict__ attn, const DwGeom g) {
    dw_attn_fwd_body<G, ITERS, 4>(sp, q, out, attn, g);
}
template <int G, int ITERS, int UNROLL>
__global__ void __launch_bounds__(kDwThreads, UNROLL == 1 ? 6 : 5)
dw_attn_fwd_small_kernel(const SlotPtrs sp, const float* __restrict__ q, float* __restrict__ out,
                         float* __restrict__ attn, const DwGeom g) {
    dw_attn_fwd_body<G, ITERS, UNROLL>(sp, q, out, attn, g);
}

// ---------------------------------------------------------------------------------------
// Forward, version 2: persistent, bulk-copy staged.  The register-streaming kernel above tops
// out near 5 TB/s (ncu: profiles/r01_ncu_dw_attn_v1_summary.txt -- 25 % occupancy, every warp
// stalled on its own loads, and a third partial wave).  Here one CTA per SM walks tiles of 8
// pixels; ONE thread feeds a 16-stage shared-memory ring with cp.async.bulk copies of whole
// value tiles (8 rows, ~9 KB, contiguous in HBM), four "score" warps compute the softmax
// weights of the NEXT tile from q and the key rows while four "stream" warps multiply-accumulate
// the staged value tiles of the current one.  ~150 KB per SM are in flight independently of
// register pressure, DRAM sees multi-KB sequential bursts, and there is no wave tail.
// ---------------------------------------------------------------------------------------
constexpr int kV2TP = 8;                  // pixels per tile
constexpr int kV2Stages = 16;
constexpr int kV2Threads = 288;           // warp 0: producer | warps 1-4: scores | warps 5-8: stream
constexpr int kV2Slack = 32;              // floats of row padding a single-span copy may carry along

struct V2Bars {
    uint64_t v_full[kV2Stages], v_empty[kV2Stages];
    uint64_t a_full[2], a_empty[2];
};

__device__ __forceinline__ uint32_t v2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void v2_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(v2_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void v2_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(v2_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void v2_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(v2_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __noinline__ void v2_wait_timed_out(const uint64_t* bar, uint32_t parity) {
    printf("davi dw_attn: mbarrier wait timed out (block %d thread %d barrier offset %u parity %u)\n", (int)blockIdx.x,
           (int)threadIdx.x, v2_smem_u32(bar), parity);
    __trap();                                                  // never hang the GPU: fail the launch instead
}
__device__ __forceinline__ void v2_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(v2_smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (t0 == 0) t0 = clock64();                           // the clock is only read once the first try failed
        else if (clock64() - t0 > 2000000000ll) v2_wait_timed_out(bar, parity);
    }
}
__device__ __forceinline__ void v2_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            v2_smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(v2_smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void v2_cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(v2_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// per-slot staging plan: a tile of slot j is either one contiguous span (row stride kept) or,
// when the slot's rows are far apart, kV2TP row copies packed densely
__device__ __forceinline__ bool v2_single_span(int v_sp, int C) { return v_sp - C <= kV2Slack; }

template <int ITERS>
__global__ void __launch_bounds__(kV2Threads, 1)
dw_attn_fwd_v2_kernel(const SlotPtrs sp, const float* __restrict__ q, float* __restrict__ out,
                      float* __restrict__ attn, const DwGeom g, const int ntiles) {
    constexpr int G = 16, SPL = DAVI_MAX_SLOTS / G;
    extern __shared__ __align__(128) uint8_t v2_smem[];
    __shared__ SlotTable tb;
    __shared__ V2Bars bars;
    __shared__ float s_a[2][kV2TP][DAVI_MAX_SLOTS];
    const int C = g.C4 * 4;
    const int stage_floats = kV2TP * (C + kV2Slack);
    float* ring = reinterpret_cast<float*>(v2_smem);

    const int tid = threadIdx.x, warp = tid >> 5;
    DW_TRACE_CLEAR();
    if (tid == 0) DW_STAMP(0);                                  // entry
    if (tid < DAVI_MAX_SLOTS) {
        const int j = tid;
        tb.k[j] = sp.k[j]; tb.v[j] = sp.v[j];
        tb.k_sb[j] = sp.k_sb[j]; tb.k_sp[j] = sp.k_sp[j];
        tb.v_sb[j] = sp.v_sb[j]; tb.v_sp[j] = sp.v_sp[j];
    }
    // barrier init spread over the producer warp's lanes (one thread doing all 36 took ~1 us of every launch)
    if (tid >= 32 && tid < 32 + kV2Stages) { v2_mbar_init(&bars.v_full[tid - 32], 1); v2_mbar_init(&bars.v_empty[tid - 32], 4); }
    if (tid >= 64 && tid < 66) { v2_mbar_init(&bars.a_full[tid - 64], 128); v2_mbar_init(&bars.a_empty[tid - 64], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (tid == 0) DW_STAMP(1);                                  // barriers + tables ready

    if (warp == 0) {
        // ---------------- producer: one thread issues every bulk copy ----------------
        if (tid == 0) {
            uint32_t gc = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int pix0 = t * kV2TP;
                const int b = pix0 / g.P, p0 = pix0 - b * g.P;
                for (int j = 0; j < g.n; ++j, ++gc) {
                    const int s = gc % kV2Stages;
                    if (gc == kV2Stages) DW_STAMP(2);           // producer: ring filled once
                    if (gc >= kV2Stages) DW_WAIT(4, v2_mbar_wait(&bars.v_empty[s], ((gc / kV2Stages) - 1) & 1));
                    const int vsp = tb.v_sp[j];
                    const float* src = tb.v[j] + (int64_t)b * tb.v_sb[j] + (int64_t)p0 * vsp;
                    float* dst = ring + (size_t)s * stage_floats;
                    if (v2_single_span(vsp, C)) {
                        const uint32_t bytes = (uint32_t)((kV2TP - 1) * vsp + C) * 4u;
                        v2_mbar_expect_tx(&bars.v_full[s], bytes);
                        v2_bulk_load(dst, src, bytes, &bars.v_full[s]);
                    } else {
                        v2_mbar_expect_tx(&bars.v_full[s], (uint32_t)(kV2TP * C) * 4u);
                        for (int r = 0; r < kV2TP; ++r)
                            v2_bulk_load(dst + r * C, src + (int64_t)r * vsp, (uint32_t)C * 4u, &bars.v_full[s]);
                    }
                }
            }
        }
    } else if (warp <= 4) {
        // ---------------- score warps: softmax weights of each tile -> s_a ----------------
        const int lt = tid - 32;
        const int pixel = lt >> 4, l = lt & 15;
        const unsigned gm = group_mask<G>();
        const float scale = g.scale_ptr ? __ldg(g.scale_ptr) : 1.f;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int pix = t * kV2TP + pixel;
            const int b = pix / g.P, p = pix - b * g.P;
            float a[SPL];
            dw_scores<G, SPL>(tb, g, l, b, p, q + (int64_t)pix * g.q_sp, a, nullptr, gm, scale);
            const int buf = it & 1;
            if (it >= 2) {
                DW_WAIT_IF(lt == 0, 6, v2_mbar_wait(&bars.a_empty[buf], ((it >> 1) - 1) & 1));
            }
#pragma unroll
            for (int u = 0; u < SPL; ++u) {
                const int j = l + u * G;
                if (j < g.n) {
                    s_a[buf][pixel][j] = a[u];
                    if (attn) attn[(int64_t)pix * g.n + j] = a[u];
                }
            }
            v2_mbar_arrive(&bars.a_full[buf]);          // arrive has release semantics for the stores above
            if (lt == 0 && it < 12) DW_STAMP(8 + it);           // score: weights of tile it published
        }
    } else {
        // ---------------- stream warps: out = sum_j a_j * V_j over the staged tiles ----------------
        const int lt = tid - 160;
        const int pixel = lt >> 4, l = lt & 15;
        const bool last_ok = (l + (ITERS - 1) * G) < g.C4;
        uint32_t gc = 0, it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int buf = it & 1;
            DW_WAIT_IF(lt == 0, 61, v2_mbar_wait(&bars.a_full[buf], (it >> 1) & 1));
            float4 acc[ITERS];
#pragma unroll
            for (int i = 0; i < ITERS; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j0 = 0; j0 < g.n; j0 += 2) {      // two staged tiles per round: their waits and loads overlap
                float4 v[2][ITERS];
                float aj[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int j = j0 + u;
                    aj[u] = 0.f;
                    if (j < g.n) {
                        const uint32_t c = gc + u;
                        const int s = c % kV2Stages;
                        const int vsp = tb.v_sp[j];
                        const int rs = v2_single_span(vsp, C) ? vsp : C;
                        aj[u] = s_a[buf][pixel][j];
                        DW_WAIT_IF(lt == 0, 62, v2_mbar_wait(&bars.v_full[s], (c / kV2Stages) & 1));
                        const float* row = ring + (size_t)s * stage_floats + pixel * rs + 4 * l;
#pragma unroll
                        for (int i = 0; i < ITERS; ++i)
                            v[u][i] = (i < ITERS - 1 || last_ok) ? *reinterpret_cast<const float4*>(row + 4 * i * G)
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                        __syncwarp();                              // one arrival per warp: all lanes have read the stage
                        if ((tid & 31) == 0) v2_mbar_arrive(&bars.v_empty[s]);
                    } else {
#pragma unroll
                        for (int i = 0; i < ITERS; ++i) v[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int i = 0; i < ITERS; ++i) fma4(acc[i], aj[u], v[u][i]);
                gc += min(2, g.n - j0);
            }
            v2_mbar_arrive(&bars.a_empty[buf]);
            float* op = out + (int64_t)(t * kV2TP + pixel) * g.o_sp + 4 * l;
#pragma unroll
            for (int i = 0; i < ITERS; ++i)
                if (i < ITERS - 1 || last_ok) st_stream4(op + 4 * i * G, acc[i]);
            if (lt == 0 && it < 12) DW_STAMP(50 + it);          // stream: tile it consumed and written
        }
    }
}

// Backward.  Per pixel: recompute a; pass 1 streams V once for da_j = <V_j, g>;
// ds_j = a_j (da_j - sum_k a_k da_k); pass 2 only writes: dV_j = a_j g,
// dK_j = scale ds_j q, dq = scale sum_j ds_j K_j (K rows re-read from L1/L2).
template <int G, int ITERS, int UNROLL>
__device__ __forceinline__ void dw_attn_bwd_body(const SlotPtrs& sp, const SlotGradPtrs& gp,
                                                 const float* __restrict__ q, const float* __restrict__ dout,
                                                 float* __restrict__ dq, float* __restrict__ dscale,
                                                 const DwGeom& g) {
    constexpr int SPL = DAVI_MAX_SLOTS / G;
    __shared__ SlotTable tb;
    __shared__ SlotGradPtrs tg;
    __shared__ float red[32];
    if (threadIdx.x < DAVI_MAX_SLOTS) {
        const int j = threadIdx.x;
        tb.k[j] = sp.k[j]; tb.v[j] = sp.v[j];
        tb.k_sb[j] = sp.k_sb[j]; tb.k_sp[j] = sp.k_sp[j];
        tb.v_sb[j] = sp.v_sb[j]; tb.v_sp[j] = sp.v_sp[j];
        tg.dk[j] = gp.dk[j]; tg.dv[j] = gp.dv[j];
        tg.k_sb[j] = gp.k_sb[j]; tg.k_sp[j] = gp.k_sp[j];
        tg.v_sb[j] = gp.v_sb[j]; tg.v_sp[j] = gp.v_sp[j];
    }
    __syncthreads();

    const int l = threadIdx.x & (G - 1);
    const int gbase = (threadIdx.x & 31) & ~(G - 1);
    const int pix = (blockIdx.x * kDwThreads + threadIdx.x) / G;
    const bool active = pix < g.total_pix;
    float dsc_local = 0.f;

    if (active) {
        const int b = pix / g.P, p = pix - b * g.P;
        auto vrow = [&](int j) { return tb.v[j] + (int64_t)b * tb.v_sb[j] + (int64_t)p * tb.v_sp[j] + 4 * l; };
        const float* qp = q + (int64_t)pix * g.q_sp;

        float a[SPL], r[SPL], da[SPL];
        const unsigned gm = group_mask<G>();
        const float scale = g.scale_ptr ? __ldg(g.scale_ptr) : 1.f;
        dw_scores<G, SPL>(tb, g, l, b, p, qp, a, r, gm, scale);

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
            for (; j + UNROLL <= jend; j += UNROLL) {
                float4 v[UNROLL][ITERS];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const float* vp = vrow(j + u);
#pragma unroll
                    for (int it = 0; it < ITERS; ++it) {
                        if (it < ITERS - 1 || last_ok) v[u][it] = ld_stream4(vp + 4 * it * G);
                        else v[u][it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    float part = 0.f;
#pragma unroll
                    for (int it = 0; it < ITERS; ++it) part += dot4(v[u][it], go[it]);
                    part = group_sum<G>(part, gm);
                    if (((j + u) & (G - 1)) == l) da[t] = part;
                }
            }
            for (; j < jend; ++j) {
                const float* vp = vrow(j);
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
        if (g.attn_out) {   // saved for the gather of the shared contexts' gradients (dw_gather_kernel)
#pragma unroll
            for (int t = 0; t < SPL; ++t) {
                const int j = l + t * G;
                if (j < g.n) {
                    g.attn_out[(int64_t)pix * g.n + j] = a[t];
                    g.ds_out[(int64_t)pix * g.n + j] = ds[t] * scale;
                }
            }
        }

        // pass 2: writes.  lanes < d4 own one float4 column of q / dq / dK rows.  The key rows are
        // re-read (L1/L2 hits) in batches of UNROLL BEFORE the stores of the batch, so that the
        // loads are independent instead of one dependent round trip per slot.
        const bool kcol = l < g.d4;
        float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), dqv = qv;
        if (kcol) qv = ld4(qp + 4 * l);
#pragma unroll
        for (int t = 0; t < SPL; ++t) {
            const int jend = min(g.n, (t + 1) * G);
            for (int j0 = t * G; j0 < jend; j0 += UNROLL) {
                float4 kv[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const int j = j0 + u;
                    kv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (kcol && j < jend)
                        kv[u] = ld4(tb.k[j] + (int64_t)b * tb.k_sb[j] + (int64_t)p * tb.k_sp[j] + 4 * l);
                }
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const int j = j0 + u;
                    if (j >= jend) break;
                    const int src = gbase + (j & (G - 1));
                    const float aj = __shfl_sync(gm, a[t], src);
                    const float dsj = __shfl_sync(gm, ds[t], src) * scale;
                    const bool acc = (g.acc_mask >> j) & 1u;
                    float* dvp = tg.dv[j];
                    if (dvp) {
                        dvp += (int64_t)b * tg.v_sb[j] + (int64_t)p * tg.v_sp[j] + 4 * l;
#pragma unroll
                        for (int it = 0; it < ITERS; ++it) {
                            if (it < ITERS - 1 || last_ok) {
                                float4 o;
                                if (acc) {
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
                        fma4(dqv, dsj, kv[u]);
                        float* dkp = tg.dk[j];
                        if (dkp) {
                            dkp += (int64_t)b * tg.k_sb[j] + (int64_t)p * tg.k_sp[j] + 4 * l;
                            float4 o;
                            if (acc) {
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
        }
        if (kcol && dq) *reinterpret_cast<float4*>(dq + (int64_t)pix * g.q_sp + 4 * l) = dqv;
    }

    if (dscale) {  // uniform branch: every thread of the block reaches the barrier
        const float tot = block_sum(dsc_local, red);
        if (threadIdx.x == 0) atomicAdd(dscale, tot);
    }
}

template <int G, int ITERS>
__global__ void __launch_bounds__(kDwThreads)
dw_attn_bwd_kernel(const SlotPtrs sp, const SlotGradPtrs gp, const float* __restrict__ q,
                   const float* __restrict__ dout, float* __restrict__ dq, float* __restrict__ dscale,
                   const DwGeom g) {
    dw_attn_bwd_body<G, ITERS, 4>(sp, gp, q, dout, dq, dscale, g);
}
template <int G, int ITERS, int UNROLL>
__global__ void __launch_bounds__(kDwThreads, UNROLL == 1 ? 6 : 5)
dw_attn_bwd_small_kernel(const SlotPtrs sp, const SlotGradPtrs gp, const float* __restrict__ q,
                         const float* __restrict__ dout, float* __restrict__ dq, float* __restrict__ dscale,
                         const DwGeom g) {
    dw_attn_bwd_body<G, ITERS, UNROLL>(sp, gp, q, dout, dq, dscale, g);
}



// ---------------------------------------------------------------------------------------
// Backward, version 2: the same persistent bulk-copy pipeline for the gradient (n >= 8, large calls).
// Per tile of 8 pixels the producer thread stages the dOut tile and then the n value tiles; the four
// "stream" warps take <V_j, g> of every staged tile (da_j, reduced over the 16 lanes of a pixel) and
// write dV_j = a_j g for the slots whose gradient this launch owns; the four "score" warps recompute
// the softmax of the NEXT tile while the current one streams and then finish the previous one from
// da: ds_j = a_j (da_j - sum a da), dq, dK_j, the saved a / scale*ds for the gather, dscale.
// No read-modify-write variant: accumulate masks stay on the register-streaming kernel.
// ---------------------------------------------------------------------------------------
struct V2BwdBars {
    uint64_t v_full[kV2Stages], v_empty[kV2Stages];
    uint64_t a_full[2], a_empty[2], da_full[2], da_empty[2], go_full[2], go_empty[2];
};

template <int ITERS>
__global__ void __launch_bounds__(kV2Threads, 1)
dw_attn_bwd_v2_kernel(const SlotPtrs sp, const SlotGradPtrs gp, const float* __restrict__ q,
                      const float* __restrict__ dout, float* __restrict__ dq, float* __restrict__ dscale,
                      const DwGeom g, const int ntiles) {
    constexpr int G = 16, SPL = DAVI_MAX_SLOTS / G;
    extern __shared__ __align__(128) uint8_t v2_smem[];
    __shared__ SlotTable tb;
    __shared__ SlotGradPtrs tg;
    __shared__ V2BwdBars bars;
    __shared__ float s_a[2][kV2TP][DAVI_MAX_SLOTS];
    __shared__ float s_da[2][kV2TP][DAVI_MAX_SLOTS];
    const int C = g.C4 * 4;
    const int stage_floats = kV2TP * (C + kV2Slack);
    float* ring = reinterpret_cast<float*>(v2_smem);
    // key and query rows of three tiles, [3][kV2TP][n + 1][d] (row n = the query): filled by cp.async one tile ahead
    // of the softmax phase, read by the softmax phase and again by the finish phase one tile later
    const int kst_pixel = (g.n + 1) * g.d4 * 4;
    float* gbuf = ring + (size_t)kV2Stages * stage_floats;      // dOut tiles of the two tiles in flight
    float* kstash = gbuf + (size_t)2 * stage_floats;

    const int tid = threadIdx.x, warp = tid >> 5;
    DW_TRACE_CLEAR();
    if (tid == 0) DW_STAMP(0);                                  // entry
    if (tid < DAVI_MAX_SLOTS) {
        const int j = tid;
        tb.k[j] = sp.k[j]; tb.v[j] = sp.v[j];
        tb.k_sb[j] = sp.k_sb[j]; tb.k_sp[j] = sp.k_sp[j];
        tb.v_sb[j] = sp.v_sb[j]; tb.v_sp[j] = sp.v_sp[j];
        tg.dk[j] = gp.dk[j]; tg.dv[j] = gp.dv[j];
        tg.k_sb[j] = gp.k_sb[j]; tg.k_sp[j] = gp.k_sp[j];
        tg.v_sb[j] = gp.v_sb[j]; tg.v_sp[j] = gp.v_sp[j];
    }
    if (tid >= 32 && tid < 32 + kV2Stages) { v2_mbar_init(&bars.v_full[tid - 32], 1); v2_mbar_init(&bars.v_empty[tid - 32], 1); }
    if (tid >= 64 && tid < 66) {
        const int i = tid - 64;
        v2_mbar_init(&bars.a_full[i], 128); v2_mbar_init(&bars.a_empty[i], 128);
        v2_mbar_init(&bars.da_full[i], 128); v2_mbar_init(&bars.da_empty[i], 128);
        v2_mbar_init(&bars.go_full[i], 1); v2_mbar_init(&bars.go_empty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (tid == 0) DW_STAMP(1);                                  // barriers + tables ready

    if (warp == 0) {
        // ---------------- producer: per tile the dOut tile, then the n value tiles ----------------
        if (tid == 0) {
            uint32_t gc = 0, it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int pix0 = t * kV2TP;
                const int b = pix0 / g.P, p0 = pix0 - b * g.P;
                {   // dOut rows (dense stride o_sp) -> the tile's own buffer
                    const int gb = it & 1;
                    if (it >= 2) v2_mbar_wait(&bars.go_empty[gb], ((it >> 1) - 1) & 1);
                    const int osp = (int)g.o_sp;
                    const float* src = dout + (int64_t)pix0 * osp;
                    float* dst = gbuf + (size_t)gb * stage_floats;
                    if (v2_single_span(osp, C)) {
                        const uint32_t bytes = (uint32_t)((kV2TP - 1) * osp + C) * 4u;
                        v2_mbar_expect_tx(&bars.go_full[gb], bytes);
                        v2_bulk_load(dst, src, bytes, &bars.go_full[gb]);
                    } else {
                        v2_mbar_expect_tx(&bars.go_full[gb], (uint32_t)(kV2TP * C) * 4u);
                        for (int r = 0; r < kV2TP; ++r)
                            v2_bulk_load(dst + r * C, src + (int64_t)r * osp, (uint32_t)C * 4u, &bars.go_full[gb]);
                    }
                }
                for (int j = 0; j < g.n; ++j, ++gc) {
                    const int s = gc % kV2Stages;
                    if (gc == kV2Stages) DW_STAMP(2);           // producer: ring filled once
                    if (gc >= kV2Stages) DW_WAIT(4, v2_mbar_wait(&bars.v_empty[s], ((gc / kV2Stages) - 1) & 1));
                    float* dst = ring + (size_t)s * stage_floats;
                    const int vsp = tb.v_sp[j];
                    const float* src = tb.v[j] + (int64_t)b * tb.v_sb[j] + (int64_t)p0 * vsp;
                    if (v2_single_span(vsp, C)) {
                        const uint32_t bytes = (uint32_t)((kV2TP - 1) * vsp + C) * 4u;
                        v2_mbar_expect_tx(&bars.v_full[s], bytes);
                        v2_bulk_load(dst, src, bytes, &bars.v_full[s]);
                    } else {
                        v2_mbar_expect_tx(&bars.v_full[s], (uint32_t)(kV2TP * C) * 4u);
                        for (int r = 0; r < kV2TP; ++r)
                            v2_bulk_load(dst + r * C, src + (int64_t)r * vsp, (uint32_t)C * 4u, &bars.v_full[s]);
                    }
                }
            }
            DW_STAMP(3);                                        // producer: everything issued
        }
    } else if (warp <= 4) {
        // ---------------- score warps: softmax of tile it, then the finish of tile it - 1 ----------------
        const int lt = tid - 32;
        const int pixel = lt >> 4, l = lt & 15;
        const int gbase = (tid & 31) & ~(G - 1);
        const unsigned gm = group_mask<G>();
        const float scale = g.scale_ptr ? __ldg(g.scale_ptr) : 1.f;
        const bool kcol = l < g.d4;
        float a_prev[SPL], r_prev[SPL];
        float dsc_local = 0.f;
        int prev_pix = -1;

        auto stash = [&](uint32_t itx) { return kstash + (size_t)((itx % 3) * kV2TP + pixel) * kst_pixel; };
        // cp.async of this lane's key rows (slots l, l + 16) and of float4 column l of the query row of tile t
        auto prefetch = [&](int t, uint32_t itx) {
            const int pix = t * kV2TP + pixel;
            const int b = pix / g.P, p = pix - b * g.P;
            float* st = stash(itx);
#pragma unroll
            for (int u = 0; u < SPL; ++u) {
                const int j = l + u * G;
                if (j < g.n) {
                    const float* kp = tb.k[j] + (int64_t)b * tb.k_sb[j] + (int64_t)p * tb.k_sp[j];
                    for (int c = 0; c < g.d4; ++c) v2_cp_async16(st + (j * g.d4 + c) * 4, kp + 4 * c);
                }
            }
            for (int c = l; c < g.d4; c += G) v2_cp_async16(st + (g.n * g.d4 + c) * 4, q + (int64_t)pix * g.q_sp + 4 * c);
            asm volatile("cp.async.commit_group;" ::: "memory");
        };

        auto finish = [&](int pix, uint32_t itp) {
            const int buf = itp & 1;
            const int b = pix / g.P, p = pix - b * g.P;
            DW_WAIT_IF(lt == 0, 5, v2_mbar_wait(&bars.da_full[buf], (itp >> 1) & 1));
            float da[SPL], ds[SPL];
            float dsum = 0.f;
#pragma unroll
            for (int u = 0; u < SPL; ++u) {
                const int j = l + u * G;
                da[u] = (j < g.n) ? s_da[buf][pixel][j] : 0.f;
                dsum += a_prev[u] * da[u];
            }
            v2_mbar_arrive(&bars.da_empty[buf]);
            dsum = group_sum<G>(dsum, gm);
#pragma unroll
            for (int u = 0; u < SPL; ++u) {
                ds[u] = a_prev[u] * (da[u] - dsum);
                dsc_local += ds[u] * r_prev[u];
                const int j = l + u * G;
                if (g.attn_out && j < g.n) {
                    g.attn_out[(int64_t)pix * g.n + j] = a_prev[u];
                    g.ds_out[(int64_t)pix * g.n + j] = ds[u] * scale;
                }
            }
            const float* ks = stash(itp);
            float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), dqv = qv;
            if (kcol) qv = *reinterpret_cast<const float4*>(ks + (g.n * g.d4 + l) * 4);
#pragma unroll
            for (int u = 0; u < SPL; ++u) {
                const int jend = min(g.n, (u + 1) * G);
                for (int j0 = u * G; j0 < jend; j0 += 4) {
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const int j = j0 + w;
                        if (j >= jend) break;
                        const float dsj = __shfl_sync(gm, ds[u], gbase + (j & (G - 1))) * scale;
                        if (kcol) {
                            fma4(dqv, dsj, *reinterpret_cast<const float4*>(ks + (j * g.d4 + l) * 4));
                            float* dkp = tg.dk[j];
                            if (dkp) {
                                dkp += (int64_t)b * tg.k_sb[j] + (int64_t)p * tg.k_sp[j] + 4 * l;
                                *reinterpret_cast<float4*>(dkp) =
                                    make_float4(dsj * qv.x, dsj * qv.y, dsj * qv.z, dsj * qv.w);
                            }
                        }
                    }
                }
            }
            if (kcol && dq) *reinterpret_cast<float4*>(dq + (int64_t)pix * g.q_sp + 4 * l) = dqv;
        };

        uint32_t it = 0;
        if ((int)blockIdx.x < ntiles) prefetch(blockIdx.x, 0);
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int pix = t * kV2TP + pixel;
            float a[SPL], r[SPL];
            const int buf = it & 1;
            // the stash buffer of tile it + 1 was last read by finish(it - 2), one iteration ago
            __syncwarp();
            if (t + (int)gridDim.x < ntiles) {
                prefetch(t + gridDim.x, it + 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncwarp();                                       // the pixel's 16 lanes share one warp
            {   // softmax over the staged key rows: lane l owns slots l, l + 16
                const float* st = stash(it);
                const float* qs = st + g.n * g.d4 * 4;
                float sc[SPL];
                float m = -INFINITY;
#pragma unroll
                for (int u = 0; u < SPL; ++u) {
                    const int j = l + u * G;
                    sc[u] = -INFINITY;
                    r[u] = 0.f;
                    if (j < g.n) {
                        float dot = 0.f;
                        for (int c = 0; c < g.d4; ++c)
                            dot += dot4(*reinterpret_cast<const float4*>(st + (j * g.d4 + c) * 4),
                                        *reinterpret_cast<const float4*>(qs + 4 * c));
                        r[u] = dot;
                        sc[u] = dot * scale;
                    }
                    m = fmaxf(m, sc[u]);
                }
                m = group_max<G>(m, gm);
                float sum = 0.f;
#pragma unroll
                for (int u = 0; u < SPL; ++u) {
                    a[u] = (l + u * G < g.n) ? __expf(sc[u] - m) : 0.f;
                    sum += a[u];
                }
                sum = group_sum<G>(sum, gm);
                const float inv = 1.f / sum;
#pragma unroll
                for (int u = 0; u < SPL; ++u) a[u] *= inv;
            }
            if (it >= 2) {
                DW_WAIT_IF(lt == 0, 6, v2_mbar_wait(&bars.a_empty[buf], ((it >> 1) - 1) & 1));
            }
#pragma unroll
            for (int u = 0; u < SPL; ++u) {
                const int j = l + u * G;
                if (j < g.n) s_a[buf][pixel][j] = a[u];
            }
            v2_mbar_arrive(&bars.a_full[buf]);
            if (lt == 0 && it < 12) DW_STAMP(8 + it);           // score: weights of tile it published
            if (it >= 1) finish(prev_pix, it - 1);
            if (lt == 0 && it >= 1 && it < 13) DW_STAMP(20 + it - 1);   // score: tile it - 1 finished (dq, dK written)
#pragma unroll
            for (int u = 0; u < SPL; ++u) { a_prev[u] = a[u]; r_prev[u] = r[u]; }
            prev_pix = pix;
        }
        if (it >= 1) finish(prev_pix, it - 1);
        if (lt == 0) DW_STAMP(33);                              // score: last tile finished
        if (dscale) {
            dsc_local = group_sum<32>(dsc_local);
            if ((tid & 31) == 0) atomicAdd(dscale, dsc_local);
        }
    } else {
        // ---------------- stream warps: da_j = <V_j, g>, dV_j = a_j g over the staged tiles ----------------
        // ONE warp owns a whole staged tile (ring stage c goes to warp c % 4), so four tiles are consumed
        // concurrently and the per-stage overhead (wait, release, reduction, addressing: ~130 instructions of a
        // lone warp, which bound the first version at ~700 cycles / stage) is paid once per 9 KB instead of four
        // times.  Lane = (pixel, quarter): a lane covers float4 columns quarter, quarter + 4, ... of its pixel's
        // row, the dot product is finished by two shuffles.
        const int lt = tid - 160;
        const int w = lt >> 5, lane = lt & 31;
        const int pixel = lane >> 2, qd = lane & 3;
        constexpr int NC = 4 * ITERS;                           // float4 columns per lane (C4 <= 16 ITERS)
        const int orow = v2_single_span((int)g.o_sp, C) ? (int)g.o_sp : C;
        uint32_t gc0 = 0, it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int buf = it & 1;
            const int pix = t * kV2TP + pixel;
            const int b = pix / g.P, p = pix - b * g.P;
            float4 go[NC];
            {
                DW_WAIT_IF(lt == 0, 60, v2_mbar_wait(&bars.go_full[buf], (it >> 1) & 1));
                const float* row = gbuf + (size_t)buf * stage_floats + pixel * orow + 4 * qd;
#pragma unroll
                for (int i = 0; i < NC; ++i)
                    go[i] = (qd + 4 * i < g.C4) ? *reinterpret_cast<const float4*>(row + 16 * i)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                __syncwarp();
                if (lane == 0) v2_mbar_arrive(&bars.go_empty[buf]);
            }
            if (lt == 0 && it < 12) DW_STAMP(36 + it);          // stream: dOut tile of tile it received
            DW_WAIT_IF(lt == 0, 61, v2_mbar_wait(&bars.a_full[buf], (it >> 1) & 1));
            if (it >= 2) v2_mbar_wait(&bars.da_empty[buf], ((it >> 1) - 1) & 1);
            // warp w takes the ring stages c with c % 4 == w (NOT the slots j with j % 4 == w): every ring stage then
            // has ONE consumer warp for all of its uses, which sees them in order.  With slot-indexed ownership the
            // next use of a stage could belong to another warp, and a parity wait issued a whole phase early (an
            // older copy landing after younger ones) passes at once -- a rare stale read + double release.
            for (int j = (int)((w - gc0) & 3u); j < g.n; j += 4) {
                const uint32_t c = gc0 + j;
                const int s = c % kV2Stages;
                const int vsp = tb.v_sp[j];
                const int rs = v2_single_span(vsp, C) ? vsp : C;
                float* dvp = tg.dv[j];
                const float aj = s_a[buf][pixel][j];
                DW_WAIT_IF(lt == 0, 62, v2_mbar_wait(&bars.v_full[s], (c / kV2Stages) & 1));
                const float* row = ring + (size_t)s * stage_floats + pixel * rs + 4 * qd;
                float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
                for (int i = 0; i < NC; i += 4) {
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        v[u] = (qd + 4 * (i + u) < g.C4) ? *reinterpret_cast<const float4*>(row + 16 * (i + u))
                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                    p0 += dot4(v[0], go[i]); p1 += dot4(v[1], go[i + 1]);
                    p2 += dot4(v[2], go[i + 2]); p3 += dot4(v[3], go[i + 3]);
                }
                __syncwarp();                                   // every lane has read the stage
                if (lane == 0) v2_mbar_arrive(&bars.v_empty[s]);
                float part = (p0 + p1) + (p2 + p3);
                part += __shfl_xor_sync(0xffffffffu, part, 1);
                part += __shfl_xor_sync(0xffffffffu, part, 2);
                if (qd == 0) s_da[buf][pixel][j] = part;
                if (dvp) {
                    dvp += (int64_t)b * tg.v_sb[j] + (int64_t)p * tg.v_sp[j] + 4 * qd;
#pragma unroll
                    for (int i = 0; i < NC; ++i)
                        if (qd + 4 * i < g.C4)
                            st_stream4(dvp + 16 * i, make_float4(aj * go[i].x, aj * go[i].y, aj * go[i].z, aj * go[i].w));
                }
            }
            gc0 += g.n;
            v2_mbar_arrive(&bars.a_empty[buf]);
            v2_mbar_arrive(&bars.da_full[buf]);
            if (lt == 0 && it < 12) DW_STAMP(50 + it);          // stream: tile it consumed
        }
    }
}

// ---------------------------------------------------------------------------------------
// n <= 5 (the first layers: 24-73 MB per call).  These calls are bound by the length of a CTA's dependent
// chain, not by bytes: the streaming kernels above take one DRAM round trip for the key rows, then one per batch
// of UNROLL value rows.  Here EVERY load of the pixel -- value rows of all slots, dOut, key rows, query -- is
// issued before the first use (nothing depends on the scores but the arithmetic), so a CTA pays ONE round trip.
// Lane l owns the score of slot l (n <= NMAX <= G); the slot table is read straight from the launch parameters.
// ---------------------------------------------------------------------------------------
template <int G, int ITERS, int NMAX>
__global__ void __launch_bounds__(kDwThreads, NMAX <= 2 ? 4 : 2)
dw_attn_fwd_tiny_kernel(const __grid_constant__ SlotPtrs sp, const float* __restrict__ q, float* __restrict__ out,
                        float* __restrict__ attn, const __grid_constant__ DwGeom g) {
    const int l = threadIdx.x & (G - 1);
    const int gbase = (threadIdx.x & 31) & ~(G - 1);
    const int pix = (blockIdx.x * kDwThreads + threadIdx.x) / G;
    if (pix >= g.total_pix) return;
    const int b = pix / g.P, p = pix - b * g.P;
    const bool last_ok = (l + (ITERS - 1) * G) < g.C4;
    const unsigned gm = group_mask<G>();

    float4 v[NMAX][ITERS];
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
        const float* vp = (j < g.n) ? sp.v[j] + (int64_t)b * sp.v_sb[j] + (int64_t)p * sp.v_sp[j] + 4 * l : nullptr;
#pragma unroll
        for (int it = 0; it < ITERS; ++it)
            v[j][it] = (vp && (it < ITERS - 1 || last_ok)) ? ld_stream4(vp + 4 * it * G) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float scale = g.scale_ptr ? __ldg(g.scale_ptr) : 1.f;
    float s = -INFINITY;
    if (l < g.n) {      // d <= 32 (dispatch): the whole key row and query row are requested at once
        const float* kp = sp.k[l] + (int64_t)b * sp.k_sb[l] + (int64_t)p * sp.k_sp[l];
        const float* qp = q + (int64_t)pix * g.q_sp;
        float4 kv[8], qv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c < g.d4) { kv[c] = ld4(kp + 4 * c); qv[c] = ld4(qp + 4 * c); }
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c < g.d4) dot += dot4(kv[c], qv[c]);
        s = dot * scale;
    }
    const float m = group_max<G>(s, gm);
    float a = (l < g.n) ? __expf(s - m) : 0.f;
    a *= 1.f / group_sum<G>(a, gm);
    if (attn && l < g.n) attn[(int64_t)pix * g.n + l] = a;

    float4 acc[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
        const float aj = __shfl_sync(gm, a, gbase + j);
#pragma unroll
        for (int it = 0; it < ITERS; ++it) fma4(acc[it], aj, v[j][it]);
    }
    float* op = out + (int64_t)pix * g.o_sp + 4 * l;
#pragma unroll
    for (int it = 0; it < ITERS; ++it)
        if (it < ITERS - 1 || last_ok) st_stream4(op + 4 * it * G, acc[it]);
}

template <int G, int ITERS, int NMAX>
__global__ void __launch_bounds__(kDwThreads, NMAX <= 2 ? 3 : 2)
dw_attn_bwd_tiny_kernel(const __grid_constant__ SlotPtrs sp, const __grid_constant__ SlotGradPtrs gp,
                        const float* __restrict__ q, const float* __restrict__ dout, float* __restrict__ dq,
                        float* __restrict__ dscale, const __grid_constant__ DwGeom g) {
    __shared__ float red[32];
    const int l = threadIdx.x & (G - 1);
    const int gbase = (threadIdx.x & 31) & ~(G - 1);
    const int pix = (blockIdx.x * kDwThreads + threadIdx.x) / G;
    float dsc_local = 0.f;
    if (pix < g.total_pix) {
        const int b = pix / g.P, p = pix - b * g.P;
        const bool last_ok = (l + (ITERS - 1) * G) < g.C4;
        const bool kcol = l < g.d4;
        const unsigned gm = group_mask<G>();
        const float* qp = q + (int64_t)pix * g.q_sp;

        float4 go[ITERS], v[NMAX][ITERS], kc[NMAX];
        {
            const float* gpx = dout + (int64_t)pix * g.o_sp + 4 * l;
#pragma unroll
            for (int it = 0; it < ITERS; ++it)
                go[it] = (it < ITERS - 1 || last_ok) ? ld_stream4(gpx + 4 * it * G) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            const bool on = j < g.n;
            const float* vp = on ? sp.v[j] + (int64_t)b * sp.v_sb[j] + (int64_t)p * sp.v_sp[j] + 4 * l : nullptr;
#pragma unroll
            for (int it = 0; it < ITERS; ++it)
                v[j][it] = (on && (it < ITERS - 1 || last_ok)) ? ld_stream4(vp + 4 * it * G) : make_float4(0.f, 0.f, 0.f, 0.f);
            // float4 column l of key row j, for dq (the same lines the score loads below touch)
            kc[j] = (on && kcol) ? ld4(sp.k[j] + (int64_t)b * sp.k_sb[j] + (int64_t)p * sp.k_sp[j] + 4 * l)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float scale = g.scale_ptr ? __ldg(g.scale_ptr) : 1.f;
        float r = 0.f, s = -INFINITY;
        if (l < g.n) {
            const float* kp = sp.k[l] + (int64_t)b * sp.k_sb[l] + (int64_t)p * sp.k_sp[l];
            float4 kv[8], qw[8];
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c < g.d4) { kv[c] = ld4(kp + 4 * c); qw[c] = ld4(qp + 4 * c); }
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c < g.d4) r += dot4(kv[c], qw[c]);
            s = r * scale;
        }
        const float4 qv = kcol ? ld4(qp + 4 * l) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float m = group_max<G>(s, gm);
        float a = (l < g.n) ? __expf(s - m) : 0.f;
        a *= 1.f / group_sum<G>(a, gm);

        float da = 0.f;
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            float part = 0.f;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) part += dot4(v[j][it], go[it]);
            part = group_sum<G>(part, gm);
            if (l == j) da = part;
        }
        const float dsum = group_sum<G>(a * da, gm);
        const float ds = a * (da - dsum);
        dsc_local = ds * r;
        if (g.attn_out && l < g.n) {
            g.attn_out[(int64_t)pix * g.n + l] = a;
            g.ds_out[(int64_t)pix * g.n + l] = ds * scale;
        }
        float4 dqv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            if (j >= g.n) break;
            const float aj = __shfl_sync(gm, a, gbase + j);
            const float dsj = __shfl_sync(gm, ds, gbase + j) * scale;
            float* dvp = gp.dv[j];
            if (dvp) {
                dvp += (int64_t)b * gp.v_sb[j] + (int64_t)p * gp.v_sp[j] + 4 * l;
#pragma unroll
                for (int it = 0; it < ITERS; ++it)
                    if (it < ITERS - 1 || last_ok)
                        st_stream4(dvp + 4 * it * G, make_float4(aj * go[it].x, aj * go[it].y, aj * go[it].z, aj * go[it].w));
            }
            if (kcol) {
                fma4(dqv, dsj, kc[j]);
                float* dkp = gp.dk[j];
                if (dkp)
                    *reinterpret_cast<float4*>(dkp + (int64_t)b * gp.k_sb[j] + (int64_t)p * gp.k_sp[j] + 4 * l) =
                        make_float4(dsj * qv.x, dsj * qv.y, dsj * qv.z, dsj * qv.w);
            }
        }
        if (kcol && dq) *reinterpret_cast<float4*>(dq + (int64_t)pix * g.q_sp + 4 * l) = dqv;
    }
    if (dscale) {  // uniform branch: every thread of the block reaches the barrier
        const float tot = block_sum(dsc_local, red);
        if (threadIdx.x == 0) atomicAdd(dscale, tot);
    }
}

// ---------------------------------------------------------------------------------------
// Gradient GATHER of the shared contexts (SURVEY.md 8f3, section 7 "backward bytes").  A context
// c[k] is a slot of every later layer i; its gradient is
//     dV_k = sum_i a_k^(i) * dOut^(i)          dK_k = sum_i (scale ds_k^(i)) * q^(i)
// With the weights a^(i), scale*ds^(i) [B,P,n_i] saved by each layer's backward kernel, this is the
// forward contraction again (a weighted sum of rows over a slot list) with GIVEN weights: every
// gradient row is WRITTEN ONCE instead of being read-modify-written by each of the L-1-k consumers.
// One launch handles several contexts (blockIdx.y = problem): the bottom-up tensors e[1..L] are all
// finished by layer 0's backward and are gathered together.
// ---------------------------------------------------------------------------------------
constexpr int kGatherMaxProblems = 32;
constexpr int kGatherMaxSlots = 320;

struct GatherSlot {
    const float* v;    // consumer's dOut rows [B*P, C], pixel stride v_sp
    const float* k;    // consumer's query rows [B*P, d], pixel stride k_sp
    const float* wa;   // consumer's a      + slot index, pixel stride wn
    const float* wd;   // consumer's scale*ds + slot index, pixel stride wn
    int v_sp, k_sp, wn, pad_;
};
struct GatherProblem {
    float* out_v;      // [B*P, C] pixel stride ov_sp
    float* out_k;      // [B*P, d] pixel stride ok_sp (nullable)
    int ov_sp, ok_sp, first, m;
};
struct GatherParams {
    GatherProblem prob[kGatherMaxProblems];
    GatherSlot slot[kGatherMaxSlots];
};

template <int G, int ITERS>
__global__ void __launch_bounds__(kDwThreads)
dw_gather_kernel(const __grid_constant__ GatherParams gp, const int total_pix, const int d4, const int C4) {
    constexpr int U = 4;
    const GatherProblem& pr = gp.prob[blockIdx.y];
    const int l = threadIdx.x & (G - 1);
    const int pix = (blockIdx.x * kDwThreads + threadIdx.x) / G;
    if (pix >= total_pix) return;
    const GatherSlot* sl = gp.slot + pr.first;
    const int m = pr.m;
    const bool last_ok = (l + (ITERS - 1) * G) < C4;

    float4 acc[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    int j = 0;
    for (; j + U <= m; j += U) {
        float4 v[U][ITERS];
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const GatherSlot& s = sl[j + u];
            w[u] = __ldg(s.wa + (int64_t)pix * s.wn);
            const float* vp = s.v + (int64_t)pix * s.v_sp + 4 * l;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                if (it < ITERS - 1 || last_ok) v[u][it] = ld_stream4(vp + 4 * it * G);
                else v[u][it] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int it = 0; it < ITERS; ++it) fma4(acc[it], w[u], v[u][it]);
    }
    for (; j < m; ++j) {
        const GatherSlot& s = sl[j];
        const float w = __ldg(s.wa + (int64_t)pix * s.wn);
        const float* vp = s.v + (int64_t)pix * s.v_sp + 4 * l;
        float4 v[ITERS];
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            if (it < ITERS - 1 || last_ok) v[it] = ld_stream4(vp + 4 * it * G);
            else v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int it = 0; it < ITERS; ++it) fma4(acc[it], w, v[it]);
    }
    float* op = pr.out_v + (int64_t)pix * pr.ov_sp + 4 * l;
#pragma unroll
    for (int it = 0; it < ITERS; ++it)
        if (it < ITERS - 1 || last_ok) st_stream4(op + 4 * it * G, acc[it]);

    if (pr.out_k) {   // key part: d floats per pixel; the float4 columns are spread over the group's lanes
        for (int c = l; c < d4; c += G) {
            float4 kacc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int jj = 0; jj < m; ++jj) {
                const GatherSlot& s = sl[jj];
                fma4(kacc, __ldg(s.wd + (int64_t)pix * s.wn), ld4(s.k + (int64_t)pix * s.k_sp + 4 * c));
            }
            *reinterpret_cast<float4*>(pr.out_k + (int64_t)pix * pr.ok_sp + 4 * c) = kacc;
        }
    }
}

template <int G, int ITERS>
static void launch_gather(const GatherParams& gp, int nprob, int total_pix, int d4, int C4, cudaStream_t st) {
    const int64_t threads = (int64_t)total_pix * G;
    const dim3 grid((unsigned)((threads + kDwThreads - 1) / kDwThreads), (unsigned)nprob);
    dw_gather_kernel<G, ITERS><<<grid, kDwThreads, 0, st>>>(gp, total_pix, d4, C4);
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
    if (!mult4(g.q_sp) || !mult4(g.o_sp) || g.q_sp < d || g.o_sp < C) {
        set_error("%s: q/out strides must be multiples of 4 floats and >= the row width", fn);
        return DAVI_EINVAL;
    }
    if ((int64_t)B * P > 0x7fffffff / 32) {
        set_error("%s: B*P too large", fn);
        return DAVI_EINVAL;
    }
    return DAVI_OK;
}

// one slot's strides -> the int fields of the launch structs (validated)
static bool put_strides(int64_t sb, int64_t sp, int width, int* o_sb, int* o_sp) {
    if (!mult4(sb) || !mult4(sp) || sp < width || sb < 0 || sb > 0x7fffffff || sp > 0x7fffffff) return false;
    *o_sb = (int)sb; *o_sp = (int)sp;
    return true;
}

template <int G, int ITERS>
static void launch_fwd(const SlotPtrs& sp, const float* q, float* out, float* attn, const DwGeom& g,
                       cudaStream_t st) {
    const int64_t threads = (int64_t)g.total_pix * G;
    const int grid = (int)((threads + kDwThreads - 1) / kDwThreads);
    if (g.d4 <= 8 && g.n <= 2) dw_attn_fwd_tiny_kernel<G, ITERS, 2><<<grid, kDwThreads, 0, st>>>(sp, q, out, attn, g);
    else if (g.d4 <= 8 && g.n <= 5) dw_attn_fwd_tiny_kernel<G, ITERS, 5><<<grid, kDwThreads, 0, st>>>(sp, q, out, attn, g);
    else if (g.n <= 2) dw_attn_fwd_small_kernel<G, ITERS, 1><<<grid, kDwThreads, 0, st>>>(sp, q, out, attn, g);
    else if (g.n <= 5) dw_attn_fwd_small_kernel<G, ITERS, 2><<<grid, kDwThreads, 0, st>>>(sp, q, out, attn, g);
    else dw_attn_fwd_kernel<G, ITERS><<<grid, kDwThreads, 0, st>>>(sp, q, out, attn, g);
}
template <int G, int ITERS>
static void launch_bwd(const SlotPtrs& sp, const SlotGradPtrs& gp, const float* q, const float* dout,
                       float* dq, float* dscale, const DwGeom& g, cudaStream_t st) {
    const int64_t threads = (int64_t)g.total_pix * G;
    const int grid = (int)((threads + kDwThreads - 1) / kDwThreads);
    if (g.acc_mask == 0 && g.d4 <= 8 && g.n <= 2) dw_attn_bwd_tiny_kernel<G, ITERS, 2><<<grid, kDwThreads, 0, st>>>(sp, gp, q, dout, dq, dscale, g);
    else if (g.acc_mask == 0 && g.d4 <= 8 && g.n <= 5) dw_attn_bwd_tiny_kernel<G, ITERS, 5><<<grid, kDwThreads, 0, st>>>(sp, gp, q, dout, dq, dscale, g);
    else if (g.n <= 2) dw_attn_bwd_small_kernel<G, ITERS, 1><<<grid, kDwThreads, 0, st>>>(sp, gp, q, dout, dq, dscale, g);
    else if (g.n <= 5) dw_attn_bwd_small_kernel<G, ITERS, 2><<<grid, kDwThreads, 0, st>>>(sp, gp, q, dout, dq, dscale, g);
    else dw_attn_bwd_kernel<G, ITERS><<<grid, kDwThreads, 0, st>>>(sp, gp, q, dout, dq, dscale, g);
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

// davi_set_option(DAVI_OPT_DW_IMPL, 1) selects the register-streaming kernels for every shape
static bool dw_v2_enabled() { return get_option(DAVI_OPT_DW_IMPL) != 1; }

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
    // v2 (bulk-copy staged, persistent) wins once a call moves >= ~130 MB (measured: n >= 10 at B*P = 10240);
    // below that the register-streaming kernel has the shorter start-up.  16 ring stages must fit 227 KB.
    if (dw_v2_enabled() && g.C4 > 32 && g.C4 <= 80 && P % kV2TP == 0 && n >= 8 &&
        (int64_t)n * g.total_pix >= 10 * 10240) {
        int dev = 0, sms = 0;
        DAVI_CUDA(cudaGetDevice(&dev));
        DAVI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int ntiles = g.total_pix / kV2TP;
        const int grid = ntiles < sms ? ntiles : sms;
        const size_t smem = (size_t)kV2Stages * kV2TP * (C + kV2Slack) * sizeof(float);
        switch ((g.C4 + 15) / 16) {
#define V2_CASE(I)                                                                                         \
    case I:                                                                                                \
        DAVI_CUDA(cudaFuncSetAttribute(dw_attn_fwd_v2_kernel<I>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)smem));                                                        \
        dw_attn_fwd_v2_kernel<I><<<grid, kV2Threads, smem, st>>>(sp, q, out, attn, g, ntiles);             \
        break;
            V2_CASE(3) V2_CASE(4) V2_CASE(5) V2_CASE(6) V2_CASE(7) V2_CASE(8)
#undef V2_CASE
        }
        return check_launch(fn);
    }
    DW_DISPATCH(launch_fwd, sp, q, out, attn, g, st);
    return check_launch(fn);
}

static int dw_bwd_impl(const char* fn, const SlotPtrs& sp, const SlotGradPtrs& gp, const float* q,
                       const float* dout, float* dq, float* dscale, int B, int n, int P, int d, int C,
                       DwGeom g, cudaStream_t st) {
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
    // same crossover as the forward; the bulk-copy kernel has no read-modify-write (accumulate) variant
    if (dw_v2_enabled() && g.acc_mask == 0 && g.C4 > 32 && g.C4 <= 80 && P % kV2TP == 0 && n >= 6 &&
        (int64_t)n * g.total_pix >= 6 * 10240) {
        int dev = 0, sms = 0;
        DAVI_CUDA(cudaGetDevice(&dev));
        DAVI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int ntiles = g.total_pix / kV2TP;
        const int grid = ntiles < sms ? ntiles : sms;
        const size_t smem = (size_t)(kV2Stages + 2) * kV2TP * (C + kV2Slack) * sizeof(float) +
                            (size_t)3 * kV2TP * (n + 1) * d * sizeof(float);
        if (smem <= 220 * 1024)
        switch ((g.C4 + 15) / 16) {
#define V2B_CASE(I)                                                                                        \
    case I:                                                                                                \
        DAVI_CUDA(cudaFuncSetAttribute(dw_attn_bwd_v2_kernel<I>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)smem));                                                        \
        dw_attn_bwd_v2_kernel<I><<<grid, kV2Threads, smem, st>>>(sp, gp, q, dout, dq, dscale, g, ntiles);  \
        break;
            V2B_CASE(3) V2B_CASE(4) V2B_CASE(5) V2B_CASE(6) V2B_CASE(7) V2B_CASE(8)
#undef V2B_CASE
        }
        if (smem <= 220 * 1024) return check_launch(fn);
    }
    DW_DISPATCH(launch_bwd, sp, gp, q, dout, dq, dscale, g, st);
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
    for (int j = 0; j < n; ++j) {
        sp.k[j] = keys + j * k_sn; sp.v[j] = values + j * v_sn;
        DAVI_REQUIRE(put_strides(k_sb, k_sp, d, &sp.k_sb[j], &sp.k_sp[j]) &&
                     put_strides(v_sb, v_sp, C, &sp.v_sb[j], &sp.v_sp[j]), "bad key/value strides");
    }
    DwGeom g{}; g.q_sp = q_sp; g.o_sp = o_sp; g.scale_ptr = scale;
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
        DAVI_REQUIRE(put_strides(k_sb, k_sp, d, &sp.k_sb[j], &sp.k_sp[j]) &&
                     put_strides(v_sb, v_sp, C, &sp.v_sb[j], &sp.v_sp[j]), "bad key/value strides");
        gp.k_sb[j] = sp.k_sb[j]; gp.k_sp[j] = sp.k_sp[j]; gp.v_sb[j] = sp.v_sb[j]; gp.v_sp[j] = sp.v_sp[j];
    }
    DwGeom g{}; g.q_sp = q_sp; g.o_sp = o_sp; g.scale_ptr = scale;
    return dw_bwd_impl(__func__, sp, gp, q, dout, dq, dscale, B, n, P, d, C, g, (cudaStream_t)stream);
}

extern "C" int davi_dw_attn_slots_fwd(const float* q, const float* const* k_ptrs,
                                      const float* const* v_ptrs, const int64_t* k_sps, const int64_t* v_sps,
                                      float* out, float* attn, int B, int n, int P, int d, int C,
                                      int64_t q_sp, int64_t o_sp, const float* scale, davi_stream_t stream) {
    DAVI_REQUIRE(k_ptrs && v_ptrs && k_sps && v_sps, "null slot tables");
    DAVI_REQUIRE(n >= 1 && n <= DAVI_MAX_SLOTS, "n out of range");
    SlotPtrs sp{};
    for (int j = 0; j < n; ++j) {
        sp.k[j] = k_ptrs[j]; sp.v[j] = v_ptrs[j];
        DAVI_REQUIRE(put_strides(k_sps[j] * P, k_sps[j], d, &sp.k_sb[j], &sp.k_sp[j]) &&
                     put_strides(v_sps[j] * P, v_sps[j], C, &sp.v_sb[j], &sp.v_sp[j]), "bad slot strides");
    }
    DwGeom g{}; g.q_sp = q_sp; g.o_sp = o_sp; g.scale_ptr = scale;
    return dw_fwd_impl(__func__, sp, q, out, attn, B, n, P, d, C, g, (cudaStream_t)stream);
}

extern "C" int davi_dw_attn_slots_bwd(const float* q, const float* const* k_ptrs,
                                      const float* const* v_ptrs, const int64_t* k_sps, const int64_t* v_sps,
                                      const float* dout, float* dq, float* const* dk_ptrs, float* const* dv_ptrs,
                                      const int64_t* dk_sps, const int64_t* dv_sps, float* dscale, int B, int n,
                                      int P, int d, int C, int64_t q_sp, int64_t o_sp, const float* scale,
                                      uint32_t accumulate_mask, float* attn_out, float* ds_out,
                                      davi_stream_t stream) {
    DAVI_REQUIRE(k_ptrs && v_ptrs && k_sps && v_sps && dk_ptrs && dv_ptrs && dk_sps && dv_sps, "null slot tables");
    DAVI_REQUIRE((attn_out == nullptr) == (ds_out == nullptr), "attn_out and ds_out come together");
    DAVI_REQUIRE(n >= 1 && n <= DAVI_MAX_SLOTS, "n out of range");
    SlotPtrs sp{}; SlotGradPtrs gp{};
    for (int j = 0; j < n; ++j) {
        sp.k[j] = k_ptrs[j]; sp.v[j] = v_ptrs[j];
        gp.dk[j] = dk_ptrs[j]; gp.dv[j] = dv_ptrs[j];
        DAVI_REQUIRE(put_strides(k_sps[j] * P, k_sps[j], d, &sp.k_sb[j], &sp.k_sp[j]) &&
                     put_strides(v_sps[j] * P, v_sps[j], C, &sp.v_sb[j], &sp.v_sp[j]), "bad slot strides");
        DAVI_REQUIRE((!gp.dk[j] || put_strides(dk_sps[j] * P, dk_sps[j], d, &gp.k_sb[j], &gp.k_sp[j])) &&
                     (!gp.dv[j] || put_strides(dv_sps[j] * P, dv_sps[j], C, &gp.v_sb[j], &gp.v_sp[j])),
                     "bad slot gradient strides");
    }
    DwGeom g{}; g.q_sp = q_sp; g.o_sp = o_sp; g.scale_ptr = scale; g.acc_mask = accumulate_mask;
    g.attn_out = attn_out; g.ds_out = ds_out;
    return dw_bwd_impl(__func__, sp, gp, q, dout, dq, dscale, B, n, P, d, C, g, (cudaStream_t)stream);
}

extern "C" int davi_dw_attn_gather(int nprob, const int* m, float* const* out_v, float* const* out_k,
                                   const int64_t* out_v_sps, const int64_t* out_k_sps,
                                   const float* const* v_ptrs, const int64_t* v_sps,
                                   const float* const* k_ptrs, const int64_t* k_sps,
                                   const float* const* wa_ptrs, const float* const* wd_ptrs,
                                   const int64_t* w_sps, int B, int P, int d, int C, davi_stream_t stream) {
    DAVI_REQUIRE(nprob >= 1 && nprob <= kGatherMaxProblems, "1 <= nprob <= 32");
    DAVI_REQUIRE(m && out_v && out_k && out_v_sps && out_k_sps && v_ptrs && v_sps && k_ptrs && k_sps && wa_ptrs &&
                 wd_ptrs && w_sps, "null table");
    DAVI_REQUIRE(B > 0 && P > 0 && d > 0 && C > 0 && !(d & 3) && !(C & 3) && d <= 128 && C <= 512, "bad d / C");
    DAVI_REQUIRE((int64_t)B * P <= 0x7fffffff / 32, "B*P too large");
    GatherParams gp{};
    int tot = 0;
    for (int p = 0; p < nprob; ++p) {
        DAVI_REQUIRE(m[p] >= 1 && tot + m[p] <= kGatherMaxSlots, "consumer counts out of range");
        DAVI_REQUIRE(out_v[p] && aligned16(out_v[p]) && aligned16(out_k[p]), "output pointer null / unaligned");
        DAVI_REQUIRE(mult4(out_v_sps[p]) && out_v_sps[p] >= C && out_v_sps[p] <= 0x7fffffff &&
                     (!out_k[p] || (mult4(out_k_sps[p]) && out_k_sps[p] >= d && out_k_sps[p] <= 0x7fffffff)),
                     "bad output strides");
        gp.prob[p].out_v = out_v[p]; gp.prob[p].out_k = out_k[p];
        gp.prob[p].ov_sp = (int)out_v_sps[p]; gp.prob[p].ok_sp = out_k[p] ? (int)out_k_sps[p] : 0;
        gp.prob[p].first = tot; gp.prob[p].m = m[p];
        for (int j = 0; j < m[p]; ++j, ++tot) {
            GatherSlot& s = gp.slot[tot];
            s.v = v_ptrs[tot]; s.k = k_ptrs[tot]; s.wa = wa_ptrs[tot]; s.wd = wd_ptrs[tot];
            DAVI_REQUIRE(s.v && s.wa && aligned16(s.v) && (!out_k[p] || (s.k && s.wd && aligned16(s.k))),
                         "consumer pointer null / unaligned");
            DAVI_REQUIRE(mult4(v_sps[tot]) && v_sps[tot] >= C && v_sps[tot] <= 0x7fffffff &&
                         (!out_k[p] || (mult4(k_sps[tot]) && k_sps[tot] >= d && k_sps[tot] <= 0x7fffffff)) &&
                         w_sps[tot] >= 1 && w_sps[tot] <= DAVI_MAX_SLOTS, "bad consumer strides");
            s.v_sp = (int)v_sps[tot]; s.k_sp = out_k[p] ? (int)k_sps[tot] : 0; s.wn = (int)w_sps[tot];
        }
    }
    DwGeom g{}; g.C4 = C / 4;
    const int total_pix = B * P, d4 = d / 4;
    DW_DISPATCH(launch_gather, gp, nprob, total_pix, d4, g.C4, (cudaStream_t)stream);
    return check_launch(__func__);
}

#ifdef DAVI_DW_TRACE
// trace build only (tools/dw_trace.py): copies the [2][64] clock64 stamps of the last persistent launch to the host
extern "C" int davi_dw_trace_read(long long* host_2x64) {
    return cudaMemcpyFromSymbol(host_2x64, davi::g_dw_trace, sizeof(long long) * 128) == cudaSuccess ? 0 : -1;
}
#endif
