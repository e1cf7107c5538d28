// Row f1: the dense convolutions of the ResNet cells (reference layers/resnet.py:589-601, 613-633: the
// Keras Conv2D layers behind WeightNorm, layers/weight_norm.py:126-135) as an implicit GEMM on the
// 5th-gen tensor cores -- forward, input gradient (the forward kernel on transposed + flipped weights)
// and weight gradient.  Stride 1, "same" padding, channels-last fp32.
//
// Precision: the reference convolves in fp32 (TF 2.3 has no TF32), so every product is error-compensated
// exactly like the nonlocal kernels (nonlocal_tc.cu): v = hi + lo, hi = the fp32 bit pattern (the tensor
// core truncates it to tf32), lo = rna_tf32(v - trunc(v)); hi*hi' + lo*hi' + hi*lo' accumulated in fp32 in
// TMEM.  The lo images are computed ON CHIP from the tiles TMA delivered -- the A tile goes to TMEM as hi | lo
// (tcgen05.st; the MMAs read A from TMEM, which took 100 KB per chunk of operand reads off shared memory), the lo
// image of B is a shared -> shared pass -- no split tensors in HBM (the cuDNN path needed davi_split_tf32: 24 B
// written per element) and no tripled contraction axis.  Weights that a caller keeps across many calls can bring their
// lo image along (davi_tf32_lo once per optimiser step, davi_conv2d_fwd_pre): the TMA producer then loads the lo tile
// straight into the lo ring slot and the lo warps only do the A pass -- same bits, taken when the ring has three slots.
//
// GEMM view.  MODE 0 (forward / input gradient):
//     D[m = pixel, n = c_out] = sum_{tap, c_in} X[pixel + tap offset, c_in] * Wt[c_out, tap, c_in]
//   A = activations, K-major: ONE 4-D TMA box [32 ch x W x 128/W rows x 1 image] per (tap, channel chunk),
//       start coordinate shifted by the tap -- the out-of-bounds zero fill of TMA IS the padding;
//   B = weights [c_out][tap][c_in] (OHWI, the layout the trainer keeps them in), K-major.
// MODE 1 (weight gradient):
//     D[m = (tap, c_in), n = c_out] = sum_{pixel} X[pixel + tap offset, c_in] * GY[pixel, c_out]
//   both operands have the channels contiguous and the pixels strided: B = gy as MN-major TMA boxes [32 ch x 32 pixels]
//   (32-byte-atom swizzle, the only MN-major layout 32-bit operands have, tc_common.cuh); A = plain boxes of x that
//   the lo warps transpose on their way into TMEM (thread = channel reads its 32 pixels).
//
// Work decomposition.  A CTA PAIR (cluster of 2, tcgen05.mma.cta_group::2, M = 256) owns 256 rows of D
// and ALL of its columns (N <= 384 fp32 columns of TMEM; the rest holds the A ring): each CTA stages its own 128 rows of A and HALF
// of B, so B is fetched from L2 once per pair (the single-CTA form would need ~30 B/clk/SM of L2
// bandwidth for the 3xTF32 operand stream, 70 % of the measured L2 limit).  A 16x16 model at batch 40
// has only 40 such tiles for 74 pairs, so the K loop is cut stream-K style: the (tile, k-chunk) units
// are dealt out in equal contiguous runs, a pair that starts in the middle of a tile writes its partial
// accumulator to the workspace, and the pair that holds the tile's first chunk adds those in a fixed
// order (deterministic) before the epilogue.
//
// Roles per CTA (448 threads):  warps 0-7 epilogue (TMEM lane quadrant = warp % 4 <-> 32 rows of the tile; the two
// warps of a quadrant take alternate groups of 32 columns), warps 8-11 lo pass (A tile -> TMEM as hi | lo, lo image of
// B in shared memory), warp 12 TMA producer, warp 13 MMA issuer (leader CTA only; also owns TMEM).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "davi_common.cuh"
#include "tc_common.cuh"

namespace davi {

using namespace tc;

// nonlocal_tc.cu; swizzle: 0 = 128B (K-major tiles), 1 = 128B with 32-byte atoms (MN-major tiles), 2 = none
int make_tmap_nd(CUtensorMap* tm, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, int swizzle);

// Optional event trace for profiling sessions (tools/conv_trace.py builds a SECOND library with -DDAVI_CV_TRACE; the
// product build compiles none of it): %globaltimer stamps of pair 0, [rank][role][chunk].
#ifdef DAVI_CV_TRACE
#ifndef DAVI_CV_TRACE_PAIR
#define DAVI_CV_TRACE_PAIR 0
#endif
__device__ unsigned long long g_cv_trace[2][12][64];
#define CV_STAMP(role, idx)                                                          \
    do {                                                                             \
        if (pair == DAVI_CV_TRACE_PAIR && (idx) < 64) {                              \
            unsigned long long t__;                                                  \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                   \
            g_cv_trace[rank][role][idx] = t__;                                       \
        }                                                                            \
    } while (0)
#else
#define CV_STAMP(role, idx) \
    do {                    \
    } while (0)
#endif

constexpr int kCvThreads = 448;
constexpr int kCvMaxStages = 6;                     // hi ring (TMA targets)
constexpr int kCvMaxSlots = 4;                      // lo ring
constexpr int kCvABytes = 128 * 32 * 4;             // one [128 x 32] fp32 operand tile (either major-ness)
constexpr int kCvFlagBytes = 1024;                  // workspace header: one uint32 per (pair, rank)
constexpr int kCvMaxPairs = 80;
constexpr int kCvStageBytes = 8 * 32 * 17 * 4;     // epilogue staging: 8 warps x [32][17] floats

struct CvGeom {
    int mode;                   // 0 forward / input gradient, 1 weight gradient
    int tiles, kch;             // 256-row tiles of D, k-chunks per tile
    int pairs;                  // CTA pairs launched
    int n_total, np0, np1;      // TMEM columns used; columns of the (up to) two MMAs per k-step
    int nb;                     // B rows (MODE 0) / columns (MODE 1) staged per CTA = n_total / 2
    int stages, slots;          // hi ring (TMA targets: A | B); lo ring: B lo image in smem + A hi|lo in TMEM
    int tile_ok;                // the staged [128 x Cout] output tile fits the operand rings
    int reduce;                 // 1: every segment is written to the workspace, conv_reduce_kernel sums them
    int blo;                    // MODE 0: 1 = the lo image of B comes from memory (davi_tf32_lo of the weights, loaded by
                                //   TMA straight into the lo ring) instead of the lo warps' shared -> shared pass
    int debug;                  // DAVI_OPT_CONV_DEBUG bits (profiling only): 1 no lo arithmetic, 2 one MMA of three, 4 no stores
    // convolution
    int H, W, Cin, Cout, kh, kw, pt, pl, taps, cch, last_ksteps;
    int rows_cta;               // image rows per CTA tile (MODE 0): 128 / W
    int tiles_per_image;        // MODE 0: H * W / 256
    int chunk_rows;             // MODE 1: image rows per 32-pixel chunk: 32 / W
    int chunks_per_image;       // MODE 1: H * W / 32
    int mblocks;                // MODE 1: taps * cch 32-row blocks of D
    int64_t out_sp;             // MODE 0: pixel stride of the output
};

struct CvArgs {
    int act;                    // 1: the input goes through swish on its way into the tensor core (conv(swish(x)),
                                //    layers/resnet.py:624-627); selects the ACT = 1 instance of the kernel
    const float* bias;          // MODE 0, nullable
    float* out;                 // MODE 0: [pixels, Cout];  MODE 1: gw [Cout][taps][Cin]
    uint32_t* flags;            // workspace header
    float* partials;            // workspace body: [slot][rank][n_pad / 4][128] float4 (n_pad = n_total rounded up to 32)
};

struct CvBars {
    uint64_t full[kCvMaxStages], empty[kCvMaxStages];
    uint64_t ready[kCvMaxSlots], lo_free[kCvMaxSlots];
    uint64_t acc_full, acc_free;
    uint32_t tmem_base;
};

// ---- cluster / pair plumbing (validated by tools/umma_probe3.cu) -------------------------------------
__device__ __forceinline__ uint32_t cv_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cv_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cv_tmem_alloc2(uint32_t* dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cv_tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void cv_mma2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// the same with the A operand (each CTA's own 128 rows x 8 columns per k-step) read from TMEM
__device__ __forceinline__ void cv_mma2_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// registers -> TMEM, 32 lanes x 32 consecutive columns (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void cv_tmem_st32(uint32_t taddr, const float* v) {
    const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
        "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]),
        "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]), "r"(u[17]),
        "r"(u[18]), "r"(u[19]), "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]), "r"(u[25]), "r"(u[26]),
        "r"(u[27]), "r"(u[28]), "r"(u[29]), "r"(u[30]), "r"(u[31])
        : "memory");
}
// the barrier at the same offset in BOTH CTAs arrives once every MMA issued so far has completed
__device__ __forceinline__ void cv_commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                     "r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
// arrive on the barrier at this offset in CTA `target` of the cluster.  Default semantics (release, CTA scope) on the
// mapped address, the form CUTLASS' ClusterBarrier::arrive uses between the CTAs of a pair: the consumers are the
// tensor core and threads that acquire the barrier afterwards; the data itself was pushed out by fence.proxy.async /
// tcgen05.wait::st before.  (An explicit .release.cluster here cost 0.4-0.9 us per arrive, profiles/r02_conv_trace*.)
__device__ __forceinline__ void cv_arrive_at(uint64_t* bar, uint32_t target, uint32_t self) {
    if (target == self) {
        mbar_arrive(bar);
        return;
    }
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(target));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ bool cv_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// waits for arrivals that may come from the peer CTA; traps (launch failure, never a hang) after ~2 s
__device__ __forceinline__ void cv_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (cv_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!cv_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void cv_tma_4d(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3,
                                          uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void cv_flag_wait(const uint32_t* f) {
    uint32_t v;
    const long long t0 = clock64();
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v) return;
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}

__device__ __forceinline__ float cv_swish(float v) { return __fdividef(v, 1.f + __expf(-v)); }
// ---- stream-K bookkeeping: every role of both CTAs walks the same list -------------------------------
struct CvSeg {
    int tile, k0, k1;
};
__device__ __forceinline__ int64_t cv_first_unit(int pair, int pairs, int64_t units) {
    return (int64_t)pair * units / pairs;
}
__device__ __forceinline__ int cv_pair_of(int64_t unit, int pairs, int64_t units) {
    int p = (int)(unit * pairs / units);
    while (p + 1 < pairs && cv_first_unit(p + 1, pairs, units) <= unit) ++p;
    while (p > 0 && cv_first_unit(p, pairs, units) > unit) --p;
    return p;
}
struct CvWalk {
    int64_t u, u1;
    int kch;
    __device__ CvWalk(int pair, const CvGeom& g) : kch(g.kch) {
        const int64_t units = (int64_t)g.tiles * g.kch;
        u = cv_first_unit(pair, g.pairs, units);
        u1 = cv_first_unit(pair + 1, g.pairs, units);
    }
    __device__ bool next(CvSeg& s) {
        if (u >= u1) return false;
        s.tile = (int)(u / kch);
        s.k0 = (int)(u - (int64_t)s.tile * kch);
        const int64_t left = u1 - u;
        s.k1 = (int)(left < kch - s.k0 ? s.k0 + left : kch);
        u += s.k1 - s.k0;
        return true;
    }
};

template <int MODE, int ACT>
// (14 warps: 4 + 4 + 3 + 3 per SM sub-partition of 16 K registers -> at most 128 registers per thread)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kCvThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmL,
                 const __grid_constant__ CUtensorMap tmL1, const CvArgs a, const CvGeom g) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ CvBars bars;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cv_ctarank();
    const int pair = blockIdx.x >> 1;
    const int b_bytes = g.nb * 128;
    const int half_bytes = kCvABytes + b_bytes;            // one stage of the hi ring: A | B as TMA wrote them
    auto sA = [&](int s) { return smem + (size_t)s * half_bytes; };
    auto sB = [&](int s) { return smem + (size_t)s * half_bytes + kCvABytes; };
    auto sLo = [&](int l) { return smem + (size_t)g.stages * half_bytes + (size_t)l * b_bytes; };   // lo image of B
    // epilogue staging: one [32 rows x 16 columns] block per epilogue warp, row pitch 17 floats (conflict-free both ways)
    float* sStage = reinterpret_cast<float*>(smem + (size_t)g.stages * half_bytes + (size_t)g.slots * b_bytes) + warp * (32 * 17);
    // TMEM: accumulator [0, n_total) | A ring: slot l = columns [acol + 64 l, +32) hi, [+32, +64) lo
    const uint32_t acol = (uint32_t)(512 - 64 * g.slots);
    const int ncol4 = (g.n_total + 31) / 32 * 8;           // float4 columns of one partial slot
    const int tile_pitch = ((g.Cout + 3) / 4 | 1) * 4;     // floats per row of the output tile staged in the rings

    if (tid == 0) {
        for (int s = 0; s < kCvMaxStages; ++s) {
            mbar_init(&bars.full[s], 1);
            mbar_init(&bars.empty[s], 1);
        }
        for (int l = 0; l < kCvMaxSlots; ++l) {
            mbar_init(&bars.ready[l], 8);                  // 4 lo warps x 2 CTAs (used in the leader only)
            mbar_init(&bars.lo_free[l], 1);
        }
        mbar_init(&bars.acc_full, 1);
        mbar_init(&bars.acc_free, 16);                     // 8 epilogue warps x 2 CTAs (leader only)
        mbar_init_fence();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmB1);
        tma_prefetch_desc(&tmL);
        tma_prefetch_desc(&tmL1);
    }
    if (warp == 13) cv_tmem_alloc2(&bars.tmem_base, 512);  // the same warp of BOTH CTAs
    tc_fence_before();
    __syncthreads();
    cv_cluster_sync();                                     // barriers of both CTAs live before any remote arrive
    tc_fence_after();
    const uint32_t tmem = bars.tmem_base;

    if (warp < 8) {
        // =====================================================================
        // epilogue: thread <-> TMEM lane r <-> row 128 * rank + r of the pair's tile; column groups c0 = 32 * (2 j + warp / 4)
        // =====================================================================
        const int r = (warp & 3) * 32 + lane, chalf = warp >> 2;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const float* __restrict__ bias = a.bias;
        const int64_t units = (int64_t)g.tiles * g.kch;
        CvWalk walk(pair, g);
        CvSeg sg;
        uint32_t si = 0;
        while (walk.next(sg)) {
            // reduce mode: segment `si` (0 or 1) of this pair goes to slot 2 * pair + si, nobody waits
            const bool owner = sg.k0 == 0 && !g.reduce;
            const bool last_seg = owner && g.tile_ok && walk.u >= walk.u1 && !(g.debug & 128);   // nothing follows: the rings are free
            float4* my_part = reinterpret_cast<float4*>(a.partials) +
                              ((size_t)(g.reduce ? 2 * pair + (int)si : pair) * 2 + rank) * ncol4 * 128;
            int ncontrib = 0;
            if (owner && sg.k1 < g.kch)
                ncontrib = cv_pair_of((int64_t)(sg.tile + 1) * g.kch - 1, g.pairs, units) - pair;
            if (g.debug & 64) ncontrib = 0;
            mbar_wait(&bars.acc_full, si & 1);
            tc_fence_after();
            if (tid == 0) CV_STAMP(6, 2 * si);
            if (owner)
                for (int q = 1; q <= ncontrib; ++q) cv_flag_wait(a.flags + (pair + q) * 2 + rank);
            if (tid == 0) CV_STAMP(6, 8 + 8 * si);
            // MODE 0: row r = pixel;  MODE 1: row r = (tap, c_in) of block 4 * rank + r / 32 of the tile
            int64_t out_row = 0;
            bool row_ok = true;
            if (MODE == 0) {
                out_row = ((int64_t)sg.tile * 256 + rank * 128 + r) * g.out_sp;
            } else {
                const int mb = sg.tile * 8 + (int)rank * 4 + (r >> 5);
                const int tap = mb / g.cch, cc = mb - tap * g.cch;
                const int ci = cc * 32 + (r & 31);
                row_ok = mb < g.mblocks && ci < g.Cin;
                out_row = (int64_t)tap * g.Cin + ci;
            }
            const float4* part1 = reinterpret_cast<const float4*>(a.partials) + ((size_t)(pair + 1) * 2 + rank) * ncol4 * 128 + r;
            float4 wn[8];
            if (owner && ncontrib > 0 && 32 * chalf < g.n_total) {
#pragma unroll
                for (int q = 0; q < 8; ++q) wn[q] = __ldcg(part1 + (size_t)(8 * chalf + q) * 128);
            }
            for (int c0 = 32 * chalf; c0 < g.n_total; c0 += 64) {
                float v[32];
                tmem_ld32(trow + c0, v);
                if (g.debug & 4) { tmem_ld_wait(); continue; }
                if (!owner) {
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        __stcg(my_part + (size_t)(c0 / 4 + q) * 128 + r,
                               make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
                    continue;
                }
                // the first contribution of this column group was requested one group ahead (wn), the second one (a tile cut
                // into three segments: the usual case at 74 pairs for 40 tiles) and the next group's first go out now,
                // before anything of this group is used: one L2 round trip per group instead of one per contribution
                float4 w2[8];
                const float4* part = part1 + (size_t)(c0 / 4) * 128;
                if (ncontrib > 1) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) w2[q] = __ldcg(part + (size_t)2 * ncol4 * 128 + q * 128);
                }
                tmem_ld_wait();
                if (ncontrib > 0) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        v[4 * q] += wn[q].x; v[4 * q + 1] += wn[q].y; v[4 * q + 2] += wn[q].z; v[4 * q + 3] += wn[q].w;
                    }
                    if (c0 + 64 < g.n_total) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) wn[q] = __ldcg(part + 16 * 128 + q * 128);
                    }
                }
                for (int p = 2; p <= ncontrib; ++p) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        v[4 * q] += w2[q].x; v[4 * q + 1] += w2[q].y; v[4 * q + 2] += w2[q].z; v[4 * q + 3] += w2[q].w;
                    }
                    if (p < ncontrib) {                                    // (a tile cut into more than three segments)
#pragma unroll
                        for (int q = 0; q < 8; ++q) w2[q] = __ldcg(part + (size_t)2 * ncol4 * 128 * p + q * 128);
                    }
                }
                if (MODE == 0 && last_seg) {
                    // the pair's last segment: the operand rings are idle, the whole [128 x Cout] tile is staged in them
                    // (row pitch = 4 * odd floats: conflict-free 16-byte accesses) and streamed out below
                    float* trow_s = reinterpret_cast<float*>(smem) + (size_t)r * tile_pitch + c0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        const int c = c0 + 4 * q;
                        if (bias && c < g.Cout) {
                            const float4 bb = ld4(bias + c);
                            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                        }
                        if (c < g.Cout) *reinterpret_cast<float4*>(trow_s + 4 * q) = o;
                    }
                    if (tid == 0) CV_STAMP(6, 9 + 8 * si + c0 / 64);
                } else if (MODE == 0) {
                    // mid-run tiles: transpose 16-column blocks through a small per-warp buffer so that every store
                    // instruction writes two 64-byte runs instead of 32 scattered 16-byte pieces
                    float* out0 = a.out + ((int64_t)sg.tile * 256 + rank * 128 + (warp & 3) * 32) * g.out_sp;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) sStage[lane * 17 + j] = v[16 * h + j];
                        __syncwarp();
                        const int c = c0 + 16 * h + (lane & 15);
                        if (c < g.Cout) {
                            const float bc = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int rr = 2 * j + (lane >> 4);
                                out0[(int64_t)rr * g.out_sp + c] = sStage[rr * 17 + (lane & 15)] + bc;
                            }
                        }
                        __syncwarp();
                    }
                } else if (row_ok) {
                    // lanes = 32 consecutive c_in of one (c_out, tap): every store instruction is one 128-byte line
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < g.Cout) a.out[(int64_t)(c0 + j) * g.taps * g.Cin + out_row] = v[j];
                }
            }
            if (MODE == 0 && last_seg) {
                asm volatile("bar.sync 1, 256;" ::: "memory");                 // the tile is complete in shared memory
                // flat float4 index over the [128 x Cout] tile: consecutive threads write consecutive 16-byte pieces.
                // (Fusing the block's shortcut or the activation gradient in here was measured and dropped: one CTA per
                // SM has too few loads in flight -- +20 us per launch against 6-7 us for a separate elementwise kernel.)
                const int C4 = g.Cout >> 2, total = 128 * C4;
                const float* tile_s = reinterpret_cast<const float*>(smem);
                float* out0 = a.out + ((int64_t)sg.tile * 256 + rank * 128) * g.out_sp;
                for (int i = tid; i < total; i += 256) {
                    const int rr = i / C4, c = (i - rr * C4) * 4;
                    st_stream4(out0 + (int64_t)rr * g.out_sp + c, *reinterpret_cast<const float4*>(tile_s + (size_t)rr * tile_pitch + c));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) cv_arrive_at(&bars.acc_free, 0, rank);            // TMEM may be overwritten
            if (tid == 0) CV_STAMP(6, 2 * si + 1);
            if (g.reduce) {
                // nothing to hand over inside this launch
            } else if (!owner) {
                __threadfence();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (tid == 0) {
                    uint32_t one = 1;
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.flags + pair * 2 + rank), "r"(one) : "memory");
                }
            } else if (ncontrib > 0) {
                asm volatile("bar.sync 1, 256;" ::: "memory");         // every thread has read the partials
                if (tid < ncontrib) a.flags[(pair + 1 + tid) * 2 + rank] = 0;   // clean for the next launch
            }
            ++si;
        }
    } else if (warp < 12) {
        // =====================================================================
        // lo pass: one sweep over the hi half (A | B) of the landed stage into its lo half
        // =====================================================================
        const int lt = tid - 256;                          // = row of this CTA's A tile = TMEM lane
        const int n16 = b_bytes / 16;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + acol;
        CvWalk walk(pair, g);
        CvSeg sg;
        uint32_t it = 0;
        while (walk.next(sg)) {
            for (int kc = sg.k0; kc < sg.k1; ++kc, ++it) {
                const int s = it % g.stages, l = it % g.slots;
                mbar_wait(&bars.full[s], (it / g.stages) & 1);
                if (lt == 0) CV_STAMP(1, it);
                if (it >= (uint32_t)g.slots) mbar_wait(&bars.lo_free[l], ((it / g.slots) - 1) & 1);
                tc_fence_after();
                if (lt == 0) CV_STAMP(7, it);
                // A: this thread's row (MODE 0: 32 channels of one pixel; MODE 1: 32 pixels of one channel) -> TMEM
                {
                    float hi[32], lo[32];
                    if (MODE == 0) {
                        const uint8_t* row = sA(s) + lt * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 v = *reinterpret_cast<const float4*>(row + ((j ^ (lt & 7)) << 4));
                            hi[4 * j] = v.x; hi[4 * j + 1] = v.y; hi[4 * j + 2] = v.z; hi[4 * j + 3] = v.w;
                        }
                    } else {
                        const float* col = reinterpret_cast<const float*>(sA(s) + (lt >> 5) * 4096) + (lt & 31);
#pragma unroll
                        for (int k = 0; k < 32; ++k) hi[k] = col[k * 32];
                    }
                    if (ACT) {                                             // swish(0) = 0: the zero padding stays zero
#pragma unroll
                        for (int k = 0; k < 32; ++k) hi[k] = cv_swish(hi[k]);
                    }
                    if (!(g.debug & 1)) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) lo[k] = lo_tf32(hi[k]);
                    }
                    cv_tmem_st32(trow + l * 64, hi);
                    cv_tmem_st32(trow + l * 64 + 32, lo);
                }
                if (lt == 0) CV_STAMP(8, it);
                // B: lo image, shared -> shared (same swizzled offsets)
                const float4* __restrict__ src = reinterpret_cast<const float4*>(sB(s));
                float4* __restrict__ dst = reinterpret_cast<float4*>(sLo(l));
                int i = ((g.debug & 1) || g.blo) ? n16 : lt;               // blo: TMA delivered the lo image already
                for (; i + 384 < n16; i += 512) {
                    const float4 v0 = src[i], v1 = src[i + 128], v2 = src[i + 256], v3 = src[i + 384];
                    dst[i] = lo4_tf32(v0);
                    dst[i + 128] = lo4_tf32(v1);
                    dst[i + 256] = lo4_tf32(v2);
                    dst[i + 384] = lo4_tf32(v3);
                }
                for (; i < n16; i += 128) dst[i] = lo4_tf32(src[i]);
                if (lt == 0) CV_STAMP(9, it);
                tmem_st_wait();
                if (lt == 0) CV_STAMP(10, it);
                fence_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lt == 0) CV_STAMP(11, it);
                if (lane == 0) cv_arrive_at(&bars.ready[l], 0, rank);
                if (lt == 0) CV_STAMP(2, it);
            }
        }
    } else if (warp == 12) {
        if (lane == 0) {
            // =================================================================
            // TMA producer: this CTA's 128 rows of A and its half of B
            // =================================================================
            CvWalk walk(pair, g);
            CvSeg sg;
            uint32_t it = 0;
            const int nbp0 = g.np0 / 2, nbp1 = g.np1 / 2;
            while (walk.next(sg)) {
                for (int kc = sg.k0; kc < sg.k1; ++kc, ++it) {
                    const int s = it % g.stages;
                    if (it >= (uint32_t)g.stages) mbar_wait(&bars.empty[s], ((it / g.stages) - 1) & 1);
                    // precomputed lo image of B: it lands in lo slot l with the stage's barrier, so the slot must have
                    // been released by the MMAs of its previous use (the lo warps wait for the same phase before they
                    // overwrite the slot's TMEM half)
                    const int l = it % g.slots;
                    const bool blo = MODE == 0 && g.blo;
                    if (blo && it >= (uint32_t)g.slots) mbar_wait(&bars.lo_free[l], ((it / g.slots) - 1) & 1);
                    CV_STAMP(0, it);
                    if (MODE == 0) {
                        const int tap = kc / g.cch, cc = kc - tap * g.cch;
                        const int ty = tap / g.kw, tx = tap - ty * g.kw;
                        const int img = sg.tile / g.tiles_per_image;
                        const int y0 = (sg.tile - img * g.tiles_per_image) * 2 * g.rows_cta + (int)rank * g.rows_cta;
                        const bool la = !(g.debug & 16), lb = !(g.debug & 32);
                        mbar_expect_tx(&bars.full[s], (la ? kCvABytes : 0) + (lb ? (blo ? 2 * b_bytes : b_bytes) : 0));
                        if (la) cv_tma_4d(sA(s), &tmA, cc * 32, tx - g.pl, y0 + ty - g.pt, img, &bars.full[s]);
                        if (lb) tma_load_3d(sB(s), &tmB, cc * 32, tap, (int)rank * nbp0, &bars.full[s]);
                        if (lb && nbp1 > 0)
                            tma_load_3d(sB(s) + nbp0 * 128, &tmB1, cc * 32, tap, g.np0 + (int)rank * nbp1, &bars.full[s]);
                        if (lb && blo) {
                            tma_load_3d(sLo(l), &tmL, cc * 32, tap, (int)rank * nbp0, &bars.full[s]);
                            if (nbp1 > 0)
                                tma_load_3d(sLo(l) + nbp0 * 128, &tmL1, cc * 32, tap, g.np0 + (int)rank * nbp1, &bars.full[s]);
                        }
                    } else {
                        const int img = kc / g.chunks_per_image;
                        const int y0 = (kc - img * g.chunks_per_image) * g.chunk_rows;
                        int nblk = g.mblocks - (sg.tile * 8 + (int)rank * 4);
                        nblk = nblk < 0 ? 0 : (nblk > 4 ? 4 : nblk);
                        mbar_expect_tx(&bars.full[s], nblk * 4096 + b_bytes);
                        for (int i = 0; i < nblk; ++i) {
                            const int mb = sg.tile * 8 + (int)rank * 4 + i;
                            const int tap = mb / g.cch, cc = mb - tap * g.cch;
                            const int ty = tap / g.kw, tx = tap - ty * g.kw;
                            cv_tma_4d(sA(s) + i * 4096, &tmA, cc * 32, tx - g.pl, y0 + ty - g.pt, img, &bars.full[s]);
                        }
                        for (int i = 0; i < nbp0 / 32; ++i)
                            cv_tma_4d(sB(s) + i * 4096, &tmB, (int)rank * nbp0 + i * 32, 0, y0, img, &bars.full[s]);
                        for (int i = 0; i < nbp1 / 32; ++i)
                            cv_tma_4d(sB(s) + (nbp0 / 32 + i) * 4096, &tmB, g.np0 + (int)rank * nbp1 + i * 32, 0, y0, img,
                                      &bars.full[s]);
                    }
                }
            }
        }
        __syncwarp();
    } else {
        if (rank == 0 && lane == 0) {
            // =================================================================
            // MMA issuer (leader CTA, one thread): operands of BOTH CTAs, accumulators in both TMEMs
            // =================================================================
            const bool mn = MODE == 1;
            const uint32_t idesc0 = idesc_tf32(256, g.np0, false, mn);
            const uint32_t idesc1 = idesc_tf32(256, g.np1 > 0 ? g.np1 : 32, false, mn);
            const uint32_t kstep = mn ? 64u : 2u;                     // 1024 / 32 bytes >> 4
            const uint32_t b1_off = (uint32_t)(g.np0 / 2 * 128) >> 4;
            // ring positions are carried, not recomputed: between the last MMA of a chunk and the first of the next the
            // tensor pipe lives off its queue, and ONE thread issues everything in between
            CvWalk walk(pair, g);
            CvSeg sg;
            uint32_t it = 0, si = 0;
            int s = 0, l = 0;
            uint32_t l_par = 0;                                      // parity of ready[l]'s current use
            uint32_t bh_s[kCvMaxStages], bl_l[kCvMaxSlots];
#pragma unroll
            for (int q = 0; q < kCvMaxStages; ++q)
                bh_s[q] = mn ? mnmajor_lo(smem_u32(sB(q)), 4096) : kmajor_lo(smem_u32(sB(q)));
#pragma unroll
            for (int q = 0; q < kCvMaxSlots; ++q)
                bl_l[q] = mn ? mnmajor_lo(smem_u32(sLo(q)), 4096) : kmajor_lo(smem_u32(sLo(q)));
            const bool one = g.debug & 2;
            const bool two = g.np1 > 0;
            while (walk.next(sg)) {
                if (si > 0) mbar_wait(&bars.acc_free, (si - 1) & 1);
                tc_fence_after();
                int cc = MODE == 0 ? sg.k0 % g.cch : 0;
                for (int kc = sg.k0; kc < sg.k1; ++kc, ++it) {
                    mbar_wait(&bars.ready[l], l_par);
                    tc_fence_after();
                    CV_STAMP(3, it);
                    int ksteps = 4;
                    if (MODE == 0) {
                        if (cc == g.cch - 1) { ksteps = g.last_ksteps; cc = 0; }
                        else ++cc;
                    }
                    uint32_t bh = 0, bl = 0;
#pragma unroll
                    for (int q = 0; q < kCvMaxStages; ++q) bh = s == q ? bh_s[q] : bh;
#pragma unroll
                    for (int q = 0; q < kCvMaxSlots; ++q) bl = l == q ? bl_l[q] : bl;
                    const uint32_t a_hi = tmem + acol + l * 64, a_lo = a_hi + 32;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t o = ks * kstep, ak = ks * 8;
                        const bool acc = kc != sg.k0 || ks != 0;
                        const uint64_t dh0 = mn ? desc_mn(bh + o) : desc_k(bh + o);
                        const uint64_t dl0 = mn ? desc_mn(bl + o) : desc_k(bl + o);
                        cv_mma2_ts(tmem, a_lo + ak, dh0, idesc0, acc);
                        if (!one) {
                            cv_mma2_ts(tmem, a_hi + ak, dl0, idesc0, true);
                            cv_mma2_ts(tmem, a_hi + ak, dh0, idesc0, true);
                        }
                        if (two) {
                            const uint64_t dh1 = mn ? desc_mn(bh + b1_off + o) : desc_k(bh + b1_off + o);
                            const uint64_t dl1 = mn ? desc_mn(bl + b1_off + o) : desc_k(bl + b1_off + o);
                            cv_mma2_ts(tmem + g.np0, a_lo + ak, dh1, idesc1, acc);
                            if (!one) {
                                cv_mma2_ts(tmem + g.np0, a_hi + ak, dl1, idesc1, true);
                                cv_mma2_ts(tmem + g.np0, a_hi + ak, dh1, idesc1, true);
                            }
                        }
                    }
                    cv_commit2(&bars.empty[s]);
                    cv_commit2(&bars.lo_free[l]);
                    CV_STAMP(4, it);
                    if (++s == g.stages) s = 0;
                    if (++l == g.slots) { l = 0; l_par ^= 1; }
                }
                cv_commit2(&bars.acc_full);
                ++si;
            }
            tc_fence_before();
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    cv_cluster_sync();                                     // the peer may still read its TMEM half / our smem
    if (warp == 13) {
        tc_fence_after();
        cv_tmem_dealloc2(tmem, 512);
    }
}

// wt[ci][taps - 1 - tap][co] = w[co][tap][ci]: the input gradient of a stride-1 "same" convolution is the
// forward convolution of gy with these weights
__global__ void conv_wt_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int taps, int Cin) {
    __shared__ float tile[32][33];
    const int tap = blockIdx.z;
    const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int co = co0 + i, ci = ci0 + threadIdx.x;
        tile[i][threadIdx.x] = (co < Cout && ci < Cin) ? w[((int64_t)co * taps + tap) * Cin + ci] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int ci = ci0 + i, co = co0 + threadIdx.x;
        if (ci < Cin && co < Cout) wt[((int64_t)ci * taps + (taps - 1 - tap)) * Cout + co] = tile[threadIdx.x][i];
    }
}

// The same for every convolution of a model in ONE launch (the trainer calls it once per step, right after the
// weight norm): kernels live at offs[i] in a flat array, their transposes at the same offsets of a second one;
// tile_prefix[i] = number of 32 x 32 tiles of the kernels before i (tile_prefix[count] = grid size).
__global__ void conv_wt_batched_kernel(const float* __restrict__ w_flat, float* __restrict__ wt_flat,
                                       const int64_t* __restrict__ offs, const int* __restrict__ couts,
                                       const int* __restrict__ taps_, const int* __restrict__ cins,
                                       const int64_t* __restrict__ tile_prefix, int count) {
    __shared__ float tile[32][33];
    const int64_t b = blockIdx.x;
    int lo = 0, hi = count - 1;
    while (lo < hi) {                                      // last i with tile_prefix[i] <= b
        const int mid = (lo + hi + 1) >> 1;
        if (tile_prefix[mid] <= b) lo = mid; else hi = mid - 1;
    }
    const int Cout = couts[lo], taps = taps_[lo], Cin = cins[lo];
    const int tci = (Cin + 31) / 32, tco = (Cout + 31) / 32;
    int64_t local = b - tile_prefix[lo];
    const int tap = (int)(local / (tci * tco));
    local -= (int64_t)tap * tci * tco;
    const int co0 = (int)(local / tci) * 32, ci0 = (int)(local % tci) * 32;
    const float* w = w_flat + offs[lo];
    float* wt = wt_flat + offs[lo];
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int co = co0 + i, ci = ci0 + threadIdx.x;
        tile[i][threadIdx.x] = (co < Cout && ci < Cin) ? w[((int64_t)co * taps + tap) * Cin + ci] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int ci = ci0 + i, co = co0 + threadIdx.x;
        if (ci < Cin && co < Cout) wt[((int64_t)ci * taps + (taps - 1 - tap)) * Cout + co] = tile[threadIdx.x][i];
    }
}

// dst = lo part of the 3xTF32 split of src (tc_common.cuh lo_tf32: v = trunc_tf32(v) + lo, lo rounded to tf32) -- exactly
// what the lo warps compute from a landed B tile; a trainer runs it once per step over all its kernels and their
// transposes and hands the image to davi_conv2d_fwd_pre (bit-identical results, no shared -> shared pass per chunk).
__global__ void __launch_bounds__(256) tf32_lo_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
    const int64_t n4 = n >> 2, stride = (int64_t)gridDim.x * blockDim.x;
    const float4* __restrict__ s4 = reinterpret_cast<const float4*>(src);
    float4* __restrict__ d4 = reinterpret_cast<float4*>(dst);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        const float4 v0 = __ldcs(s4 + i), v1 = __ldcs(s4 + i + stride), v2 = __ldcs(s4 + i + 2 * stride), v3 = __ldcs(s4 + i + 3 * stride);
        d4[i] = lo4_tf32(v0);
        d4[i + stride] = lo4_tf32(v1);
        d4[i + 2 * stride] = lo4_tf32(v2);
        d4[i + 3 * stride] = lo4_tf32(v3);
    }
    for (; i < n4; i += stride) d4[i] = lo4_tf32(__ldcs(s4 + i));
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[(n4 << 2) + threadIdx.x] = lo_tf32(src[(n4 << 2) + threadIdx.x]);
}

// Weight gradients with few tiles and a long pixel axis (1x1 kernels: 2 tiles x 320 chunks) are cut into dozens of
// segments per tile; one owner adding them serially would take longer than the MMAs.  In `reduce` mode every
// segment lands in the workspace and this kernel sums the segments of a tile in pair order (deterministic).
// grid = (tiles * 2 ranks, float4 column groups); block = 128 threads <-> rows; one float4 per thread and segment.
template <int MODE>
__global__ void __launch_bounds__(128) conv_reduce_kernel(const CvArgs a, const CvGeom g) {
    const int tile = blockIdx.x >> 1, rank = blockIdx.x & 1, c4 = blockIdx.y, r = threadIdx.x;
    const int64_t units = (int64_t)g.tiles * g.kch;
    const int ncol4 = (g.n_total + 31) / 32 * 8;
    const int p0 = cv_pair_of((int64_t)tile * g.kch, g.pairs, units);
    const int p1 = cv_pair_of((int64_t)(tile + 1) * g.kch - 1, g.pairs, units);
    // segment 0 of the first pair lies in this tile iff that pair starts inside it; otherwise this tile is its segment 1
    const int seg0 = (int)(cv_first_unit(p0, g.pairs, units) / g.kch) == tile ? 0 : 1;
    const size_t slot = (size_t)2 * ncol4 * 128;                                  // float4s per (pair, segment)
    const float4* part = reinterpret_cast<const float4*>(a.partials) + (size_t)rank * ncol4 * 128 + (size_t)c4 * 128 + r;
    float4 v = __ldcg(part + (size_t)(2 * p0 + seg0) * slot);
    int p = p0 + 1;                                                // every later pair STARTS inside this tile: its segment 0
    for (; p + 3 <= p1; p += 4) {
        const float4 w0 = __ldcg(part + (size_t)(2 * p) * slot), w1 = __ldcg(part + (size_t)(2 * p + 2) * slot);
        const float4 w2 = __ldcg(part + (size_t)(2 * p + 4) * slot), w3 = __ldcg(part + (size_t)(2 * p + 6) * slot);
        v.x += w0.x; v.y += w0.y; v.z += w0.z; v.w += w0.w;
        v.x += w1.x; v.y += w1.y; v.z += w1.z; v.w += w1.w;
        v.x += w2.x; v.y += w2.y; v.z += w2.z; v.w += w2.w;
        v.x += w3.x; v.y += w3.y; v.z += w3.z; v.w += w3.w;
    }
    for (; p <= p1; ++p) {
        const float4 w = __ldcg(part + (size_t)(2 * p) * slot);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    const int c = c4 * 4;
    if (c >= g.Cout) return;
    if (MODE == 0) {
        const int64_t out_row = ((int64_t)tile * 256 + rank * 128 + r) * g.out_sp;
        if (a.bias) {
            const float4 bb = ld4(a.bias + c);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        *reinterpret_cast<float4*>(a.out + out_row + c) = v;
    } else {
        const int mb = tile * 8 + rank * 4 + (r >> 5);
        const int tap = mb / g.cch, cc = mb - tap * g.cch;
        const int ci = cc * 32 + (r & 31);
        if (mb < g.mblocks && ci < g.Cin) {
            float* o = a.out + (int64_t)c * g.taps * g.Cin + (int64_t)tap * g.Cin + ci;
            const int64_t cs = (int64_t)g.taps * g.Cin;
            o[0] = v.x; o[cs] = v.y; o[2 * cs] = v.z; o[3 * cs] = v.w;
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------------

static int cv_pairs_for(int tiles, int kch, int sms) {
    // DAVI_OPT_CONV_SM_RESERVE: SMs left free for kernels that run beside the convolutions (the NCCL kernels of the
    // overlapped gradient exchange): a CTA pair that finds no free TPC starts only when another pair has finished
    int reserve = get_option(DAVI_OPT_CONV_SM_RESERVE);
    if (reserve < 0) reserve = 0;
    if (reserve > sms - 8) reserve = sms - 8;
    int pmax = (sms - reserve) / 2;
    if (pmax > kCvMaxPairs) pmax = kCvMaxPairs;
    if (pmax < 1) pmax = 1;
    const int64_t units = (int64_t)tiles * kch;
    // short K loops: whole tiles per pair (no partial accumulators) unless that leaves most pairs idle
    if (kch < 16 && tiles >= pmax / 2) return tiles < pmax ? tiles : pmax;
    int64_t p = units / 4;                                 // at least 4 chunks per pair
    if (p < 1) p = 1;
    return (int)(p < pmax ? p : pmax);
}

static bool cv_common(CvGeom& g, int B, int H, int W, int Cin, int Cout, int kh, int kw) {
    if (B <= 0 || (W != 8 && W != 16 && W != 32) || H <= 0 || (H * W) % 256 || Cin % 4 || Cout % 4 || Cin < 4 ||
        Cout < 4 || Cout > 384 || Cin > 384 || kh < 1 || kw < 1 || !(kh & 1) || !(kw & 1) || kh > 7 || kw > 7)
        return false;
    g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.kh = kh; g.kw = kw;
    g.pt = (kh - 1) / 2; g.pl = (kw - 1) / 2;
    g.taps = kh * kw;
    g.cch = (Cin + 31) / 32;
    g.last_ksteps = (Cin - 32 * (g.cch - 1) + 7) / 8;
    g.rows_cta = 128 / W;
    g.tiles_per_image = H * W / 256;
    g.chunk_rows = 32 / W;
    g.chunks_per_image = H * W / 32;
    g.mblocks = g.taps * g.cch;
    return true;
}

static bool cv_finish(CvGeom& g, int sms) {
    g.nb = g.n_total / 2;
    const int b_bytes = g.nb * 128, half = kCvABytes + b_bytes;
    int slots = (512 - g.n_total) / 64;                     // TMEM columns left of the accumulator: 64 per A slot
    if (slots < 2) return false;
    if (slots > 3) slots = 3;
    int stages = (227 * 1024 - 1024 - kCvStageBytes - slots * b_bytes) / half;
    if (stages > kCvMaxStages) stages = kCvMaxStages;
    if (stages < 2) return false;
    g.stages = stages;
    g.slots = slots;
    g.tile_ok = (size_t)128 * (((g.Cout + 3) / 4 | 1) * 4) * 4 <= (size_t)stages * half + (size_t)slots * b_bytes ? 1 : 0;
    g.pairs = cv_pairs_for(g.tiles, g.kch, sms);
    g.debug = get_option(DAVI_OPT_CONV_DEBUG);
    // more than a handful of segments per tile: sum them with a second, parallel kernel.  Needs every pair to hold
    // at most two segments (its run of units is no longer than a tile)
    const int64_t units = (int64_t)g.tiles * g.kch;
    const int64_t per_pair = (units + g.pairs - 1) / g.pairs;
    g.reduce = (per_pair * 6 < g.kch && per_pair <= g.kch) ? 1 : 0;
    return true;
}

bool conv_tc_supported(int B, int H, int W, int Cin, int Cout, int kh, int kw) {
    CvGeom g{};
    return cv_common(g, B, H, W, Cin, Cout, kh, kw);
}

static int cv_sms() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return sms;
}

size_t conv_tc_workspace_bytes() {
    return (size_t)kCvFlagBytes + (size_t)kCvMaxPairs * 2 * 2 * 128 * 512 * 4;
}

template <int MODE>
static int cv_launch(const char* fn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmB1,
                     const CUtensorMap& tmL, const CUtensorMap& tmL1, CvArgs a, const CvGeom& g,
                     void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const size_t n_pad = (size_t)(g.n_total + 31) / 32 * 32;
    const size_t need = (size_t)kCvFlagBytes + (size_t)g.pairs * (g.reduce ? 2 : 1) * 2 * 128 * n_pad * 4;
    if (!workspace || workspace_bytes < need || !aligned16(workspace)) {
        set_error("%s: workspace of %zu bytes needed (16-byte aligned, zero-filled once), got %zu", fn, need, workspace_bytes);
        return DAVI_EWORKSPACE;
    }
    a.flags = reinterpret_cast<uint32_t*>(workspace);
    a.partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + kCvFlagBytes);
    const size_t smem = 1024 + (size_t)g.stages * (kCvABytes + g.nb * 128) + (size_t)g.slots * g.nb * 128 + kCvStageBytes;
    if (a.act) {
        DAVI_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<MODE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_gemm_kernel<MODE, 1><<<2 * g.pairs, kCvThreads, smem, st>>>(tmA, tmB, tmB1, tmL, tmL1, a, g);
    } else {
        DAVI_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<MODE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_gemm_kernel<MODE, 0><<<2 * g.pairs, kCvThreads, smem, st>>>(tmA, tmB, tmB1, tmL, tmL1, a, g);
    }
    int rc = check_launch(fn);
    if (rc || !g.reduce) return rc;
    conv_reduce_kernel<MODE><<<dim3(2 * g.tiles, (g.n_total + 31) / 32 * 8), 128, 0, st>>>(a, g);
    return check_launch(fn);
}

// geometry of a forward / input-gradient launch (mode 0) or a weight-gradient launch (mode 1) on `sms` SMs
static bool cv_geom(CvGeom& g, int mode, int B, int H, int W, int Cin, int Cout, int kh, int kw, int sms) {
    if (!cv_common(g, B, H, W, Cin, Cout, kh, kw) || sms < 2) return false;
    g.mode = mode;
    if (mode == 0) {
        g.tiles = B * g.tiles_per_image;
        g.kch = g.taps * g.cch;
        // the TMEM-operand pair MMA takes N in multiples of 32
        g.n_total = (Cout + 31) / 32 * 32;
        if (g.n_total <= 256) { g.np0 = g.n_total; g.np1 = 0; }
        else if (get_option(DAVI_OPT_CONV_DEBUG) & 8) { g.np0 = (g.n_total / 2 + 31) / 32 * 32; g.np1 = g.n_total - g.np0; }
        else { g.np0 = 256; g.np1 = g.n_total - 256; }     // one full-width MMA + the rest: measured 3 % faster than two halves
    } else {
        g.tiles = (g.mblocks + 7) / 8;
        g.kch = B * g.chunks_per_image;
        g.n_total = (Cout + 63) / 64 * 64;                 // MN-major B: whole 32-column blocks per CTA
        g.np0 = g.n_total < 256 ? g.n_total : 256;
        g.np1 = g.n_total - g.np0;
    }
    return cv_finish(g, sms);
}

int conv_fwd_tc_launch(const float* x, const float* w, const float* w_lo, const float* bias, float* y, int B, int H,
                       int W, int Cin, int Cout, int kh, int kw, int64_t x_sp, int64_t y_sp, int act, void* workspace,
                       size_t workspace_bytes, cudaStream_t st) {
    CvGeom g{};
    if (!cv_geom(g, 0, B, H, W, Cin, Cout, kh, kw, cv_sms())) {
        set_error("davi_conv2d_fwd: shape not supported by the tensor-core kernel");
        return DAVI_EINVAL;
    }
    g.out_sp = y_sp;
    // the lo image travels through the lo ring, whose depth is the prefetch distance of BOTH weight streams then: with
    // two slots (more than 320 accumulator columns: the packed q|k|v projections) the on-chip pass is faster (measured
    // 29.1 vs 31.6 us at 276 -> 340 channels, 1x1), with three the loaded image is (3x3, 316 channels: 89.5 vs 91.4 us)
    g.blo = (w_lo && g.slots >= 3) ? 1 : 0;
    CUtensorMap tmA, tmB, tmB1, tmL, tmL1;
    int rc;
    {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t str[3] = {(uint64_t)x_sp * 4, (uint64_t)W * x_sp * 4, (uint64_t)H * W * x_sp * 4};
        const uint32_t box[4] = {32, (uint32_t)W, (uint32_t)g.rows_cta, 1};
        if ((rc = make_tmap_nd(&tmA, x, 4, dims, str, box, 0))) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)g.taps, (uint64_t)Cout};
        const uint64_t str[2] = {(uint64_t)Cin * 4, (uint64_t)g.taps * Cin * 4};
        const uint32_t box[3] = {32, 1, (uint32_t)(g.np0 / 2)};
        const uint32_t box1[3] = {32, 1, (uint32_t)(g.np1 > 0 ? g.np1 / 2 : g.np0 / 2)};
        if ((rc = make_tmap_nd(&tmB, w, 3, dims, str, box, 0))) return rc;
        if ((rc = make_tmap_nd(&tmB1, w, 3, dims, str, box1, 0))) return rc;
        tmL = tmB;
        tmL1 = tmB1;
        if (w_lo) {                                            // same boxes over the precomputed lo image
            if ((rc = make_tmap_nd(&tmL, w_lo, 3, dims, str, box, 0))) return rc;
            if ((rc = make_tmap_nd(&tmL1, w_lo, 3, dims, str, box1, 0))) return rc;
        }
    }
    CvArgs a{};
    a.bias = bias; a.out = y;
    a.act = act;
    return cv_launch<0>("davi_conv2d_fwd", tmA, tmB, tmB1, tmL, tmL1, a, g, workspace, workspace_bytes, st);
}

int conv_wgrad_tc_launch(const float* x, const float* gy, float* gw, int B, int H, int W, int Cin, int Cout, int kh,
                         int kw, int64_t x_sp, int64_t gy_sp, int act, void* workspace, size_t workspace_bytes,
                         cudaStream_t st) {
    CvGeom g{};
    if (!cv_geom(g, 1, B, H, W, Cin, Cout, kh, kw, cv_sms())) {
        set_error("davi_conv2d_wgrad: shape not supported by the tensor-core kernel");
        return DAVI_EINVAL;
    }
    CUtensorMap tmA, tmB;
    int rc;
    const uint32_t box[4] = {32, (uint32_t)W, (uint32_t)g.chunk_rows, 1};
    {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t str[3] = {(uint64_t)x_sp * 4, (uint64_t)W * x_sp * 4, (uint64_t)H * W * x_sp * 4};
        if ((rc = make_tmap_nd(&tmA, x, 4, dims, str, box, 2))) return rc;     // plain rows: read by threads, not by the MMA
    }
    {
        const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t str[3] = {(uint64_t)gy_sp * 4, (uint64_t)W * gy_sp * 4, (uint64_t)H * W * gy_sp * 4};
        if ((rc = make_tmap_nd(&tmB, gy, 4, dims, str, box, 1))) return rc;
    }
    CvArgs a{};
    a.out = gw;
    a.act = act;
    return cv_launch<1>("davi_conv2d_wgrad", tmA, tmB, tmB, tmB, tmB, a, g, workspace, workspace_bytes, st);
}

// plan[0..11] = tiles, k-chunks per tile, CTA pairs, accumulator columns, columns of MMA 0, of MMA 1, hi-ring stages,
// lo-ring slots, reduce mode, dynamic shared memory bytes, workspace bytes needed, staged-tile epilogue available
int conv_plan(int mode, int B, int H, int W, int Cin, int Cout, int kh, int kw, int sms, int64_t* plan) {
    CvGeom g{};
    if (!cv_geom(g, mode, B, H, W, Cin, Cout, kh, kw, sms)) return DAVI_EINVAL;
    const int64_t n_pad = (g.n_total + 31) / 32 * 32;
    plan[0] = g.tiles; plan[1] = g.kch; plan[2] = g.pairs; plan[3] = g.n_total; plan[4] = g.np0; plan[5] = g.np1;
    plan[6] = g.stages; plan[7] = g.slots; plan[8] = g.reduce;
    plan[9] = 1024 + (int64_t)g.stages * (kCvABytes + g.nb * 128) + (int64_t)g.slots * g.nb * 128 + kCvStageBytes;
    plan[10] = kCvFlagBytes + (int64_t)g.pairs * (g.reduce ? 2 : 1) * 2 * 128 * n_pad * 4;
    plan[11] = g.tile_ok;
    return DAVI_OK;
}

int conv_wt_launch(const float* w, float* wt, int Cout, int taps, int Cin, cudaStream_t st) {
    dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, taps);
    conv_wt_kernel<<<grid, dim3(32, 8), 0, st>>>(w, wt, Cout, taps, Cin);
    return check_launch("davi_conv2d_wt");
}

}  // namespace davi

using namespace davi;

#ifdef DAVI_CV_TRACE
extern "C" int davi_cv_trace_read(unsigned long long* out) {
    return cudaMemcpyFromSymbol(out, g_cv_trace, sizeof(g_cv_trace)) == cudaSuccess ? 0 : DAVI_ECUDA;
}
#endif

extern "C" size_t davi_conv2d_workspace_bytes(void) { return conv_tc_workspace_bytes(); }

extern "C" int davi_conv2d_wt_batched(const float* w_flat, float* wt_flat, const int64_t* offs, const int* couts,
                                      const int* taps, const int* cins, const int64_t* tile_prefix, int count,
                                      int64_t total_tiles, davi_stream_t stream) {
    DAVI_REQUIRE(w_flat && wt_flat && w_flat != wt_flat && offs && couts && taps && cins && tile_prefix, "null / aliased pointer");
    DAVI_REQUIRE(count > 0 && total_tiles > 0 && total_tiles < (int64_t)1 << 31, "table size");
    conv_wt_batched_kernel<<<(unsigned)total_tiles, dim3(32, 8), 0, (cudaStream_t)stream>>>(w_flat, wt_flat, offs, couts, taps,
                                                                                          cins, tile_prefix, count);
    return check_launch("davi_conv2d_wt_batched");
}

extern "C" int davi_fill_zero(float* dst, int64_t n, davi_stream_t stream) {
    DAVI_REQUIRE(dst && n >= 0, "null pointer / negative size");
    DAVI_CUDA(cudaMemsetAsync(dst, 0, (size_t)n * sizeof(float), (cudaStream_t)stream));
    return DAVI_OK;
}

extern "C" int davi_conv2d_plan(int mode, int B, int H, int W, int Cin, int Cout, int kh, int kw, int sms, int64_t* plan) {
    DAVI_REQUIRE(plan && (mode == 0 || mode == 1) && sms >= 2, "arguments");
    const int rc = conv_plan(mode, B, H, W, Cin, Cout, kh, kw, sms, plan);
    if (rc) set_error("davi_conv2d_plan: shape outside the envelope of the tensor-core kernels");
    return rc;
}

extern "C" int davi_conv2d_supported(int B, int H, int W, int Cin, int Cout, int kh, int kw) {
    return conv_tc_supported(B, H, W, Cin, Cout, kh, kw) ? 1 : 0;
}

extern "C" int davi_conv2d_fwd_ex(const float* x, const float* w, const float* bias, float* y, int B, int H, int W,
                                  int Cin, int Cout, int kh, int kw, int64_t x_sp, int64_t y_sp, int act, void* workspace,
                                  size_t workspace_bytes, davi_stream_t stream) {
    DAVI_REQUIRE(x && w && y, "null pointer");
    DAVI_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && (!bias || aligned16(bias)), "alignment");
    DAVI_REQUIRE(mult4(x_sp) && mult4(y_sp) && x_sp >= Cin && y_sp >= Cout, "pixel strides");
    DAVI_REQUIRE(act == 0 || act == 1, "act");
    return conv_fwd_tc_launch(x, w, nullptr, bias, y, B, H, W, Cin, Cout, kh, kw, x_sp, y_sp, act, workspace, workspace_bytes,
                              (cudaStream_t)stream);
}

extern "C" int davi_conv2d_fwd_pre(const float* x, const float* w, const float* w_lo, const float* bias, float* y, int B,
                                   int H, int W, int Cin, int Cout, int kh, int kw, int64_t x_sp, int64_t y_sp, int act,
                                   void* workspace, size_t workspace_bytes, davi_stream_t stream) {
    DAVI_REQUIRE(x && w && w_lo && y && w_lo != w, "null / aliased pointer");
    DAVI_REQUIRE(aligned16(x) && aligned16(w) && aligned16(w_lo) && aligned16(y) && (!bias || aligned16(bias)), "alignment");
    DAVI_REQUIRE(mult4(x_sp) && mult4(y_sp) && x_sp >= Cin && y_sp >= Cout, "pixel strides");
    DAVI_REQUIRE(act == 0 || act == 1, "act");
    return conv_fwd_tc_launch(x, w, w_lo, bias, y, B, H, W, Cin, Cout, kh, kw, x_sp, y_sp, act, workspace, workspace_bytes,
                              (cudaStream_t)stream);
}

extern "C" int davi_tf32_lo(const float* src, float* dst, int64_t n, davi_stream_t stream) {
    DAVI_REQUIRE(src && dst && src != dst && n >= 0, "null / aliased pointer, negative size");
    DAVI_REQUIRE(aligned16(src) && aligned16(dst), "alignment");
    if (n == 0) return DAVI_OK;
    int64_t blocks = (n / 4 + 4 * 256 - 1) / (4 * 256);
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    tf32_lo_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    return check_launch("davi_tf32_lo");
}

extern "C" int davi_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W,
                               int Cin, int Cout, int kh, int kw, int64_t x_sp, int64_t y_sp, void* workspace,
                               size_t workspace_bytes, davi_stream_t stream) {
    return davi_conv2d_fwd_ex(x, w, bias, y, B, H, W, Cin, Cout, kh, kw, x_sp, y_sp, 0, workspace, workspace_bytes, stream);
}

extern "C" int davi_conv2d_wgrad_ex(const float* x, const float* gy, float* gw, int B, int H, int W, int Cin, int Cout,
                                    int kh, int kw, int64_t x_sp, int64_t gy_sp, int act, void* workspace,
                                    size_t workspace_bytes, davi_stream_t stream) {
    DAVI_REQUIRE(x && gy && gw, "null pointer");
    DAVI_REQUIRE(aligned16(x) && aligned16(gy) && aligned16(gw), "alignment");
    DAVI_REQUIRE(mult4(x_sp) && mult4(gy_sp) && x_sp >= Cin && gy_sp >= Cout, "pixel strides");
    DAVI_REQUIRE(act == 0 || act == 1, "act");
    return conv_wgrad_tc_launch(x, gy, gw, B, H, W, Cin, Cout, kh, kw, x_sp, gy_sp, act, workspace, workspace_bytes,
                                (cudaStream_t)stream);
}

extern "C" int davi_conv2d_wgrad(const float* x, const float* gy, float* gw, int B, int H, int W, int Cin, int Cout,
                                 int kh, int kw, int64_t x_sp, int64_t gy_sp, void* workspace, size_t workspace_bytes,
                                 davi_stream_t stream) {
    return davi_conv2d_wgrad_ex(x, gy, gw, B, H, W, Cin, Cout, kh, kw, x_sp, gy_sp, 0, workspace, workspace_bytes, stream);
}

extern "C" int davi_conv2d_wt(const float* w, float* wt, int Cout, int taps, int Cin, davi_stream_t stream) {
    DAVI_REQUIRE(w && wt && w != wt, "null / aliased pointer");
    DAVI_REQUIRE(Cout > 0 && taps > 0 && Cin > 0 && taps <= 65535, "shape");
    return conv_wt_launch(w, wt, Cout, taps, Cin, (cudaStream_t)stream);
}
