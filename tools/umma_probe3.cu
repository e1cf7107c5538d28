// Dev probe 3 (GPU box; NOT part of libdavi, not run by tests): the CTA-pair form of the tensor-core
// MMA, tcgen05.mma.cta_group::2 (M = 256 over two SMs of a TPC), for tf32 operands.
//   P1  correctness with exact small-integer matrices: D[256 x N] = A[256 x K] * B[N x K]^T, K-major
//       SWIZZLE_128B tiles filled by the threads; CTA r of the pair holds rows [128 r, 128 r + 128) of A
//       and rows [N/2 r, N/2 r + N/2) of B at the SAME shared-memory offsets, and lanes 0..127 of its own
//       TMEM hold rows [128 r, 128 r + 128) of D.
//   P2  issue cost: cycles per instruction for N = 64 / 128 / 256, K = 8 per instruction, against the
//       ~173 cycles of the single-CTA SS form (profiles/r01_umma_probe2.txt).  If the pair form costs the
//       same per instruction, every MMA-issue-bound kernel of this repo (nonlocal attention at N = 256,
//       a future implicit-GEMM convolution) doubles its tensor throughput by pairing CTAs.
// Why a probe: the pair form needs a cluster launch, TMEM allocated with cta_group::2 by a warp of BOTH
// CTAs, the MMA issued by the leader only, and a multicast commit -- none of which the round-1 kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I deep_attentive_vi_b200/csrc -I include \
//        tools/umma_probe3.cu -o tools/_bin/umma_probe3 && timeout 60 tools/_bin/umma_probe3
// Status: written at the end of round 1 without GPU time left -- compiles, never executed.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace davi::tc;

constexpr int kM = 256;          // rows of D over the pair
constexpr int kKs = 4;           // k-steps of 8 -> K = 32 = one 128-byte swizzle row
constexpr int kK = 8 * kKs;
constexpr int kThreads = 128;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// leader only: D (both CTAs' TMEM) (+)= A (both CTAs' smem halves) * B (both CTAs' smem halves)
__device__ __forceinline__ void mma2_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
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
// leader only: the barrier at the same offset in every CTA of `mask` arrives once the MMAs issued so far are done
__device__ __forceinline__ void tc_commit2(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                     "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

// K-major SWIZZLE_128B byte offset of element (r, c) of a [rows x 32] tf32 tile (tc_common.cuh)
__device__ __host__ inline uint32_t sw128(int r, int c) {
    return (uint32_t)((r / 8) * 1024 + (r % 8) * 128 + (((c / 4) ^ (r % 8)) * 16) + (c % 4) * 4);
}

struct Shared {
    uint64_t done;
    uint32_t tmem_base;
};

// a: [256 x K], b: [N x K] row-major in global memory; d: [256 x N]; cyc: cycles per instruction (rank 0)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pair_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, int N, int repeat,
            long long* __restrict__ cyc) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Shared sh;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                       // [128 x 32] tf32 = 16 KB
    uint8_t* sB = smem + 16384;               // [N/2 x 32] tf32 <= 16 KB
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const int half = N / 2;

    for (int i = tid; i < 128 * kK; i += kThreads) {
        const int r = i / kK, c = i % kK;
        *reinterpret_cast<float*>(sA + sw128(r, c)) = a[(size_t)(rank * 128 + r) * kK + c];
    }
    for (int i = tid; i < half * kK; i += kThreads) {
        const int r = i / kK, c = i % kK;
        *reinterpret_cast<float*>(sB + sw128(r, c)) = b[(size_t)(rank * half + r) * kK + c];
    }
    fence_async_smem();
    if (tid == 0) {
        mbar_init(&sh.done, 1);
        mbar_init_fence();
    }
    if (warp == 0) tmem_alloc2(&sh.tmem_base, 256);          // the same warp of BOTH CTAs
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                          // both halves of A / B written, both barriers live
    tc_fence_after();
    const uint32_t tmem = sh.tmem_base;

    if (rank == 0 && tid == 0) {
        const uint32_t idesc = idesc_tf32(kM, N, false, false);
        const long long t0 = clock64();
        for (int rep = 0; rep < repeat; ++rep)
            for (int ks = 0; ks < kKs; ++ks)
                mma2_tf32(tmem, desc_kmajor(smem_u32(sA) + ks * 32), desc_kmajor(smem_u32(sB) + ks * 32), idesc,
                          (rep | ks) != 0);
        tc_commit2(&sh.done, 0b11);
        mbar_wait(&sh.done, 0);
        const long long t1 = clock64();
        if (cyc) *cyc = (t1 - t0) / ((long long)repeat * kKs);
    }
    mbar_wait(&sh.done, 0);                                  // every thread of both CTAs
    tc_fence_after();
    // thread t <-> TMEM lane t <-> row rank * 128 + t of D
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(trow + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) d[(size_t)(rank * 128 + tid) * N + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                          // the peer may still be reading its TMEM half
    if (warp == 0) tmem_dealloc2(tmem, 256);
}

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e = (x);                                                          \
        if (e != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main() {
    CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024));
    for (int N : {64, 128, 256}) {
        std::vector<float> ha((size_t)kM * kK), hb((size_t)N * kK), hd((size_t)kM * N), want((size_t)kM * N);
        for (int i = 0; i < kM; ++i)
            for (int k = 0; k < kK; ++k) ha[(size_t)i * kK + k] = (float)((i * 3 + k * 5) % 7 - 3);
        for (int j = 0; j < N; ++j)
            for (int k = 0; k < kK; ++k) hb[(size_t)j * kK + k] = (float)((j * 2 + k * 3) % 5 - 2);
        for (int i = 0; i < kM; ++i)
            for (int j = 0; j < N; ++j) {
                float s = 0.f;
                for (int k = 0; k < kK; ++k) s += ha[(size_t)i * kK + k] * hb[(size_t)j * kK + k];
                want[(size_t)i * N + j] = s;
            }
        float *da, *db, *dd;
        long long* dc;
        CK(cudaMalloc(&da, ha.size() * 4));
        CK(cudaMalloc(&db, hb.size() * 4));
        CK(cudaMalloc(&dd, hd.size() * 4));
        CK(cudaMalloc(&dc, 8));
        CK(cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
        // P1: one pass over K
        pair_kernel<<<2, kThreads, 34 * 1024>>>(da, db, dd, N, 1, nullptr);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < hd.size(); ++i) bad += hd[i] != want[i];
        printf("P1 N=%3d  cta_group::2 M=256: %zu / %zu mismatches (D[0,0]=%g want %g; D[128,1]=%g want %g)\n", N, bad,
               hd.size(), hd[0], want[0], hd[(size_t)128 * N + 1], want[(size_t)128 * N + 1]);
        // P2: issue cost
        long long hc = 0;
        pair_kernel<<<2, kThreads, 34 * 1024>>>(da, db, dd, N, 256, dc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost));
        printf("P2 N=%3d  %lld cycles per cta_group::2 tf32 MMA (M=256, K=8): %.0f flop/cycle/SM\n", N, hc,
               2.0 * kM * N * 8 / (double)hc / 2.0);
        cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dc);
    }
    return 0;
}
