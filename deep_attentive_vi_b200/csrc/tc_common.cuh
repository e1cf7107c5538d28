// Blackwell (sm_100a) tensor-core plumbing shared by the nonlocal attention
// kernels: tcgen05.mma (kind::tf32) issue, TMEM alloc/load, mbarriers, UMMA
// shared-memory / instruction descriptors.  Inline PTX only -- no CUTLASS.
//
// Operand layouts used throughout (SWIZZLE_NONE "interleaved" canonical forms,
// 16-byte units = 4 tf32 values):
//   K-major  tile [R rows x Kd]  : byte(r, k)  = (k/4)*(R*16) + r*16 + (k%4)*4
//                                   -> LBO (next 4-k chunk) = R*16, SBO (next 8 rows) = 128
//   MN-major tile [Kd x Cn cols] : byte(k, c)  = (k/8)*(Cn*32) + (c/4)*128 + (k%8)*16 + (c%4)*4
//                                   -> LBO (next 8-k group) = Cn*32, SBO (next 4 cols) = 128
// One kind::tf32 instruction consumes K = 8 (two 16-byte chunks / one 8-k group).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace davi {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Waits for the phase with the given parity to complete.  A watchdog traps the
// kernel (launch failure, never a hang) if a barrier is not reached in ~2 s.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- tcgen05 ----------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// whole warp; writes the TMEM base address to *dst (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// mbarrier arrives once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem], fp32 accumulate, tf32 operands.  One thread issues.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// shared-memory matrix descriptor, SWIZZLE_NONE (layout_type 0), Blackwell version 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor for kind::tf32, fp32 accumulator
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                       // D format: f32
           | (2u << 7)                     // A format: tf32
           | (2u << 10)                    // B format: tf32
           | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16)
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// 32 lanes x 32-bit, 16 / 32 consecutive columns per thread (lane = 32*(warp%4) + laneid)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
          "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]),
          "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]),
          "=r"(u[30]), "=r"(u[31])
        : "r"(taddr)
        : "memory");
}

// ---- swizzled layouts, TMA, TMEM-resident A operands (validated by tools/umma_probe2.cu) --------
// K-major  SWIZZLE_128B        : tile [rows x 32 tf32]; byte(r,c) = (r/8)*1024 + (r%8)*128 + (((c/4)^(r%8))*16) + (c%4)*4
//                                descriptor layout 2, SBO = 1024, one k-step (8 tf32) = +32 bytes
// MN-major SWIZZLE_128B_BASE32B: blocks of 32 contiguous MN elements; block = k rows of 128 bytes,
//                                byte(k,c) = blk*(rows*128) + k*128 + (((c%32/8)^(k%4))*32) + (c%8)*4
//                                descriptor layout 1, LBO = rows*128 (next 32-element block), SBO = 512,
//                                one k-step (8 rows) = +1024 bytes.  (The only MN-major layout of 32-bit operands.)
// Both are exactly what a TMA box [32 floats x rows] writes with CU_TENSOR_MAP_SWIZZLE_128B /
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  The tensor core TRUNCATES fp32 bit patterns to tf32, so the raw
// tile is its own "hi" part and lo = rna_tf32(v - trunc(v)) is written by a separate elementwise pass.
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;

__device__ __forceinline__ uint64_t smem_desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                 uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) { return smem_desc_sw(saddr, 16, 1024, kLayoutSw128); }
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, uint32_t block_bytes) {
    return smem_desc_sw(saddr, block_bytes, 512, kLayoutSw128Base32);
}

// The same two descriptors as (low word, constant high word): inside an MMA issue loop the operand moves by a
// multiple of 16 bytes, so the next descriptor is ONE 32-bit add of (bytes >> 4) to the low word instead of the
// add / shift / mask / or chain on the uniform datapath (6-13 instead of 16-17 instructions between MMAs; worth
// 1-2 % in the kernels, which are bound by the softmax -> P -> MMA dependency chain, not by issue).
// Shared-memory addresses stay below 2^18, so the 14-bit address field never carries into the LBO field.
constexpr uint32_t kDescHiKmajor = (1024u >> 4) | (1u << 14) | (kLayoutSw128 << 29);
constexpr uint32_t kDescHiMnmajor = (512u >> 4) | (1u << 14) | (kLayoutSw128Base32 << 29);
__device__ __forceinline__ uint32_t kmajor_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3fffu) | (1u << 16); }
__device__ __forceinline__ uint32_t mnmajor_lo(uint32_t saddr, uint32_t block_bytes) {
    return ((saddr >> 4) & 0x3fffu) | (((block_bytes >> 4) & 0x3fffu) << 16);
}
__device__ __forceinline__ uint64_t desc_k(uint32_t lo) { return ((uint64_t)kDescHiKmajor << 32) | lo; }
__device__ __forceinline__ uint64_t desc_mn(uint32_t lo) { return ((uint64_t)kDescHiMnmajor << 32) | lo; }

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M=128 lanes x 8 columns per k-step) is read from TMEM
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// registers -> TMEM, 32 lanes x 16 consecutive columns (lane = 32*(warp%4) + laneid)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
        "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]),
        "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// TMA: one elected thread arms the barrier with the byte count, then issues box loads
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// lo part of the truncation split: v = trunc_tf32(v) + lo, lo rounded to tf32
__device__ __forceinline__ float lo_tf32(float v) {
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    uint32_t l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hi));
    return __uint_as_float(l);
}
__device__ __forceinline__ float4 lo4_tf32(const float4& v) {
    return make_float4(lo_tf32(v.x), lo_tf32(v.y), lo_tf32(v.z), lo_tf32(v.w));
}

// fp32 -> (hi, lo) with hi exactly representable in tf32 (round to nearest) and
// lo = v - hi (exact); hi*hi' + lo*hi' + hi*lo' recovers the fp32 product to ~2^-21.
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    lo = v - hi;
}
__device__ __forceinline__ void split4(const float4& v, float4& hi, float4& lo) {
    split_tf32(v.x, hi.x, lo.x);
    split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z);
    split_tf32(v.w, hi.w, lo.w);
}

}  // namespace tc
}  // namespace davi
