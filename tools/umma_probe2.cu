// Dev probe 2 (GPU box): validates, with exact small-integer matrices, every building
// block the tensor-core nonlocal kernels rely on, and times tcgen05.mma issue rates:
//   T1  K-major SWIZZLE_128B operands (rows of 32 tf32 = one 128-byte swizzle row)
//   T2  MN-major SWIZZLE_128B B operand (channel-contiguous value rows, no transpose)
//   T3  A operand from TMEM (tcgen05.mma ... [d], [a_tmem], b_desc)
//   T4  does the tensor core truncate or round fp32 bit patterns fed as tf32?
//   T5  TMA (cp.async.bulk.tensor, SWIZZLE_128B) lands tiles in exactly those layouts,
//       with out-of-bounds zero fill for ragged channel counts
//   T6  MMA throughput: SWIZZLE_NONE vs SWIZZLE_128B, SS vs TS, N = 128/144/160/256
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I deep_attentive_vi_b200/csrc -I include \
//        tools/umma_probe2.cu -o tools/_bin/umma_probe2 && tools/_bin/umma_probe2
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_common.cuh"

using namespace davi::tc;

struct Cfg {
    int N, ksteps, a_tmem, repeat;
    int a_mn, b_mn;
    uint32_t a_off, a_lbo, a_sbo, a_layout, a_kadv;   // bytes (a_kadv in TMEM columns when a_tmem)
    uint32_t b_off, b_lbo, b_sbo, b_layout, b_kadv;
    int a_bytes, b_bytes;
};

__device__ __forceinline__ uint64_t desc_of(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}

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

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
        "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]),
        "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]),
        "r"(u[17]), "r"(u[18]), "r"(u[19]), "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]),
        "r"(u[25]), "r"(u[26]), "r"(u[27]), "r"(u[28]), "r"(u[29]), "r"(u[30]), "r"(u[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

constexpr int kSmemBytes = 200 * 1024;   // usable; +1 KB is requested for the 1024-byte alignment
constexpr int kACol = 256;   // TMEM column of a TMEM-resident A operand

// a_img: smem byte image of A (SS) or plain row-major [128][32] floats (TS).
__global__ void __launch_bounds__(128) probe(const float* a_img, const float* b_img, float* d_out,
                                             long long* cycles, Cfg t) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* sa = reinterpret_cast<float*>(smem + t.a_off);
    float* sb = reinterpret_cast<float*>(smem + t.b_off);
    if (!t.a_tmem)
        for (int i = tid; i < t.a_bytes / 4; i += blockDim.x) sa[i] = a_img[i];
    for (int i = tid; i < t.b_bytes / 4; i += blockDim.x) sb[i] = b_img[i];
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc(&tbase, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tbase;
    const uint32_t trow = tm + ((uint32_t)(warp * 32) << 16);
    if (t.a_tmem) {
        float v[32];
        for (int j = 0; j < 32; ++j) v[j] = a_img[tid * 32 + j];
        tmem_st32(trow + kACol, v);
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    long long t0 = 0;
    if (tid == 0) {
        const uint32_t idesc = idesc_tf32(128, t.N, t.a_mn, t.b_mn);
        const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
        t0 = clock64();
        for (int rep = 0; rep < t.repeat; ++rep)
            for (int ks = 0; ks < t.ksteps; ++ks) {
                const uint64_t bd = desc_of(b0 + ks * t.b_kadv, t.b_lbo, t.b_sbo, t.b_layout);
                const bool acc = (rep | ks) != 0;
                if (t.a_tmem) mma_tf32_ts(tm, tm + kACol + ks * t.a_kadv, bd, idesc, acc);
                else mma_tf32(tm, desc_of(a0 + ks * t.a_kadv, t.a_lbo, t.a_sbo, t.a_layout), bd, idesc, acc);
            }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (tid == 0 && cycles) *cycles = clock64() - t0;
    tc_fence_after();
    for (int c = 0; c < t.N; c += 16) {
        float v[16];
        tmem_ld16(trow + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) d_out[tid * 256 + c + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- host-side layout builders -------------------------------------------------
// K-major SWIZZLE_128B image of X[rows][32]
static std::vector<float> kmajor_sw128(const std::vector<float>& X, int rows) {
    std::vector<float> img(rows * 32, 0.f);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < 32; ++c) {
            const int byte = (r / 8) * 1024 + (r % 8) * 128 + (((c / 4) ^ (r % 8)) * 16) + (c % 4) * 4;
            img[byte / 4] = X[r * 32 + c];
        }
    return img;
}
// MN-major image of V[keys][C] (C padded to a multiple of 32): blocks of 32 channels, each block =
// keys rows of 128 bytes.  base32 = true : SWIZZLE_128B_BASE32B (Swizzle<2,5,2>: 32-byte chunk ^= row % 4),
// the only MN-major layout of 32-bit operands; base32 = false: plain SWIZZLE_128B (16-byte chunk ^= row % 8).
static std::vector<float> mnmajor_sw128(const std::vector<float>& V, int keys, int C, bool base32 = true) {
    const int Cp = (C + 31) / 32 * 32;
    std::vector<float> img(keys * Cp, 0.f);
    for (int k = 0; k < keys; ++k)
        for (int c = 0; c < C; ++c) {
            int byte = (c / 32) * (keys * 128) + k * 128;
            if (base32) byte += ((((c % 32) / 8) ^ (k % 4)) * 32) + (c % 8) * 4;
            else byte += ((((c % 32) / 4) ^ (k % 8)) * 16) + (c % 4) * 4;
            img[byte / 4] = V[k * C + c];
        }
    return img;
}
// SWIZZLE_NONE K-major image of X[rows][K] (K multiple of 8): chunk-planar
static std::vector<float> kmajor_none(const std::vector<float>& X, int rows, int K) {
    std::vector<float> img(rows * K, 0.f);
    for (int r = 0; r < rows; ++r)
        for (int k = 0; k < K; ++k) img[(k / 4) * rows * 4 + r * 4 + (k % 4)] = X[r * K + k];
    return img;
}

struct Dev {
    float *a, *b, *d;
    long long* cyc;
    Dev() {
        cudaMalloc(&a, kSmemBytes); cudaMalloc(&b, kSmemBytes); cudaMalloc(&d, 128 * 256 * 4);
        cudaMalloc(&cyc, 8);
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024);
    }
};

static std::vector<float> launch(Dev& dv, const std::vector<float>& a_img, const std::vector<float>& b_img, Cfg t,
                                 long long* cycles = nullptr) {
    cudaMemset(dv.a, 0, kSmemBytes); cudaMemset(dv.b, 0, kSmemBytes);
    cudaMemcpy(dv.a, a_img.data(), a_img.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dv.b, b_img.data(), b_img.size() * 4, cudaMemcpyHostToDevice);
    t.a_bytes = (int)a_img.size() * 4; t.b_bytes = (int)b_img.size() * 4;
    if (t.repeat < 1) t.repeat = 1;
    probe<<<1, 128, kSmemBytes + 1024>>>(dv.a, dv.b, dv.d, dv.cyc, t);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
    std::vector<float> D(128 * 256);
    cudaMemcpy(D.data(), dv.d, D.size() * 4, cudaMemcpyDeviceToHost);
    if (cycles) cudaMemcpy(cycles, dv.cyc, 8, cudaMemcpyDeviceToHost);
    return D;
}

static int compare(const char* name, const std::vector<float>& D, const std::vector<double>& want, int N) {
    int bad = 0, first = -1;
    for (int i = 0; i < 128; ++i)
        for (int n = 0; n < N; ++n)
            if ((double)D[i * 256 + n] != want[i * N + n]) { if (first < 0) first = i * 256 + n; ++bad; }
    printf("%-58s N=%3d mismatches %5d / %d", name, N, bad, 128 * N);
    if (bad) printf("  (first row %d col %d: got %g want %g)", first / 256, first % 256, D[first],
                    want[(first / 256) * N + first % 256]);
    printf("\n");
    return bad;
}

static float trunc_tf32(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xffffe000u; memcpy(&v, &u, 4); return v; }
static float rna_tf32(float v) { uint32_t u; memcpy(&u, &v, 4); u += 0x1000u; u &= 0xffffe000u; memcpy(&v, &u, 4); return v; }

// ---- TMA --------------------------------------------------------------------------
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128) tma_probe(const __grid_constant__ CUtensorMap tmap, float* out, int boxes,
                                                 int box_bytes, int c1, int c2) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)),
                     "r"((uint32_t)(boxes * box_bytes))
                     : "memory");
        for (int bx = 0; bx < boxes; ++bx)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem + bx * box_bytes)),
                "l"(&tmap), "r"(bx * 32), "r"(c1), "r"(c2), "r"(smem_u32(&bar))
                : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = tid; i < boxes * box_bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

int main() {
    Dev dv;
    srand(1);
    auto rnd = [](std::vector<float>& v) { for (auto& x : v) x = (float)(rand() % 7 - 3); };
    int fails = 0;

    // ---------------- T1: K-major SW128, 4 k-steps (K = 32) ----------------
    {
        const int N = 128;
        std::vector<float> A(128 * 32), B(N * 32);
        rnd(A); rnd(B);
        std::vector<double> want(128 * N);
        for (int i = 0; i < 128; ++i) for (int n = 0; n < N; ++n) {
            double w = 0; for (int k = 0; k < 32; ++k) w += (double)A[i * 32 + k] * B[n * 32 + k];
            want[i * N + n] = w;
        }
        Cfg t{}; t.N = N; t.ksteps = 4; t.a_off = 0; t.b_off = 65536;
        t.a_layout = t.b_layout = 2; t.a_sbo = t.b_sbo = 1024; t.a_lbo = t.b_lbo = 16; t.a_kadv = t.b_kadv = 32;
        fails += compare("T1 K-major SW128 (SBO=1024, +32B per k-step)", launch(dv, kmajor_sw128(A, 128), kmajor_sw128(B, N), t), want, N);
    }
    // ---------------- T2/T3: D = P[128 x 32 keys] * V[32 keys x C] -----------
    for (int C : {96, 160, 256}) {
        const int keys = 32;
        std::vector<float> P(128 * keys), V(keys * C);
        rnd(P); rnd(V);
        std::vector<double> want(128 * C);
        for (int i = 0; i < 128; ++i) for (int c = 0; c < C; ++c) {
            double w = 0; for (int k = 0; k < keys; ++k) w += (double)P[i * keys + k] * V[k * C + c];
            want[i * C + c] = w;
        }
        Cfg t{}; t.N = C; t.ksteps = 4; t.a_off = 0; t.b_off = 65536; t.b_mn = 1;
        t.a_layout = 2; t.a_sbo = 1024; t.a_lbo = 16; t.a_kadv = 32;
        t.b_layout = 1; t.b_lbo = keys * 128; t.b_sbo = 512; t.b_kadv = 1024;
        fails += compare("T2 B MN-major SW128_BASE32B (LBO=blk, SBO=512, +1024B/k-step)",
                         launch(dv, kmajor_sw128(P, 128), mnmajor_sw128(V, keys, C), t), want, C);
        Cfg s = t; s.b_lbo = 512; s.b_sbo = keys * 128;
        compare("   (swapped LBO/SBO, expected to FAIL)", launch(dv, kmajor_sw128(P, 128), mnmajor_sw128(V, keys, C), s), want, C);
        Cfg u = t; u.a_tmem = 1; u.a_kadv = 8;
        fails += compare("T3 A from TMEM (+8 columns per k-step), B MN-major SW128",
                         launch(dv, P, mnmajor_sw128(V, keys, C), u), want, C);
    }
    // A MN-major (dS^T read as A with M contiguous): A^T stored [K=32 rows x M=128] MN-major SW128
    {
        const int N = 32, K = 32;
        std::vector<float> At(K * 128), Bt(K * N);          // At[k][m], Bt[k][n]
        rnd(At); rnd(Bt);
        std::vector<double> want(128 * N);
        for (int i = 0; i < 128; ++i) for (int n = 0; n < N; ++n) {
            double w = 0; for (int k = 0; k < K; ++k) w += (double)At[k * 128 + i] * Bt[k * N + n];
            want[i * N + n] = w;
        }
        Cfg t{}; t.N = N; t.ksteps = 4; t.a_off = 0; t.b_off = 65536; t.a_mn = 1; t.b_mn = 1;
        t.a_layout = 1; t.a_lbo = K * 128; t.a_sbo = 512; t.a_kadv = 1024;
        t.b_layout = 1; t.b_lbo = K * 128; t.b_sbo = 512; t.b_kadv = 1024;
        fails += compare("T2b A and B both MN-major SW128_BASE32B", launch(dv, mnmajor_sw128(At, K, 128), mnmajor_sw128(Bt, K, N), t), want, N);
    }
    // ---------------- T4: truncate or round? ----------------
    {
        const int N = 128;
        std::vector<float> A(128 * 32, 0.f), B(N * 32, 0.f);
        for (int i = 0; i < 128; ++i) A[i * 32] = 1.0f + (float)(i + 1) * 1.1920929e-7f * 37.f;   // low mantissa bits set
        for (int n = 0; n < N; ++n) B[n * 32] = 1.0f;
        auto D = launch(dv, kmajor_sw128(A, 128), kmajor_sw128(B, N), [&] {
            Cfg t{}; t.N = N; t.ksteps = 4; t.b_off = 65536; t.a_layout = t.b_layout = 2; t.a_sbo = t.b_sbo = 1024;
            t.a_lbo = t.b_lbo = 16; t.a_kadv = t.b_kadv = 32; return t; }());
        int ntr = 0, nrn = 0, nraw = 0;
        for (int i = 0; i < 128; ++i) {
            const float g = D[i * 256];
            ntr += g == trunc_tf32(A[i * 32]); nrn += g == rna_tf32(A[i * 32]); nraw += g == A[i * 32];
        }
        printf("T4 fp32 bits fed as tf32: matches truncation %d/128, round-to-nearest %d/128, raw fp32 %d/128\n", ntr, nrn, nraw);
    }
    // ---------------- T5: TMA SW128 boxes ----------------
    {
        EncodeTiled enc = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres);
        if (!enc) { printf("T5: cuTensorMapEncodeTiled not found\n"); return 1; }
        const int Bn = 2, Nn = 64, rs = 340, C = 276, keys = 16, c_off = 64;   // V = columns [64, 340) of a [B,N,340] tensor
        std::vector<float> T((size_t)Bn * Nn * rs);
        for (size_t i = 0; i < T.size(); ++i) T[i] = (float)(i % 1000) + 1.f;
        float* dT; cudaMalloc(&dT, T.size() * 4); cudaMemcpy(dT, T.data(), T.size() * 4, cudaMemcpyHostToDevice);
        CUtensorMap tm;
        cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)Nn, (cuuint64_t)Bn};
        cuuint64_t strides[2] = {(cuuint64_t)rs * 4, (cuuint64_t)Nn * rs * 4};
        cuuint32_t box[3] = {32, (cuuint32_t)keys, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        for (int base32 = 0; base32 < 2; ++base32) {
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dT + c_off, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         base32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("T5 cuTensorMapEncodeTiled(base32=%d) -> %d\n", base32, (int)r);
        const int boxes = (C + 31) / 32, box_bytes = 32 * keys * 4;
        float* dout; cudaMalloc(&dout, boxes * box_bytes);
        cudaFuncSetAttribute(tma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024);
        const int key0 = 24, img = 1;
        tma_probe<<<1, 128, kSmemBytes + 1024>>>(tm, dout, boxes, box_bytes, key0, img);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("T5 CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> got(boxes * box_bytes / 4);
        cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
        std::vector<float> V(keys * C);
        for (int k = 0; k < keys; ++k) for (int c = 0; c < C; ++c)
            V[k * C + c] = T[((size_t)img * Nn + key0 + k) * rs + c_off + c];
        auto want = mnmajor_sw128(V, keys, C, base32 != 0);
        int bad = 0;
        for (size_t i = 0; i < want.size(); ++i) bad += want[i] != got[i];
        printf("T5 TMA %s boxes [32ch x %d keys] x %d == manual image (zero OOB fill): mismatches %d / %zu\n",
               base32 ? "SWIZZLE_128B_ATOM_32B" : "SWIZZLE_128B", keys, boxes, bad, want.size());
        fails += bad != 0;
        cudaFree(dout);
        }
    }
    // ---------------- T6: issue-rate timing ----------------
    {
        std::vector<float> zeros(kSmemBytes / 8, 0.f);
        const int reps = 64;
        auto time_cfg = [&](const char* name, Cfg t) {
            t.repeat = reps;
            long long cyc = 0;
            launch(dv, t.a_tmem ? std::vector<float>(128 * 32, 0.f) : zeros, zeros, t, &cyc);
            const double per = (double)cyc / (reps * t.ksteps);
            printf("T6 %-52s N=%3d  %7.1f cyc/MMA (floor %5.1f)  -> %6.0f FMA/clk\n", name, t.N, per, t.N / 2.0,
                   128.0 * t.N * 8 / per);
        };
        for (int N : {32, 64, 128, 160, 256}) {
            Cfg t{}; t.N = N; t.ksteps = 4; t.b_off = 98304;
            t.a_layout = 0; t.a_lbo = 128 * 16; t.a_sbo = 128; t.a_kadv = 2 * 128 * 16;
            t.b_layout = 0; t.b_lbo = N * 16; t.b_sbo = 128; t.b_kadv = 2 * N * 16;
            time_cfg("SS  A,B K-major SWIZZLE_NONE (current kernel)", t);
            t.a_layout = t.b_layout = 2; t.a_sbo = t.b_sbo = 1024; t.a_lbo = t.b_lbo = 16; t.a_kadv = t.b_kadv = 32;
            time_cfg("SS  A,B K-major SWIZZLE_128B", t);
            t.b_mn = 1; t.b_layout = 1; t.b_lbo = 32 * 128; t.b_sbo = 512; t.b_kadv = 1024;
            time_cfg("SS  A K-major SW128, B MN-major BASE32B", t);
            t.a_tmem = 1; t.a_kadv = 8;
            time_cfg("TS  A in TMEM, B MN-major SW128", t);
        }
    }
    printf(fails ? "PROBE: %d FAILED\n" : "PROBE: all layout tests passed\n", fails);
    return 0;
}
