// Dev probe (GPU box): which shared-memory layout / descriptor convention does
// tcgen05.mma kind::tf32 use with SWIZZLE_NONE?  One MMA (M=128, N, K=8) per test
// on small-integer matrices, D compared exactly with the CPU product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I deep_attentive_vi_b200/csrc -I include \
//        tools/umma_probe.cu -o gpurun_out/umma_probe && gpurun_out/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_common.cuh"

using namespace davi::tc;

struct Test {
    int N;                 // MMA N
    int a_mn, b_mn;        // operand major-ness flags for the instruction descriptor
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    int a_words, b_words;  // words of smem used by each operand image
};

__global__ void probe(const float* a_img, const float* b_img, float* d_out, Test t) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    float* sa = reinterpret_cast<float*>(smem);
    float* sb = reinterpret_cast<float*>(smem + 65536);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < t.a_words; i += blockDim.x) sa[i] = a_img[i];
    for (int i = tid; i < t.b_words; i += blockDim.x) sb[i] = b_img[i];
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc(&tbase, 256);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tbase;
    if (tid == 0) {
        const uint32_t idesc = idesc_tf32(128, t.N, t.a_mn, t.b_mn);
        mma_tf32(tm, smem_desc(smem_u32(sa), t.a_lbo, t.a_sbo), smem_desc(smem_u32(sb), t.b_lbo, t.b_sbo),
                 idesc, false);
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t trow = tm + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < t.N; c += 16) {
        float v[16];
        tmem_ld16(trow + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) d_out[tid * 256 + c + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}

static int run(const char* name, const std::vector<float>& A, const std::vector<float>& B,   // logical A[128][8], B[N][8]
               const std::vector<float>& a_img, const std::vector<float>& b_img, Test t) {
    float *da, *db, *dd;
    cudaMalloc(&da, 65536); cudaMalloc(&db, 65536); cudaMalloc(&dd, 128 * 256 * 4);
    cudaMemset(da, 0, 65536); cudaMemset(db, 0, 65536);
    cudaMemcpy(da, a_img.data(), a_img.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b_img.data(), b_img.size() * 4, cudaMemcpyHostToDevice);
    t.a_words = (int)a_img.size(); t.b_words = (int)b_img.size();
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    probe<<<1, 128, 131072>>>(da, db, dd, t);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<float> D(128 * 256);
    cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; int first = -1;
    for (int i = 0; i < 128; ++i)
        for (int n = 0; n < t.N; ++n) {
            float w = 0; for (int k = 0; k < 8; ++k) w += A[i * 8 + k] * B[n * 8 + k];
            if (D[i * 256 + n] != w) { if (first < 0) first = i * 256 + n; ++bad; }
        }
    printf("%-44s N=%3d  mismatches %5d / %d", name, t.N, bad, 128 * t.N);
    if (bad) printf("  (first at row %d col %d: got %g)", first / 256, first % 256, D[first]);
    printf("\n");
    cudaFree(da); cudaFree(db); cudaFree(dd);
    return bad;
}

// Decode the hardware's address map for the B operand: A[i][k] = (i%8 == k), B image
// has ONE non-zero word w; D[i][n] != 0 exactly where the hardware read word w as B(n, k=i%8).
static void decode(int b_mn, uint32_t lbo, uint32_t sbo, int N, const std::vector<int>& words) {
    std::vector<float> A(128 * 8, 0.f);
    for (int i = 0; i < 128; ++i) A[i * 8 + (i % 8)] = 1.f;
    std::vector<float> a_img(128 * 8);
    for (int r = 0; r < 128; ++r) for (int k = 0; k < 8; ++k) a_img[(k / 4) * 128 * 4 + r * 4 + (k % 4)] = A[r * 8 + k];
    printf("decode b_mn=%d LBO=%u SBO=%u N=%d\n", b_mn, lbo, sbo, N);
    float *da, *db, *dd;
    cudaMalloc(&da, 65536); cudaMalloc(&db, 65536); cudaMalloc(&dd, 128 * 256 * 4);
    cudaMemcpy(da, a_img.data(), a_img.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    std::vector<float> D(128 * 256);
    for (int w : words) {
        std::vector<float> b_img(16384, 0.f);
        b_img[w] = 1.f;
        cudaMemcpy(db, b_img.data(), b_img.size() * 4, cudaMemcpyHostToDevice);
        Test t{}; t.N = N; t.a_mn = 0; t.b_mn = b_mn; t.a_lbo = 128 * 16; t.a_sbo = 128; t.b_lbo = lbo; t.b_sbo = sbo;
        t.a_words = 1024; t.b_words = 16384;
        probe<<<1, 128, 131072>>>(da, db, dd, t);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error\n"); exit(1); }
        cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
        printf("  word %5d (byte %6d):", w, w * 4);
        int found = 0;
        for (int i = 0; i < 8; ++i) for (int n = 0; n < N; ++n)
            if (D[i * 256 + n] != 0.f) { printf(" (n=%d,k=%d)", n, i); ++found; }
        if (!found) printf(" -");
        printf("\n");
    }
    cudaFree(da); cudaFree(db); cudaFree(dd);
}

int main(int argc, char** argv) {
    if (argc > 1) {
        std::vector<int> words;
        for (int w = 0; w < 40; ++w) words.push_back(w);
        for (int w : {64, 68, 96, 128, 132, 256, 260, 512, 516, 1024, 1028, 2048, 2052}) words.push_back(w);
        decode(0, 64 * 16, 128, 64, words);      // sanity: K-major, known-good
        decode(1, 4096, 128, 64, words);
        decode(1, 128, 4096, 64, words);
        decode(1, 2048, 256, 64, words);
        return 0;
    }

    srand(1);
    const int N = 64;
    std::vector<float> A(128 * 8), B(256 * 8);
    for (auto& v : A) v = (float)(rand() % 7 - 3);
    for (auto& v : B) v = (float)(rand() % 7 - 3);

    // K-major image: byte(r,k) = (k/4)*(R*16) + r*16 + (k%4)*4   [chunk-planar]
    auto kmajor = [](const std::vector<float>& M, int R) {
        std::vector<float> img(R * 8);
        for (int r = 0; r < R; ++r) for (int k = 0; k < 8; ++k) img[(k / 4) * R * 4 + r * 4 + (k % 4)] = M[r * 8 + k];
        return img;
    };
    // MN-major image of B[N][8]: byte(k,c) = (c/4)*128 + (k%8)*16 + (c%4)*4
    auto mnmajor = [](const std::vector<float>& M, int R) {
        std::vector<float> img(R * 8);
        for (int c = 0; c < R; ++c) for (int k = 0; k < 8; ++k) img[(c / 4) * 32 + k * 4 + (c % 4)] = M[c * 8 + k];
        return img;
    };
    Test t{};
    t.N = N;
    // 1: both K-major, LBO = chunk stride, SBO = 8-row-group stride
    t.a_mn = 0; t.b_mn = 0; t.a_lbo = 128 * 16; t.a_sbo = 128; t.b_lbo = N * 16; t.b_sbo = 128;
    run("K-major  LBO=chunk SBO=rowgroup", A, B, kmajor(A, 128), kmajor(B, N), t);
    // 2: swapped
    t.a_lbo = 128; t.a_sbo = 128 * 16; t.b_lbo = 128; t.b_sbo = N * 16;
    run("K-major  LBO=rowgroup SBO=chunk (swapped)", A, B, kmajor(A, 128), kmajor(B, N), t);
    // N = 128 and 256 with the K-major convention of test 1
    for (int n : {128, 256}) {
        t.N = n; t.a_lbo = 128 * 16; t.a_sbo = 128; t.b_lbo = n * 16; t.b_sbo = 128;
        run("K-major  conv.1 wide N", A, B, kmajor(A, 128), kmajor(B, n), t);
    }
    // 3: B MN-major: SBO = 4-column-group stride (128), LBO = 8-k group stride (unused for K=8)
    for (int aconv = 0; aconv < 2; ++aconv)
    for (int n : {64, 160, 256}) {
        t.N = n; t.a_mn = 0; t.b_mn = 1;
        t.a_lbo = aconv ? 128 : 128 * 16; t.a_sbo = aconv ? 128 * 16 : 128;
        printf("A conv %d: ", aconv + 1);
        t.b_lbo = n * 32; t.b_sbo = 128;
        run("B MN-major  SBO=colgroup(128) LBO=kgroup", A, B, kmajor(A, 128), mnmajor(B, n), t);
        printf("A conv %d: ", aconv + 1);
        t.b_lbo = 128; t.b_sbo = n * 32;
        run("B MN-major  LBO=colgroup(128) SBO=kgroup (swapped)", A, B, kmajor(A, 128), mnmajor(B, n), t);
    }
    return 0;
}
