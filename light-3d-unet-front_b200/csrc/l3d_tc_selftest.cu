// Self-test of the tcgen05 building blocks (l3d_tc.cuh): D[128*MT][N] = A[128*MT][K] . W[N][K]^T with fp16 operands
// staged in the no-swizzle K-major layout and fp32 accumulation in TMEM.  Used by tests/test_gpu_tc.py.
#include "l3d_common.cuh"
#include "l3d_tc.cuh"

namespace {
__global__ void __launch_bounds__(128) tc_selftest_kernel(const float *__restrict__ A, const float *__restrict__ Wt, int MT, int K, int N,
                                                          float *__restrict__ D) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    unsigned char *sA = smem;                               // MT * 128 * K * 2
    unsigned char *sB = smem + (size_t)MT * 128 * K * 2;    // N * K * 2
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int ncols = 32;
    while (ncols < MT * N) ncols <<= 1;
    if (warp == 0) tc::tmem_alloc(&s_tmem, ncols);
    if (tid == 32) tc::mbar_init(&s_bar, 1);
    for (int i = tid; i < MT * 128 * K; i += 128) {
        const int k = i % K, r = i / K;
        *reinterpret_cast<__half *>(sA + (size_t)(r >> 7) * 128 * K * 2 + tc::tile_off(r & 127, k, 128)) = __float2half_rn(A[i]);
    }
    for (int i = tid; i < N * K; i += 128) {
        const int k = i % K, n = i / K;
        *reinterpret_cast<__half *>(sB + tc::tile_off(n, k, N)) = __float2half_rn(Wt[i]);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_f16_m128(N);
        for (int m = 0; m < MT; ++m)
            for (int j = 0; j < K / 16; ++j) {
                const uint64_t ad = tc::smem_desc(tc::smem_u32(sA) + m * 128 * K * 2 + 2 * j * 2048, 2048, 128);
                const uint64_t bd = tc::smem_desc(tc::smem_u32(sB) + 2 * j * N * 16, N * 16, 128);
                tc::mma_f16(tmem + m * N, ad, bd, idesc, j > 0);
            }
        tc::mma_commit(&s_bar);
    }
    tc::mbar_wait(&s_bar, 0);
    tc::fence_after_sync();
    for (int m = 0; m < MT; ++m)
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + m * N + c0, v);
            const int row = m * 128 + warp * 32 + lane;
            for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = v[j];
        }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}
}  // namespace

extern "C" int l3d_tc_selftest(const float *A, const float *Wt, int MT, int K, int N, float *D, void *stream) {
    L3D_REQUIRE(A && Wt && D && MT >= 1 && MT <= 4 && K % 16 == 0 && K >= 16 && N % 16 == 0 && N >= 16 && N <= 256 && MT * N <= 512,
                "l3d_tc_selftest: bad shape");
    const size_t smem = (size_t)MT * 128 * K * 2 + (size_t)N * K * 2;
    L3D_REQUIRE(smem <= 200 * 1024, "l3d_tc_selftest: tile too large");
    if (smem > 48 * 1024) cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, Wt, MT, K, N, D);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_tc_selftest launch");
    return 0;
}
