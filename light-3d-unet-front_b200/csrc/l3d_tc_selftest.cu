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

// ---- kind::tf32 self-test: D[128][N] = G[128][K] . W[N][K]^T, K-major operands in the planar layout [K/4][rows][4].
// (MN-major tf32 operands return zeros on this part; the voxel reductions of the backward kernels use bf16, below.)
namespace {
__global__ void __launch_bounds__(128) tc_selftest_tf32_kernel(const float *__restrict__ G, const float *__restrict__ WU, int mode, int M,
                                                               int K, int N, float *__restrict__ D) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int ncols = 32;
    while (ncols < N) ncols <<= 1;
    if (warp == 0) tc::tmem_alloc(&s_tmem, ncols);
    if (tid == 32) tc::mbar_init(&s_bar, 1);
    unsigned char *sA = smem;                       // up to 64 KB: 128-row planes
    unsigned char *sB = smem + 64 * 1024;
    for (int i = tid; i < 128 * K; i += 128) { const int k = i % K, r = i / K; *reinterpret_cast<float *>(sA + tc::tile_off32(r, k, 128)) = G[i]; }
    for (int i = tid; i < N * K; i += 128) { const int k = i % K, n = i / K; *reinterpret_cast<float *>(sB + tc::tile_off32(n, k, N)) = WU[i]; }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        {
            const uint32_t idesc = tc::idesc_tf32_m128(N, false, false);
            for (int j = 0; j < K / 8; ++j) {
                const uint64_t ad = tc::smem_desc(tc::smem_u32(sA) + 2 * j * 128 * 16, 128 * 16, 128);
                const uint64_t bd = tc::smem_desc(tc::smem_u32(sB) + 2 * j * N * 16, N * 16, 128);
                tc::mma_tf32(tmem, ad, bd, idesc, j > 0);
            }
        }
        tc::mma_commit(&s_bar);
    }
    tc::mbar_wait(&s_bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        const int row = warp * 32 + lane;
        for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}
}  // namespace

extern "C" int l3d_tc_selftest_tf32(const float *G, const float *WU, int mode, int M, int K, int N, float *D, void *stream) {
    L3D_REQUIRE(G && WU && D && mode == 0 && N % 16 == 0 && N >= 16 && N <= 256 && K % 8 == 0 && K >= 8 && K <= 128, "l3d_tc_selftest_tf32: bad shape");
    (void)M;
    const size_t smem = 192 * 1024;
    cudaFuncSetAttribute(tc_selftest_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_selftest_tf32_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(G, WU, mode, M, K, N, D);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_tc_selftest_tf32 launch");
    return 0;
}

// ---- 16-bit MN-major voxel reduction: D[m][n] = sum_{v<128} G[v][m] * U[v][n], operands bf16 in voxel-planar tiles
// [channel/8][128 voxels][8 channels] (the layout the conv kernels already use), both operands MN-major.
namespace {
__global__ void __launch_bounds__(128) tc_selftest_mn16_kernel(const float *__restrict__ G, const float *__restrict__ U, int M, int N, int var,
                                                               float *__restrict__ D) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int ncols = 32;
    while (ncols < N) ncols <<= 1;
    if (warp == 0) tc::tmem_alloc(&s_tmem, ncols);
    if (tid == 32) tc::mbar_init(&s_bar, 1);
    unsigned char *sA = smem, *sB = smem + 64 * 1024;
    // both operands bf16: a kind::f16 MMA whose A and B formats differ (bf16 x fp16) raises "illegal instruction" on
    // sm_100a (measured), so the backward kernels split the stored fp16 activations exactly into bf16 hi + lo
    for (int i = tid; i < 128 * M; i += 128) { const int m = i % M, v = i / M; *reinterpret_cast<__nv_bfloat16 *>(sA + tc::tile_off(v, m, 128)) = __float2bfloat16_rn(G[i]); }
    for (int i = tid; i < 128 * N; i += 128) { const int n = i % N, v = i / N; *reinterpret_cast<__nv_bfloat16 *>(sB + tc::tile_off(v, n, 128)) = __float2bfloat16_rn(U[i]); }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_16b_m128(N, 1, 1, true, true);
        for (int j = 0; j < 128 / 16; ++j) {          // 16 voxels per MMA = two 8-voxel groups, 128 B apart
            const uint32_t lbo = (var & 1) ? 128u * 16u : 128u, sbo = (var & 1) ? 128u : 128u * 16u;
            const uint64_t ad = tc::smem_desc(tc::smem_u32(sA) + j * 256, lbo, sbo);
            const uint64_t bd = tc::smem_desc(tc::smem_u32(sB) + j * 256, lbo, sbo);
            tc::mma_f16(tmem, ad, bd, idesc, j > 0);
        }
        tc::mma_commit(&s_bar);
    }
    tc::mbar_wait(&s_bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        const int row = warp * 32 + lane;
        for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}
}  // namespace

extern "C" int l3d_tc_selftest_mn16(const float *G, const float *U, int M, int N, int var, float *D, void *stream) {
    L3D_REQUIRE(G && U && D && N % 16 == 0 && N >= 16 && N <= 256 && M % 8 == 0 && M >= 8 && M <= 128, "l3d_tc_selftest_mn16: bad shape");
    const size_t smem = 192 * 1024;
    cudaFuncSetAttribute(tc_selftest_mn16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_selftest_mn16_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(G, U, M, N, var, D);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_tc_selftest_mn16 launch");
    return 0;
}
