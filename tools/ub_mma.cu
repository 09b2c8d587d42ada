// Micro-benchmark (development tool): cycles per tcgen05.mma (M=128, K=16, fp16, SS mode) as a function of N and of
// the A-operand shared-memory layout (aligned canonical tile vs. the shifted-window layout of the implicit-GEMM conv).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I light-3d-unet-front_b200/csrc -o tools/ub_mma tools/ub_mma.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "l3d_tc.cuh"

struct Variant { int n; int sbo; int lbo; int chains; };

template <int MODE>
__global__ void __launch_bounds__(128) ub_kernel(Variant v, int iters, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
    if (tid == 32) tc::mbar_init(&s_bar, 1);
    for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x00010001u;   // tiny
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_f16_m128(v.n);
        const uint32_t sA = tc::smem_u32(smem), sB = sA + 96 * 1024;
        const uint64_t bd = tc::smem_desc(sB, v.n * 16, 128);
        const uint64_t ad0 = tc::smem_desc(sA, v.lbo, v.sbo);
        const uint32_t rowp = (uint32_t)v.sbo >> 4;
        const uint32_t d1 = tmem + (v.chains > 1 ? v.n : 0);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int tap = 0; tap < 27; ++tap) {
                const int dz = tap / 9, dy = (tap / 3) % 3, dx = tap % 3;
                uint32_t off = 0;          // in 16-byte units
                if (MODE == 1) off = (dz * 18 + dy) * rowp + dx;        // 27 taps
                if (MODE == 2) off = dx;                                // dx only
                if (MODE == 3) off = 1;                                 // constant +16 B
                if (MODE == 4) off = (dz * 18 + dy) * rowp;             // row-aligned shifts only
                tc::mma_f16(tmem, ad0 + off, bd, idesc, 1u);
                tc::mma_f16(d1, ad0 + off + 18 * rowp, bd, idesc, 1u);
            }
        }
        tc::mma_commit(&s_bar);
    }
    tc::mbar_wait(&s_bar, 0);
    if (tid == 0) { t1 = clock64(); out[blockIdx.x] = t1 - t0; }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// MN-major bf16 operands over voxel-planar tiles (the weight-gradient form): cycles per MMA
__global__ void __launch_bounds__(128) ub_mn_kernel(int n, int mn, int iters, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
    if (tid == 32) tc::mbar_init(&s_bar, 1);
    for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x00010001u;
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_16b_m128(n, 1, 1, mn != 0, mn != 0);
        const uint32_t sA = tc::smem_u32(smem), sB = sA + 96 * 1024;
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint64_t ad = mn ? tc::smem_desc(sA + j * 256, 128, 2048) : tc::smem_desc(sA + 2 * j * 2048, 2048, 128);
                const uint64_t bd = mn ? tc::smem_desc(sB + j * 256, 128, 2048) : tc::smem_desc(sB + 2 * j * n * 16, n * 16, 128);
                tc::mma_f16(tmem, ad, bd, idesc, 1u);
            }
        }
        tc::mma_commit(&s_bar);
    }
    tc::mbar_wait(&s_bar, 0);
    if (tid == 0) { t1 = clock64(); out[blockIdx.x] = t1 - t0; }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

template <int MODE>
static double run(Variant v, int iters, long long *d_out) {
    long long h_out[148];
    cudaFuncSetAttribute(ub_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    ub_kernel<MODE><<<148, 128, 200 * 1024>>>(v, iters, d_out);
    ub_kernel<MODE><<<148, 128, 200 * 1024>>>(v, iters, d_out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
    cudaMemcpy(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h_out[i] > mx ? h_out[i] : mx;
    return (double)mx / (iters * 54.0);
}

int main() {
    long long *d_out;
    cudaMalloc(&d_out, sizeof(long long) * 148);
    const int iters = 40;
    const int ns[] = {16, 32, 48, 64, 96, 128, 256};
    struct L { const char *name; int sbo, lbo, mode; } layouts[] = {
        {"canonical sbo128 lbo2048 aligned", 128, 2048, 0},
        {"canonical +16B offset", 128, 2048, 3},
        {"conv: sbo160 27 taps (current kernel)", 160, 11584, 1},
        {"conv: sbo160 dx only", 160, 11584, 2},
        {"conv: sbo160 dz/dy only", 160, 11584, 4},
        {"conv: sbo160 no shift", 160, 11584, 0},
        {"conv: sbo256 27 taps", 256, 18432, 1},
        {"conv: sbo256 dz/dy only", 256, 18432, 4},
        {"conv: sbo256 no shift", 256, 18432, 0},
    };
    cudaFuncSetAttribute(ub_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int mn = 0; mn <= 1; ++mn) {
        printf("%-40s          :", mn ? "bf16 MN-major A and B (voxel-planar)" : "bf16 K-major A and B");
        for (int n : {16, 32, 64, 128}) {
            long long h_out[148];
            ub_mn_kernel<<<148, 128, 200 * 1024>>>(n, mn, 100, d_out);
            ub_mn_kernel<<<148, 128, 200 * 1024>>>(n, mn, 100, d_out);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            cudaMemcpy(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < 148; ++i) mx = h_out[i] > mx ? h_out[i] : mx;
            printf(" N%-3d %6.1f", n, (double)mx / 800.0);
        }
        printf("\n");
    }
    for (int chains = 1; chains <= 1; ++chains)
        for (auto &l : layouts) {
            printf("%-40s chains=%d :", l.name, chains);
            for (int n : ns) {
                if (chains * n > 512) { printf("     -"); continue; }
                Variant v{n, l.sbo, l.lbo, chains};
                double c = 0;
                switch (l.mode) {
                    case 0: c = run<0>(v, iters, d_out); break;
                    case 1: c = run<1>(v, iters, d_out); break;
                    case 2: c = run<2>(v, iters, d_out); break;
                    case 3: c = run<3>(v, iters, d_out); break;
                    default: c = run<4>(v, iters, d_out); break;
                }
                printf(" N%-3d %6.1f", n, c);
            }
            printf("\n");
        }
    return 0;
}
