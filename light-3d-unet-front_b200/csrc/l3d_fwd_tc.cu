// Tensor-core forward of the depthwise-separable conv with the depthwise stage on the CUDA cores: the TRAINING forward
// (it saves the depthwise output u for the pointwise weight gradient) in fp16 or fp32 storage, and the fp32-storage
// inference path.
//
//   TMA halo tile (raw) -> [InstanceNorm + LeakyReLU + Dropout3d] -> depthwise 3x3x3 (fp32 FMA, CUDA cores)
//   -> pointwise 1x1x1 (+ the block's 1x1x1 shortcut) as tcgen05.mma (fp16 operands -- hi + lo pairs in fp32 storage --
//      fp32 accumulate in TMEM)
//   -> store + InstanceNorm statistics in the epilogue                  (unet3d.py:20-23, 70-72, 80-87)
//
// One persistent CTA (256 threads) walks 4x8x8-voxel tiles = two 128-row MMA tiles, 16 input channels at a time:
//   * one thread issues a 5-D TMA load of the 6x10x10x16 raw halo box (out-of-volume voxels arrive as zeros) for
//     the NEXT work item while the CUDA cores work on the current one (mbarrier complete_tx);
//   * an activation pass turns the raw box into the fp32 stencil tile (norm/act applied once per element, zero
//     outside the volume: the conv pads the *activated* tensor);
//   * the stencil writes its output straight into the K-major fp16 operand tile of the chunk (double-buffered: the MMAs
//     of chunk i run under the activation pass and the stencil of chunk i + 1, an mbarrier per buffer guards its reuse);
//     one thread issues the K=16 MMA step(s) for that chunk;
//   * the epilogue reads the accumulators with tcgen05.ld (thread = voxel row, 16 channels per load).
#include <cuda.h>

#include "l3d_common.cuh"
#include "l3d_tc.cuh"

namespace {

constexpr int TZ = 4, TY = 8, TX = 8, TV = TZ * TY * TX;
constexpr int HZ = TZ + 2, HY = TY + 2, HX = TX + 2;
constexpr int HXP = HX + 1, HPLANE = HY * HXP + 1, HVOX = HZ * HPLANE;
constexpr int CK = 16, NT = 256, MT = 2;
constexpr int ACT_ITEMS = HZ * HY * HX * 2;             // 8-channel vectors in the raw box
constexpr int ACT_PER_THREAD = (ACT_ITEMS + NT - 1) / NT;
constexpr int CHUNK_TILE = MT * 128 * CK * 2;           // 8192: one fp16 operand tile of a 16-channel chunk, [m][k/8][128 rows][8]

// Storage traits.  fp16 storage: single fp16 operands (the stored value IS the operand precision).  fp32 storage: every
// operand -- activations and weights -- is split into fp16 hi + lo (22 significand bits) and each product is three MMAs
// (hi.hi + lo.hi + hi.lo), so the pointwise GEMM carries ~fp32 accuracy on the tensor cores: this is the mode whose
// gradients match the reference tensor by tensor (tests/test_gpu_configs.py).
template <typename T> struct St;
template <> struct St<h16> { static constexpr int NP = 1, ES = 2; };
template <> struct St<float> { static constexpr int NP = 2, ES = 4; };

struct TcArgs {
    int Cin; NormDev xn;
    int N, D, H, W;
    const float *dw_w, *pw_w, *sc_w; int Cout;
    void *t; int ldt; double *t_stats;
    void *r; int ldr; double *r_stats;
    void *u; int ldu;
    int tmem_cols, wide_st;
};

__device__ __forceinline__ void split_f16(float v, __half &hi, __half &lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}
// fp32 storage: the operands of a product are scaled by powers of two before the hi + lo split and the accumulators scaled
// back in the epilogue (exact).  An fp16 `lo` part is a NORMAL number only when |v| > 2^-14 * 2^11 = 0.125: below that the pair
// carries fewer than 22 significand bits, and most weights and many activations are below that.  Unscaled, the forward GEMMs
// were 2e-6 from fp32 per layer and the parameter gradients 3.3x further from the float64 oracle than the fp32 reference is
// (median 1.6e-3 vs 4.9e-4 at 8 x 48^3; the CUDA-core forward: 4.8e-4).  Activations x 2^5 (finite up to |v| = 2047, values
// beyond saturate), weights x 2^8 (|w| < 255).
constexpr float SPLIT_AS = 32.f, SPLIT_WS = 256.f, SPLIT_INV = 1.f / (32.f * 256.f);
__device__ __forceinline__ void split_f16_scaled(float v, float s, __half &hi, __half &lo) {
    v = fminf(fmaxf(v * s, -65504.f), 65504.f);
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}
// 16 accumulator values of one voxel -> global (values as stored are returned in v for the statistics)
__device__ __forceinline__ void store16(h16 *p, float (&v)[16], bool valid) {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        pk[j] = valid ? pack_h16x2(v[2 * j], v[2 * j + 1]) : 0u;
        v[2 * j] = h16_lo(pk[j]); v[2 * j + 1] = h16_hi(pk[j]);
    }
    if (valid) {
        *reinterpret_cast<uint4 *>(p) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4 *>(p + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
}
// The same as ONE 256-bit store (STG.E.256, sm_100): p must be 32-byte aligned.  A lane's 16 channels are one 32-byte
// sector; with two 16-byte stores a warp whose lanes write to 32 different 128-byte lines pays 64 LSU wavefronts per
// voxel row instead of 32 (ConvTranspose pixel shuffle: ncu counted 76 % of the peak LSU wavefront rate at 37 % DRAM).
__device__ __forceinline__ void store16_256(h16 *p, const float (&v)[16], bool valid) {
    if (valid) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_h16x2(v[2 * j], v[2 * j + 1]);
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]),
                     "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
    }
}
// store16 with 256-bit stores (p 32-byte aligned): fp16 one, fp32 two instructions instead of two / four
__device__ __forceinline__ void store16_wide(h16 *p, float (&v)[16], bool valid) {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        pk[j] = valid ? pack_h16x2(v[2 * j], v[2 * j + 1]) : 0u;
        v[2 * j] = h16_lo(pk[j]); v[2 * j + 1] = h16_hi(pk[j]);
    }
    if (valid) st_global_256(p, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
}
__device__ __forceinline__ void store16_wide(float *p, float (&v)[16], bool valid) {
    if (valid) {
        st_global_256(p, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]),
                      __float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
        st_global_256(p + 8, __float_as_uint(v[8]), __float_as_uint(v[9]), __float_as_uint(v[10]), __float_as_uint(v[11]),
                      __float_as_uint(v[12]), __float_as_uint(v[13]), __float_as_uint(v[14]), __float_as_uint(v[15]));
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
}
__device__ __forceinline__ void store16(float *p, float (&v)[16], bool valid) {
    if (valid) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4 *>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
}

__device__ __forceinline__ void store16_256(float *p, float (&v)[16], bool valid) { store16_wide(p, v, valid); }

template <typename T>
__global__ void __launch_bounds__(NT) dwpw_tc_kernel(const __grid_constant__ CUtensorMap tmap, TcArgs A) {
    constexpr int NP = St<T>::NP, ES = St<T>::ES;
    constexpr int RAW_BYTES = HZ * HY * HX * CK * ES;      // the TMA box, dense [z][y][x][c]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_mma[2], s_tma_bar;
    __shared__ uint32_t s_tmem;
    const int Cin = A.Cin, Cout = A.Cout;
    const bool has_sc = A.sc_w != nullptr;
    const int nkind = has_sc ? 2 : 1;
    const uint32_t abuf_bytes = (uint32_t)(nkind * NP) * CHUNK_TILE, b_bytes = (uint32_t)Cout * Cin * 2;
    unsigned char *s_raw = smem_raw;                                           // RAW_BYTES (128-B aligned for TMA)
    unsigned char *sA = s_raw + RAW_BYTES;                                     // 2 chunk buffers x [kind][part] operand tiles
    unsigned char *sB = sA + 2 * abuf_bytes;                                   // [kind][part] weights, K-major [Cout x Cin]
    float *s_in = reinterpret_cast<float *>(sB + (size_t)(nkind * NP) * b_bytes);   // HVOX*CK
    float *s_dw = s_in + HVOX * CK;                                           // CK*27 (current chunk)
    float *s_scale = s_dw + CK * 27;                                          // Cin
    float *s_shift = s_scale + Cin;
    // {sum, sum of squares} of t and r, 4*Cout DOUBLES: every partial sum that is combined in a run-dependent order (the
    // shared-memory atomics of the warps, then the global ones) is double, so that the statistics -- and with them the stored
    // activations, the LeakyReLU / max-pool decisions of the backward pass and the gradients -- do not depend on the run
    // (fp32 here: forward tensors differed by 2e-6 between repetitions of one training step, individual gradient tensors by
    // 1e-3; tools/diag_bwd_determinism.py)
    double *s_stat = reinterpret_cast<double *>(s_shift + Cin + ((reinterpret_cast<uintptr_t>(s_shift + Cin) & 4) ? 1 : 0));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc::tmem_alloc(&s_tmem, (uint32_t)A.tmem_cols);
    if (tid == 32) { tc::mbar_init(&s_mma[0], 1); tc::mbar_init(&s_mma[1], 1); tc::mbar_init(&s_tma_bar, 1); }
    // stage the pointwise (+ shortcut) weights once as fp16 (hi [+ lo]) K-major operand tiles
    for (int kd = 0; kd < nkind; ++kd) {
        const float *src = kd == 0 ? A.pw_w : A.sc_w;
        unsigned char *dst = sB + (size_t)(kd * NP) * b_bytes;
        for (int i = tid; i < Cout * Cin; i += NT) {
            const int k = i % Cin, n = i / Cin;
            __half hi, lo;
            if (NP == 2) split_f16_scaled(src[i], SPLIT_WS, hi, lo); else split_f16(src[i], hi, lo);
            const uint32_t off = tc::tile_off(n, k, Cout);
            *reinterpret_cast<__half *>(dst + off) = hi;
            if (NP == 2) *reinterpret_cast<__half *>(dst + b_bytes + off) = lo;
        }
    }
    for (int i = tid; i < 4 * Cout; i += NT) s_stat[i] = 0.0;
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = tc::idesc_f16_m128(Cout);
    const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB);

    const int tilesX = (A.W + TX - 1) / TX, tilesY = (A.H + TY - 1) / TY, tilesZ = (A.D + TZ - 1) / TZ;
    const long long tiles_per_sample = (long long)tilesX * tilesY * tilesZ;
    const long long total_tiles = tiles_per_sample * A.N;
    const int nchunks = Cin / CK;
    // 32-bit arithmetic (the host refuses >= 2^30 tiles): 64-bit divisions are ~100 instructions each, per thread and tile
    const uint32_t tps32 = (uint32_t)tiles_per_sample;
    auto tile_coord = [&](long long tile, int &n, int &z0, int &y0, int &x0) {
        const uint32_t t32 = (uint32_t)tile;
        n = (int)(t32 / tps32);
        uint32_t b = t32 - (uint32_t)n * tps32;
        x0 = (int)(b % (uint32_t)tilesX) * TX; b /= (uint32_t)tilesX;
        y0 = (int)(b % (uint32_t)tilesY) * TY; b /= (uint32_t)tilesY;
        z0 = (int)b * TZ;
    };
    // activation-pass role: fixed set of 8-channel vectors of the raw box (same for every work item)
    uint32_t act_item[ACT_PER_THREAD];
#pragma unroll
    for (int k = 0; k < ACT_PER_THREAD; ++k) {
        const int item = tid + k * NT;
        int hv = item >> 1;
        const int q = item & 1;
        const int hx = hv % HX; hv /= HX;
        const int hy = hv % HY;
        const int hz = hv / HY;
        const uint32_t so = (uint32_t)((hz * HPLANE + hy * HXP + hx) * CK + q * 8);   // float index into s_in
        act_item[k] = item < ACT_ITEMS ? (so | ((uint32_t)hx << 16) | ((uint32_t)hy << 20) | ((uint32_t)hz << 24) | ((uint32_t)q << 28))
                                       : 0xffffffffu;
    }
    // stencil role: channel c of the chunk, rows (lz, ly0) and (lz, ly0+1)
    const int c = tid & 15, g = tid >> 4;
    const int lz = (g & 1) + 2 * (g >> 3);
    const int ly0 = 2 * ((g >> 1) & 3);
    // epilogue role: voxel row of MMA tile `em`
    const int em = warp >> 2, erow = (warp & 3) * 32 + lane;
    const int ev = em * 128 + erow;
    const int elx = ev & 7, ely = (ev >> 3) & 7, elz = ev >> 6;
    uint32_t tphase = 0;
    int cur_n = -1;
    int it = 0;                           // work items (tile, chunk) done by this CTA: operand buffer it & 1, its use number it >> 1

    auto flush_stats = [&](int n) {
        if (n < 0) return;
        for (int i = tid; i < 2 * Cout; i += NT) {
            const int isq = i >= Cout, cc = isq ? i - Cout : i;
            atomicAdd(&A.t_stats[(size_t)isq * A.N * Cout + (size_t)n * Cout + cc], s_stat[i]);
            s_stat[i] = 0.0;
            if (has_sc) {
                atomicAdd(&A.r_stats[(size_t)isq * A.N * Cout + (size_t)n * Cout + cc], s_stat[2 * Cout + i]);
                s_stat[2 * Cout + i] = 0.0;
            }
        }
    };

    float dwr0 = A.dw_w[tid], dwr1 = (tid + NT < CK * 27) ? A.dw_w[tid + NT] : 0.f;
    if (tid == 0 && (long long)blockIdx.x < total_tiles) {
        int n, z0, y0, x0;
        tile_coord(blockIdx.x, n, z0, y0, x0);
        tc::mbar_expect_tx(&s_tma_bar, RAW_BYTES);
        tc::tma_load_5d(s_raw, &tmap, &s_tma_bar, 0, x0 - 1, y0 - 1, z0 - 1, n);
    }

    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int n, z0, y0, x0;
        tile_coord(tile, n, z0, y0, x0);
        if (n != cur_n) {
            __syncthreads();            // epilogue atomics of the previous tile are in s_stat
            flush_stats(cur_n);
            cur_n = n;
            for (int cc = tid; cc < Cin; cc += NT) {
                float sc, sh;
                norm_scale_shift(A.xn, A.N, Cin, n, cc, sc, sh);
                s_scale[cc] = sc; s_shift[cc] = sh;
            }
            __syncthreads();
        }
        for (int ch = 0; ch < nchunks; ++ch, ++it) {
            const int c0 = ch * CK, buf = it & 1;
            unsigned char *Ab = sA + (size_t)buf * abuf_bytes;
            // ---- activation pass: raw box -> fp32 stencil tile
            // depthwise taps of this chunk were prefetched into registers one chunk ago
            s_dw[tid] = dwr0;
            if (tid + NT < CK * 27) s_dw[tid + NT] = dwr1;
            {
                const int nc0 = (ch + 1 < nchunks) ? c0 + CK : 0;
                dwr0 = A.dw_w[(size_t)nc0 * 27 + tid];
                if (tid + NT < CK * 27) dwr1 = A.dw_w[(size_t)nc0 * 27 + tid + NT];
            }
            tc::mbar_wait(&s_tma_bar, tphase);
            tphase ^= 1u;
#pragma unroll
            for (int k = 0; k < ACT_PER_THREAD; ++k) {
                const uint32_t ai = act_item[k];
                if (ai != 0xffffffffu) {
                    const int hx = (ai >> 16) & 15, hy = (ai >> 20) & 15, hz = (ai >> 24) & 15, q = (ai >> 28) & 1;
                    const int gz = z0 + hz - 1, gy = y0 + hy - 1, gx = x0 + hx - 1;
                    float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
                    if (gz >= 0 && gz < A.D && gy >= 0 && gy < A.H && gx >= 0 && gx < A.W) {
                        float f[8];
                        if (NP == 1) {
                            const uint4 rw = *reinterpret_cast<const uint4 *>(s_raw + (size_t)(tid + k * NT) * 16);
                            f[0] = h16_lo(rw.x); f[1] = h16_hi(rw.x); f[2] = h16_lo(rw.y); f[3] = h16_hi(rw.y);
                            f[4] = h16_lo(rw.z); f[5] = h16_hi(rw.z); f[6] = h16_lo(rw.w); f[7] = h16_hi(rw.w);
                        } else {
                            const float4 a = *reinterpret_cast<const float4 *>(s_raw + (size_t)(tid + k * NT) * 32);
                            const float4 b = *reinterpret_cast<const float4 *>(s_raw + (size_t)(tid + k * NT) * 32 + 16);
                            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
                        }
                        const float4 sc0 = *reinterpret_cast<const float4 *>(s_scale + c0 + q * 8);
                        const float4 sc1 = *reinterpret_cast<const float4 *>(s_scale + c0 + q * 8 + 4);
                        const float4 sh0 = *reinterpret_cast<const float4 *>(s_shift + c0 + q * 8);
                        const float4 sh1 = *reinterpret_cast<const float4 *>(s_shift + c0 + q * 8 + 4);
                        const float sl = A.xn.slope;
                        o0.x = lrelu(fmaf(f[0], sc0.x, sh0.x), sl); o0.y = lrelu(fmaf(f[1], sc0.y, sh0.y), sl);
                        o0.z = lrelu(fmaf(f[2], sc0.z, sh0.z), sl); o0.w = lrelu(fmaf(f[3], sc0.w, sh0.w), sl);
                        o1.x = lrelu(fmaf(f[4], sc1.x, sh1.x), sl); o1.y = lrelu(fmaf(f[5], sc1.y, sh1.y), sl);
                        o1.z = lrelu(fmaf(f[6], sc1.z, sh1.z), sl); o1.w = lrelu(fmaf(f[7], sc1.w, sh1.w), sl);
                    }
                    float *dst = s_in + (ai & 0xffffu);
                    *reinterpret_cast<float4 *>(dst) = o0;
                    *reinterpret_cast<float4 *>(dst + 4) = o1;
                }
            }
            __syncthreads();             // stencil tile complete, raw box consumed
            // ---- prefetch the next work item's raw box (overlaps the stencil / MMA / epilogue below)
            if (tid == 0) {
                int nn = n, nz = z0, ny = y0, nx = x0, nc = c0 + CK;
                bool more = true;
                if (ch + 1 == nchunks) {
                    const long long nt = tile + gridDim.x;
                    more = nt < total_tiles;
                    if (more) tile_coord(nt, nn, nz, ny, nx);
                    nc = 0;
                }
                if (more) {
                    tc::mbar_expect_tx(&s_tma_bar, RAW_BYTES);
                    tc::tma_load_5d(s_raw, &tmap, &s_tma_bar, nc, nx - 1, ny - 1, nz - 1, nn);
                }
            }
            // the MMAs that read operand buffer `buf` two work items ago must have completed before it is rewritten
            if (it >= 2) tc::mbar_wait(&s_mma[buf], (uint32_t)(((it >> 1) - 1) & 1));
            // ---- depthwise stencil: 2 rows x 8 voxels of channel c0+c
            {
                float wreg[27];
#pragma unroll
                for (int k = 0; k < 27; ++k) wreg[k] = s_dw[c * 27 + k];
                float acc0[TX], acc1[TX];
#pragma unroll
                for (int i = 0; i < TX; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }
                float ctr0[TX], ctr1[TX];
#pragma unroll
                for (int dz = 0; dz < 3; ++dz) {
#pragma unroll
                    for (int hy = 0; hy < 4; ++hy) {
                        const float *rowp = s_in + (size_t)((lz + dz) * HPLANE + (ly0 + hy) * HXP) * CK + c;
                        float row[HX];
#pragma unroll
                        for (int hx = 0; hx < HX; ++hx) row[hx] = rowp[hx * CK];
                        if (hy <= 2) {
                            const float w0 = wreg[dz * 9 + hy * 3], w1 = wreg[dz * 9 + hy * 3 + 1], w2 = wreg[dz * 9 + hy * 3 + 2];
#pragma unroll
                            for (int i = 0; i < TX; ++i) acc0[i] = fmaf(w2, row[i + 2], fmaf(w1, row[i + 1], fmaf(w0, row[i], acc0[i])));
                        }
                        if (hy >= 1) {
                            const float w0 = wreg[dz * 9 + (hy - 1) * 3], w1 = wreg[dz * 9 + (hy - 1) * 3 + 1], w2 = wreg[dz * 9 + (hy - 1) * 3 + 2];
#pragma unroll
                            for (int i = 0; i < TX; ++i) acc1[i] = fmaf(w2, row[i + 2], fmaf(w1, row[i + 1], fmaf(w0, row[i], acc1[i])));
                        }
                        if (dz == 1 && hy == 1) {
#pragma unroll
                            for (int i = 0; i < TX; ++i) ctr0[i] = row[i + 1];
                        }
                        if (dz == 1 && hy == 2) {
#pragma unroll
                            for (int i = 0; i < TX; ++i) ctr1[i] = row[i + 1];
                        }
                    }
                }
                // operand tiles of this chunk: [m][k/8][128 rows][8]; row = voxel within the 128-row MMA tile
                const int m = lz >> 1;
                const int r0 = ((lz & 1) * TY + ly0) * TX;          // row of (lz, ly0, lx = 0), a multiple of 8
                const uint32_t o0 = (uint32_t)m * 4096 + (uint32_t)(c >> 3) * 2048 + (uint32_t)(r0 >> 3) * 128 + (uint32_t)(c & 7) * 2, o1 = o0 + 128;
#pragma unroll
                for (int i = 0; i < TX; ++i) {
                    __half h0, l0, h1, l1;
                    if (NP == 2) { split_f16_scaled(acc0[i], SPLIT_AS, h0, l0); split_f16_scaled(acc1[i], SPLIT_AS, h1, l1); }
                    else { split_f16(acc0[i], h0, l0); split_f16(acc1[i], h1, l1); }
                    *reinterpret_cast<__half *>(Ab + o0 + i * 16) = h0;
                    *reinterpret_cast<__half *>(Ab + o1 + i * 16) = h1;
                    if (NP == 2) {
                        *reinterpret_cast<__half *>(Ab + CHUNK_TILE + o0 + i * 16) = l0;
                        *reinterpret_cast<__half *>(Ab + CHUNK_TILE + o1 + i * 16) = l1;
                    }
                }
                if (has_sc) {
                    unsigned char *As = Ab + NP * CHUNK_TILE;
#pragma unroll
                    for (int i = 0; i < TX; ++i) {
                        __half h0, l0, h1, l1;
                        if (NP == 2) { split_f16_scaled(ctr0[i], SPLIT_AS, h0, l0); split_f16_scaled(ctr1[i], SPLIT_AS, h1, l1); }
                        else { split_f16(ctr0[i], h0, l0); split_f16(ctr1[i], h1, l1); }
                        *reinterpret_cast<__half *>(As + o0 + i * 16) = h0;
                        *reinterpret_cast<__half *>(As + o1 + i * 16) = h1;
                        if (NP == 2) {
                            *reinterpret_cast<__half *>(As + CHUNK_TILE + o0 + i * 16) = l0;
                            *reinterpret_cast<__half *>(As + CHUNK_TILE + o1 + i * 16) = l1;
                        }
                    }
                }
            }
            tc::fence_async_smem();
            __syncthreads();             // operand chunk complete; s_in / s_dw free for the next chunk
            if (tid == 0) {
                tc::fence_after_sync();
                const uint32_t a_u = sA_u + (uint32_t)buf * abuf_bytes;
#pragma unroll
                for (int kd = 0; kd < 2; ++kd) {
                    if (kd < nkind) {
#pragma unroll
                        for (int m = 0; m < MT; ++m) {
                            const uint32_t d = tmem + (uint32_t)((kd * MT + m) * Cout);
                            const uint64_t ah = tc::smem_desc(a_u + (uint32_t)(kd * NP) * CHUNK_TILE + m * 4096, 2048, 128);
                            const uint64_t bh = tc::smem_desc(sB_u + (uint32_t)(kd * NP) * b_bytes + 2 * ch * Cout * 16, Cout * 16, 128);
                            tc::mma_f16(d, ah, bh, idesc, ch > 0 ? 1u : 0u);
                            if (NP == 2) {
                                const uint64_t al = tc::smem_desc(a_u + (uint32_t)(kd * NP + 1) * CHUNK_TILE + m * 4096, 2048, 128);
                                const uint64_t bl = tc::smem_desc(sB_u + (uint32_t)(kd * NP + 1) * b_bytes + 2 * ch * Cout * 16, Cout * 16, 128);
                                tc::mma_f16(d, al, bh, idesc, 1u);
                                tc::mma_f16(d, ah, bl, idesc, 1u);
                            }
                        }
                    }
                }
                tc::mma_commit(&s_mma[buf]);
            }
            // ---- optional save of the depthwise output for the backward pass, straight from the operand tile of this chunk
            if (A.u != nullptr) {
#pragma unroll
                for (int k = 0; k < MT * 128 * 2 / NT; ++k) {
                    const int item = tid + k * NT, q = item & 1, v = item >> 1;
                    const int lx = v & 7, ly = (v >> 3) & 7, lzz = v >> 6;
                    const int gz = z0 + lzz, gy = y0 + ly, gx = x0 + lx;
                    if (gz < A.D && gy < A.H && gx < A.W) {
                        const uint32_t off = (uint32_t)(v >> 7) * 4096 + (uint32_t)q * 2048 + (uint32_t)((v & 127) >> 3) * 128 + (uint32_t)(v & 7) * 16;
                        const uint4 h = *reinterpret_cast<const uint4 *>(Ab + off);
                        T *up = reinterpret_cast<T *>(A.u) + ((((size_t)n * A.D + gz) * A.H + gy) * A.W + gx) * (size_t)A.ldu + c0 + q * 8;
                        if (NP == 1) {
                            *reinterpret_cast<uint4 *>(up) = h;               // the fp16 operand is the stored value
                        } else {
                            const uint4 l = *reinterpret_cast<const uint4 *>(Ab + CHUNK_TILE + off);
                            float4 a, b;
                            a.x = h16_lo(h.x) + h16_lo(l.x); a.y = h16_hi(h.x) + h16_hi(l.x); a.z = h16_lo(h.y) + h16_lo(l.y); a.w = h16_hi(h.y) + h16_hi(l.y);
                            b.x = h16_lo(h.z) + h16_lo(l.z); b.y = h16_hi(h.z) + h16_hi(l.z); b.z = h16_lo(h.w) + h16_lo(l.w); b.w = h16_hi(h.w) + h16_hi(l.w);
                            constexpr float ia = 1.f / SPLIT_AS;        // the operand tile holds the scaled values
                            a.x *= ia; a.y *= ia; a.z *= ia; a.w *= ia; b.x *= ia; b.y *= ia; b.z *= ia; b.w *= ia;
                            reinterpret_cast<float4 *>(up)[0] = a;             // hi + lo: the value the pointwise GEMM multiplied
                            reinterpret_cast<float4 *>(up)[1] = b;
                        }
                    }
                }
            }
        }
        // ---- epilogue: TMEM -> global + statistics (the commit of the last chunk covers every MMA of the tile)
        tc::mbar_wait(&s_mma[(it - 1) & 1], (uint32_t)(((it - 1) >> 1) & 1));
        tc::fence_after_sync();
        {
            const int gz = z0 + elz, gy = y0 + ely, gx = x0 + elx;
            const bool valid = gz < A.D && gy < A.H && gx < A.W;
            const size_t vox = (((size_t)n * A.D + gz) * A.H + gy) * A.W + gx;
            const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            const bool wide_st = A.wide_st != 0;
            for (int a = 0; a < nkind; ++a) {
                T *outp = a == 0 ? reinterpret_cast<T *>(A.t) + vox * (size_t)A.ldt : reinterpret_cast<T *>(A.r) + vox * (size_t)A.ldr;
                double *stat = s_stat + a * 2 * Cout;
                for (int cb = 0; cb < Cout; cb += 16) {
                    float v[16];
                    tc::tmem_ld16(trow + (uint32_t)((a * MT + em) * Cout + cb), v);
                    if (NP == 2) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] *= SPLIT_INV;
                    }
                    if (wide_st) store16_wide(outp + cb, v, valid); else store16(outp + cb, v, valid);
                    float sv[32];
#pragma unroll
                    for (int j = 0; j < 16; ++j) { sv[j] = v[j]; sv[16 + j] = v[j] * v[j]; }
                    warp_transpose_sum<32>(sv, lane);
                    const int idx = warp_transpose_owner<32>(lane);      // 0..15 sums, 16..31 squares
                    atomicAdd(&stat[(idx >= 16 ? Cout + idx - 16 : idx) + cb], (double)sv[0]);
                }
            }
        }
        tc::fence_before_sync();
        __syncthreads();                 // TMEM drained, statistics in s_stat
    }
    flush_stats(cur_n);
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)A.tmem_cols);
}

static size_t tc_smem_bytes(int Cin, int Cout, bool has_sc, int NP, int ES) {
    const size_t nk = (has_sc ? 2 : 1) * (size_t)NP;
    return (size_t)HZ * HY * HX * CK * ES + 2 * nk * CHUNK_TILE + nk * (size_t)Cout * Cin * 2 +
           sizeof(float) * ((size_t)HVOX * CK + 27 * CK + 2 * (size_t)Cin + 2 + 8 * (size_t)Cout);   /* statistics: 4 x Cout doubles (+ alignment) */
}

}  // namespace

// Returns -1 when the tensor-core path does not apply (the caller falls back to the generic kernel).
int l3d_dwpw_fwd_tc(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                    const float *dw_w, const float *pw_w, const float *sc_w,
                    const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats,
                    const l3d_act *u, void *stream) {
    if (L3D_ENV_INT("L3D_NO_TC", 0) == 1) return -1;
    const int Cin = x->C, Cout = t->C;
    const bool has_sc = sc_w != nullptr, has_u = !act_null(u);
    const bool f32 = x->dtype == L3D_F32;
    if (dw_w == nullptr || t->dtype != x->dtype || (has_sc && r->dtype != x->dtype) || (has_u && u->dtype != x->dtype)) return -1;
    if (Cin % 16 != 0 || Cout % 16 != 0 || Cout > 256) return -1;
    const int cols_needed = MT * Cout * (has_sc ? 2 : 1);
    if (cols_needed > 512) return -1;
    const int es = f32 ? 4 : 2;
    const size_t smem = tc_smem_bytes(Cin, Cout, has_sc, f32 ? 2 : 1, es);
    if (smem > 226 * 1024) return -1;
    const int vec = 16 / es;            // elements per 16 bytes: TMA needs a 16-byte aligned base and strides, the stores 16-byte vectors
    auto aligned = [vec](const l3d_act *a) { return (a->ldc % vec == 0) && (reinterpret_cast<uintptr_t>(a->ptr) % 16 == 0); };
    if (!aligned(x) || !aligned(t) || (has_sc && !aligned(r)) || (has_u && !aligned(u))) return -1;
    if ((long long)N * D * H * W * x->ldc * es >= (1ll << 40)) return -1;
    int cols = 32;
    while (cols < cols_needed) cols <<= 1;

    // 5-D tensor map over the channels-last input view: (C, W, H, D, N), box (16, 10, 10, 6, 1)
    CUtensorMap tmap;
    {
        const cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        const cuuint64_t ld = (cuuint64_t)x->ldc;
        const cuuint64_t strides[4] = {ld * es, (cuuint64_t)W * ld * es, (cuuint64_t)H * W * ld * es, (cuuint64_t)D * H * W * ld * es};
        const cuuint32_t box[5] = {CK, HX, HY, HZ, 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        if (l3d_encode_tiled(&tmap, (int)(f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16), 5, x->ptr,
                             (const unsigned long long *)dims, (const unsigned long long *)strides, (const unsigned *)box, (const unsigned *)estr)) return 3;
    }
    TcArgs A;
    A.Cin = Cin; A.xn = norm_dev(xn);
    A.N = N; A.D = D; A.H = H; A.W = W;
    A.dw_w = dw_w; A.pw_w = pw_w; A.sc_w = sc_w; A.Cout = Cout;
    A.t = t->ptr; A.ldt = t->ldc; A.t_stats = t_stats;
    A.r = has_sc ? r->ptr : nullptr; A.ldr = has_sc ? r->ldc : 0; A.r_stats = r_stats;
    A.u = has_u ? u->ptr : nullptr; A.ldu = has_u ? u->ldc : 0;
    A.tmem_cols = cols;
    {
        // 16 output channels of a voxel = 32 (fp16) / 64 (fp32) bytes: 256-bit stores when every block starts on a 32-byte boundary
        const int per32 = f32 ? 8 : 16;
        auto al32 = [per32](const l3d_act *a) { return a->ldc % per32 == 0 && reinterpret_cast<uintptr_t>(a->ptr) % 32 == 0; };
        A.wide_st = (al32(t) && (sc_w == nullptr || al32(r)) && L3D_ENV_INT("L3D_ST256", 1) != 0) ? 1 : 0;
    }
    {
        cudaError_t e = f32 ? cudaFuncSetAttribute(dwpw_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)
                            : cudaFuncSetAttribute(dwpw_tc_kernel<h16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) { l3d_set_error("dwpw_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 3; }
    }
    int occ = (int)((227 * 1024) / (smem + 2048));
    if (occ > 2) occ = 2;                        // 256 threads x ~120 registers
    if (occ < 1) occ = 1;
    if (occ * cols > 512) occ = 512 / cols;      // TMEM columns are a per-SM resource
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (long long)N * ((D + TZ - 1) / TZ) * ((H + TY - 1) / TY) * ((W + TX - 1) / TX);
    if (tiles >= (1ll << 30)) return -1;
    long long grid = (long long)sms * occ;
    if (grid > tiles) grid = tiles;
    if (f32) dwpw_tc_kernel<float><<<(unsigned)grid, NT, smem, (cudaStream_t)stream>>>(tmap, A);
    else dwpw_tc_kernel<h16><<<(unsigned)grid, NT, smem, (cudaStream_t)stream>>>(tmap, A);
    l3d_count_launch();
    l3d_note_kernel("dwpw_tc_kernel");
    L3D_CUDA_OK("l3d_dwpw_fwd (tcgen05) launch");
    return 0;
}

// =============================================================================================
// ConvTranspose3d(k=2, s=2) + bias on tensor cores (bf16 storage).  Non-overlapping taps make it one GEMM per
// 128 input voxels: D[128][8*Cout] = X[128][Cin] . Wt[8*Cout][Cin]^T (row tap*Cout+co of Wt = W[:, co, tap]),
// followed by a 2x2x2 pixel shuffle into the lower channel half of the skip-concat buffer (unet3d.py:127-141).
namespace {

struct CtArgs {
    const void *x; int ldx; int Cin;
    int N, d, h, w;
    const float *wgt, *bias; int Cout;
    void *out; int ldo; int OD, OH, OW, oz, oy, ox;
    int tmem_cols, wide_st;
};

// eight fp32 values -> fp16 hi / lo vectors
__device__ __forceinline__ void split8_f16(const float4 &a, const float4 &b, uint4 &hi, uint4 &lo) {
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __half h0, l0, h1, l1;
        split_f16_scaled(f[2 * j], SPLIT_AS, h0, l0); split_f16_scaled(f[2 * j + 1], SPLIT_AS, h1, l1);       // (fp32 storage only)
        h[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        l[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]); lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <typename T>
__global__ void __launch_bounds__(NT) convt_tc_kernel(CtArgs A) {
    constexpr int NP = St<T>::NP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int Cin = A.Cin, Cout = A.Cout, NB = 8 * Cout;
    const uint32_t a_bytes = 128u * Cin * 2, b_bytes = (uint32_t)NB * Cin * 2;
    unsigned char *sB = smem_raw;                         // NP parts of NB * Cin * 2
    unsigned char *sA = sB + (size_t)NP * b_bytes;        // 2 buffers x NP parts of a_bytes
    float *s_bias = reinterpret_cast<float *>(sA + 2 * NP * a_bytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc::tmem_alloc(&s_tmem, (uint32_t)A.tmem_cols);
    if (tid == 32) tc::mbar_init(&s_bar, 1);
    for (int i = tid; i < Cin * Cout * 8; i += NT) {      // wgt[ci][co][tap]
        const int tap = i & 7, co = (i >> 3) % Cout, ci = (i >> 3) / Cout;
        __half hi, lo;
        if (NP == 2) split_f16_scaled(A.wgt[i], SPLIT_WS, hi, lo); else split_f16(A.wgt[i], hi, lo);
        const uint32_t off = tc::tile_off(tap * Cout + co, ci, NB);
        *reinterpret_cast<__half *>(sB + off) = hi;
        if (NP == 2) *reinterpret_cast<__half *>(sB + b_bytes + off) = lo;
    }
    for (int i = tid; i < Cout; i += NT) s_bias[i] = A.bias[i];
    const long long nvox = (long long)A.N * A.d * A.h * A.w;
    const long long ntiles = (nvox + 127) / 128;
    const int kq = Cin >> 3;
    const int kq_sh = __ffs(kq) - 1;                                   // Cin / 8 is a power of two (host check): shifts instead of divisions
    const T *xin = reinterpret_cast<const T *>(A.x);
    auto load_a = [&](long long tile, int buf) {
        unsigned char *dst = sA + (size_t)buf * NP * a_bytes;
        const long long v0 = tile * 128;
        for (int item = tid; item < 128 * kq; item += NT) {
            const int q = item & (kq - 1), v = item >> kq_sh;
            uint4 o = make_uint4(0u, 0u, 0u, 0u), ol = o;
            if (v0 + v < nvox) {
                const T *src = xin + (size_t)(v0 + v) * A.ldx + q * 8;
                if (NP == 1) {
                    o = *reinterpret_cast<const uint4 *>(src);   // stored fp16 = MMA operand
                } else {
                    split8_f16(reinterpret_cast<const float4 *>(src)[0], reinterpret_cast<const float4 *>(src)[1], o, ol);
                }
            }
            const uint32_t off = (uint32_t)q * 2048 + (uint32_t)(v >> 3) * 128 + (uint32_t)(v & 7) * 16;
            *reinterpret_cast<uint4 *>(dst + off) = o;
            if (NP == 2) *reinterpret_cast<uint4 *>(dst + a_bytes + off) = ol;
        }
    };
    if ((long long)blockIdx.x < ntiles) load_a(blockIdx.x, 0);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const int n_mma = NB > 256 ? NB / 256 : 1, n_each = NB > 256 ? 256 : NB;
    const uint32_t idesc = tc::idesc_f16_m128(n_each);
    const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB);
    uint32_t phase = 0;
    int buf = 0;
    const int erow = (warp & 3) * 32 + lane, tap0 = (warp >> 2) * 4;
    T *outp = reinterpret_cast<T *>(A.out);
    const bool wide_st = A.wide_st != 0;                 // 32-byte aligned 16-channel blocks (host check)
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        if (tid == 0) {
            tc::fence_after_sync();
            for (int j = 0; j < Cin / 16; ++j)
                for (int hN = 0; hN < n_mma; ++hN) {
                    const uint32_t a0 = sA_u + (uint32_t)buf * NP * a_bytes + 2 * j * 2048, b0 = sB_u + 2 * j * NB * 16 + hN * (256 / 8) * 128;
                    const uint64_t ah = tc::smem_desc(a0, 2048, 128), bh = tc::smem_desc(b0, NB * 16, 128);
                    tc::mma_f16(tmem + hN * 256, ah, bh, idesc, j > 0);
                    if (NP == 2) {
                        tc::mma_f16(tmem + hN * 256, tc::smem_desc(a0 + a_bytes, 2048, 128), bh, idesc, 1u);
                        tc::mma_f16(tmem + hN * 256, ah, tc::smem_desc(b0 + b_bytes, NB * 16, 128), idesc, 1u);
                    }
                }
            tc::mma_commit(&s_bar);
        }
        const long long nt = tile + gridDim.x;
        if (nt < ntiles) load_a(nt, buf ^ 1);        // overlaps the MMAs of this tile
        tc::mbar_wait(&s_bar, phase);
        phase ^= 1u;
        tc::fence_after_sync();
        {
            // tcgen05.ld is warp-collective (.sync.aligned): every lane runs the same loads, only the stores are predicated
            const long long gv = tile * 128 + erow;
            const bool row_ok = gv < nvox;
            uint32_t rem = row_ok ? (uint32_t)gv : 0u;      // nvox < 2^31 (host check)
            const int ix = (int)(rem % (uint32_t)A.w); rem /= (uint32_t)A.w;
            const int iy = (int)(rem % (uint32_t)A.h); rem /= (uint32_t)A.h;
            const int iz = (int)(rem % (uint32_t)A.d);
            const int n = (int)(rem / (uint32_t)A.d);
            const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            for (int tp = tap0; tp < tap0 + 4; ++tp) {
                const int Z = A.oz + 2 * iz + (tp >> 2), Y = A.oy + 2 * iy + ((tp >> 1) & 1), X = A.ox + 2 * ix + (tp & 1);
                const bool ok = row_ok && Z >= 0 && Z < A.OD && Y >= 0 && Y < A.OH && X >= 0 && X < A.OW;
                T *op = outp + ((((size_t)n * A.OD + Z) * A.OH + Y) * A.OW + X) * (size_t)A.ldo;
                for (int cb = 0; cb < Cout; cb += 16) {
                    float v[16];
                    tc::tmem_ld16(trow + (uint32_t)(tp * Cout + cb), v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = (NP == 2 ? v[j] * SPLIT_INV : v[j]) + s_bias[cb + j];
                    if (wide_st) store16_256(op + cb, v, ok); else store16(op + cb, v, ok);
                }
            }
        }
        tc::fence_async_smem();          // next tile's operand stores (issued above) -> async proxy
        tc::fence_before_sync();
        __syncthreads();
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)A.tmem_cols);
}

}  // namespace

int l3d_convt_fwd_tc(const l3d_act *x, int N, int d, int h, int w_, const float *w, const float *b,
                     const l3d_act *out, int OD, int OH, int OW, int oz, int oy, int ox, void *stream) {
    if (L3D_ENV_INT("L3D_NO_TC", 0) == 1) return -1;
    const int Cin = x->C, Cout = out->C;
    const bool f32 = x->dtype == L3D_F32;
    const int np = f32 ? 2 : 1, es = f32 ? 4 : 2;
    if (out->dtype != x->dtype || Cin % 16 != 0 || Cout % 16 != 0 || 8 * Cout > 512) return -1;
    if ((Cin & (Cin - 1)) != 0) return -1;           // the A-tile loader decodes (voxel, channel group) with shifts
    if (8 * Cout > 256 && (8 * Cout) % 256 != 0) return -1;
    const int vec = 16 / es;
    auto aligned = [vec](const l3d_act *a) { return (a->ldc % vec == 0) && (reinterpret_cast<uintptr_t>(a->ptr) % 16 == 0); };
    if (!aligned(x) || !aligned(out)) return -1;
    const size_t smem = (size_t)np * 8 * Cout * Cin * 2 + 2 * (size_t)np * 128 * Cin * 2 + sizeof(float) * Cout;
    if (smem > 226 * 1024) return -1;           // fp32 storage, 128 -> 64 (1728 voxels per 48^3 patch): generic kernel
    int cols = 32;
    while (cols < 8 * Cout) cols <<= 1;
    CtArgs A;
    A.x = x->ptr; A.ldx = x->ldc; A.Cin = Cin;
    A.N = N; A.d = d; A.h = h; A.w = w_;
    A.wgt = w; A.bias = b; A.Cout = Cout;
    A.out = out->ptr; A.ldo = out->ldc; A.OD = OD; A.OH = OH; A.OW = OW; A.oz = oz; A.oy = oy; A.ox = ox;
    A.tmem_cols = cols;
    A.wide_st = (out->ldc % (f32 ? 8 : 16) == 0 && reinterpret_cast<uintptr_t>(out->ptr) % 32 == 0 && L3D_ENV_INT("L3D_ST256", 1) != 0) ? 1 : 0;
    {
        cudaError_t e = f32 ? cudaFuncSetAttribute(convt_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)
                            : cudaFuncSetAttribute(convt_tc_kernel<h16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) { l3d_set_error("convt_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 3; }
    }
    int occ = (int)((227 * 1024) / (smem + 2048));
    if (occ > 4) occ = 4;                        // 56 registers; the TMEM columns (8 * Cout per CTA) bound it below
    if (occ < 1) occ = 1;
    if (occ * cols > 512) occ = 512 / cols;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = ((long long)N * d * h * w_ + 127) / 128;
    if (tiles >= (1ll << 24)) return -1;             // 32-bit voxel indices in the kernel
    long long grid = (long long)sms * occ;
    if (grid > tiles) grid = tiles;
    if (f32) convt_tc_kernel<float><<<(unsigned)grid, NT, smem, (cudaStream_t)stream>>>(A);
    else convt_tc_kernel<h16><<<(unsigned)grid, NT, smem, (cudaStream_t)stream>>>(A);
    l3d_count_launch();
    l3d_note_kernel("convt_tc_kernel");
    L3D_CUDA_OK("l3d_convt_fwd (tcgen05) launch");
    return 0;
}
