// Tensor-core backward of the pointwise (1x1x1) stage for bf16 storage (unet3d.py:18, 70-72): both GEMMs of the stage
// run on tcgen05 from the SAME two staged tiles.
//
//   stage   g_t[v][c] = a_c*gz[v][c] + b_c*t[v][c] + d_c   (InstanceNorm backward on load, fp32) and the activated
//           u[v][k], for 128 voxels, as bf16 in the voxel-planar layout [channel/8][128 voxels][8 channels];
//           g_t is split into hi + lo bf16 parts (the weights too), so the products carry ~16 mantissa bits; the stored
//           fp16 activation u (11-bit significand) is split EXACTLY into bf16 hi + lo (a tcgen05.mma.kind::f16 whose A and
//           B formats differ -- bf16 gradient x fp16 activation -- raises "illegal instruction" on sm_100a, measured)
//   dgrad   D1[v][k] = sum_c g_t[v][c] * W[c][k]      A = G tile read K-major  (rows = voxels, K = channels)
//   wgrad   D2[c][k] += sum_v g_t[v][c] * u[v][k]     A = G tile read MN-major (M = channels, K = voxels),
//                                                     B = U tile read MN-major (N = channels, K = voxels)
//           D2 stays in TMEM over all tiles of the CTA and is flushed once with one atomic per weight.
//
// The MN-major reading of the planar tile (LBO = 128 B: next 8 voxels, SBO = 2048 B: next 8 channels) was validated
// with l3d_tc_selftest_mn16; kind::tf32 accepts only K-major operands on this part (MN-major yields zeros), which is
// why the operands are bf16 hi/lo pairs rather than tf32.
#include "l3d_common.cuh"
#include "l3d_tc.cuh"

namespace {


constexpr int NT = 256, TV = 128;
constexpr int PLANE = TV * 16;          // bytes of one 8-channel group of a tile

struct PwTcArgs {
    const float *gz; int ldg;
    const void *t; int ldt; NormDev nt; const double *red;      // stored activations: fp16 or fp32 (template T)
    const void *u; int ldu; NormDev un;
    int N; long long vox;
    int Cg, Cu;
    const float *w; float *g_w;
    float *g_u; int ldgu; int accumulate;
    int tmem_cols;
};

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}
// hi = bf16(x), lo = bf16(x - hi): two values at a time
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
    hi = pack2_bf16(a, b);
    lo = pack2_bf16(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}
// eight stored fp16 values -> bf16 hi / lo vectors (exact: 11 significand bits fit 8 + 8)
__device__ __forceinline__ void split_h16x8(const uint4 &r, uint4 &hi, uint4 &lo) {
    split2(h16_lo(r.x), h16_hi(r.x), hi.x, lo.x); split2(h16_lo(r.y), h16_hi(r.y), hi.y, lo.y);
    split2(h16_lo(r.z), h16_hi(r.z), hi.z, lo.z); split2(h16_lo(r.w), h16_hi(r.w), hi.w, lo.w);
}
// eight consecutive stored channels of one voxel, fp16 (16 bytes) or fp32 (32 bytes)
template <typename T> struct V8;
template <> struct V8<h16> {
    uint4 r;
    __device__ __forceinline__ void load(const h16 *p) { r = *reinterpret_cast<const uint4 *>(p); }
    __device__ __forceinline__ void unpack(float (&f)[8]) const {
        f[0] = h16_lo(r.x); f[1] = h16_hi(r.x); f[2] = h16_lo(r.y); f[3] = h16_hi(r.y);
        f[4] = h16_lo(r.z); f[5] = h16_hi(r.z); f[6] = h16_lo(r.w); f[7] = h16_hi(r.w);
    }
};
template <> struct V8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float *p) { a = reinterpret_cast<const float4 *>(p)[0]; b = reinterpret_cast<const float4 *>(p)[1]; }
    __device__ __forceinline__ void unpack(float (&f)[8]) const {
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};
// eight values -> bf16 hi / lo planes (exact for stored fp16 values, 16 significand bits for fp32)
__device__ __forceinline__ void store_split8(unsigned char *sh, unsigned char *sl, size_t off, const float (&f)[8], bool ok) {
    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
    if (ok) {
        split2(f[0], f[1], hi.x, lo.x); split2(f[2], f[3], hi.y, lo.y);
        split2(f[4], f[5], hi.z, lo.z); split2(f[6], f[7], hi.w, lo.w);
    }
    *reinterpret_cast<uint4 *>(sh + off) = hi;
    *reinterpret_cast<uint4 *>(sl + off) = lo;
}
template <typename T>
__device__ __forceinline__ void store_u_split(unsigned char *sUh, unsigned char *sUl, size_t off, const V8<T> &r, bool ok) {
    float f[8];
    r.unpack(f);
    store_split8(sUh, sUl, off, f, ok);
}

// GI / UI: G / U staging items (one 8-channel group of one voxel) per thread, held in registers one tile ahead so the
// global loads of tile T+1 are in flight during the MMAs and the epilogue of tile T.  GI == 0: no prefetch (wide layers).
template <typename T, int GI, int UI>
__global__ void __launch_bounds__(NT) pw_bwd_tc_kernel(PwTcArgs A) {
    const T *At = reinterpret_cast<const T *>(A.t), *Au = reinterpret_cast<const T *>(A.u);
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int Cg = A.Cg, Cu = A.Cu, gq = Cg >> 3, uq = Cu >> 3;
    unsigned char *sGh = smem;                              // gq planes
    unsigned char *sGl = sGh + (size_t)gq * PLANE;
    unsigned char *sUh = sGl + (size_t)gq * PLANE;          // uq planes (hi), uq planes (lo)
    unsigned char *sUl = sUh + (size_t)uq * PLANE;
    unsigned char *sWh = sUl + (size_t)uq * PLANE;          // dgrad B operand: [N = Cu][K = Cg] K-major
    unsigned char *sWl = sWh + (size_t)Cg * Cu * 2;
    double *s_ca = reinterpret_cast<double *>(sWl + (size_t)Cg * Cu * 2);       // a, b, d of the InstanceNorm backward, in double (in_bwd_apply)
    double *s_cb = s_ca + Cg, *s_cd = s_cb + Cg;
    float *s_us = reinterpret_cast<float *>(s_cd + Cg), *s_uh = s_us + Cu;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool has_gu = A.g_u != nullptr, has_gw = A.g_w != nullptr;
    if (warp == 0) tc::tmem_alloc(&s_tmem, (uint32_t)A.tmem_cols);
    if (tid == 32) tc::mbar_init(&s_bar, 1);
    for (int i = tid; i < Cg * Cu; i += NT) {               // w[c][k]
        const int k = i % Cu, c = i / Cu;
        const float wv = A.w[i];
        const __nv_bfloat16 hi = __float2bfloat16_rn(wv);
        const uint32_t off = tc::tile_off(k, c, Cu);
        *reinterpret_cast<__nv_bfloat16 *>(sWh + off) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(sWl + off) = __float2bfloat16_rn(wv - __bfloat162float(hi));
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t d1 = tmem, d2 = tmem + (uint32_t)Cu;
    const uint32_t id_k = tc::idesc_16b_m128(Cu, 1, 1, false, false), id_mn = tc::idesc_16b_m128(Cu, 1, 1, true, true);
    const uint32_t sGh_u = tc::smem_u32(sGh), sGl_u = tc::smem_u32(sGl), sUh_u = tc::smem_u32(sUh), sUl_u = tc::smem_u32(sUl), sWh_u = tc::smem_u32(sWh), sWl_u = tc::smem_u32(sWl);

    const long long tiles_per_sample = (A.vox + TV - 1) / TV;
    const uint32_t tps32 = (uint32_t)tiles_per_sample;
    const long long total_tiles = tiles_per_sample * A.N;
    const bool has_nt = A.nt.stats != nullptr, u_ident = A.un.stats == nullptr;
    int cur_n = -1;
    uint32_t phase = 0;
    bool first = true;
    float4 pg0[GI > 0 ? GI : 1], pg1[GI > 0 ? GI : 1];
    V8<T> pt[GI > 0 ? GI : 1], pu[UI > 0 ? UI : 1];
    auto prefetch = [&](long long tl) {
        const int pn = (int)((uint32_t)tl / tps32);                                  // 32-bit: the host checks tiles < 2^31
        const long long pv0 = (long long)((uint32_t)tl - (uint32_t)pn * tps32) * TV;
#pragma unroll
        for (int i = 0; i < GI; ++i) {
            const int item = tid + i * NT, v = item & (TV - 1), q = item >> 7;
            if (pv0 + v < A.vox) {
                const size_t gv = (size_t)pn * A.vox + pv0 + v;
                pg0[i] = *reinterpret_cast<const float4 *>(A.gz + gv * (size_t)A.ldg + q * 8);
                pg1[i] = *reinterpret_cast<const float4 *>(A.gz + gv * (size_t)A.ldg + q * 8 + 4);
                if (has_nt) pt[i].load(At + gv * (size_t)A.ldt + q * 8);
            }
        }
        if (has_gw) {
#pragma unroll
            for (int i = 0; i < UI; ++i) {
                const int item = tid + i * NT, v = item & (TV - 1), q = item >> 7;
                if (pv0 + v < A.vox) pu[i].load(Au + ((size_t)pn * A.vox + pv0 + v) * (size_t)A.ldu + q * 8);
            }
        }
    };
    if (GI > 0 && (long long)blockIdx.x < total_tiles) prefetch(blockIdx.x);
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = (int)((uint32_t)tile / tps32);
        const long long v0 = (long long)((uint32_t)tile - (uint32_t)n * tps32) * TV;
        if (n != cur_n) {
            cur_n = n;
            for (int c = tid; c < Cg; c += NT) {
                double a, b, d;
                in_bwd_coef_d(A.nt, A.red, A.N, Cg, n, c, a, b, d);
                s_ca[c] = a; s_cb[c] = b; s_cd[c] = d;
            }
            for (int k = tid; k < Cu; k += NT) {
                float sc, sh;
                norm_scale_shift(A.un, A.N, Cu, n, k, sc, sh);
                s_us[k] = sc; s_uh[k] = sh;
            }
            __syncthreads();
        }
        if (GI > 0) {
            // ---- registers -> shared memory (the loads were issued one tile ago)
#pragma unroll
            for (int i = 0; i < GI; ++i) {
                const int item = tid + i * NT, v = item & (TV - 1), q = item >> 7;
                uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
                if (v0 + v < A.vox) {
                    float g[8] = {pg0[i].x, pg0[i].y, pg0[i].z, pg0[i].w, pg1[i].x, pg1[i].y, pg1[i].z, pg1[i].w};
                    if (has_nt) {
                        float tv[8];
                        pt[i].unpack(tv);
#pragma unroll
                        for (int j = 0; j < 8; ++j) g[j] = in_bwd_apply(s_ca[q * 8 + j], s_cb[q * 8 + j], s_cd[q * 8 + j], g[j], tv[j]);
                    }
                    split2(g[0], g[1], hi.x, lo.x); split2(g[2], g[3], hi.y, lo.y);
                    split2(g[4], g[5], hi.z, lo.z); split2(g[6], g[7], hi.w, lo.w);
                }
                *reinterpret_cast<uint4 *>(sGh + (size_t)q * PLANE + (size_t)v * 16) = hi;
                *reinterpret_cast<uint4 *>(sGl + (size_t)q * PLANE + (size_t)v * 16) = lo;
            }
            if (has_gw) {
#pragma unroll
                for (int i = 0; i < UI; ++i) {
                    const int item = tid + i * NT, v = item & (TV - 1), q = item >> 7;
                    store_u_split(sUh, sUl, (size_t)q * PLANE + (size_t)v * 16, pu[i], v0 + v < A.vox);
                }
            }
            // ---- issue the loads of this CTA's next tile
            if (tile + gridDim.x < total_tiles) prefetch(tile + gridDim.x);
        } else {
            // ---- stage G (hi / lo): item = (8-channel group q, voxel v), consecutive threads = consecutive voxels
            for (int item = tid; item < gq * TV; item += NT) {
                const int v = item & (TV - 1), q = item >> 7;
                uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
                if (v0 + v < A.vox) {
                    const size_t gv = (size_t)n * A.vox + v0 + v;
                    const float4 g0 = *reinterpret_cast<const float4 *>(A.gz + gv * (size_t)A.ldg + q * 8);
                    const float4 g1 = *reinterpret_cast<const float4 *>(A.gz + gv * (size_t)A.ldg + q * 8 + 4);
                    float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                    if (has_nt) {
                        V8<T> tr;
                        tr.load(At + gv * (size_t)A.ldt + q * 8);
                        float tv[8];
                        tr.unpack(tv);
#pragma unroll
                        for (int j = 0; j < 8; ++j) g[j] = in_bwd_apply(s_ca[q * 8 + j], s_cb[q * 8 + j], s_cd[q * 8 + j], g[j], tv[j]);
                    }
                    split2(g[0], g[1], hi.x, lo.x); split2(g[2], g[3], hi.y, lo.y);
                    split2(g[4], g[5], hi.z, lo.z); split2(g[6], g[7], hi.w, lo.w);
                }
                *reinterpret_cast<uint4 *>(sGh + (size_t)q * PLANE + (size_t)v * 16) = hi;
                *reinterpret_cast<uint4 *>(sGl + (size_t)q * PLANE + (size_t)v * 16) = lo;
            }
            // ---- stage U (activated; bf16 hi / lo of the fp16 value)
            if (has_gw) {
                for (int item = tid; item < uq * TV; item += NT) {
                    const int v = item & (TV - 1), q = item >> 7;
                    float uf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (v0 + v < A.vox) {
                        const size_t gv = (size_t)n * A.vox + v0 + v;
                        V8<T> o;
                        o.load(Au + gv * (size_t)A.ldu + q * 8);
                        o.unpack(uf);
                        if (!u_ident) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) uf[j] = lrelu(fmaf(uf[j], s_us[q * 8 + j], s_uh[q * 8 + j]), A.un.slope);
                        }
                    }
                    store_split8(sUh, sUl, (size_t)q * PLANE + (size_t)v * 16, uf, true);
                }
            }
        }
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            if (has_gu) {
                // D1 = Gh.Wh + Gl.Wh + Gh.Wl    (K = Cg in steps of 16 = two channel-group planes)
                for (int j = 0; j < Cg / 16; ++j) {
                    const uint64_t agh = tc::smem_desc(sGh_u + 2 * j * PLANE, PLANE, 128), agl = tc::smem_desc(sGl_u + 2 * j * PLANE, PLANE, 128);
                    const uint64_t bwh = tc::smem_desc(sWh_u + 2 * j * Cu * 16, Cu * 16, 128), bwl = tc::smem_desc(sWl_u + 2 * j * Cu * 16, Cu * 16, 128);
                    tc::mma_f16(d1, agh, bwh, id_k, j > 0 ? 1u : 0u);
                    tc::mma_f16(d1, agl, bwh, id_k, 1u);
                    tc::mma_f16(d1, agh, bwl, id_k, 1u);
                }
            }
            if (has_gw) {
                // D2 += Gh^T.Uh + Gl^T.Uh + Gh^T.Ul   (K = 128 voxels in steps of 16 = two 8-voxel groups)
                for (int j = 0; j < TV / 16; ++j) {
                    const uint64_t buh = tc::smem_desc(sUh_u + j * 256, 128, PLANE), bul = tc::smem_desc(sUl_u + j * 256, 128, PLANE);
                    const uint64_t agh = tc::smem_desc(sGh_u + j * 256, 128, PLANE);
                    tc::mma_f16(d2, agh, buh, id_mn, (first && j == 0) ? 0u : 1u);
                    tc::mma_f16(d2, tc::smem_desc(sGl_u + j * 256, 128, PLANE), buh, id_mn, 1u);
                    tc::mma_f16(d2, agh, bul, id_mn, 1u);
                }
            }
            tc::mma_commit(&s_bar);
        }
        first = false;
        tc::mbar_wait(&s_bar, phase);
        phase ^= 1u;
        tc::fence_after_sync();
        // ---- epilogue: D1 -> g_u (fp32); thread = voxel row, the two warp groups split the column blocks
        if (has_gu) {
            const int v = (warp & 3) * 32 + lane;
            const bool ok = v0 + v < A.vox;
            float *op = A.g_u + ((size_t)n * A.vox + v0 + (ok ? v : 0)) * (size_t)A.ldgu;
            const uint32_t trow = d1 + ((uint32_t)((warp & 3) * 32) << 16);
            for (int cb = (warp >> 2) * 16; cb < Cu; cb += 32) {
                float r[16];
                tc::tmem_ld16(trow + (uint32_t)cb, r);
                if (ok) store16_f32(op + cb, r, A.accumulate != 0, (reinterpret_cast<uintptr_t>(op) & 31) == 0);
            }
        }
        tc::fence_before_sync();
        __syncthreads();                 // tiles and D1 free for the next work item
    }
    // ---- flush the weight gradient: row = output channel c, column = input channel k
    if (has_gw && !first) {
        tc::fence_after_sync();
        const int c = (warp & 3) * 32 + lane;
        const uint32_t trow = d2 + ((uint32_t)((warp & 3) * 32) << 16);
        if ((warp & 3) * 32 < Cg) {      // warp-uniform: this lane quarter holds valid rows
            for (int cb = (warp >> 2) * 16; cb < Cu; cb += 32) {
                float r[16];
                tc::tmem_ld16(trow + (uint32_t)cb, r);
                if (c < Cg) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) atomicAdd(&A.g_w[(size_t)c * Cu + cb + j], r[j]);
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)A.tmem_cols);
}

// Software-pipelined variant for the narrow layers (GI, UI >= 1; everything fits twice): the staged tiles are
// double-buffered, so while the tensor pipe works on tile T the CTA converts tile T+1 (whose global loads were issued a
// full iteration earlier) and issues the loads of T+2; the only exposed step per tile is the D1 epilogue.  Tiles are
// assigned in contiguous ranges (one or two sample changes per CTA).
template <typename T, int GI, int UI>
__global__ void __launch_bounds__(NT) pw_bwd_tc_pipe_kernel(PwTcArgs A) {
    const T *At = reinterpret_cast<const T *>(A.t), *Au = reinterpret_cast<const T *>(A.u);
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int Cg = A.Cg, Cu = A.Cu, gq = Cg >> 3, uq = Cu >> 3;
    // channel-group counts are powers of two in every configured model: shifts instead of run-time divisions per staged vector
    const int gq_sh = __ffs(gq) - 1, uq_sh = __ffs(uq) - 1;      // powers of two (host check)
    const uint32_t tile_bytes = (uint32_t)(2 * gq + 2 * uq) * PLANE;   // [Gh | Gl | Uh | Ul]
    unsigned char *sWh = smem + 2 * tile_bytes;
    unsigned char *sWl = sWh + (size_t)Cg * Cu * 2;
    double *s_ca = reinterpret_cast<double *>(sWl + (size_t)Cg * Cu * 2);
    double *s_cb = s_ca + Cg, *s_cd = s_cb + Cg;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool has_gu = A.g_u != nullptr, has_gw = A.g_w != nullptr, has_nt = A.nt.stats != nullptr;
    if (warp == 0) tc::tmem_alloc(&s_tmem, (uint32_t)A.tmem_cols);
    if (tid == 32) tc::mbar_init(&s_bar, 1);
    for (int i = tid; i < Cg * Cu; i += NT) {
        const int k = i % Cu, c = i / Cu;
        const float wv = A.w[i];
        const __nv_bfloat16 hi = __float2bfloat16_rn(wv);
        const uint32_t off = tc::tile_off(k, c, Cu);
        *reinterpret_cast<__nv_bfloat16 *>(sWh + off) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(sWl + off) = __float2bfloat16_rn(wv - __bfloat162float(hi));
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t d1 = tmem, d2 = tmem + (uint32_t)Cu;
    const uint32_t id_k = tc::idesc_16b_m128(Cu, 1, 1, false, false), id_mn = tc::idesc_16b_m128(Cu, 1, 1, true, true);
    const uint32_t smem_u = tc::smem_u32(smem), sWh_u = tc::smem_u32(sWh), sWl_u = tc::smem_u32(sWl);
    const long long tiles_per_sample = (A.vox + TV - 1) / TV;
    const uint32_t tps32 = (uint32_t)tiles_per_sample;
    const long long total_tiles = tiles_per_sample * A.N;
    const long long per = (total_tiles + gridDim.x - 1) / gridDim.x;
    const long long t_begin = (long long)blockIdx.x * per, t_end = t_begin + per < total_tiles ? t_begin + per : total_tiles;
    float4 pg0[GI], pg1[GI];
    V8<T> pt[GI], pu[UI];
    auto prefetch = [&](long long tl) {
        const int pn = (int)((uint32_t)tl / tps32);                                  // 32-bit: the host checks tiles < 2^31
        const long long pv0 = (long long)((uint32_t)tl - (uint32_t)pn * tps32) * TV;
#pragma unroll
        for (int i = 0; i < GI; ++i) {
            const int item = tid + i * NT, q = item & (gq - 1), v = item >> gq_sh;      // channel group fastest: a warp reads contiguous voxel rows
            if (pv0 + v < A.vox) {
                const size_t gv = (size_t)pn * A.vox + pv0 + v;
                pg0[i] = *reinterpret_cast<const float4 *>(A.gz + gv * (size_t)A.ldg + q * 8);
                pg1[i] = *reinterpret_cast<const float4 *>(A.gz + gv * (size_t)A.ldg + q * 8 + 4);
                if (has_nt) pt[i].load(At + gv * (size_t)A.ldt + q * 8);
            }
        }
        if (has_gw) {
#pragma unroll
            for (int i = 0; i < UI; ++i) {
                const int item = tid + i * NT, q = item & (uq - 1), v = item >> uq_sh;
                if (pv0 + v < A.vox) pu[i].load(Au + ((size_t)pn * A.vox + pv0 + v) * (size_t)A.ldu + q * 8);
            }
        }
    };
    int cur_n = -1;
    auto stage = [&](long long tl, int b) {          // registers (tile tl) -> tile buffer b
        const int n = (int)((uint32_t)tl / tps32);
        const long long v0 = (long long)((uint32_t)tl - (uint32_t)n * tps32) * TV;
        if (n != cur_n) {                              // uniform over the CTA; the previous stage() ended before a CTA barrier
            cur_n = n;
            for (int c = tid; c < Cg; c += NT) {
                double a, bb, d;
                in_bwd_coef_d(A.nt, A.red, A.N, Cg, n, c, a, bb, d);
                s_ca[c] = a; s_cb[c] = bb; s_cd[c] = d;
            }
            __syncthreads();
        }
        unsigned char *sGh = smem + (size_t)b * tile_bytes, *sGl = sGh + (size_t)gq * PLANE, *sUh = sGl + (size_t)gq * PLANE, *sUl = sUh + (size_t)uq * PLANE;
#pragma unroll
        for (int i = 0; i < GI; ++i) {
            const int item = tid + i * NT, q = item & (gq - 1), v = item >> gq_sh;
            uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
            if (v0 + v < A.vox) {
                float g[8] = {pg0[i].x, pg0[i].y, pg0[i].z, pg0[i].w, pg1[i].x, pg1[i].y, pg1[i].z, pg1[i].w};
                if (has_nt) {
                    float tv[8];
                    pt[i].unpack(tv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) g[j] = in_bwd_apply(s_ca[q * 8 + j], s_cb[q * 8 + j], s_cd[q * 8 + j], g[j], tv[j]);
                }
                split2(g[0], g[1], hi.x, lo.x); split2(g[2], g[3], hi.y, lo.y);
                split2(g[4], g[5], hi.z, lo.z); split2(g[6], g[7], hi.w, lo.w);
            }
            *reinterpret_cast<uint4 *>(sGh + (size_t)q * PLANE + (size_t)v * 16) = hi;
            *reinterpret_cast<uint4 *>(sGl + (size_t)q * PLANE + (size_t)v * 16) = lo;
        }
        if (has_gw) {
#pragma unroll
            for (int i = 0; i < UI; ++i) {
                const int item = tid + i * NT, q = item & (uq - 1), v = item >> uq_sh;
                store_u_split(sUh, sUl, (size_t)q * PLANE + (size_t)v * 16, pu[i], v0 + v < A.vox);
            }
        }
    };
    bool first = true;
    auto issue = [&](int b) {                          // one thread, after a CTA barrier that followed fence_async_smem
        tc::fence_after_sync();
        const uint32_t sGh_u = smem_u + (uint32_t)b * tile_bytes, sGl_u = sGh_u + (uint32_t)gq * PLANE, sUh_u = sGl_u + (uint32_t)gq * PLANE, sUl_u = sUh_u + (uint32_t)uq * PLANE;
        if (has_gu) {
            for (int j = 0; j < Cg / 16; ++j) {
                const uint64_t agh = tc::smem_desc(sGh_u + 2 * j * PLANE, PLANE, 128), agl = tc::smem_desc(sGl_u + 2 * j * PLANE, PLANE, 128);
                const uint64_t bwh = tc::smem_desc(sWh_u + 2 * j * Cu * 16, Cu * 16, 128), bwl = tc::smem_desc(sWl_u + 2 * j * Cu * 16, Cu * 16, 128);
                tc::mma_f16(d1, agh, bwh, id_k, j > 0 ? 1u : 0u);
                tc::mma_f16(d1, agl, bwh, id_k, 1u);
                tc::mma_f16(d1, agh, bwl, id_k, 1u);
            }
        }
        if (has_gw) {
            for (int j = 0; j < TV / 16; ++j) {
                const uint64_t buh = tc::smem_desc(sUh_u + j * 256, 128, PLANE), bul = tc::smem_desc(sUl_u + j * 256, 128, PLANE);
                const uint64_t agh = tc::smem_desc(sGh_u + j * 256, 128, PLANE);
                tc::mma_f16(d2, agh, buh, id_mn, (first && j == 0) ? 0u : 1u);
                tc::mma_f16(d2, tc::smem_desc(sGl_u + j * 256, 128, PLANE), buh, id_mn, 1u);
                tc::mma_f16(d2, agh, bul, id_mn, 1u);
            }
        }
        tc::mma_commit(&s_bar);
    };
    uint32_t phase = 0;
    if (t_begin < t_end) {
        prefetch(t_begin);
        stage(t_begin, 0);
        if (t_begin + 1 < t_end) prefetch(t_begin + 1);
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) issue(0);
        first = false;
    }
    for (long long tile = t_begin; tile < t_end; ++tile) {
        const int b = (int)((tile - t_begin) & 1);
        if (tile + 1 < t_end) {
            stage(tile + 1, b ^ 1);                    // overlaps the MMAs of `tile`
            tc::fence_async_smem();                    // here, not after the epilogue: the fence also drains this thread's global stores
            if (tile + 2 < t_end) prefetch(tile + 2);
        }
        tc::mbar_wait(&s_bar, phase);
        phase ^= 1u;
        tc::fence_after_sync();
        if (has_gu) {
            const int n = (int)((uint32_t)tile / tps32);
            const long long v0 = (long long)((uint32_t)tile - (uint32_t)n * tps32) * TV;
            const int v = (warp & 3) * 32 + lane;
            const bool ok = v0 + v < A.vox;
            float *op = A.g_u + ((size_t)n * A.vox + v0 + (ok ? v : 0)) * (size_t)A.ldgu;
            const uint32_t trow = d1 + ((uint32_t)((warp & 3) * 32) << 16);
            for (int cb = (warp >> 2) * 16; cb < Cu; cb += 32) {
                float r[16];
                tc::tmem_ld16(trow + (uint32_t)cb, r);
                if (ok) store16_f32(op + cb, r, A.accumulate != 0, (reinterpret_cast<uintptr_t>(op) & 31) == 0);
            }
        }
        tc::fence_before_sync();
        __syncthreads();                               // D1 drained; tile buffer b^1 complete
        if (tile + 1 < t_end && tid == 0) issue(b ^ 1);
    }
    if (has_gw && t_begin < t_end) {
        tc::fence_after_sync();
        const int c = (warp & 3) * 32 + lane;
        const uint32_t trow = d2 + ((uint32_t)((warp & 3) * 32) << 16);
        if ((warp & 3) * 32 < Cg) {
            for (int cb = (warp >> 2) * 16; cb < Cu; cb += 32) {
                float r[16];
                tc::tmem_ld16(trow + (uint32_t)cb, r);
                if (c < Cg) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) atomicAdd(&A.g_w[(size_t)c * Cu + cb + j], r[j]);
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)A.tmem_cols);
}

}  // namespace

// Returns -1 when the tensor-core path does not apply (the caller falls back to the CUDA-core kernel).
int l3d_pw_bwd_tc(const l3d_act *gz, const l3d_act *t, const l3d_norm *nt, const double *red, const l3d_act *u,
                  const l3d_norm *un, int N, long long vox, const float *w, float *g_w, const l3d_act *g_u,
                  int accumulate_gu, void *stream) {
    if (L3D_ENV_INT("L3D_NO_TC_BWD", 0) == 1) return -1;
    const int Cg = gz->C, Cu = u->C;
    const bool has_nt = nt != nullptr && nt->stats != nullptr, has_gu = !act_null(g_u);
    const bool f32 = u->dtype == L3D_F32;                 // storage type of the activations t and u (gradients are always fp32)
    if (gz->dtype != L3D_F32 || (has_nt && t->dtype != u->dtype)) return -1;
    if (Cg % 16 != 0 || Cu % 16 != 0 || Cg > 128 || Cu > 256) return -1;
    auto al = [](const l3d_act *a, int elems, int bytes) { return a->ldc % elems == 0 && reinterpret_cast<uintptr_t>(a->ptr) % bytes == 0; };
    const int sv = f32 ? 4 : 8;
    if (!al(gz, 4, 16) || !al(u, sv, 16) || (has_nt && !al(t, sv, 16)) || (has_gu && !al(g_u, 4, 16))) return -1;
    if (g_w == nullptr && !has_gu) return -1;
    // an activated u is not exactly representable in bf16 (it would need a hi/lo pair like g_t); every caller on the
    // U-Net path passes a stored bf16 tensor with the identity norm, so the general case stays on the CUDA-core kernel
    if (un != nullptr && un->stats != nullptr && g_w != nullptr) return -1;
    size_t smem = (size_t)(2 * (Cg / 8) + 2 * (Cu / 8)) * PLANE + 2 * (size_t)Cg * Cu * 2 + sizeof(double) * 3 * (size_t)Cg + sizeof(float) * 2 * (size_t)Cu;
    // the MN-major A operand always spans 128 rows (16 channel groups): keep its over-read inside the allocation
    const size_t span = (size_t)(Cg / 8) * PLANE + 16 * (size_t)PLANE + 256;
    if (smem < span) smem = span;
    if (smem > 226 * 1024) return -1;
    int cols = 32;
    while (cols < 2 * Cu) cols <<= 1;
    PwTcArgs A;
    A.gz = (const float *)gz->ptr; A.ldg = gz->ldc;
    A.t = has_nt ? t->ptr : nullptr; A.ldt = has_nt ? t->ldc : 0; A.nt = norm_dev(nt); A.red = red;
    A.u = u->ptr; A.ldu = u->ldc; A.un = norm_dev(un);
    A.N = N; A.vox = vox; A.Cg = Cg; A.Cu = Cu;
    A.w = w; A.g_w = g_w;
    A.g_u = has_gu ? (float *)g_u->ptr : nullptr; A.ldgu = has_gu ? g_u->ldc : 0; A.accumulate = accumulate_gu;
    A.tmem_cols = cols;
    int occ = (int)((227 * 1024) / (smem + 2048));
    if (occ > 4) occ = 4;
    if (occ < 1) occ = 1;
    if (occ * cols > 512) occ = 512 / cols;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = ((vox + TV - 1) / TV) * N;
    if (tiles >= (1ll << 31)) return -1;          // 32-bit tile arithmetic in the kernels
    // pipelined variant: two tile buffers + weights + tables, and the MN-major over-read of the second buffer's G planes
    size_t smem_p = 2 * (size_t)(2 * (Cg / 8) + 2 * (Cu / 8)) * PLANE + 2 * (size_t)Cg * Cu * 2 + sizeof(double) * 3 * (size_t)Cg;
    {
        const size_t span_p = (size_t)(2 * (Cg / 8) + 2 * (Cu / 8)) * PLANE + (size_t)(Cg / 8) * PLANE + 16 * (size_t)PLANE + 256;
        if (smem_p < span_p) smem_p = span_p;
    }
    // the pipelined kernel decodes (voxel, channel group) with shifts: power-of-two channel counts only
    const bool use_pipe = L3D_ENV_INT("L3D_NO_PWB_PIPE", 0) != 1 && (Cg & (Cg - 1)) == 0 && (Cu & (Cu - 1)) == 0;
#define L3D_PWTC_T(TT, GIV, UIV)                                                                                               \
    do {                                                                                                                        \
        {                                                                                                                       \
            cudaError_t e = cudaFuncSetAttribute(pw_bwd_tc_kernel<TT, GIV, UIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); \
            if (e == cudaSuccess && GIV > 0) e = cudaFuncSetAttribute(pw_bwd_tc_pipe_kernel<TT, (GIV > 0 ? GIV : 1), (UIV > 0 ? UIV : 1)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); \
            if (e != cudaSuccess) { l3d_set_error("pw_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 3; }    \
        }                                                                                                                       \
        static int occ_regs = 0, occ_regs_p = 0;                                                                                \
        if (occ_regs == 0) {                                                                                                    \
            cudaFuncAttributes fa;                                                                                              \
            occ_regs = occ_regs_p = 1;                                                                                          \
            if (cudaFuncGetAttributes(&fa, pw_bwd_tc_kernel<TT, GIV, UIV>) == cudaSuccess && fa.numRegs > 0)                        \
                occ_regs = 65536 / (((fa.numRegs + 7) / 8 * 8) * NT);                                                           \
            if (cudaFuncGetAttributes(&fa, pw_bwd_tc_pipe_kernel<TT, (GIV > 0 ? GIV : 1), (UIV > 0 ? UIV : 1)>) == cudaSuccess && fa.numRegs > 0) \
                occ_regs_p = 65536 / (((fa.numRegs + 7) / 8 * 8) * NT);                                                         \
            if (occ_regs < 1) occ_regs = 1;                                                                                     \
            if (occ_regs_p < 1) occ_regs_p = 1;                                                                                 \
        }                                                                                                                       \
        if (GIV > 0 && use_pipe && smem_p <= 100 * 1024) {                                                                      \
            int occ_p = (int)((227 * 1024) / (smem_p + 2048));                                                                  \
            if (occ_p > 4) occ_p = 4;                                                                                           \
            if (occ_p * cols > 512) occ_p = 512 / cols;                                                                         \
            if (occ_p > occ_regs_p) occ_p = occ_regs_p;                                                                         \
            if (occ_p < 1) occ_p = 1;                                                                                           \
            long long grid_p = (long long)sms * occ_p;                                                                          \
            if (grid_p > tiles) grid_p = tiles;                                                                                 \
            pw_bwd_tc_pipe_kernel<TT, (GIV > 0 ? GIV : 1), (UIV > 0 ? UIV : 1)><<<(unsigned)grid_p, NT, smem_p, (cudaStream_t)stream>>>(A); \
        } else {                                                                                                                \
            const int occ_k = occ_regs < occ ? occ_regs : occ;                                                                  \
            long long grid_k = (long long)sms * occ_k;                                                                          \
            if (grid_k > tiles) grid_k = tiles;                                                                                 \
            pw_bwd_tc_kernel<TT, GIV, UIV><<<(unsigned)grid_k, NT, smem, (cudaStream_t)stream>>>(A);                                \
        }                                                                                                                       \
    } while (0)
#define L3D_PWTC(GIV, UIV) do { if (f32) L3D_PWTC_T(float, GIV, UIV); else L3D_PWTC_T(h16, GIV, UIV); } while (0)
    const int gi = Cg / 16, ui = Cu / 16;            // staging items per thread (TV * C / 8 / NT)
    if (gi == 1 && ui == 1) L3D_PWTC(1, 1);
    else if (gi == 1 && ui == 2) L3D_PWTC(1, 2);
    else if (gi == 2 && ui == 1) L3D_PWTC(2, 1);
    else if (gi == 2 && ui == 2) L3D_PWTC(2, 2);
    else if (gi == 2 && ui == 4) L3D_PWTC(2, 4);
    else if (gi == 4 && ui == 2) L3D_PWTC(4, 2);
    else if (gi == 4 && ui == 4) L3D_PWTC(4, 4);
    else L3D_PWTC(0, 0);
#undef L3D_PWTC
#undef L3D_PWTC_T
    L3D_CUDA_OK("l3d_pw_bwd (tcgen05) launch");
    return 0;
}

// =============================================================================================================
// ConvTranspose3d(k=2, s=2) backward on tensor cores (bf16 storage; unet3d.py:119, 127-141).  The eight taps are
// eight independent pointwise maps  out[up(v, tap)][co] = sum_ci x[v][ci] * W[ci][co][tap] + b[co],  so per tile of
// 128 input voxels and per tap:
//   stage   G_tap[v][co] = g_out[up(v, tap)][co] (zero outside the output volume) as bf16 hi / lo, voxel-planar
//   dgrad   D1[v][ci]  += G_tap . W_tap            (K-major;  accumulated over the 8 taps, then stored to g_x)
//   wgrad   D2_tap[co][ci | 1] += G_tap^T . [X | 1] (MN-major; the extra ones-column of the X tile yields the bias
//                                                   gradient; D2 stays in TMEM over all tiles of the CTA)
// TMEM holds D1 and TP of the eight D2 accumulators; blockIdx.y selects the group of TP taps whose weight gradient
// this CTA owns, and only blockIdx.y == 0 computes the input gradient (over all taps).
namespace {

struct CtTcArgs {
    const float *g; int ldg; int OD, OH, OW, oz, oy, ox;
    const void *x; int ldx; int N, d, h, w;
    int Cin, Cout;
    const float *wgt; float *g_w; float *g_b;
    float *g_x; int ldgx; int accumulate;
    int TP, w_resident, tmem_cols;
};

template <typename T>
__global__ void __launch_bounds__(NT) convt_bwd_tc_kernel(CtTcArgs A) {
    const T *Ax = reinterpret_cast<const T *>(A.x);
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int Cin = A.Cin, Cout = A.Cout, gq = Cout >> 3, xq = (Cin >> 3) + 2, NX = Cin + 16;
    unsigned char *sGh = smem;                                   // gq planes
    unsigned char *sGl = sGh + (size_t)gq * PLANE;
    unsigned char *sXh = sGl + (size_t)gq * PLANE;               // xq planes: Cin channels + [1, 0 x 15]; bf16 hi, then lo
    unsigned char *sXl = sXh + (size_t)xq * PLANE;
    unsigned char *sWh = sXl + (size_t)xq * PLANE;               // dgrad B operand per tap: [N = Cin][K = Cout] K-major
    const size_t wtap_bytes = (size_t)Cin * Cout * 2;
    unsigned char *sWl = sWh + (A.w_resident ? 8 : 1) * wtap_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pass = blockIdx.y, TP = A.TP;
    const bool do_dgrad = pass == 0 && A.g_x != nullptr;
    const int tap_lo = pass * TP, tap_hi = tap_lo + TP;          // weight-gradient taps of this CTA
    const int tap_begin = do_dgrad ? 0 : tap_lo, tap_end = do_dgrad ? 8 : tap_hi;
    if (warp == 0) tc::tmem_alloc(&s_tmem, (uint32_t)A.tmem_cols);
    if (tid == 32) tc::mbar_init(&s_bar, 1);
    auto stage_w = [&](int tap, int slot) {                      // wgt[ci][co][tap] -> B[n = ci][k = co]
        for (int i = tid; i < Cin * Cout; i += NT) {
            const int co = i % Cout, ci = i / Cout;
            const float wv = A.wgt[(size_t)i * 8 + tap];
            const __nv_bfloat16 hi = __float2bfloat16_rn(wv);
            const uint32_t off = tc::tile_off(ci, co, Cin);
            *reinterpret_cast<__nv_bfloat16 *>(sWh + slot * wtap_bytes + off) = hi;
            *reinterpret_cast<__nv_bfloat16 *>(sWl + slot * wtap_bytes + off) = __float2bfloat16_rn(wv - __bfloat162float(hi));
        }
    };
    if (A.w_resident && do_dgrad) for (int tap = 0; tap < 8; ++tap) stage_w(tap, tap);
    // constant part of the X tile: channel Cin = 1, channels Cin+1 .. Cin+15 = 0
    for (int i = tid; i < 2 * TV; i += NT) {
        const int v = i & (TV - 1), q = (Cin >> 3) + (i >> 7);
        *reinterpret_cast<uint4 *>(sXh + (size_t)q * PLANE + (size_t)v * 16) = make_uint4(i < TV ? 0x00003f80u : 0u, 0u, 0u, 0u);   // bf16 1.0
        *reinterpret_cast<uint4 *>(sXl + (size_t)q * PLANE + (size_t)v * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t d1 = tmem;                                     // Cin columns
    const uint32_t d2 = tmem + (uint32_t)Cin;                     // TP accumulators of NX columns
    const uint32_t id_k = tc::idesc_16b_m128(Cin, 1, 1, false, false), id_mn = tc::idesc_16b_m128(NX, 1, 1, true, true);
    const uint32_t sGh_u = tc::smem_u32(sGh), sGl_u = tc::smem_u32(sGl), sXh_u = tc::smem_u32(sXh), sXl_u = tc::smem_u32(sXl), sWh_u = tc::smem_u32(sWh), sWl_u = tc::smem_u32(sWl);
    const long long nvox = (long long)A.N * A.d * A.h * A.w;
    const long long ntiles = (nvox + TV - 1) / TV;
    uint32_t phase = 0;
    bool first = true;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long v0 = tile * TV;
        // ---- X tile (stored fp16, identity norm -> exact bf16 hi / lo): one copy per tile
        for (int item = tid; item < (Cin >> 3) * TV; item += NT) {
            const int v = item & (TV - 1), q = item >> 7;
            V8<T> o;
            const bool ok = v0 + v < nvox;
            if (ok) o.load(Ax + (size_t)(v0 + v) * A.ldx + q * 8);
            float xf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (ok) o.unpack(xf);
            store_split8(sXh, sXl, (size_t)q * PLANE + (size_t)v * 16, xf, ok);
        }
        for (int tap = tap_begin; tap < tap_end; ++tap) {
            // ---- G_tap tile: gather from the up-sampled grid
            for (int item = tid; item < gq * TV; item += NT) {
                const int v = item & (TV - 1), q = item >> 7;
                uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
                if (v0 + v < nvox) {
                    uint32_t rem = (uint32_t)(v0 + v);             // nvox < 2^31 (host check): 32-bit divisions
                    const int ix = (int)(rem % (uint32_t)A.w); rem /= (uint32_t)A.w;
                    const int iy = (int)(rem % (uint32_t)A.h); rem /= (uint32_t)A.h;
                    const int iz = (int)(rem % (uint32_t)A.d);
                    const int n = (int)(rem / (uint32_t)A.d);
                    const int Z = A.oz + 2 * iz + (tap >> 2), Y = A.oy + 2 * iy + ((tap >> 1) & 1), X = A.ox + 2 * ix + (tap & 1);
                    if (Z >= 0 && Z < A.OD && Y >= 0 && Y < A.OH && X >= 0 && X < A.OW) {
                        const float *gp = A.g + ((((size_t)n * A.OD + Z) * A.OH + Y) * A.OW + X) * (size_t)A.ldg + q * 8;
                        const float4 g0 = *reinterpret_cast<const float4 *>(gp), g1 = *reinterpret_cast<const float4 *>(gp + 4);
                        split2(g0.x, g0.y, hi.x, lo.x); split2(g0.z, g0.w, hi.y, lo.y);
                        split2(g1.x, g1.y, hi.z, lo.z); split2(g1.z, g1.w, hi.w, lo.w);
                    }
                }
                *reinterpret_cast<uint4 *>(sGh + (size_t)q * PLANE + (size_t)v * 16) = hi;
                *reinterpret_cast<uint4 *>(sGl + (size_t)q * PLANE + (size_t)v * 16) = lo;
            }
            if (!A.w_resident && do_dgrad) stage_w(tap, 0);
            tc::fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                tc::fence_after_sync();
                if (do_dgrad) {
                    const uint32_t wslot = (uint32_t)((A.w_resident ? tap : 0) * wtap_bytes);
                    for (int j = 0; j < Cout / 16; ++j) {
                        const uint64_t agh = tc::smem_desc(sGh_u + 2 * j * PLANE, PLANE, 128), agl = tc::smem_desc(sGl_u + 2 * j * PLANE, PLANE, 128);
                        const uint64_t bwh = tc::smem_desc(sWh_u + wslot + 2 * j * Cin * 16, Cin * 16, 128);
                        const uint64_t bwl = tc::smem_desc(sWl_u + wslot + 2 * j * Cin * 16, Cin * 16, 128);
                        tc::mma_f16(d1, agh, bwh, id_k, (tap > 0 || j > 0) ? 1u : 0u);
                        tc::mma_f16(d1, agl, bwh, id_k, 1u);
                        tc::mma_f16(d1, agh, bwl, id_k, 1u);
                    }
                }
                if (tap >= tap_lo && tap < tap_hi) {
                    const uint32_t dacc = d2 + (uint32_t)((tap - tap_lo) * NX);
                    for (int j = 0; j < TV / 16; ++j) {
                        const uint64_t bxh = tc::smem_desc(sXh_u + j * 256, 128, PLANE), bxl = tc::smem_desc(sXl_u + j * 256, 128, PLANE);
                        const uint64_t agh = tc::smem_desc(sGh_u + j * 256, 128, PLANE);
                        tc::mma_f16(dacc, agh, bxh, id_mn, (first && j == 0) ? 0u : 1u);
                        tc::mma_f16(dacc, tc::smem_desc(sGl_u + j * 256, 128, PLANE), bxh, id_mn, 1u);
                        tc::mma_f16(dacc, agh, bxl, id_mn, 1u);
                    }
                }
                tc::mma_commit(&s_bar);
            }
            tc::mbar_wait(&s_bar, phase);          // G tile (and the per-tap weights) free again
            phase ^= 1u;
            tc::fence_after_sync();
        }
        first = false;
        // ---- epilogue: D1 -> g_x (fp32)
        if (do_dgrad) {
            const int v = (warp & 3) * 32 + lane;
            const bool ok = v0 + v < nvox;
            float *op = A.g_x + (size_t)(v0 + (ok ? v : 0)) * A.ldgx;
            const uint32_t trow = d1 + ((uint32_t)((warp & 3) * 32) << 16);
            for (int cb = (warp >> 2) * 16; cb < Cin; cb += 32) {
                float r[16];
                tc::tmem_ld16(trow + (uint32_t)cb, r);
                if (ok) store16_f32(op + cb, r, A.accumulate != 0, (reinterpret_cast<uintptr_t>(op) & 31) == 0);
            }
        }
        tc::fence_before_sync();
        __syncthreads();
    }
    // ---- flush: D2_tap[co][ci] -> g_w[ci][co][tap], D2_tap[co][Cin] -> g_b[co]
    if (!first) {
        tc::fence_after_sync();
        const int co = (warp & 3) * 32 + lane;
        if ((warp & 3) * 32 < Cout) {
            for (int tl = 0; tl < TP; ++tl) {
                const int tap = tap_lo + tl;
                const uint32_t trow = d2 + (uint32_t)(tl * NX) + ((uint32_t)((warp & 3) * 32) << 16);
                for (int cb = (warp >> 2) * 16; cb < NX; cb += 32) {
                    float r[16];
                    tc::tmem_ld16(trow + (uint32_t)cb, r);
                    if (co < Cout) {
                        if (cb < Cin) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) atomicAdd(&A.g_w[((size_t)(cb + j) * Cout + co) * 8 + tap], r[j]);
                        } else if (A.g_b != nullptr) {
                            atomicAdd(&A.g_b[co], r[0]);
                        }
                    }
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)A.tmem_cols);
}

}  // namespace

// Returns -1 when the tensor-core path does not apply.
int l3d_convt_bwd_tc(const l3d_act *g_out, int OD, int OH, int OW, int oz, int oy, int ox, const l3d_act *x, int N, int d, int h, int w_,
                     const float *w, float *g_w, float *g_b, const l3d_act *g_x, int accumulate_gx, void *stream) {
    if (L3D_ENV_INT("L3D_NO_TC_BWD", 0) == 1) return -1;
    const int Cin = x->C, Cout = g_out->C;
    const bool has_gx = !act_null(g_x);
    const bool f32 = x->dtype == L3D_F32;
    if (g_out->dtype != L3D_F32 || g_w == nullptr) return -1;
    if (Cin % 16 != 0 || Cout % 16 != 0 || Cin > 128 || Cout > 128) return -1;
    auto al = [](const l3d_act *a, int elems, int bytes) { return a->ldc % elems == 0 && reinterpret_cast<uintptr_t>(a->ptr) % bytes == 0; };
    if (!al(g_out, 4, 16) || !al(x, f32 ? 4 : 8, 16) || (has_gx && !al(g_x, 4, 16))) return -1;
    const int NX = Cin + 16;
    int TP = 8;
    while (TP > 1 && TP * NX + Cin > 512) TP >>= 1;
    if (TP * NX + Cin > 512) return -1;
    const size_t wtap = (size_t)Cin * Cout * 2;
    const size_t base = (size_t)(2 * (Cout / 8) + 2 * (Cin / 8 + 2)) * PLANE;
    int w_resident = base + 16 * wtap <= 160 * 1024 ? 1 : 0;
    size_t smem = base + (w_resident ? 16 : 2) * wtap;
    const size_t span = (size_t)(Cout / 8) * PLANE + 16 * (size_t)PLANE + 256;     // MN-major A over-read (128 rows)
    if (smem < span) smem = span;
    if (smem > 226 * 1024) return -1;
    int cols = 32;
    while (cols < TP * NX + Cin) cols <<= 1;
    CtTcArgs A;
    A.g = (const float *)g_out->ptr; A.ldg = g_out->ldc; A.OD = OD; A.OH = OH; A.OW = OW; A.oz = oz; A.oy = oy; A.ox = ox;
    A.x = x->ptr; A.ldx = x->ldc; A.N = N; A.d = d; A.h = h; A.w = w_;
    A.Cin = Cin; A.Cout = Cout; A.wgt = w; A.g_w = g_w; A.g_b = g_b;
    A.g_x = has_gx ? (float *)g_x->ptr : nullptr; A.ldgx = has_gx ? g_x->ldc : 0; A.accumulate = accumulate_gx;
    A.TP = TP; A.w_resident = w_resident; A.tmem_cols = cols;
    {
        cudaError_t e = f32 ? cudaFuncSetAttribute(convt_bwd_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)
                            : cudaFuncSetAttribute(convt_bwd_tc_kernel<h16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) { l3d_set_error("convt_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 3; }
    }
    int occ = (int)((227 * 1024) / (smem + 2048));
    if (occ > 3) occ = 3;
    if (occ < 1) occ = 1;
    if (occ * cols > 512) occ = 512 / cols;
    if (occ < 1) occ = 1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int npass = 8 / TP;
    const long long tiles = ((long long)N * d * h * w_ + TV - 1) / TV;
    if (tiles >= (1ll << 24)) return -1;          // 32-bit voxel indices in the kernel
    long long gx = ((long long)sms * occ + npass - 1) / npass;
    if (gx > tiles) gx = tiles;
    if (gx < 1) gx = 1;
    if (f32) convt_bwd_tc_kernel<float><<<dim3((unsigned)gx, (unsigned)npass), NT, smem, (cudaStream_t)stream>>>(A);
    else convt_bwd_tc_kernel<h16><<<dim3((unsigned)gx, (unsigned)npass), NT, smem, (cudaStream_t)stream>>>(A);
    L3D_CUDA_OK("l3d_convt_bwd (tcgen05) launch");
    return 0;
}
