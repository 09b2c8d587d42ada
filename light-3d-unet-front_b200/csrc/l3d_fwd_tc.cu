// Tensor-core forward of the depthwise-separable conv (bf16 storage): the hot kernel of the 217K-parameter model.
//
//   TMA halo tile (raw bf16) -> [InstanceNorm + LeakyReLU + Dropout3d] -> depthwise 3x3x3 (fp32 FMA, CUDA cores)
//   -> pointwise 1x1x1 (+ the block's 1x1x1 shortcut) as tcgen05.mma (fp16 operands, fp32 accumulate in TMEM)
//   -> bf16 store + InstanceNorm statistics in the epilogue            (unet3d.py:20-23, 70-72, 80-87)
//
// One persistent CTA (256 threads) walks 4x8x8-voxel tiles = two 128-row MMA tiles, 16 input channels at a time:
//   * one thread issues a 5-D TMA load of the 6x10x10x16 raw halo box (out-of-volume voxels arrive as zeros) for
//     the NEXT work item while the CUDA cores work on the current one (mbarrier complete_tx);
//   * an activation pass turns the raw box into the fp32 stencil tile (norm/act applied once per element, zero
//     outside the volume: the conv pads the *activated* tensor);
//   * the stencil writes its output straight into the K-major fp16 operand tile; one thread issues the K=16 MMA
//     step for that chunk, which runs asynchronously under the next chunk's work;
//   * the epilogue reads the accumulators with tcgen05.ld (thread = voxel row, 16 channels per load).
#include <cuda.h>

#include "l3d_common.cuh"
#include "l3d_tc.cuh"

namespace {

constexpr int TZ = 4, TY = 8, TX = 8, TV = TZ * TY * TX;
constexpr int HZ = TZ + 2, HY = TY + 2, HX = TX + 2;
constexpr int HXP = HX + 1, HPLANE = HY * HXP + 1, HVOX = HZ * HPLANE;
constexpr int CK = 16, NT = 256, MT = 2;
constexpr int RAW_BYTES = HZ * HY * HX * CK * 2;        // 19200: the TMA box, dense [z][y][x][c] bf16
constexpr int ACT_ITEMS = HZ * HY * HX * 2;             // 16-byte (8-channel) vectors in the box
constexpr int ACT_PER_THREAD = (ACT_ITEMS + NT - 1) / NT;

struct TcArgs {
    int Cin; NormDev xn;
    int N, D, H, W;
    const float *dw_w, *pw_w, *sc_w; int Cout;
    h16 *t; int ldt; double *t_stats;
    h16 *r; int ldr; double *r_stats;
    h16 *u; int ldu;
    int tmem_cols;
};


__global__ void __launch_bounds__(NT) dwpw_tc_kernel(const __grid_constant__ CUtensorMap tmap, TcArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar, s_tma_bar;
    __shared__ uint32_t s_tmem;
    const int Cin = A.Cin, Cout = A.Cout;
    const bool has_sc = A.sc_w != nullptr;
    const uint32_t a_bytes = (uint32_t)MT * 128 * Cin * 2, b_bytes = (uint32_t)Cout * Cin * 2;
    unsigned char *s_raw = smem_raw;                                           // RAW_BYTES (128-B aligned for TMA)
    unsigned char *sA = s_raw + RAW_BYTES;
    unsigned char *sA2 = sA + a_bytes;
    unsigned char *sB = sA2 + (has_sc ? a_bytes : 0);
    unsigned char *sB2 = sB + b_bytes;
    float *s_in = reinterpret_cast<float *>(sB2 + (has_sc ? b_bytes : 0));   // HVOX*CK
    float *s_dw = s_in + HVOX * CK;                                           // CK*27 (current chunk)
    float *s_scale = s_dw + CK * 27;                                          // Cin
    float *s_shift = s_scale + Cin;
    float *s_stat = s_shift + Cin;                                            // 2*Cout (t) + 2*Cout (r)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc::tmem_alloc(&s_tmem, (uint32_t)A.tmem_cols);
    if (tid == 32) { tc::mbar_init(&s_bar, 1); tc::mbar_init(&s_tma_bar, 1); }
    // stage the pointwise (+ shortcut) weights once as fp16 K-major operand tiles
    {
        // float4 loads, 4 in flight per thread (the weights are the only cold global reads of this kernel)
        const int nvec = Cout * Cin / 4;
        const int nsrc = has_sc ? 2 : 1;
        for (int s = 0; s < nsrc; ++s) {
            const float4 *src = reinterpret_cast<const float4 *>(s == 0 ? A.pw_w : A.sc_w);
            unsigned char *dst = s == 0 ? sB : sB2;
            for (int i0 = tid; i0 < nvec; i0 += 4 * NT) {
                float4 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) if (i0 + j * NT < nvec) v[j] = src[i0 + j * NT];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = i0 + j * NT;
                    if (i < nvec) {
                        const int k = (i * 4) % Cin, n = (i * 4) / Cin;
                        const __half2 h0 = __floats2half2_rn(v[j].x, v[j].y), h1 = __floats2half2_rn(v[j].z, v[j].w);
                        uint2 o;
                        o.x = *reinterpret_cast<const uint32_t *>(&h0);
                        o.y = *reinterpret_cast<const uint32_t *>(&h1);
                        *reinterpret_cast<uint2 *>(dst + tc::tile_off(n, k, Cout)) = o;
                    }
                }
            }
        }
    }
    for (int i = tid; i < 4 * Cout; i += NT) s_stat[i] = 0.f;
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = tc::idesc_f16_m128(Cout);
    const uint32_t sA_u = tc::smem_u32(sA), sA2_u = tc::smem_u32(sA2), sB_u = tc::smem_u32(sB), sB2_u = tc::smem_u32(sB2);

    const int tilesX = (A.W + TX - 1) / TX, tilesY = (A.H + TY - 1) / TY, tilesZ = (A.D + TZ - 1) / TZ;
    const long long tiles_per_sample = (long long)tilesX * tilesY * tilesZ;
    const long long total_tiles = tiles_per_sample * A.N;
    const int nchunks = Cin / CK;
    auto tile_coord = [&](long long tile, int &n, int &z0, int &y0, int &x0) {
        n = (int)(tile / tiles_per_sample);
        int b = (int)(tile % tiles_per_sample);
        x0 = (b % tilesX) * TX; b /= tilesX;
        y0 = (b % tilesY) * TY; b /= tilesY;
        z0 = b * TZ;
    };
    // activation-pass role: fixed set of 16-byte vectors of the raw box (same for every work item)
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
    uint32_t phase = 0, tphase = 0;
    int cur_n = -1;

    auto flush_stats = [&](int n) {
        if (n < 0) return;
        for (int i = tid; i < 2 * Cout; i += NT) {
            const int isq = i >= Cout, cc = isq ? i - Cout : i;
            atomicAdd(&A.t_stats[(size_t)isq * A.N * Cout + (size_t)n * Cout + cc], (double)s_stat[i]);
            s_stat[i] = 0.f;
            if (has_sc) {
                atomicAdd(&A.r_stats[(size_t)isq * A.N * Cout + (size_t)n * Cout + cc], (double)s_stat[2 * Cout + i]);
                s_stat[2 * Cout + i] = 0.f;
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
        for (int ch = 0; ch < nchunks; ++ch) {
            const int c0 = ch * CK;
            // ---- activation pass: raw bf16 box -> fp32 stencil tile
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
                const uint32_t it = act_item[k];
                if (it != 0xffffffffu) {
                    const int hx = (it >> 16) & 15, hy = (it >> 20) & 15, hz = (it >> 24) & 15, q = (it >> 28) & 1;
                    const int gz = z0 + hz - 1, gy = y0 + hy - 1, gx = x0 + hx - 1;
                    float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
                    if (gz >= 0 && gz < A.D && gy >= 0 && gy < A.H && gx >= 0 && gx < A.W) {
                        const uint4 rw = *reinterpret_cast<const uint4 *>(s_raw + (size_t)(tid + k * NT) * 16);
                        const float4 sc0 = *reinterpret_cast<const float4 *>(s_scale + c0 + q * 8);
                        const float4 sc1 = *reinterpret_cast<const float4 *>(s_scale + c0 + q * 8 + 4);
                        const float4 sh0 = *reinterpret_cast<const float4 *>(s_shift + c0 + q * 8);
                        const float4 sh1 = *reinterpret_cast<const float4 *>(s_shift + c0 + q * 8 + 4);
                        const float sl = A.xn.slope;
                        o0.x = lrelu(fmaf(h16_lo(rw.x), sc0.x, sh0.x), sl);
                        o0.y = lrelu(fmaf(h16_hi(rw.x), sc0.y, sh0.y), sl);
                        o0.z = lrelu(fmaf(h16_lo(rw.y), sc0.z, sh0.z), sl);
                        o0.w = lrelu(fmaf(h16_hi(rw.y), sc0.w, sh0.w), sl);
                        o1.x = lrelu(fmaf(h16_lo(rw.z), sc1.x, sh1.x), sl);
                        o1.y = lrelu(fmaf(h16_hi(rw.z), sc1.y, sh1.y), sl);
                        o1.z = lrelu(fmaf(h16_lo(rw.w), sc1.z, sh1.z), sl);
                        o1.w = lrelu(fmaf(h16_hi(rw.w), sc1.w, sh1.w), sl);
                    }
                    float *dst = s_in + (it & 0xffffu);
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
                // operand tiles: row = voxel within the 128-row MMA tile, column = channel
                const int m = lz >> 1;
                const int r0 = ((lz & 1) * TY + ly0) * TX;          // row of (lz, ly0, lx = 0), a multiple of 8
                const uint32_t base = (uint32_t)m * 128 * Cin * 2 + (uint32_t)((c0 + c) >> 3) * 2048 + (uint32_t)((c0 + c) & 7) * 2;
                const uint32_t o0 = base + (uint32_t)(r0 >> 3) * 128, o1 = o0 + 128;
#pragma unroll
                for (int i = 0; i < TX; ++i) {
                    *reinterpret_cast<__half *>(sA + o0 + i * 16) = __float2half_rn(acc0[i]);
                    *reinterpret_cast<__half *>(sA + o1 + i * 16) = __float2half_rn(acc1[i]);
                }
                if (has_sc) {
#pragma unroll
                    for (int i = 0; i < TX; ++i) {
                        *reinterpret_cast<__half *>(sA2 + o0 + i * 16) = __float2half_rn(ctr0[i]);
                        *reinterpret_cast<__half *>(sA2 + o1 + i * 16) = __float2half_rn(ctr1[i]);
                    }
                }
            }
            tc::fence_async_smem();
            __syncthreads();             // operand chunk complete; s_in / s_dw free for the next chunk
            if (tid == 0) {
                tc::fence_after_sync();
                const uint32_t acc = ch > 0 ? 1u : 0u;
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const uint64_t ad = tc::smem_desc(sA_u + m * 128 * Cin * 2 + 2 * ch * 2048, 2048, 128);
                    const uint64_t bd = tc::smem_desc(sB_u + 2 * ch * Cout * 16, Cout * 16, 128);
                    tc::mma_f16(tmem + m * Cout, ad, bd, idesc, acc);
                    if (has_sc) {
                        const uint64_t ad2 = tc::smem_desc(sA2_u + m * 128 * Cin * 2 + 2 * ch * 2048, 2048, 128);
                        const uint64_t bd2 = tc::smem_desc(sB2_u + 2 * ch * Cout * 16, Cout * 16, 128);
                        tc::mma_f16(tmem + (MT + m) * Cout, ad2, bd2, idesc, acc);
                    }
                }
                if (ch + 1 == nchunks) tc::mma_commit(&s_bar);
            }
        }
        // ---- optional save of the depthwise output (bf16) for the backward pass, straight from the operand tile
        if (A.u != nullptr) {
            const int kq = Cin >> 3;
            for (int item = tid; item < TV * kq; item += NT) {
                const int q = item % kq, v = item / kq;
                const int lx = v & 7, ly = (v >> 3) & 7, lzz = v >> 6;
                const int gz = z0 + lzz, gy = y0 + ly, gx = x0 + lx;
                if (gz < A.D && gy < A.H && gx < A.W) {
                    const uint4 h = *reinterpret_cast<const uint4 *>(sA + (uint32_t)(v >> 7) * 128 * Cin * 2 + (uint32_t)q * 2048 +
                                                                      (uint32_t)((v & 127) >> 3) * 128 + (uint32_t)(v & 7) * 16);
                    *reinterpret_cast<uint4 *>(A.u + ((((size_t)n * A.D + gz) * A.H + gy) * A.W + gx) * (size_t)A.ldu + q * 8) = h;   // the fp16 operand is the stored value
                }
            }
        }
        // ---- epilogue: TMEM -> bf16 global + statistics
        tc::mbar_wait(&s_bar, phase);
        phase ^= 1u;
        tc::fence_after_sync();
        {
            const int gz = z0 + elz, gy = y0 + ely, gx = x0 + elx;
            const bool valid = gz < A.D && gy < A.H && gx < A.W;
            const size_t vox = (((size_t)n * A.D + gz) * A.H + gy) * A.W + gx;
            const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            const int nacc = has_sc ? 2 : 1;
            for (int a = 0; a < nacc; ++a) {
                h16 *outp = (a == 0 ? A.t + vox * (size_t)A.ldt : A.r + vox * (size_t)A.ldr);
                float *stat = s_stat + a * 2 * Cout;
                for (int cb = 0; cb < Cout; cb += 16) {
                    float v[16];
                    tc::tmem_ld16(trow + (uint32_t)((a * MT + em) * Cout + cb), v);
                    float sv[32];
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        pk[j] = valid ? pack_h16x2(v[2 * j], v[2 * j + 1]) : 0u;
                        const float r0 = h16_lo(pk[j]);
                        const float r1 = h16_hi(pk[j]);
                        sv[2 * j] = r0; sv[2 * j + 1] = r1;
                        sv[16 + 2 * j] = r0 * r0; sv[16 + 2 * j + 1] = r1 * r1;
                    }
                    if (valid) {
                        *reinterpret_cast<uint4 *>(outp + cb) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4 *>(outp + cb + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                    warp_transpose_sum<32>(sv, lane);
                    const int idx = warp_transpose_owner<32>(lane);      // 0..15 sums, 16..31 squares
                    atomicAdd(&stat[(idx >= 16 ? Cout + idx - 16 : idx) + cb], sv[0]);
                }
            }
        }
        tc::fence_before_sync();
        __syncthreads();                 // TMEM drained, operand tiles free, statistics in s_stat
    }
    flush_stats(cur_n);
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)A.tmem_cols);
}

static size_t tc_smem_bytes(int Cin, int Cout, bool has_sc) {
    const size_t a = (size_t)MT * 128 * Cin * 2, b = (size_t)Cout * Cin * 2;
    return RAW_BYTES + (has_sc ? 2 : 1) * (a + b) + sizeof(float) * ((size_t)HVOX * CK + 27 * CK + 2 * (size_t)Cin + 4 * (size_t)Cout);
}

}  // namespace

// Returns -1 when the tensor-core path does not apply (the caller falls back to the generic kernel).
int l3d_dwpw_fwd_tc(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                    const float *dw_w, const float *pw_w, const float *sc_w,
                    const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats,
                    const l3d_act *u, void *stream) {
    static int disabled = -1;
    if (disabled < 0) { const char *e = getenv("L3D_NO_TC"); disabled = (e && e[0] == '1') ? 1 : 0; }
    if (disabled) return -1;
    const int Cin = x->C, Cout = t->C;
    const bool has_sc = sc_w != nullptr, has_u = !act_null(u);
    if (x->dtype != L3D_F16 || dw_w == nullptr) return -1;
    if (Cin % 16 != 0 || Cout % 16 != 0 || Cout > 256) return -1;
    const int cols_needed = MT * Cout * (has_sc ? 2 : 1);
    if (cols_needed > 512) return -1;
    const size_t smem = tc_smem_bytes(Cin, Cout, has_sc);
    if (smem > 226 * 1024) return -1;
    auto aligned = [](const l3d_act *a, int mult) {
        return (a->ldc % mult == 0) && (reinterpret_cast<uintptr_t>(a->ptr) % (2 * mult) == 0);
    };
    // TMA: 16-byte aligned base and strides
    if (!aligned(x, 8) || !aligned(t, 8) || (has_sc && !aligned(r, 8)) || (has_u && !aligned(u, 8))) return -1;
    if ((long long)N * D * H * W * x->ldc * 2 >= (1ll << 40)) return -1;
    int cols = 32;
    while (cols < cols_needed) cols <<= 1;

    // 5-D tensor map over the channels-last input view: (C, W, H, D, N), box (16, 10, 10, 6, 1)
    CUtensorMap tmap;
    {
        const cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        const cuuint64_t es = 2, ld = (cuuint64_t)x->ldc;
        const cuuint64_t strides[4] = {ld * es, (cuuint64_t)W * ld * es, (cuuint64_t)H * W * ld * es, (cuuint64_t)D * H * W * ld * es};
        const cuuint32_t box[5] = {CK, HX, HY, HZ, 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        if (l3d_encode_tiled(&tmap, (int)CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, x->ptr, (const unsigned long long *)dims,
                             (const unsigned long long *)strides, (const unsigned *)box, (const unsigned *)estr)) return 3;
    }
    TcArgs A;
    A.Cin = Cin; A.xn = norm_dev(xn);
    A.N = N; A.D = D; A.H = H; A.W = W;
    A.dw_w = dw_w; A.pw_w = pw_w; A.sc_w = sc_w; A.Cout = Cout;
    A.t = (h16 *)t->ptr; A.ldt = t->ldc; A.t_stats = t_stats;
    A.r = has_sc ? (h16 *)r->ptr : nullptr; A.ldr = has_sc ? r->ldc : 0; A.r_stats = r_stats;
    A.u = has_u ? (h16 *)u->ptr : nullptr; A.ldu = has_u ? u->ldc : 0;
    A.tmem_cols = cols;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(dwpw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) { l3d_set_error("dwpw_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 3; }
        attr_set = true;
    }
    int occ = (int)((227 * 1024) / (smem + 2048));
    if (occ > 2) occ = 2;                        // 256 threads x ~120 registers
    if (occ < 1) occ = 1;
    if (occ * cols > 512) occ = 512 / cols;      // TMEM columns are a per-SM resource
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (long long)N * ((D + TZ - 1) / TZ) * ((H + TY - 1) / TY) * ((W + TX - 1) / TX);
    long long grid = (long long)sms * occ;
    if (grid > tiles) grid = tiles;
    dwpw_tc_kernel<<<(unsigned)grid, NT, smem, (cudaStream_t)stream>>>(tmap, A);
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
    const h16 *x; int ldx; int Cin;
    int N, d, h, w;
    const float *wgt, *bias; int Cout;
    h16 *out; int ldo; int OD, OH, OW, oz, oy, ox;
    int tmem_cols;
};

__global__ void __launch_bounds__(NT) convt_tc_kernel(CtArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int Cin = A.Cin, Cout = A.Cout, NB = 8 * Cout;
    const uint32_t a_bytes = 128u * Cin * 2;
    unsigned char *sB = smem_raw;                         // NB * Cin * 2
    unsigned char *sA = sB + (size_t)NB * Cin * 2;        // 2 buffers of a_bytes
    float *s_bias = reinterpret_cast<float *>(sA + 2 * a_bytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc::tmem_alloc(&s_tmem, (uint32_t)A.tmem_cols);
    if (tid == 32) tc::mbar_init(&s_bar, 1);
    for (int i = tid; i < Cin * Cout * 8; i += NT) {      // wgt[ci][co][tap]
        const int tap = i & 7, co = (i >> 3) % Cout, ci = (i >> 3) / Cout;
        *reinterpret_cast<__half *>(sB + tc::tile_off(tap * Cout + co, ci, NB)) = __float2half_rn(A.wgt[i]);
    }
    for (int i = tid; i < Cout; i += NT) s_bias[i] = A.bias[i];
    const long long nvox = (long long)A.N * A.d * A.h * A.w;
    const long long ntiles = (nvox + 127) / 128;
    const int kq = Cin >> 3;
    auto load_a = [&](long long tile, int buf) {
        unsigned char *dst = sA + (size_t)buf * a_bytes;
        const long long v0 = tile * 128;
        for (int item = tid; item < 128 * kq; item += NT) {
            const int q = item % kq, v = item / kq;
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (v0 + v < nvox) {
                o = *reinterpret_cast<const uint4 *>(A.x + (size_t)(v0 + v) * A.ldx + q * 8);   // stored fp16 = MMA operand
            }
            *reinterpret_cast<uint4 *>(dst + (uint32_t)q * 2048 + (uint32_t)(v >> 3) * 128 + (uint32_t)(v & 7) * 16) = o;
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
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        if (tid == 0) {
            tc::fence_after_sync();
            for (int j = 0; j < Cin / 16; ++j)
                for (int hN = 0; hN < n_mma; ++hN) {
                    const uint64_t ad = tc::smem_desc(sA_u + buf * a_bytes + 2 * j * 2048, 2048, 128);
                    const uint64_t bd = tc::smem_desc(sB_u + 2 * j * NB * 16 + hN * (256 / 8) * 128, NB * 16, 128);
                    tc::mma_f16(tmem + hN * 256, ad, bd, idesc, j > 0);
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
            long long rem = row_ok ? gv : 0;
            const int ix = (int)(rem % A.w); rem /= A.w;
            const int iy = (int)(rem % A.h); rem /= A.h;
            const int iz = (int)(rem % A.d);
            const int n = (int)(rem / A.d);
            const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            for (int tp = tap0; tp < tap0 + 4; ++tp) {
                const int Z = A.oz + 2 * iz + (tp >> 2), Y = A.oy + 2 * iy + ((tp >> 1) & 1), X = A.ox + 2 * ix + (tp & 1);
                const bool ok = row_ok && Z >= 0 && Z < A.OD && Y >= 0 && Y < A.OH && X >= 0 && X < A.OW;
                h16 *op = A.out + ((((size_t)n * A.OD + Z) * A.OH + Y) * A.OW + X) * (size_t)A.ldo;
                for (int cb = 0; cb < Cout; cb += 16) {
                    float v[16];
                    tc::tmem_ld16(trow + (uint32_t)(tp * Cout + cb), v);
                    if (ok) {
                        uint32_t pk[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) pk[j] = pack_h16x2(v[2 * j] + s_bias[cb + 2 * j], v[2 * j + 1] + s_bias[cb + 2 * j + 1]);
                        *reinterpret_cast<uint4 *>(op + cb) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4 *>(op + cb + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
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
    static int disabled = -1;
    if (disabled < 0) { const char *e = getenv("L3D_NO_TC"); disabled = (e && e[0] == '1') ? 1 : 0; }
    if (disabled) return -1;
    const int Cin = x->C, Cout = out->C;
    if (x->dtype != L3D_F16 || Cin % 16 != 0 || Cout % 16 != 0 || 8 * Cout > 512) return -1;
    if (8 * Cout > 256 && (8 * Cout) % 256 != 0) return -1;
    auto aligned = [](const l3d_act *a, int mult) {
        return (a->ldc % mult == 0) && (reinterpret_cast<uintptr_t>(a->ptr) % (2 * mult) == 0);
    };
    if (!aligned(x, 8) || !aligned(out, 8)) return -1;
    const size_t smem = (size_t)8 * Cout * Cin * 2 + 2 * (size_t)128 * Cin * 2 + sizeof(float) * Cout;
    if (smem > 226 * 1024) return -1;
    int cols = 32;
    while (cols < 8 * Cout) cols <<= 1;
    CtArgs A;
    A.x = (const h16 *)x->ptr; A.ldx = x->ldc; A.Cin = Cin;
    A.N = N; A.d = d; A.h = h; A.w = w_;
    A.wgt = w; A.bias = b; A.Cout = Cout;
    A.out = (h16 *)out->ptr; A.ldo = out->ldc; A.OD = OD; A.OH = OH; A.OW = OW; A.oz = oz; A.oy = oy; A.ox = ox;
    A.tmem_cols = cols;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(convt_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) { l3d_set_error("convt_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 3; }
        attr_set = true;
    }
    int occ = (int)((227 * 1024) / (smem + 2048));
    if (occ > 4) occ = 4;                        // 56 registers; the TMEM columns (8 * Cout per CTA) bound it below
    if (occ < 1) occ = 1;
    if (occ * cols > 512) occ = 512 / cols;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = ((long long)N * d * h * w_ + 127) / 128;
    long long grid = (long long)sms * occ;
    if (grid > tiles) grid = tiles;
    convt_tc_kernel<<<(unsigned)grid, NT, smem, (cudaStream_t)stream>>>(A);
    l3d_count_launch();
    l3d_note_kernel("convt_tc_kernel");
    L3D_CUDA_OK("l3d_convt_fwd (tcgen05) launch");
    return 0;
}
