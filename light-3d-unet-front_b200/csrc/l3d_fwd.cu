// Forward kernels of the 3D U-Net hot path (sm_100a).
//
//  dwpw_fwd_kernel   fused [norm+LeakyReLU+dropout on load] -> depthwise 3x3x3 -> pointwise 1x1x1
//                    (+ the block's 1x1x1 shortcut from the same tile) with InstanceNorm statistics
//                    accumulated in the epilogue          (unet3d.py:20-23, 70-72, 80-87)
//  conv3_fwd_kernel  dense / grouped 3x3x3 direct convolution, same prologue/epilogue
//                    (unet3d.py:30, 49, 60)
//  merge_fwd_kernel  lrelu(IN(t2) + IN(r)) + fused MaxPool3d(2) / 1x1x1 head + sigmoid
//                    (unet3d.py:87-91, 109, 220-221)
//  convt_fwd_kernel  ConvTranspose3d(k=2,s=2)+bias scattered into the skip-concat buffer
//                    (unet3d.py:127-141)
#include "l3d_common.cuh"

namespace {

// ---- spatial tile shared by the stencil kernels -------------------------------------------
constexpr int TZ = 4, TY = 8, TX = 8, TV = TZ * TY * TX;  // 256 output voxels per CTA
constexpr int HZ = TZ + 2, HY = TY + 2, HX = TX + 2;      // halo tile 6 x 10 x 10
constexpr int HXP = HX + 1;                               // x pitch 11 (odd)
constexpr int HPLANE = HY * HXP + 1;                      // plane pitch 111 (odd => z-neighbour lanes hit the other 16 banks)
constexpr int HVOX = HZ * HPLANE;                         // 666 padded halo voxels
constexpr int CK = 16;                                    // input channels staged per pass
constexpr int NT = 256;                                   // threads per CTA

struct TileCoord {
    int n, z0, y0, x0;
};
__device__ __forceinline__ TileCoord decode_tile(int D, int H, int W) {
    const int tilesX = (W + TX - 1) / TX, tilesY = (H + TY - 1) / TY, tilesZ = (D + TZ - 1) / TZ;
    int b = blockIdx.x;
    TileCoord t;
    t.x0 = (b % tilesX) * TX; b /= tilesX;
    t.y0 = (b % tilesY) * TY; b /= tilesY;
    t.z0 = (b % tilesZ) * TZ; b /= tilesZ;
    t.n = b;
    return t;
}
static inline int64_t num_tiles(int N, int D, int H, int W) {
    return (int64_t)N * ((D + TZ - 1) / TZ) * ((H + TY - 1) / TY) * ((W + TX - 1) / TX);
}

// 16-byte channel vectors: 8 bf16 or 4 fp32
template <typename T> struct VecW;
template <> struct VecW<float> { static constexpr int V = 4; };
template <> struct VecW<h16> { static constexpr int V = 8; };
__device__ __forceinline__ void ldv(const float *p, float (&v)[4]) {
    const float4 f = *reinterpret_cast<const float4 *>(p);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}
__device__ __forceinline__ void ldv(const h16 *p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4 *>(p);
    v[0] = h16_lo(r.x); v[1] = h16_hi(r.x);
    v[2] = h16_lo(r.y); v[3] = h16_hi(r.y);
    v[4] = h16_lo(r.z); v[5] = h16_hi(r.z);
    v[6] = h16_lo(r.w); v[7] = h16_hi(r.w);
}
// raw 16-byte loads (conversion deferred, keeps the register footprint of in-flight loads small)
__device__ __forceinline__ uint4 ldraw(const float *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ uint4 ldraw(const h16 *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ void cvt_raw(const float *, const uint4 &r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
}
__device__ __forceinline__ void cvt_raw(const h16 *, const uint4 &r, float (&v)[8]) {
    v[0] = h16_lo(r.x); v[1] = h16_hi(r.x);
    v[2] = h16_lo(r.y); v[3] = h16_hi(r.y);
    v[4] = h16_lo(r.z); v[5] = h16_hi(r.z);
    v[6] = h16_lo(r.w); v[7] = h16_hi(r.w);
}
__device__ __forceinline__ void stv(float *p, const float (&v)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void stv(h16 *p, const float (&v)[8]) {
    *reinterpret_cast<uint4 *>(p) = make_uint4(pack_h16x2(v[0], v[1]), pack_h16x2(v[2], v[3]), pack_h16x2(v[4], v[5]), pack_h16x2(v[6], v[7]));
}

// per-channel prologue scale/shift into shared memory
__device__ __forceinline__ void setup_prologue(const NormDev &nd, int N, int C, int n, float *s_scale, float *s_shift) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sc, sh;
        norm_scale_shift(nd, N, C, n, c, sc, sh);
        s_scale[c] = sc;
        s_shift[c] = sh;
    }
}

// Load the (TZ+2)x(TY+2)x(TX+2) halo tile of channels [c0, c0+CK) into s_in[hvox][CK] as activated fp32
// (zero outside the volume and for channels >= C: the conv pads the *activated* tensor with zeros).
template <typename T>
__device__ __forceinline__ void load_halo_chunk(const T *__restrict__ x, int ldc, int C, int c0, bool vec_ok,
                                                const TileCoord &tc, int D, int H, int W,
                                                const float *s_scale, const float *s_shift, float slope,
                                                float *s_in) {
    for (int item = threadIdx.x; item < HZ * HY * HX * (CK / 4); item += NT) {
        const int q = item & (CK / 4 - 1);
        int hv = item / (CK / 4);
        const int hx = hv % HX; hv /= HX;
        const int hy = hv % HY;
        const int hz = hv / HY;
        const int gz = tc.z0 + hz - 1, gy = tc.y0 + hy - 1, gx = tc.x0 + hx - 1;
        const int c = c0 + q * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W && c < C) {
            const T *p = x + ((((size_t)tc.n * D + gz) * H + gy) * W + gx) * (size_t)ldc + c;
            if (vec_ok && c + 3 < C) {
                const float4 f = ld4(p);
                v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (c + j < C) v[j] = ld1(p + j);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c + j < C) v[j] = lrelu(v[j] * s_scale[c + j] + s_shift[c + j], slope);
            }
        }
        float4 o; o.x = v[0]; o.y = v[1]; o.z = v[2]; o.w = v[3];
        *reinterpret_cast<float4 *>(s_in + (size_t)(hz * HPLANE + hy * HXP + hx) * CK + q * 4) = o;
    }
}

// Fill s_u[v][k] (pitch P) with the activated centre voxels of the tile, all C channels.
template <typename T>
__device__ __forceinline__ void load_center_all(const T *__restrict__ x, int ldc, int C, bool vec_ok, const TileCoord &tc,
                                                int D, int H, int W, const float *s_scale, const float *s_shift,
                                                float slope, float *s_u, int P) {
    const int groups = (C + 3) / 4;
    for (int item = threadIdx.x; item < TV * groups; item += NT) {
        const int q = item % groups;
        const int v = item / groups;
        const int lx = v & 7, ly = (v >> 3) & 7, lz = v >> 6;
        const int gz = tc.z0 + lz, gy = tc.y0 + ly, gx = tc.x0 + lx;
        const int c = q * 4;
        float val[4] = {0.f, 0.f, 0.f, 0.f};
        if (gz < D && gy < H && gx < W) {
            const T *p = x + ((((size_t)tc.n * D + gz) * H + gy) * W + gx) * (size_t)ldc + c;
            if (vec_ok && c + 3 < C) {
                const float4 f = ld4(p);
                val[0] = f.x; val[1] = f.y; val[2] = f.z; val[3] = f.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (c + j < C) val[j] = ld1(p + j);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c + j < C) val[j] = lrelu(val[j] * s_scale[c + j] + s_shift[c + j], slope);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (c + j < C) s_u[(size_t)v * P + c + j] = val[j];
    }
}

// Pointwise GEMM phase: out[v][co] = sum_k s_u[v][k] * w[co][k] for the 256 voxels of the tile, written to
// `out` (raw) with {sum, sumsq} accumulated into stats[2][N][Cout].  Each thread owns 2 voxels x CPT outputs.
template <typename T, int CPT>
__device__ __forceinline__ void pw_phase(const float *s_u, int P, int Cin, const float *__restrict__ w, int Cout,
                                         float *s_w, double *s_stat, T *__restrict__ out, int ldo,
                                         double *__restrict__ stats, int N, const TileCoord &tc, int D, int H, int W) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int vp = tid & 127, half = tid >> 7;
    // the two voxels of this thread
    size_t goff[2];
    bool valid[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int v = vp + i * 128;
        const int lx = v & 7, ly = (v >> 3) & 7, lz = v >> 6;
        const int gz = tc.z0 + lz, gy = tc.y0 + ly, gx = tc.x0 + lx;
        valid[i] = gz < D && gy < H && gx < W;
        goff[i] = ((((size_t)tc.n * D + gz) * H + gy) * W + gx) * (size_t)ldo;
    }
    for (int i = tid; i < 2 * Cout; i += NT) s_stat[i] = 0.0;
    constexpr int CB = 2 * CPT;  // output channels per pass
    for (int cb = 0; cb < Cout; cb += CB) {
        __syncthreads();  // previous pass done with s_w (and s_u/s_stat initialised on first pass)
        for (int i = tid; i < Cin * CB; i += NT) {
            const int j = i % CB, k = i / CB;
            s_w[i] = w[(size_t)(cb + j) * Cin + k];
        }
        __syncthreads();
        float a0[CPT], a1[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
        const float *u0 = s_u + (size_t)vp * P;
        const float *u1 = s_u + (size_t)(vp + 128) * P;
        const float *wp = s_w + half * CPT;
        for (int k = 0; k < Cin; ++k) {
            const float x0 = u0[k], x1 = u1[k];
#pragma unroll
            for (int j4 = 0; j4 < CPT; j4 += 4) {
                const float4 wv = *reinterpret_cast<const float4 *>(wp + (size_t)k * CB + j4);
                a0[j4 + 0] += x0 * wv.x; a0[j4 + 1] += x0 * wv.y; a0[j4 + 2] += x0 * wv.z; a0[j4 + 3] += x0 * wv.w;
                a1[j4 + 0] += x1 * wv.x; a1[j4 + 1] += x1 * wv.y; a1[j4 + 2] += x1 * wv.z; a1[j4 + 3] += x1 * wv.w;
            }
        }
        const int co = cb + half * CPT;
        // store + statistics on the values as stored
        float sv[CB];  // [0,CPT): sums, [CPT,2CPT): sums of squares
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const float r0 = valid[0] ? round_as(out, a0[j]) : 0.f;
            const float r1 = valid[1] ? round_as(out, a1[j]) : 0.f;
            sv[j] = r0 + r1;
            sv[CPT + j] = r0 * r0 + r1 * r1;
        }
#pragma unroll
        for (int j4 = 0; j4 < CPT; j4 += 4) {
            if (valid[0]) st4(out + goff[0] + co + j4, make_float4(a0[j4], a0[j4 + 1], a0[j4 + 2], a0[j4 + 3]));
            if (valid[1]) st4(out + goff[1] + co + j4, make_float4(a1[j4], a1[j4 + 1], a1[j4 + 2], a1[j4 + 3]));
        }
        warp_transpose_sum<CB>(sv, lane);
        {
            constexpr int DUP = 32 / CB;  // lanes sharing one value index
            if ((lane % DUP) == 0) {
                const int idx = warp_transpose_owner<CB>(lane);  // 0..CB-1
                const int isq = idx >= CPT;
                const int c = co + (isq ? idx - CPT : idx);
                atomicAdd(&s_stat[isq * Cout + c], (double)sv[0]);      // double from the first run-order-dependent addition on
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < 2 * Cout; i += NT) {
        const int isq = i >= Cout;
        const int c = isq ? i - Cout : i;
        atomicAdd(&stats[(size_t)isq * N * Cout + (size_t)tc.n * Cout + c], s_stat[i]);
    }
}

// -------------------------------------------------------------------------------------------
template <typename T, int CPT>
__global__ void __launch_bounds__(NT) dwpw_fwd_kernel(
    const T *__restrict__ x, int ldx, int Cin, NormDev xn, int N, int D, int H, int W,
    const float *__restrict__ dw_w, const float *__restrict__ pw_w, const float *__restrict__ sc_w, int Cout,
    T *__restrict__ t, int ldt, double *__restrict__ t_stats,
    T *__restrict__ r, int ldr, double *__restrict__ r_stats,
    T *__restrict__ u, int ldu, int vec_ok) {
    extern __shared__ __align__(16) float smem[];
    const int P = Cin | 1;
    float *s_in = smem;                        // HVOX*CK
    float *s_u = s_in + HVOX * CK;             // TV*P
    float *s_w = s_u + (size_t)TV * P;         // Cin*2*CPT   (16B aligned: see host-side padding)
    s_w = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(s_w) + 15) & ~(uintptr_t)15);
    float *s_scale = s_w + (size_t)Cin * 2 * CPT;
    float *s_shift = s_scale + Cin;
    double *s_stat = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(s_shift + Cin) + 7) & ~(uintptr_t)7);   // 2*Cout doubles

    const TileCoord tc = decode_tile(D, H, W);
    const int tid = threadIdx.x;
    setup_prologue(xn, N, Cin, tc.n, s_scale, s_shift);
    __syncthreads();

    if (dw_w != nullptr) {
        const int c = tid & 15, g = tid >> 4;
        const int lz = (g & 1) + 2 * (g >> 3);
        const int ly0 = 2 * ((g >> 1) & 3);
        for (int c0 = 0; c0 < Cin; c0 += CK) {
            if (c0 > 0) __syncthreads();  // everyone finished reading the previous chunk
            load_halo_chunk<T>(x, ldx, Cin, c0, vec_ok != 0, tc, D, H, W, s_scale, s_shift, xn.slope, s_in);
            __syncthreads();
            if (c0 + c < Cin) {
                float wreg[27];
#pragma unroll
                for (int k = 0; k < 27; ++k) wreg[k] = dw_w[(size_t)(c0 + c) * 27 + k];
                float acc0[TX], acc1[TX];
#pragma unroll
                for (int i = 0; i < TX; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }
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
                            for (int i = 0; i < TX; ++i) acc0[i] += w0 * row[i] + w1 * row[i + 1] + w2 * row[i + 2];
                        }
                        if (hy >= 1) {
                            const float w0 = wreg[dz * 9 + (hy - 1) * 3], w1 = wreg[dz * 9 + (hy - 1) * 3 + 1], w2 = wreg[dz * 9 + (hy - 1) * 3 + 2];
#pragma unroll
                            for (int i = 0; i < TX; ++i) acc1[i] += w0 * row[i] + w1 * row[i + 1] + w2 * row[i + 2];
                        }
                    }
                }
                float *up = s_u + (size_t)((lz * TY + ly0) * TX) * P + c0 + c;
#pragma unroll
                for (int i = 0; i < TX; ++i) {
                    up[(size_t)i * P] = acc0[i];
                    up[(size_t)(TX + i) * P] = acc1[i];
                }
            }
        }
    } else {
        load_center_all<T>(x, ldx, Cin, vec_ok != 0, tc, D, H, W, s_scale, s_shift, xn.slope, s_u, P);
    }
    __syncthreads();

    // optional save of the depthwise output for the backward pass (values as the GEMM consumes them)
    if (u != nullptr) {
        for (int item = tid; item < TV * Cin; item += NT) {
            const int k = item % Cin, v = item / Cin;
            const int lx = v & 7, ly = (v >> 3) & 7, lz = v >> 6;
            const int gz = tc.z0 + lz, gy = tc.y0 + ly, gx = tc.x0 + lx;
            if (gz < D && gy < H && gx < W)
                st1(u + ((((size_t)tc.n * D + gz) * H + gy) * W + gx) * (size_t)ldu + k, s_u[(size_t)v * P + k]);
        }
    }

    pw_phase<T, CPT>(s_u, P, Cin, pw_w, Cout, s_w, s_stat, t, ldt, t_stats, N, tc, D, H, W);

    if (sc_w != nullptr) {
        __syncthreads();  // all reads of s_u by the main GEMM are done
        if (dw_w != nullptr)
            load_center_all<T>(x, ldx, Cin, vec_ok != 0, tc, D, H, W, s_scale, s_shift, xn.slope, s_u, P);
        __syncthreads();
        pw_phase<T, CPT>(s_u, P, Cin, sc_w, Cout, s_w, s_stat, r, ldr, r_stats, N, tc, D, H, W);
    }
}

static size_t dwpw_smem_bytes(int Cin, int Cout, int CPT) {
    const size_t P = (size_t)(Cin | 1);
    size_t fl = (size_t)HVOX * CK + (size_t)TV * P + 4 /*align slack*/ + (size_t)Cin * 2 * CPT + 2 * (size_t)Cin + 2 + 4 * (size_t)Cout;
    return fl * sizeof(float);
}

// -------------------------------------------------------------------------------------------
// Dense / grouped 3x3x3 direct convolution on CUDA cores (generic path; every channel count).
constexpr int C3_CK = 8;   // input channels per pass
template <typename T, int CC>  // CC output channels per pass, each thread: 1 voxel x CC outputs
__global__ void __launch_bounds__(NT) conv3_fwd_kernel(
    const T *__restrict__ x, int ldx, int Cin, NormDev xn, int N, int D, int H, int W,
    const float *__restrict__ wgt, int groups, int Cout, T *__restrict__ t, int ldt, double *__restrict__ t_stats) {
    extern __shared__ __align__(16) float smem[];
    float *s_in = smem;                                   // C3_CK * HZ*HY*HX (channel-major)
    float *s_w = s_in + C3_CK * HZ * HY * HX;             // 27 * C3_CK * CC
    float *s_scale = s_w + 27 * C3_CK * CC;
    float *s_shift = s_scale + Cin;
    double *s_stat = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(s_shift + Cin) + 7) & ~(uintptr_t)7);   // 2*Cout doubles

    const TileCoord tc = decode_tile(D, H, W);
    const int tid = threadIdx.x, lane = tid & 31;
    setup_prologue(xn, N, Cin, tc.n, s_scale, s_shift);
    for (int i = tid; i < 2 * Cout; i += NT) s_stat[i] = 0.0;

    const int v = tid;
    const int lx = v & 7, ly = (v >> 3) & 7, lz = v >> 6;
    const int gz = tc.z0 + lz, gy = tc.y0 + ly, gx = tc.x0 + lx;
    const bool valid = gz < D && gy < H && gx < W;
    const size_t goff = ((((size_t)tc.n * D + gz) * H + gy) * W + gx) * (size_t)ldt;
    const int cin_g = Cin / groups, cout_g = Cout / groups;

    for (int cb = 0; cb < Cout; cb += CC) {
        float acc[CC];
#pragma unroll
        for (int j = 0; j < CC; ++j) acc[j] = 0.f;
        // input channels that can reach outputs [cb, cb+CC)
        const int g_lo = cb / cout_g, g_hi = (min(cb + CC, Cout) - 1) / cout_g;
        const int ci_lo = g_lo * cin_g, ci_hi = (g_hi + 1) * cin_g;
        for (int c0 = ci_lo; c0 < ci_hi; c0 += C3_CK) {
            __syncthreads();
            // stage activated input, channel-major
            for (int item = tid; item < C3_CK * HZ * HY * HX; item += NT) {
                const int ci = item / (HZ * HY * HX);
                int hv = item % (HZ * HY * HX);
                const int hx = hv % HX; hv /= HX;
                const int hy = hv % HY;
                const int hz = hv / HY;
                const int iz = tc.z0 + hz - 1, iy = tc.y0 + hy - 1, ix = tc.x0 + hx - 1;
                const int c = c0 + ci;
                float val = 0.f;
                if (c < ci_hi && iz >= 0 && iz < D && iy >= 0 && iy < H && ix >= 0 && ix < W) {
                    val = ld1(x + ((((size_t)tc.n * D + iz) * H + iy) * W + ix) * (size_t)ldx + c);
                    val = lrelu(val * s_scale[c] + s_shift[c], xn.slope);
                }
                s_in[item] = val;
            }
            // stage weights [tap][ci][j], expanding groups to a zero-filled dense block
            for (int item = tid; item < 27 * C3_CK * CC; item += NT) {
                const int j = item % CC;
                const int ci = (item / CC) % C3_CK;
                const int tap = item / (CC * C3_CK);
                const int co = cb + j, c = c0 + ci;
                float wv = 0.f;
                if (co < Cout && c < ci_hi) {
                    const int g = co / cout_g;
                    const int cl = c - g * cin_g;
                    if (cl >= 0 && cl < cin_g) wv = wgt[((size_t)co * cin_g + cl) * 27 + tap];
                }
                s_w[item] = wv;
            }
            __syncthreads();
#pragma unroll 1
            for (int ci = 0; ci < C3_CK; ++ci) {
                const float *ip = s_in + ci * (HZ * HY * HX) + (lz * HY + ly) * HX + lx;
#pragma unroll
                for (int tap = 0; tap < 27; ++tap) {
                    const int dz = tap / 9, dy = (tap / 3) % 3, dx = tap % 3;
                    const float a = ip[(dz * HY + dy) * HX + dx];
                    const float *wp = s_w + (tap * C3_CK + ci) * CC;
#pragma unroll
                    for (int j4 = 0; j4 < CC; j4 += 4) {
                        const float4 wv = *reinterpret_cast<const float4 *>(wp + j4);
                        acc[j4] += a * wv.x; acc[j4 + 1] += a * wv.y; acc[j4 + 2] += a * wv.z; acc[j4 + 3] += a * wv.w;
                    }
                }
            }
        }
        // epilogue for this output-channel pass
        float sv[CC];
#pragma unroll
        for (int j = 0; j < CC; ++j) sv[j] = valid ? round_as(t, acc[j]) : 0.f;
        if (valid) {
#pragma unroll
            for (int j4 = 0; j4 < CC; j4 += 4)
                if (cb + j4 < Cout) st4(t + goff + cb + j4, make_float4(acc[j4], acc[j4 + 1], acc[j4 + 2], acc[j4 + 3]));
        }
        float sq[CC];
#pragma unroll
        for (int j = 0; j < CC; ++j) sq[j] = sv[j] * sv[j];
        warp_transpose_sum<CC>(sv, lane);
        warp_transpose_sum<CC>(sq, lane);
        constexpr int DUP = 32 / CC;
        if ((lane % DUP) == 0) {
            const int c = cb + warp_transpose_owner<CC>(lane);
            if (c < Cout) {
                atomicAdd(&s_stat[c], (double)sv[0]);
                atomicAdd(&s_stat[Cout + c], (double)sq[0]);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < 2 * Cout; i += NT) {
        const int isq = i >= Cout;
        const int c = isq ? i - Cout : i;
        atomicAdd(&t_stats[(size_t)isq * N * Cout + (size_t)tc.n * Cout + c], s_stat[i]);
    }
}

// -------------------------------------------------------------------------------------------
// Residual merge (+ optional 2x2x2 max-pool of the result).  One thread = one 2x2x2 cell x one 16-byte channel
// vector; grid.y = sample, per-(n,c) scale/shift tables in shared memory; all loads of a cell are issued first.
// R1: the shortcut tensor is not materialised -- it is the rank-1 map r[v][c] = r1_w[c] * x[v] of a single-channel
// tensor x (the first block's 1x1x1 shortcut conv of a 1-channel image): `r` then points at x (1 channel, stride ldr).
// ncu (325 windows): at 97-121 registers only two CTAs (16 warps, 24 % of the SM's warp slots) are resident and the kernel
// sits on long-scoreboard stalls at 3.7 TB/s.  With the scale / shift tables read from shared memory both variants fit
// three CTAs in 80 registers without spills (rank-1: 0.68 -> 0.57 ms at 4.4 TB/s; four CTAs at 64 registers spill and
// are no faster).
template <typename T, bool R1>
__global__ void __launch_bounds__(256, 3) merge_fwd_kernel(
    const T *__restrict__ t2, int ld2, NormDev n2, const T *__restrict__ r, int ldr, NormDev nr, const float *__restrict__ r1_w,
    int N, int C, int D, int H, int W, float slope,
    T *__restrict__ out, int ldo, T *__restrict__ pooled, int ldp) {
    constexpr int V = VecW<T>::V;
    extern __shared__ float sm[];
    float *s_sc2 = sm, *s_sh2 = sm + C, *s_scr = sm + 2 * C, *s_shr = sm + 3 * C;
    const int n = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        norm_scale_shift(n2, N, C, n, c, s_sc2[c], s_sh2[c]);
        norm_scale_shift(nr, N, C, n, c, s_scr[c], s_shr[c]);
        s_sh2[c] += s_shr[c];                       // the two shifts only ever appear as their sum
        if (R1) s_scr[c] *= r1_w[c];
    }
    __syncthreads();
    const int CD = (D + 1) / 2, CH = (H + 1) / 2, CW = (W + 1) / 2, CQ = C / V;
    const int PD = D / 2, PH = H / 2, PW = W / 2;
    // per-sample cell count (grid.y = sample): 32-bit index arithmetic (the host checks the range) -- the 64-bit divisions
    // this loop started with are ~100 instructions each, ahead of the first load
    const uint32_t total = (uint32_t)CD * CH * CW * CQ;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        uint32_t rem = idx;
        const int q = (int)(rem % (uint32_t)CQ); rem /= (uint32_t)CQ;
        const int cx = (int)(rem % (uint32_t)CW); rem /= (uint32_t)CW;
        const int cy = (int)(rem % (uint32_t)CH);
        const int cz = (int)(rem / (uint32_t)CH);
        const int c = q * V;
        // scale / shift tables are read from shared memory where they are used (24 registers less: one more CTA per SM)
        const float *sc2 = s_sc2 + c, *sh2 = s_sh2 + c, *scr = s_scr + c;
        float mx[V];
#pragma unroll
        for (int j = 0; j < V; ++j) mx[j] = -INFINITY;
#pragma unroll
        for (int half = 0; half < 2; ++half) {        // one z-plane of the cell at a time: 8 raw loads in flight
            uint4 ra[4], rb[4];
            float xr[4];
            bool ok[4];
            size_t vox[4];
            const int z = cz * 2 + half;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int y = cy * 2 + (k >> 1), xx = cx * 2 + (k & 1);
                ok[k] = z < D && y < H && xx < W;
                vox[k] = (((size_t)n * D + z) * H + y) * W + xx;
                if (ok[k]) {
                    ra[k] = ldraw(t2 + vox[k] * ld2 + c);
                    if (R1) xr[k] = ld1(r + vox[k] * ldr); else rb[k] = ldraw(r + vox[k] * ldr + c);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (ok[k]) {
                    float a[V], b[V], o[V];
                    cvt_raw(t2, ra[k], a);
                    if (R1) {
#pragma unroll
                        for (int j = 0; j < V; ++j) b[j] = xr[k];
                    } else {
                        cvt_raw(t2, rb[k], b);
                    }
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        o[j] = lrelu(fmaf(a[j], sc2[j], fmaf(b[j], scr[j], sh2[j])), slope);
                        // pool the values as stored, so the backward arg-max sees the same numbers
                        mx[j] = fmaxf(mx[j], round_as(t2, o[j]));
                    }
                    if (out != nullptr) stv(out + vox[k] * ldo + c, o);
                }
            }
        }
        if (pooled != nullptr && cz < PD && cy < PH && cx < PW) {
            const size_t pv = (((size_t)n * PD + cz) * PH + cy) * PW + cx;
            stv(pooled + pv * ldp + c, mx);
        }
    }
}

// The same with one thread per x-COLUMN of a cell (its 2 x 2 (z, y) voxels) and 16-byte channel vector: the lanes of a
// warp then cover CONSECUTIVE voxels of a row, so a load instruction reads 512 contiguous bytes (the cell-per-thread
// mapping above strides over every other voxel: ncu, 325 windows at 48^3, counted 64 % of the peak LSU wavefront rate
// at 53 % of the DRAM rate -- half of every 128-byte line requested by an instruction belonged to the next one).  The
// pooled maximum needs the other x-column of the cell: one shuffle with lane ^ CQ (CQ = C / V lanes per voxel, a power of
// two <= 16; the host falls back to the cell-per-thread kernel otherwise).
template <typename T, bool R1>
__global__ void __launch_bounds__(256, 3) merge_col_fwd_kernel(
    const T *__restrict__ t2, int ld2, NormDev n2, const T *__restrict__ r, int ldr, NormDev nr, const float *__restrict__ r1_w,
    int N, int C, int D, int H, int W, float slope,
    T *__restrict__ out, int ldo, T *__restrict__ pooled, int ldp) {
    constexpr int V = VecW<T>::V;
    extern __shared__ float sm[];
    float *s_sc2 = sm, *s_sh2 = sm + C, *s_scr = sm + 2 * C, *s_shr = sm + 3 * C;
    const int n = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        norm_scale_shift(n2, N, C, n, c, s_sc2[c], s_sh2[c]);
        norm_scale_shift(nr, N, C, n, c, s_scr[c], s_shr[c]);
        s_sh2[c] += s_shr[c];
        if (R1) s_scr[c] *= r1_w[c];
    }
    __syncthreads();
    const int CD = (D + 1) / 2, CH = (H + 1) / 2, CW2 = 2 * ((W + 1) / 2), CQ = C / V;
    const int PD = D / 2, PH = H / 2, PW = W / 2;
    const int cq_sh = __ffs(CQ) - 1;
    const uint32_t total = (uint32_t)CD * CH * CW2 * CQ;             // a multiple of 2 * CQ; rounded up to whole warps below
    const uint32_t total_round = (total + 31u) & ~31u;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total_round; idx += gridDim.x * blockDim.x) {
        const bool in_range = idx < total;
        uint32_t rem = idx;
        const int q = (int)(rem & (uint32_t)(CQ - 1)); rem >>= cq_sh;
        const int xx = (int)(rem % (uint32_t)CW2); rem /= (uint32_t)CW2;
        const int cy = (int)(rem % (uint32_t)CH);
        const int cz = (int)(rem / (uint32_t)CH);
        const int c = q * V;
        const float *sc2 = s_sc2 + c, *sh2 = s_sh2 + c, *scr = s_scr + c;
        float mx[V];
#pragma unroll
        for (int j = 0; j < V; ++j) mx[j] = -INFINITY;
        uint4 ra[4], rb[4];
        float xr[4];
        bool ok[4];
        size_t vox[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                 // all raw loads of the column are issued first
            const int z = cz * 2 + (k >> 1), y = cy * 2 + (k & 1);
            ok[k] = in_range && z < D && y < H && xx < W;
            vox[k] = (((size_t)n * D + z) * H + y) * W + xx;
            if (ok[k]) {
                ra[k] = ldraw(t2 + vox[k] * ld2 + c);
                if (R1) xr[k] = ld1(r + vox[k] * ldr); else rb[k] = ldraw(r + vox[k] * ldr + c);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (ok[k]) {
                float a[V], b[V], o[V];
                cvt_raw(t2, ra[k], a);
                if (R1) {
#pragma unroll
                    for (int j = 0; j < V; ++j) b[j] = xr[k];
                } else {
                    cvt_raw(t2, rb[k], b);
                }
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    o[j] = lrelu(fmaf(a[j], sc2[j], fmaf(b[j], scr[j], sh2[j])), slope);
                    mx[j] = fmaxf(mx[j], round_as(t2, o[j]));      // pool the values as stored (the backward arg-max sees them)
                }
                if (out != nullptr) stv(out + vox[k] * ldo + c, o);
            }
        }
        if (pooled != nullptr) {                        // warp-uniform: every lane takes part in the exchange
#pragma unroll
            for (int j = 0; j < V; ++j) mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], CQ));
            const int cx = xx >> 1;
            if (in_range && (xx & 1) == 0 && cz < PD && cy < PH && cx < PW) {
                const size_t pv = (((size_t)n * PD + cz) * PH + cy) * PW + cx;
                stv(pooled + pv * ldp + c, mx);
            }
        }
    }
}

// Residual merge + 1x1x1 head + sigmoid: one thread = one voxel, all channels (C <= 64).
template <typename T, int CMAX>
__global__ void __launch_bounds__(256) merge_head_fwd_kernel(
    const T *__restrict__ t2, int ld2, NormDev n2, const T *__restrict__ r, int ldr, NormDev nr,
    int N, int C, size_t nvox, float slope, T *__restrict__ out, int ldo,
    const float *__restrict__ head_w, const float *__restrict__ head_b, int OC,
    float *__restrict__ prob, float *__restrict__ logits) {
    constexpr int V = VecW<T>::V;
    __shared__ float s_sc2[CMAX], s_sh2[CMAX], s_scr[CMAX], s_hw[4 * CMAX];
    const int n = blockIdx.y;
    for (int c = threadIdx.x; c < CMAX; c += blockDim.x) {
        float a = 0.f, b = 0.f, cc = 0.f, d = 0.f;
        if (c < C) {
            norm_scale_shift(n2, N, C, n, c, a, b);
            norm_scale_shift(nr, N, C, n, c, cc, d);
        }
        s_sc2[c] = a; s_sh2[c] = b + d; s_scr[c] = cc;
        for (int oc = 0; oc < 4; ++oc) s_hw[oc * CMAX + c] = (c < C && oc < OC) ? head_w[(size_t)oc * C + c] : 0.f;
    }
    __syncthreads();
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
        const size_t vox = (size_t)n * nvox + v;
        float o[CMAX];
        uint4 ra[CMAX / V], rb[CMAX / V];
#pragma unroll
        for (int i = 0; i < CMAX / V; ++i) {
            if (i * V < C) { ra[i] = ldraw(t2 + vox * ld2 + i * V); rb[i] = ldraw(r + vox * ldr + i * V); }
        }
#pragma unroll
        for (int i = 0; i < CMAX / V; ++i) {
            if (i * V < C) {
                float a[V], b[V], ov[V];
                cvt_raw(t2, ra[i], a);
                cvt_raw(t2, rb[i], b);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    ov[j] = lrelu(fmaf(a[j], s_sc2[i * V + j], fmaf(b[j], s_scr[i * V + j], s_sh2[i * V + j])), slope);
                    o[i * V + j] = ov[j];
                }
                if (out != nullptr) stv(out + vox * ldo + i * V, ov);
            }
        }
        for (int oc = 0; oc < OC; ++oc) {
            float acc = head_b[oc];
            if (oc < 4) {
#pragma unroll
                for (int c = 0; c < CMAX; ++c)
                    if (c < C) acc = fmaf(round_as(t2, o[c]), s_hw[oc * CMAX + c], acc);
            } else {
#pragma unroll
                for (int c = 0; c < CMAX; ++c)
                    if (c < C) acc = fmaf(round_as(t2, o[c]), head_w[(size_t)oc * C + c], acc);
            }
            const size_t oi = ((size_t)n * OC + oc) * nvox + v;
            if (logits != nullptr) logits[oi] = acc;
            prob[oi] = 1.f / (1.f + expf(-acc));
        }
    }
}

// The same for C = 16 in bf16 storage (the configured model's head, unet3d.py:201-202,220-221): thread = HALF a voxel
// (8 channels = one 16-byte vector per tensor), so every load / store instruction of a warp covers 512 contiguous bytes,
// the lane pair adds its two partial head sums with one shuffle, and 40 registers keep the SM's warp slots full (the
// voxel-per-thread kernel above: 119 registers, 24 % of the warp slots, 3.6 TB/s).  Two voxel halves per thread in flight.
template <int U>
__global__ void __launch_bounds__(256, (U > 2 ? 3 : 4)) merge_head16_fwd_kernel(
    const h16 *__restrict__ t2, int ld2, NormDev n2, const h16 *__restrict__ r, int ldr, NormDev nr,
    int N, size_t nvox, float slope, h16 *__restrict__ out, int ldo,
    const float *__restrict__ head_w, const float *__restrict__ head_b, int OC,
    float *__restrict__ prob, float *__restrict__ logits) {
    constexpr int C = 16;
    __shared__ float s_sc2[C], s_sh2[C], s_scr[C], s_hw[4 * C];
    const int n = blockIdx.y;
    if (threadIdx.x < C) {
        const int c = threadIdx.x;
        float a, b, cc, d;
        norm_scale_shift(n2, N, C, n, c, a, b);
        norm_scale_shift(nr, N, C, n, c, cc, d);
        s_sc2[c] = a; s_sh2[c] = b + d; s_scr[c] = cc;
        for (int oc = 0; oc < 4; ++oc) s_hw[oc * C + c] = oc < OC ? head_w[(size_t)oc * C + c] : 0.f;
    }
    __syncthreads();
    const int h = threadIdx.x & 1;
    const float *sc2 = s_sc2 + h * 8, *sh2 = s_sh2 + h * 8, *scr = s_scr + h * 8;     // read from shared memory in the loop: no spills at 64 registers
    const size_t total = 2 * nvox, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += U * stride) {
        uint4 ra[U], rb[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + u * stride;
            ok[u] = i < total;                       // both lanes of a pair share the voxel, so they agree
            if (ok[u]) {
                const size_t vox = (size_t)n * nvox + (i >> 1);
                ra[u] = ldraw(t2 + vox * ld2 + h * 8);
                rb[u] = ldraw(r + vox * ldr + h * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const size_t v = (i0 + u * stride) >> 1, vox = (size_t)n * nvox + v;
            float a[8], b[8], o[8];
            cvt_raw(t2, ra[u], a);
            cvt_raw(t2, rb[u], b);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = lrelu(fmaf(a[j], sc2[j], fmaf(b[j], scr[j], sh2[j])), slope);
            if (out != nullptr) stv(out + vox * ldo + h * 8, o);
            for (int oc = 0; oc < OC; ++oc) {
                float acc = 0.f;
                if (oc < 4) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc = fmaf(round_as(t2, o[j]), s_hw[oc * C + h * 8 + j], acc);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc = fmaf(round_as(t2, o[j]), head_w[(size_t)oc * C + h * 8 + j], acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                if (h == 0) {
                    acc += head_b[oc];
                    const size_t oi = ((size_t)n * OC + oc) * nvox + v;
                    if (logits != nullptr) logits[oi] = acc;
                    prob[oi] = 1.f / (1.f + expf(-acc));
                }
            }
        }
    }
}

// ---- 256-bit global accesses (LDG / STG.E.256, sm_100): one instruction moves the 16 fp16 channels of a voxel ----------
struct U8 { uint32_t w[8]; };
__device__ __forceinline__ U8 ld256(const h16 *p) {
    U8 r;
    asm("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(h16 *p, const U8 &v) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]),
                 "r"(v.w[4]), "r"(v.w[5]), "r"(v.w[6]), "r"(v.w[7]) : "memory");
}

// Column-per-thread residual merge (+ MaxPool3d(2)) in fp16 storage with 32-byte channel vectors: thread = one x-column of a
// cell (2 x 2 voxels in z, y) x 16 channels, C / 16 lanes per voxel (a power of two <= 16).  Half the load / store
// instructions of merge_col_fwd_kernel<h16>; same arithmetic, bit-identical results.
__global__ void __launch_bounds__(256, 2) merge_col32_fwd_kernel(
    const h16 *__restrict__ t2, int ld2, NormDev n2, const h16 *__restrict__ r, int ldr, NormDev nr,
    int N, int C, int D, int H, int W, float slope, h16 *__restrict__ out, int ldo, h16 *__restrict__ pooled, int ldp) {
    extern __shared__ float sm[];
    float *s_sc2 = sm, *s_sh2 = sm + C, *s_scr = sm + 2 * C, *s_shr = sm + 3 * C;
    const int n = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        norm_scale_shift(n2, N, C, n, c, s_sc2[c], s_sh2[c]);
        norm_scale_shift(nr, N, C, n, c, s_scr[c], s_shr[c]);
        s_sh2[c] += s_shr[c];
    }
    __syncthreads();
    const int CD = (D + 1) / 2, CH = (H + 1) / 2, CW2 = 2 * ((W + 1) / 2), CQ = C / 16;
    const int PD = D / 2, PH = H / 2, PW = W / 2;
    const int cq_sh = __ffs(CQ) - 1;
    const uint32_t total = (uint32_t)CD * CH * CW2 * CQ;
    const uint32_t total_round = (total + 31u) & ~31u;
    const uint32_t HW = (uint32_t)H * W;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total_round; idx += gridDim.x * blockDim.x) {
        const bool in_range = idx < total;
        uint32_t rem = idx;
        const int q = (int)(rem & (uint32_t)(CQ - 1)); rem >>= cq_sh;
        const int xx = (int)(rem % (uint32_t)CW2); rem /= (uint32_t)CW2;
        const int cy = (int)(rem % (uint32_t)CH);
        const int cz = (int)(rem / (uint32_t)CH);
        const int c = q * 16;
        const float *sc2 = s_sc2 + c, *sh2 = s_sh2 + c, *scr = s_scr + c;
        const size_t v00 = (((size_t)n * D + cz * 2) * H + cy * 2) * W + xx;
        U8 ra[4], rb[4];
        bool ok[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int z = cz * 2 + (k >> 1), y = cy * 2 + (k & 1);
            ok[k] = in_range && z < D && y < H && xx < W;
            const size_t vk = v00 + (k >> 1) * HW + (k & 1) * W;
            if (ok[k]) { ra[k] = ld256(t2 + vk * ld2 + c); rb[k] = ld256(r + vk * ldr + c); }
        }
        __half2 mx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) mx[j] = __float2half2_rn(-INFINITY);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (ok[k]) {
                U8 pk;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float a0 = h16_lo(ra[k].w[j]), a1 = h16_hi(ra[k].w[j]), b0 = h16_lo(rb[k].w[j]), b1 = h16_hi(rb[k].w[j]);
                    const float o0 = lrelu(fmaf(a0, sc2[2 * j], fmaf(b0, scr[2 * j], sh2[2 * j])), slope);
                    const float o1 = lrelu(fmaf(a1, sc2[2 * j + 1], fmaf(b1, scr[2 * j + 1], sh2[2 * j + 1])), slope);
                    pk.w[j] = pack_h16x2(o0, o1);
                    mx[j] = __hmax2(mx[j], *reinterpret_cast<const __half2 *>(&pk.w[j]));     // pool the values as stored
                }
                if (out != nullptr) st256(out + (v00 + (k >> 1) * HW + (k & 1) * W) * ldo + c, pk);
            }
        }
        if (pooled != nullptr) {                        // warp-uniform: every lane takes part in the exchange
            U8 pm;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(&mx[j]);
                const uint32_t o = __shfl_xor_sync(0xffffffffu, w, CQ);
                const __half2 m = __hmax2(mx[j], *reinterpret_cast<const __half2 *>(&o));
                pm.w[j] = *reinterpret_cast<const uint32_t *>(&m);
            }
            const int cx = xx >> 1;
            if (in_range && (xx & 1) == 0 && cz < PD && cy < PH && cx < PW) {
                const size_t pv = (((size_t)n * PD + cz) * PH + cy) * PW + cx;
                st256(pooled + pv * ldp + c, pm);
            }
        }
    }
}

// Residual merge + head for C = 16, fp16 storage, one thread per VOXEL: two 32-byte loads and (optionally) one 32-byte
// store per voxel, no lane-pair exchange for the head sum.  Needs 32-byte aligned voxels (host check).
__global__ void __launch_bounds__(256, 3) merge_head16v_fwd_kernel(
    const h16 *__restrict__ t2, int ld2, NormDev n2, const h16 *__restrict__ r, int ldr, NormDev nr,
    int N, size_t nvox, float slope, h16 *__restrict__ out, int ldo,
    const float *__restrict__ head_w, const float *__restrict__ head_b, int OC,
    float *__restrict__ prob, float *__restrict__ logits) {
    constexpr int C = 16;
    __shared__ __align__(16) float s_sc2[C], s_sh2[C], s_scr[C], s_hw[4 * C];
    const int n = blockIdx.y;
    if (threadIdx.x < C) {
        const int c = threadIdx.x;
        float a, b, cc, d;
        norm_scale_shift(n2, N, C, n, c, a, b);
        norm_scale_shift(nr, N, C, n, c, cc, d);
        s_sc2[c] = a; s_sh2[c] = b + d; s_scr[c] = cc;
        for (int oc = 0; oc < 4; ++oc) s_hw[oc * C + c] = oc < OC ? head_w[(size_t)oc * C + c] : 0.f;
    }
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvox; i0 += 2 * stride) {
        U8 ra[2], rb[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const size_t i = i0 + u * stride;
            ok[u] = i < nvox;
            if (ok[u]) {
                const size_t vox = (size_t)n * nvox + i;
                ra[u] = ld256(t2 + vox * ld2);
                rb[u] = ld256(r + vox * ldr);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!ok[u]) continue;
            const size_t v = i0 + u * stride, vox = (size_t)n * nvox + v;
            float o[16];
            U8 pk;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float a0 = h16_lo(ra[u].w[j]), a1 = h16_hi(ra[u].w[j]), b0 = h16_lo(rb[u].w[j]), b1 = h16_hi(rb[u].w[j]);
                const float o0 = lrelu(fmaf(a0, s_sc2[2 * j], fmaf(b0, s_scr[2 * j], s_sh2[2 * j])), slope);
                const float o1 = lrelu(fmaf(a1, s_sc2[2 * j + 1], fmaf(b1, s_scr[2 * j + 1], s_sh2[2 * j + 1])), slope);
                pk.w[j] = pack_h16x2(o0, o1);
                o[2 * j] = h16_lo(pk.w[j]); o[2 * j + 1] = h16_hi(pk.w[j]);      // the head sees the values as stored
            }
            if (out != nullptr) st256(out + vox * ldo, pk);
            for (int oc = 0; oc < OC; ++oc) {
                float acc = 0.f;
                if (oc < 4) {
                    // same summation order as the half-voxel kernel: channels 0-7 and 8-15 separately, then added
                    float lo = 0.f, hi = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { lo = fmaf(o[j], s_hw[oc * C + j], lo); hi = fmaf(o[8 + j], s_hw[oc * C + 8 + j], hi); }
                    acc = lo + hi;
                } else {
                    float lo = 0.f, hi = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { lo = fmaf(o[j], head_w[(size_t)oc * C + j], lo); hi = fmaf(o[8 + j], head_w[(size_t)oc * C + 8 + j], hi); }
                    acc = lo + hi;
                }
                acc += head_b[oc];
                const size_t oi = ((size_t)n * OC + oc) * nvox + v;
                if (logits != nullptr) logits[oi] = acc;
                prob[oi] = 1.f / (1.f + expf(-acc));
            }
        }
    }
}

// Rank-1 residual merge (+ MaxPool3d(2)) for C = 16, fp16 storage: one thread per x-column of a cell (its 2 x 2 (z, y)
// voxels), all 16 channels -- a warp's load covers 32 consecutive voxels (1 KB), every store is a whole 32-byte sector of
// the concat buffer, and the pooled maximum needs one exchange with the neighbouring lane.
__global__ void __launch_bounds__(256, 3) merge_col16_r1_fwd_kernel(
    const h16 *__restrict__ t2, int ld2, NormDev n2, const h16 *__restrict__ x, int ldx, NormDev nr, const float *__restrict__ r1_w,
    int N, int D, int H, int W, float slope, h16 *__restrict__ out, int ldo, h16 *__restrict__ pooled, int ldp) {
    constexpr int C = 16;
    __shared__ __align__(16) float s_sc2[C], s_sh2[C], s_scr[C];
    const int n = blockIdx.y;
    if (threadIdx.x < C) {
        const int c = threadIdx.x;
        float a, b, cc, d;
        norm_scale_shift(n2, N, C, n, c, a, b);
        norm_scale_shift(nr, N, C, n, c, cc, d);
        s_sc2[c] = a; s_sh2[c] = b + d; s_scr[c] = cc * r1_w[c];
    }
    __syncthreads();
    const int CD = (D + 1) / 2, CH = (H + 1) / 2, CW2 = 2 * ((W + 1) / 2);
    const int PD = D / 2, PH = H / 2, PW = W / 2;
    const uint32_t total = (uint32_t)CD * CH * CW2;
    const uint32_t total_round = (total + 31u) & ~31u;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total_round; idx += gridDim.x * blockDim.x) {
        const bool in_range = idx < total;
        uint32_t rem = idx;
        const int xx = (int)(rem % (uint32_t)CW2); rem /= (uint32_t)CW2;
        const int cy = (int)(rem % (uint32_t)CH);
        const int cz = (int)(rem / (uint32_t)CH);
        U8 ra[4];
        float xr[4];
        bool ok[4];
        const size_t v00 = (((size_t)n * D + cz * 2) * H + cy * 2) * W + xx;      // voxel k of the column: v00 + (k >> 1) * H * W + (k & 1) * W
        const uint32_t HW = (uint32_t)H * W;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int z = cz * 2 + (k >> 1), y = cy * 2 + (k & 1);
            ok[k] = in_range && z < D && y < H && xx < W;
            const size_t vk = v00 + (k >> 1) * HW + (k & 1) * W;
            if (ok[k]) {
                ra[k] = ld256(t2 + vk * ld2);
                xr[k] = __half2float(x[vk * ldx]);
            }
        }
        __half2 mx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) mx[j] = __float2half2_rn(-INFINITY);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (ok[k]) {
                U8 pk;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float a0 = h16_lo(ra[k].w[j]), a1 = h16_hi(ra[k].w[j]);
                    const float o0 = lrelu(fmaf(a0, s_sc2[2 * j], fmaf(xr[k], s_scr[2 * j], s_sh2[2 * j])), slope);
                    const float o1 = lrelu(fmaf(a1, s_sc2[2 * j + 1], fmaf(xr[k], s_scr[2 * j + 1], s_sh2[2 * j + 1])), slope);
                    pk.w[j] = pack_h16x2(o0, o1);
                    mx[j] = __hmax2(mx[j], *reinterpret_cast<const __half2 *>(&pk.w[j]));     // pool the values as stored
                }
                if (out != nullptr) st256(out + (v00 + (k >> 1) * HW + (k & 1) * W) * ldo, pk);
            }
        }
        if (pooled != nullptr) {                        // warp-uniform: every lane takes part in the exchange
            U8 pm;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t w = *reinterpret_cast<const uint32_t *>(&mx[j]);
                const uint32_t o = __shfl_xor_sync(0xffffffffu, w, 1);
                const __half2 m = __hmax2(mx[j], *reinterpret_cast<const __half2 *>(&o));
                pm.w[j] = *reinterpret_cast<const uint32_t *>(&m);
            }
            const int cx = xx >> 1;
            if (in_range && (xx & 1) == 0 && cz < PD && cy < PH && cx < PW) {
                const size_t pv = (((size_t)n * PD + cz) * PH + cy) * PW + cx;
                st256(pooled + pv * ldp, pm);
            }
        }
    }
}

// -------------------------------------------------------------------------------------------
// ConvTranspose3d k=2 s=2: every input voxel produces a 2x2x2 block of outputs.
constexpr int CT_VOX = 64;  // input voxels per CTA
template <typename T, int CPT>
__global__ void __launch_bounds__(NT) convt_fwd_kernel(
    const T *__restrict__ x, int ldx, int Cin, int N, int d, int h, int w_,
    const float *__restrict__ wgt, const float *__restrict__ bias, int Cout,
    T *__restrict__ out, int ldo, int OD, int OH, int OW, int oz, int oy, int ox, int vec_ok) {
    extern __shared__ __align__(16) float smem[];
    const int P = Cin | 1;
    float *s_x = smem;                       // CT_VOX * P
    float *s_w = s_x + (size_t)CT_VOX * P;   // Cin * Cout  (one tap)
    s_w = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(s_w) + 15) & ~(uintptr_t)15);
    const int tid = threadIdx.x;
    const size_t nvox = (size_t)N * d * h * w_;
    const size_t v0 = (size_t)blockIdx.x * CT_VOX;
    // stage inputs
    const int groups = (Cin + 3) / 4;
    for (int item = tid; item < CT_VOX * groups; item += NT) {
        const int q = item % groups, v = item / groups;
        const int c = q * 4;
        float val[4] = {0.f, 0.f, 0.f, 0.f};
        if (v0 + v < nvox) {
            const T *p = x + (v0 + v) * (size_t)ldx + c;
            if (vec_ok && c + 3 < Cin) {
                const float4 f = ld4(p);
                val[0] = f.x; val[1] = f.y; val[2] = f.z; val[3] = f.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (c + j < Cin) val[j] = ld1(p + j);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) if (c + j < Cin) s_x[(size_t)v * P + c + j] = val[j];
    }
    // thread -> (voxel, slice of CPT output channels)
    const int nslice = Cout / CPT;           // host guarantees CT_VOX * nslice is a multiple of... handled by loop
    for (int tap = 0; tap < 8; ++tap) {
        __syncthreads();
        for (int i = tid; i < Cin * Cout; i += NT) {
            const int co = i % Cout, ci = i / Cout;
            s_w[i] = wgt[((size_t)ci * Cout + co) * 8 + tap];
        }
        __syncthreads();
        const int dz = tap >> 2, dy = (tap >> 1) & 1, dx = tap & 1;
        for (int work = tid; work < CT_VOX * nslice; work += NT) {
            const int v = work % CT_VOX, sl = work / CT_VOX;
            const size_t gv = v0 + v;
            if (gv >= nvox) continue;
            float acc[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[j] = bias[sl * CPT + j];
            const float *xp = s_x + (size_t)v * P;
            for (int k = 0; k < Cin; ++k) {
                const float a = xp[k];
                const float *wp = s_w + (size_t)k * Cout + sl * CPT;
#pragma unroll
                for (int j4 = 0; j4 < CPT; j4 += 4) {
                    const float4 wv = *reinterpret_cast<const float4 *>(wp + j4);
                    acc[j4] += a * wv.x; acc[j4 + 1] += a * wv.y; acc[j4 + 2] += a * wv.z; acc[j4 + 3] += a * wv.w;
                }
            }
            size_t rem = gv;
            const int ix = (int)(rem % w_); rem /= w_;
            const int iy = (int)(rem % h); rem /= h;
            const int iz = (int)(rem % d);
            const int n = (int)(rem / d);
            const int Z = oz + 2 * iz + dz, Y = oy + 2 * iy + dy, X = ox + 2 * ix + dx;
            if (Z < 0 || Z >= OD || Y < 0 || Y >= OH || X < 0 || X >= OW) continue;
            T *op = out + ((((size_t)n * OD + Z) * OH + Y) * OW + X) * (size_t)ldo + sl * CPT;
#pragma unroll
            for (int j4 = 0; j4 < CPT; j4 += 4) st4(op + j4, make_float4(acc[j4], acc[j4 + 1], acc[j4 + 2], acc[j4 + 3]));
        }
    }
}


// -------------------------------------------------------------------------------------------
// Single-input-channel variant of the fused depthwise->pointwise(+shortcut) conv (the network's first conv,
// unet3d.py:168,209): K = 1, so the "GEMM" is an outer product t[v][co] = pw[co] * u[v] and the kernel is a pure
// HBM write stream.  thread = one voxel, all COUT outputs of t (and r); persistent CTAs.
// InstanceNorm statistics: sum_v t[v][co] = pw[co] * sum_v u[v] and sum_v t^2 = pw[co]^2 * sum_v u^2, so only the
// four scalars {sum u, sum u^2, sum x, sum x^2} are reduced per tile (in fp32 storage this is exact; in bf16 storage
// it ignores the zero-mean rounding of the stored values, a < 1e-5 relative effect on mean / variance).
template <typename T, int COUT>
__global__ void __launch_bounds__(NT) dwpw_c1_kernel(
    const T *__restrict__ x, int ldx, NormDev xn, int N, int D, int H, int W,
    const float *__restrict__ dw_w, const float *__restrict__ pw_w, const float *__restrict__ sc_w,
    T *__restrict__ t, int ldt, double *__restrict__ t_stats, T *__restrict__ r, int ldr, double *__restrict__ r_stats,
    T *__restrict__ u, int ldu) {
    __shared__ float s_in[HZ][HY][HX + 1];
    __shared__ double s_sum[4];       // double: the warps' partial sums meet in a run-dependent order
    __shared__ float s_w[27 + 2 * COUT];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 27; i += NT) s_w[i] = dw_w[i];
    for (int i = tid; i < COUT; i += NT) { s_w[27 + i] = pw_w[i]; s_w[27 + COUT + i] = sc_w != nullptr ? sc_w[i] : 0.f; }
    if (tid < 4) s_sum[tid] = 0.0;
    const int tilesX = (W + TX - 1) / TX, tilesY = (H + TY - 1) / TY, tilesZ = (D + TZ - 1) / TZ;
    const int tiles_per_sample = tilesX * tilesY * tilesZ;
    const int total_tiles = tiles_per_sample * N;
    const int per = (total_tiles + gridDim.x - 1) / gridDim.x;      // contiguous tile range: few sample changes
    const int tile_begin = blockIdx.x * per, tile_end = min(total_tiles, tile_begin + per);
    const int lx = tid & 7, ly = (tid >> 3) & 7, lz = tid >> 6;
    int cur_n = -1;
    float sc = 1.f, sh = 0.f;
    auto flush = [&](int n) {
        if (n < 0) return;
        for (int i = tid; i < 2 * COUT; i += NT) {
            const int isq = i >= COUT, c = isq ? i - COUT : i;
            const float wt = s_w[27 + c], wr = s_w[27 + COUT + c];
            atomicAdd(&t_stats[(size_t)isq * N * COUT + (size_t)n * COUT + c], isq ? (double)wt * wt * s_sum[1] : (double)wt * s_sum[0]);
            if (sc_w != nullptr)
                atomicAdd(&r_stats[(size_t)isq * N * COUT + (size_t)n * COUT + c], isq ? (double)wr * wr * s_sum[3] : (double)wr * s_sum[2]);
        }
        __syncthreads();
        if (tid < 4) s_sum[tid] = 0.0;
    };
    for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int n = tile / tiles_per_sample;
        int b = tile - n * tiles_per_sample;
        const int x0 = (b % tilesX) * TX; b /= tilesX;
        const int y0 = (b % tilesY) * TY; b /= tilesY;
        const int z0 = b * TZ;
        __syncthreads();                      // previous tile done with s_in, its sums are in s_sum
        if (n != cur_n) {
            flush(cur_n);
            cur_n = n;
            norm_scale_shift(xn, N, 1, n, 0, sc, sh);
        }
        for (int item = tid; item < HZ * HY * HX; item += NT) {
            int hv = item;
            const int hx = hv % HX; hv /= HX;
            const int hy = hv % HY;
            const int hz = hv / HY;
            const int gz = z0 + hz - 1, gy = y0 + hy - 1, gx = x0 + hx - 1;
            float v = 0.f;
            if (gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W)
                v = lrelu(ld1(x + ((((size_t)n * D + gz) * H + gy) * W + gx) * (size_t)ldx) * sc + sh, xn.slope);
            s_in[hz][hy][hx] = v;
        }
        __syncthreads();
        float uacc = 0.f;
#pragma unroll
        for (int dz = 0; dz < 3; ++dz)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) uacc = fmaf(s_w[dz * 9 + dy * 3 + dx], s_in[lz + dz][ly + dy][lx + dx], uacc);
        const float xc = s_in[lz + 1][ly + 1][lx + 1];
        const int gz = z0 + lz, gy = y0 + ly, gx = x0 + lx;
        const bool valid = gz < D && gy < H && gx < W;
        const size_t vox = (((size_t)n * D + gz) * H + gy) * W + gx;
        if (u != nullptr && valid) st1(u + vox * (size_t)ldu, uacc);
        {
            constexpr int V = VecW<T>::V;
            // this kernel is a pure write stream: lanes 2j / 2j+1 (x-adjacent voxels) exchange their inputs so that one store
            // instruction covers both 16-byte halves of a voxel's 32-byte sector (even lane: first vector, odd lane: second)
            const bool pair_ok = (V == 8) && (COUT % 16 == 0);
            const float uacc_p = __shfl_xor_sync(0xffffffffu, uacc, 1), xc_p = __shfl_xor_sync(0xffffffffu, xc, 1);
            const bool valid_p = __shfl_xor_sync(0xffffffffu, valid ? 1 : 0, 1) != 0;
            const bool odd = (lane & 1) != 0;
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                if (a == 1 && (sc_w == nullptr || r == nullptr)) break;     // r == NULL: statistics only
                const float *wv = s_w + 27 + a * COUT;
                const int ld = a == 0 ? ldt : ldr;
                T *own = (a == 0 ? t + vox * (size_t)ldt : r + vox * (size_t)ldr);
                if (pair_ok) {
                    const float in_e = odd ? (a == 0 ? uacc_p : xc_p) : (a == 0 ? uacc : xc);     // even voxel of the pair
                    const float in_o = odd ? (a == 0 ? uacc : xc) : (a == 0 ? uacc_p : xc_p);     // odd voxel of the pair
                    T *pe = odd ? own - ld : own, *po = odd ? own : own + ld;
                    const bool ve = odd ? valid_p : valid, vo = odd ? valid : valid_p;
#pragma unroll
                    for (int cb = 0; cb < COUT; cb += 16) {
                        const int c = cb + (odd ? 8 : 0);
                        float oe[V], oo[V];
#pragma unroll
                        for (int j = 0; j < V; ++j) { oe[j] = in_e * wv[c + j]; oo[j] = in_o * wv[c + j]; }
                        if (ve) stv(pe + c, oe);
                        if (vo) stv(po + c, oo);
                    }
                } else if (valid) {
                    const float in = a == 0 ? uacc : xc;
#pragma unroll
                    for (int cb = 0; cb < COUT; cb += V) {
                        float o[V];
#pragma unroll
                        for (int j = 0; j < V; ++j) o[j] = in * wv[cb + j];
                        stv(own + cb, o);
                    }
                }
            }
        }
        float s0 = valid ? uacc : 0.f, s2 = valid ? xc : 0.f;
        float s1 = s0 * s0, s3 = s2 * s2;
        s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
        if (lane == 0) { atomicAdd(&s_sum[0], (double)s0); atomicAdd(&s_sum[1], (double)s1); atomicAdd(&s_sum[2], (double)s2); atomicAdd(&s_sum[3], (double)s3); }
    }
    __syncthreads();
    flush(cur_n);
}

// Depthwise stage alone of the single-input-channel first conv: u = dw * act(x) as one fp32 channel, plus the analytic
// statistics of t[v][co] = pw[co] * u[v] and r[v][co] = sc[co] * x[v] (neither tensor is written).  For consumers that
// evaluate the rank-1 pointwise stage on the fly (l3d_dwpw_fwd_rank1, l3d_merge_fwd_rank1): 4 B / voxel instead of 2 * Cout.
// thread = a z-column of ZC1 output voxels at one (y, x); lanes run along the flattened (y, x) plane, so every load
// and the store are coalesced; each input plane is loaded once (9 values) and feeds up to three output planes.
constexpr int ZC1 = 8;
template <typename T>
__global__ void __launch_bounds__(256) dw_c1_kernel(const T *__restrict__ x, int ldx, NormDev xn, int N, int D, int H, int W,
                                                    const float *__restrict__ dw_w, const float *__restrict__ pw_w,
                                                    const float *__restrict__ sc_w, int Cout, float *__restrict__ u,
                                                    double *__restrict__ t_stats, double *__restrict__ r_stats) {
    __shared__ float s_w[27];
    __shared__ double s_sum[4];       // double: the warps' partial sums meet in a run-dependent order
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 27) s_w[tid] = dw_w[tid];
    if (tid < 4) s_sum[tid] = 0.0;
    __syncthreads();
    const int n = blockIdx.z, z0 = blockIdx.y * ZC1;
    const int p = blockIdx.x * 256 + tid;
    const bool pvalid = p < H * W;
    const int y = pvalid ? p / W : 0, xx = pvalid ? p - (p / W) * W : 0;
    float sc, sh;
    norm_scale_shift(xn, N, 1, n, 0, sc, sh);
    float wr[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) wr[k] = s_w[k];
    float acc[ZC1];
#pragma unroll
    for (int k = 0; k < ZC1; ++k) acc[k] = 0.f;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    bool okx[3], oky[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { okx[d] = xx + d - 1 >= 0 && xx + d - 1 < W; oky[d] = y + d - 1 >= 0 && y + d - 1 < H; }
#pragma unroll
    for (int zi = 0; zi < ZC1 + 2; ++zi) {
        const int gz = z0 + zi - 1;
        float v[9];
        const bool okz = pvalid && gz >= 0 && gz < D;
        const T *plane = x + (((size_t)n * D + (okz ? gz : 0)) * H) * W * (size_t)ldx;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                float a = 0.f;
                if (okz && oky[dy] && okx[dx]) a = lrelu(fmaf(ld1(plane + ((size_t)(y + dy - 1) * W + (xx + dx - 1)) * ldx), sc, sh), xn.slope);
                v[dy * 3 + dx] = a;
            }
#pragma unroll
        for (int dz = 0; dz < 3; ++dz) {
            const int k = zi - dz;                    // output plane z0 + k reads input plane z0 + k + dz - 1 = gz
            if (k >= 0 && k < ZC1) {
#pragma unroll
                for (int t9 = 0; t9 < 9; ++t9) acc[k] = fmaf(wr[dz * 9 + t9], v[t9], acc[k]);
            }
        }
        if (zi >= 1 && zi <= ZC1 && okz) { s2 += v[4]; s3 = fmaf(v[4], v[4], s3); }
    }
    if (pvalid) {
#pragma unroll
        for (int k = 0; k < ZC1; ++k)
            if (z0 + k < D) {
                u[(((size_t)n * D + z0 + k) * H) * W + p] = acc[k];
                s0 += acc[k]; s1 = fmaf(acc[k], acc[k], s1);
            }
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
    if (lane == 0) { atomicAdd(&s_sum[0], (double)s0); atomicAdd(&s_sum[1], (double)s1); atomicAdd(&s_sum[2], (double)s2); atomicAdd(&s_sum[3], (double)s3); }
    __syncthreads();
    for (int i = tid; i < 2 * Cout; i += 256) {
        const int isq = i >= Cout, c = isq ? i - Cout : i;
        const float wt = pw_w[c];
        atomicAdd(&t_stats[(size_t)isq * N * Cout + (size_t)n * Cout + c], isq ? (double)wt * wt * s_sum[1] : (double)wt * s_sum[0]);
        if (sc_w != nullptr && r_stats != nullptr) {
            const float wq = sc_w[c];
            atomicAdd(&r_stats[(size_t)isq * N * Cout + (size_t)n * Cout + c], isq ? (double)wq * wq * s_sum[3] : (double)wq * s_sum[2]);
        }
    }
}

// bf16, W % 8 == 0, contiguous single-channel input: thread = 8 x-consecutive voxels x ZC1V planes.  One 16-byte load per
// input row (+ the two x-halo scalars), so ~1.7 loads per voxel instead of 11; lanes run along the flattened (y, x) plane.
constexpr int ZC1V = 4;
__global__ void __launch_bounds__(256) dw_c1_vec_kernel(const h16 *__restrict__ x, NormDev xn, int N, int D, int H, int W,
                                                        const float *__restrict__ dw_w, const float *__restrict__ pw_w,
                                                        const float *__restrict__ sc_w, int Cout, float *__restrict__ u,
                                                        double *__restrict__ t_stats, double *__restrict__ r_stats) {
    __shared__ float s_w[27];
    __shared__ double s_sum[4];       // double: the warps' partial sums meet in a run-dependent order
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 27) s_w[tid] = dw_w[tid];
    if (tid < 4) s_sum[tid] = 0.0;
    __syncthreads();
    const int n = blockIdx.z, z0 = blockIdx.y * ZC1V;
    const int W8 = W >> 3;
    const int q = blockIdx.x * 256 + tid;                 // 8-voxel group within the plane
    const bool qvalid = q < H * W8;
    const int y = qvalid ? q / W8 : 0, x0 = qvalid ? (q - (q / W8) * W8) * 8 : 0;
    float sc, sh;
    norm_scale_shift(xn, N, 1, n, 0, sc, sh);
    const float slope = xn.slope;
    const bool ident = xn.stats == nullptr && slope == 1.f;
    float wr[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) wr[k] = s_w[k];
    float acc[ZC1V][8];
#pragma unroll
    for (int k = 0; k < ZC1V; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
    float s2 = 0.f, s3 = 0.f;
    const bool okl = x0 > 0, okr = x0 + 8 < W;
#pragma unroll
    for (int zi = 0; zi < ZC1V + 2; ++zi) {
        const int gz = z0 + zi - 1;
        const bool okz = qvalid && gz >= 0 && gz < D;
        const h16 *plane = x + (((size_t)n * D + (okz ? gz : 0)) * H) * W;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int gy = y + dy - 1;
            const bool ok = okz && gy >= 0 && gy < H;
            float v[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) v[i] = 0.f;
            if (ok) {
                const h16 *row = plane + (size_t)gy * W + x0;
                const uint4 r4 = *reinterpret_cast<const uint4 *>(row);
                v[1] = h16_lo(r4.x); v[2] = h16_hi(r4.x);
                v[3] = h16_lo(r4.y); v[4] = h16_hi(r4.y);
                v[5] = h16_lo(r4.z); v[6] = h16_hi(r4.z);
                v[7] = h16_lo(r4.w); v[8] = h16_hi(r4.w);
                if (okl) v[0] = __half2float(row[-1]);
                if (okr) v[9] = __half2float(row[8]);
                if (!ident) {                  // the network input carries no norm / activation: v * 1 + 0 through lrelu(., 1) is v
#pragma unroll
                    for (int i = 0; i < 10; ++i) v[i] = lrelu(fmaf(v[i], sc, sh), slope);
                    if (!okl) v[0] = 0.f;
                    if (!okr) v[9] = 0.f;
                }
            }
#pragma unroll
            for (int dz = 0; dz < 3; ++dz) {
                const int k = zi - dz;                    // output plane z0 + k reads input plane gz with tap dz
                if (k >= 0 && k < ZC1V) {
                    const float w0 = wr[dz * 9 + dy * 3], w1 = wr[dz * 9 + dy * 3 + 1], w2 = wr[dz * 9 + dy * 3 + 2];
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[k][i] = fmaf(w2, v[i + 2], fmaf(w1, v[i + 1], fmaf(w0, v[i], acc[k][i])));
                }
            }
            if (dy == 1 && zi >= 1 && zi <= ZC1V && ok) {
#pragma unroll
                for (int i = 1; i <= 8; ++i) { s2 += v[i]; s3 = fmaf(v[i], v[i], s3); }
            }
        }
    }
    float s0 = 0.f, s1 = 0.f;
    if (qvalid) {
#pragma unroll
        for (int k = 0; k < ZC1V; ++k)
            if (z0 + k < D) {
                float4 *dst = reinterpret_cast<float4 *>(u + (((size_t)n * D + z0 + k) * H + y) * W + x0);
                dst[0] = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
                dst[1] = make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
#pragma unroll
                for (int i = 0; i < 8; ++i) { s0 += acc[k][i]; s1 = fmaf(acc[k][i], acc[k][i], s1); }
            }
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
    if (lane == 0) { atomicAdd(&s_sum[0], (double)s0); atomicAdd(&s_sum[1], (double)s1); atomicAdd(&s_sum[2], (double)s2); atomicAdd(&s_sum[3], (double)s3); }
    __syncthreads();
    for (int i = tid; i < 2 * Cout; i += 256) {
        const int isq = i >= Cout, c = isq ? i - Cout : i;
        const float wt = pw_w[c];
        atomicAdd(&t_stats[(size_t)isq * N * Cout + (size_t)n * Cout + c], isq ? (double)wt * wt * s_sum[1] : (double)wt * s_sum[0]);
        if (sc_w != nullptr && r_stats != nullptr) {
            const float wq = sc_w[c];
            atomicAdd(&r_stats[(size_t)isq * N * Cout + (size_t)n * Cout + c], isq ? (double)wq * wq * s_sum[3] : (double)wq * s_sum[2]);
        }
    }
}

// Dense 3x3x3 conv with a single input channel (the first conv of the dense / grouped variants, unet3d.py:49 with
// in_channels = 1): 27 taps x COUT filters per voxel on CUDA cores, thread = one voxel, all COUT outputs; optionally the
// block's 1x1x1 shortcut (r[c] = sc[c] * x) from the same staged tile.  Statistics by the transposing warp reduction.
template <typename T, int COUT>
__global__ void __launch_bounds__(NT) conv3_c1_kernel(
    const T *__restrict__ x, int ldx, NormDev xn, int N, int D, int H, int W, const float *__restrict__ wgt, const float *__restrict__ sc_w,
    T *__restrict__ t, int ldt, double *__restrict__ t_stats, T *__restrict__ r, int ldr, double *__restrict__ r_stats) {
    __shared__ float s_in[HZ][HY][HX + 1];
    __shared__ __align__(16) float s_w[27 * COUT + COUT];        // [tap][co], then the shortcut weights
    __shared__ double s_stat[4 * COUT];      // double: combined in run-dependent order
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 27 * COUT; i += NT) { const int tap = i / COUT, co = i % COUT; s_w[i] = wgt[(size_t)co * 27 + tap]; }
    for (int i = tid; i < COUT; i += NT) s_w[27 * COUT + i] = sc_w != nullptr ? sc_w[i] : 0.f;
    for (int i = tid; i < 4 * COUT; i += NT) s_stat[i] = 0.0;
    const int tilesX = (W + TX - 1) / TX, tilesY = (H + TY - 1) / TY, tilesZ = (D + TZ - 1) / TZ;
    const int tiles_per_sample = tilesX * tilesY * tilesZ;
    const int total_tiles = tiles_per_sample * N;
    const int per = (total_tiles + gridDim.x - 1) / gridDim.x;
    const int tile_begin = blockIdx.x * per, tile_end = min(total_tiles, tile_begin + per);
    const int lx = tid & 7, ly = (tid >> 3) & 7, lz = tid >> 6;
    int cur_n = -1;
    float sc = 1.f, sh = 0.f;
    auto flush = [&](int n) {
        if (n < 0) return;
        for (int i = tid; i < 2 * COUT; i += NT) {
            const int isq = i >= COUT, c = isq ? i - COUT : i;
            atomicAdd(&t_stats[(size_t)isq * N * COUT + (size_t)n * COUT + c], s_stat[i]);
            if (sc_w != nullptr) atomicAdd(&r_stats[(size_t)isq * N * COUT + (size_t)n * COUT + c], s_stat[2 * COUT + i]);
        }
        __syncthreads();
        for (int i = tid; i < 4 * COUT; i += NT) s_stat[i] = 0.0;
    };
    for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int n = tile / tiles_per_sample;
        int b = tile - n * tiles_per_sample;
        const int x0 = (b % tilesX) * TX; b /= tilesX;
        const int y0 = (b % tilesY) * TY; b /= tilesY;
        const int z0 = b * TZ;
        __syncthreads();
        if (n != cur_n) {
            flush(cur_n);
            cur_n = n;
            norm_scale_shift(xn, N, 1, n, 0, sc, sh);
        }
        for (int item = tid; item < HZ * HY * HX; item += NT) {
            int hv = item;
            const int hx = hv % HX; hv /= HX;
            const int hy = hv % HY;
            const int hz = hv / HY;
            const int gz = z0 + hz - 1, gy = y0 + hy - 1, gx = x0 + hx - 1;
            float v = 0.f;
            if (gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W)
                v = lrelu(ld1(x + ((((size_t)n * D + gz) * H + gy) * W + gx) * (size_t)ldx) * sc + sh, xn.slope);
            s_in[hz][hy][hx] = v;
        }
        __syncthreads();
        float acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
#pragma unroll
        for (int tap = 0; tap < 27; ++tap) {
            const float a = s_in[lz + tap / 9][ly + (tap / 3) % 3][lx + tap % 3];
#pragma unroll
            for (int c4 = 0; c4 < COUT; c4 += 4) {
                const float4 wv = *reinterpret_cast<const float4 *>(s_w + tap * COUT + c4);
                acc[c4] = fmaf(a, wv.x, acc[c4]); acc[c4 + 1] = fmaf(a, wv.y, acc[c4 + 1]);
                acc[c4 + 2] = fmaf(a, wv.z, acc[c4 + 2]); acc[c4 + 3] = fmaf(a, wv.w, acc[c4 + 3]);
            }
        }
        const float xc = s_in[lz + 1][ly + 1][lx + 1];
        const int gz = z0 + lz, gy = y0 + ly, gx = x0 + lx;
        const bool valid = gz < D && gy < H && gx < W;
        const size_t vox = (((size_t)n * D + gz) * H + gy) * W + gx;
        constexpr int V = VecW<T>::V;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            if (a == 1 && sc_w == nullptr) break;
            T *op = (a == 0 ? t + vox * (size_t)ldt : r + vox * (size_t)ldr);
            double *stat = s_stat + a * 2 * COUT;
#pragma unroll
            for (int cb = 0; cb < COUT; cb += 16) {
                float sv[32];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float o = a == 0 ? acc[cb + j] : xc * s_w[27 * COUT + cb + j];
                    const float q = valid ? round_as(t, o) : 0.f;      // statistics of the stored values
                    sv[j] = q; sv[16 + j] = q * q;
                }
                if (valid) {
#pragma unroll
                    for (int j0 = 0; j0 < 16; j0 += V) {
                        float o[V];
#pragma unroll
                        for (int j = 0; j < V; ++j) o[j] = a == 0 ? acc[cb + j0 + j] : xc * s_w[27 * COUT + cb + j0 + j];
                        stv(op + cb + j0, o);
                    }
                }
                warp_transpose_sum<32>(sv, lane);
                const int idx = warp_transpose_owner<32>(lane);
                atomicAdd(&stat[(idx >= 16 ? COUT + idx - 16 : idx) + cb], (double)sv[0]);
            }
        }
    }
    __syncthreads();
    flush(cur_n);
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) { l3d_set_error("cudaFuncSetAttribute(%zu B smem): %s", bytes, cudaGetErrorString(e)); return 1; }
    }
    return 0;
}

static bool vec4_ok(const l3d_act *a) {
    const size_t es = a->dtype == L3D_F32 ? 4 : 2;
    return (a->C % 4 == 0) && (a->ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(a->ptr) % (4 * es)) == 0);
}

}  // namespace

// tensor-core path for bf16 storage (l3d_fwd_tc.cu); returns -1 when it does not apply
int l3d_dwpw_fwd_tc(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                    const float *dw_w, const float *pw_w, const float *sc_w,
                    const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats,
                    const l3d_act *u, void *stream);

int l3d_conv3_tc(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                 const float *w, int groups, const float *dw_w, const float *pw_w, const float *sc_w,
                 const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats, int stat_ld, int co0, int cout_total,
                 void *stream);
int l3d_dwpw_fwd_slab(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                      const float *dw_w, const float *pw_w, const float *sc_w,
                      const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats, void *stream);
int l3d_convt_fwd_tc(const l3d_act *x, int N, int d, int h, int w_, const float *w, const float *b,
                     const l3d_act *out, int OD, int OH, int OW, int oz, int oy, int ox, void *stream);

// ================================================================= C ABI ====================
extern "C" int l3d_dwpw_fwd(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                            const float *dw_w, const float *pw_w, const float *sc_w,
                            const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats,
                            const l3d_act *u, void *stream) {
    L3D_REQUIRE(!act_null(x) && !act_null(t) && pw_w && t_stats, "l3d_dwpw_fwd: null argument");
    L3D_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0, "l3d_dwpw_fwd: bad dims");
    const int Cin = x->C, Cout = t->C;
    L3D_REQUIRE(Cout % 8 == 0, "l3d_dwpw_fwd: Cout=%d must be a multiple of 8", Cout);
    L3D_REQUIRE(t->dtype == x->dtype, "l3d_dwpw_fwd: dtype mismatch");
    L3D_REQUIRE(vec4_ok(t), "l3d_dwpw_fwd: output view must be 4-channel aligned");
    const bool has_r = sc_w != nullptr;
    // Cin == 1: the shortcut output may be left out (r = {NULL}) -- its statistics are still produced (analytically), for
    // consumers that evaluate the rank-1 shortcut on the fly (l3d_merge_fwd_rank1)
    const bool r_stats_only = has_r && act_null(r) && Cin == 1 && dw_w != nullptr && r_stats != nullptr;
    if (has_r && !r_stats_only) {
        L3D_REQUIRE(!act_null(r) && r_stats && r->C == Cout && r->dtype == x->dtype && vec4_ok(r), "l3d_dwpw_fwd: bad shortcut output");
    }
    const bool has_u = !act_null(u);
    if (has_u) L3D_REQUIRE(u->C == Cin && u->dtype == x->dtype, "l3d_dwpw_fwd: bad u view");
    // inference, narrow layers: depthwise o pointwise composed into one implicit GEMM (27x the pointwise MACs on the
    // tensor pipe; measured 2.2x faster than the CUDA-core stencil + GEMM kernel at 16/32 channels, on par at 32 -> 32 and
    // slower once the 27 weight tiles (27*Cin*Cout*2 B) crowd the operand buffers out of shared memory: Cin*Cout <= 1024)
    const int igemm_max_env = L3D_ENV_INT("L3D_DWS_IGEMM_MAX", 0);
    int igemm_max = igemm_max_env > 0 ? igemm_max_env : 1024;
    // 64 -> 32 (+ shortcut) at >= 16^3: one launch with a shorter tile (the 27 weight tiles take 110 KB) beats two launches over
    // 16-channel output slices that each re-read and re-activate the input (measured 974 vs 1162 us for 325 windows at 24^3)
    if (igemm_max_env <= 0 && D >= 16 && H >= 16 && Cin * Cout <= 2048 && Cout <= 32) igemm_max = 2048;
    if (dw_w != nullptr && !has_u && Cin * Cout <= igemm_max) {
        const int rc = l3d_conv3_tc(x, xn, N, D, H, W, nullptr, 1, dw_w, pw_w, sc_w, t, t_stats, r, r_stats, Cout, 0, Cout, stream);
        if (rc >= 0) return rc;
    }
    // wider layers: several launches over output-channel slices, each with weights that fit next to the operand buffers
    // (the input is re-read per slice).  Measured faster than the stencil kernel at 24^3 (64 -> 32: 1.57 -> 1.24 ms for
    // 325 windows); at 12^3 / 6^3 the 16 x 8-voxel MMA tiles waste too many rows, so small volumes keep the stencil.
    {
        const int min_dim = L3D_ENV_INT("L3D_DWS_SLICE_MIN_DIM", 16);
        if (dw_w != nullptr && !has_u && Cin * Cout > igemm_max && Cin * 16 <= 2 * igemm_max && Cout % 16 == 0 && D >= min_dim && H >= min_dim &&
            x->dtype == L3D_F16) {
            for (int Cs : {32, 16}) {
                if (Cout % Cs != 0 || Cout <= Cs || Cin * Cs > igemm_max) continue;
                int rc = 0;
                for (int c0 = 0; c0 < Cout && rc == 0; c0 += Cs) {
                    l3d_act th = *t, rh;
                    th.ptr = (char *)t->ptr + (size_t)c0 * 2; th.C = Cs;
                    if (has_r) { rh = *r; rh.ptr = (char *)r->ptr + (size_t)c0 * 2; rh.C = Cs; }
                    rc = l3d_conv3_tc(x, xn, N, D, H, W, nullptr, 1, dw_w, pw_w + (size_t)c0 * Cin, has_r ? sc_w + (size_t)c0 * Cin : nullptr,
                                      &th, t_stats + c0, has_r ? &rh : nullptr, has_r ? r_stats + c0 : nullptr, Cout, 0, Cs, stream);
                    if (rc < 0 && c0 > 0) { l3d_set_error("l3d_dwpw_fwd: implicit GEMM accepted one channel slice but not the next"); return 3; }
                }
                if (rc >= 0) return rc;
            }
        }
    }
    // small volumes (the 12^3 / 6^3 levels of a 48^3 window): whole-plane slabs, no halo or tile padding (l3d_fwd_slab.cu)
    if (dw_w != nullptr && !has_u && !r_stats_only && (long long)H * W <= 256) {
        const int rc = l3d_dwpw_fwd_slab(x, xn, N, D, H, W, dw_w, pw_w, sc_w, t, t_stats, r, r_stats, stream);
        if (rc >= 0) return rc;
    }
    {
        const int rc = l3d_dwpw_fwd_tc(x, xn, N, D, H, W, dw_w, pw_w, sc_w, t, t_stats, r, r_stats, u, stream);
        if (rc >= 0) return rc;
    }
    L3D_REQUIRE(!r_stats_only || Cout == 16 || Cout == 32, "l3d_dwpw_fwd: statistics-only shortcut needs Cout of 16 or 32");
    if (Cin == 1 && dw_w != nullptr && (Cout == 16 || Cout == 32)) {
        const int64_t tiles1 = num_tiles(N, D, H, W);
        const unsigned grid1 = (unsigned)(tiles1 < 148 * 8 ? tiles1 : 148 * 8);
        const NormDev nd1 = norm_dev(xn);
        cudaStream_t st1_ = (cudaStream_t)stream;
#define LAUNCH_C1(T, CO)                                                                                              \
        dwpw_c1_kernel<T, CO><<<grid1, NT, 0, st1_>>>((const T *)x->ptr, x->ldc, nd1, N, D, H, W, dw_w, pw_w, sc_w,      \
                                                      (T *)t->ptr, t->ldc, t_stats, (has_r && !r_stats_only) ? (T *)r->ptr : nullptr, \
                                                      (has_r && !r_stats_only) ? r->ldc : 0, r_stats, has_u ? (T *)u->ptr : nullptr, has_u ? u->ldc : 0)
        L3D_DISPATCH_DTYPE(x->dtype, T, { if (Cout == 16) LAUNCH_C1(T, 16); else LAUNCH_C1(T, 32); });
#undef LAUNCH_C1
        l3d_count_launch();
        l3d_note_kernel("dwpw_c1_kernel");
        L3D_CUDA_OK("l3d_dwpw_fwd (Cin=1) launch");
        return 0;
    }
    const int CPT = (Cout % 32 == 0) ? 16 : (Cout % 16 == 0) ? 8 : 4;
    const size_t smem = dwpw_smem_bytes(Cin, Cout, CPT);
    L3D_REQUIRE(smem <= 227 * 1024, "l3d_dwpw_fwd: Cin=%d needs %zu B shared memory", Cin, smem);
    const int64_t tiles = num_tiles(N, D, H, W);
    L3D_REQUIRE(tiles < (1ll << 31), "l3d_dwpw_fwd: grid too large");
    const NormDev nd = norm_dev(xn);
    const int vok = vec4_ok(x) ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_DWPW(T, CPTV)                                                                                   \
    do {                                                                                                       \
        auto kern = dwpw_fwd_kernel<T, CPTV>;                                                                  \
        if (set_smem(kern, smem)) return 3;                                                                    \
        kern<<<(unsigned)tiles, NT, smem, st>>>((const T *)x->ptr, x->ldc, Cin, nd, N, D, H, W, dw_w, pw_w, sc_w, Cout, \
                                                (T *)t->ptr, t->ldc, t_stats, has_r ? (T *)r->ptr : nullptr,   \
                                                has_r ? r->ldc : 0, r_stats, has_u ? (T *)u->ptr : nullptr,    \
                                                has_u ? u->ldc : 0, vok);                                      \
    } while (0)
    L3D_DISPATCH_DTYPE(x->dtype, T, {
        if (CPT == 16) LAUNCH_DWPW(T, 16);
        else if (CPT == 8) LAUNCH_DWPW(T, 8);
        else LAUNCH_DWPW(T, 4);
    });
#undef LAUNCH_DWPW
    l3d_count_launch();
    l3d_note_kernel("dwpw_fwd_kernel");
    L3D_CUDA_OK("l3d_dwpw_fwd launch");
    return 0;
}

extern "C" int l3d_conv3_fwd(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                             const float *w, int groups, const l3d_act *t, double *t_stats,
                             const float *sc_w, const l3d_act *r, double *r_stats, void *stream) {
    L3D_REQUIRE(!act_null(x) && !act_null(t) && w && t_stats, "l3d_conv3_fwd: null argument");
    const int Cin = x->C, Cout = t->C;
    L3D_REQUIRE(groups >= 1 && Cin % groups == 0 && Cout % groups == 0, "l3d_conv3_fwd: bad groups");
    L3D_REQUIRE(Cout % 8 == 0 && vec4_ok(t), "l3d_conv3_fwd: Cout=%d must be a multiple of 8 and aligned", Cout);
    L3D_REQUIRE(t->dtype == x->dtype, "l3d_conv3_fwd: dtype mismatch");
    const bool has_sc = sc_w != nullptr;
    if (has_sc) L3D_REQUIRE(!act_null(r) && r_stats && r->C == Cout && r->dtype == x->dtype && vec4_ok(r), "l3d_conv3_fwd: bad shortcut output");
    const size_t w_bytes = (size_t)27 * Cin * Cout * 2;      // fp16 weight tiles of the whole layer
    if (w_bytes <= 60 * 1024) {
        const int rc = l3d_conv3_tc(x, xn, N, D, H, W, w, groups, nullptr, nullptr, sc_w, t, t_stats, r, r_stats, Cout, 0, Cout, stream);
        if (rc >= 0) return rc;
    }
    // wide layers: the 27 weight tiles of all output channels would crowd the operand buffers out of shared memory, so
    // the layer runs as several launches over output-channel slices (each re-reads the input; still tensor-core bound)
    if (x->dtype == L3D_F16 && Cin % 16 == 0 && Cout % 16 == 0) {
        for (int Cs : {32, 16}) {
            if (Cout % Cs != 0 || Cout <= Cs) continue;
            if (Cs > 16 && (size_t)27 * Cin * Cs * 2 > 60 * 1024) continue;     // leave room for tall operand tiles
            int rc = 0;
            for (int c0 = 0; c0 < Cout && rc == 0; c0 += Cs) {
                l3d_act ts = *t, rs;
                ts.ptr = (char *)t->ptr + (size_t)c0 * 2; ts.C = Cs;
                if (has_sc) { rs = *r; rs.ptr = (char *)r->ptr + (size_t)c0 * 2; rs.C = Cs; }
                rc = l3d_conv3_tc(x, xn, N, D, H, W, w, groups, nullptr, nullptr, has_sc ? sc_w + (size_t)c0 * Cin : nullptr, &ts, t_stats + c0,
                                  has_sc ? &rs : nullptr, has_sc ? r_stats + c0 : nullptr, Cout, c0, Cout, stream);
                if (rc < 0 && c0 > 0) { l3d_set_error("l3d_conv3_fwd: implicit GEMM accepted one channel slice but not the next"); return 3; }
            }
            if (rc >= 0) return rc;
        }
    }
    if (w_bytes > 60 * 1024) {
        const int rc = l3d_conv3_tc(x, xn, N, D, H, W, w, groups, nullptr, nullptr, sc_w, t, t_stats, r, r_stats, Cout, 0, Cout, stream);
        if (rc >= 0) return rc;
    }
    if (Cin == 1 && groups == 1 && (Cout == 16 || Cout == 32)) {
        const int64_t tiles1 = num_tiles(N, D, H, W);
        const unsigned grid1 = (unsigned)(tiles1 < 148 * 4 ? tiles1 : 148 * 4);
        const NormDev nd1 = norm_dev(xn);
        cudaStream_t st1_ = (cudaStream_t)stream;
#define LAUNCH_C3C1(T, CO)                                                                                               \
        conv3_c1_kernel<T, CO><<<grid1, NT, 0, st1_>>>((const T *)x->ptr, x->ldc, nd1, N, D, H, W, w, sc_w, (T *)t->ptr, t->ldc, t_stats, \
                                                       has_sc ? (T *)r->ptr : nullptr, has_sc ? r->ldc : 0, r_stats)
        L3D_DISPATCH_DTYPE(x->dtype, T, { if (Cout == 16) LAUNCH_C3C1(T, 16); else LAUNCH_C3C1(T, 32); });
#undef LAUNCH_C3C1
        l3d_count_launch();
        l3d_note_kernel("conv3_c1_kernel");
        L3D_CUDA_OK("l3d_conv3_fwd (Cin=1) launch");
        return 0;
    }
    if (has_sc) {     // generic path: the shortcut is a separate pointwise launch
        const int rc = l3d_dwpw_fwd(x, xn, N, D, H, W, nullptr, sc_w, nullptr, r, r_stats, nullptr, nullptr, nullptr, stream);
        if (rc) return rc;
    }
    l3d_note_kernel("conv3_fwd_kernel");
    const int CC = (Cout % 32 == 0) ? 32 : (Cout % 16 == 0) ? 16 : 8;
    const size_t smem = sizeof(float) * ((size_t)C3_CK * HZ * HY * HX + 27 * C3_CK * CC + 2 * (size_t)Cin + 2 + 4 * (size_t)Cout);
    const int64_t tiles = num_tiles(N, D, H, W);
    L3D_REQUIRE(tiles < (1ll << 31), "l3d_conv3_fwd: grid too large");
    const NormDev nd = norm_dev(xn);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_C3(T, CCV)                                                                                       \
    do {                                                                                                        \
        auto kern = conv3_fwd_kernel<T, CCV>;                                                                   \
        if (set_smem(kern, smem)) return 3;                                                                     \
        kern<<<(unsigned)tiles, NT, smem, st>>>((const T *)x->ptr, x->ldc, Cin, nd, N, D, H, W, w, groups, Cout, \
                                                (T *)t->ptr, t->ldc, t_stats);                                  \
    } while (0)
    L3D_DISPATCH_DTYPE(x->dtype, T, {
        if (CC == 32) LAUNCH_C3(T, 32);
        else if (CC == 16) LAUNCH_C3(T, 16);
        else LAUNCH_C3(T, 8);
    });
#undef LAUNCH_C3
    l3d_count_launch();
    L3D_CUDA_OK("l3d_conv3_fwd launch");
    return 0;
}

extern "C" int l3d_merge_fwd(const l3d_act *t2, const l3d_norm *n2, const l3d_act *r, const l3d_norm *nr,
                             int N, int D, int H, int W, float slope,
                             const l3d_act *out, const l3d_act *pooled,
                             const float *head_w, const float *head_b, int OC, float *prob, float *logits,
                             void *stream) {
    L3D_REQUIRE(!act_null(t2) && !act_null(r), "l3d_merge_fwd: null argument");
    const int C = t2->C;
    L3D_REQUIRE(r->C == C && r->dtype == t2->dtype, "l3d_merge_fwd: shortcut view mismatch");
    L3D_REQUIRE(C % 4 == 0 && vec4_ok(t2) && vec4_ok(r), "l3d_merge_fwd: views must be 4-channel aligned");
    const bool has_out = !act_null(out), has_pool = !act_null(pooled);
    if (has_out) L3D_REQUIRE(out->C == C && out->dtype == t2->dtype && vec4_ok(out), "l3d_merge_fwd: bad out view");
    if (has_pool) L3D_REQUIRE(pooled->C == C && pooled->dtype == t2->dtype && vec4_ok(pooled), "l3d_merge_fwd: bad pooled view");
    const NormDev d2 = norm_dev(n2), dr = norm_dev(nr);
    cudaStream_t st = (cudaStream_t)stream;
    if (head_w != nullptr) {
        L3D_REQUIRE(!has_pool, "l3d_merge_fwd: head and pool cannot be combined");
        L3D_REQUIRE(C <= 64 && OC >= 1 && prob && head_b, "l3d_merge_fwd: head needs C <= 64 (got %d)", C);
        {
            const int V = t2->dtype == L3D_F32 ? 4 : 8;
            auto vec_ok = [V](const l3d_act *a) {
                return a->C % V == 0 && a->ldc % V == 0 && (reinterpret_cast<uintptr_t>(a->ptr) % 16) == 0;
            };
            L3D_REQUIRE(vec_ok(t2) && vec_ok(r) && (!has_out || vec_ok(out)), "l3d_merge_fwd: views must be 16-byte aligned with C a multiple of %d", V);
        }
        const size_t nvox = (size_t)D * H * W;
        // a few CTAs per sample, each striding over many voxels: the per-CTA prologue (norm tables in double precision,
        // head weights) would otherwise cost more than the 256 voxels a one-shot CTA handles
        size_t gx_ = (nvox + 255) / 256;
        const size_t cap_ = (148 * 16 + (size_t)N - 1) / (size_t)N;
        if (gx_ > cap_) gx_ = cap_;
        const unsigned gx = (unsigned)gx_;
        dim3 grid(gx, (unsigned)N);
        if (C == 16 && t2->dtype == L3D_F16) {
            size_t gh = (2 * nvox + 511) / 512;
            const size_t caph = (148 * 16 + (size_t)N - 1) / (size_t)N;
            if (gh > caph) gh = caph;
            dim3 gridh((unsigned)gh, (unsigned)N);
            auto al32 = [](const l3d_act *a) { return a->ldc % 16 == 0 && reinterpret_cast<uintptr_t>(a->ptr) % 32 == 0; };
            if (al32(t2) && al32(r) && (!has_out || al32(out)) && L3D_ENV_INT("L3D_MERGE_256", 1) != 0) {
                size_t gv = (nvox + 511) / 512;
                if (gv > caph) gv = caph;
                dim3 gridv((unsigned)gv, (unsigned)N);
                merge_head16v_fwd_kernel<<<gridv, 256, 0, st>>>((const h16 *)t2->ptr, t2->ldc, d2, (const h16 *)r->ptr, r->ldc, dr, N, nvox, slope,
                                                               has_out ? (h16 *)out->ptr : nullptr, has_out ? out->ldc : 0, head_w, head_b, OC, prob, logits);
            } else
            // (four loads in flight per thread instead of two changed nothing: 537 vs 556 us -- the half-voxel kernel is not latency-bound)
            merge_head16_fwd_kernel<2><<<gridh, 256, 0, st>>>((const h16 *)t2->ptr, t2->ldc, d2, (const h16 *)r->ptr, r->ldc, dr, N, nvox, slope,
                                                          has_out ? (h16 *)out->ptr : nullptr, has_out ? out->ldc : 0, head_w, head_b, OC, prob, logits);
            l3d_count_launch();
            L3D_CUDA_OK("l3d_merge_fwd (head) launch");
            return 0;
        }
        L3D_DISPATCH_DTYPE(t2->dtype, T, {
            if (C <= 16)
                merge_head_fwd_kernel<T, 16><<<grid, 256, 0, st>>>((const T *)t2->ptr, t2->ldc, d2, (const T *)r->ptr, r->ldc, dr, N, C, nvox, slope,
                                                                   has_out ? (T *)out->ptr : nullptr, has_out ? out->ldc : 0, head_w, head_b, OC, prob, logits);
            else if (C <= 32)
                merge_head_fwd_kernel<T, 32><<<grid, 256, 0, st>>>((const T *)t2->ptr, t2->ldc, d2, (const T *)r->ptr, r->ldc, dr, N, C, nvox, slope,
                                                                   has_out ? (T *)out->ptr : nullptr, has_out ? out->ldc : 0, head_w, head_b, OC, prob, logits);
            else
                merge_head_fwd_kernel<T, 64><<<grid, 256, 0, st>>>((const T *)t2->ptr, t2->ldc, d2, (const T *)r->ptr, r->ldc, dr, N, C, nvox, slope,
                                                                   has_out ? (T *)out->ptr : nullptr, has_out ? out->ldc : 0, head_w, head_b, OC, prob, logits);
        });
    } else {
        L3D_REQUIRE(has_out || has_pool, "l3d_merge_fwd: nothing to write");
        const int V = t2->dtype == L3D_F32 ? 4 : 8;
        auto vec_ok = [V](const l3d_act *a) {
            return a->C % V == 0 && a->ldc % V == 0 && (reinterpret_cast<uintptr_t>(a->ptr) % 16) == 0;
        };
        L3D_REQUIRE(vec_ok(t2) && vec_ok(r) && (!has_out || vec_ok(out)) && (!has_pool || vec_ok(pooled)),
                    "l3d_merge_fwd: views must be 16-byte aligned with C a multiple of %d", V);
        const size_t total = (size_t)((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2) * (C / V);
        L3D_REQUIRE(total < (1ull << 31), "l3d_merge_fwd: sample too large for 32-bit cell indices");
        size_t blocks = (total + 255) / 256;
        const size_t cap = (148 * 32 + N - 1) / N;
        if (blocks > cap) blocks = cap;
        dim3 grid((unsigned)blocks, (unsigned)N);
        const int CQ = C / V;
        auto al32 = [](const l3d_act *a) { return a->ldc % 16 == 0 && reinterpret_cast<uintptr_t>(a->ptr) % 32 == 0; };
        const int CQ2 = C / 16;
        const bool col_ok = total < (1ull << 29);       // the column kernels count 2x the cells in 32 bits, plus the grid stride
        if (col_ok && t2->dtype == L3D_F16 && C % 16 == 0 && (CQ2 & (CQ2 - 1)) == 0 && CQ2 <= 16 && al32(t2) && al32(r) && (!has_out || al32(out)) &&
            (!has_pool || al32(pooled)) && L3D_ENV_INT("L3D_MERGE_256", 1) != 0 && L3D_ENV_INT("L3D_MERGE_CELL", 0) == 0) {
            // 32-byte channel vectors: one thread per x-column of a cell and 16 channels
            const size_t items = (size_t)((D + 1) / 2) * ((H + 1) / 2) * (2 * ((W + 1) / 2)) * CQ2;
            size_t bl = (items + 255) / 256;
            if (bl > cap) bl = cap;
            dim3 grid3((unsigned)bl, (unsigned)N);
            merge_col32_fwd_kernel<<<grid3, 256, sizeof(float) * 4 * C, st>>>((const h16 *)t2->ptr, t2->ldc, d2, (const h16 *)r->ptr, r->ldc, dr, N, C, D, H, W, slope,
                                                                              has_out ? (h16 *)out->ptr : nullptr, has_out ? out->ldc : 0,
                                                                              has_pool ? (h16 *)pooled->ptr : nullptr, has_pool ? pooled->ldc : 0);
        } else
        if (col_ok && (CQ & (CQ - 1)) == 0 && CQ <= 16 && L3D_ENV_INT("L3D_MERGE_CELL", 0) <= 0) {
            // thread = x-column of a cell: twice the threads of the cell mapping (L3D_MERGE_CELL=1: cell per thread)
            size_t blocks2 = (2 * total + 255) / 256;
            if (blocks2 > cap) blocks2 = cap;
            dim3 grid2((unsigned)blocks2, (unsigned)N);
            L3D_DISPATCH_DTYPE(t2->dtype, T, {
                merge_col_fwd_kernel<T, false><<<grid2, 256, sizeof(float) * 4 * C, st>>>((const T *)t2->ptr, t2->ldc, d2, (const T *)r->ptr, r->ldc, dr, nullptr, N, C, D, H, W, slope,
                                                                       has_out ? (T *)out->ptr : nullptr, has_out ? out->ldc : 0,
                                                                       has_pool ? (T *)pooled->ptr : nullptr, has_pool ? pooled->ldc : 0);
            });
        } else
        L3D_DISPATCH_DTYPE(t2->dtype, T, {
            merge_fwd_kernel<T, false><<<grid, 256, sizeof(float) * 4 * C, st>>>((const T *)t2->ptr, t2->ldc, d2, (const T *)r->ptr, r->ldc, dr, nullptr, N, C, D, H, W, slope,
                                                                   has_out ? (T *)out->ptr : nullptr, has_out ? out->ldc : 0,
                                                                   has_pool ? (T *)pooled->ptr : nullptr, has_pool ? pooled->ldc : 0);
        });
    }
    l3d_count_launch();
    L3D_CUDA_OK("l3d_merge_fwd launch");
    return 0;
}

int l3d_conv3_tc_ex(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                    const float *w, int groups, const float *dw_w, const float *pw_w, const float *sc_w,
                    const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats, int stat_ld, int co0, int cout_total,
                    const float *r1_w, void *stream);

extern "C" int l3d_dw_c1_fwd(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                             const float *dw_w, const float *pw_w, const float *sc_w, int Cout,
                             float *u, double *t_stats, double *r_stats, void *stream) {
    L3D_REQUIRE(!act_null(x) && dw_w && pw_w && u && t_stats, "l3d_dw_c1_fwd: null argument");
    L3D_REQUIRE(x->C == 1, "l3d_dw_c1_fwd: the input view must have one channel");
    L3D_REQUIRE(N > 0 && N <= 65535 && D > 0 && H > 0 && W > 0 && Cout > 0, "l3d_dw_c1_fwd: bad dims");
    L3D_REQUIRE(sc_w == nullptr || r_stats != nullptr, "l3d_dw_c1_fwd: shortcut weights without a statistics buffer");
    const NormDev nd = norm_dev(xn);
    if (x->dtype == L3D_F16 && W % 8 == 0 && x->ldc == 1 && reinterpret_cast<uintptr_t>(x->ptr) % 16 == 0 &&
        reinterpret_cast<uintptr_t>(u) % 16 == 0) {
        dim3 gridv((unsigned)(((size_t)H * (W / 8) + 255) / 256), (unsigned)((D + ZC1V - 1) / ZC1V), (unsigned)N);
        dw_c1_vec_kernel<<<gridv, 256, 0, (cudaStream_t)stream>>>((const h16 *)x->ptr, nd, N, D, H, W, dw_w, pw_w, sc_w, Cout, u, t_stats, r_stats);
        l3d_count_launch();
        l3d_note_kernel("dw_c1_vec_kernel");
        L3D_CUDA_OK("l3d_dw_c1_fwd launch");
        return 0;
    }
    dim3 grid((unsigned)(((size_t)H * W + 255) / 256), (unsigned)((D + ZC1 - 1) / ZC1), (unsigned)N);
    L3D_DISPATCH_DTYPE(x->dtype, T, {
        dw_c1_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T *)x->ptr, x->ldc, nd, N, D, H, W, dw_w, pw_w, sc_w, Cout, u,
                                                                 t_stats, r_stats);
    });
    l3d_count_launch();
    l3d_note_kernel("dw_c1_kernel");
    L3D_CUDA_OK("l3d_dw_c1_fwd launch");
    return 0;
}

int l3d_conv3_tc_ex2(const l3d_act *x, const l3d_act *x2, const l3d_norm *xn, int N, int D, int H, int W,
                     const float *w, int groups, const float *dw_w, const float *pw_w, const float *sc_w,
                     const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats, int stat_ld, int co0, int cout_total,
                     const float *r1_w, void *stream);

extern "C" int l3d_dwpw_fwd2(const l3d_act *x_lo, const l3d_act *x_hi, const l3d_norm *xn, int N, int D, int H, int W,
                             const float *dw_w, const float *pw_w, const float *sc_w,
                             const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats, void *stream) {
    L3D_REQUIRE(!act_null(x_lo) && !act_null(x_hi) && !act_null(t) && dw_w && pw_w && t_stats, "l3d_dwpw_fwd2: null argument");
    L3D_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0, "l3d_dwpw_fwd2: bad dims");
    L3D_REQUIRE(x_lo->dtype == L3D_F16 && x_hi->dtype == L3D_F16 && t->dtype == L3D_F16, "l3d_dwpw_fwd2: fp16 storage only");
    L3D_REQUIRE(x_lo->C == 16 && x_hi->C == 16 && x_lo->ldc == 16 && x_hi->ldc == 16, "l3d_dwpw_fwd2: both inputs must be dense 16-channel tensors");
    if (sc_w != nullptr) L3D_REQUIRE(!act_null(r) && r_stats && r->C == t->C && r->dtype == t->dtype, "l3d_dwpw_fwd2: bad shortcut output");
    const int rc = l3d_conv3_tc_ex2(x_lo, x_hi, xn, N, D, H, W, nullptr, 1, dw_w, pw_w, sc_w, t, t_stats, r, r_stats, t->C, 0, t->C, nullptr, stream);
    if (rc < 0) { l3d_set_error("l3d_dwpw_fwd2: the implicit-GEMM kernel does not take this shape / alignment"); return 3; }
    return rc;
}

extern "C" int l3d_dwpw_fwd_rank1(const float *u, const float *r1_w, int Cin, const l3d_norm *xn, int N, int D, int H, int W,
                                  const float *dw_w, const float *pw_w, const l3d_act *t, double *t_stats, void *stream) {
    L3D_REQUIRE(u && r1_w && dw_w && pw_w && !act_null(t) && t_stats, "l3d_dwpw_fwd_rank1: null argument");
    L3D_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0, "l3d_dwpw_fwd_rank1: bad dims");
    L3D_REQUIRE(Cin == 16 && W % 4 == 0 && t->dtype == L3D_F16 && t->C % 16 == 0 && t->C <= 64,
                "l3d_dwpw_fwd_rank1: needs Cin = 16, W %% 4 == 0, bf16 output with 16..64 channels (got Cin=%d W=%d Cout=%d)", Cin, W, t->C);
    l3d_act xv;
    xv.ptr = const_cast<float *>(u); xv.C = Cin; xv.ldc = Cin; xv.dtype = L3D_F16; xv.pad_ = 0;
    const int rc = l3d_conv3_tc_ex(&xv, xn, N, D, H, W, nullptr, 1, dw_w, pw_w, nullptr, t, t_stats, nullptr, nullptr, t->C, 0, t->C, r1_w, stream);
    if (rc < 0) { l3d_set_error("l3d_dwpw_fwd_rank1: the implicit-GEMM kernel does not take this shape / alignment"); return 3; }
    return rc;
}

extern "C" int l3d_merge_fwd_rank1(const l3d_act *t2, const l3d_norm *n2, const l3d_act *x1, const float *r1_w, const l3d_norm *nr,
                                   int N, int D, int H, int W, float slope, const l3d_act *out, const l3d_act *pooled, void *stream) {
    L3D_REQUIRE(!act_null(t2) && !act_null(x1) && r1_w, "l3d_merge_fwd_rank1: null argument");
    const int C = t2->C;
    L3D_REQUIRE(x1->C == 1 && x1->dtype == t2->dtype, "l3d_merge_fwd_rank1: x1 must be a single-channel view of the same dtype");
    const bool has_out = !act_null(out), has_pool = !act_null(pooled);
    L3D_REQUIRE(has_out || has_pool, "l3d_merge_fwd_rank1: nothing to write");
    const int V = t2->dtype == L3D_F32 ? 4 : 8;
    auto vec_ok = [V](const l3d_act *a) { return a->C % V == 0 && a->ldc % V == 0 && (reinterpret_cast<uintptr_t>(a->ptr) % 16) == 0; };
    L3D_REQUIRE(vec_ok(t2) && (!has_out || (vec_ok(out) && out->C == C && out->dtype == t2->dtype)) &&
                (!has_pool || (vec_ok(pooled) && pooled->C == C && pooled->dtype == t2->dtype)), "l3d_merge_fwd_rank1: bad views");
    const NormDev d2 = norm_dev(n2), dr = norm_dev(nr);
    const size_t total = (size_t)((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2) * (C / V);
    L3D_REQUIRE(total < (1ull << 31), "l3d_merge_fwd_rank1: sample too large for 32-bit cell indices");
    size_t blocks = (total + 255) / 256;
    const size_t cap = (148 * 32 + N - 1) / N;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)N);
    const int CQ = C / V;
    {
        auto al32 = [](const l3d_act *a) { return a->ldc % 16 == 0 && reinterpret_cast<uintptr_t>(a->ptr) % 32 == 0; };
        if (total < (1ull << 29) && C == 16 && t2->dtype == L3D_F16 && al32(t2) && (!has_out || al32(out)) && (!has_pool || al32(pooled)) &&
            L3D_ENV_INT("L3D_MERGE_256", 1) != 0 && L3D_ENV_INT("L3D_MERGE_CELL", 0) == 0) {
            const size_t cols = (size_t)((D + 1) / 2) * ((H + 1) / 2) * (2 * ((W + 1) / 2));
            size_t bl = (cols + 255) / 256;
            if (bl > cap) bl = cap;
            dim3 gridc((unsigned)bl, (unsigned)N);
            merge_col16_r1_fwd_kernel<<<gridc, 256, 0, (cudaStream_t)stream>>>(
                (const h16 *)t2->ptr, t2->ldc, d2, (const h16 *)x1->ptr, x1->ldc, dr, r1_w, N, D, H, W, slope,
                has_out ? (h16 *)out->ptr : nullptr, has_out ? out->ldc : 0, has_pool ? (h16 *)pooled->ptr : nullptr, has_pool ? pooled->ldc : 0);
            l3d_count_launch();
            L3D_CUDA_OK("l3d_merge_fwd_rank1 launch");
            return 0;
        }
    }
    // column-per-thread mapping: measured SLOWER for the rank-1 shortcut (325 windows of 48^3: 634 vs 568 us; the plain merge
    // gains 18 %, 730 -> 617 us), so it is opt-in here (L3D_MERGE_CELL=-1) and the default above
    if (total < (1ull << 29) && (CQ & (CQ - 1)) == 0 && CQ <= 16 && L3D_ENV_INT("L3D_MERGE_CELL", 0) == -1) {
        size_t blocks2 = (2 * total + 255) / 256;
        if (blocks2 > cap) blocks2 = cap;
        dim3 grid2((unsigned)blocks2, (unsigned)N);
        L3D_DISPATCH_DTYPE(t2->dtype, T, {
            merge_col_fwd_kernel<T, true><<<grid2, 256, sizeof(float) * 4 * C, (cudaStream_t)stream>>>(
                (const T *)t2->ptr, t2->ldc, d2, (const T *)x1->ptr, x1->ldc, dr, r1_w, N, C, D, H, W, slope,
                has_out ? (T *)out->ptr : nullptr, has_out ? out->ldc : 0, has_pool ? (T *)pooled->ptr : nullptr, has_pool ? pooled->ldc : 0);
        });
    } else
    L3D_DISPATCH_DTYPE(t2->dtype, T, {
        merge_fwd_kernel<T, true><<<grid, 256, sizeof(float) * 4 * C, (cudaStream_t)stream>>>(
            (const T *)t2->ptr, t2->ldc, d2, (const T *)x1->ptr, x1->ldc, dr, r1_w, N, C, D, H, W, slope,
            has_out ? (T *)out->ptr : nullptr, has_out ? out->ldc : 0, has_pool ? (T *)pooled->ptr : nullptr, has_pool ? pooled->ldc : 0);
    });
    l3d_count_launch();
    L3D_CUDA_OK("l3d_merge_fwd_rank1 launch");
    return 0;
}

extern "C" int l3d_convt_fwd(const l3d_act *x, int N, int d, int h, int w_, const float *w, const float *b,
                             const l3d_act *out, int OD, int OH, int OW, int oz, int oy, int ox, void *stream) {
    L3D_REQUIRE(!act_null(x) && !act_null(out) && w && b, "l3d_convt_fwd: null argument");
    const int Cin = x->C, Cout = out->C;
    L3D_REQUIRE(Cout % 4 == 0 && vec4_ok(out), "l3d_convt_fwd: Cout=%d must be a multiple of 4 and aligned", Cout);
    L3D_REQUIRE(out->dtype == x->dtype, "l3d_convt_fwd: dtype mismatch");
    {
        const int rc = l3d_convt_fwd_tc(x, N, d, h, w_, w, b, out, OD, OH, OW, oz, oy, ox, stream);
        if (rc >= 0) return rc;
    }
    const int CPT = (Cout % 16 == 0) ? 16 : (Cout % 8 == 0) ? 8 : 4;
    const size_t smem = sizeof(float) * ((size_t)CT_VOX * (Cin | 1) + 4 + (size_t)Cin * Cout);
    L3D_REQUIRE(smem <= 227 * 1024, "l3d_convt_fwd: Cin=%d Cout=%d needs %zu B shared memory", Cin, Cout, smem);
    const size_t nvox = (size_t)N * d * h * w_;
    const size_t blocks = (nvox + CT_VOX - 1) / CT_VOX;
    const int vok = vec4_ok(x) ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_CT(T, CPTV)                                                                                      \
    do {                                                                                                        \
        auto kern = convt_fwd_kernel<T, CPTV>;                                                                  \
        if (set_smem(kern, smem)) return 3;                                                                     \
        kern<<<(unsigned)blocks, NT, smem, st>>>((const T *)x->ptr, x->ldc, Cin, N, d, h, w_, w, b, Cout,       \
                                                 (T *)out->ptr, out->ldc, OD, OH, OW, oz, oy, ox, vok);         \
    } while (0)
    L3D_DISPATCH_DTYPE(x->dtype, T, {
        if (CPT == 16) LAUNCH_CT(T, 16);
        else if (CPT == 8) LAUNCH_CT(T, 8);
        else LAUNCH_CT(T, 4);
    });
#undef LAUNCH_CT
    l3d_count_launch();
    L3D_CUDA_OK("l3d_convt_fwd launch");
    return 0;
}
