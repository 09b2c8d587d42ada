// Shared device/host helpers for libl3d (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/l3d.h"

// 16-bit activation storage: IEEE fp16 (11-bit significand).  The MMA operands of the forward kernels are fp16 anyway, so
// storing the raw (pre-norm) tensors in the same format costs the same 2 B / element as bf16 and carries 8x less rounding
// error through the InstanceNorms (measured: pre-sigmoid logits 2.0-3.4e-2 -> 3-4e-3 rel-L2 against the fp32 oracle).
// Stores saturate to +-65504 instead of producing inf.
typedef __half h16;

// two stored values <-> floats: the low half-word is the lower channel
__device__ __forceinline__ float h16_lo(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); }
__device__ __forceinline__ float h16_hi(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }
__device__ __forceinline__ uint32_t pack_h16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// one 256-bit global store (STG.E.256, sm_100); p must be 32-byte aligned
__device__ __forceinline__ void st_global_256(void *p, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5, uint32_t w6, uint32_t w7) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w0), "r"(w1), "r"(w2), "r"(w3), "r"(w4), "r"(w5), "r"(w6), "r"(w7) : "memory");
}
// 16 fp32 values of a voxel row -> global, optionally accumulated onto what is there: two 256-bit accesses per direction when
// the row is 32-byte aligned (wide), four 128-bit ones otherwise
__device__ __forceinline__ void store16_f32(float *p, const float (&r)[16], bool accumulate, bool wide) {
    if (wide) {
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = r[j + k];
            if (accumulate) {
                uint32_t q[8];
                asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) : "l"(p + j) : "memory");
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] += __uint_as_float(q[k]);
            }
            st_global_256(p + j, __float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]),
                          __float_as_uint(o[4]), __float_as_uint(o[5]), __float_as_uint(o[6]), __float_as_uint(o[7]));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            float4 o = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
            if (accumulate) { const float4 q = *reinterpret_cast<const float4 *>(p + j); o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
            *reinterpret_cast<float4 *>(p + j) = o;
        }
    }
}
__device__ __forceinline__ h16 to_h16(float v) { return __ushort_as_half((unsigned short)(pack_h16x2(v, 0.f) & 0xffffu)); }

// ------------------------------------------------------------------ errors --
void l3d_set_error(const char *fmt, ...);
void l3d_count_launch(int n = 1);
// name of the kernel a dispatching entry point launched last on this thread (l3d_last_kernel(), for per-kernel timing)
void l3d_note_kernel(const char *name);
// tiled CUtensorMap encode through a runtime-resolved driver entry point (no libcuda link dependency)
int l3d_encode_tiled(void *tmap, int dtype, unsigned rank, void *base, const unsigned long long *dims,
                     const unsigned long long *strides, const unsigned *box, const unsigned *estr);

// Tuning / test knobs from the environment, read ONCE per call site (getenv walks the whole environment block; it used to
// run on every l3d_dwpw_fwd).  l3d_env_refresh() invalidates the cached values (tests flip knobs between calls).
int l3d_env_generation();
int l3d_env_read(const char *name, int dflt);
#define L3D_ENV_INT(name, dflt)                                                                    \
    ([]() -> int {                                                                                 \
        static int v_ = 0, gen_ = -1;                                                              \
        const int g_ = l3d_env_generation();                                                       \
        if (gen_ != g_) { v_ = l3d_env_read(name, dflt); gen_ = g_; }                              \
        return v_;                                                                                 \
    }())

#define L3D_REQUIRE(cond, ...)                  \
    do {                                        \
        if (!(cond)) {                          \
            l3d_set_error(__VA_ARGS__);         \
            return 1;                           \
        }                                       \
    } while (0)

#define L3D_CUDA_OK(what)                                                            \
    do {                                                                             \
        cudaError_t e_ = cudaGetLastError();                                         \
        if (e_ != cudaSuccess) {                                                     \
            l3d_set_error("%s: %s", what, cudaGetErrorString(e_));                   \
            return 2;                                                                \
        }                                                                            \
    } while (0)

// Dispatch a templated launch on the activation dtype.
#define L3D_DISPATCH_DTYPE(dt, T, ...)                        \
    do {                                                      \
        if ((dt) == L3D_F32) { typedef float T; __VA_ARGS__; } \
        else { typedef h16 T; __VA_ARGS__; }                  \
    } while (0)

static inline bool act_null(const l3d_act *a) { return a == nullptr || a->ptr == nullptr; }

// ------------------------------------------------------------ element I/O --
__device__ __forceinline__ float ld1(const float *p) { return *p; }
__device__ __forceinline__ float ld1(const h16 *p) { return __half2float(*p); }
__device__ __forceinline__ void st1(float *p, float v) { *p = v; }
__device__ __forceinline__ void st1(h16 *p, float v) { *p = to_h16(v); }

// 4 consecutive channels; p must be aligned to 4 elements.
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 ld4(const h16 *p) {
    const uint2 r = *reinterpret_cast<const uint2 *>(p);
    return make_float4(h16_lo(r.x), h16_hi(r.x), h16_lo(r.y), h16_hi(r.y));
}
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ void st4(h16 *p, float4 v) {
    *reinterpret_cast<uint2 *>(p) = make_uint2(pack_h16x2(v.x, v.y), pack_h16x2(v.z, v.w));
}
// value as it will be read back from storage (fp16 rounding), used so that statistics and
// arg-max decisions are taken on exactly the stored numbers
__device__ __forceinline__ float round_as(const float *, float v) { return v; }
__device__ __forceinline__ float round_as(const h16 *, float v) { return __half2float(to_h16(v)); }

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// ----------------------------------------------------- norm prologue setup --
// scale/shift so that activated = lrelu(x*scale + shift, slope) reproduces
// InstanceNorm(affine) -> LeakyReLU -> Dropout3d (the keep-scale m >= 0 commutes with lrelu).
struct NormDev {
    const double *stats;
    const float *gamma, *beta, *drop;
    float eps, slope;
    int count;
};
static inline NormDev norm_dev(const l3d_norm *n) {
    NormDev d;
    if (n == nullptr || n->stats == nullptr) {
        d.stats = nullptr; d.gamma = d.beta = d.drop = nullptr; d.eps = 0.f; d.slope = 1.f; d.count = 1;
    } else {
        d.stats = n->stats; d.gamma = n->gamma; d.beta = n->beta; d.drop = n->drop;
        d.eps = n->eps; d.slope = n->slope; d.count = n->count;
    }
    return d;
}
// mean / rstd of channel c of sample n from the {sum, sumsq} buffer ([2][N][C])
__device__ __forceinline__ void norm_mean_rstd(const NormDev &nd, int N, int C, int n, int c, float &mean, float &rstd) {
    const double s = nd.stats[(size_t)n * C + c];
    const double q = nd.stats[(size_t)N * C + (size_t)n * C + c];
    const double m = s / (double)nd.count;
    double var = q / (double)nd.count - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)nd.eps));
}
__device__ __forceinline__ void norm_scale_shift(const NormDev &nd, int N, int C, int n, int c, float &scale, float &shift) {
    if (nd.stats == nullptr) { scale = 1.f; shift = 0.f; return; }
    float mean, rstd;
    norm_mean_rstd(nd, N, C, n, c, mean, rstd);
    float g = nd.gamma[c] * rstd;
    float b = nd.beta[c] - mean * g;
    if (nd.drop != nullptr) { const float m = nd.drop[(size_t)n * C + c]; g *= m; b *= m; }
    scale = g; shift = b;
}

// g_t = a*gz + b*t + d : InstanceNorm(affine) backward with the two reductions red = {sum gz, sum gz*xhat}
__device__ __forceinline__ void in_bwd_coef(const NormDev &nd, const double *__restrict__ red, int N, int C, int n, int c,
                                            float &a, float &b, float &d) {
    if (nd.stats == nullptr) { a = 1.f; b = 0.f; d = 0.f; return; }
    float mean, rstd;
    norm_mean_rstd(nd, N, C, n, c, mean, rstd);
    const double inv = 1.0 / (double)nd.count;
    const float k1 = (float)(red[(size_t)n * C + c] * inv);
    const float k2 = (float)(red[(size_t)N * C + (size_t)n * C + c] * inv);
    const float gr = nd.gamma[c] * rstd;
    a = gr;
    b = -gr * rstd * k2;
    d = gr * (mean * rstd * k2 - k1);
}

// The same coefficients in double, for a double-precision evaluation of g_t (in_bwd_apply).  The three terms cancel: the loss
// gradient that reaches the last blocks is nearly constant over a window (|mean gz| up to ~1e5 x |gz - mean gz|), so one ulp of
// an fp32 `d` is a per-cent error of g_t -- and which way the coefficients round depends on the order of the double atomics
// behind `red`, i.e. on the run (measured, 8 x 24^3: repetitions of one step differed by 2.5e-2 on individual gradient
// tensors, 8e-3 median, with fp32 coefficients; tools/repro_grad_race.py).  PyTorch's CPU kernels, which the oracle runs,
// evaluate the same expression with double accumulators (at::acc_type<float, false>).
__device__ __forceinline__ void in_bwd_coef_d(const NormDev &nd, const double *__restrict__ red, int N, int C, int n, int c,
                                              double &a, double &b, double &d) {
    if (nd.stats == nullptr) { a = 1.0; b = 0.0; d = 0.0; return; }
    const double inv = 1.0 / (double)nd.count;
    const double mean = nd.stats[(size_t)n * C + c] * inv;
    double var = nd.stats[(size_t)N * C + (size_t)n * C + c] * inv - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)nd.eps);
    const double k1 = red[(size_t)n * C + c] * inv, k2 = red[(size_t)N * C + (size_t)n * C + c] * inv;
    const double gr = (double)nd.gamma[c] * rstd;
    a = gr;
    b = -gr * rstd * k2;
    d = gr * (mean * rstd * k2 - k1);
}
__device__ __forceinline__ float in_bwd_apply(double a, double b, double d, float gz, float t) {
    return (float)fma(a, (double)gz, fma(b, (double)t, d));
}

// ------------------------------------------------------- warp reductions --
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// Transposing butterfly: every lane holds NV values (NV power of two <= 32); on return
// vals[0] of lane l holds the warp-wide total of value index (l / (32/NV)) ... i.e. lane l owns
// index l >> log2(32/NV); lanes sharing an index hold identical totals.  NV-1 + log2(32/NV) shuffles.
template <int NV>
__device__ __forceinline__ void warp_transpose_sum(float (&vals)[NV], int lane) {
    int cnt = NV;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        if (cnt > 1) {
            const int h = cnt >> 1;
            const bool up = (lane & s) != 0;
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) {
                if (i < h) {
                    const float send = up ? vals[i] : vals[i + h];
                    const float keep = up ? vals[i + h] : vals[i];
                    vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
                }
            }
            cnt = h;
        } else {
            vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], s);
        }
    }
}
// The same over the 16 lanes that share this lane's parity (lane bit 0 is never crossed): NV = 16 values per lane, on return
// vals[0] of lane l is the total, over the lanes of l's parity, of value index warp_transpose_owner<16>(l).  15 shuffles.
__device__ __forceinline__ void warp_transpose_sum_parity16(float (&vals)[16], int lane) {
    int cnt = 16;
#pragma unroll
    for (int s = 16; s >= 2; s >>= 1) {
        const int h = cnt >> 1;
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < h) {
                const float send = up ? vals[i] : vals[i + h];
                const float keep = up ? vals[i + h] : vals[i];
                vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
            }
        }
        cnt = h;
    }
}
// index owned by `lane` after warp_transpose_sum<NV>: the exchange steps consumed the top log2(NV)
// lane bits, most significant first, each selecting the upper half.
template <int NV>
__device__ __forceinline__ int warp_transpose_owner(int lane) {
    int idx = 0, cnt = NV;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        if (cnt > 1) { cnt >>= 1; if (lane & s) idx += cnt; }
    }
    return idx;
}
