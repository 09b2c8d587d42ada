// Focal Tversky loss kernels (losses.py:40-52) + library-wide error/launch bookkeeping.
#include <stdarg.h>

#include "l3d_common.cuh"
#include <cuda.h>

// ------------------------------------------------------------ bookkeeping --
static thread_local char g_err[512] = "";
static int64_t g_launches = 0;

void l3d_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static int g_env_gen = 0;
int l3d_env_generation() { return __atomic_load_n(&g_env_gen, __ATOMIC_RELAXED); }
int l3d_env_read(const char *name, int dflt) {
    const char *e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}
extern "C" void l3d_env_refresh(void) { __atomic_fetch_add(&g_env_gen, 1, __ATOMIC_RELAXED); }

void l3d_count_launch(int n) { __atomic_fetch_add(&g_launches, (int64_t)n, __ATOMIC_RELAXED); }
extern "C" void l3d_add_launch_count(int64_t n) { __atomic_fetch_add(&g_launches, n, __ATOMIC_RELAXED); }
static thread_local const char *g_last_kernel = "";
void l3d_note_kernel(const char *name) { g_last_kernel = name; }
extern "C" const char *l3d_last_kernel(void) { return g_last_kernel; }


// cuTensorMapEncodeTiled is resolved through the runtime at first use, so libl3d.so has no link-time dependency on
// libcuda.so.1 (it must load -- symbols only -- on hosts without a driver).
int l3d_encode_tiled(void *tmap, int dtype, unsigned rank, void *base, const unsigned long long *dims,
                     const unsigned long long *strides, const unsigned *box, const unsigned *estr) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr) {
            l3d_set_error("cuTensorMapEncodeTiled is not available from the driver");
            return 3;
        }
        fn = (encode_fn)p;
    }
    const CUresult cr = fn((CUtensorMap *)tmap, (CUtensorMapDataType)dtype, rank, base, (const cuuint64_t *)dims,
                           (const cuuint64_t *)strides, (const cuuint32_t *)box, (const cuuint32_t *)estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { l3d_set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr); return 3; }
    return 0;
}

extern "C" const char *l3d_last_error(void) { return g_err; }
extern "C" int l3d_abi_version(void) { return L3D_ABI_VERSION; }
extern "C" int64_t l3d_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

namespace {

// sums[0] += sum p*t, sums[1] += sum p, sums[2] += sum t.   FP = sum p - TP, FN = sum t - TP.
// Grid-stride, float4 loads, fp32 per-thread partials over a bounded span, double across threads.
__global__ void __launch_bounds__(256) ftl_sums_kernel(const float *__restrict__ p, const float *__restrict__ t,
                                                       int64_t n, double *__restrict__ sums) {
    const int64_t n4 = n >> 2;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(t)) & 15) == 0;
    double d_pt = 0.0, d_p = 0.0, d_t = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        float f_pt = 0.f, f_p = 0.f, f_t = 0.f;
        int k = 0;
        for (int64_t i = gid; i < n4; i += stride) {
            const float4 a = reinterpret_cast<const float4 *>(p)[i];
            const float4 b = reinterpret_cast<const float4 *>(t)[i];
            f_pt += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
            f_p += (a.x + a.y) + (a.z + a.w);
            f_t += (b.x + b.y) + (b.z + b.w);
            if (++k == 64) { d_pt += f_pt; d_p += f_p; d_t += f_t; f_pt = f_p = f_t = 0.f; k = 0; }
        }
        d_pt += f_pt; d_p += f_p; d_t += f_t;
        for (int64_t i = (n4 << 2) + gid; i < n; i += stride) { d_pt += (double)p[i] * t[i]; d_p += p[i]; d_t += t[i]; }
    } else {
        for (int64_t i = gid; i < n; i += stride) { d_pt += (double)p[i] * t[i]; d_p += p[i]; d_t += t[i]; }
    }
    d_pt = warp_sum(d_pt); d_p = warp_sum(d_p); d_t = warp_sum(d_t);
    __shared__ double s[3][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s[0][wid] = d_pt; s[1][wid] = d_p; s[2][wid] = d_t; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int i = 0; i < 8; ++i) v += s[threadIdx.x][i];
        atomicAdd(&sums[threadIdx.x], v);
    }
}

__global__ void ftl_finish_kernel(const double *__restrict__ sums, float alpha, float beta, float gamma, float smooth,
                                  float *__restrict__ loss, float *__restrict__ coef) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double tp = sums[0], sp = sums[1], st = sums[2];
    const double fp = sp - tp, fn = st - tp;
    const double a = alpha, b = beta, g = gamma, s = smooth;
    const double num = tp + s;
    const double den = tp + a * fn + b * fp + s;
    const double ti = num / den;
    const double om = 1.0 - ti;
    *loss = (float)pow(om, g);
    // dL/dp_i = -g*om^(g-1) * (t_i*den - num*dden_i)/den^2,   dden_i = (1-a-b)*t_i + b
    double k = 0.0;
    if (om > 0.0) k = -g * pow(om, g - 1.0) / (den * den);
    else if (g == 1.0) k = -1.0 / (den * den);
    const double c0 = k * (-num * b);
    const double c1 = k * (den - num * (1.0 - a - b));
    coef[0] = (float)c0;
    coef[1] = (float)c1;
}

__global__ void __launch_bounds__(256) ftl_grad_kernel(const float *__restrict__ t, int64_t n, const float *__restrict__ coef,
                                                       const float *__restrict__ g_loss, float *__restrict__ grad) {
    const float g = g_loss ? g_loss[0] : 1.f;
    const float c0 = coef[0] * g, c1 = coef[1] * g;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) grad[i] = c0 + c1 * t[i];
}

}  // namespace

extern "C" int l3d_ftl_sums(const float *pred, const float *target, int64_t n, double *sums, void *stream) {
    L3D_REQUIRE(pred && target && sums && n >= 0, "l3d_ftl_sums: null argument");
    if (n == 0) return 0;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    ftl_sums_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pred, target, n, sums);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_ftl_sums launch");
    return 0;
}

extern "C" int l3d_ftl_finish(const double *sums, float alpha, float beta, float gamma, float smooth,
                              float *loss, float *coef, void *stream) {
    L3D_REQUIRE(sums && loss && coef, "l3d_ftl_finish: null argument");
    ftl_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, alpha, beta, gamma, smooth, loss, coef);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_ftl_finish launch");
    return 0;
}

extern "C" int l3d_ftl_grad(const float *target, int64_t n, const float *coef, const float *g_loss, float *grad,
                            void *stream) {
    L3D_REQUIRE(target && coef && grad && n >= 0, "l3d_ftl_grad: null argument");
    if (n == 0) return 0;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ftl_grad_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(target, n, coef, g_loss, grad);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_ftl_grad launch");
    return 0;
}
