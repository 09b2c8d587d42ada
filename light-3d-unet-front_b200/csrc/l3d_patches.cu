// Training-patch pipeline on the device: crop a batch of patches out of volumes that stay resident in
// HBM and augment them there.  The reference re-reads two whole NIfTI volumes from disk per patch and runs
// scipy.ndimage.rotate / zoom on the host (patch_dataset.py:114-134, :156-220).
//
// Every kernel is one thread per output voxel over the whole batch, with per-sample parameters in small device tables
// (a sample whose augmentation is switched off is copied).  The resampling kernels restate scipy's arithmetic for
// spline order 1 (image) / 0 (label), mode='constant', cval=0 operation for operation in float64 with separately rounded
// multiplies and adds (ni_interpolation.c: coordinate = offset + sum o_j * m_j, weights (1 - f, f), value = sum of
// v * w_a * w_b [* w_c] in the scan order of the neighbours, cast to float32), so patches are bit-identical to the
// reference's for the same augmentation parameters (oracle/augment_ref.py == scipy, tests/test_augment_ref.py).
#include "l3d_common.cuh"

namespace {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

// scipy 'constant' mode along one axis: outside [0, n-1] -> cval; order 1: floor / floor + 1 (mirrored at the far edge,
// weight 0 there) with weights (1 - f, f); order 0: floor(cc + 0.5)
struct Axis1 { bool inside; int i0, i1; double w0, w1; };
__device__ __forceinline__ Axis1 axis_order1(double cc, int n) {
    Axis1 a;
    a.inside = !(cc < 0.0 || cc > (double)(n - 1));
    const double c = a.inside ? cc : 0.0;
    const double fl = floor(c);
    const double f = c - fl;
    a.i0 = (int)fl;
    int i1 = a.i0 + 1;
    if (i1 >= n) { i1 = 2 * n - 2 - i1; if (i1 < 0) i1 = 0; }
    a.i1 = i1;
    a.w0 = 1.0 - f; a.w1 = f;
    return a;
}
__device__ __forceinline__ int axis_order0(double cc, int n, bool &inside) {
    inside = !(cc < 0.0 || cc > (double)(n - 1));
    const double c = inside ? cc : 0.0;
    int i = (int)floor(c + 0.5);
    if (i >= n) { i = 2 * n - 2 - i; if (i < 0) i = 0; }
    return i;
}

__global__ void __launch_bounds__(256) patch_extract_kernel(const float *const *__restrict__ imgs, const float *const *__restrict__ labs,
                                                            const int32_t *__restrict__ dims, const int32_t *__restrict__ start,
                                                            int pd, int ph, int pw, float *__restrict__ out_img, float *__restrict__ out_lab) {
    const int b = blockIdx.y;
    const int D = dims[b * 3], H = dims[b * 3 + 1], W = dims[b * 3 + 2];
    const int z0 = start[b * 3], y0 = start[b * 3 + 1], x0 = start[b * 3 + 2];
    const float *img = imgs[b], *lab = labs[b];
    const int per = pd * ph * pw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
        const int x = i % pw, y = (i / pw) % ph, z = i / (pw * ph);
        const int gz = z0 + z, gy = y0 + y, gx = x0 + x;
        float vi = 0.f, vl = 0.f;
        if (gz < D && gy < H && gx < W) {                          // the window is clipped at the far edge and zero-padded at the end
            const size_t gi = ((size_t)gz * H + gy) * W + gx;
            vi = img[gi]; vl = lab[gi];
        }
        out_img[(size_t)b * per + i] = vi;
        out_lab[(size_t)b * per + i] = vl;
    }
}

__global__ void __launch_bounds__(256) patch_flip_kernel(const float *__restrict__ in_img, const float *__restrict__ in_lab,
                                                         float *__restrict__ out_img, float *__restrict__ out_lab,
                                                         const int32_t *__restrict__ axis, int pd, int ph, int pw) {
    const int b = blockIdx.y, ax = axis[b];
    const int per = pd * ph * pw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
        int x = i % pw, y = (i / pw) % ph, z = i / (pw * ph);
        if (ax == 0) z = pd - 1 - z; else if (ax == 1) y = ph - 1 - y; else if (ax == 2) x = pw - 1 - x;
        const size_t s = (size_t)b * per + ((size_t)z * ph + y) * pw + x;
        out_img[(size_t)b * per + i] = in_img[s];
        out_lab[(size_t)b * per + i] = in_lab[s];
    }
}

// scipy.ndimage.rotate(reshape=False): a 2-D affine map in the plane of axes (a0 < a1), applied plane by plane
__global__ void __launch_bounds__(256) patch_rotate_kernel(const float *__restrict__ in_img, const float *__restrict__ in_lab,
                                                           float *__restrict__ out_img, float *__restrict__ out_lab,
                                                           const int32_t *__restrict__ axes, const double *__restrict__ coef,
                                                           int pd, int ph, int pw) {
    const int b = blockIdx.y;
    const int a0 = axes[b * 2], a1 = axes[b * 2 + 1];
    const int per = pd * ph * pw;
    const int dim[3] = {pd, ph, pw};
    const size_t str[3] = {(size_t)ph * pw, (size_t)pw, 1};
    const double m00 = coef[b * 6], m01 = coef[b * 6 + 1], m10 = coef[b * 6 + 2], m11 = coef[b * 6 + 3], f0 = coef[b * 6 + 4], f1 = coef[b * 6 + 5];
    const float *si = in_img + (size_t)b * per, *sl = in_lab + (size_t)b * per;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
        if (a0 < 0) { out_img[(size_t)b * per + i] = si[i]; out_lab[(size_t)b * per + i] = sl[i]; continue; }
        const int o[3] = {i / (pw * ph), (i / pw) % ph, i % pw};
        const int a2 = 3 - a0 - a1;
        const double cc0 = dadd(dadd(f0, dmul((double)o[a0], m00)), dmul((double)o[a1], m01));
        const double cc1 = dadd(dadd(f1, dmul((double)o[a0], m10)), dmul((double)o[a1], m11));
        const size_t base = (size_t)o[a2] * str[a2];
        // image: order 1, four neighbours in scan order (0,0) (0,1) (1,0) (1,1), each v * w_a0 * w_a1
        const Axis1 p = axis_order1(cc0, dim[a0]), q = axis_order1(cc1, dim[a1]);
        float vi = 0.f;
        if (p.inside && q.inside) {
            double t = dmul(dmul((double)si[base + p.i0 * str[a0] + q.i0 * str[a1]], p.w0), q.w0);
            t = dadd(t, dmul(dmul((double)si[base + p.i0 * str[a0] + q.i1 * str[a1]], p.w0), q.w1));
            t = dadd(t, dmul(dmul((double)si[base + p.i1 * str[a0] + q.i0 * str[a1]], p.w1), q.w0));
            t = dadd(t, dmul(dmul((double)si[base + p.i1 * str[a0] + q.i1 * str[a1]], p.w1), q.w1));
            vi = (float)t;
        }
        bool in0, in1;
        const int n0 = axis_order0(cc0, dim[a0], in0), n1 = axis_order0(cc1, dim[a1], in1);
        out_img[(size_t)b * per + i] = vi;
        out_lab[(size_t)b * per + i] = (in0 && in1) ? sl[base + n0 * str[a0] + n1 * str[a1]] : 0.f;
    }
}

// scipy.ndimage.zoom followed by the reference's centre-crop / end-pad back to the patch size (patch_dataset.py:186-208):
// geo[b] = {on, zoomed dims (3), crop starts (3)}, zf[b] = per-axis coordinate scale (n - 1) / (zoomed - 1)
__global__ void __launch_bounds__(256) patch_zoom_kernel(const float *__restrict__ in_img, const float *__restrict__ in_lab,
                                                         float *__restrict__ out_img, float *__restrict__ out_lab,
                                                         const int32_t *__restrict__ geo, const double *__restrict__ zf,
                                                         int pd, int ph, int pw) {
    const int b = blockIdx.y;
    const int per = pd * ph * pw;
    const int32_t *g = geo + b * 7;
    const float *si = in_img + (size_t)b * per, *sl = in_lab + (size_t)b * per;
    const size_t sz = (size_t)ph * pw, sy = (size_t)pw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
        if (!g[0]) { out_img[(size_t)b * per + i] = si[i]; out_lab[(size_t)b * per + i] = sl[i]; continue; }
        const int qz = i / (pw * ph) + g[4], qy = (i / pw) % ph + g[5], qx = i % pw + g[6];      // index in the zoomed array
        float vi = 0.f, vl = 0.f;
        if (qz < g[1] && qy < g[2] && qx < g[3]) {                                               // else: the zero pad at the end
            const double cz = dmul((double)qz, zf[b * 3]), cy = dmul((double)qy, zf[b * 3 + 1]), cx = dmul((double)qx, zf[b * 3 + 2]);
            const Axis1 A = axis_order1(cz, pd), B = axis_order1(cy, ph), C = axis_order1(cx, pw);
            if (A.inside && B.inside && C.inside) {
                const int zi[2] = {A.i0, A.i1}, yi[2] = {B.i0, B.i1}, xi[2] = {C.i0, C.i1};
                const double wz[2] = {A.w0, A.w1}, wy[2] = {B.w0, B.w1}, wx[2] = {C.w0, C.w1};
                double t = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int kz = k >> 2, ky = (k >> 1) & 1, kx = k & 1;
                    const double v = dmul(dmul(dmul((double)si[zi[kz] * sz + yi[ky] * sy + xi[kx]], wz[kz]), wy[ky]), wx[kx]);
                    t = k == 0 ? v : dadd(t, v);
                }
                vi = (float)t;
            }
            bool iz, iy, ix;
            const int nz = axis_order0(cz, pd, iz), ny = axis_order0(cy, ph, iy), nx = axis_order0(cx, pw, ix);
            if (iz && iy && ix) vl = sl[nz * sz + ny * sy + nx];
        }
        out_img[(size_t)b * per + i] = vi;
        out_lab[(size_t)b * per + i] = vl;
    }
}

// intensity shift (float32 add, clip to [0, 1]) then Gaussian noise (float64 add, clip, cast) -- patch_dataset.py:210-218
__global__ void __launch_bounds__(256) patch_intensity_kernel(float *__restrict__ img, const float *__restrict__ shift, const int32_t *__restrict__ shift_on,
                                                              const double *__restrict__ noise, const int32_t *__restrict__ noise_on, int per) {
    const int b = blockIdx.y;
    const bool s_on = shift_on[b] != 0, n_on = noise != nullptr && noise_on[b] != 0;
    if (!s_on && !n_on) return;
    const float sh = shift[b];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
        float v = img[(size_t)b * per + i];
        if (s_on) v = fminf(fmaxf(__fadd_rn(v, sh), 0.f), 1.f);
        if (n_on) v = (float)fmin(fmax(dadd((double)v, noise[(size_t)b * per + i]), 0.0), 1.0);
        img[(size_t)b * per + i] = v;
    }
}

static dim3 patch_grid(int per, int B) {
    int bx = (per + 255) / 256;
    if (bx > 148 * 4) bx = 148 * 4;
    return dim3((unsigned)bx, (unsigned)B);
}

}  // namespace

extern "C" int l3d_patch_extract(const void *const *img_ptrs, const void *const *lab_ptrs, const int32_t *dims, const int32_t *start,
                                 int B, int pd, int ph, int pw, float *out_img, float *out_lab, void *stream) {
    L3D_REQUIRE(img_ptrs && lab_ptrs && dims && start && out_img && out_lab && B > 0 && pd > 0 && ph > 0 && pw > 0, "l3d_patch_extract: bad argument");
    patch_extract_kernel<<<patch_grid(pd * ph * pw, B), 256, 0, (cudaStream_t)stream>>>((const float *const *)img_ptrs, (const float *const *)lab_ptrs, dims,
                                                                                     start, pd, ph, pw, out_img, out_lab);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_patch_extract launch");
    return 0;
}

extern "C" int l3d_patch_flip(const float *in_img, const float *in_lab, float *out_img, float *out_lab, const int32_t *axis,
                              int B, int pd, int ph, int pw, void *stream) {
    L3D_REQUIRE(in_img && in_lab && out_img && out_lab && axis && B > 0 && in_img != out_img && in_lab != out_lab, "l3d_patch_flip: bad argument");
    patch_flip_kernel<<<patch_grid(pd * ph * pw, B), 256, 0, (cudaStream_t)stream>>>(in_img, in_lab, out_img, out_lab, axis, pd, ph, pw);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_patch_flip launch");
    return 0;
}

extern "C" int l3d_patch_rotate(const float *in_img, const float *in_lab, float *out_img, float *out_lab, const int32_t *axes,
                                const double *coef, int B, int pd, int ph, int pw, void *stream) {
    L3D_REQUIRE(in_img && in_lab && out_img && out_lab && axes && coef && B > 0 && in_img != out_img && in_lab != out_lab, "l3d_patch_rotate: bad argument");
    patch_rotate_kernel<<<patch_grid(pd * ph * pw, B), 256, 0, (cudaStream_t)stream>>>(in_img, in_lab, out_img, out_lab, axes, coef, pd, ph, pw);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_patch_rotate launch");
    return 0;
}

extern "C" int l3d_patch_zoom(const float *in_img, const float *in_lab, float *out_img, float *out_lab, const int32_t *geo,
                              const double *zf, int B, int pd, int ph, int pw, void *stream) {
    L3D_REQUIRE(in_img && in_lab && out_img && out_lab && geo && zf && B > 0 && in_img != out_img && in_lab != out_lab, "l3d_patch_zoom: bad argument");
    patch_zoom_kernel<<<patch_grid(pd * ph * pw, B), 256, 0, (cudaStream_t)stream>>>(in_img, in_lab, out_img, out_lab, geo, zf, pd, ph, pw);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_patch_zoom launch");
    return 0;
}

extern "C" int l3d_patch_intensity(float *img, const float *shift, const int32_t *shift_on, const double *noise, const int32_t *noise_on,
                                   int B, int64_t per, void *stream) {
    L3D_REQUIRE(img && shift && shift_on && B > 0 && per > 0 && per < (1ll << 31) && (noise == nullptr || noise_on != nullptr), "l3d_patch_intensity: bad argument");
    patch_intensity_kernel<<<patch_grid((int)per, B), 256, 0, (cudaStream_t)stream>>>(img, shift, shift_on, noise, noise_on, (int)per);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_patch_intensity launch");
    return 0;
}
