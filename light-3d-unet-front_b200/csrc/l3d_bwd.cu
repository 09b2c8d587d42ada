// Backward kernels of the 3D U-Net hot path (sm_100a).
//
//  merge_bwd_kernel       d/d(pre-activation) of the residual merge (+ max-pool routing), with the InstanceNorm
//  merge_head_bwd_kernel  backward reductions {sum gz, sum gz*xhat} of norm2 / the shortcut norm accumulated in
//                         the same pass; the head variant also does sigmoid' and the 1x1x1 head dgrad / wgrad
//                         (autograd of unet3d.py:87-91, 109, 220-221)
//  pw_bwd_kernel          InstanceNorm backward applied on load (g_t = a*gz + b*t + d), then dgrad
//                         g_u = g_t . W and wgrad g_W += g_t^T . u of a 1x1x1 conv; the same kernel in gather
//                         mode is the ConvTranspose3d(k=2,s=2) backward (8 taps)      (unet3d.py:18, 70-72, 119)
//  dw_bwd_kernel          depthwise 3x3x3 dgrad (flipped stencil) + wgrad, then back through
//                         Dropout3d / LeakyReLU of the producer with its norm reductions  (unet3d.py:16-17, 80-85)
//  conv3 backward         dense / grouped 3x3x3: g_t materialisation, direct dgrad, direct wgrad
#include "l3d_common.cuh"

namespace {

constexpr int NT = 256;

struct NormCoef {
    float mean, rstd, gamma, beta, m;
};
__device__ __forceinline__ NormCoef norm_coef(const NormDev &nd, int N, int C, int n, int c) {
    NormCoef k;
    if (nd.stats == nullptr) { k.mean = 0.f; k.rstd = 1.f; k.gamma = 1.f; k.beta = 0.f; k.m = 1.f; return k; }
    norm_mean_rstd(nd, N, C, n, c, k.mean, k.rstd);
    k.gamma = nd.gamma[c];
    k.beta = nd.beta[c];
    k.m = nd.drop != nullptr ? nd.drop[(size_t)n * C + c] : 1.f;
    return k;
}

template <typename T>
__device__ __forceinline__ void ld4a(const T *p, float (&v)[4]) {
    const float4 f = ld4(p);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}

// ===========================================================================================
// Residual-merge backward.  One thread = one 2x2x2 cell x 4 channels (so that the max-pool arg-max tie-break,
// first maximum in z,y,x scan order like ATen, is a thread-local decision).  grid = (blocks, N).
template <typename T>
__global__ void __launch_bounds__(NT) merge_bwd_kernel(
    const float *__restrict__ g_out, int ldgo, const float *__restrict__ pooled_g, int ldpg,
    const T *__restrict__ out, int ldo, const T *__restrict__ pooled, int ldp,
    const T *__restrict__ t2, int ld2, NormDev n2, const T *__restrict__ r, int ldr, NormDev nr,
    int N, int C, int D, int H, int W, float slope, float *__restrict__ gz, int ldgz,
    double *__restrict__ red2, double *__restrict__ redr) {
    // The two reductions of the InstanceNorm backward (sum g, sum g * xhat) are accumulated in DOUBLE from the first addition on:
    // sum g * xhat cancels to a small fraction of its terms, the backward of the next norm amplifies what is left, and every
    // fp32 partial sum combined in atomic (= run-dependent) order showed up as run-to-run differences of up to 2.5e-2 in
    // individual gradient tensors (8 x 24^3, tools/repro_grad_race.py, tools/diag_bwd_determinism.py: `red` differed by 6e-5
    // after this kernel).  PyTorch's CPU kernels, which the oracle runs, accumulate these sums in double as well.
    extern __shared__ __align__(16) float sm[];
    float *s_mean2 = sm, *s_rstd2 = sm + C, *s_meanr = sm + 2 * C, *s_rstdr = sm + 3 * C;
    double *s_red = reinterpret_cast<double *>(sm + 4 * C);      // s_red[3][C]  (16 * C bytes in: 8-byte aligned)
    const int n = blockIdx.y, tid = threadIdx.x;
    const bool has_r = nr.stats != nullptr;
    for (int c = tid; c < C; c += NT) {
        float m, rs;
        norm_mean_rstd(n2, N, C, n, c, m, rs);
        s_mean2[c] = m; s_rstd2[c] = rs;
        if (has_r) { norm_mean_rstd(nr, N, C, n, c, m, rs); s_meanr[c] = m; s_rstdr[c] = rs; }
        else { s_meanr[c] = 0.f; s_rstdr[c] = 0.f; }
        s_red[c] = 0.0; s_red[C + c] = 0.0; s_red[2 * C + c] = 0.0;
    }
    __syncthreads();
    const int CD = (D + 1) / 2, CH = (H + 1) / 2, CW = (W + 1) / 2, CQ = C / 4;
    const int PD = D / 2, PH = H / 2, PW = W / 2;
    const size_t total = (size_t)CD * CH * CW * CQ;
    double acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = 0.0;
    int cur_q = -1;
    auto flush = [&]() {
        if (cur_q < 0) return;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&s_red[cur_q * 4 + j], acc[j]);
            atomicAdd(&s_red[C + cur_q * 4 + j], acc[4 + j]);
            if (has_r) atomicAdd(&s_red[2 * C + cur_q * 4 + j], acc[8 + j]);
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) acc[j] = 0.0;
    };
    // per-sample cell index: 32-bit arithmetic (host check) instead of four 64-bit divisions per item
    for (uint32_t idx = blockIdx.x * NT + tid; idx < (uint32_t)total; idx += gridDim.x * NT) {
        uint32_t rem = idx;
        const int q = (int)(rem % (uint32_t)CQ); rem /= (uint32_t)CQ;
        const int cx = (int)(rem % (uint32_t)CW); rem /= (uint32_t)CW;
        const int cy = (int)(rem % (uint32_t)CH);
        const int cz = (int)(rem / (uint32_t)CH);
        if (q != cur_q) { flush(); cur_q = q; }
        const int c = q * 4;
        const bool has_pool = pooled_g != nullptr && cz < PD && cy < PH && cx < PW;
        float pv[4] = {0.f, 0.f, 0.f, 0.f}, pg[4] = {0.f, 0.f, 0.f, 0.f};
        bool taken[4] = {false, false, false, false};
        if (has_pool) {
            const size_t pvox = (((size_t)n * PD + cz) * PH + cy) * PW + cx;
            ld4a(pooled + pvox * ldp + c, pv);
            ld4a(pooled_g + pvox * ldpg + c, pg);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int z = cz * 2 + (k >> 2), y = cy * 2 + ((k >> 1) & 1), xx = cx * 2 + (k & 1);
            if (z < D && y < H && xx < W) {
                const size_t vox = (((size_t)n * D + z) * H + y) * W + xx;
                float o[4], g[4] = {0.f, 0.f, 0.f, 0.f}, a[4], b[4] = {0.f, 0.f, 0.f, 0.f};
                ld4a(out + vox * ldo + c, o);
                if (g_out != nullptr) ld4a(g_out + vox * ldgo + c, g);
                ld4a(t2 + vox * ld2 + c, a);
                if (has_r) ld4a(r + vox * ldr + c, b);
                float gv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (has_pool && !taken[j] && o[j] == pv[j]) { g[j] += pg[j]; taken[j] = true; }
                    gv[j] = g[j] * (o[j] > 0.f ? 1.f : slope);
                }
                st4(gz + vox * ldgz + c, make_float4(gv[0], gv[1], gv[2], gv[3]));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double gr = (double)round_as(gz, gv[j]);
                    acc[j] += gr;
                    acc[4 + j] = fma(gr, (double)((a[j] - s_mean2[c + j]) * s_rstd2[c + j]), acc[4 + j]);
                    acc[8 + j] = fma(gr, (double)((b[j] - s_meanr[c + j]) * s_rstdr[c + j]), acc[8 + j]);
                }
            }
        }
    }
    flush();
    __syncthreads();
    for (int c = tid; c < C; c += NT) {
        atomicAdd(&red2[(size_t)n * C + c], s_red[c]);
        atomicAdd(&red2[(size_t)N * C + (size_t)n * C + c], s_red[C + c]);
        if (has_r) {
            atomicAdd(&redr[(size_t)n * C + c], s_red[c]);
            atomicAdd(&redr[(size_t)N * C + (size_t)n * C + c], s_red[2 * C + c]);
        }
    }
}

// Residual merge + head backward: one thread = one voxel, all C (<= CMAX) channels, OC <= OCM heads.
template <typename T, int CMAX, int OCM>
__global__ void __launch_bounds__(NT) merge_head_bwd_kernel(
    const float *__restrict__ g_out, int ldgo, const T *__restrict__ out, int ldo,
    const T *__restrict__ t2, int ld2, NormDev n2, const T *__restrict__ r, int ldr, NormDev nr,
    int N, int C, size_t nvox, float slope, const float *__restrict__ head_w, int OC,
    const float *__restrict__ g_prob, const float *__restrict__ prob,
    float *__restrict__ g_head_w, float *__restrict__ g_head_b,
    float *__restrict__ gz, int ldgz, double *__restrict__ red2, double *__restrict__ redr) {
    __shared__ float s_mean2[CMAX], s_rstd2[CMAX], s_meanr[CMAX], s_rstdr[CMAX], s_hw[OCM * CMAX];
    __shared__ float s_acc[(3 + OCM) * CMAX + OCM];
    __shared__ double s_accd[3 * CMAX];                 // the InstanceNorm-backward reductions, in double (see merge_bwd_kernel)
    const int n = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const bool has_r = nr.stats != nullptr;
    for (int c = tid; c < CMAX; c += NT) {
        float m = 0.f, rs = 0.f, mr = 0.f, rr = 0.f;
        if (c < C) {
            norm_mean_rstd(n2, N, C, n, c, m, rs);
            if (has_r) norm_mean_rstd(nr, N, C, n, c, mr, rr);
        }
        s_mean2[c] = m; s_rstd2[c] = rs; s_meanr[c] = mr; s_rstdr[c] = rr;
        for (int oc = 0; oc < OCM; ++oc) s_hw[oc * CMAX + c] = (c < C && oc < OC) ? head_w[(size_t)oc * C + c] : 0.f;
    }
    for (int i = tid; i < (3 + OCM) * CMAX + OCM; i += NT) s_acc[i] = 0.f;
    for (int i = tid; i < 3 * CMAX; i += NT) s_accd[i] = 0.0;
    __syncthreads();
    double a1[CMAX], a2[CMAX], a3[CMAX];
    float ahw[OCM][CMAX], ahb[OCM];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        a1[c] = a2[c] = a3[c] = 0.0;
#pragma unroll
        for (int oc = 0; oc < OCM; ++oc) ahw[oc][c] = 0.f;
    }
#pragma unroll
    for (int oc = 0; oc < OCM; ++oc) ahb[oc] = 0.f;
    for (size_t v = (size_t)blockIdx.x * NT + tid; v < nvox; v += (size_t)gridDim.x * NT) {
        const size_t vox = (size_t)n * nvox + v;
        float gl[OCM];
#pragma unroll
        for (int oc = 0; oc < OCM; ++oc) {
            gl[oc] = 0.f;
            if (oc < OC) {
                const size_t oi = ((size_t)n * OC + oc) * nvox + v;
                const float p = prob[oi];
                gl[oc] = g_prob[oi] * p * (1.f - p);
                ahb[oc] += gl[oc];
            }
        }
#pragma unroll
        for (int c4 = 0; c4 < CMAX; c4 += 4) {
            if (c4 < C) {
                float o[4], g[4] = {0.f, 0.f, 0.f, 0.f}, a[4], b[4] = {0.f, 0.f, 0.f, 0.f}, gv[4];
                ld4a(out + vox * ldo + c4, o);
                if (g_out != nullptr) ld4a(g_out + vox * ldgo + c4, g);
                ld4a(t2 + vox * ld2 + c4, a);
                if (has_r) ld4a(r + vox * ldr + c4, b);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int oc = 0; oc < OCM; ++oc) {
                        g[j] += gl[oc] * s_hw[oc * CMAX + c4 + j];
                        ahw[oc][c4 + j] += gl[oc] * o[j];
                    }
                    gv[j] = g[j] * (o[j] > 0.f ? 1.f : slope);
                }
                st4(gz + vox * ldgz + c4, make_float4(gv[0], gv[1], gv[2], gv[3]));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double gr = (double)round_as(gz, gv[j]);
                    a1[c4 + j] += gr;
                    a2[c4 + j] = fma(gr, (double)((a[j] - s_mean2[c4 + j]) * s_rstd2[c4 + j]), a2[c4 + j]);
                    a3[c4 + j] = fma(gr, (double)((b[j] - s_meanr[c4 + j]) * s_rstdr[c4 + j]), a3[c4 + j]);
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        const double v1 = warp_sum(a1[c]), v2 = warp_sum(a2[c]), v3 = warp_sum(a3[c]);
        if (lane == 0) { atomicAdd(&s_accd[c], v1); atomicAdd(&s_accd[CMAX + c], v2); atomicAdd(&s_accd[2 * CMAX + c], v3); }
#pragma unroll
        for (int oc = 0; oc < OCM; ++oc) {
            const float vh = warp_sum(ahw[oc][c]);
            if (lane == 0) atomicAdd(&s_acc[(3 + oc) * CMAX + c], vh);
        }
    }
#pragma unroll
    for (int oc = 0; oc < OCM; ++oc) {
        const float vb = warp_sum(ahb[oc]);
        if (lane == 0) atomicAdd(&s_acc[(3 + OCM) * CMAX + oc], vb);
    }
    __syncthreads();
    for (int c = tid; c < C; c += NT) {
        atomicAdd(&red2[(size_t)n * C + c], s_accd[c]);
        atomicAdd(&red2[(size_t)N * C + (size_t)n * C + c], s_accd[CMAX + c]);
        if (has_r) {
            atomicAdd(&redr[(size_t)n * C + c], s_accd[c]);
            atomicAdd(&redr[(size_t)N * C + (size_t)n * C + c], s_accd[2 * CMAX + c]);
        }
        for (int oc = 0; oc < OC; ++oc) atomicAdd(&g_head_w[(size_t)oc * C + c], s_acc[(3 + oc) * CMAX + c]);
    }
    if (tid < OC) atomicAdd(&g_head_b[tid], s_acc[(3 + OCM) * CMAX + tid]);
}

// ===========================================================================================
// Pointwise (1x1x1) backward, and ConvTranspose3d(k2,s2) backward in gather mode.
//   plain : g[v][c] = a_c*gz[v][c] + b_c*t[v][c] + d_c  (c < Cg = Cout),  u[v][k] (k < Cu = Cin), w[c][k]
//   convT : for tap in 0..7: g[v][c] = g_out[up-voxel(v,tap)][c] (c < Cg = Cout_T), u = x (Cu = Cin_T),
//           w_tap[c][k] = W[k][c][tap];  g_x[v][k] = sum_tap sum_c g*w_tap,  g_W[k][c][tap] += sum_v u*g
constexpr int PB_V = 128;     // voxels per tile
constexpr int PB_MAXKC = 4;   // dgrad register chunks per thread

struct PwBwdArgs {
    const void *gz; int ldg;          // plain: gz ; convT: g_out (up-sampled grid)
    const void *t; int ldt;           // plain only
    NormDev nt; const double *red;    // plain only
    const void *u; int ldu; NormDev un;
    int N; long long vox;             // voxels per sample of the (input-resolution) grid
    int Cg, Cu;
    const float *w; float *g_w; float *g_b;
    void *g_u; int ldgu; int accumulate;
    // convT geometry
    int d, h, w_, OD, OH, OW, oz, oy, ox;
    int tap_split;   // convT: blockIdx.y selects one tap; its weight-gradient slice is accumulated over all tiles of the CTA
};

template <typename T, int KC, int MAXT, bool CONVT>
__global__ void __launch_bounds__(NT) pw_bwd_kernel(PwBwdArgs A) {
    extern __shared__ __align__(16) float sm[];
    const int Cg = A.Cg, Cu = A.Cu;
    const int PG = Cg | 1, PU = Cu | 1;
    float *s_w = sm;                                    // Cg*Cu   (16B aligned)
    float *s_g = s_w + (((size_t)Cg * Cu + 3) & ~(size_t)3);   // PB_V*PG
    float *s_u = s_g + (size_t)PB_V * PG;               // PB_V*PU
    // Cg x3 doubles : a, b, d of the InstanceNorm backward (evaluated in double: in_bwd_apply), 8-byte aligned
    double *s_ca = reinterpret_cast<double *>(sm + (((s_u + (size_t)PB_V * PU) - sm + 1) & ~(size_t)1));
    double *s_cb = s_ca + Cg, *s_cd = s_cb + Cg;
    float *s_us = reinterpret_cast<float *>(s_cd + Cg), *s_uh = s_us + Cu;         // Cu x2 : prologue scale/shift of u
    float *s_gb = s_uh + Cu;                            // Cg : bias gradient (convT)
    const int tid = threadIdx.x;
    const float *gz = (const float *)A.gz;
    const T *tt = (const T *)A.t, *uu = (const T *)A.u;
    float *gu = (float *)A.g_u;
    constexpr int TK = KC >= 4 ? 4 : 1;
    const int nkt = Cu / TK, nct = Cg / 4, ntile = nct * nkt;
    const int G = ntile <= NT ? NT / ntile : 1;
    const long long tiles_per_sample = (A.vox + PB_V - 1) / PB_V;
    const long long total_tiles = tiles_per_sample * A.N;
    constexpr int NTAP = CONVT ? 8 : 1;
    const int tap0 = (CONVT && A.tap_split) ? (int)blockIdx.y : 0;
    const int tap1 = (CONVT && A.tap_split) ? tap0 + 1 : NTAP;

    float wacc[MAXT][4][TK];
#pragma unroll
    for (int i = 0; i < MAXT; ++i)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < TK; ++b) wacc[i][a][b] = 0.f;
    if (CONVT) for (int c = tid; c < Cg; c += NT) s_gb[c] = 0.f;
    if (!CONVT) {
        for (int i = tid; i < Cg * Cu; i += NT) s_w[i] = A.w[i];
    }
    int cur_n = -1;
    const int v = tid & (PB_V - 1), half = tid >> 7;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = (int)(tile / tiles_per_sample);
        const long long v0 = (tile % tiles_per_sample) * PB_V;
        __syncthreads();   // previous tile fully consumed
        if (n != cur_n) {
            cur_n = n;
            for (int c = tid; c < Cg; c += NT) {
                double a = 1.0, b = 0.0, d = 0.0;
                if (!CONVT) in_bwd_coef_d(A.nt, A.red, A.N, Cg, n, c, a, b, d);
                s_ca[c] = a; s_cb[c] = b; s_cd[c] = d;
            }
            for (int k = tid; k < Cu; k += NT) {
                float sc, sh;
                norm_scale_shift(A.un, A.N, Cu, n, k, sc, sh);
                s_us[k] = sc; s_uh[k] = sh;
            }
            __syncthreads();
        }
        // ---- stage u tile (activated); 4 independent loads in flight per thread
        {
            const int groups = (Cu + 3) / 4;
            const bool vec = (Cu % 4 == 0) && (A.ldu % 4 == 0);
            const int nitem = PB_V * groups;
            for (int item0 = tid; item0 < nitem; item0 += 4 * NT) {
                float val[4][4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int item = item0 + b * NT;
                    val[b][0] = val[b][1] = val[b][2] = val[b][3] = 0.f;
                    if (item < nitem) {
                        const int q = item % groups, lv = item / groups;
                        if (v0 + lv < A.vox) {
                            const T *p = uu + ((size_t)n * A.vox + v0 + lv) * (size_t)A.ldu + q * 4;
                            if (vec) ld4a(p, val[b]);
                            else for (int j = 0; j < 4; ++j) if (q * 4 + j < Cu) val[b][j] = ld1(p + j);
                        }
                    }
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int item = item0 + b * NT;
                    if (item < nitem) {
                        const int q = item % groups, lv = item / groups;
                        const int k = q * 4;
                        const bool inb = v0 + lv < A.vox;
                        for (int j = 0; j < 4; ++j)
                            if (k + j < Cu) s_u[(size_t)lv * PU + k + j] = inb ? lrelu(val[b][j] * s_us[k + j] + s_uh[k + j], A.un.slope) : 0.f;
                    }
                }
            }
        }
        float dacc[PB_MAXKC][KC];
#pragma unroll
        for (int i = 0; i < PB_MAXKC; ++i)
#pragma unroll
            for (int j = 0; j < KC; ++j) dacc[i][j] = 0.f;
#pragma unroll 1
        for (int tap = tap0; tap < tap1; ++tap) {
            if (CONVT) {
                __syncthreads();   // previous tap's s_g / s_w consumed
                for (int i = tid; i < Cg * Cu; i += NT) {
                    const int k = i % Cu, c = i / Cu;
                    s_w[i] = A.w[((size_t)k * Cg + c) * 8 + tap];
                }
            }
            // ---- stage g tile; 4 independent (gz, t) load pairs in flight per thread
            {
                const int groups = Cg / 4;
                const int nitem = PB_V * groups;
                for (int item0 = tid; item0 < nitem; item0 += 4 * NT) {
                    float g4[4][4], t4[4][4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int item = item0 + b * NT;
                        g4[b][0] = g4[b][1] = g4[b][2] = g4[b][3] = 0.f;
                        t4[b][0] = t4[b][1] = t4[b][2] = t4[b][3] = 0.f;
                        if (item < nitem) {
                            const int q = item % groups, lv = item / groups;
                            const int c = q * 4;
                            if (v0 + lv < A.vox) {
                                if (CONVT) {
                                    long long rem = v0 + lv;
                                    const int ix = (int)(rem % A.w_); rem /= A.w_;
                                    const int iy = (int)(rem % A.h);
                                    const int iz = (int)(rem / A.h);
                                    const int Z = A.oz + 2 * iz + (tap >> 2), Y = A.oy + 2 * iy + ((tap >> 1) & 1), X = A.ox + 2 * ix + (tap & 1);
                                    if (Z >= 0 && Z < A.OD && Y >= 0 && Y < A.OH && X >= 0 && X < A.OW)
                                        ld4a(gz + ((((size_t)n * A.OD + Z) * A.OH + Y) * A.OW + X) * (size_t)A.ldg + c, g4[b]);
                                } else {
                                    const size_t gv = (size_t)n * A.vox + v0 + lv;
                                    ld4a(gz + gv * (size_t)A.ldg + c, g4[b]);
                                    if (A.nt.stats != nullptr) ld4a(tt + gv * (size_t)A.ldt + c, t4[b]);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int item = item0 + b * NT;
                        if (item < nitem) {
                            const int q = item % groups, lv = item / groups;
                            const int c = q * 4;
                            const bool inb = v0 + lv < A.vox;
                            for (int j = 0; j < 4; ++j) {
                                float val = g4[b][j];
                                if (!CONVT && A.nt.stats != nullptr) val = in_bwd_apply(s_ca[c + j], s_cb[c + j], s_cd[c + j], g4[b][j], t4[b][j]);
                                s_g[(size_t)lv * PG + c + j] = inb ? val : 0.f;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (CONVT) {
                for (int c = tid; c < Cg; c += NT) {
                    float s = 0.f;
                    for (int lv = 0; lv < PB_V; ++lv) s += s_g[(size_t)lv * PG + c];
                    s_gb[c] += s;
                }
            }
            // ---- dgrad: g_u[v][k] += sum_c g[v][c] * w[c][k]
            if (gu != nullptr) {
                const float *gp = s_g + (size_t)v * PG;
#pragma unroll
                for (int i = 0; i < PB_MAXKC; ++i) {
                    const int k0 = (2 * i + half) * KC;
                    if (k0 < Cu) {
                        for (int c = 0; c < Cg; ++c) {
                            const float g = gp[c];
                            const float *wp = s_w + (size_t)c * Cu + k0;
                            if (KC >= 4) {
#pragma unroll
                                for (int j4 = 0; j4 < KC; j4 += 4) {
                                    const float4 wv = *reinterpret_cast<const float4 *>(wp + j4);
                                    dacc[i][j4] += g * wv.x; dacc[i][j4 + 1] += g * wv.y;
                                    dacc[i][j4 + 2] += g * wv.z; dacc[i][j4 + 3] += g * wv.w;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < KC; ++j) dacc[i][j] += g * wp[j];
                            }
                        }
                    }
                }
            }
            // ---- wgrad: g_w[c][k] += sum_v g[v][c] * u[v][k]
            if (A.g_w != nullptr) {
#pragma unroll
                for (int i = 0; i < MAXT; ++i) {
                    const int tl = ntile <= NT ? tid % ntile : tid + i * NT;
                    const int grp = ntile <= NT ? tid / ntile : 0;
                    if (tl < ntile && grp < G && (ntile > NT || i == 0)) {
                        const int ct = tl / nkt, kt = tl % nkt;
                        const float *gp = s_g + ct * 4;
                        const float *up = s_u + kt * TK;
                        for (int lv = grp; lv < PB_V; lv += G) {
                            const float g0 = gp[(size_t)lv * PG], g1 = gp[(size_t)lv * PG + 1], g2 = gp[(size_t)lv * PG + 2], g3 = gp[(size_t)lv * PG + 3];
#pragma unroll
                            for (int b = 0; b < TK; ++b) {
                                const float uv = up[(size_t)lv * PU + b];
                                wacc[i][0][b] += g0 * uv; wacc[i][1][b] += g1 * uv;
                                wacc[i][2][b] += g2 * uv; wacc[i][3][b] += g3 * uv;
                            }
                        }
                    }
                }
                if (CONVT && !A.tap_split) {   // per-tap flush: the weight slice changes with the tap
#pragma unroll
                    for (int i = 0; i < MAXT; ++i) {
                        const int tl = ntile <= NT ? tid % ntile : tid + i * NT;
                        const int grp = ntile <= NT ? tid / ntile : 0;
                        if (tl < ntile && grp < G && (ntile > NT || i == 0)) {
                            const int ct = tl / nkt, kt = tl % nkt;
#pragma unroll
                            for (int a = 0; a < 4; ++a)
#pragma unroll
                                for (int b = 0; b < TK; ++b) {
                                    atomicAdd(&A.g_w[((size_t)(kt * TK + b) * Cg + ct * 4 + a) * 8 + tap], wacc[i][a][b]);
                                    wacc[i][a][b] = 0.f;
                                }
                        }
                    }
                }
            }
        }
        // ---- store g_u
        if (gu != nullptr && v0 + v < A.vox) {
            float *op = gu + ((size_t)n * A.vox + v0 + v) * (size_t)A.ldgu;
#pragma unroll
            for (int i = 0; i < PB_MAXKC; ++i) {
                const int k0 = (2 * i + half) * KC;
                if (k0 < Cu) {
                    if (KC >= 4) {
#pragma unroll
                        for (int j4 = 0; j4 < KC; j4 += 4) {
                            float4 o = make_float4(dacc[i][j4], dacc[i][j4 + 1], dacc[i][j4 + 2], dacc[i][j4 + 3]);
                            if (A.accumulate) { const float4 p = ld4(op + k0 + j4); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
                            st4(op + k0 + j4, o);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < KC; ++j) {
                            float o = dacc[i][j];
                            if (A.accumulate) o += ld1(op + k0 + j);
                            st1(op + k0 + j, o);
                        }
                    }
                }
            }
        }
    }
    // ---- flush weight gradients: reduce the voxel groups of this CTA in shared memory, then one atomic per output
    if ((!CONVT || A.tap_split) && A.g_w != nullptr) {
        __syncthreads();                                   // s_w (the weights) is no longer needed
        for (int i = tid; i < Cg * Cu; i += NT) s_w[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < MAXT; ++i) {
            const int tl = ntile <= NT ? tid % ntile : tid + i * NT;
            const int grp = ntile <= NT ? tid / ntile : 0;
            if (tl < ntile && grp < G && (ntile > NT || i == 0)) {
                const int ct = tl / nkt, kt = tl % nkt;
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < TK; ++b) {
                        if (G > 1) atomicAdd(&s_w[(size_t)(ct * 4 + a) * Cu + kt * TK + b], wacc[i][a][b]);
                        else s_w[(size_t)(ct * 4 + a) * Cu + kt * TK + b] = wacc[i][a][b];
                    }
            }
        }
        __syncthreads();
        if (CONVT) {
            for (int i = tid; i < Cg * Cu; i += NT) {
                const int c = i / Cu, k = i % Cu;
                atomicAdd(&A.g_w[((size_t)k * Cg + c) * 8 + tap0], s_w[i]);
            }
        } else {
            for (int i = tid; i < Cg * Cu; i += NT) atomicAdd(&A.g_w[i], s_w[i]);
        }
    }
    if (CONVT && A.g_b != nullptr) {
        __syncthreads();
        for (int c = tid; c < Cg; c += NT) atomicAdd(&A.g_b[c], s_gb[c]);
    }
}

// Pointwise backward with a single input channel and no input gradient (the first conv and its shortcut,
// unet3d.py:168,209): g_w[c] = sum_v g_t[v][c] * u[v] is a plain reduction; one thread owns one voxel.
template <typename T, int CG>
__global__ void __launch_bounds__(256) pw_bwd_cu1_kernel(const float *__restrict__ gz, int ldg, const T *__restrict__ t, int ldt, NormDev nt,
                                                         const double *__restrict__ red, const T *__restrict__ u, int ldu, NormDev un,
                                                         int N, long long vox, float *__restrict__ g_w) {
    // g_w[c] = sum_v g_t[v][c] * u[v] is, for a conv that feeds an InstanceNorm, the small residue of terms that cancel (IN(w x)
    // does not depend on |w| but for eps): accumulated in double inside the CTA; one fp32 atomic per CTA and channel at the end
    __shared__ double s_acc[CG];
    __shared__ double s_ca[CG], s_cb[CG], s_cd[CG];
    const int n = blockIdx.y;
    if (threadIdx.x < CG) {
        s_acc[threadIdx.x] = 0.0;
        double a, b, d;
        in_bwd_coef_d(nt, red, N, CG, n, threadIdx.x, a, b, d);
        s_ca[threadIdx.x] = a; s_cb[threadIdx.x] = b; s_cd[threadIdx.x] = d;
    }
    __syncthreads();
    float usc, ush;
    norm_scale_shift(un, N, 1, n, 0, usc, ush);
    const bool has_nt = nt.stats != nullptr;
    double acc[CG];
#pragma unroll
    for (int c = 0; c < CG; ++c) acc[c] = 0.0;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < vox; v += (long long)gridDim.x * blockDim.x) {
        const size_t gv = (size_t)n * vox + v;
        const float uv = lrelu(ld1(u + gv * (size_t)ldu) * usc + ush, un.slope);
#pragma unroll
        for (int c4 = 0; c4 < CG; c4 += 4) {
            float g4[4], t4[4] = {0.f, 0.f, 0.f, 0.f};
            ld4a(gz + gv * (size_t)ldg + c4, g4);
            if (has_nt) ld4a(t + gv * (size_t)ldt + c4, t4);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[c4 + j] = fma(fma(s_ca[c4 + j], (double)g4[j], fma(s_cb[c4 + j], (double)t4[j], s_cd[c4 + j])), (double)uv, acc[c4 + j]);
        }
    }
#pragma unroll
    for (int c = 0; c < CG; ++c) {
        const double s = warp_sum(acc[c]);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[c], s);
    }
    __syncthreads();
    if (threadIdx.x < CG) atomicAdd(&g_w[threadIdx.x], (float)s_acc[threadIdx.x]);
}

static size_t pw_bwd_smem(int Cg, int Cu) {
    size_t fl = (((size_t)Cg * Cu + 3) & ~(size_t)3) + (size_t)PB_V * (Cg | 1) + (size_t)PB_V * (Cu | 1) + 7 * (size_t)Cg + 2 + 2 * (size_t)Cu;   /* 3 x Cg doubles (+ alignment) + the bias-gradient row */
    return fl * sizeof(float);
}

// ===========================================================================================
// Depthwise backward.  Same 4x8x8 tile / 6x10x10 halo geometry as the forward stencil.
constexpr int TZ = 4, TY = 8, TX = 8, TV = TZ * TY * TX;
constexpr int HZ = TZ + 2, HY = TY + 2, HX = TX + 2;
constexpr int HXP = HX + 1, HPLANE = HY * HXP + 1, HVOX = HZ * HPLANE;
constexpr int CK = 16;

template <typename T>
__global__ void __launch_bounds__(NT, 2) dw_bwd_kernel(
    const float *__restrict__ g_u, int ldgu, const T *__restrict__ x, int ldx, NormDev xn, int C,
    int N, int D, int H, int W, const float *__restrict__ dw_w, float *__restrict__ g_dw,
    float *__restrict__ gy, int ldgy, int accumulate, double *__restrict__ redx) {
    extern __shared__ __align__(16) float sm[];
    float *s_gu = sm;                        // HVOX*CK
    float *s_a = s_gu + HVOX * CK;           // HVOX*CK   (re-used as the gy staging tile [TV][CK])
    float *s_scale = s_a + HVOX * CK;        // C  (combined scale incl. dropout keep-scale)
    float *s_shift = s_scale + C;
    float *s_m = s_shift + C;                // dropout keep-scale
    float *s_mean = s_m + C, *s_rstd = s_mean + C;
    // 2*C doubles: the reductions of the producer's InstanceNorm backward (double from the first atomic on, see merge_bwd_kernel)
    double *s_red = reinterpret_cast<double *>(sm + ((2 * HVOX * CK + 5 * C + 1) & ~1));
    float *s_gdw = reinterpret_cast<float *>(s_red + 2 * C);            // C*27 (accumulated over all tiles of this CTA)
    const int tid = threadIdx.x;
    const bool has_norm = xn.stats != nullptr;
    const int tilesX = (W + TX - 1) / TX, tilesY = (H + TY - 1) / TY, tilesZ = (D + TZ - 1) / TZ;
    const long long tiles_per_sample = (long long)tilesX * tilesY * tilesZ;
    const long long total_tiles = tiles_per_sample * N;
    for (int i = tid; i < C * 27; i += NT) s_gdw[i] = 0.f;
    for (int i = tid; i < 2 * C; i += NT) s_red[i] = 0.0;
    const int c = tid & 15, g = tid >> 4;
    const int lz = (g & 1) + 2 * (g >> 3);
    const int ly0 = 2 * ((g >> 1) & 3);
    int cur_n = -1;
    auto flush_red = [&](int n) {
        if (!has_norm || redx == nullptr || n < 0) return;
        for (int i = tid; i < 2 * C; i += NT) {
            const int isq = i >= C;
            const int cc = isq ? i - C : i;
            atomicAdd(&redx[(size_t)isq * N * C + (size_t)n * C + cc], s_red[i]);
            s_red[i] = 0.0;
        }
    };
    // work item = (tile, 16-channel chunk): the deep layers have few tiles and many chunks, so the chunks of one tile
    // spread over the CTAs instead of running back to back in one
    const int nchunks = (C + CK - 1) / CK;
    const long long total_items = total_tiles * nchunks;
    // 32-bit work-item arithmetic (the host checks total_items < 2^31): the 64-bit divisions this loop started with cost every
    // thread a few hundred cycles per 256-voxel item
    const uint32_t n_items32 = (uint32_t)total_items, tps32 = (uint32_t)tiles_per_sample;
    for (uint32_t item = blockIdx.x; item < n_items32; item += gridDim.x) {
        const uint32_t tile = item / (uint32_t)nchunks;
        const int c0 = (int)(item - tile * (uint32_t)nchunks) * CK;
        const int n = (int)(tile / tps32);
        uint32_t b = tile - (uint32_t)n * tps32;
        const int x0 = (int)(b % (uint32_t)tilesX) * TX; b /= (uint32_t)tilesX;
        const int y0 = (int)(b % (uint32_t)tilesY) * TY; b /= (uint32_t)tilesY;
        const int z0 = (int)b * TZ;
        __syncthreads();
        flush_red(cur_n);      // per item: keeps the fp32 shared-memory partial sums short (double beyond this point)
        if (n != cur_n) {
            cur_n = n;
            for (int cc = tid; cc < C; cc += NT) {
                const NormCoef k = norm_coef(xn, N, C, n, cc);
                float sc = 1.f, sh = 0.f;
                if (has_norm) { sc = k.gamma * k.rstd * k.m; sh = (k.beta - k.mean * k.gamma * k.rstd) * k.m; }
                s_scale[cc] = sc; s_shift[cc] = sh; s_m[cc] = k.m; s_mean[cc] = k.mean; s_rstd[cc] = k.rstd;
            }
            __syncthreads();
        }
        {
            // stage halo tiles of g_u and of the activated input a: (voxel, 4-channel quad) items, 4 in flight per thread
            {
                const bool vec = (C % 4 == 0) && (ldx % 4 == 0) && (ldgu % 4 == 0);
                constexpr int NITEM = HZ * HY * HX * (CK / 4);
                // item = (halo voxel, 4-channel quad); consecutive items of a thread are NT / 4 = 64 voxels apart, so the halo
                // coordinates advance by (+6 rows, +4 columns) with carries instead of being re-derived by divisions
                static_assert(NT / (CK / 4) == 64 && HX == 10 && HY == 10, "incremental halo walk assumes 64 voxels per step of a 10 x 10 plane");
                const int q = tid & (CK / 4 - 1);
                int whx = (tid >> 2) % HX, why = ((tid >> 2) / HX) % HY, whz = (tid >> 2) / (HX * HY);
                for (int item0 = tid; item0 < NITEM; item0 += 4 * NT) {
                    float gv[4][4], xv[4][4];
                    bool inb[4];
                    int hxs[4], hys[4], hzs[4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int item = item0 + b * NT;
                        gv[b][0] = gv[b][1] = gv[b][2] = gv[b][3] = 0.f;
                        xv[b][0] = xv[b][1] = xv[b][2] = xv[b][3] = 0.f;
                        inb[b] = false;
                        hxs[b] = whx; hys[b] = why; hzs[b] = whz;
                        whx += 4; if (whx >= HX) { whx -= HX; ++why; }
                        why += 6; if (why >= HY) { why -= HY; ++whz; }
                        if (item < NITEM) {
                            const int hx = hxs[b], hy = hys[b], hz = hzs[b];
                            const int gz_ = z0 + hz - 1, gy_ = y0 + hy - 1, gx_ = x0 + hx - 1;
                            const int cc = c0 + q * 4;
                            if (gz_ >= 0 && gz_ < D && gy_ >= 0 && gy_ < H && gx_ >= 0 && gx_ < W && cc < C) {
                                inb[b] = true;
                                const size_t vox = (((size_t)n * D + gz_) * H + gy_) * W + gx_;
                                if (vec) {
                                    ld4a(g_u + vox * (size_t)ldgu + cc, gv[b]);
                                    ld4a(x + vox * (size_t)ldx + cc, xv[b]);
                                } else {
                                    for (int j = 0; j < 4; ++j)
                                        if (cc + j < C) { gv[b][j] = ld1(g_u + vox * (size_t)ldgu + cc + j); xv[b][j] = ld1(x + vox * (size_t)ldx + cc + j); }
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int item = item0 + b * NT;
                        if (item < NITEM) {
                            const int hx = hxs[b], hy = hys[b], hz = hzs[b];
                            const int cc = c0 + q * 4;
                            float4 go = make_float4(0.f, 0.f, 0.f, 0.f), ao = go;
                            if (inb[b]) {
                                float av[4];
                                for (int j = 0; j < 4; ++j)
                                    av[j] = (cc + j < C) ? lrelu(xv[b][j] * s_scale[min(cc + j, C - 1)] + s_shift[min(cc + j, C - 1)], xn.slope) : 0.f;
                                go = make_float4(gv[b][0], gv[b][1], gv[b][2], gv[b][3]);
                                ao = make_float4(av[0], av[1], av[2], av[3]);
                            }
                            const int so = (hz * HPLANE + hy * HXP + hx) * CK + q * 4;
                            *reinterpret_cast<float4 *>(s_gu + so) = go;
                            *reinterpret_cast<float4 *>(s_a + so) = ao;
                        }
                    }
                }
            }
            __syncthreads();
            float ga0[TX], ga1[TX];
            float asign0[TX], asign1[TX];
            const bool act = c0 + c < C;
            if (act) {
                float wreg[27];
#pragma unroll
                for (int k = 0; k < 27; ++k) wreg[k] = dw_w[(size_t)(c0 + c) * 27 + k];
                float wacc[27];
#pragma unroll
                for (int k = 0; k < 27; ++k) wacc[k] = 0.f;
#pragma unroll
                for (int i = 0; i < TX; ++i) { ga0[i] = 0.f; ga1[i] = 0.f; }
                // centre g_u values of this thread's 2 rows x 8 voxels
                float gc0[TX], gc1[TX];
                {
                    const float *p0 = s_gu + (size_t)((lz + 1) * HPLANE + (ly0 + 1) * HXP + 1) * CK + c;
                    const float *p1 = p0 + (size_t)HXP * CK;
                    const float *q0 = s_a + (size_t)((lz + 1) * HPLANE + (ly0 + 1) * HXP + 1) * CK + c;
                    const float *q1 = q0 + (size_t)HXP * CK;
#pragma unroll
                    for (int i = 0; i < TX; ++i) {
                        gc0[i] = p0[i * CK]; gc1[i] = p1[i * CK];
                        asign0[i] = q0[i * CK]; asign1[i] = q1[i * CK];
                    }
                }
#pragma unroll
                for (int dz = 0; dz < 3; ++dz) {
#pragma unroll
                    for (int hy = 0; hy < 4; ++hy) {
                        const size_t ro = (size_t)((lz + dz) * HPLANE + (ly0 + hy) * HXP) * CK + c;
                        float grow[HX], arow[HX];
#pragma unroll
                        for (int hx = 0; hx < HX; ++hx) { grow[hx] = s_gu[ro + hx * CK]; arow[hx] = s_a[ro + hx * CK]; }
                        if (hy <= 2) {
                            // dgrad with the flipped stencil: weight index (2-dz, 2-hy, 2-dx)
                            const float w0 = wreg[(2 - dz) * 9 + (2 - hy) * 3 + 2], w1 = wreg[(2 - dz) * 9 + (2 - hy) * 3 + 1], w2 = wreg[(2 - dz) * 9 + (2 - hy) * 3];
#pragma unroll
                            for (int i = 0; i < TX; ++i) ga0[i] = fmaf(w2, grow[i + 2], fmaf(w1, grow[i + 1], fmaf(w0, grow[i], ga0[i])));
                            // wgrad: tap (dz, hy, dx) pairs centre g_u with a at offset
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                float s = wacc[dz * 9 + hy * 3 + dx];
#pragma unroll
                                for (int i = 0; i < TX; ++i) s = fmaf(gc0[i], arow[i + dx], s);
                                wacc[dz * 9 + hy * 3 + dx] = s;
                            }
                        }
                        if (hy >= 1) {
                            const int dy = hy - 1;
                            const float w0 = wreg[(2 - dz) * 9 + (2 - dy) * 3 + 2], w1 = wreg[(2 - dz) * 9 + (2 - dy) * 3 + 1], w2 = wreg[(2 - dz) * 9 + (2 - dy) * 3];
#pragma unroll
                            for (int i = 0; i < TX; ++i) ga1[i] = fmaf(w2, grow[i + 2], fmaf(w1, grow[i + 1], fmaf(w0, grow[i], ga1[i])));
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                float s = wacc[dz * 9 + dy * 3 + dx];
#pragma unroll
                                for (int i = 0; i < TX; ++i) s = fmaf(gc1[i], arow[i + dx], s);
                                wacc[dz * 9 + dy * 3 + dx] = s;
                            }
                        }
                    }
                }
                if (g_dw != nullptr) {
#pragma unroll
                    for (int k = 0; k < 27; ++k) atomicAdd(&s_gdw[(c0 + c) * 27 + k], wacc[k]);
                }
            }
            __syncthreads();   // every thread is done with s_gu / s_a of this chunk
            if (gy != nullptr) {
                // through Dropout3d / LeakyReLU of the producer, staged for coalesced stores
                float *s_stage = s_a;   // [TV][CK]
                if (act) {
                    const float m = s_m[c0 + c];
#pragma unroll
                    for (int i = 0; i < TX; ++i) {
                        float v0 = ga0[i], v1 = ga1[i];
                        if (has_norm) {
                            v0 *= m * (asign0[i] > 0.f ? 1.f : xn.slope);
                            v1 *= m * (asign1[i] > 0.f ? 1.f : xn.slope);
                        }
                        s_stage[(size_t)((lz * TY + ly0) * TX + i) * CK + c] = v0;
                        s_stage[(size_t)((lz * TY + ly0 + 1) * TX + i) * CK + c] = v1;
                    }
                }
                __syncthreads();
                // thread -> (voxel, 4-channel group); the group is the same for all items of a thread (NT % 4 == 0), so
                // the norm reductions are kept in registers and combined across the 8 lanes sharing a group
                float rs[4] = {0.f, 0.f, 0.f, 0.f}, rx[4] = {0.f, 0.f, 0.f, 0.f};
                const int q = tid & (CK / 4 - 1);
                const int cc = c0 + q * 4;
                for (int item = tid; item < TV * (CK / 4); item += NT) {
                    const int lv = item / (CK / 4);
                    const int lx = lv & 7, ly = (lv >> 3) & 7, lzz = lv >> 6;
                    const int gz_ = z0 + lzz, gy_ = y0 + ly, gx_ = x0 + lx;
                    if (gz_ < D && gy_ < H && gx_ < W && cc < C) {
                        const size_t vox = (((size_t)n * D + gz_) * H + gy_) * W + gx_;
                        float val[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) val[j] = s_stage[(size_t)lv * CK + q * 4 + j];
                        float *op = gy + vox * (size_t)ldgy + cc;
                        const bool vec = (cc + 3 < C) && (ldgy % 4 == 0) && (C % 4 == 0);
                        if (vec) {
                            float4 o = make_float4(val[0], val[1], val[2], val[3]);
                            if (accumulate) { const float4 p = ld4(op); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
                            st4(op, o);
                        } else {
                            for (int j = 0; j < 4; ++j) if (cc + j < C) {
                                float o = val[j];
                                if (accumulate) o += ld1(op + j);
                                st1(op + j, o);
                            }
                        }
                        if (has_norm && redx != nullptr) {
                            const T *xp = x + vox * (size_t)ldx + cc;
                            float xr[4] = {0.f, 0.f, 0.f, 0.f};
                            if ((cc + 3 < C) && (ldx % 4 == 0)) ld4a(xp, xr);
                            else for (int j = 0; j < 4; ++j) if (cc + j < C) xr[j] = ld1(xp + j);
#pragma unroll
                            for (int j = 0; j < 4; ++j) if (cc + j < C) {
                                rs[j] += val[j];
                                rx[j] += val[j] * ((xr[j] - s_mean[cc + j]) * s_rstd[cc + j]);
                            }
                        }
                    }
                }
                if (has_norm && redx != nullptr) {
                    // the thread's own (fixed-order, few-term) fp32 partial sums go on in double: everything combined in a
                    // run-dependent order is double
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double ds = (double)rs[j], dx = (double)rx[j];
#pragma unroll
                        for (int sft = 4; sft <= 16; sft <<= 1) {
                            ds += __shfl_xor_sync(0xffffffffu, ds, sft);
                            dx += __shfl_xor_sync(0xffffffffu, dx, sft);
                        }
                        if ((tid & 31) < 4 && cc + j < C) {
                            atomicAdd(&s_red[cc + j], ds);
                            atomicAdd(&s_red[C + cc + j], dx);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    flush_red(cur_n);
    if (g_dw != nullptr) {
        for (int i = tid; i < C * 27; i += NT) {
            const float vv = s_gdw[i];
            if (vv != 0.f) atomicAdd(&g_dw[i], vv);
        }
    }
}

// Single-channel depthwise backward without an input gradient (the network's first conv, unet3d.py:168,209: the
// image needs no gradient): only g_dw[27] = sum_v g_u[v] * a[v + tap].  The generic kernel would run a 16-channel
// chunk with one live lane in sixteen; here one thread owns one voxel.
template <typename T>
__global__ void __launch_bounds__(256) dw_bwd_c1_wgrad_kernel(const float *__restrict__ g_u, int ldgu, const T *__restrict__ x, int ldx,
                                                              NormDev xn, int N, int D, int H, int W, float *__restrict__ g_dw) {
    __shared__ float s_acc[27];
    if (threadIdx.x < 27) s_acc[threadIdx.x] = 0.f;
    __syncthreads();
    const size_t vox_per = (size_t)D * H * W, total = vox_per * N;
    float acc[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float g = g_u[i * (size_t)ldgu];
        const int n = (int)((uint32_t)i / (uint32_t)vox_per);            // total < 2^31 (host check)
        uint32_t rem = (uint32_t)i - (uint32_t)n * (uint32_t)vox_per;
        const int xx = (int)(rem % (uint32_t)W); rem /= (uint32_t)W;
        const int yy = (int)(rem % (uint32_t)H);
        const int zz = (int)(rem / (uint32_t)H);
        float sc, sh;
        norm_scale_shift(xn, N, 1, n, 0, sc, sh);
#pragma unroll
        for (int dz = 0; dz < 3; ++dz)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int z = zz + dz - 1, y = yy + dy - 1, xq = xx + dx - 1;
                    if (z >= 0 && z < D && y >= 0 && y < H && xq >= 0 && xq < W) {
                        const float a = lrelu(ld1(x + ((((size_t)n * D + z) * H + y) * W + xq) * (size_t)ldx) * sc + sh, xn.slope);
                        acc[dz * 9 + dy * 3 + dx] = fmaf(g, a, acc[dz * 9 + dy * 3 + dx]);
                    }
                }
    }
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        const float v = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[k], v);
    }
    __syncthreads();
    if (threadIdx.x < 27) atomicAdd(&g_dw[threadIdx.x], s_acc[threadIdx.x]);
}

static size_t dw_bwd_smem(int C) {
    return sizeof(float) * (2 * (size_t)HVOX * CK + 5 * (size_t)C + 2 + 4 * (size_t)C + 27 * (size_t)C);   /* 2 x C doubles (+ alignment) for the reductions */
}

__global__ void norm_param_grad_kernel(const double *__restrict__ red, int N, int C, float *__restrict__ g_gamma, float *__restrict__ g_beta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double sb = 0.0, sg = 0.0;
    for (int n = 0; n < N; ++n) { sb += red[(size_t)n * C + c]; sg += red[(size_t)N * C + (size_t)n * C + c]; }
    g_beta[c] += (float)sb;
    g_gamma[c] += (float)sg;
}

// all InstanceNorm affine gradients of one backward pass in a single launch: blockIdx.x = norm
constexpr int NPG_MAX = 32;
struct NormGradBatch {
    const double *red[NPG_MAX];
    float *g_gamma[NPG_MAX], *g_beta[NPG_MAX];
    int C[NPG_MAX];
};
__global__ void norm_param_grad_batch_kernel(NormGradBatch B, int N) {
    const int i = blockIdx.x, C = B.C[i];
    const double *red = B.red[i];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double sb = 0.0, sg = 0.0;
        for (int n = 0; n < N; ++n) { sb += red[(size_t)n * C + c]; sg += red[(size_t)N * C + (size_t)n * C + c]; }
        B.g_beta[i][c] += (float)sb;
        B.g_gamma[i][c] += (float)sg;
    }
}

// ===========================================================================================
// Dense / grouped 3x3x3 backward (generic CUDA-core path).
// (1) g_t = a*gz + b*t + d materialised once
template <typename T>
__global__ void __launch_bounds__(NT) c3_gt_kernel(const float *__restrict__ gz, int ldg, const T *__restrict__ t, int ldt, NormDev nt,
                                                   const double *__restrict__ red, int N, int C, size_t vox, float *__restrict__ gt) {
    const size_t total = (size_t)N * vox * C;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int c = (int)(i % C);
        const size_t v = i / C;
        const int n = (int)(v / vox);
        double a, b, d;
        in_bwd_coef_d(nt, red, N, C, n, c, a, b, d);
        const float tv = nt.stats != nullptr ? ld1(t + v * ldt + c) : 0.f;
        st1(gt + i, in_bwd_apply(a, b, d, ld1(gz + v * ldg + c), tv));
    }
}
// (2) flipped / transposed weights for the dgrad-as-forward-conv: wT[ci][col][tap] = w[co][cil][26-tap]
__global__ void c3_flip_w_kernel(const float *__restrict__ w, int Cin, int Cout, int groups, float *__restrict__ wT) {
    const int cin_g = Cin / groups, cout_g = Cout / groups;
    const int total = Cin * cout_g * 27;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tap = i % 27;
        const int col = (i / 27) % cout_g;
        const int ci = i / (27 * cout_g);
        const int g = ci / cin_g, cil = ci - g * cin_g;
        const int co = g * cout_g + col;
        wT[i] = w[((size_t)co * cin_g + cil) * 27 + (26 - tap)];
    }
}
// (3) back through Dropout3d / LeakyReLU of the producer + its norm reductions (elementwise)
template <typename T>
__global__ void __launch_bounds__(NT) c3_act_bwd_kernel(const float *__restrict__ ga, const T *__restrict__ x, int ldx, NormDev xn,
                                                        int N, int C, size_t vox, float *__restrict__ gy, int ldgy, int accumulate,
                                                        double *__restrict__ redx) {
    extern __shared__ __align__(16) float sm[];
    double *s_red = reinterpret_cast<double *>(sm);   // 2*C doubles (reductions in double: see merge_bwd_kernel)
    float *s_mean = sm + 4 * C, *s_rstd = sm + 5 * C, *s_gam = sm + 6 * C, *s_bet = sm + 7 * C, *s_m = sm + 8 * C;
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < 2 * C; i += NT) s_red[i] = 0.0;
    for (int c = threadIdx.x; c < C; c += NT) {
        const NormCoef k = norm_coef(xn, N, C, n, c);
        s_mean[c] = k.mean; s_rstd[c] = k.rstd; s_gam[c] = k.gamma; s_bet[c] = k.beta; s_m[c] = k.m;
    }
    __syncthreads();
    const bool has_norm = xn.stats != nullptr;
    const size_t total = vox * C;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int c = (int)(i % C);
        const size_t v = (size_t)n * vox + i / C;
        float g = ld1(ga + v * C + c);
        if (has_norm) {
            const float xv = ld1(x + v * ldx + c);
            const float xh = (xv - s_mean[c]) * s_rstd[c];
            const float y = s_gam[c] * xh + s_bet[c];
            g *= s_m[c] * (y > 0.f ? 1.f : xn.slope);
            const float gr = round_as(gy, g);
            atomicAdd(&s_red[c], (double)gr);
            atomicAdd(&s_red[C + c], (double)gr * (double)xh);
        }
        float *op = gy + v * ldgy + c;
        if (accumulate) g += ld1(op);
        st1(op, g);
    }
    __syncthreads();
    if (has_norm && redx != nullptr)
        for (int i = threadIdx.x; i < 2 * C; i += NT) {
            const int isq = i >= C, c = isq ? i - C : i;
            atomicAdd(&redx[(size_t)isq * N * C + (size_t)n * C + c], s_red[i]);
        }
}
// (4) wgrad: g_w[co][cil][tap] += sum_vox g_t[vox][co] * a[vox + off(tap)][ci].  A CTA owns an 8x8 block of
// (co, ci) pairs and a slice of the spatial tiles; thread = one pair x 8 of the tile's 32 x-rows.
constexpr int WG_C = 8;
template <typename T>
__global__ void __launch_bounds__(NT) c3_wgrad_kernel(const float *__restrict__ gt, const T *__restrict__ x, int ldx, NormDev xn,
                                                      int N, int Cin, int Cout, int groups, int D, int H, int W,
                                                      const int2 *__restrict__ pairs, int npairs, int splits,
                                                      float *__restrict__ g_w) {
    __shared__ float s_a[WG_C][HZ * HY * HX];
    __shared__ float s_g[WG_C][TV];
    __shared__ float s_acc[WG_C * WG_C * 27];
    __shared__ float s_sc[WG_C], s_sh[WG_C];
    const int tid = threadIdx.x;
    const int pair = blockIdx.x % npairs, split = blockIdx.x / npairs;
    const int cb = pairs[pair].x, c0 = pairs[pair].y;     // first output / input channel of the block
    const int cin_g = Cin / groups, cout_g = Cout / groups;
    const int tilesX = (W + TX - 1) / TX, tilesY = (H + TY - 1) / TY, tilesZ = (D + TZ - 1) / TZ;
    const long long tiles_per_sample = (long long)tilesX * tilesY * tilesZ;
    const long long total_tiles = tiles_per_sample * N;
    const int pj = tid & 63, rg = tid >> 6;               // pair index within the block, row group
    const int jo = pj >> 3, ji = pj & 7;                  // output / input channel offsets
    float acc[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = 0.f;
    for (int i = tid; i < WG_C * WG_C * 27; i += NT) s_acc[i] = 0.f;
    for (long long tile = split; tile < total_tiles; tile += splits) {
        const int n = (int)(tile / tiles_per_sample);
        int b = (int)(tile % tiles_per_sample);
        const int x0 = (b % tilesX) * TX; b /= tilesX;
        const int y0 = (b % tilesY) * TY; b /= tilesY;
        const int z0 = b * TZ;
        __syncthreads();
        if (tid < WG_C) {
            float sc = 0.f, sh = 0.f;
            if (c0 + tid < Cin) norm_scale_shift(xn, N, Cin, n, c0 + tid, sc, sh);
            s_sc[tid] = sc; s_sh[tid] = sh;
        }
        __syncthreads();
        for (int item = tid; item < WG_C * HZ * HY * HX; item += NT) {
            const int ci = item / (HZ * HY * HX);
            int hv = item % (HZ * HY * HX);
            const int hx = hv % HX; hv /= HX;
            const int hy = hv % HY;
            const int hz = hv / HY;
            const int iz = z0 + hz - 1, iy = y0 + hy - 1, ix = x0 + hx - 1;
            const int c = c0 + ci;
            float val = 0.f;
            if (c < Cin && iz >= 0 && iz < D && iy >= 0 && iy < H && ix >= 0 && ix < W) {
                val = lrelu(ld1(x + ((((size_t)n * D + iz) * H + iy) * W + ix) * (size_t)ldx + c) * s_sc[ci] + s_sh[ci], xn.slope);
            }
            s_a[ci][item % (HZ * HY * HX)] = val;
        }
        for (int item = tid; item < WG_C * TV; item += NT) {
            const int co = item % WG_C, lv = item / WG_C;
            const int lx = lv & 7, ly = (lv >> 3) & 7, lz = lv >> 6;
            const int gz_ = z0 + lz, gy_ = y0 + ly, gx_ = x0 + lx;
            float val = 0.f;
            if (cb + co < Cout && gz_ < D && gy_ < H && gx_ < W)
                val = ld1(gt + ((((size_t)n * D + gz_) * H + gy_) * W + gx_) * (size_t)Cout + cb + co);
            s_g[co][lv] = val;
        }
        __syncthreads();
        for (int row = rg; row < TZ * TY; row += 4) {
            const int lz = row >> 3, ly = row & 7;
            float g[TX];
#pragma unroll
            for (int i = 0; i < TX; ++i) g[i] = s_g[jo][row * TX + i];
#pragma unroll
            for (int dz = 0; dz < 3; ++dz)
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const float *ap = &s_a[ji][((lz + dz) * HY + ly + dy) * HX];
                    float a[HX];
#pragma unroll
                    for (int i = 0; i < HX; ++i) a[i] = ap[i];
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        float s = 0.f;
#pragma unroll
                        for (int i = 0; i < TX; ++i) s += g[i] * a[i + dx];
                        acc[dz * 9 + dy * 3 + dx] += s;
                    }
                }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 27; ++k) atomicAdd(&s_acc[pj * 27 + k], acc[k]);
    __syncthreads();
    for (int i = tid; i < WG_C * WG_C * 27; i += NT) {
        const int k = i % 27, p2 = i / 27;
        const int co = cb + (p2 >> 3), ci = c0 + (p2 & 7);
        if (co < Cout && ci < Cin && co / cout_g == ci / cin_g)
            atomicAdd(&g_w[((size_t)co * cin_g + (ci - (ci / cin_g) * cin_g)) * 27 + k], s_acc[i]);
    }
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

template <typename T, bool CONVT>
static int launch_pw_bwd(const PwBwdArgs &A, cudaStream_t st) {
    const int Cg = A.Cg, Cu = A.Cu;
    const int KC = (Cu % 16 == 0) ? 16 : (Cu % 4 == 0) ? 4 : 1;
    L3D_REQUIRE(Cu <= 2 * KC * PB_MAXKC, "pointwise backward: Cin=%d is not supported by the %d-wide register tiling", Cu, KC);
    const int TK = KC >= 4 ? 4 : 1;
    const int ntile = (Cg / 4) * (Cu / TK);
    L3D_REQUIRE(ntile <= 4 * NT, "pointwise backward: Cout*Cin=%d too large", Cg * Cu);
    const int MAXT = ntile <= NT ? 1 : 4;
    const size_t smem = pw_bwd_smem(Cg, Cu);
    L3D_REQUIRE(smem <= 227 * 1024, "pointwise backward: Cout=%d Cin=%d needs %zu B shared memory", Cg, Cu, smem);
    const long long tiles = ((A.vox + PB_V - 1) / PB_V) * A.N;
    const int ctas_per_sm = smem > 110 * 1024 ? 1 : 2;     // 256 threads x up to 128 registers
    long long grid = tiles < 148ll * ctas_per_sm ? tiles : 148ll * ctas_per_sm;
    if (A.tap_split) grid = (grid + 7) / 8;                // 8 taps share the machine
    if (grid < 1) grid = 1;
    const dim3 grid3((unsigned)grid, A.tap_split ? 8u : 1u);
#define L3D_PWB(KCV, MT)                                                             \
    do {                                                                             \
        auto kern = pw_bwd_kernel<T, KCV, MT, CONVT>;                                \
        if (set_smem(kern, smem)) return 3;                                          \
        kern<<<grid3, NT, smem, st>>>(A);                                            \
    } while (0)
    if (KC == 16) { if (MAXT == 1) L3D_PWB(16, 1); else L3D_PWB(16, 4); }
    else if (KC == 4) { if (MAXT == 1) L3D_PWB(4, 1); else L3D_PWB(4, 4); }
    else { if (MAXT == 1) L3D_PWB(1, 1); else L3D_PWB(1, 4); }
#undef L3D_PWB
    return 0;
}

}  // namespace

// ================================================================= C ABI ====================
extern "C" int l3d_merge_bwd(const l3d_act *g_out, const l3d_act *pooled_g, const l3d_act *out, const l3d_act *pooled,
                             const l3d_act *t2, const l3d_norm *n2, const l3d_act *r, const l3d_norm *nr,
                             int N, int D, int H, int W, float slope,
                             const float *head_w, int OC, const float *g_prob, const float *prob,
                             float *g_head_w, float *g_head_b,
                             const l3d_act *gz, double *red2, double *redr, void *stream) {
    L3D_REQUIRE(!act_null(out) && !act_null(t2) && !act_null(r) && !act_null(gz) && red2 && n2 && n2->stats, "l3d_merge_bwd: null argument");
    const int C = t2->C;
    const bool has_go = !act_null(g_out), has_pg = !act_null(pooled_g);
    const bool has_rn = nr != nullptr && nr->stats != nullptr;
    L3D_REQUIRE(C % 4 == 0 && out->C == C && r->C == C && gz->C == C, "l3d_merge_bwd: channel mismatch");
    L3D_REQUIRE(vec4_ok(out) && vec4_ok(t2) && vec4_ok(r) && vec4_ok(gz), "l3d_merge_bwd: views must be 4-channel aligned");
    L3D_REQUIRE(!has_rn || redr, "l3d_merge_bwd: shortcut norm needs redr");
    if (has_go) L3D_REQUIRE(g_out->C == C && vec4_ok(g_out) && g_out->dtype == L3D_F32, "l3d_merge_bwd: bad g_out (gradients are fp32)");
    if (has_pg) L3D_REQUIRE(!act_null(pooled) && pooled->C == C && pooled_g->C == C && vec4_ok(pooled) && vec4_ok(pooled_g) && pooled_g->dtype == L3D_F32, "l3d_merge_bwd: bad pooled views");
    L3D_REQUIRE(out->dtype == t2->dtype && r->dtype == t2->dtype && gz->dtype == L3D_F32, "l3d_merge_bwd: dtype mismatch (gz is fp32)");
    const NormDev d2 = norm_dev(n2), dr = norm_dev(nr);
    cudaStream_t st = (cudaStream_t)stream;
    if (head_w != nullptr) {
        L3D_REQUIRE(!has_pg, "l3d_merge_bwd: head and pool cannot be combined");
        L3D_REQUIRE(C <= 64 && OC >= 1 && OC <= 4 && g_prob && prob && g_head_w && g_head_b, "l3d_merge_bwd: head needs C <= 64 and OC <= 4 (got %d, %d)", C, OC);
        const size_t nvox = (size_t)D * H * W;
        size_t gx = (nvox + NT - 1) / NT;
        const size_t cap = (148 * 8 + N - 1) / N;
        if (gx > cap) gx = cap;
        dim3 grid((unsigned)gx, (unsigned)N);
#define L3D_MHB(CM, OM)                                                                                                   \
        merge_head_bwd_kernel<T, CM, OM><<<grid, NT, 0, st>>>(has_go ? (const float *)g_out->ptr : nullptr, has_go ? g_out->ldc : 0, \
            (const T *)out->ptr, out->ldc, (const T *)t2->ptr, t2->ldc, d2, (const T *)r->ptr, r->ldc, dr, N, C, nvox, slope, \
            head_w, OC, g_prob, prob, g_head_w, g_head_b, (float *)gz->ptr, gz->ldc, red2, redr)
        L3D_DISPATCH_DTYPE(t2->dtype, T, {
            if (OC == 1) { if (C <= 16) L3D_MHB(16, 1); else if (C <= 32) L3D_MHB(32, 1); else L3D_MHB(64, 1); }
            else { if (C <= 16) L3D_MHB(16, 4); else if (C <= 32) L3D_MHB(32, 4); else L3D_MHB(64, 4); }
        });
#undef L3D_MHB
    } else {
        L3D_REQUIRE(has_go || has_pg, "l3d_merge_bwd: no incoming gradient");
        const size_t total = (size_t)((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
        L3D_REQUIRE(total < (1ull << 31), "l3d_merge_bwd: sample too large for 32-bit cell indices");
        size_t gx = (total + NT - 1) / NT;
        const size_t cap = (148 * 16 + N - 1) / N;
        if (gx > cap) gx = cap;
        dim3 grid((unsigned)gx, (unsigned)N);
        const size_t smem = sizeof(float) * 4 * (size_t)C + sizeof(double) * 3 * (size_t)C;
        L3D_DISPATCH_DTYPE(t2->dtype, T, {
            merge_bwd_kernel<T><<<grid, NT, smem, st>>>(has_go ? (const float *)g_out->ptr : nullptr, has_go ? g_out->ldc : 0,
                has_pg ? (const float *)pooled_g->ptr : nullptr, has_pg ? pooled_g->ldc : 0, (const T *)out->ptr, out->ldc,
                has_pg ? (const T *)pooled->ptr : nullptr, has_pg ? pooled->ldc : 0, (const T *)t2->ptr, t2->ldc, d2,
                (const T *)r->ptr, r->ldc, dr, N, C, D, H, W, slope, (float *)gz->ptr, gz->ldc, red2, redr);
        });
    }
    l3d_count_launch();
    L3D_CUDA_OK("l3d_merge_bwd launch");
    return 0;
}

int l3d_pw_bwd_tc(const l3d_act *gz, const l3d_act *t, const l3d_norm *nt, const double *red, const l3d_act *u,
                  const l3d_norm *un, int N, long long vox, const float *w, float *g_w, const l3d_act *g_u,
                  int accumulate_gu, void *stream);

int l3d_convt_bwd_tc(const l3d_act *g_out, int OD, int OH, int OW, int oz, int oy, int ox, const l3d_act *x, int N, int d, int h, int w_,
                     const float *w, float *g_w, float *g_b, const l3d_act *g_x, int accumulate_gx, void *stream);

extern "C" int l3d_pw_bwd(const l3d_act *gz, const l3d_act *t, const l3d_norm *nt, const double *red,
                          const l3d_act *u, const l3d_norm *un, int N, int D, int H, int W,
                          const float *w, float *g_w, const l3d_act *g_u, int accumulate_gu, void *stream) {
    L3D_REQUIRE(!act_null(gz) && !act_null(u) && w, "l3d_pw_bwd: null argument");
    const bool has_nt = nt != nullptr && nt->stats != nullptr;
    if (has_nt) L3D_REQUIRE(!act_null(t) && red && t->C == gz->C && vec4_ok(t) && t->dtype == u->dtype, "l3d_pw_bwd: bad t / red");
    L3D_REQUIRE(gz->C % 4 == 0 && vec4_ok(gz), "l3d_pw_bwd: gz must be 4-channel aligned");
    L3D_REQUIRE(gz->dtype == L3D_F32, "l3d_pw_bwd: gradient tensors are fp32");
    const bool has_gu = !act_null(g_u);
    if (has_gu) L3D_REQUIRE(g_u->C == u->C && g_u->dtype == L3D_F32 && (u->C % 4 != 0 || vec4_ok(g_u)), "l3d_pw_bwd: bad g_u view");
    PwBwdArgs A;
    memset(&A, 0, sizeof(A));
    A.gz = gz->ptr; A.ldg = gz->ldc;
    A.t = has_nt ? t->ptr : nullptr; A.ldt = has_nt ? t->ldc : 0;
    A.nt = norm_dev(nt); A.red = red;
    A.u = u->ptr; A.ldu = u->ldc; A.un = norm_dev(un);
    if (!vec4_ok(u)) A.ldu = u->ldc | 0;   // scalar path is selected in-kernel from Cu / ldu
    A.N = N; A.vox = (long long)D * H * W;
    A.Cg = gz->C; A.Cu = u->C;
    A.w = w; A.g_w = g_w; A.g_b = nullptr;
    A.g_u = has_gu ? g_u->ptr : nullptr; A.ldgu = has_gu ? g_u->ldc : 0; A.accumulate = accumulate_gu;
    // the vector path of the u loader needs an aligned base as well
    if (u->C % 4 == 0 && !vec4_ok(u)) { l3d_set_error("l3d_pw_bwd: u view must be 4-channel aligned"); return 1; }
    if (u->C == 1 && !has_gu && g_w != nullptr && (gz->C == 16 || gz->C == 32) && (!has_nt || vec4_ok(t))) {
        const unsigned gx = (unsigned)((148 * 8 + N - 1) / N);
        dim3 grid1(gx < 1 ? 1 : gx, (unsigned)N);
        L3D_DISPATCH_DTYPE(u->dtype, T, {
            if (gz->C == 16)
                pw_bwd_cu1_kernel<T, 16><<<grid1, 256, 0, (cudaStream_t)stream>>>((const float *)gz->ptr, gz->ldc, has_nt ? (const T *)t->ptr : nullptr, has_nt ? t->ldc : 0,
                                                                                    norm_dev(nt), red, (const T *)u->ptr, u->ldc, norm_dev(un), N, A.vox, g_w);
            else
                pw_bwd_cu1_kernel<T, 32><<<grid1, 256, 0, (cudaStream_t)stream>>>((const float *)gz->ptr, gz->ldc, has_nt ? (const T *)t->ptr : nullptr, has_nt ? t->ldc : 0,
                                                                                    norm_dev(nt), red, (const T *)u->ptr, u->ldc, norm_dev(un), N, A.vox, g_w);
        });
        l3d_count_launch();
        L3D_CUDA_OK("l3d_pw_bwd (single input channel) launch");
        return 0;
    }
    {   // tensor-core path (bf16 storage, 16-aligned channel counts)
        const int rc_tc = l3d_pw_bwd_tc(gz, t, nt, red, u, un, N, A.vox, w, g_w, g_u, accumulate_gu, stream);
        if (rc_tc == 0) { l3d_count_launch(); return 0; }
        if (rc_tc > 0) return rc_tc;
    }
    int rc = 0;
    L3D_DISPATCH_DTYPE(u->dtype, T, { rc = launch_pw_bwd<T, false>(A, (cudaStream_t)stream); });
    if (rc) return rc;
    l3d_count_launch();
    L3D_CUDA_OK("l3d_pw_bwd launch");
    return 0;
}

extern "C" int l3d_convt_bwd(const l3d_act *g_out, int OD, int OH, int OW, int oz, int oy, int ox,
                             const l3d_act *x, int N, int d, int h, int w_, const float *w,
                             float *g_w, float *g_b, const l3d_act *g_x, int accumulate_gx, void *stream) {
    L3D_REQUIRE(!act_null(g_out) && !act_null(x) && w, "l3d_convt_bwd: null argument");
    L3D_REQUIRE(g_out->C % 4 == 0 && vec4_ok(g_out) && g_out->dtype == L3D_F32, "l3d_convt_bwd: g_out must be fp32 and 4-channel aligned");
    L3D_REQUIRE(x->C % 4 != 0 || vec4_ok(x), "l3d_convt_bwd: x view must be 4-channel aligned");
    const bool has_gx = !act_null(g_x);
    if (has_gx) L3D_REQUIRE(g_x->C == x->C && g_x->dtype == L3D_F32 && (x->C % 4 != 0 || vec4_ok(g_x)), "l3d_convt_bwd: bad g_x view");
    PwBwdArgs A;
    memset(&A, 0, sizeof(A));
    A.gz = g_out->ptr; A.ldg = g_out->ldc;
    A.nt = norm_dev(nullptr); A.un = norm_dev(nullptr);
    A.u = x->ptr; A.ldu = x->ldc;
    A.N = N; A.vox = (long long)d * h * w_;
    A.Cg = g_out->C; A.Cu = x->C;
    A.w = w; A.g_w = g_w; A.g_b = g_b;
    A.g_u = has_gx ? g_x->ptr : nullptr; A.ldgu = has_gx ? g_x->ldc : 0; A.accumulate = accumulate_gx;
    A.d = d; A.h = h; A.w_ = w_; A.OD = OD; A.OH = OH; A.OW = OW; A.oz = oz; A.oy = oy; A.ox = ox;
    {   // tensor-core path (bf16 storage, 16-aligned channel counts)
        const int rc_tc = l3d_convt_bwd_tc(g_out, OD, OH, OW, oz, oy, ox, x, N, d, h, w_, w, g_w, g_b, g_x, accumulate_gx, stream);
        if (rc_tc == 0) { l3d_count_launch(); return 0; }
        if (rc_tc > 0) return rc_tc;
    }
    int rc = 0;
    // launch 1: input gradient (all taps) + bias gradient; launch 2: weight gradient, one tap per blockIdx.y with the
    // slice accumulated in registers over all tiles of the CTA (one atomic per output per CTA)
    float *gw_saved = A.g_w;
    if (has_gx || g_b != nullptr) {
        A.g_w = nullptr; A.tap_split = 0;
        L3D_DISPATCH_DTYPE(x->dtype, T, { rc = launch_pw_bwd<T, true>(A, (cudaStream_t)stream); });
        if (rc) return rc;
        l3d_count_launch();
    }
    if (gw_saved != nullptr) {
        A.g_w = gw_saved; A.g_b = nullptr; A.g_u = nullptr; A.tap_split = 1;
        L3D_DISPATCH_DTYPE(x->dtype, T, { rc = launch_pw_bwd<T, true>(A, (cudaStream_t)stream); });
        if (rc) return rc;
        l3d_count_launch();
    }
    L3D_CUDA_OK("l3d_convt_bwd launch");
    return 0;
}

extern "C" int l3d_dw_bwd(const l3d_act *g_u, const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                          const float *dw_w, float *g_dw_w, const l3d_act *gy, int accumulate_gy, double *redx,
                          void *stream) {
    L3D_REQUIRE(!act_null(g_u) && !act_null(x) && dw_w, "l3d_dw_bwd: null argument");
    const int C = x->C;
    L3D_REQUIRE(g_u->C == C && g_u->dtype == L3D_F32, "l3d_dw_bwd: g_u / x mismatch (gradients are fp32)");
    const bool has_gy = !act_null(gy);
    if (has_gy) L3D_REQUIRE(gy->C == C && gy->dtype == L3D_F32 && (C % 4 != 0 || vec4_ok(gy)), "l3d_dw_bwd: bad gy view");
    const bool has_norm = xn != nullptr && xn->stats != nullptr;
    if (has_norm && has_gy) L3D_REQUIRE(redx != nullptr, "l3d_dw_bwd: normalised producer needs redx");
    if (C == 1 && !has_gy && g_dw_w != nullptr) {
        const NormDev nd1 = norm_dev(xn);
        const long long nv1 = (long long)N * D * H * W;
        L3D_REQUIRE(nv1 < (1ll << 31), "l3d_dw_bwd (single channel): too many voxels for 32-bit indices");
        const unsigned grid1 = (unsigned)(nv1 / 256 + 1 < 148 * 8 ? nv1 / 256 + 1 : 148 * 8);
        L3D_DISPATCH_DTYPE(x->dtype, T, {
            dw_bwd_c1_wgrad_kernel<T><<<grid1, 256, 0, (cudaStream_t)stream>>>((const float *)g_u->ptr, g_u->ldc, (const T *)x->ptr, x->ldc, nd1, N, D, H, W, g_dw_w);
        });
        l3d_count_launch();
        L3D_CUDA_OK("l3d_dw_bwd (single channel) launch");
        return 0;
    }
    const size_t smem = dw_bwd_smem(C);
    L3D_REQUIRE(smem <= 227 * 1024, "l3d_dw_bwd: C=%d needs %zu B shared memory", C, smem);
    const long long tiles = (long long)N * ((D + TZ - 1) / TZ) * ((H + TY - 1) / TY) * ((W + TX - 1) / TX) * ((C + CK - 1) / CK);
    L3D_REQUIRE(tiles < (1ll << 31), "l3d_dw_bwd: too many work items for 32-bit indices");
    long long grid = tiles < 148 * 2 ? tiles : 148 * 2;
    const NormDev nd = norm_dev(xn);
    cudaStream_t st = (cudaStream_t)stream;
    L3D_DISPATCH_DTYPE(x->dtype, T, {
        auto kern = dw_bwd_kernel<T>;
        if (set_smem(kern, smem)) return 3;
        kern<<<(unsigned)grid, NT, smem, st>>>((const float *)g_u->ptr, g_u->ldc, (const T *)x->ptr, x->ldc, nd, C, N, D, H, W, dw_w, g_dw_w,
                                               has_gy ? (float *)gy->ptr : nullptr, has_gy ? gy->ldc : 0, accumulate_gy, redx);
    });
    l3d_count_launch();
    L3D_CUDA_OK("l3d_dw_bwd launch");
    return 0;
}

extern "C" int l3d_norm_param_grad(const double *red, int N, int C, float *g_gamma, float *g_beta, void *stream) {
    L3D_REQUIRE(red && g_gamma && g_beta && N > 0 && C > 0, "l3d_norm_param_grad: bad argument");
    norm_param_grad_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(red, N, C, g_gamma, g_beta);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_norm_param_grad launch");
    return 0;
}

extern "C" int l3d_norm_param_grad_batch(int count, const double *const *red, const int *C, float *const *g_gamma,
                                         float *const *g_beta, int N, void *stream) {
    L3D_REQUIRE(count >= 0 && red && C && g_gamma && g_beta && N > 0, "l3d_norm_param_grad_batch: bad argument");
    for (int i0 = 0; i0 < count; i0 += NPG_MAX) {
        NormGradBatch B;
        const int nb = count - i0 < NPG_MAX ? count - i0 : NPG_MAX;
        for (int i = 0; i < nb; ++i) {
            L3D_REQUIRE(red[i0 + i] && g_gamma[i0 + i] && g_beta[i0 + i] && C[i0 + i] > 0, "l3d_norm_param_grad_batch: null entry %d", i0 + i);
            B.red[i] = red[i0 + i]; B.g_gamma[i] = g_gamma[i0 + i]; B.g_beta[i] = g_beta[i0 + i]; B.C[i] = C[i0 + i];
        }
        norm_param_grad_batch_kernel<<<nb, 128, 0, (cudaStream_t)stream>>>(B, N);
        l3d_count_launch();
    }
    L3D_CUDA_OK("l3d_norm_param_grad_batch launch");
    return 0;
}

// ===========================================================================================
// Dense / grouped 3x3x3 backward on the tensor cores, tap by tap (unet3d.py:30,49,60):
//   forward   out[v][co] = sum_t sum_ci a[v + s_t][ci] * w[co][ci][t]
//   wgrad     g_w[co][ci][t] = sum_v g_t[v][co] * a[v + s_t][ci]        27 voxel-reduction GEMMs  [Cout x Cin]
//   dgrad     g_a[v + s_t][ci] += sum_co g_t[v][co] * w[co][ci][t]      27 pointwise GEMMs, accumulated at shifted addresses
// On zero-padded copies of g_t and a (one voxel of halo on every side, flat [N * (D+2)(H+2)(W+2)] voxel rows) a tap shift
// s_t is a constant pointer offset, so every tap is exactly the pointwise backward problem: both GEMMs of a tap run in
// ONE launch of pw_bwd_tc_kernel (l3d_bwd_tc.cu: bf16 hi + lo operand pairs on tcgen05, fp32 accumulation in TMEM) from
// the same staged tiles.  Halo rows of g_t are zero, so they add nothing to either product.
template <typename T>
__global__ void __launch_bounds__(NT) c3_pad_kernel(const float *__restrict__ gz, int ldg, const T *__restrict__ t, int ldt, NormDev nt,
                                                    const double *__restrict__ red, const T *__restrict__ x, int ldx, NormDev xn,
                                                    int N, int Cout, int Cin, int D, int H, int W,
                                                    float *__restrict__ gpad, float *__restrict__ apad) {
    // interior voxels only (the buffers were zeroed): g_t = a*gz + b*t + d and the activated conv input
    const size_t vox = (size_t)D * H * W;
    const int C = Cout + Cin;
    const size_t total = (size_t)N * vox * C;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int c = (int)(i % C);
        const size_t v = i / C;
        const int n = (int)(v / vox);
        size_t r = v - (size_t)n * vox;
        const int xw = (int)(r % W); r /= W;
        const int yh = (int)(r % H);
        const int zd = (int)(r / H);
        const size_t pv = (((size_t)n * (D + 2) + zd + 1) * (H + 2) + yh + 1) * (W + 2) + xw + 1;
        if (c < Cout) {
            double a, b, d;
            in_bwd_coef_d(nt, red, N, Cout, n, c, a, b, d);
            const float tv = nt.stats != nullptr ? ld1(t + v * ldt + c) : 0.f;
            gpad[pv * Cout + c] = in_bwd_apply(a, b, d, gz[v * ldg + c], tv);
        } else {
            const int ci = c - Cout;
            float sc, sh;
            norm_scale_shift(xn, N, Cin, n, ci, sc, sh);
            apad[pv * Cin + ci] = lrelu(ld1(x + v * ldx + ci) * sc + sh, xn.slope);
        }
    }
}
// wtap[t][co][ci] = w[co][ci - g*cin_g][t] inside the group of co, 0 across groups
__global__ void c3_tap_weights_kernel(const float *__restrict__ w, int Cin, int Cout, int groups, float *__restrict__ wtap) {
    const int cin_g = Cin / groups, cout_g = Cout / groups;
    const int total = 27 * Cout * Cin;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ci = i % Cin, co = (i / Cin) % Cout, tap = i / (Cin * Cout);
        const int g = co / cout_g, cil = ci - g * cin_g;
        wtap[i] = (cil >= 0 && cil < cin_g) ? w[((size_t)co * cin_g + cil) * 27 + tap] : 0.f;
    }
}
__global__ void c3_gw_scatter_kernel(const float *__restrict__ gwtap, int Cin, int Cout, int groups, float *__restrict__ g_w) {
    const int cin_g = Cin / groups, cout_g = Cout / groups;
    const int total = Cout * cin_g * 27;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tap = i % 27, cil = (i / 27) % cin_g, co = i / (27 * cin_g);
        g_w[i] += gwtap[((size_t)tap * Cout + co) * Cin + (co / cout_g) * cin_g + cil];
    }
}
__global__ void __launch_bounds__(NT) c3_unpad_kernel(const float *__restrict__ gapad, int N, int C, int D, int H, int W, float *__restrict__ ga) {
    const size_t vox = (size_t)D * H * W, total = (size_t)N * vox * (C / 4);
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int q = (int)(i % (C / 4));
        const size_t v = i / (C / 4);
        const int n = (int)(v / vox);
        size_t r = v - (size_t)n * vox;
        const int xw = (int)(r % W); r /= W;
        const int yh = (int)(r % H);
        const int zd = (int)(r / H);
        const size_t pv = (((size_t)n * (D + 2) + zd + 1) * (H + 2) + yh + 1) * (W + 2) + xw + 1;
        reinterpret_cast<float4 *>(ga + v * C)[q] = reinterpret_cast<const float4 *>(gapad + pv * C)[q];
    }
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static size_t c3_guard_vox(int H, int W) { return ((size_t)(H + 2) * (W + 2) + (W + 2) + 1 + 127) / 128 * 128; }
// Narrow layers (the 16 / 32-channel 48^3 level) keep the CUDA-core kernels: an [M = 128 x N = Cin] MMA tile is mostly
// empty at Cin = 16 and the 27 tap launches re-read the padded tensors 27 times (measured on the dense variant, batch 8 of
// 48^3, fp32 storage: every layer on the tap path 41.6 ms per step, none 38.4 ms; L3D_C3_BWD_TC_MIN moves the threshold).
static bool c3_bwd_tc_ok(int Cin, int Cout) {
    return L3D_ENV_INT("L3D_C3_BWD_TC", 1) != 0 && Cin % 16 == 0 && Cout % 16 == 0 && Cout <= 128 && Cin <= 256 &&
           Cin * Cout >= L3D_ENV_INT("L3D_C3_BWD_TC_MIN", 1024);
}
static size_t c3_bwd_tc_bytes(int N, int D, int H, int W, int Cin, int Cout) {
    const size_t vp = (size_t)N * (D + 2) * (H + 2) * (W + 2) + 2 * c3_guard_vox(H, W);
    return align256(vp * Cout * 4) + 2 * align256(vp * Cin * 4) + 2 * align256((size_t)27 * Cin * Cout * 4);
}
extern "C" int64_t l3d_conv3_bwd_workspace_bytes(int N, int D, int H, int W, int Cin, int Cout, int elem_size) {
    const size_t vox = (size_t)N * D * H * W;
    const size_t maxc = (size_t)(Cin > Cout ? Cin : Cout);
    return (int64_t)(align256(vox * Cout * elem_size) + align256(vox * Cin * elem_size) + align256((size_t)Cin * Cout * 27 * 4) +
                     align256(2 * (size_t)N * maxc * 8) + align256((size_t)((Cout + 7) / 8) * ((Cin + 7) / 8) * 8) +
                     (c3_bwd_tc_ok(Cin, Cout) ? c3_bwd_tc_bytes(N, D, H, W, Cin, Cout) : 0));
}

extern "C" int l3d_conv3_fwd(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                             const float *w, int groups, const l3d_act *t, double *t_stats,
                             const float *sc_w, const l3d_act *r, double *r_stats, void *stream);

extern "C" int l3d_conv3_bwd(const l3d_act *gz, const l3d_act *t, const l3d_norm *nt, const double *red,
                             const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                             const float *w, int groups, float *g_w,
                             const l3d_act *gy, int accumulate_gy, double *redx, void *work, int64_t work_bytes, void *stream) {
    L3D_REQUIRE(!act_null(gz) && !act_null(x) && w && g_w && work, "l3d_conv3_bwd: null argument");
    const int Cin = x->C, Cout = gz->C;
    const bool has_nt = nt != nullptr && nt->stats != nullptr;
    if (has_nt) L3D_REQUIRE(!act_null(t) && red && t->C == Cout, "l3d_conv3_bwd: bad t / red");
    L3D_REQUIRE(groups >= 1 && Cin % groups == 0 && Cout % groups == 0, "l3d_conv3_bwd: bad groups");
    L3D_REQUIRE(gz->dtype == L3D_F32, "l3d_conv3_bwd: gradient tensors are fp32");
    const size_t es = 4;
    L3D_REQUIRE(work_bytes >= l3d_conv3_bwd_workspace_bytes(N, D, H, W, Cin, Cout, (int)es), "l3d_conv3_bwd: workspace too small");
    const bool has_gy = !act_null(gy);
    const size_t vox = (size_t)D * H * W;
    char *wp = (char *)work;
    void *gt = wp; wp += align256((size_t)N * vox * Cout * es);
    void *ga = wp; wp += align256((size_t)N * vox * Cin * es);
    float *wT = (float *)wp; wp += align256((size_t)Cin * Cout * 27 * 4);
    double *dummy_stats = (double *)wp; wp += align256(2 * (size_t)N * (Cin > Cout ? Cin : Cout) * 8);
    int2 *pairs = (int2 *)wp; wp += align256((size_t)((Cout + 7) / 8) * ((Cin + 7) / 8) * 8);
    cudaStream_t st = (cudaStream_t)stream;
    const NormDev dnt = norm_dev(nt), dxn = norm_dev(xn);
    // ---- tensor-core path: 27 tap launches of the pointwise backward kernel over zero-padded copies (see above)
    if (c3_bwd_tc_ok(Cin, Cout) && (x->dtype == L3D_F32 || x->dtype == L3D_F16) && (!has_nt || t->dtype == x->dtype) &&
        (!has_gy || (gy->C == Cin && gy->dtype == L3D_F32)) && (long long)N * (D + 2) * (H + 2) * (W + 2) < (1ll << 31)) {
        const size_t guard = c3_guard_vox(H, W);
        const size_t vp = (size_t)N * (D + 2) * (H + 2) * (W + 2);
        float *gpad = (float *)wp; wp += align256((vp + 2 * guard) * Cout * 4);
        float *apad = (float *)wp; wp += align256((vp + 2 * guard) * Cin * 4);
        float *gapad = (float *)wp; wp += align256((vp + 2 * guard) * Cin * 4);
        float *wtap = (float *)wp; wp += align256((size_t)27 * Cin * Cout * 4);
        float *gwtap = (float *)wp;
        cudaMemsetAsync(gpad, 0, (vp + 2 * guard) * Cout * 4, st);
        cudaMemsetAsync(apad, 0, (vp + 2 * guard) * Cin * 4, st);
        if (has_gy) cudaMemsetAsync(gapad, 0, (vp + 2 * guard) * Cin * 4, st);
        cudaMemsetAsync(gwtap, 0, (size_t)27 * Cin * Cout * 4, st);
        {
            const size_t total = (size_t)N * vox * (Cin + Cout);
            size_t blocks = (total + NT - 1) / NT;
            if (blocks > 148 * 16) blocks = 148 * 16;
            L3D_DISPATCH_DTYPE(x->dtype, T, {
                c3_pad_kernel<T><<<(unsigned)blocks, NT, 0, st>>>((const float *)gz->ptr, gz->ldc, has_nt ? (const T *)t->ptr : nullptr, has_nt ? t->ldc : 0,
                                                                  dnt, red, (const T *)x->ptr, x->ldc, dxn, N, Cout, Cin, D, H, W,
                                                                  gpad + guard * Cout, apad + guard * Cin);
            });
            c3_tap_weights_kernel<<<(27 * Cin * Cout + 255) / 256, 256, 0, st>>>(w, Cin, Cout, groups, wtap);
            l3d_count_launch(2);
        }
        for (int tap = 0; tap < 27; ++tap) {
            const long long sh = ((long long)(tap / 9 - 1) * (H + 2) + (tap / 3 % 3 - 1)) * (W + 2) + (tap % 3 - 1);
            l3d_act a_g = {gpad + guard * Cout, Cout, Cout, L3D_F32, 0};
            l3d_act a_u = {apad + (long long)(guard + sh) * Cin, Cin, Cin, L3D_F32, 0};
            l3d_act a_gu = {has_gy ? gapad + (long long)(guard + sh) * Cin : nullptr, Cin, Cin, L3D_F32, 0};
            const int rc = l3d_pw_bwd_tc(&a_g, nullptr, nullptr, nullptr, &a_u, nullptr, 1, (long long)vp, wtap + (size_t)tap * Cout * Cin,
                                         gwtap + (size_t)tap * Cout * Cin, &a_gu, 1, stream);
            if (rc != 0) { if (rc < 0) l3d_set_error("l3d_conv3_bwd: tensor-core tap launch rejected (Cin=%d Cout=%d)", Cin, Cout); return rc < 0 ? 3 : rc; }
            l3d_count_launch();
        }
        c3_gw_scatter_kernel<<<(Cout * (Cin / groups) * 27 + 255) / 256, 256, 0, st>>>(gwtap, Cin, Cout, groups, g_w);
        l3d_count_launch();
        if (has_gy) {
            size_t blocks = ((size_t)N * vox * (Cin / 4) + NT - 1) / NT;
            if (blocks > 148 * 16) blocks = 148 * 16;
            c3_unpad_kernel<<<(unsigned)blocks, NT, 0, st>>>(gapad + guard * Cin, N, Cin, D, H, W, (float *)ga);
            size_t gx = (vox * Cin + NT - 1) / NT;
            const size_t cap = (148 * 16 + N - 1) / N;
            if (gx > cap) gx = cap;
            dim3 grid((unsigned)gx, (unsigned)N);
            L3D_DISPATCH_DTYPE(x->dtype, T, {
                c3_act_bwd_kernel<T><<<grid, NT, sizeof(float) * 9 * Cin, st>>>((const float *)ga, (const T *)x->ptr, x->ldc, dxn, N, Cin, vox,
                                                                                 (float *)gy->ptr, gy->ldc, accumulate_gy, redx);
            });
            l3d_count_launch(2);
        }
        l3d_note_kernel("pw_bwd_tc_kernel");
        L3D_CUDA_OK("l3d_conv3_bwd (tensor-core taps) launch");
        return 0;
    }
    // (1) g_t
    {
        const size_t total = (size_t)N * vox * Cout;
        size_t blocks = (total + NT - 1) / NT;
        if (blocks > 148 * 16) blocks = 148 * 16;
        L3D_DISPATCH_DTYPE(x->dtype, T, {
            c3_gt_kernel<T><<<(unsigned)blocks, NT, 0, st>>>((const float *)gz->ptr, gz->ldc, has_nt ? (const T *)t->ptr : nullptr,
                                                             has_nt ? t->ldc : 0, dnt, red, N, Cout, vox, (float *)gt);
        });
        l3d_count_launch();
    }
    // (2) wgrad over the (co-block, ci-block) pairs that share a group
    {
        const int cin_g = Cin / groups, cout_g = Cout / groups;
        int2 host_pairs[1024];
        int np = 0;
        for (int cb = 0; cb < Cout; cb += WG_C) {
            const int g_lo = cb / cout_g, g_hi = ((cb + WG_C < Cout ? cb + WG_C : Cout) - 1) / cout_g;
            const int lo = (g_lo * cin_g) / WG_C * WG_C, hi = (g_hi + 1) * cin_g;
            for (int c0 = lo; c0 < hi; c0 += WG_C) {
                L3D_REQUIRE(np < 1024, "l3d_conv3_bwd: too many channel-block pairs");
                host_pairs[np++] = make_int2(cb, c0);
            }
        }
        L3D_REQUIRE((size_t)np * 8 <= align256((size_t)((Cout + 7) / 8) * ((Cin + 7) / 8) * 8), "l3d_conv3_bwd: pair table overflow");
        cudaError_t e = cudaMemcpyAsync(pairs, host_pairs, sizeof(int2) * np, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { l3d_set_error("l3d_conv3_bwd: pair table copy: %s", cudaGetErrorString(e)); return 2; }
        cudaStreamSynchronize(st);   // host_pairs lives on this stack frame
        const long long tiles = (long long)N * ((D + TZ - 1) / TZ) * ((H + TY - 1) / TY) * ((W + TX - 1) / TX);
        long long splits = (148 * 4 + np - 1) / np;
        if (splits > tiles) splits = tiles;
        if (splits < 1) splits = 1;
        L3D_DISPATCH_DTYPE(x->dtype, T, {
            c3_wgrad_kernel<T><<<(unsigned)(np * splits), NT, 0, st>>>((const float *)gt, (const T *)x->ptr, x->ldc, dxn, N, Cin, Cout, groups,
                                                                        D, H, W, pairs, np, (int)splits, g_w);
        });
        l3d_count_launch();
    }
    // (3) dgrad = forward 3x3x3 conv of g_t with the flipped, transposed weights, then the producer's activation
    if (has_gy) {
        L3D_REQUIRE(gy->C == Cin && gy->dtype == L3D_F32, "l3d_conv3_bwd: bad gy view");
        L3D_REQUIRE(Cin % 8 == 0, "l3d_conv3_bwd: input-gradient path needs Cin %% 8 == 0 (got %d)", Cin);
        const int total = Cin * (Cout / groups) * 27;
        c3_flip_w_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, Cin, Cout, groups, wT);
        cudaMemsetAsync(dummy_stats, 0, 2 * (size_t)N * Cin * 8, st);
        l3d_count_launch();
        l3d_act a_gt = {gt, Cout, Cout, L3D_F32, 0};
        l3d_act a_ga = {ga, Cin, Cin, L3D_F32, 0};
        int rc = l3d_conv3_fwd(&a_gt, nullptr, N, D, H, W, wT, groups, &a_ga, dummy_stats, nullptr, nullptr, nullptr, stream);
        if (rc) return rc;
        size_t gx = (vox * Cin + NT - 1) / NT;
        const size_t cap = (148 * 16 + N - 1) / N;
        if (gx > cap) gx = cap;
        dim3 grid((unsigned)gx, (unsigned)N);
        L3D_DISPATCH_DTYPE(x->dtype, T, {
            c3_act_bwd_kernel<T><<<grid, NT, sizeof(float) * 9 * Cin, st>>>((const float *)ga, (const T *)x->ptr, x->ldc, dxn, N, Cin, vox,
                                                                             (float *)gy->ptr, gy->ldc, accumulate_gy, redx);
        });
        l3d_count_launch();
    }
    l3d_note_kernel("c3_wgrad_kernel");
    L3D_CUDA_OK("l3d_conv3_bwd launch");
    return 0;
}
