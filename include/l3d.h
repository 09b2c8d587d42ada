/* l3d.h -- C-ABI of the B200-native Light-3D-Unet hot path (libl3d.so).
 *
 * The reference (xxxxxxyp/Light-3D-Unet-Front) is pure Python/PyTorch and has no
 * FFI layer; its boundary for this path is the Python API of light_unet.models,
 * light_unet.utils and light_unet.core.inferencer.  The
 * drop-in Python package keeps that API and drives these entry points; each one
 * replaces the stock ATen/cuDNN/NumPy/SciPy call sequence issued by the cited
 * reference lines.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - activations are channels-last NDHWC ("voxel rows"): element (n,z,y,x,c) of a
 *    view lives at ptr[(((n*D+z)*H+y)*W+x)*ldc + c], ldc >= C (ldc > C lets a
 *    producer write straight into one half of a skip-concat buffer);
 *  - parameters and their gradients use the reference's own (PyTorch) layouts in
 *    fp32, so state_dict tensors are passed without repacking;
 *  - functions are asynchronous on `stream` (a cudaStream_t passed as void*),
 *    allocate nothing, and return 0 on success or a non-zero code whose text is
 *    available from l3d_last_error();
 *  - InstanceNorm statistics travel as double[2][N][C] = {sum, sum of squares} of
 *    the raw (pre-norm) tensor, accumulated by the producer's epilogue and
 *    turned into scale/shift by each consumer's prologue (no standalone norm pass).
 */
#ifndef L3D_H_
#define L3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L3D_ABI_VERSION 5

/* activation storage: fp32, or IEEE fp16 (stores saturate to +-65504); accumulation and statistics are fp32 / double */
enum { L3D_F32 = 0, L3D_F16 = 1 };

/* A channels-last activation view (spatial dims are passed per call). */
typedef struct {
    void   *ptr;
    int32_t C;      /* channels in this view */
    int32_t ldc;    /* elements between consecutive voxels */
    int32_t dtype;  /* L3D_F32 | L3D_F16 */
    int32_t pad_;
} l3d_act;

/* How a consumer must read a raw tensor: y = lrelu((x-mean)*rstd*gamma+beta)*drop.
 * stats == NULL means "already activated" (identity).  Mirrors
 * InstanceNorm3d(affine) -> LeakyReLU(0.01) -> Dropout3d of unet3d.py:80-85. */
typedef struct {
    const double *stats;  /* [2][N][C] sum, sumsq; NULL => identity */
    const float  *gamma;  /* [C] */
    const float  *beta;   /* [C] */
    const float  *drop;   /* [N][C] keep-mask already scaled by 1/(1-p), or NULL */
    float   eps;          /* 1e-5 */
    float   slope;        /* LeakyReLU negative slope; 1.0f => no activation */
    int32_t count;        /* voxels per (n,c) the stats were reduced over */
    int32_t pad_;
} l3d_norm;

const char *l3d_last_error(void);
int l3d_abi_version(void);
/* number of kernel launches issued through this library since load (for bench.py's gpu_launches) */
int64_t l3d_launch_count(void);
/* a CUDA graph captured through this library replays its kernels without passing through the entry points: the host adds
 * the number of launches the capture recorded, once per replay */
void l3d_add_launch_count(int64_t n);

/* ---------------------------------------------------------------- forward -- */

/* Fused depthwise 3x3x3 (pad 1) -> pointwise 1x1x1, optionally with the block's
 * 1x1x1 shortcut conv computed from the same input tile.
 *   replaces DepthwiseSeparableConv3d.forward (unet3d.py:20-23) + the shortcut
 *   conv (unet3d.py:70-71) + the stat pass of the InstanceNorm3d that follows
 *   (unet3d.py:51,62,72), with the preceding norm/LeakyReLU/Dropout3d applied on
 *   load (xn).
 * dw_w: [Cin][1][3][3][3] or NULL (pointwise only); pw_w: [Cout][Cin];
 * sc_w: [Cout][Cin] or NULL.  t/r: raw outputs (Cout channels) with their stats
 * buffers (must be zeroed by the caller); r may be {NULL}.  u: optional save of the
 * depthwise output (Cin channels) for the backward pass, may be {NULL}. */
int l3d_dwpw_fwd(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                 const float *dw_w, const float *pw_w, const float *sc_w,
                 const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats,
                 const l3d_act *u, void *stream);

/* Depthwise stage alone of the first conv of a single-channel image (DepthwiseSeparableConv3d with in_channels = 1,
 * unet3d.py:168,209): u[v] = (dw * act(x))[v] as one fp32 channel [N][D][H][W], plus the analytic InstanceNorm statistics
 * of t[v][c] = pw_w[c] * u[v] (the conv's output, unet3d.py:22-23) and, when sc_w is given, of the block's shortcut
 * r[v][c] = sc_w[c] * x[v] (unet3d.py:70-71) -- neither t nor r is written.  Inference only.  The stats buffers must be
 * zeroed by the caller. */
int l3d_dw_c1_fwd(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                  const float *dw_w, const float *pw_w, const float *sc_w, int Cout,
                  float *u, double *t_stats, double *r_stats, void *stream);

/* l3d_dwpw_fwd on the channel concatenation [x_lo | x_hi] of two DENSE 16-channel fp16 tensors (inference): the decoder
 * block of the top level reads the ConvTranspose3d output and the encoder skip as two tensors, so torch.cat
 * (unet3d.py:140) is not materialised AND both producers write whole cache lines (an interleaved [up | skip] buffer is
 * written as 32-byte halves of 64-byte voxels: measured 368 vs 274 us for the transposed conv, 549 vs 476 us for the
 * encoder merge at 325 windows of 48^3).  Implicit-GEMM kernel only; returns an error when it does not take the shape. */
int l3d_dwpw_fwd2(const l3d_act *x_lo, const l3d_act *x_hi, const l3d_norm *xn, int N, int D, int H, int W,
                  const float *dw_w, const float *pw_w, const float *sc_w,
                  const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats, void *stream);

/* l3d_dwpw_fwd on an input that is never materialised: x[v][c] = r1_w[c] * u[v] (Cin = 16 channels) with u from
 * l3d_dw_c1_fwd and xn the norm of that rank-1 tensor (its statistics come from l3d_dw_c1_fwd).  Replaces the second
 * DepthwiseSeparableConv3d of the first block (unet3d.py:55-63,82-86) without the 16-channel intermediate ever touching
 * HBM.  bf16 output, W % 4 == 0; returns an error otherwise (the caller then uses l3d_dwpw_fwd twice). */
int l3d_dwpw_fwd_rank1(const float *u, const float *r1_w, int Cin, const l3d_norm *xn, int N, int D, int H, int W,
                       const float *dw_w, const float *pw_w, const l3d_act *t, double *t_stats, void *stream);

/* Dense / grouped 3x3x3 convolution, pad 1, no bias (nn.Conv3d at unet3d.py:30,49,60).
 * w: [Cout][Cin/groups][3][3][3].  Optionally also the block's 1x1x1 shortcut conv on the same activated input
 * (unet3d.py:70-71): sc_w [Cout][Cin] -> r, r_stats (all three NULL when the block has no shortcut conv). */
int l3d_conv3_fwd(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                  const float *w, int groups, const l3d_act *t, double *t_stats,
                  const float *sc_w, const l3d_act *r, double *r_stats, void *stream);

/* Residual merge: out = lrelu(IN2(t2) + (INs(r) | r)) (unet3d.py:87-91), optionally also
 * emitting MaxPool3d(2,2) of out (unet3d.py:109) and/or the 1x1x1 head + sigmoid
 * (unet3d.py:220-221).  out / pooled may be {NULL}.  head_w: [OC][C] or NULL; prob and
 * logits are fp32 NCDHW [N][OC][D][H][W]; logits may be NULL. */
int l3d_merge_fwd(const l3d_act *t2, const l3d_norm *n2, const l3d_act *r, const l3d_norm *nr,
                  int N, int D, int H, int W, float slope,
                  const l3d_act *out, const l3d_act *pooled,
                  const float *head_w, const float *head_b, int OC, float *prob, float *logits,
                  void *stream);

/* ConvTranspose3d(Cin, Cout, k=2, s=2) + bias (unet3d.py:119,127), written at spatial
 * offset (oz,oy,ox) of an [N][OD][OH][OW] channels-last buffer (the centre pad of
 * unet3d.py:130-138; the caller zeroes the buffer when the offsets/sizes leave a rim).
 * w: [Cin][Cout][2][2][2]. */
int l3d_convt_fwd(const l3d_act *x, int N, int d, int h, int w_, const float *w, const float *b,
                  const l3d_act *out, int OD, int OH, int OW, int oz, int oy, int ox, void *stream);

/* --------------------------------------------------------------- backward -- */

/* The same with a shortcut that is never materialised: r[v][c] = r1_w[c] * x1[v] for a single-channel tensor x1 (the
 * first block's 1x1x1 shortcut conv of a 1-channel image, unet3d.py:70-71,168); nr holds r's statistics, which
 * l3d_dwpw_fwd produces even when its r output is {NULL}.  Inference only (the backward pass reads a stored r). */
int l3d_merge_fwd_rank1(const l3d_act *t2, const l3d_norm *n2, const l3d_act *x1, const float *r1_w, const l3d_norm *nr,
                        int N, int D, int H, int W, float slope, const l3d_act *out, const l3d_act *pooled, void *stream);

/* Backward of l3d_merge_fwd.  gz = g_out * lrelu'(out) (+ head gradient when head_w given:
 * g_out += g_logit * head_w; g_logit = g_prob * p * (1-p)).  Emits
 *   gz (C channels, raw gradient w.r.t. the pre-activation sum),
 *   red2/redr: double[2][N][C] = {sum gz, sum gz*xhat} for norm2 / the shortcut norm
 *   (must be zeroed), and head parameter gradients (accumulated, fp32).
 * g_out may be {NULL} when the only consumer is the head.  pooled_g: gradient arriving
 * through the fused max-pool (routed to the arg-max voxel, first max in z,y,x scan order as
 * ATen does), may be {NULL}. */
int l3d_merge_bwd(const l3d_act *g_out, const l3d_act *pooled_g, const l3d_act *out, const l3d_act *pooled,
                  const l3d_act *t2, const l3d_norm *n2, const l3d_act *r, const l3d_norm *nr,
                  int N, int D, int H, int W, float slope,
                  const float *head_w, int OC, const float *g_prob, const float *prob,
                  float *g_head_w, float *g_head_b,
                  const l3d_act *gz, double *red2, double *redr, void *stream);

/* Backward of the pointwise stage: given gz (gradient w.r.t. the normalised output y of a
 * raw tensor t = u . W^T), applies the InstanceNorm backward on the fly
 *   g_t = gamma*rstd*(gz - mean(gz) - xhat*mean(gz*xhat))
 * and produces g_u = g_t . W (Cin channels) and accumulates g_W += g_t^T . u.
 * For the identity norm (nt.stats == NULL) g_t = gz.  u is the saved depthwise output
 * (or the activated block input for a shortcut / pointwise-only conv, read through un). */
int l3d_pw_bwd(const l3d_act *gz, const l3d_act *t, const l3d_norm *nt, const double *red,
               const l3d_act *u, const l3d_norm *un, int N, int D, int H, int W,
               const float *w, float *g_w, const l3d_act *g_u, int accumulate_gu, void *stream);

/* Backward of the depthwise stage: g_a = dw^T(g_u) (flipped 27-tap stencil), g_dw += sum g_u*a,
 * then through Dropout3d/LeakyReLU of the producer: gy = g_a * drop * lrelu'(y) written to
 * `gy` (raw gradient w.r.t. the producer's normalised output) together with its reduction
 * redx = {sum gy, sum gy*xhat} (zeroed by caller).  With xn.stats == NULL the input was a
 * materialised activation: gy = g_a is accumulated/written as the gradient of that tensor. */
int l3d_dw_bwd(const l3d_act *g_u, const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
               const float *dw_w, float *g_dw_w, const l3d_act *gy, int accumulate_gy, double *redx,
               void *stream);

/* Backward of l3d_conv3_fwd (dense / grouped): g_x (through the producer's activation, like
 * l3d_dw_bwd) and g_w.  work: device scratch of l3d_conv3_bwd_workspace_bytes(...) bytes (holds the
 * materialised g_t, the raw input gradient, the flipped weights and a channel-block table). */
int64_t l3d_conv3_bwd_workspace_bytes(int N, int D, int H, int W, int Cin, int Cout, int elem_size);
int l3d_conv3_bwd(const l3d_act *gz, const l3d_act *t, const l3d_norm *nt, const double *red,
                  const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                  const float *w, int groups, float *g_w,
                  const l3d_act *gy, int accumulate_gy, double *redx, void *work, int64_t work_bytes, void *stream);

/* Backward of l3d_convt_fwd: g_x = sum_taps g_out . W^T, g_w += x^T . g_out, g_b += sum g_out. */
int l3d_convt_bwd(const l3d_act *g_out, int OD, int OH, int OW, int oz, int oy, int ox,
                  const l3d_act *x, int N, int d, int h, int w_, const float *w,
                  float *g_w, float *g_b, const l3d_act *g_x, int accumulate_gx, void *stream);

/* InstanceNorm affine gradients from the reductions: g_gamma[c] += sum_n red[1][n][c],
 * g_beta[c] += sum_n red[0][n][c]. */
int l3d_norm_param_grad(const double *red, int N, int C, float *g_gamma, float *g_beta, void *stream);
/* The same for `count` norms in one launch; red / C / g_gamma / g_beta are HOST arrays of `count` device pointers / sizes. */
int l3d_norm_param_grad_batch(int count, const double *const *red, const int *C, float *const *g_gamma,
                              float *const *g_beta, int N, void *stream);

/* ------------------------------------------------------------------- loss -- */

/* Focal Tversky (losses.py:40-52): sums[0..2] += {sum p*t, sum p, sum t} (double, zeroed by caller). */
int l3d_ftl_sums(const float *pred, const float *target, int64_t n, double *sums, void *stream);
/* loss = (1-TI)^gamma from (possibly all-reduced) sums; coef[0..1] = dL/dp for t=0 and the slope in t
 * (dL/dp_i = coef[0] + coef[1]*t_i). */
int l3d_ftl_finish(const double *sums, float alpha, float beta, float gamma, float smooth,
                   float *loss, float *coef, void *stream);
/* grad[i] = g_loss[0] * (coef[0] + coef[1]*target[i]) */
int l3d_ftl_grad(const float *target, int64_t n, const float *coef, const float *g_loss, float *grad,
                 void *stream);

/* --------------------------------------------------------- sliding window -- */

/* Cut windows out of a [D][H][W] fp32 volume (zero-padded at the far end, utils.py:91-112) into a
 * [nwin][pd][ph][pw] single-channel batch of `dtype`.  pos: int32 [nwin][3] (z,y,x). */
int l3d_gather_windows(const float *vol, int D, int H, int W, const int32_t *pos, int nwin,
                       int pd, int ph, int pw, void *out, int dtype, void *stream);

/* Gaussian-weighted overlap stitching (utils.py:126-137) as a per-voxel gather over the window
 * grid, adding contributions in the reference's z->y->x window order with separately rounded
 * fp32 multiply and add, then prob/cnt.  preds: fp32 [nz*ny*nx][pd][ph][pw] in window order.
 * importance: fp32 [pd][ph][pw].  body_mask (uint8 [D][H][W]) or NULL multiplies the result
 * (inferencer.py:161-162).  mask_out (int32 [D][H][W]) or NULL receives prob >= threshold
 * (inferencer.py:64). */
int l3d_stitch(const float *preds, const int32_t *zpos, int nz, const int32_t *ypos, int ny,
               const int32_t *xpos, int nx, int pd, int ph, int pw, const float *importance,
               int D, int H, int W, const uint8_t *body_mask, float *prob,
               float threshold, int32_t *mask_out, void *stream);

/* Slab form of l3d_stitch for window-level sharding of ONE volume over several GPUs (the loop of
 * utils.py:86-134 split along the longest axis): stitches the voxels x in [x0, x1) only.  Window (a, b, c) of the
 * (z, y, x) grid lives at preds[(a*wstride_z + b*wstride_y + c*wstride_x) * pd*ph*pw] -- a rank keeps [x-position][z][y]
 * blocks, the seam positions received from its neighbours first -- and voxel (z, y, x) is written to
 * prob[(z*H + y) * out_pitch + out_xoff + (x - x0)] (same for mask_out).  xpos holds the nx x-positions present in
 * `preds`.  Candidate windows are added in the same z -> y -> x order, so the slab is bit-identical to the same voxels of
 * the single-GPU map given identical predictions.  body_mask is indexed in the full volume. */
int l3d_stitch_slab(const float *preds, const int32_t *zpos, int nz, const int32_t *ypos, int ny,
                    const int32_t *xpos, int nx, int64_t wstride_z, int64_t wstride_y, int64_t wstride_x,
                    int pd, int ph, int pw, const float *importance,
                    int D, int H, int W, int x0, int x1, int out_pitch, int out_xoff,
                    const uint8_t *body_mask, float *prob, float threshold, int32_t *mask_out, void *stream);

/* mask[i] = prob[i] >= threshold (inferencer.py:64). */
int l3d_threshold(const float *prob, int64_t n, float threshold, int32_t *mask, void *stream);

/* 6-connected component labelling with minimum-size filter and raster-order renumbering
 * (metrics.py:50-61; replaces scipy.ndimage.label x2 + bincount).  labels: int32 [D][H][W] out.
 * work: int32 scratch of l3d_ccl_workspace_elems(D*H*W) elements.  n_out: device int32. */
int64_t l3d_ccl_workspace_elems(int64_t nvox);
int l3d_ccl_label(const int32_t *mask, int D, int H, int W, int min_size, int32_t *labels,
                  int32_t *n_out, int32_t *work, void *stream);

/* Per-component reduction (inferencer.py:74-100): table[id-1] = {zmin,zmax,ymin,ymax,xmin,xmax,count,
 * bits(max prob)} for ids 1..cap (components beyond cap are ignored).  table must be initialised by
 * l3d_bbox_init. */
int l3d_bbox_init(int32_t *table, int cap, void *stream);
int l3d_bbox_reduce(const int32_t *labels, const float *prob, int D, int H, int W, int32_t *table, int cap,
                    void *stream);

/* Lesion-matching statistics of two label maps in one pass (metrics.py:127-229, :311-404; replaces np.bincount over
 * pred_id * (nb + 1) + target_id, the size bincounts and ndimage.center_of_mass): counts[a * (nb + 1) + b] += 1 for voxels
 * labelled a in A and b in B (both > 0); mom_x[id * 4 + {0, 1, 2, 3}] += {1, z, y, x} per labelled voxel.  counts
 * (int32 [(na + 1) * (nb + 1)]), mom_a (int64 [(na + 1) * 4]) and mom_b must be zeroed by the caller; labels_b / counts /
 * mom_b may be NULL (moments of one map only). */
int l3d_label_pair_stats(const int32_t *labels_a, const int32_t *labels_b, int D, int H, int W, int na, int nb,
                         int32_t *counts, int64_t *mom_a, int64_t *mom_b, void *stream);

/* ------------------------------------------------------ training patches -- */

/* Training-patch pipeline on the device (replaces PatchDataset.__getitem__'s disk reads + numpy / scipy work,
 * patch_dataset.py:114-220).  Patches are fp32 [B][pd][ph][pw]; per-sample parameters are small DEVICE tables.
 * l3d_patch_extract: img_ptrs / lab_ptrs = device arrays of B device pointers to fp32 [D][H][W] volumes resident in HBM,
 *   dims / start = int32 [B][3]; the window start..start+p is clipped at the far edge and zero-padded at the end (:136-154).
 * l3d_patch_flip: axis[b] in {0, 1, 2} or -1 (copy)                                                              (:160-165)
 * l3d_patch_rotate: scipy.ndimage.rotate(reshape=False, mode='constant'), order 1 for the image and 0 for the label;
 *   axes[b] = {a0 < a1} or {-1, -1} (copy), coef[b] = {m00, m01, m10, m11, off0, off1} with in = off + m . out          (:167-174)
 * l3d_patch_zoom: scipy.ndimage.zoom (order 1 / 0, mode='constant') + centre-crop / end-pad back to the patch size;
 *   geo[b] = {on, zoomed dims[3], crop start[3]}, zf[b] = (n - 1) / (zoomed - 1) per axis                             (:176-208)
 * l3d_patch_intensity: in place, image = clip(image + shift[b], 0, 1) in fp32 if shift_on[b], then
 *   image = float(clip(double(image) + noise[b][i], 0, 1)) if noise_on[b] (noise may be NULL)                         (:210-218)
 * Results are bit-identical to the reference's for the same parameters (tests/test_gpu_patches.py). */
int l3d_patch_extract(const void *const *img_ptrs, const void *const *lab_ptrs, const int32_t *dims, const int32_t *start,
                      int B, int pd, int ph, int pw, float *out_img, float *out_lab, void *stream);
int l3d_patch_flip(const float *in_img, const float *in_lab, float *out_img, float *out_lab, const int32_t *axis,
                   int B, int pd, int ph, int pw, void *stream);
int l3d_patch_rotate(const float *in_img, const float *in_lab, float *out_img, float *out_lab, const int32_t *axes,
                     const double *coef, int B, int pd, int ph, int pw, void *stream);
int l3d_patch_zoom(const float *in_img, const float *in_lab, float *out_img, float *out_lab, const int32_t *geo,
                   const double *zf, int B, int pd, int ph, int pw, void *stream);
int l3d_patch_intensity(float *img, const float *shift, const int32_t *shift_on, const double *noise, const int32_t *noise_on,
                        int B, int64_t per, void *stream);

/* ------------------------------------------------------------- diagnostics -- */

/* The library reads its tuning / test knobs (L3D_* environment variables) once per call site; call this after changing
 * one inside a running process. */
void l3d_env_refresh(void);

/* Name of the kernel the last dispatching entry point (l3d_dwpw_fwd, l3d_conv3_fwd, l3d_convt_fwd) launched on this
 * thread: "conv3_tc_kernel", "dwpw_tc_kernel", "dwpw_c1_kernel", ... ("" before the first launch). */
const char *l3d_last_kernel(void);

/* Self-test of the tcgen05 building blocks: D[128*MT][N] (fp32) = A[128*MT][K] . Wt[N][K]^T with fp16 operands
 * and fp32 accumulation in TMEM.  A, Wt, D are fp32 device arrays. */
int l3d_tc_selftest(const float *A, const float *Wt, int MT, int K, int N, float *D, void *stream);

/* Self-test of kind::tf32 with K-major operands: D[128][N] = G[128][K] . WU[N][K]^T (mode must be 0, M is ignored).  fp32
 * device arrays; operands are rounded to tf32 by the tensor core. */
int l3d_tc_selftest_tf32(const float *G, const float *WU, int mode, int M, int K, int N, float *D, void *stream);

/* Self-test of the 16-bit MN-major voxel reduction used by the tensor-core weight gradients:
 * D[m][n] = sum_{v<128} G[v][m] * U[v][n] with bf16 operands (rows >= M of D are don't-care). */
int l3d_tc_selftest_mn16(const float *G, const float *U, int M, int N, int var, float *D, void *stream);

/* Development aid: per-work-item clock64 stamps of CTA 0 of the last implicit-GEMM conv launched with
 * L3D_C3_DEBUG_SKIP & 8 ({worker: box landed, operand buffer free, operand written, accumulators ready, epilogue
 * done; issuer: operand ready, accumulators free, MMAs issued} x up to 128 items); n int64 values are copied. */
int l3d_conv3_debug_read(long long *host, int n);

#ifdef __cplusplus
}
#endif
#endif /* L3D_H_ */
