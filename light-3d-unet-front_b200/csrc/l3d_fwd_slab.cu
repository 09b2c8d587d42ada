// Depthwise-separable conv forward for SMALL volumes (bf16 storage, inference): the 12^3 / 6^3 levels of a 48^3 window,
// where a sample plane is only 36 .. 144 voxels and the 4x8x8-voxel tiles of dwpw_tc_kernel waste > 2x of their stencil
// work, halo traffic and MMA rows (unet3d.py:20-23, 70-72, 80-87).
//
// Work item = a SLAB: SZ whole z-planes of one sample (H x W voxels each).  Whole planes mean
//   * no x / y halo: the raw box of a 16-channel chunk is (SZ + 2) contiguous planes, one 5-D TMA box;
//   * the slab's output voxels are contiguous in memory and fill MT 128-row MMA tiles back to back;
//   * no padded tile voxels: every stencil output is a real voxel.
// Per (slab, 16-channel chunk): TMA raw box (double-buffered, requested one item ahead) -> activation pass (InstanceNorm +
// LeakyReLU + Dropout3d of the producer) into a zero-bordered fp32 tile -> depthwise 3x3x3 on CUDA cores, each thread
// 2 channels x 2 rows x XT x-consecutive voxels with 8-byte shared-memory loads -> fp16 K-major operand tile
// (double-buffered) -> tcgen05.mma for the pointwise (+ shortcut) stage, accumulating over the chunks in TMEM ->
// epilogue: tcgen05.ld, bf16 store, InstanceNorm statistics.
#include <cuda.h>

#include "l3d_common.cuh"
#include "l3d_tc.cuh"

namespace {

constexpr int CK = 16;
constexpr int ACT_SLOTS = 8;

struct SlabArgs {
    int Cin; NormDev xn;
    int N, D, H, W;
    const float *dw_w, *pw_w, *sc_w; int Cout;
    h16 *t; int ldt; double *t_stats;
    h16 *r; int ldr; double *r_stats;
    int SZ, MT, RP, PP;      // slab height, MMA tiles per slab, padded row / plane pitch of the stencil tile (voxels, odd)
    int tmem_cols;
    uint32_t raw_bytes, raw_stride, in_bytes;
    int wide_st;             // outputs are 32-byte aligned 16-channel blocks: 256-bit stores
    int dbg;                 // development aid (L3D_SLAB_SKIP): 1 = no activation pass, 2 = no stencil, 4 = no epilogue, 8 = no TMA
};

__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ void fma4(float4 &a, const float4 &w, const float4 &v) {
    a.x = fmaf(w.x, v.x, a.x); a.y = fmaf(w.y, v.y, a.y); a.z = fmaf(w.z, v.z, a.z); a.w = fmaf(w.w, v.w, a.w);
}

__device__ __forceinline__ void fma2(float2 &a, const float2 &w, const float2 &v) {
    a.x = fmaf(w.x, v.x, a.x); a.y = fmaf(w.y, v.y, a.y);
}

// NT threads: 256 (two CTAs per SM where shared memory allows) or 384 (slabs with more than 256 stencil tasks)
template <int XT, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1) dwpw_slab_kernel(const __grid_constant__ CUtensorMap tmap, SlabArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_tma_full[2], s_mma_done[2];
    __shared__ uint32_t s_tmem;
    const int Cin = A.Cin, Cout = A.Cout, MT = A.MT, SZ = A.SZ, H = A.H, W = A.W, RP = A.RP, PP = A.PP;
    const bool has_sc = A.sc_w != nullptr;
    const int nacc = has_sc ? 2 : 1;
    const uint32_t a_bytes = (uint32_t)MT * 128 * CK * 2;                 // one chunk operand tile [2 K groups][MT*128 rows][8]
    const uint32_t b_bytes = (uint32_t)Cout * Cin * 2;
    unsigned char *s_raw = smem_raw;                                       // 2 x raw_stride (TMA destinations)
    float *s_in = reinterpret_cast<float *>(s_raw + 2 * A.raw_stride);     // (SZ+2) x PP voxels x 16 fp32, zero borders
    unsigned char *sA = reinterpret_cast<unsigned char *>(s_in) + A.in_bytes;   // [buf][main | shortcut] x a_bytes
    unsigned char *sB = sA + 2 * nacc * a_bytes;
    unsigned char *sB2 = sB + b_bytes;
    float *s_dw = reinterpret_cast<float *>(sB2 + (has_sc ? b_bytes : 0));  // [2][27][16]
    float *s_scale = s_dw + 2 * 27 * CK;                                    // Cin
    float *s_shift = s_scale + Cin;
    float *s_stat = s_shift + Cin;                                          // 2*Cout (t) + 2*Cout (r)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc::tmem_alloc(&s_tmem, (uint32_t)A.tmem_cols);
    if (tid == 32) {
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&s_tma_full[i], 1); tc::mbar_init(&s_mma_done[i], 1); }
    }
    // pointwise (+ shortcut) weights once per CTA as fp16 K-major operand tiles
    {
        const int nvec = Cout * Cin / 4;
        for (int s = 0; s < nacc; ++s) {
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
                        uint2 o;
                        o.x = pack_f16x2(v[j].x, v[j].y);
                        o.y = pack_f16x2(v[j].z, v[j].w);
                        *reinterpret_cast<uint2 *>(dst + tc::tile_off(n, k, Cout)) = o;
                    }
                }
            }
        }
    }
    for (int i = tid; i < (int)(A.in_bytes / 16); i += NT) reinterpret_cast<float4 *>(s_in)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 4 * Cout; i += NT) s_stat[i] = 0.f;
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = tc::idesc_f16_m128(Cout);
    const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB), sB2_u = tc::smem_u32(sB2);

    const int zslabs = (A.D + SZ - 1) / SZ;
    const int total_slabs = zslabs * A.N;
    const int nchunks = Cin / CK;
    const int my_slabs = (int)blockIdx.x < total_slabs ? (total_slabs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int n_items = my_slabs * nchunks;
    const int HW = H * W;

    // ---- fixed per-thread roles
    // activation pass: 16-byte vectors (voxel, 8-channel half) of the raw box -> float offset of the voxel in s_in
    const int act_items = (SZ + 2) * HW * 2;
    uint32_t act_slot[ACT_SLOTS];
#pragma unroll
    for (int k = 0; k < ACT_SLOTS; ++k) {
        const int item = tid + k * NT;
        uint32_t v = 0xffffffffu;
        if (item < act_items) {
            int vx = item >> 1;
            const int q = item & 1;
            const int px = vx % W; vx /= W;
            const int py = vx % H;
            const int pz = vx / H;
            v = (uint32_t)((pz * PP + (py + 1) * RP + px + 1) * CK + q * 8) | ((uint32_t)pz << 24);
        }
        act_slot[k] = v;
    }
    // stencil: task = (xb, y pair, z, channel pair), z fastest: the two tasks of a half warp are an odd number of
    // voxels apart (PP is odd), so their 8-byte loads fall into different bank halves; a thread's tasks all share its
    // channel pair cp = tid & 7 (NT is a multiple of 8).  Two channels per task (not four): twice the tasks, so that a
    // 12^3 slab of 3 planes keeps 9 warps busy instead of 4.5 -- the stencil is latency-bound, not FMA-bound
    const int cp = tid & 7;
    const int xblocks = W / XT;
    const int ntasks = SZ * (H / 2) * xblocks * 8;
    // epilogue: TMEM lane quarter warp & 3; the two warp groups alternate over the (accumulator, tile, 16-column) jobs
    const int eq = warp & 3, eg = warp >> 2;
    constexpr int NG = NT / 128;

    // depthwise taps of the first chunk (register prefetch, one chunk ahead)
    float dwr0 = A.dw_w[tid], dwr1 = (tid + NT < CK * 27) ? A.dw_w[tid + NT] : 0.f;

    auto item_coord = [&](int it, int &n, int &z0, int &ch) {
        const int sl = it / nchunks;
        ch = it - sl * nchunks;
        const int slab = (int)blockIdx.x + sl * (int)gridDim.x;
        n = slab / zslabs;
        z0 = (slab - n * zslabs) * SZ;
    };
    auto issue_tma = [&](int it) {
        if (it >= n_items || (A.dbg & 8)) return;
        int n, z0, ch;
        item_coord(it, n, z0, ch);
        tc::mbar_expect_tx(&s_tma_full[it & 1], A.raw_bytes);
        tc::tma_load_5d(s_raw + (size_t)(it & 1) * A.raw_stride, &tmap, &s_tma_full[it & 1], ch * CK, 0, 0, z0 - 1, n);
    };
    if (tid == 0) issue_tma(0);

    // the first two stencil tasks of this thread, decoded once (integer divisions by run-time extents cost a few hundred
    // cycles each and sat between two CTA barriers of every work item): packed {vox, row}
    int task_vox[2], task_row[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        int b = (tid + k * NT) >> 3;
        const int z = b % SZ; b /= SZ;
        const int y = (b % (H / 2)) * 2;
        const int xb = b / (H / 2);
        task_vox[k] = z * PP + y * RP + xb * XT;
        task_row[k] = (z * H + y) * W + xb * XT;
    }
    int cur_n = 0, cur_z0 = 0, cur_ch = 0, cur_sl = 0;
    if (n_items > 0) item_coord(0, cur_n, cur_z0, cur_ch);
    for (int it = 0; it < n_items; ++it) {
        const int n = cur_n, z0 = cur_z0, ch = cur_ch;
        if (++cur_ch == nchunks) {          // coordinates of the next item: one division per slab, none per chunk
            cur_ch = 0; ++cur_sl;
            const int slab = (int)blockIdx.x + cur_sl * (int)gridDim.x;
            cur_n = slab / zslabs;
            cur_z0 = (slab - cur_n * zslabs) * SZ;
        }
        const int buf = it & 1;
        // request the next item's raw box: its buffer was consumed by the activation pass of item it-1
        if (tid == 0) issue_tma(it + 1);
        // depthwise taps of this chunk -> s_dw[buf] as [tap][16 channels]; prefetch the next chunk's
        {
            float *dwb = s_dw + buf * 27 * CK;
            { const int c = tid / 27, tap = tid - c * 27; dwb[tap * CK + c] = dwr0; }
            if (tid + NT < CK * 27) { const int e = tid + NT, c = e / 27, tap = e - c * 27; dwb[tap * CK + c] = dwr1; }
            const int nc0 = (ch + 1 < nchunks) ? (ch + 1) * CK : 0;
            dwr0 = A.dw_w[(size_t)nc0 * 27 + tid];
            if (tid + NT < CK * 27) dwr1 = A.dw_w[(size_t)nc0 * 27 + tid + NT];
        }
        if (ch == 0) {
            for (int cc = tid; cc < Cin; cc += NT) {
                float sc, sh;
                norm_scale_shift(A.xn, A.N, Cin, n, cc, sc, sh);
                s_scale[cc] = sc; s_shift[cc] = sh;
            }
            __syncthreads();
        }
        // ---- activation pass: raw bf16 box [pz][y][x][16] -> fp32 stencil tile (zero outside the volume: the conv pads
        // the ACTIVATED tensor; the x / y borders of the tile are never written and stay zero)
        if (!(A.dbg & 8)) tc::mbar_wait(&s_tma_full[buf], (uint32_t)((it >> 1) & 1));
        // operand tiles `buf` are free once the MMAs of item it-2 have completed
        if (it >= 2) tc::mbar_wait(&s_mma_done[buf], (uint32_t)(((it >> 1) - 1) & 1));
        if (!(A.dbg & 1)) {
            unsigned char *As = sA + (size_t)buf * nacc * a_bytes + a_bytes;
            const unsigned char *Rb = s_raw + (size_t)buf * A.raw_stride;
            const float sl = A.xn.slope;
            const int c0 = ch * CK;
#pragma unroll
            for (int k = 0; k < ACT_SLOTS; ++k) {
                const uint32_t as = act_slot[k];
                if (as != 0xffffffffu) {
                    const int pz = (int)(as >> 24), q = (int)((as >> 3) & 1);
                    const int gz = z0 + pz - 1;
                    float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
                    if (gz >= 0 && gz < A.D) {
                        const uint4 rw = *reinterpret_cast<const uint4 *>(Rb + (size_t)(tid + k * NT) * 16);
                        const float4 sc0 = *reinterpret_cast<const float4 *>(s_scale + c0 + q * 8);
                        const float4 sc1 = *reinterpret_cast<const float4 *>(s_scale + c0 + q * 8 + 4);
                        const float4 sh0 = *reinterpret_cast<const float4 *>(s_shift + c0 + q * 8);
                        const float4 sh1 = *reinterpret_cast<const float4 *>(s_shift + c0 + q * 8 + 4);
                        o0.x = lrelu(fmaf(h16_lo(rw.x), sc0.x, sh0.x), sl);
                        o0.y = lrelu(fmaf(h16_hi(rw.x), sc0.y, sh0.y), sl);
                        o0.z = lrelu(fmaf(h16_lo(rw.y), sc0.z, sh0.z), sl);
                        o0.w = lrelu(fmaf(h16_hi(rw.y), sc0.w, sh0.w), sl);
                        o1.x = lrelu(fmaf(h16_lo(rw.z), sc1.x, sh1.x), sl);
                        o1.y = lrelu(fmaf(h16_hi(rw.z), sc1.y, sh1.y), sl);
                        o1.z = lrelu(fmaf(h16_lo(rw.w), sc1.z, sh1.z), sl);
                        o1.w = lrelu(fmaf(h16_hi(rw.w), sc1.w, sh1.w), sl);
                    }
                    float *dst = s_in + (as & 0xffffffu);
                    *reinterpret_cast<float4 *>(dst) = o0;
                    *reinterpret_cast<float4 *>(dst + 4) = o1;
                    if (has_sc && pz >= 1 && pz <= SZ) {
                        // the shortcut conv's operand is the activated tensor itself: row = voxel of the slab, K group = q
                        const int rr = ((tid + k * NT) >> 1) - HW;
                        *reinterpret_cast<uint4 *>(As + (uint32_t)q * MT * 2048 + (uint32_t)(rr >> 3) * 128 + (uint32_t)(rr & 7) * 16) =
                            make_uint4(pack_f16x2(o0.x, o0.y), pack_f16x2(o0.z, o0.w), pack_f16x2(o1.x, o1.y), pack_f16x2(o1.z, o1.w));
                    }
                }
            }
        }
        __syncthreads();                 // stencil tile complete, raw box consumed
        // ---- depthwise stencil: 4 channels x 2 rows x XT voxels per task
        if (!(A.dbg & 2)) {
            const float2 *in2 = reinterpret_cast<const float2 *>(s_in);
            const float2 *w2p = reinterpret_cast<const float2 *>(s_dw + buf * 27 * CK) + cp;
            unsigned char *Am = sA + (size_t)buf * nacc * a_bytes;
            const uint32_t kg_off = (uint32_t)(cp >> 2) * MT * 2048 + (uint32_t)(cp & 3) * 4;
#pragma unroll 1
            int tk = 0;
            for (int tsk = tid; tsk < ntasks; tsk += NT, ++tk) {
                int vox, row;
                if (tk < 2) { vox = tk == 0 ? task_vox[0] : task_vox[1]; row = tk == 0 ? task_row[0] : task_row[1]; }
                else {
                    int b = tsk >> 3;
                    const int z = b % SZ; b /= SZ;
                    const int y = (b % (H / 2)) * 2;
                    const int xb = b / (H / 2);
                    vox = z * PP + y * RP + xb * XT; row = (z * H + y) * W + xb * XT;
                }
                // output rows y and y + 1: every input row is loaded once and feeds both
                float2 acc0[XT], acc1[XT];
#pragma unroll
                for (int i = 0; i < XT; ++i) { acc0[i] = make_float2(0.f, 0.f); acc1[i] = acc0[i]; }
#pragma unroll
                for (int dz = 0; dz < 3; ++dz) {
#pragma unroll
                    for (int hy = 0; hy < 4; ++hy) {
                        const float2 *rp = in2 + (size_t)(vox + dz * PP + hy * RP) * 8 + cp;
                        float2 v[XT + 2];
#pragma unroll
                        for (int x = 0; x < XT + 2; ++x) v[x] = rp[x * 8];
                        if (hy <= 2) {
                            const float2 w0 = w2p[((dz * 3 + hy) * 3 + 0) * 8], w1 = w2p[((dz * 3 + hy) * 3 + 1) * 8], w2 = w2p[((dz * 3 + hy) * 3 + 2) * 8];
#pragma unroll
                            for (int i = 0; i < XT; ++i) { fma2(acc0[i], w0, v[i]); fma2(acc0[i], w1, v[i + 1]); fma2(acc0[i], w2, v[i + 2]); }
                        }
                        if (hy >= 1) {
                            const float2 w0 = w2p[((dz * 3 + hy - 1) * 3 + 0) * 8], w1 = w2p[((dz * 3 + hy - 1) * 3 + 1) * 8], w2 = w2p[((dz * 3 + hy - 1) * 3 + 2) * 8];
#pragma unroll
                            for (int i = 0; i < XT; ++i) { fma2(acc1[i], w0, v[i]); fma2(acc1[i], w1, v[i + 1]); fma2(acc1[i], w2, v[i + 2]); }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < XT; ++i) {
                    const int r0 = row + i, r1 = row + W + i;
                    *reinterpret_cast<uint32_t *>(Am + kg_off + (uint32_t)(r0 >> 3) * 128 + (uint32_t)(r0 & 7) * 16) = pack_f16x2(acc0[i].x, acc0[i].y);
                    *reinterpret_cast<uint32_t *>(Am + kg_off + (uint32_t)(r1 >> 3) * 128 + (uint32_t)(r1 & 7) * 16) = pack_f16x2(acc1[i].x, acc1[i].y);
                }
            }
        }
        tc::fence_async_smem();
        __syncthreads();                 // operand chunk complete; s_in free for the next item
        if (tid == 0) {
            tc::fence_after_sync();
            const uint32_t accf = ch > 0 ? 1u : 0u;
            const uint32_t a_u = sA_u + (uint32_t)buf * nacc * a_bytes;
            const uint64_t bd = tc::smem_desc(sB_u + 2 * ch * Cout * 16, Cout * 16, 128);
            const uint64_t bd2 = tc::smem_desc(sB2_u + 2 * ch * Cout * 16, Cout * 16, 128);
            for (int m = 0; m < MT; ++m) {
                tc::mma_f16(tmem + m * Cout, tc::smem_desc(a_u + m * 2048, MT * 2048, 128), bd, idesc, accf);
                if (has_sc) tc::mma_f16(tmem + (MT + m) * Cout, tc::smem_desc(a_u + a_bytes + m * 2048, MT * 2048, 128), bd2, idesc, accf);
            }
            tc::mma_commit(&s_mma_done[buf]);
        }
        if (ch + 1 < nchunks) continue;
        // ---- epilogue of the slab: TMEM -> bf16 global + statistics
        tc::mbar_wait(&s_mma_done[buf], (uint32_t)((it >> 1) & 1));
        tc::fence_after_sync();
        if (!(A.dbg & 4)) {
            const int zv = min(SZ, A.D - z0);
            const int rows_valid = zv * HW;
            const size_t vox0 = ((size_t)n * A.D + z0) * HW;
            const uint32_t trow = tmem + ((uint32_t)(eq * 32) << 16);
            const int cbn = Cout >> 4;
            const int njobs = nacc * MT * cbn;
            const bool wide_st = A.wide_st != 0;
            // job j = (a * MT + m) * cbn + cbi, walked in steps of NG with carries (no divisions)
            int cbi = eg % cbn, m = (eg / cbn) % MT, a = eg / (cbn * MT);
            for (int j = eg; j < njobs; j += NG) {
                const int cb = cbi * 16, m_ = m, a_ = a;
#pragma unroll
                for (int st = 0; st < NG; ++st) if (++cbi == cbn) { cbi = 0; if (++m == MT) { m = 0; ++a; } }
                if (m_ * 128 + eq * 32 >= rows_valid) continue;                  // warp-uniform: no valid row in this quarter
                const int rr = m_ * 128 + eq * 32 + lane;
                const bool valid = rr < rows_valid;
                float v[16];
                tc::tmem_ld16(trow + (uint32_t)((a_ * MT + m_) * Cout + cb), v);
                float sv[32];
                uint32_t pk[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    pk[jj] = valid ? pack_h16x2(v[2 * jj], v[2 * jj + 1]) : 0u;
                    const float r0 = h16_lo(pk[jj]);
                    const float r1 = h16_hi(pk[jj]);
                    sv[2 * jj] = r0; sv[2 * jj + 1] = r1;
                    sv[16 + 2 * jj] = r0 * r0; sv[16 + 2 * jj + 1] = r1 * r1;
                }
                if (valid) {
                    h16 *outp = a_ == 0 ? A.t + (vox0 + rr) * (size_t)A.ldt : A.r + (vox0 + rr) * (size_t)A.ldr;
                    if (wide_st) {        // a lane's 16 channels are one 32-byte sector: one 256-bit store (half the LSU wavefronts)
                        st_global_256(outp + cb, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
                    } else {
                        *reinterpret_cast<uint4 *>(outp + cb) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4 *>(outp + cb + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
                warp_transpose_sum<32>(sv, lane);
                const int idx = warp_transpose_owner<32>(lane);      // 0..15 sums, 16..31 squares
                atomicAdd(&s_stat[a_ * 2 * Cout + (idx >= 16 ? Cout + idx - 16 : idx) + cb], sv[0]);
            }
        }
        tc::fence_before_sync();
        __syncthreads();                 // TMEM drained, statistics in s_stat
        for (int i = tid; i < 2 * Cout; i += NT) {
            const int isq = i >= Cout, cc = isq ? i - Cout : i;
            atomicAdd(&A.t_stats[(size_t)isq * A.N * Cout + (size_t)n * Cout + cc], (double)s_stat[i]);
            s_stat[i] = 0.f;
            if (has_sc) {
                atomicAdd(&A.r_stats[(size_t)isq * A.N * Cout + (size_t)n * Cout + cc], (double)s_stat[2 * Cout + i]);
                s_stat[2 * Cout + i] = 0.f;
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)A.tmem_cols);
}

struct SlabPlan { int SZ, MT, RP, PP, cols, occ, nt; size_t smem; uint32_t raw_bytes, raw_stride, in_bytes; };

static bool slab_plan(int Cin, int Cout, bool has_sc, int D, int H, int W, int XT, SlabPlan &best) {
    const int nacc = has_sc ? 2 : 1;
    double best_score = 0.0;
    bool found = false;
    const int force_sz = L3D_ENV_INT("L3D_SLAB_SZ", 0);           // tuning / test knob: force the slab height
    for (int SZ = (D < 8 ? D : 8); SZ >= 1; --SZ) {
        if (force_sz && SZ != force_sz) continue;
        const int rows = SZ * H * W;
        const int MT = (rows + 127) / 128;
        if (MT > 4 || MT * Cout * nacc > 512) continue;
        SlabPlan p;
        p.nt = SZ * (H / 2) * (W / XT) * 8 > 256 ? 384 : 256;
        if ((SZ + 2) * H * W * 2 > ACT_SLOTS * p.nt) continue;
        p.SZ = SZ; p.MT = MT;
        p.RP = (W + 2) | 1;
        p.PP = ((H + 2) * p.RP) | 1;
        if ((SZ + 2) * p.PP * CK >= (1 << 24)) continue;
        p.raw_bytes = (uint32_t)((SZ + 2) * H * W * CK * 2);
        p.raw_stride = (p.raw_bytes + 127u) & ~127u;
        p.in_bytes = (uint32_t)((((size_t)(SZ + 2) * p.PP * CK * 4) + 127) & ~(size_t)127);
        p.smem = 2 * (size_t)p.raw_stride + p.in_bytes + 2 * (size_t)nacc * MT * 128 * CK * 2 + (size_t)nacc * Cout * Cin * 2 +
                 sizeof(float) * (2 * 27 * CK + 2 * (size_t)Cin + 4 * (size_t)Cout);
        if (p.smem > 226 * 1024) continue;
        p.cols = 32;
        while (p.cols < MT * Cout * nacc) p.cols <<= 1;
        p.occ = (int)((227 * 1024) / (p.smem + 2048));
        if (p.occ > 2) p.occ = 2;
        if (p.nt > 256) p.occ = 1;
        if (p.occ * p.cols > 512) p.occ = 512 / p.cols;
        if (p.occ < 1) continue;
        const int zs = (D + SZ - 1) / SZ;
        // useful fraction of the MMA rows x useful fraction of the staged planes x (two resident CTAs hide the per-item barriers)
        const double score = ((double)D * H * W / ((double)zs * MT * 128)) * ((double)D / ((double)zs * (SZ + 2))) * (p.occ >= 2 ? 1.3 : 1.0);
        if (!found || score > best_score) { best = p; best_score = score; found = true; }
    }
    return found;
}

}  // namespace

// Returns -1 when the slab kernel does not apply (the caller falls back to dwpw_tc_kernel).
int l3d_dwpw_fwd_slab(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                      const float *dw_w, const float *pw_w, const float *sc_w,
                      const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats, void *stream) {
    if (L3D_ENV_INT("L3D_NO_SLAB", 0) == 1) return -1;
    const int Cin = x->C, Cout = t->C;
    const bool has_sc = sc_w != nullptr;
    if (x->dtype != L3D_F16 || t->dtype != L3D_F16 || dw_w == nullptr) return -1;
    if (Cin % 16 != 0 || Cout % 16 != 0 || Cout > 256) return -1;
    if (H > 256 || W > 256 || H * W > 1024 || H % 2 != 0) return -1;
    const int XT = (W % 6 == 0) ? 6 : (W % 4 == 0) ? 4 : 0;
    if (XT == 0) return -1;
    auto aligned = [](const l3d_act *a, int mult) {
        return (a->ldc % mult == 0) && (reinterpret_cast<uintptr_t>(a->ptr) % (2 * mult) == 0);
    };
    if (!aligned(x, 8) || !aligned(t, 8) || (has_sc && (act_null(r) || !aligned(r, 8)))) return -1;
    if ((long long)N * D * H * W * x->ldc * 2 >= (1ll << 40)) return -1;
    SlabPlan p;
    if (!slab_plan(Cin, Cout, has_sc, D, H, W, XT, p)) return -1;
    const long long slabs = (long long)N * ((D + p.SZ - 1) / p.SZ);
    if (slabs >= (1ll << 30)) return -1;

    CUtensorMap tmap;
    {
        const cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        const cuuint64_t es = 2, ld = (cuuint64_t)x->ldc;
        const cuuint64_t strides[4] = {ld * es, (cuuint64_t)W * ld * es, (cuuint64_t)H * W * ld * es, (cuuint64_t)D * H * W * ld * es};
        const cuuint32_t box[5] = {CK, (cuuint32_t)W, (cuuint32_t)H, (cuuint32_t)(p.SZ + 2), 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        if (l3d_encode_tiled(&tmap, (int)CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, x->ptr, (const unsigned long long *)dims,
                             (const unsigned long long *)strides, (const unsigned *)box, (const unsigned *)estr)) return 3;
    }
    SlabArgs A;
    A.Cin = Cin; A.xn = norm_dev(xn);
    A.N = N; A.D = D; A.H = H; A.W = W;
    A.dw_w = dw_w; A.pw_w = pw_w; A.sc_w = sc_w; A.Cout = Cout;
    A.t = (h16 *)t->ptr; A.ldt = t->ldc; A.t_stats = t_stats;
    A.r = has_sc ? (h16 *)r->ptr : nullptr; A.ldr = has_sc ? r->ldc : 0; A.r_stats = r_stats;
    A.SZ = p.SZ; A.MT = p.MT; A.RP = p.RP; A.PP = p.PP; A.tmem_cols = p.cols;
    A.dbg = L3D_ENV_INT("L3D_SLAB_SKIP", 0);
    {
        auto al32 = [](const l3d_act *a) { return a->ldc % 16 == 0 && reinterpret_cast<uintptr_t>(a->ptr) % 32 == 0; };
        A.wide_st = (al32(t) && (!has_sc || al32(r)) && L3D_ENV_INT("L3D_ST256", 1) != 0) ? 1 : 0;
    }
    A.raw_bytes = p.raw_bytes; A.raw_stride = p.raw_stride; A.in_bytes = p.in_bytes;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = (long long)sms * p.occ;
    if (grid > slabs) grid = slabs;
#define L3D_SLAB_LAUNCH(XTV, NTV)                                                                                             \
    do {                                                                                                                      \
        static bool attr_set_[64] = {};  /* the attribute is per device */                                                                                                                   \
        if (dev < 0 || dev >= 64 || !attr_set_[dev]) {                                                                                                      \
            cudaError_t e = cudaFuncSetAttribute(dwpw_slab_kernel<XTV, NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); \
            if (e != cudaSuccess) { l3d_set_error("dwpw_slab: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 3; }   \
            if (dev >= 0 && dev < 64) attr_set_[dev] = true;                                                                                                  \
        }                                                                                                                     \
        dwpw_slab_kernel<XTV, NTV><<<(unsigned)grid, NTV, p.smem, (cudaStream_t)stream>>>(tmap, A);                           \
    } while (0)
    if (XT == 6) { if (p.nt == 384) L3D_SLAB_LAUNCH(6, 384); else L3D_SLAB_LAUNCH(6, 256); }
    else         { if (p.nt == 384) L3D_SLAB_LAUNCH(4, 384); else L3D_SLAB_LAUNCH(4, 256); }
#undef L3D_SLAB_LAUNCH
    l3d_count_launch();
    l3d_note_kernel("dwpw_slab_kernel");
    L3D_CUDA_OK("l3d_dwpw_fwd (slab) launch");
    return 0;
}
