// 3x3x3 convolution as an implicit GEMM on tcgen05 (bf16 storage) with TMA-staged halo tiles.
//
//   D[128 voxels][Cout] (fp32, TMEM) += A_tap[128][16] . B_tap[16][Cout]       for 27 taps x Cin/16 channel chunks
//
// * A_tap is not materialised: the activated halo tile lives in shared memory as fp16 in a planar layout
//   [8-channel group][z][y][x][8 ch], and the MMA descriptor of tap (dz,dy,dx) simply starts (dz,dy,dx) voxels
//   further into that tile (rows 16 B apart along x, row groups one y-row apart, K halves one plane apart).
//   An MMA tile is one z-plane of 16 (y) x 8 (x) voxels; a CTA tile is two planes.
// * B_tap: dense / grouped conv: W[co][ci][tap] (unet3d.py:30,49,60).  Depthwise-separable conv: the depthwise
//   and pointwise stages are both linear, so their composition is a 3x3x3 conv with W[co][ci][tap] =
//   pw[co][ci] * dw[ci][tap] (unet3d.py:20-23) -- 27x the pointwise FLOPs, all on the tensor pipe, and no
//   intermediate tensor.  The block's 1x1x1 shortcut conv (unet3d.py:70-71) is one more MMA on the centre tap.
// * Pipeline per (tile, 16-channel chunk): TMA raw box (bf16, zero-filled outside the volume) -> activation pass
//   (InstanceNorm + LeakyReLU + Dropout3d of the producer, fp16) into one of two A buffers -> one thread issues the
//   27(+1) MMAs -> tcgen05.commit.  Accumulators are double-buffered in TMEM so the epilogue of tile T (tcgen05.ld,
//   bf16 store, InstanceNorm statistics) overlaps the MMAs of tile T+1.
#include <cuda.h>

#include "l3d_common.cuh"
#include "l3d_tc.cuh"

namespace {

constexpr int TY = 16, TX = 8;                          // one MMA tile (128 rows) = one z-plane of 16 (y) x 8 (x) voxels
constexpr int HY = TY + 2, HX = TX + 2;                 // 18 x 10 halo plane
constexpr int CK = 16;
constexpr int ROWPITCH = HX * 16;                       // 160 B between y rows of one 8-channel group

// CTA tile: TZ output planes; the halo tile has TZ + 2 planes
template <int TZ>
struct Geo {
    static constexpr int HZ = TZ + 2;
    static constexpr int HVOX = HZ * HY * HX;
    static constexpr int RAW_BYTES = HVOX * CK * 2;          // TMA box, dense [z][y][x][16] bf16
    static constexpr int PLANE = HVOX * 16 + 64;             // bytes of one 8-channel group of the A tile (fp16); +64 so that the
                                                             // two groups written by a lane pair land in different bank halves
    static constexpr int A_BYTES = 2 * PLANE;
    static constexpr int ACT_ITEMS = HVOX * 2;
};

// development aid: per-tile clock64 stamps of CTA 0 (L3D_C3_DEBUG_SKIP & 8), read back with l3d_conv3_debug_read
__device__ long long g_c3_dbg[128 * 8];

struct C3Args {
    int Cin; NormDev xn;
    int N, D, H, W;
    const float *w;          // dense / grouped weights [Cout][Cin/groups][27], or NULL
    int groups;
    const float *dw_w, *pw_w;   // depthwise-separable: [Cin][27], [Cout][Cin]
    const float *sc_w;          // shortcut [Cout][Cin] or NULL
    int Cout;
    h16 *t; int ldt; double *t_stats;
    h16 *r; int ldr; double *r_stats;
    int co0, cout_total;     // dense / grouped weights: first output channel of this launch and the layer's full Cout
    int stat_ld;             // channels per (n) row of the statistics arrays (>= Cout: the launch may own a channel slice)
    int tmem_cols, nraw, nsets, merge, merged_cx, dbg;
    int tma_split;           // z-slices a halo box is requested in (several TMA operations in flight per box)
    int rot;                 // three rotating 57 KB slots instead of {raw, A0, A1}: see the kernel comment
    int nabuf;               // operand tiles: 2 (activation of item it+1 under the MMAs of item it) or 1 (half the shared memory: two CTAs per SM)
    int split2;              // prod only: the (single) raw box is requested as two z-halves with their own barriers -- see the producer warp
    int prod;                // halo boxes are requested by the dedicated producer warp (12-worker TMA variants) instead of the MMA issuer
    const void *xp; int ldx; // LD (loader warps instead of TMA): the input view / the fp32 tensor u of the rank-1 input
    const float *r1_w;       // rank-1 input (template R1): x[v][c] = r1_w[c] * u[v], u = single-channel fp32 tensor behind the tensor map
};
constexpr int R1_BOXW = 16, R1_X0 = 4;                   // rank-1 TMA box: x0 - 4 .. x0 + 11, so that the box starts on a 16-byte boundary

__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
// 16-byte asynchronous copy global -> shared; src_bytes = 0 zero-fills the destination (out-of-volume halo voxels)
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}
template <int NW>
__device__ __forceinline__ void worker_bar_n() { asm volatile("bar.sync 1, %0;" ::"n"(NW) : "memory"); }

// Warp-specialised: warps 0..7 (256 threads) are workers (activation pass + epilogue), warp 8 is the issuer (TMA
// loads and tcgen05.mma).  Hand-offs are mbarriers only, so the tensor pipe, the TMA unit and the CUDA cores run
// concurrently on different work items:
//   tma_full[r]   issuer -> workers   raw box r landed (complete_tx)
//   a_full[b]     workers -> issuer   operand tile b written (8 warp arrivals); also: raw box consumed
//   mma_done[b]   issuer -> workers   MMAs that read operand tile b finished (tcgen05.commit)
//   acc_free[s]   workers -> issuer   epilogue drained accumulator set s (8 warp arrivals)
//
// z-merged MMAs: a tcgen05.mma (M=128, K=16) costs ~44 cycles for any N <= 48 (measured, tools/ub_mma.cu), so the
// three z-taps of one input plane are issued as ONE instruction: the accumulators of consecutive output planes sit
// in consecutive TMEM column blocks, and the weight tile of (chunk, dy, dx) holds the rows [dz=2 | dz=1 | dz=0], so
// input plane zi updates output planes zi-2, zi-1, zi with a single N = 3*Cout MMA.  9*(TZ+2) MMAs per 16-channel
// chunk instead of 27*TZ.
//
// R1 (rank-1 input): the 16-channel input is x[v][c] = r1_w[c] * u[v] with a single-channel fp32 tensor u (the first
// block of a 1-channel image: conv1's pointwise stage has K = 1, unet3d.py:168,209), so only u is staged -- the TMA box is
// [z][y][16] fp32 -- and the activation pass evaluates lrelu(u * (scale_c * r1_w[c]) + shift_c) for the 16 channels.
//
// LD (thread-private cp.async staging instead of TMA): measured at 325 windows, the TMA unit needs ~6.2 K cycles for the
// 180 x 320-byte rows of a 57.6 KB halo box (~14 cycles per row + 1 cycle per 16 bytes; 3.2 K for the 180 x 64-byte rows of
// the rank-1 box), only ONE box fits next to the two operand tiles at TZ = 8, and the next box can only be requested once
// the activation pass has consumed the current one -- so the pipeline period was box latency + activation pass (7.7 K
// cycles) with the tensor pipe and the workers waiting.  In LD mode every worker thread fetches exactly the 16-byte
// vectors IT will activate (cp.async, zero-filled outside the volume) right after its activation pass and before the
// epilogue of the previous tile; no barrier is involved (cp.async.wait_all of the own copies).  Measured: faster than TMA
// only where a tile has many chunks (64 -> 32 at 24^3: -10 %), slower on the 48^3 layers -- the host picks per layer.
//
// rot (three rotating slots; used where only ONE raw box fits): with fixed roles {raw, A0, A1} the box of item it+1 can only
// be requested after the activation pass of item it has consumed the raw buffer, so the period is box latency (6.2 K
// cycles) + activation pass (1.4 K).  The three buffers have the same size, so the roles rotate instead: item it lands in
// slot X(it) and is activated into slot Y(it), with X(it+1) = Y(it-1) and Y(it+1) = X(it).  X(it+1) is free as soon as the
// MMAs of item it-1 have completed -- BEFORE item it is activated -- so warp 0 of the workers requests box it+1 at the top of
// iteration it and the TMA unit always has the next box queued.  (Measured slower than the fixed roles, see the host side.)
template <int TZ, bool MERGE, int NWARPS, bool R1 = false, bool LD = false>
__global__ void __launch_bounds__(NWARPS * 32 + 32 + (NWARPS == 12 && !LD ? 32 : 0), (TZ > 6 || NWARPS > 8 ? 1 : 2)) conv3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap2, C3Args A) {
    using G = Geo<TZ>;
    constexpr bool PRODW = NWARPS == 12 && !LD;            // a third role: one warp that only requests halo boxes
    constexpr int NW = NWARPS * 32, NT = NW + 32 + (PRODW ? 32 : 0);   // worker warps + 1 issuer warp (+ 1 producer warp)
    constexpr int EPG = NWARPS / 4;                        // worker warps per TMEM lane quarter: plane stride of the epilogue
    auto worker_bar = [] { worker_bar_n<NW>(); };
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_tma_full[4], s_a_full[2], s_mma_done[2], s_acc_free[2], s_raw_free[4];
    __shared__ uint32_t s_tmem;
    const int Cin = A.Cin, Cout = A.Cout;
    const bool has_sc = A.sc_w != nullptr;
    const int nchunks = Cin / CK, nraw = A.nraw, nsets = A.nsets;
    const int ns_m = nsets - 1;                                            // nsets is 1 or 2: set of tile tj = tj & ns_m, its use count tj >> ns_m
    const int ab_m = A.nabuf - 1, ab_sh = A.nabuf == 1 ? 0 : 1;             // operand tile of item it: it & ab_m, its use count: it >> ab_sh
    const uint32_t btile_bytes = (uint32_t)Cout * 3 * 32;                 // one [3*Cout x 16] fp16 operand tile: (chunk, dy, dx)
    const uint32_t bsc_bytes = (uint32_t)Cout * 32;
    const uint32_t b_bytes = (uint32_t)nchunks * 9 * btile_bytes;
    constexpr int RAW_STRIDE = R1 ? G::HZ * HY * R1_BOXW * 4 : G::RAW_BYTES;   // bytes between the TMA destinations
    unsigned char *s_raw = smem_raw;                                       // nraw x RAW_STRIDE (TMA destinations)
    const bool rot = !R1 && !LD && A.rot != 0;                             // slots k = 0..2 at smem_raw + k * A_BYTES
    unsigned char *sA = rot ? smem_raw : s_raw + (size_t)nraw * RAW_STRIDE;   // 2 x A_BYTES
    unsigned char *sB = rot ? smem_raw + 3 * (size_t)G::A_BYTES : sA + (size_t)A.nabuf * G::A_BYTES;   // [chunk][dy*3+dx][3*Cout x 16]
    auto slot_x = [](int it) { const int m = it % 3; return m == 0 ? 0 : m == 1 ? 2 : 1; };   // raw slot of item it
    auto slot_y = [](int it) { const int m = it % 3; return m == 0 ? 1 : m == 1 ? 0 : 2; };   // operand slot of item it
    unsigned char *sB2 = sB + b_bytes;                                     // shortcut: [chunk][Cout x 16]
    float *s_scale = reinterpret_cast<float *>(sB2 + (has_sc ? nchunks * bsc_bytes : 0));
    float *s_shift = s_scale + Cin;
    float *s_stat = s_shift + Cin;                                         // 2*Cout (t) + 2*Cout (r)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == NW / 32) tc::tmem_alloc(&s_tmem, (uint32_t)A.tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { tc::mbar_init(&s_tma_full[i], 1); tc::mbar_init(&s_raw_free[i], NW / 32); }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&s_a_full[i], NW / 32);
            tc::mbar_init(&s_mma_done[i], 1); tc::mbar_init(&s_acc_free[i], NW / 32);
        }
    }
    // ---- stage the (effective) 3x3x3 weights once per CTA as fp16 K-major operand tiles
    {
        const int cin_g = Cin / A.groups, cout_g = A.cout_total / A.groups;
        const int rows = 3 * Cout;
        // one (co, ci) pair per step: its 27 taps are contiguous in global memory and all loads are in flight together
        for (int i = tid; i < Cout * Cin; i += NT) {
            const int ci = i % Cin, co = i / Cin;
            float wv[27];
            if (A.w != nullptr) {
                const int cog = A.co0 + co;
                const int g = cog / cout_g, cl = ci - g * cin_g;
                const bool in_group = cl >= 0 && cl < cin_g;
                const float *src = A.w + ((size_t)cog * cin_g + (in_group ? cl : 0)) * 27;
#pragma unroll
                for (int tap = 0; tap < 27; ++tap) wv[tap] = in_group ? src[tap] : 0.f;
            } else {
                const float pwv = A.pw_w[(size_t)co * Cin + ci];
                const float *src = A.dw_w + (size_t)ci * 27;
#pragma unroll
                for (int tap = 0; tap < 27; ++tap) wv[tap] = src[tap];
#pragma unroll
                for (int tap = 0; tap < 27; ++tap) wv[tap] *= pwv;
            }
            const int ch = ci >> 4, k = ci & 15;
#pragma unroll
            for (int tap = 0; tap < 27; ++tap) {
                const int dz = tap / 9, t9 = tap - dz * 9;
                *reinterpret_cast<__half *>(sB + (size_t)(ch * 9 + t9) * btile_bytes + tc::tile_off((2 - dz) * Cout + co, k, rows)) = __float2half_rn(wv[tap]);
            }
        }
        if (has_sc)
            for (int i = tid; i < Cout * Cin; i += NT) {
                const int ci = i % Cin, co = i / Cin;
                *reinterpret_cast<__half *>(sB2 + (size_t)(ci >> 4) * bsc_bytes + tc::tile_off(co, ci & 15, Cout)) = __float2half_rn(A.sc_w[i]);
            }
    }
    for (int i = tid; i < 4 * Cout; i += NT) s_stat[i] = 0.f;
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB), sB2_u = tc::smem_u32(sB2);
    const int nacc = has_sc ? 2 : 1;
    const int acc_cols = TZ * Cout * nacc;          // TMEM columns of one accumulator set: [plane][Cout] main, then shortcut

    const int tilesX = (A.W + TX - 1) / TX, tilesY = (A.H + TY - 1) / TY, tilesZ = (A.D + TZ - 1) / TZ;
    const int tiles_per_sample = tilesX * tilesY * tilesZ;
    const int total_tiles = tiles_per_sample * A.N;
    // contiguous tile range per CTA (keeps a CTA inside one sample as long as possible, halo re-use in L2)
    const int per = (total_tiles + gridDim.x - 1) / gridDim.x;
    const int tile_begin = blockIdx.x * per;
    const int tile_end = min(total_tiles, tile_begin + per);
    const int n_items = (tile_end > tile_begin ? tile_end - tile_begin : 0) * nchunks;
    auto tile_coord = [&](int tile, int &n, int &z0, int &y0, int &x0) {
        n = tile / tiles_per_sample;
        int b = tile - n * tiles_per_sample;
        x0 = (b % tilesX) * TX; b /= tilesX;
        y0 = (b % tilesY) * TY; b /= tilesY;
        z0 = b * TZ;
    };

    // the tile after (n, z0, y0, x0) in this CTA's raster order (no divisions: the integer divisions of tile_coord cost the
    // worker warps ~2 K cycles per tile between the end of an epilogue and the next activation pass -- clock stamps)
    auto tile_next = [&](int &n, int &z0, int &y0, int &x0) {
        x0 += TX;
        if (x0 >= A.W) { x0 = 0; y0 += TY; if (y0 >= A.H) { y0 = 0; z0 += TZ; if (z0 >= A.D) { z0 = 0; ++n; } } }
    };
    // bit i set <=> halo coordinate c0 - 1 + i lies inside [0, ext)   (i < len <= 18)
    auto axis_mask = [](int c0, int len, int ext) -> uint32_t {
        const int lo = c0 >= 1 ? 0 : 1 - c0, hi = min(len, ext - c0 + 1);
        return hi > lo ? ((1u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
    };
    // TMA cursor (own copy per warp; used by whichever warp requests the boxes): (tile, chunk) of the next box, advanced
    // incrementally (no divisions on the issue path)
    int pf_item = 0, pf_ch = 0, pf_rb = 0, pf_n, pf_z0, pf_y0, pf_x0;
    tile_coord(tile_begin < tile_end ? tile_begin : 0, pf_n, pf_z0, pf_y0, pf_x0);
    auto issue_tma = [&]() {
        if (LD || rot || pf_item >= n_items) return;
        const int rb = pf_rb;
        if (tc::elect_one()) {
            tc::mbar_expect_tx(&s_tma_full[rb], RAW_STRIDE);
            const int nsl = A.tma_split, zs = G::HZ / nsl;           // planes per slice
            const uint32_t sl_bytes = (uint32_t)(RAW_STRIDE / nsl);
            for (int sl = 0; sl < nsl; ++sl) {
                unsigned char *dst = s_raw + (size_t)rb * RAW_STRIDE + (size_t)sl * sl_bytes;
                const int zc = pf_z0 - 1 + sl * zs;
                if (R1) tc::tma_load_4d(dst, &tmap, &s_tma_full[rb], pf_x0 - R1_X0, pf_y0 - 1, zc, pf_n);
                else if (A.merged_cx) tc::tma_load_4d(dst, pf_ch ? &tmap2 : &tmap, &s_tma_full[rb], (pf_x0 - 1) * CK, pf_y0 - 1, zc, pf_n);
                else tc::tma_load_5d(dst, &tmap, &s_tma_full[rb], pf_ch * CK, pf_x0 - 1, pf_y0 - 1, zc, pf_n);
            }
        }
        ++pf_item;
        if (++pf_rb == nraw) pf_rb = 0;
        if (++pf_ch == nchunks) {
            pf_ch = 0;
            pf_x0 += TX;
            if (pf_x0 >= A.W) { pf_x0 = 0; pf_y0 += TY; if (pf_y0 >= A.H) { pf_y0 = 0; pf_z0 += TZ; if (pf_z0 >= A.D) { pf_z0 = 0; ++pf_n; } } }
        }
    };
    const bool prod = PRODW && A.prod != 0 && !rot;
    const bool split2 = prod && !R1 && A.split2 != 0;                      // (host: nraw == 1, even halo height)
    if (PRODW && warp == NW / 32 + 1) {
        // =============================== producer warp ===============================
        // Requests the halo boxes as soon as their raw buffer has been consumed.  With the requests on the MMA issuer's
        // instruction stream the cp.async.bulk.tensor held that warp for ~1.4 K cycles per box (clock stamps) before the
        // first MMA of the tile went out, and a second box in flight blocked it for the whole TMA service time.
        //
        // split2 (one raw buffer): the box travels as two z-halves with their own {tma_full, raw_free} barriers.  The activation
        // pass walks the box in z order, so the lower half of box it+1 is requested while the upper half of box it is still being
        // activated, and the workers start on box it+1 as soon as ITS lower half has landed -- the chain "activation pass ->
        // request -> box latency" (clock stamps: 2.5 K + 4.4 K cycles of a 7.1 K period) no longer bounds the tile.
        if (prod && split2) {
            constexpr int HALF_BYTES = G::RAW_BYTES / 2;
            for (int i = 0; i < n_items; ++i) {
                for (int h = 0; h < 2; ++h) {
                    if (i >= 1) tc::mbar_wait(&s_raw_free[h], (uint32_t)((i - 1) & 1));
                    if (tc::elect_one()) {
                        tc::mbar_expect_tx(&s_tma_full[h], HALF_BYTES);
                        unsigned char *dst = s_raw + (size_t)h * HALF_BYTES;
                        const int zc = pf_z0 - 1 + h * (G::HZ / 2);
                        if (A.merged_cx) tc::tma_load_4d(dst, pf_ch ? &tmap2 : &tmap, &s_tma_full[h], (pf_x0 - 1) * CK, pf_y0 - 1, zc, pf_n);
                        else tc::tma_load_5d(dst, &tmap, &s_tma_full[h], pf_ch * CK, pf_x0 - 1, pf_y0 - 1, zc, pf_n);
                    }
                    __syncwarp();
                }
                if (++pf_ch == nchunks) { pf_ch = 0; tile_next(pf_n, pf_z0, pf_y0, pf_x0); }
            }
        } else if (prod)
            for (int i = 0; i < n_items; ++i) {
                if (i >= nraw) tc::mbar_wait(&s_raw_free[i % nraw], (uint32_t)(((i / nraw) - 1) & 1));
                issue_tma();
                __syncwarp();
            }
        else if (A.dbg & 256)      // development aid: an otherwise idle warp observes when each box lands
            for (int i = 0; i < n_items; ++i) {
                tc::mbar_wait(&s_tma_full[i % nraw], (uint32_t)((i / nraw) & 1));
                if (blockIdx.x == 0 && i < 128 && lane == 0) g_c3_dbg[i * 8 + 2] = clock64();
            }
    } else
    if (warp == NW / 32) {
        // =============================== issuer warp ===============================
        // The whole warp runs this loop converged (every value is warp-uniform, so descriptors and TMEM addresses
        // live in uniform registers); one elected lane issues the TMA loads and the tcgen05.mma / commit.  A
        // single-lane branch around the loop makes ptxas wrap every MMA in an ELECT / R2UR.BROADCAST waterfall.
        const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem);    // TMEM base (from shared memory) as a uniform value
        if (!prod)
            for (int i = 0; i < nraw; ++i) issue_tma();
        const uint32_t idesc1 = tc::idesc_f16_m128(Cout), idesc2 = tc::idesc_f16_m128(2 * Cout), idesc3 = tc::idesc_f16_m128(3 * Cout);
        const uint32_t brow = (uint32_t)(Cout >> 3) * 128 >> 4;     // descriptor units (16 B) per Cout rows of a weight tile
        const uint32_t bstep = btile_bytes >> 4;
        int tj = 0, ch = 0;                       // tile index within this CTA's range, chunk
        for (int it = 0; it < n_items; ++it) {
            const int buf = it & ab_m, set = tj & ns_m;
            tc::mbar_wait(&s_a_full[buf], (uint32_t)((it >> ab_sh) & 1));        // operand tile written, raw box consumed
            const bool stamp = (A.dbg & 8) && blockIdx.x == 0 && it < 128;
            if (stamp && lane == 0) g_c3_dbg[it * 8 + 5] = clock64();
            if (!prod) issue_tma();                                          // refill the raw box just consumed
            if (ch == 0 && tj >= nsets) tc::mbar_wait(&s_acc_free[set], (uint32_t)(((tj >> ns_m) - 1) & 1));   // epilogue of tile tj-nsets done
            tc::fence_after_sync();
            if (stamp && lane == 0) g_c3_dbg[it * 8 + 6] = clock64();
            const uint64_t ad0 = tc::smem_desc(sA_u + (rot ? slot_y(it) : buf) * G::A_BYTES, G::PLANE, ROWPITCH);
            const uint64_t bd0 = tc::smem_desc(sB_u + (uint32_t)(ch * 9) * btile_bytes, 3 * Cout * 16, 128);
            const uint32_t d_t = tmem_u + (uint32_t)(set * acc_cols);
            const bool first = ch == 0;
            if (tc::elect_one()) {
              if (!(A.dbg & 4)) {
#pragma unroll
                for (int zi = 0; zi < TZ + 2; ++zi) {
                    const int lo = zi >= 2 ? zi - 2 : 0, hi = zi <= TZ - 1 ? zi : TZ - 1;     // output planes fed by input plane zi
                    const int np = hi - lo + 1;
                    const uint64_t bdz = bd0 + (uint64_t)((uint32_t)(lo - zi + 2) * brow);     // first weight row block: dz = zi - lo
                    const uint64_t adz = ad0 + (uint64_t)((uint32_t)(zi * HY * ROWPITCH) >> 4);
                    const uint32_t d_lo = d_t + (uint32_t)(lo * Cout);
                    if (MERGE) {
                        const uint32_t idn = np == 3 ? idesc3 : np == 2 ? idesc2 : idesc1;
                        if (zi <= TZ - 1 && first) {
                            // plane zi gets its first contribution (dz = 0): overwrite it, accumulate into the other planes
                            if (np > 1) tc::mma_f16(d_lo, adz, bdz, np == 3 ? idesc2 : idesc1, 1u);
                            tc::mma_f16(d_t + (uint32_t)(zi * Cout), adz, bdz + (uint64_t)((uint32_t)(np - 1) * brow), idesc1, 0u);
                        } else {
                            tc::mma_f16(d_lo, adz, bdz, idn, 1u);
                        }
#pragma unroll
                        for (int t9 = 1; t9 < 9; ++t9) {
                            const int dy = t9 / 3, dx = t9 % 3;
                            tc::mma_f16(d_lo, adz + (uint64_t)((uint32_t)(dy * ROWPITCH + dx * 16) >> 4), bdz + (uint64_t)(t9 * bstep), idn, 1u);
                        }
                    } else {
#pragma unroll
                        for (int t9 = 0; t9 < 9; ++t9) {
                            const int dy = t9 / 3, dx = t9 % 3;
                            const uint64_t ad = adz + (uint64_t)((uint32_t)(dy * ROWPITCH + dx * 16) >> 4);
#pragma unroll
                            for (int p = lo; p <= hi; ++p)
                                tc::mma_f16(d_t + (uint32_t)(p * Cout), ad, bdz + (uint64_t)(t9 * bstep + (uint32_t)(p - lo) * brow), idesc1,
                                            (t9 == 0 && p == zi && first) ? 0u : 1u);
                        }
                    }
                }
                if (has_sc) {
                    const uint64_t bd2 = tc::smem_desc(sB2_u + (uint32_t)ch * bsc_bytes, Cout * 16, 128);
#pragma unroll
                    for (int p = 0; p < TZ; ++p) {
                        const uint32_t aoff = (uint32_t)((((p + 1) * HY + 1) * ROWPITCH + 16) >> 4);
                        tc::mma_f16(d_t + (uint32_t)((TZ + p) * Cout), ad0 + aoff, bd2, idesc1, ch > 0 ? 1u : 0u);
                    }
                }
              }
                tc::mma_commit(&s_mma_done[buf]);
            }
            __syncwarp();
            if (stamp && lane == 0) g_c3_dbg[it * 8 + 7] = clock64();
            if (++ch == nchunks) { ch = 0; ++tj; }
        }
    } else {
        // ===================================== workers =====================================
        // activation-pass role: a thread owns CPT fixed columns (halo y, halo x, 8-channel half q) of the raw box and walks them
        // in z: the vector of plane hz sits hz * COLS vectors further on in the box and in the operand tile, so the loop needs no
        // per-vector coordinates.  (It used a per-thread table of packed (x, y, z) items first: ptxas kept that table in local
        // memory, and its eleven reloads at the top of every tile missed the small L1 left beside ~200 KB of shared memory --
        // ~1 K cycles per tile between the end of an epilogue and the next activation pass, clock stamps.)
        constexpr int COLS = HY * HX * 2;                       // 16-byte vectors per halo plane
        constexpr int CPT = (COLS + NW - 1) / NW;               // columns per thread (1 with 12 worker warps, 2 with 8)
        const int aq = tid & 1;                                 // NW is even: all of a thread's columns share the channel half
        int col_hx[CPT], col_hy[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int v2 = (tid + c * NW) >> 1;                 // halo voxel within the plane
            col_hx[c] = v2 % HX; col_hy[c] = v2 / HX;          // (col_hy >= HY: no such column)
        }
        // LD: byte offset of each column's plane-0 vector from the tile's halo origin in global memory
        uint32_t g_off[LD ? CPT : 1];
        if (LD) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) g_off[c] = (uint32_t)((col_hy[c] * A.W + col_hx[c]) * A.ldx * 2 + aq * 16);
        }
        auto issue_copies = [&](int n, int z0, int y0, int x0, int chunk) {
            const uint32_t mz = axis_mask(z0, G::HZ, A.D), my = axis_mask(y0, HY, A.H), mx = axis_mask(x0, HX, A.W);
            const char *org = reinterpret_cast<const char *>(A.xp) +
                              (((((long long)n * A.D + (z0 - 1)) * A.H + (y0 - 1)) * A.W + (x0 - 1)) * A.ldx + chunk * CK) * 2;
            const uint32_t zstep = (uint32_t)(A.H * A.W * A.ldx * 2);         // bytes between z-planes (host: the box spans < 2^31 bytes)
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                if (CPT * NW == COLS || tid + c * NW < COLS) {
                    const bool xy_ok = ((mx >> col_hx[c]) & (my >> col_hy[c]) & 1u) != 0;
#pragma unroll
                    for (int hz = 0; hz < G::HZ; ++hz) {
                        const bool ok = xy_ok && ((mz >> hz) & 1u) != 0;
                        cp_async16(s_raw + (size_t)(hz * COLS + tid + c * NW) * 16, ok ? org + g_off[c] + hz * zstep : reinterpret_cast<const char *>(A.xp), ok ? 16u : 0u);
                    }
                }
            }
        };
        int wn, wz0, wy0, wx0;                                  // this warp's tile cursor
        tile_coord(tile_begin < tile_end ? tile_begin : 0, wn, wz0, wy0, wx0);
        if (LD && tile_begin < tile_end) issue_copies(wn, wz0, wy0, wx0, 0);
        // rot: request the halo box of item j = (tile tl, chunk) into slot X(j)  (one lane)
        auto rot_issue = [&](int j, int tl, int chunk) {
            int n, z0, y0, x0;
            tile_coord(tl, n, z0, y0, x0);
            const int b3 = j % 3;
            unsigned char *dst = smem_raw + (size_t)slot_x(j) * G::A_BYTES;
            tc::mbar_expect_tx(&s_tma_full[b3], G::RAW_BYTES);
            if (A.merged_cx) tc::tma_load_4d(dst, chunk ? &tmap2 : &tmap, &s_tma_full[b3], (x0 - 1) * CK, y0 - 1, z0 - 1, n);
            else tc::tma_load_5d(dst, &tmap, &s_tma_full[b3], chunk * CK, x0 - 1, y0 - 1, z0 - 1, n);
        };
        if (rot && tid == 0 && n_items > 0) {
            rot_issue(0, tile_begin, 0);
            if (n_items > 1) rot_issue(1, nchunks > 1 ? tile_begin : tile_begin + 1, nchunks > 1 ? 1 : 0);
        }
        // epilogue role: voxel row of the MMA tiles (planes) em, em + EPG, ...
        const int em = warp >> 2, erow = (warp & 3) * 32 + lane;
        const int elx = erow & 7, ely = erow >> 3;
        int stat_n = -1;

        auto flush_stats = [&](int n) {
            if (n < 0) return;
            for (int i = tid; i < 2 * Cout; i += NW) {
                const int isq = i >= Cout, cc = isq ? i - Cout : i;
                atomicAdd(&A.t_stats[(size_t)isq * A.N * A.stat_ld + (size_t)n * A.stat_ld + cc], (double)s_stat[i]);
                s_stat[i] = 0.f;
                if (has_sc) {
                    atomicAdd(&A.r_stats[(size_t)isq * A.N * A.stat_ld + (size_t)n * A.stat_ld + cc], (double)s_stat[2 * Cout + i]);
                    s_stat[2 * Cout + i] = 0.f;
                }
            }
        };
        // epilogue of tile `tile` (accumulator set `set`): TMEM -> bf16 global + statistics
        auto epilogue = [&](int n, int z0, int y0, int x0, int set) {
            if (n != stat_n) {
                worker_bar();
                flush_stats(stat_n);
                stat_n = n;
                worker_bar();
            }
            const int gy = y0 + ely, gx = x0 + elx;
            const bool valid_yx = gy < A.H && gx < A.W;
            const size_t vox0 = (((size_t)n * A.D + z0) * A.H + gy) * A.W + gx;
            const size_t zstride = (size_t)A.H * A.W;
            const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(set * acc_cols);
            if (!(A.dbg & 2))
            for (int a = 0; a < nacc; ++a) {
                h16 *outb = a == 0 ? A.t : A.r;
                const int ldo = a == 0 ? A.ldt : A.ldr;
                float *stat = s_stat + a * 2 * Cout;
                for (int cb = 0; cb < Cout; cb += 16) {
                    // Lanes 2j / 2j+1 hold x-adjacent voxels.  After the half swap below the even lane owns channels 0-7 and the odd
                    // lane channels 8-15 of BOTH voxels -- so that every store instruction fills whole 32-byte sectors, and so that a
                    // lane accumulates the statistics of 8 channels only (16 accumulators instead of 32: room for the TMEM load of
                    // the next plane to be in flight under the work on the current one).
                    float sv[16];                  // [0..7] sums, [8..15] sums of squares of channels (odd ? 8 : 0) + j
#pragma unroll
                    for (int j = 0; j < 16; ++j) sv[j] = 0.f;
                    const bool odd = (lane & 1) != 0;
                    uint32_t vr[16];
                    if (em < TZ) tc::tmem_ld16_issue(trow + (uint32_t)((a * TZ + em) * Cout + cb), vr);
                    // this lane's 16-byte store slot in plane em, advanced by EPG planes per step (no 64-bit multiplies in the loop)
                    h16 *own = outb + (vox0 + (size_t)em * zstride) * (size_t)ldo + cb + (odd ? 8 : 0);
                    const size_t own_step = (size_t)EPG * zstride * (size_t)ldo;
                    const int pair_off = odd ? -ldo : ldo;                              // the other voxel of the lane pair
#pragma unroll 1
                    for (int p = em; p < TZ; p += EPG, own += own_step) {
                        tc::tmem_ld_wait16(vr);
                        const bool valid = valid_yx && (z0 + p < A.D);
                        uint32_t pk[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) pk[j] = valid ? pack_h16x2(__uint_as_float(vr[2 * j]), __uint_as_float(vr[2 * j + 1])) : 0u;
                        if (p + EPG < TZ) tc::tmem_ld16_issue(trow + (uint32_t)((a * TZ + p + EPG) * Cout + cb), vr);
                        const uint4 h0 = make_uint4(pk[0], pk[1], pk[2], pk[3]), h1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        const uint4 snd = odd ? h0 : h1, mine = odd ? h1 : h0;
                        uint4 rcv;
                        rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, 1); rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, 1);
                        rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, 1); rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, 1);
                        const bool pvalid = __shfl_xor_sync(0xffffffffu, valid ? 1 : 0, 1) != 0;
                        if (valid) *reinterpret_cast<uint4 *>(own) = mine;                   // own voxel, own channel half
                        if (pvalid) *reinterpret_cast<uint4 *>(own + pair_off) = rcv;       // the partner's voxel, same channel half
                        // statistics of the stored (fp16-rounded) values; an invalid voxel was packed as zeros
                        const uint32_t mw[4] = {mine.x, mine.y, mine.z, mine.w}, rw4[4] = {rcv.x, rcv.y, rcv.z, rcv.w};
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            const float m0 = h16_lo(mw[w]), m1 = h16_hi(mw[w]), q0 = h16_lo(rw4[w]), q1 = h16_hi(rw4[w]);
                            sv[2 * w] += m0 + q0; sv[2 * w + 1] += m1 + q1;
                            sv[8 + 2 * w] = fmaf(m0, m0, fmaf(q0, q0, sv[8 + 2 * w])); sv[8 + 2 * w + 1] = fmaf(m1, m1, fmaf(q1, q1, sv[8 + 2 * w + 1]));
                        }
                    }
                    warp_transpose_sum_parity16(sv, lane);
                    {
                        const int idx = warp_transpose_owner<16>(lane);      // bits 16, 8, 4, 2 of the lane select the value index
                        const int chn = (odd ? 8 : 0) + (idx & 7);
                        atomicAdd(&stat[(idx >= 8 ? Cout : 0) + chn + cb], sv[0]);
                    }
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_acc_free[set]);
        };

        int cur_n = -1, it = 0;
        int w_rb = 0;                                           // raw box of item it and the parity of its use count (it % nraw, (it / nraw) & 1)
        uint32_t w_rph = 0;
        int pn = 0, pz0 = 0, py0 = 0, px0 = 0;                  // the previous tile (its epilogue runs one work item later)
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const int n = wn, z0 = wz0, y0 = wy0, x0 = wx0;
            tile_next(wn, wz0, wy0, wx0);                       // (wn, ...) is now the next tile
            if (n != cur_n) {
                cur_n = n;
                worker_bar();    // nobody still reads the previous sample's scale/shift
                for (int cc = tid; cc < Cin; cc += NW) {
                    float sc, sh;
                    norm_scale_shift(A.xn, A.N, Cin, n, cc, sc, sh);
                    if (R1) sc *= A.r1_w[cc];
                    s_scale[cc] = sc; s_shift[cc] = sh;
                }
                worker_bar();
            }
            // validity of the halo coordinates of this tile as per-axis bit masks
            const uint32_t mz = axis_mask(z0, G::HZ, A.D), my = axis_mask(y0, HY, A.H), mx = axis_mask(x0, HX, A.W);
            const int tj = tile - tile_begin;
            for (int ch = 0; ch < nchunks; ++ch, ++it) {
                const int buf = it & ab_m, rb = LD ? 0 : w_rb;
                unsigned char *Ab = sA + (size_t)(rot ? slot_y(it) : buf) * G::A_BYTES;
                const unsigned char *Rb = rot ? smem_raw + (size_t)slot_x(it) * G::A_BYTES : s_raw + (size_t)rb * RAW_STRIDE;
                if (rot && warp == 0 && it >= 1 && it + 1 < n_items) {
                    // slot X(it+1) = Y(it-1) is free once the MMAs of item it-1 have completed: request box it+1 now
                    tc::mbar_wait(&s_mma_done[(it - 1) & 1], (uint32_t)(((it - 1) >> 1) & 1));
                    if (lane == 0) {
                        if (ch + 1 < nchunks) rot_issue(it + 1, tile, ch + 1);
                        else rot_issue(it + 1, tile + 1, 0);
                    }
                    __syncwarp();
                }
                // scale / shift of this thread's 8 channels
                const float4 sc0 = *reinterpret_cast<const float4 *>(s_scale + ch * CK + aq * 8);
                const float4 sc1 = *reinterpret_cast<const float4 *>(s_scale + ch * CK + aq * 8 + 4);
                const float4 sh0 = *reinterpret_cast<const float4 *>(s_shift + ch * CK + aq * 8);
                const float4 sh1 = *reinterpret_cast<const float4 *>(s_shift + ch * CK + aq * 8 + 4);
                const __half2 sl2 = __float2half2_rn(A.xn.slope);
                const bool ident = A.xn.stats == nullptr;
                const bool stamp = (A.dbg & 8) && blockIdx.x == 0 && it < 128 && tid == 0;
                if (stamp && (A.dbg & 512)) g_c3_dbg[it * 8 + 1] = clock64();          // development aid: before the box wait
                if (LD) asm volatile("cp.async.wait_all;" ::: "memory");                   // own copies of this item landed
                else if (rot) {
                    tc::mbar_wait(&s_tma_full[it % 3], (uint32_t)((it / 3) & 1));
                    // Y(it) = X(it-1): every worker warp must be done reading the raw box of item it-1
                    if (it >= 1) tc::mbar_wait(&s_a_full[(it - 1) & 1], (uint32_t)(((it - 1) >> 1) & 1));
                }
                else if (split2) tc::mbar_wait(&s_tma_full[0], (uint32_t)(it & 1));       // lower z-half of this item's box landed
                else tc::mbar_wait(&s_tma_full[rb], w_rph);                                // raw box of this item landed
                if (stamp) g_c3_dbg[it * 8 + 0] = clock64();
                if (it > ab_m) tc::mbar_wait(&s_mma_done[buf], (uint32_t)(((it >> ab_sh) - 1) & 1));   // MMAs of the item that used A[buf] done
                if (stamp && !(A.dbg & 512)) g_c3_dbg[it * 8 + 1] = clock64();
                // ---- activation pass: raw bf16 [z][y][x][16] -> fp16 planar [q][z][y][x][8]; the shared-memory loads of a
                // batch of vectors are issued before any of them is used
                constexpr int ACT_BATCH = 4;
                if (!(A.dbg & 1))
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const bool col_ok = CPT * NW == COLS || tid + c * NW < COLS;
                    // conv zero padding of the ACTIVATED tensor: x / y validity of the column, z validity per plane below
                    const bool xy_ok = ((mx >> col_hx[c]) & (my >> col_hy[c]) & 1u) != 0;
                    const unsigned char *Rc = Rb + (size_t)(tid + c * NW) * 16;                               // plane 0 of the column
                    unsigned char *Ac = Ab + (size_t)aq * G::PLANE + (size_t)((tid + c * NW) >> 1) * 16;
                    const float *Ru = reinterpret_cast<const float *>(Rb) + col_hy[c] * R1_BOXW + col_hx[c] + (R1_X0 - 1);   // R1
#pragma unroll
                    for (int k0 = 0; k0 < G::HZ; k0 += ACT_BATCH) {
                        uint4 rw[ACT_BATCH];
                        float ru[ACT_BATCH];
                        // split2: the first batch that reaches into the upper z-half waits for it
                        if (!R1 && !LD && c == 0 && k0 + ACT_BATCH > G::HZ / 2 && k0 <= G::HZ / 2) {
                            if (split2) tc::mbar_wait(&s_tma_full[1], (uint32_t)(it & 1));
                        }
#pragma unroll
                        for (int kk = 0; kk < ACT_BATCH; ++kk) {
                            const int hz = k0 + kk;
                            rw[kk] = make_uint4(0u, 0u, 0u, 0u);
                            ru[kk] = 0.f;
                            if (hz < G::HZ && col_ok) {
                                if (R1) ru[kk] = Ru[hz * (HY * R1_BOXW)];
                                else rw[kk] = *reinterpret_cast<const uint4 *>(Rc + (size_t)hz * (COLS * 16));
                            }
                        }
#pragma unroll
                        for (int kk = 0; kk < ACT_BATCH; ++kk) {
                            const int hz = k0 + kk;
                            if (hz < G::HZ && col_ok) {
                                const uint4 r4 = rw[kk];
                                float f[8];
                                if (R1) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) f[j] = ru[kk];
                                } else {
                                    f[0] = h16_lo(r4.x); f[1] = h16_hi(r4.x);
                                    f[2] = h16_lo(r4.y); f[3] = h16_hi(r4.y);
                                    f[4] = h16_lo(r4.z); f[5] = h16_hi(r4.z);
                                    f[6] = h16_lo(r4.w); f[7] = h16_hi(r4.w);
                                }
                                uint4 o;
                                if (!R1 && ident) {        // already-activated input (a block's first conv): the stored fp16 values are the operand
                                    o = r4;
                                } else {
                                    f[0] = fmaf(f[0], sc0.x, sh0.x); f[1] = fmaf(f[1], sc0.y, sh0.y); f[2] = fmaf(f[2], sc0.z, sh0.z); f[3] = fmaf(f[3], sc0.w, sh0.w);
                                    f[4] = fmaf(f[4], sc1.x, sh1.x); f[5] = fmaf(f[5], sc1.y, sh1.y); f[6] = fmaf(f[6], sc1.z, sh1.z); f[7] = fmaf(f[7], sc1.w, sh1.w);
                                    // LeakyReLU on the packed fp16 values (0 <= slope <= 1): max(v, slope * v)
                                    __half2 h[4];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        h[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                                        h[j] = __hmax2(h[j], __hmul2(h[j], sl2));
                                    }
                                    o = make_uint4(*reinterpret_cast<uint32_t *>(&h[0]), *reinterpret_cast<uint32_t *>(&h[1]),
                                                   *reinterpret_cast<uint32_t *>(&h[2]), *reinterpret_cast<uint32_t *>(&h[3]));
                                    // An identity-norm input needs no padding test: TMA (and the zero-size cp.async of the loader mode)
                                    // already delivered zeros for the voxels outside the volume
                                    if (!(xy_ok && ((mz >> hz) & 1u))) o = make_uint4(0u, 0u, 0u, 0u);
                                }
                                *reinterpret_cast<uint4 *>(Ac + (size_t)hz * (HY * HX * 16)) = o;
                            }
                        }
                        // split2: the batch that holds the last vectors of the lower half has them in registers now
                        if (!R1 && !LD && c == CPT - 1 && k0 + ACT_BATCH >= G::HZ / 2 && k0 < G::HZ / 2) {
                            if (split2) { __syncwarp(); if (lane == 0) mbar_arrive(&s_raw_free[0]); }
                        }
                    }
                }
                tc::fence_async_smem();
                __syncwarp();
                if (stamp && !(A.dbg & 256)) g_c3_dbg[it * 8 + 2] = clock64();
                if (lane == 0) { mbar_arrive(&s_a_full[buf]); if (prod) mbar_arrive(&s_raw_free[split2 ? 1 : rb]); }
                if (++w_rb == nraw) { w_rb = 0; w_rph ^= 1u; }
                if (LD) {
                    // own raw vectors consumed: fetch the ones of the next work item (they land under the epilogue below)
                    if (ch + 1 < nchunks) issue_copies(n, z0, y0, x0, ch + 1);
                    else if (tile + 1 < tile_end) issue_copies(wn, wz0, wy0, wx0, 0);
                }
                // ---- epilogue of the previous tile (its MMAs were issued one work item ago)
                if (ch == 0 && tile > tile_begin) {
                    const int pit = it - 1;
                    tc::mbar_wait(&s_mma_done[pit & ab_m], (uint32_t)((pit >> ab_sh) & 1));
                    tc::fence_after_sync();
                    if (stamp) g_c3_dbg[it * 8 + 3] = clock64();
                    epilogue(pn, pz0, py0, px0, (tj - 1) & ns_m);
                    if (stamp) g_c3_dbg[it * 8 + 4] = clock64();
                }
            }
            pn = n; pz0 = z0; py0 = y0; px0 = x0;
        }
        if (tile_begin < tile_end) {
            const int pit = it - 1;
            tc::mbar_wait(&s_mma_done[pit & ab_m], (uint32_t)((pit >> ab_sh) & 1));
            tc::fence_after_sync();
            epilogue(pn, pz0, py0, px0, (tile_end - 1 - tile_begin) & ns_m);
        }
        worker_bar();
        flush_stats(stat_n);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == NW / 32) tc::tmem_dealloc(tmem, (uint32_t)A.tmem_cols);
}

static size_t c3_smem_bytes(int TZ, int Cin, int Cout, bool has_sc, int nraw, bool rank1 = false, int nabuf = 2) {
    const size_t nch = Cin / CK;
    const size_t hvox = (size_t)(TZ + 2) * HY * HX;
    const size_t raw = rank1 ? (size_t)(TZ + 2) * HY * R1_BOXW * 4 : hvox * CK * 2, a_bytes = 2 * (hvox * 16 + 64);
    return (size_t)nraw * raw + (size_t)nabuf * a_bytes + nch * 27 * (size_t)Cout * 32 + (has_sc ? nch * (size_t)Cout * 32 : 0) +
           sizeof(float) * (2 * (size_t)Cin + 4 * (size_t)Cout);
}


}  // namespace

extern "C" int l3d_conv3_debug_read(long long *host, int n) {
    return cudaMemcpyFromSymbol(host, g_c3_dbg, sizeof(long long) * (size_t)(n < 1024 ? n : 1024)) == cudaSuccess ? 0 : 2;
}

// Implicit-GEMM 3x3x3 conv on tcgen05.  Exactly one of {w} / {dw_w, pw_w} is given.  Returns -1 when the path
// does not apply (the caller falls back to another kernel).
// r1_w != NULL: rank-1 input -- x describes the virtual 16-channel tensor x[v][c] = r1_w[c] * u[v] and x->ptr is the
// single-channel fp32 tensor u.
int l3d_conv3_tc_ex(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                    const float *w, int groups, const float *dw_w, const float *pw_w, const float *sc_w,
                    const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats, int stat_ld, int co0, int cout_total,
                    const float *r1_w, void *stream);
int l3d_conv3_tc_ex2(const l3d_act *x, const l3d_act *x2, const l3d_norm *xn, int N, int D, int H, int W,
                     const float *w, int groups, const float *dw_w, const float *pw_w, const float *sc_w,
                     const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats, int stat_ld, int co0, int cout_total,
                     const float *r1_w, void *stream);
int l3d_conv3_tc(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                 const float *w, int groups, const float *dw_w, const float *pw_w, const float *sc_w,
                 const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats, int stat_ld, int co0, int cout_total,
                 void *stream) {
    return l3d_conv3_tc_ex(x, xn, N, D, H, W, w, groups, dw_w, pw_w, sc_w, t, t_stats, r, r_stats, stat_ld, co0, cout_total, nullptr, stream);
}
int l3d_conv3_tc_ex(const l3d_act *x, const l3d_norm *xn, int N, int D, int H, int W,
                    const float *w, int groups, const float *dw_w, const float *pw_w, const float *sc_w,
                    const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats, int stat_ld, int co0, int cout_total,
                    const float *r1_w, void *stream) {
    return l3d_conv3_tc_ex2(x, nullptr, xn, N, D, H, W, w, groups, dw_w, pw_w, sc_w, t, t_stats, r, r_stats, stat_ld, co0, cout_total, r1_w, stream);
}
// x2 != NULL: the input is the channel concatenation [x | x2] of two DENSE 16-channel tensors (a decoder block reading
// [ConvTranspose output | skip] without an interleaved concat buffer): one channel chunk per tensor, each fetched through
// its own tensor map with the merged (C, W) box rows.
int l3d_conv3_tc_ex2(const l3d_act *x, const l3d_act *x2, const l3d_norm *xn, int N, int D, int H, int W,
                     const float *w, int groups, const float *dw_w, const float *pw_w, const float *sc_w,
                     const l3d_act *t, double *t_stats, const l3d_act *r, double *r_stats, int stat_ld, int co0, int cout_total,
                     const float *r1_w, void *stream) {
    const bool rank1 = r1_w != nullptr;
    if (L3D_ENV_INT("L3D_NO_IGEMM", 0) == 1) return -1;
    const bool split = x2 != nullptr;
    if (split && (rank1 || x->C != CK || x2->C != CK || x->ldc != CK || x2->ldc != CK || x2->dtype != x->dtype ||
                  reinterpret_cast<uintptr_t>(x2->ptr) % 16 != 0)) return -1;
    const int Cin = x->C + (split ? x2->C : 0), Cout = t->C;
    const bool has_sc = sc_w != nullptr;
    if (x->dtype != L3D_F16 || t->dtype != L3D_F16) return -1;
    if (Cin % 16 != 0 || Cout % 16 != 0 || Cout > 256) return -1;
    auto aligned = [](const l3d_act *a, int mult) {
        return (a->ldc % mult == 0) && (reinterpret_cast<uintptr_t>(a->ptr) % (2 * mult) == 0);
    };
    if (rank1 && (Cin != CK || W % 4 != 0 || reinterpret_cast<uintptr_t>(x->ptr) % 16 != 0 || 3 * Cout > 256)) return -1;
    if ((!rank1 && !aligned(x, 8)) || !aligned(t, 8) || (has_sc && (act_null(r) || !aligned(r, 8)))) return -1;
    // ---- tile height: the tallest tile (fewest halo planes and MMAs per voxel) whose accumulators fit TMEM and whose
    // buffers fit shared memory; two accumulator sets (epilogue of tile T under the MMAs of tile T+1) when they fit
    const int force_tz = L3D_ENV_INT("L3D_C3_TZ", 0), force_nraw = L3D_ENV_INT("L3D_C3_NRAW", 0), force_sets = L3D_ENV_INT("L3D_C3_SETS", 0);   // tuning / test knobs
    const int nacc = has_sc ? 2 : 1;
    int TZ = 0, nraw = 0, nsets = 0, nabuf = 2;
    // L3D_C3_OCC2: two CTAs per SM, each with ONE raw box and ONE operand tile (<= 112 KB of shared memory, <= 256 TMEM columns,
    // 8 worker warps): inside a CTA the box latency, the activation pass and the MMAs of a tile are then serial, and the two
    // CTAs fill each other's gaps
    const int occ2 = rank1 ? L3D_ENV_INT("L3D_C3_OCC2_R1", 0) : L3D_ENV_INT("L3D_C3_OCC2", 0);
    if (occ2)
        for (int tz : {6, 4, 2}) {
            if (force_tz && tz != force_tz) continue;
            if (!force_tz && tz > 2 && tz > D) continue;
            const int cols1 = tz * Cout * nacc;
            int ns = 2 * cols1 <= 256 ? 2 : 1;
            if (force_sets == 1 || force_sets == 2) ns = force_sets;
            if (ns * cols1 > 256) continue;
            if (c3_smem_bytes(tz, Cin, Cout, has_sc, 1, rank1, 1) + 1024 > 113 * 1024) continue;
            TZ = tz; nraw = 1; nsets = ns; nabuf = 1;
            break;
        }
    if (TZ == 0)
    for (int tz : {8, 6, 4, 2}) {
        if (force_tz && tz != force_tz) continue;
        if (!force_tz && tz > 2 && tz > D) continue;                 // no taller than the volume
        const int cols1 = tz * Cout * nacc;
        if (cols1 > 512) continue;
        int ns = 2 * cols1 <= 512 ? 2 : 1;
        if (force_sets == 1 || force_sets == 2) ns = force_sets;
        if (ns * cols1 > 512) continue;
        // raw TMA boxes in flight: a box has microseconds of latency under load, so as many as fit (up to 4)
        int nr = 1;
        for (int c = 4; c >= 1; --c) if (c3_smem_bytes(tz, Cin, Cout, has_sc, c, rank1) <= 226 * 1024) { nr = c; break; }
        if (force_nraw) nr = force_nraw;
        if (c3_smem_bytes(tz, Cin, Cout, has_sc, nr, rank1) > 226 * 1024) continue;
        // prefer a shorter tile with double-buffered accumulators over a taller single-buffered one -- except for single-chunk
        // layers (Cin = 16), whose MMA phase per tile is short: 16 -> 32 + shortcut at 24^3, 325 windows: TZ 8 / one set 442 us,
        // TZ 6 482 us, TZ 4 / two sets 541 us
        if (ns == 1 && tz > 2 && !force_tz && !force_sets && Cin > CK) {
            const int cols_half = (tz / 2) * Cout * nacc;
            if (2 * cols_half <= 512 && tz / 2 >= 2) continue;
        }
        TZ = tz; nraw = nr; nsets = ns;
        break;
    }
    if (TZ == 0) return -1;
    const size_t smem = c3_smem_bytes(TZ, Cin, Cout, has_sc, nraw, rank1, nabuf);
    const long long tiles = (long long)N * ((D + TZ - 1) / TZ) * ((H + TY - 1) / TY) * ((W + TX - 1) / TX);
    if (tiles >= (1ll << 30)) return -1;
    int cols = 32;
    while (cols < nsets * TZ * Cout * nacc) cols <<= 1;

    // TMA moves one request per innermost box row, and a 16-channel voxel is only 32 B: when the view is a whole
    // 16-channel tensor the (C, W) axes are contiguous and merge into one axis, so a box row is a 320-B x-row
    // (10 voxels) instead of ten 32-B rows (measured: the 5-D box is TMA-issue bound).
    int tma_split = L3D_ENV_INT("L3D_C3_TMASPLIT", 1);   // measured: no effect on the box latency (1, 2, 5 or 10 slices)
    if (tma_split < 1 || (TZ + 2) % tma_split != 0) tma_split = 1;
    // CTAs per SM, worker warps and staging mode (needed here: the producer-warp modes shape the TMA box)
    int occ = (int)((227 * 1024) / (smem + 1024));
    if (occ > 3) occ = 3;
    if (nabuf == 1 && occ > 2) occ = 2;
    if (occ < 1) occ = 1;
    if (occ * cols > 512) occ = 512 / cols;
    // 12 worker warps (3 per scheduler) hide the latency of the activation pass and the epilogue when one CTA owns the SM
    const int nwarps_env = L3D_ENV_INT("L3D_C3_WARPS", 0);
    const int nwarps = nwarps_env > 0 ? nwarps_env : (occ == 1 ? 12 : 8);
    // 1: where only one TMA box fits and a tile has >= 4 channel chunks (measured at 325 windows: 64 -> 32 + shortcut at 24^3
    // 1033 -> 935 us; slower than TMA on the one- and two-chunk 48^3 layers: 1035 -> 1327 us, 1897 -> 2455 us); 2: wherever
    // possible; 0: never
    const int ld_mode = L3D_ENV_INT("L3D_C3_LOADER", 1);
    const bool use_loader = !split && (ld_mode == 2 || (ld_mode == 1 && nraw == 1 && Cin / CK >= 4)) && (long long)(TZ + 2) * H * W * x->ldc * 2 < (1ll << 31);
    const bool rot_mode = !rank1 && nabuf == 2 && !(nwarps == 12 && use_loader) && nraw == 1 && tma_split == 1 && L3D_ENV_INT("L3D_C3_ROT", 0) != 0 && smem + 128 <= 226 * 1024;
    // the producer warp exists in the 12-worker TMA variants
    const bool prod_mode = nwarps == 12 && !(use_loader && !rank1) && !rot_mode && L3D_ENV_INT("L3D_C3_PROD", 1) != 0;
    const bool split2 = prod_mode && !rank1 && nraw == 1 && tma_split == 1 && (TZ + 2) % 2 == 0 && L3D_ENV_INT("L3D_C3_SPLIT2", 1) != 0;
    const cuuint32_t box_z = (cuuint32_t)(split2 ? (TZ + 2) / 2 : (TZ + 2) / tma_split);
    const bool merged_cx = split || (!rank1 && Cin == CK && x->ldc == CK && L3D_ENV_INT("L3D_C3_NOMERGECX", 0) == 0);
    CUtensorMap tmap, tmap2;
    if (rank1) {
        const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        const cuuint64_t rowb = (cuuint64_t)W * 4;
        const cuuint64_t strides[3] = {rowb, (cuuint64_t)H * rowb, (cuuint64_t)D * H * rowb};
        const cuuint32_t box[4] = {R1_BOXW, HY, box_z, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        if (l3d_encode_tiled(&tmap, (int)CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x->ptr, (const unsigned long long *)dims,
                             (const unsigned long long *)strides, (const unsigned *)box, (const unsigned *)estr)) return 3;
    } else if (merged_cx) {
        const cuuint64_t dims[4] = {(cuuint64_t)W * CK, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        const cuuint64_t rowb = (cuuint64_t)W * CK * 2;
        const cuuint64_t strides[3] = {rowb, (cuuint64_t)H * rowb, (cuuint64_t)D * H * rowb};
        const cuuint32_t box[4] = {HX * CK, HY, box_z, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        if (l3d_encode_tiled(&tmap, (int)CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x->ptr, (const unsigned long long *)dims,
                             (const unsigned long long *)strides, (const unsigned *)box, (const unsigned *)estr)) return 3;
        if (split && l3d_encode_tiled(&tmap2, (int)CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x2->ptr, (const unsigned long long *)dims,
                                      (const unsigned long long *)strides, (const unsigned *)box, (const unsigned *)estr)) return 3;
    } else {
        const cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        const cuuint64_t es = 2, ld = (cuuint64_t)x->ldc;
        const cuuint64_t strides[4] = {ld * es, (cuuint64_t)W * ld * es, (cuuint64_t)H * W * ld * es, (cuuint64_t)D * H * W * ld * es};
        const cuuint32_t box[5] = {CK, HX, HY, box_z, 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        if (l3d_encode_tiled(&tmap, (int)CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, x->ptr, (const unsigned long long *)dims,
                             (const unsigned long long *)strides, (const unsigned *)box, (const unsigned *)estr)) return 3;
    }
    if (!split) tmap2 = tmap;
    C3Args A;
    A.Cin = Cin; A.xn = norm_dev(xn);
    A.N = N; A.D = D; A.H = H; A.W = W;
    A.w = w; A.groups = w != nullptr ? groups : 1; A.dw_w = dw_w; A.pw_w = pw_w; A.sc_w = sc_w; A.Cout = Cout;
    A.t = (h16 *)t->ptr; A.ldt = t->ldc; A.t_stats = t_stats;
    A.r = has_sc ? (h16 *)r->ptr : nullptr; A.ldr = has_sc ? r->ldc : 0; A.r_stats = r_stats;
    A.stat_ld = stat_ld > 0 ? stat_ld : Cout;
    A.co0 = co0; A.cout_total = cout_total > 0 ? cout_total : Cout;
    A.tmem_cols = cols; A.nraw = nraw; A.nsets = nsets; A.merged_cx = merged_cx ? 1 : 0; A.dbg = L3D_ENV_INT("L3D_C3_DEBUG_SKIP", 0);
    A.r1_w = r1_w; A.tma_split = tma_split; A.xp = x->ptr; A.ldx = x->ldc; A.rot = 0;
    A.prod = prod_mode ? 1 : 0;
    A.split2 = split2 ? 1 : 0;
    A.nabuf = nabuf;
    A.merge = (3 * Cout <= 256 && L3D_ENV_INT("L3D_C3_NOMERGE", 0) == 0) ? 1 : 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = (long long)sms * occ;
    if (grid > tiles) grid = tiles;
    { const int g = L3D_ENV_INT("L3D_C3_GRID", 0); if (g > 0 && g < grid) grid = g; }   // development aid: fewer CTAs than SMs
    size_t smem_launch = smem;
#define L3D_C3_LAUNCH_L(TZV, MG, NWV, R1V, LDV)                                                                                     \
    do {                                                                                                                    \
        static bool attr_set_[64] = {};  /* the attribute is per device */                                                                                                                 \
        if (dev < 0 || dev >= 64 || !attr_set_[dev]) {                                                                                                    \
            cudaError_t e = cudaFuncSetAttribute(conv3_tc_kernel<TZV, MG, NWV, R1V, LDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); \
            if (e != cudaSuccess) { l3d_set_error("conv3_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 3; } \
            if (dev >= 0 && dev < 64) attr_set_[dev] = true;                                                                                                \
        }                                                                                                                   \
        conv3_tc_kernel<TZV, MG, NWV, R1V, LDV><<<(unsigned)grid, NWV * 32 + 32 + ((NWV) == 12 && !(LDV) ? 32 : 0), smem_launch, (cudaStream_t)stream>>>(tmap, tmap2, A); \
    } while (0)
    /* thread-private cp.async staging instead of TMA: with 12 worker warps (one CTA per SM), not for the rank-1 input */
#define L3D_C3_LAUNCH_R(TZV, MG, NWV, R1V)                                                                                          \
    do {                                                                                                                    \
        if (NWV == 12 && use_loader && !R1V) L3D_C3_LAUNCH_L(TZV, MG, 12, false, true);                                     \
        else L3D_C3_LAUNCH_L(TZV, MG, NWV, R1V, false);                                                                     \
    } while (0)
#define L3D_C3_LAUNCH(TZV, MG, NWV) L3D_C3_LAUNCH_R(TZV, MG, NWV, false)
#define L3D_C3_TZ(MG, NWV)                                \
    switch (TZ) {                                         \
        case 8: L3D_C3_LAUNCH(8, MG, NWV); break;         \
        case 6: L3D_C3_LAUNCH(6, MG, NWV); break;         \
        case 4: L3D_C3_LAUNCH(4, MG, NWV); break;         \
        default: L3D_C3_LAUNCH(2, MG, NWV); break;        \
    }
    // rotating slots where only one raw box fits (and the box is requested in one piece by TMA).  Off by default: measured
    // SLOWER at 325 windows (16 -> 16 at 48^3: 1101 -> 1187 us, 32 -> 16 + shortcut: 1912 -> 2195 us, 32 -> 32 at 24^3: 406 -> 446 us)
    // -- with the next box streaming in under the MMAs, the TMA writes, the operand reads of the tensor core and the
    // activation pass compete for the same 128 B / clock of shared-memory bandwidth, which is what really bounds the tile
    if (rot_mode) {
        A.rot = 1;
        smem_launch = smem + 128;
    }
    if (rank1) {
        // merged MMAs only (3 * Cout <= 256 checked above)
        if (nwarps == 12) {
            switch (TZ) {
                case 8: L3D_C3_LAUNCH_R(8, true, 12, true); break;
                case 6: L3D_C3_LAUNCH_R(6, true, 12, true); break;
                case 4: L3D_C3_LAUNCH_R(4, true, 12, true); break;
                default: L3D_C3_LAUNCH_R(2, true, 12, true); break;
            }
        } else {
            switch (TZ) {
                case 8: L3D_C3_LAUNCH_R(8, true, 8, true); break;
                case 6: L3D_C3_LAUNCH_R(6, true, 8, true); break;
                case 4: L3D_C3_LAUNCH_R(4, true, 8, true); break;
                default: L3D_C3_LAUNCH_R(2, true, 8, true); break;
            }
        }
    } else
    if (A.merge) { if (nwarps == 12) { L3D_C3_TZ(true, 12) } else { L3D_C3_TZ(true, 8) } }
    else         { if (nwarps == 12) { L3D_C3_TZ(false, 12) } else { L3D_C3_TZ(false, 8) } }
#undef L3D_C3_TZ
#undef L3D_C3_LAUNCH
#undef L3D_C3_LAUNCH_R
#undef L3D_C3_LAUNCH_L
    l3d_count_launch();
    l3d_note_kernel("conv3_tc_kernel");
    L3D_CUDA_OK("l3d_conv3 (tcgen05 implicit GEMM) launch");
    return 0;
}
