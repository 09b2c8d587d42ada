// Sliding-window gather / stitch, threshold, 6-connected labelling and bounding-box reduction.
//   utils.py:86-137 (window loop + normalisation), inferencer.py:62-109 (extract_bboxes),
//   metrics.py:50-61 (get_connected_components; scipy.ndimage.label semantics).
#include "l3d_common.cuh"

namespace {

// ---------------------------------------------------------------- windows ------------------
__device__ __forceinline__ void stv8(float *p, const float (&v)[8]) {
    reinterpret_cast<float4 *>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4 *>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void stv8(h16 *p, const float (&v)[8]) {
    const uint32_t a = pack_h16x2(v[0], v[1]), b = pack_h16x2(v[2], v[3]), c = pack_h16x2(v[4], v[5]), d = pack_h16x2(v[6], v[7]);
    *reinterpret_cast<uint4 *>(p) = make_uint4(a, b, c, d);
}

// grid.y = window; one thread = 8 x-consecutive voxels of one window row (32-bit index arithmetic only)
template <typename T>
__global__ void __launch_bounds__(256) gather_windows_kernel(const float *__restrict__ vol, int D, int H, int W,
                                                             const int32_t *__restrict__ pos, int nwin,
                                                             int pd, int ph, int pw, T *__restrict__ out) {
    const int wi = blockIdx.y;
    const int oz = pos[wi * 3], oy = pos[wi * 3 + 1], ox = pos[wi * 3 + 2];
    const int xg = (pw + 7) >> 3;                       // 8-voxel groups per row
    const int groups = pd * ph * xg;
    T *wout = out + (size_t)wi * pd * ph * pw;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
        const int x0 = (g % xg) * 8;
        const int row = g / xg;
        const int y = row % ph, z = row / ph;
        const int gz = oz + z, gy = oy + y;
        const bool rowok = gz < D && gy < H;
        const float *src = vol + ((size_t)(rowok ? gz : 0) * H + (rowok ? gy : 0)) * W + ox + x0;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (rowok && ox + x0 + j < W && x0 + j < pw) ? src[j] : 0.f;   // zero padding at the far end (utils.py:102-112)
        T *dst = wout + (size_t)row * pw + x0;
        if (x0 + 8 <= pw && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            stv8(dst, v);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (x0 + j < pw) st1(dst + j, v[j]);
        }
    }
}

// Per-voxel gather over the windows that cover it, in the reference's z -> y -> x window order.
// The multiply and the add are rounded separately (numpy computes `pred * weights` into a temporary
// and then `+=`), so the accumulation is bit-identical to utils.py:133-134 for identical predictions.
// Per-axis coverage tables (first covering window + count per coordinate; window starts ascend, utils.py:63-81) are
// built once per CTA in shared memory; a voxel's (<= STITCH_MAXC) candidates are all loaded before the ordered
// accumulation.  Axes with more than 3 covering windows per coordinate (overlap > 2/3) take the scanning path.
// Slab form (window-level sharding of one volume over several GPUs): the launch covers the voxels
// x in [g.x0, g.x1) only, the window (a, b, c) lives at preds[(a * g.wsz + b * g.wsy + c * g.wsx) * per] (so a rank can
// keep [x-position][z][y] blocks: received seam positions first, then its own), and voxel (z, y, x) is written to
// prob[(z * H + y) * g.pitch + g.xoff + (x - g.x0)] (rank 0 writes into the full map, the other ranks a packed slab).
// The candidate order per voxel stays z -> y -> x, so the result is bit-identical to the single-launch form.
constexpr int STITCH_MAXC = 27;
struct StitchGeo { long long wsz, wsy, wsx; int x0, x1, pitch, xoff; };
__global__ void __launch_bounds__(256) stitch_kernel(const float *__restrict__ preds,
                                                     const int32_t *__restrict__ zpos, int nz,
                                                     const int32_t *__restrict__ ypos, int ny,
                                                     const int32_t *__restrict__ xpos, int nx,
                                                     int pd, int ph, int pw, const float *__restrict__ imp,
                                                     int D, int H, int W, const uint8_t *__restrict__ body,
                                                     float *__restrict__ prob, float thr, int32_t *__restrict__ mask, int use_tables,
                                                     StitchGeo g) {
    extern __shared__ int32_t s_tab[];                  // [D + H + W] : first << 8 | count
    __shared__ int s_over;
    const int Wl = g.x1 - g.x0;
    const size_t total = (size_t)D * H * Wl;
    const size_t per = (size_t)pd * ph * pw;
    if (threadIdx.x == 0) s_over = 0;
    __syncthreads();
    if (use_tables) {
        for (int i = threadIdx.x; i < D + H + W; i += blockDim.x) {
            const int32_t *p; int np, ext, c;
            if (i < D) { p = zpos; np = nz; ext = pd; c = i; }
            else if (i < D + H) { p = ypos; np = ny; ext = ph; c = i - D; }
            else { p = xpos; np = nx; ext = pw; c = i - D - H; }
            int first = 0, cnt = 0;
            for (int a = 0; a < np; ++a) {
                const int l = c - p[a];
                if (l >= 0 && l < ext) { if (cnt == 0) first = a; ++cnt; }
            }
            s_tab[i] = (first << 8) | cnt;
            if (cnt > 3) s_over = 1;                    // more than 3 covering windows on an axis: scanning path (every CTA decides alike)
        }
        __syncthreads();
        if (s_over) use_tables = 0;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = g.x0 + (int)(i % (size_t)Wl);
        const int rem = (int)(i / (size_t)Wl);
        const int y = rem % H, z = rem / H;
        const size_t oi = (size_t)rem * g.pitch + g.xoff + (x - g.x0);     // output element; body mask is indexed in the full volume
        float acc = 0.f, cnt = 0.f;
        if (use_tables) {
            const int tz = s_tab[z], ty = s_tab[D + y], tx = s_tab[D + H + x];
            const int a0 = tz >> 8, na = tz & 255, b0 = ty >> 8, nb = ty & 255, c0 = tx >> 8, nc = tx & 255;
            float pv[STITCH_MAXC], wv[STITCH_MAXC];
#pragma unroll
            for (int k = 0; k < STITCH_MAXC; ++k) {
                const int ia = k / 9, ib = (k / 3) % 3, ic = k % 3;
                pv[k] = 0.f; wv[k] = 0.f;
                if (ia < na && ib < nb && ic < nc) {
                    const int a = a0 + ia, b = b0 + ib, c = c0 + ic;
                    const size_t li = ((size_t)(z - zpos[a]) * ph + (y - ypos[b])) * pw + (x - xpos[c]);
                    wv[k] = imp[li];
                    pv[k] = preds[(size_t)(a * g.wsz + b * g.wsy + c * g.wsx) * per + li];
                }
            }
#pragma unroll
            for (int k = 0; k < STITCH_MAXC; ++k) {
                const int ia = k / 9, ib = (k / 3) % 3, ic = k % 3;
                if (ia < na && ib < nb && ic < nc) {
                    acc = __fadd_rn(acc, __fmul_rn(pv[k], wv[k]));
                    cnt = __fadd_rn(cnt, wv[k]);
                }
            }
        } else {
            for (int a = 0; a < nz; ++a) {
                const int lz = z - zpos[a];
                if (lz < 0 || lz >= pd) continue;
                for (int b = 0; b < ny; ++b) {
                    const int ly = y - ypos[b];
                    if (ly < 0 || ly >= ph) continue;
                    for (int c = 0; c < nx; ++c) {
                        const int lx = x - xpos[c];
                        if (lx < 0 || lx >= pw) continue;
                        const size_t wi = (size_t)(a * g.wsz + b * g.wsy + c * g.wsx);
                        const size_t li = ((size_t)lz * ph + ly) * pw + lx;
                        const float wgt = imp[li];
                        acc = __fadd_rn(acc, __fmul_rn(preds[wi * per + li], wgt));
                        cnt = __fadd_rn(cnt, wgt);
                    }
                }
            }
        }
        float pr = cnt > 0.f ? __fdiv_rn(acc, cnt) : acc;   // np.divide(..., where=cnt>0) (utils.py:137)
        if (body != nullptr) pr = __fmul_rn(pr, body[(size_t)rem * W + x] ? 1.f : 0.f);  // inferencer.py:161-162
        prob[oi] = pr;
        if (mask != nullptr) mask[oi] = pr >= thr ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) threshold_kernel(const float *__restrict__ prob, int64_t n, float thr,
                                                        int32_t *__restrict__ mask) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        mask[i] = prob[i] >= thr ? 1 : 0;
}

// ------------------------------------------------------------------- CCL --------------------
// Union-find over linear voxel indices; the root of a set is always its smallest index, so ranking the
// roots by index reproduces scipy.ndimage.label's raster numbering.
// (flatten pass: the forest no longer changes shape -- only links are shortened -- so the hops may come from L1: every find
// of a large component ends on the same root line, which one L2 slice would otherwise serve to all SMs)
template <bool L1>
__device__ __forceinline__ int32_t uf_find(const int32_t *parent, int32_t i) {
    int32_t p = __ldcg(parent + i);
    while (p != i) {
        i = p;
        if (L1) asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(p) : "l"(parent + i) : "memory");
        else p = __ldcg(parent + i);
    }
    return i;
}
// find with path halving: every visited node is re-pointed at its grandparent (atomicMin keeps the links monotone
// under concurrent unions: a node only ever points at a smaller index of its own set), so the long row-over-row chains of
// a large component collapse while the merge pass is still running
// link loads of the merge pass.  L1: cached in L1 (ld.global.ca) -- a stale link is an older, larger member of the same
// chain, and uf_union only ever acts on a value that an atomicMin returned, so the forest stays valid; every find of a
// component of millions of voxels ends at the same root, and read at L2 (ld.global.cg) that one line is served by one slice
template <bool L1>
__device__ __forceinline__ int32_t uf_ld(const int32_t *p) {
    int32_t v;
    if (L1) asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <bool L1>
__device__ __forceinline__ int32_t uf_find_halving(int32_t *parent, int32_t i) {
    int32_t p = uf_ld<L1>(parent + i);
    while (p != i) {
        const int32_t gp = uf_ld<L1>(parent + p);
        if (gp != p) atomicMin(&parent[i], gp);
        i = p;
        p = gp;
    }
    return i;
}
template <bool L1>
__device__ __forceinline__ void uf_union(int32_t *parent, int32_t a, int32_t b) {
    while (true) {
        a = uf_find_halving<L1>(parent, a);
        b = uf_find_halving<L1>(parent, b);
        if (a == b) return;
        if (a < b) { const int32_t t = a; a = b; b = t; }   // a > b: hook a under b
        const int32_t old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;                                             // someone else hooked a first; retry with its new parent
    }
}

// Initial forest: one warp walks one row (32 voxels per step, the open run carried across steps), so every foreground
// voxel starts out pointing at the first voxel of its whole x-run -- a row of foreground costs no union at all.
__global__ void __launch_bounds__(256) ccl_init_kernel(const int32_t *__restrict__ mask, int32_t n, int W, int32_t *__restrict__ parent,
                                                       int32_t *__restrict__ size) {
    const int lane = threadIdx.x & 31;
    const int32_t nrows = n / W;
    const int32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (int32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < nrows; row += warps) {
        const int32_t base = row * W;
        bool carry = false;
        int32_t carry_start = 0;
        for (int x0 = 0; x0 < W; x0 += 32) {
            const int x = x0 + lane;
            const bool fg = x < W && mask[base + x] != 0;
            const unsigned fgm = __ballot_sync(0xffffffffu, fg);
            // run starts inside this step: foreground whose left neighbour is background (lane 0: no open run carried in)
            const unsigned starts = fgm & ~((fgm << 1) | (carry ? 1u : 0u));
            const unsigned below = starts & (0xffffffffu >> (31 - lane));
            const int32_t my_start = below ? base + x0 + (31 - __clz(below)) : carry_start;
            if (x < W) {
                parent[base + x] = fg ? my_start : -1;
                size[base + x] = 0;
            }
            carry = (fgm >> 31) != 0;
            carry_start = __shfl_sync(0xffffffffu, my_start, 31);
        }
    }
}
// Unions only where two runs meet for the first time: for the y / z neighbour only at the first voxel of an overlap
// (if the previous voxel in x and its y / z neighbour are both foreground, that voxel has already joined the same two
// runs).
template <bool L1>
__global__ void __launch_bounds__(256) ccl_merge_kernel(int32_t *__restrict__ parent, int D, int H, int W) {
    const int32_t n = D * H * W;
    const int32_t WH = W * H;
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (parent[i] < 0) continue;
        const int x = i % W, y = (i / W) % H, z = i / WH;
        const bool left = x > 0 && parent[i - 1] >= 0;
        if (y > 0 && parent[i - W] >= 0 && !(left && parent[i - W - 1] >= 0)) uf_union<L1>(parent, i, i - W);
        if (z > 0 && parent[i - WH] >= 0 && !(left && parent[i - WH - 1] >= 0)) uf_union<L1>(parent, i, i - WH);
    }
}
template <bool L1>
__global__ void __launch_bounds__(256) ccl_flatten_count_kernel(int32_t *__restrict__ parent, int32_t n, int32_t *__restrict__ size) {
    // warp-uniform trip count so that the lanes can aggregate their size increments per root (a large
    // component would otherwise serialise millions of atomics on one address)
    // ... and the CTA keeps the count of ONE root (the first it meets) in shared memory: a component of millions of voxels
    // otherwise sends one global atomic per warp step to a single address (163 K of them for a 128 x 128 x 320 map that is
    // mostly one component), which the L2 serialises
    __shared__ int32_t s_hot, s_cnt;
    if (threadIdx.x == 0) { s_hot = -1; s_cnt = 0; }
    __syncthreads();
    const int32_t n_round = (n + 31) & ~31;
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        int32_t r = -1;
        if (i < n && parent[i] >= 0) {
            r = uf_find<L1>(parent, i);
            // roots keep parent[r] == r; non-roots may point at the root directly (no thread still needs the chain:
            // every chain ends at r and a concurrent reader following a shortened link still arrives at r)
            if (r != i) parent[i] = r;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, r);
        if (r >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) {
            int32_t hot = *reinterpret_cast<volatile int32_t *>(&s_hot);
            if (hot == -1) { const int32_t old = atomicCAS(&s_hot, -1, r); hot = old == -1 ? r : old; }
            if (hot == r) atomicAdd(&s_cnt, __popc(peers));
            else atomicAdd(&size[r], __popc(peers));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_hot >= 0 && s_cnt > 0) atomicAdd(&size[s_hot], s_cnt);
}
// flag[i] = 1 iff voxel i is the root of a surviving component
__global__ void __launch_bounds__(256) ccl_flag_kernel(const int32_t *__restrict__ parent, const int32_t *__restrict__ size,
                                                       int32_t n, int min_size, int32_t *__restrict__ flag) {
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        flag[i] = (parent[i] == i && size[i] >= min_size) ? 1 : 0;
}
// three-kernel exclusive scan of flag -> rank (int32), SCAN_BLOCK elements per CTA
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;
__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp, int &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < (SCAN_THREADS / 32) ? s_warp[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        if (lane < SCAN_THREADS / 32) s_warp[lane] = winc - w;  // exclusive per-warp offsets
        if (lane == SCAN_THREADS / 32 - 1) s_warp[SCAN_THREADS / 32] = winc;
    }
    __syncthreads();
    total = s_warp[SCAN_THREADS / 32];
    const int res = s_warp[wid] + inc - v;
    __syncthreads();
    return res;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const int32_t *__restrict__ flag, int32_t n, int32_t *__restrict__ bsum) {
    __shared__ int s_warp[SCAN_THREADS / 32 + 1];
    const int32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    int v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) if (base + k < n) v += flag[base + k];
    int total;
    block_exclusive_scan(v, s_warp, total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}
// single CTA: exclusive scan of the block sums in place; writes the grand total to *n_out
__global__ void __launch_bounds__(SCAN_THREADS) scan_top_kernel(int32_t *__restrict__ bsum, int32_t nb, int32_t *__restrict__ n_out) {
    __shared__ int s_warp[SCAN_THREADS / 32 + 1];
    int carry = 0;
    for (int32_t start = 0; start < nb; start += SCAN_THREADS) {
        const int32_t i = start + threadIdx.x;
        const int v = i < nb ? bsum[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, s_warp, total);
        if (i < nb) bsum[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *n_out = carry;
}
// rank[i] = exclusive prefix of flag (only meaningful where flag[i] == 1); written over `flag`
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(int32_t *__restrict__ flag, int32_t n, const int32_t *__restrict__ bsum) {
    __shared__ int s_warp[SCAN_THREADS / 32 + 1];
    const int32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    int f[SCAN_ITEMS];
    int v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { f[k] = (base + k < n) ? flag[base + k] : 0; v += f[k]; }
    int total;
    int run = bsum[blockIdx.x] + block_exclusive_scan(v, s_warp, total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) flag[base + k] = f[k] ? run + 1 : 0;   // 1-based id at surviving roots, 0 elsewhere
        run += f[k];
    }
}
__global__ void __launch_bounds__(256) ccl_relabel_kernel(const int32_t *__restrict__ parent, const int32_t *__restrict__ rootid,
                                                          int32_t n, int32_t *__restrict__ labels) {
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int32_t p = parent[i];
        labels[i] = p < 0 ? 0 : rootid[p];
    }
}

// ------------------------------------------------------------------ bbox --------------------
__global__ void bbox_init_kernel(int32_t *__restrict__ table, int cap) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        int32_t *t = table + (size_t)i * 8;
        t[0] = t[2] = t[4] = 0x7fffffff;
        t[1] = t[3] = t[5] = -1;
        t[6] = 0;
        t[7] = 0;  // bits of +0.0f; probabilities are >= 0 so the int ordering equals the float ordering
    }
}
// One thread walks BB_RUN consecutive voxels and keeps a private accumulator while the label does not change;
// accumulators are merged across the warp per label (match.any + redux) before touching the table, so a large
// component costs ~1/256th of the per-voxel atomics.
constexpr int BB_RUN = 8;
struct BoxAcc {
    int32_t l, z0, z1, y0, y1, x0, x1, cnt, pb;
};
// A CTA keeps a small direct-mapped cache of label -> box accumulators in shared memory: a warp merges its lanes per
// label (match.any + redux), then folds the result into the cache slot of that label; a slot holding another label is
// first written back to the global table.  A component of millions of voxels therefore costs a handful of global
// atomics per CTA instead of one set per warp run.
constexpr int BB_SLOTS = 64;
struct BoxCache {
    int32_t key[BB_SLOTS];
    int32_t v[BB_SLOTS][8];
};
__device__ __forceinline__ void box_writeback(const int32_t *v, int32_t key, int32_t *__restrict__ table) {
    int32_t *t = table + (size_t)(key - 1) * 8;
    atomicMin(&t[0], v[0]); atomicMax(&t[1], v[1]);
    atomicMin(&t[2], v[2]); atomicMax(&t[3], v[3]);
    atomicMin(&t[4], v[4]); atomicMax(&t[5], v[5]);
    atomicAdd(&t[6], v[6]);
    atomicMax(&t[7], v[7]);
}
__device__ __forceinline__ void box_flush_warp(const BoxAcc &a, int32_t *__restrict__ table, int cap, BoxCache *cache) {
    const int32_t key = (a.l > 0 && a.l <= cap) ? a.l : 0;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int z0 = __reduce_min_sync(peers, a.z0), z1 = __reduce_max_sync(peers, a.z1);
    const int y0 = __reduce_min_sync(peers, a.y0), y1 = __reduce_max_sync(peers, a.y1);
    const int x0 = __reduce_min_sync(peers, a.x0), x1 = __reduce_max_sync(peers, a.x1);
    const int cnt = __reduce_add_sync(peers, a.cnt), pb = __reduce_max_sync(peers, a.pb);
    if (key > 0 && (threadIdx.x & 31) == __ffs(peers) - 1) {
        const int slot = key & (BB_SLOTS - 1);
        // claim the slot for `key` (0 = empty); another warp may hold it for a different label
        while (true) {
            const int32_t cur = atomicCAS(&cache->key[slot], 0, key);
            if (cur == 0 || cur == key) {
                int32_t *v = cache->v[slot];
                atomicMin(&v[0], z0); atomicMax(&v[1], z1);
                atomicMin(&v[2], y0); atomicMax(&v[3], y1);
                atomicMin(&v[4], x0); atomicMax(&v[5], x1);
                atomicAdd(&v[6], cnt);
                atomicMax(&v[7], pb);
                break;
            }
            // occupied by another label: bypass the cache for this update (rare: labels colliding modulo BB_SLOTS inside one CTA)
            const int32_t tmp[8] = {z0, z1, y0, y1, x0, x1, cnt, pb};
            box_writeback(tmp, key, table);
            break;
        }
    }
}
__global__ void __launch_bounds__(256) bbox_reduce_kernel(const int32_t *__restrict__ labels, const float *__restrict__ prob,
                                                          int D, int H, int W, int32_t *__restrict__ table, int cap) {
    __shared__ BoxCache cache;
    for (int i = threadIdx.x; i < BB_SLOTS; i += blockDim.x) {
        cache.key[i] = 0;
        cache.v[i][0] = cache.v[i][2] = cache.v[i][4] = 0x7fffffff;
        cache.v[i][1] = cache.v[i][3] = cache.v[i][5] = -1;
        cache.v[i][6] = 0; cache.v[i][7] = 0;
    }
    __syncthreads();
    const int64_t n = (int64_t)D * H * W;
    const int64_t nruns = (n + BB_RUN - 1) / BB_RUN;
    const int64_t nruns_round = (nruns + 31) & ~(int64_t)31;
    // contiguous range of runs per CTA: its voxels mostly share a few labels, which the cache then absorbs
    const int64_t per_cta = ((nruns_round / 32 + gridDim.x - 1) / gridDim.x) * 32;
    const int64_t run_begin = (int64_t)blockIdx.x * per_cta, run_end = min(nruns_round, run_begin + per_cta);
    for (int64_t run = run_begin + threadIdx.x; run < run_end; run += blockDim.x) {
        BoxAcc a;
        a.l = 0; a.z0 = a.y0 = a.x0 = 0x7fffffff; a.z1 = a.y1 = a.x1 = -1; a.cnt = 0; a.pb = 0;
        const int64_t i0 = run * BB_RUN;
        for (int k = 0; k < BB_RUN; ++k) {           // same trip count on every lane: the flush is warp-collective
            const int64_t i = i0 + k;
            const int32_t l = (i < n) ? labels[i] : 0;
            const bool change = (l != a.l) && a.cnt > 0;
            if (__any_sync(0xffffffffu, change)) {
                BoxAcc f = a;
                if (!change) { f.l = 0; }
                box_flush_warp(f, table, cap, &cache);
                if (change) { a.z0 = a.y0 = a.x0 = 0x7fffffff; a.z1 = a.y1 = a.x1 = -1; a.cnt = 0; a.pb = 0; }
            }
            a.l = l;
            if (l > 0) {
                const int x = (int)(i % W), y = (int)((i / W) % H), z = (int)(i / ((int64_t)W * H));
                a.z0 = min(a.z0, z); a.z1 = max(a.z1, z);
                a.y0 = min(a.y0, y); a.y1 = max(a.y1, y);
                a.x0 = min(a.x0, x); a.x1 = max(a.x1, x);
                a.cnt += 1;
                a.pb = max(a.pb, __float_as_int(fmaxf(prob[i], 0.f)));
            }
        }
        if (a.cnt == 0) a.l = 0;
        box_flush_warp(a, table, cap, &cache);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BB_SLOTS; i += blockDim.x)
        if (cache.key[i] > 0 && cache.v[i][6] > 0) box_writeback(cache.v[i], cache.key[i], table);
}

static unsigned grid_for(int64_t n, int threads, int cap_blocks) {
    int64_t b = (n + threads - 1) / threads;
    if (b > cap_blocks) b = cap_blocks;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace

extern "C" int l3d_gather_windows(const float *vol, int D, int H, int W, const int32_t *pos, int nwin,
                                  int pd, int ph, int pw, void *out, int dtype, void *stream) {
    L3D_REQUIRE(vol && pos && out && nwin > 0 && pd > 0 && ph > 0 && pw > 0, "l3d_gather_windows: bad argument");
    L3D_REQUIRE(nwin <= 65535, "l3d_gather_windows: at most 65535 windows per call");
    const int64_t groups = (int64_t)pd * ph * ((pw + 7) / 8);
    L3D_REQUIRE(groups < (1ll << 31), "l3d_gather_windows: window too large");
    dim3 grid(grid_for(groups, 256, 64), (unsigned)nwin);
    L3D_DISPATCH_DTYPE(dtype, T, {
        gather_windows_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(vol, D, H, W, pos, nwin, pd, ph, pw, (T *)out);
    });
    l3d_count_launch();
    L3D_CUDA_OK("l3d_gather_windows launch");
    return 0;
}

extern "C" int l3d_stitch(const float *preds, const int32_t *zpos, int nz, const int32_t *ypos, int ny,
                          const int32_t *xpos, int nx, int pd, int ph, int pw, const float *importance,
                          int D, int H, int W, const uint8_t *body_mask, float *prob,
                          float threshold, int32_t *mask_out, void *stream) {
    return l3d_stitch_slab(preds, zpos, nz, ypos, ny, xpos, nx, (int64_t)ny * nx, nx, 1, pd, ph, pw, importance, D, H, W, 0, W, W, 0,
                           body_mask, prob, threshold, mask_out, stream);
}

extern "C" int l3d_stitch_slab(const float *preds, const int32_t *zpos, int nz, const int32_t *ypos, int ny,
                               const int32_t *xpos, int nx, int64_t wstride_z, int64_t wstride_y, int64_t wstride_x,
                               int pd, int ph, int pw, const float *importance,
                               int D, int H, int W, int x0, int x1, int out_pitch, int out_xoff,
                               const uint8_t *body_mask, float *prob, float threshold, int32_t *mask_out, void *stream) {
    L3D_REQUIRE(preds && zpos && ypos && xpos && importance && prob, "l3d_stitch: null argument");
    L3D_REQUIRE(nz > 0 && ny > 0 && nx > 0 && D > 0 && H > 0 && W > 0, "l3d_stitch: bad dims");
    L3D_REQUIRE(0 <= x0 && x0 < x1 && x1 <= W && out_xoff >= 0 && out_pitch >= out_xoff + (x1 - x0), "l3d_stitch: bad slab [%d, %d) of W = %d (pitch %d, offset %d)", x0, x1, W, out_pitch, out_xoff);
    L3D_REQUIRE((int64_t)D * H < (1ll << 31), "l3d_stitch: volume too large");
    StitchGeo g;
    g.wsz = wstride_z; g.wsy = wstride_y; g.wsx = wstride_x; g.x0 = x0; g.x1 = x1; g.pitch = out_pitch; g.xoff = out_xoff;
    const unsigned blocks = grid_for((int64_t)D * H * (x1 - x0), 256, 148 * 8);
    // per-axis coverage tables in shared memory when they fit (the kernel itself falls back to scanning the window lists
    // when some coordinate is covered by more than 3 windows of an axis, i.e. overlap > 2/3)
    const size_t tab_bytes = sizeof(int32_t) * ((size_t)D + H + W);
    const int use_tables = (nz < (1 << 23) && ny < (1 << 23) && nx < (1 << 23) && tab_bytes <= 40 * 1024) ? 1 : 0;
    stitch_kernel<<<blocks, 256, use_tables ? tab_bytes : 0, (cudaStream_t)stream>>>(preds, zpos, nz, ypos, ny, xpos, nx, pd, ph, pw, importance,
                                                                                   D, H, W, body_mask, prob, threshold, mask_out, use_tables, g);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_stitch launch");
    return 0;
}

extern "C" int l3d_threshold(const float *prob, int64_t n, float threshold, int32_t *mask, void *stream) {
    L3D_REQUIRE(prob && mask && n >= 0, "l3d_threshold: null argument");
    if (n == 0) return 0;
    threshold_kernel<<<grid_for(n, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(prob, n, threshold, mask);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_threshold launch");
    return 0;
}

extern "C" int64_t l3d_ccl_workspace_elems(int64_t nvox) {
    const int64_t nb = (nvox + SCAN_BLOCK - 1) / SCAN_BLOCK;
    return 3 * nvox + nb + 16;   // parent, size, flag/rootid, block sums
}

extern "C" int l3d_ccl_label(const int32_t *mask, int D, int H, int W, int min_size, int32_t *labels,
                             int32_t *n_out, int32_t *work, void *stream) {
    L3D_REQUIRE(mask && labels && n_out && work, "l3d_ccl_label: null argument");
    const int64_t n64 = (int64_t)D * H * W;
    L3D_REQUIRE(n64 > 0 && n64 < (1ll << 31) - SCAN_BLOCK, "l3d_ccl_label: volume too large for int32 indices");
    const int32_t n = (int32_t)n64;
    const int32_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    int32_t *parent = work, *size = work + n64, *flag = work + 2 * n64, *bsum = work + 3 * n64;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned g = grid_for(n, 256, 148 * 32);
    ccl_init_kernel<<<g, 256, 0, st>>>(mask, n, W, parent, size);
    if (L3D_ENV_INT("L3D_CCL_L1", 1) != 0) ccl_merge_kernel<true><<<g, 256, 0, st>>>(parent, D, H, W);
    else ccl_merge_kernel<false><<<g, 256, 0, st>>>(parent, D, H, W);
    if (L3D_ENV_INT("L3D_CCL_L1", 1) != 0) ccl_flatten_count_kernel<true><<<g, 256, 0, st>>>(parent, n, size);
    else ccl_flatten_count_kernel<false><<<g, 256, 0, st>>>(parent, n, size);
    // metrics.py:52-58 drops components with size < min_size only when min_size > 0; with min_size <= 0 every
    // component (size >= 1) survives, which `size >= min_size` also yields.
    ccl_flag_kernel<<<g, 256, 0, st>>>(parent, size, n, min_size, flag);
    scan_block_sums_kernel<<<nb, SCAN_THREADS, 0, st>>>(flag, n, bsum);
    scan_top_kernel<<<1, SCAN_THREADS, 0, st>>>(bsum, nb, n_out);
    scan_apply_kernel<<<nb, SCAN_THREADS, 0, st>>>(flag, n, bsum);
    ccl_relabel_kernel<<<g, 256, 0, st>>>(parent, flag, n, labels);
    l3d_count_launch(8);
    L3D_CUDA_OK("l3d_ccl_label launch");
    return 0;
}

extern "C" int l3d_bbox_init(int32_t *table, int cap, void *stream) {
    L3D_REQUIRE(table && cap > 0, "l3d_bbox_init: bad argument");
    bbox_init_kernel<<<grid_for(cap, 256, 1024), 256, 0, (cudaStream_t)stream>>>(table, cap);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_bbox_init launch");
    return 0;
}

extern "C" int l3d_bbox_reduce(const int32_t *labels, const float *prob, int D, int H, int W, int32_t *table, int cap,
                               void *stream) {
    L3D_REQUIRE(labels && prob && table && cap > 0, "l3d_bbox_reduce: bad argument");
    const int64_t n = (int64_t)D * H * W;
    L3D_REQUIRE(n > 0 && n < (1ll << 31), "l3d_bbox_reduce: volume too large");
    bbox_reduce_kernel<<<grid_for((n + BB_RUN - 1) / BB_RUN, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(labels, prob, D, H, W, table, cap);
    l3d_count_launch();
    L3D_CUDA_OK("l3d_bbox_reduce launch");
    return 0;
}

// ------------------------------------------------------- lesion matching statistics ---------
// Everything match_components (metrics.py:127-229) and calculate_metrics (:311-404) need from the volumes, in one
// pass over two label maps: the pair-intersection histogram (np.bincount of pred_id * (nb + 1) + target_id, :153-160),
// the component sizes (:162-163) and the first moments of the voxel coordinates per component (ndimage.center_of_mass
// with unit weights, :111-124).  All sums are integers, so the host-side arithmetic on them (IoU in float32, centres and
// distances in float64, the greedy matching) reproduces the reference bit for bit.  Background-background voxels -- the
// bulk of a PET volume -- touch no counter.
namespace {
__global__ void __launch_bounds__(256) label_pair_stats_kernel(const int32_t *__restrict__ la, const int32_t *__restrict__ lb,
                                                               int D, int H, int W, int na, int nb, int32_t *__restrict__ counts,
                                                               unsigned long long *__restrict__ mom_a, unsigned long long *__restrict__ mom_b) {
    const int64_t total = (int64_t)D * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = la[i], b = lb != nullptr ? lb[i] : 0;
        if (a == 0 && b == 0) continue;
        const int x = (int)(i % W);
        const int64_t rem = i / W;
        const int y = (int)(rem % H), z = (int)(rem / H);
        if (a > 0 && a <= na) {
            atomicAdd(&mom_a[(size_t)a * 4 + 0], 1ull); atomicAdd(&mom_a[(size_t)a * 4 + 1], (unsigned long long)z);
            atomicAdd(&mom_a[(size_t)a * 4 + 2], (unsigned long long)y); atomicAdd(&mom_a[(size_t)a * 4 + 3], (unsigned long long)x);
        }
        if (b > 0 && b <= nb && mom_b != nullptr) {
            atomicAdd(&mom_b[(size_t)b * 4 + 0], 1ull); atomicAdd(&mom_b[(size_t)b * 4 + 1], (unsigned long long)z);
            atomicAdd(&mom_b[(size_t)b * 4 + 2], (unsigned long long)y); atomicAdd(&mom_b[(size_t)b * 4 + 3], (unsigned long long)x);
        }
        if (counts != nullptr && a > 0 && a <= na && b > 0 && b <= nb) atomicAdd(&counts[(size_t)a * (nb + 1) + b], 1);
    }
}
}  // namespace

extern "C" int l3d_label_pair_stats(const int32_t *labels_a, const int32_t *labels_b, int D, int H, int W, int na, int nb,
                                    int32_t *counts, int64_t *mom_a, int64_t *mom_b, void *stream) {
    L3D_REQUIRE(labels_a && mom_a && na >= 0 && nb >= 0, "l3d_label_pair_stats: bad argument");
    L3D_REQUIRE(labels_b != nullptr || (counts == nullptr && mom_b == nullptr), "l3d_label_pair_stats: counts / mom_b need labels_b");
    const int64_t n = (int64_t)D * H * W;
    L3D_REQUIRE(n > 0, "l3d_label_pair_stats: empty volume");
    label_pair_stats_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(labels_a, labels_b, D, H, W, na, nb, counts,
                                                                                          reinterpret_cast<unsigned long long *>(mom_a),
                                                                                          reinterpret_cast<unsigned long long *>(mom_b));
    l3d_count_launch();
    L3D_CUDA_OK("l3d_label_pair_stats launch");
    return 0;
}
