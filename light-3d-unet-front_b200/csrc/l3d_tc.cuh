// tcgen05 / TMEM / mbarrier PTX wrappers for sm_100a (hand-written; no CUTLASS).
//
// Operand layout used throughout: K-major, no swizzle ("interleaved" canonical layout).  A 16-bit operand tile
// [rows x K] is stored as 8x8-element core matrices (8 rows x 16 bytes, rows 16 B apart); core matrices that
// follow each other along the row axis are SBO = 128 B apart, the ones that follow along K are LBO = rows*16 B
// apart:   byte_offset(r, k) = (k/8)*rows*16 + (r/8)*128 + (r%8)*16 + (k%8)*2
// One tcgen05.mma.kind::f16 consumes K = 16 (two core matrices along K); the descriptor start address advances by
// 2*LBO per K step.  Accumulators live in TMEM: lane = row of the 128-row tile, column = output channel.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM allocation (one full warp executes these) --------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make st.shared writes visible to the async proxy (tcgen05.mma reads its operands through it)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one lane of a converged warp (elect.sync): the idiom ptxas recognises as "single thread active"
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
// arrive on `bar` when every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// ---- TMA: 5-D tiled load global -> shared, completion signalled on `bar` (complete_tx::bytes).  Coordinates are
// innermost-first and may lie outside the tensor (those elements are zero-filled). ------------------------------
__device__ __forceinline__ void tma_load_5d(void *dst_smem, const void *tmap, uint64_t *bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst_smem)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}

__device__ __forceinline__ void tma_load_4d(void *dst_smem, const void *tmap, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst_smem)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// ---- descriptors -----------------------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell); layout type 0 = no swizzle, base offset 0
    return d;
}
// kind::f16, fp16 A and B (K-major), fp32 accumulate, M = 128
__host__ __device__ constexpr uint32_t idesc_f16_m128(int n) {
    return (1u << 4) /* D = f32 */ | (0u << 7) /* A = f16 */ | (0u << 10) /* B = f16 */ | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
// kind::f16 with selectable 16-bit formats (0 = fp16, 1 = bf16) and operand majors (MN-major: the M / N index is the
// contiguous one inside each 16-byte chunk), fp32 accumulate, M = 128, K = 16 per instruction
__host__ __device__ constexpr uint32_t idesc_16b_m128(int n, int a_fmt, int b_fmt, bool a_mn, bool b_mn) {
    return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// kind::tf32: fp32 words in shared memory (the tensor core reads the top 19 bits), fp32 accumulate, M = 128, K = 8 per
// instruction.  a_mn / b_mn select MN-major operands (bits 15 / 16): the operand's M (or N) index is the contiguous
// one inside each 16-byte chunk, as in a tile stored [channel group][voxel][4 channels] and reduced over voxels.
__host__ __device__ constexpr uint32_t idesc_tf32_m128(int n, bool a_mn, bool b_mn) {
    return (1u << 4) /* D = f32 */ | (2u << 7) /* A = tf32 */ | (2u << 10) /* B = tf32 */ | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// byte offset of element (row r, column k) of a [rows x K] 32-bit tile in the planar layout [k/4][row][4]:
// K-major operand (rows = M or N): LBO = rows*16 (next 4 K), SBO = 128 (next 8 rows), K = 8 per MMA = 2 LBO steps;
// MN-major operand over the same bytes (MN = k index here, K = row index): SBO = rows*16, LBO = 128.
__device__ __forceinline__ uint32_t tile_off32(int r, int k, int rows) {
    return (uint32_t)((k >> 2) * rows * 16 + r * 16 + (k & 3) * 4);
}

// ---- operand tile addressing -------------------------------------------------------------------
// byte offset of element (row r, column k) of a [rows x K] 16-bit tile
__device__ __forceinline__ uint32_t tile_off(int r, int k, int rows) {
    return (uint32_t)((k >> 3) * rows * 16 + (r >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

// ---- TMEM -> registers: 16 consecutive fp32 columns of this thread's lane (row) ----------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Split form: issue the load, do other work, then wait.  The wait names the destination registers as read-write operands, so
// no use of them can be scheduled ahead of it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}

}  // namespace tc
