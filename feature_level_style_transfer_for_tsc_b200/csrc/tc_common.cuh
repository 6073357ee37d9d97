// sm_100a primitives used by the tcgen05 engine: mbarrier, TMA (tensor + bulk), TMEM, UMMA descriptors.
// Written as inline PTX (no CUTLASS dependency).  Conventions follow the PTX ISA "tcgen05" chapter;
// field layouts of the shared-memory / instruction descriptors are the ones CUTLASS documents in
// cute/arch/mma_sm100_desc.hpp (read for the bit positions only).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace tsc {
namespace tc {

// watchdog word (one per translation unit): the first timed-out mbarrier wait stores
// (code << 16 | blockIdx) here; read back through tsc_debug_read_and_clear_watchdog().
static __device__ int g_watchdog = 0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait.  `dead` is sticky per thread: after one timeout every later wait returns at once, so a
// broken pipeline drains instead of hanging the GPU (results are then garbage and g_watchdog != 0).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, bool& dead, int code) {
    if (dead) return;
    for (uint32_t i = 0; i < (1u << 22); ++i) {
        if (mbar_try_wait(bar, parity)) return;
    }
    dead = true;
    atomicCAS(&g_watchdog, 0, (code << 16) | (int)(blockIdx.x & 0xffff));
}
// As mbar_wait, for a thread that waits LONG next to the MMA issuer (accumulator poller, weight producer): between two polls it
// sleeps `ns` nanoseconds instead of re-issuing the try_wait at once.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, bool& dead, int code, uint32_t ns) {
    if (dead) return;
    for (uint32_t i = 0; i < (1u << 20); ++i) {
        if (mbar_try_wait(bar, parity)) return;
        __nanosleep(ns);
    }
    dead = true;
    atomicCAS(&g_watchdog, 0, (code << 16) | (int)(blockIdx.x & 0xffff));
}

// ---- TMA ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// contiguous global -> shared bulk copy (bytes % 16 == 0, both addresses 16 B aligned)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- LDGSTS (cp.async) staging of c8 tiles ---------------------------------------------------------
// A c8 row is 16 B and consecutive positions are contiguous, so a warp's 32 copies cover 512 contiguous bytes; one
// producer warpgroup stages a tile an order of magnitude faster than a TMA box whose inner extent is a single
// 16 B row (measured: the TMA unit retires ~1 such row per cycle -- profiles/r1_conv_timeline.md).
__device__ __forceinline__ void cp_async16_zfill(uint32_t smem_dst, const void* gsrc, bool valid) {
    const uint32_t n = valid ? 16u : 0u;       // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(n) : "memory");
}
// the calling thread's earlier cp.async copies arrive on the mbarrier when they have landed (does not add to the
// barrier's expected count: initialise it with the number of producer threads)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Stage rows [l_start, l_start + R) of chunks [chunk0, chunk0 + nchunk) of sample b of a c8 bf16 tensor
// [B][C8][L][8] into smem [nchunk][R][8]; rows outside [0, L) and chunks outside [0, C8) are zero-filled
// (ConstantPad1d / channel padding).  R >= nthreads is required (one chunk wrap per step at most).
__device__ __forceinline__ void stage_c8_tile(uint8_t* smem_dst, const __nv_bfloat16* base, int b, int C8, int L, int chunk0,
                                              int nchunk, int l_start, int R, int tid, int nthreads) {
    const uint32_t dst0 = smem_u32(smem_dst);
    int c = 0, r = tid;
    while (r >= R) { r -= R; ++c; }
    const int total = nchunk * R;
    for (int i = tid; i < total; i += nthreads) {
        const int l = l_start + r, ch = chunk0 + c;
        const bool valid = l >= 0 && l < L && ch >= 0 && ch < C8;
        const __nv_bfloat16* src = base + (((size_t)b * C8 + (valid ? ch : 0)) * L + (valid ? l : 0)) * 8;
        cp_async16_zfill(dst0 + (uint32_t)i * 16u, src, valid);
        r += nthreads;
        while (r >= R) { r -= R; ++c; }
    }
}

// ---- TMEM -----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 bit, 16 consecutive columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- UMMA descriptors -----------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layout: 8x(16 B) core matrices
// stored as 128 contiguous bytes.
//   K-major operand : LBO = byte distance between core matrices adjacent in K,
//                     SBO = byte distance between core matrices adjacent in M/N.
//   MN-major operand: SBO = byte distance between 8-element groups along M/N,
//                     LBO = byte distance between groups of 8 along K.
// bits [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1, [61,64) layout = 0.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// Instruction descriptor for kind::f16 with BF16 operands and FP32 accumulation, M = 128.
// bits [4,6) c_format=1 (F32), [7,10) a_format=1 (BF16), [10,13) b_format=1 (BF16), 13/14 negate a/b,
// 15 a_major (0 = K, 1 = MN), 16 b_major, [17,23) N>>3, [24,29) M>>4.
__device__ __forceinline__ uint32_t make_idesc_bf16(uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major, bool neg_a) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= 1u << 7;
    d |= 1u << 10;
    d |= (neg_a ? 1u : 0u) << 13;
    d |= (a_mn_major ? 1u : 0u) << 15;
    d |= (b_mn_major ? 1u : 0u) << 16;
    d |= (N >> 3) << 17;
    d |= (M >> 4) << 24;
    return d;
}
// Same for kind::tf32 (fp32 data in shared memory read as TF32): a/b format = 2.
__device__ __forceinline__ uint32_t make_idesc_tf32(uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major, bool neg_a) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= 2u << 7;
    d |= 2u << 10;
    d |= (neg_a ? 1u : 0u) << 13;
    d |= (a_mn_major ? 1u : 0u) << 15;
    d |= (b_mn_major ? 1u : 0u) << 16;
    d |= (N >> 3) << 17;
    d |= (M >> 4) << 24;
    return d;
}
// D[tmem] (+)= A[smem] * B[smem]; one thread issues for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// host: cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

}  // namespace tc
}  // namespace tsc
