// tcgen05 engine, forward / dgrad: the omni-scale convolution as an implicit GEMM on 5th-gen tensor cores.
//
//   positions (128 per CTA)  -> MMA M   (accumulator rows = TMEM lanes)
//   output channels (<= 256) -> MMA N   (the whole channel axis is one accumulator tile in TMEM)
//   (tap, input channel)     -> MMA K   (16 channels per instruction)
//
// * The activation tile with its halo, [kc][128 + Kmax - 1 rows][8 ch] bf16, is staged ONCE per CTA by
//   TMA from the c8 tensor (4-D tensor map, out-of-bounds rows zero-filled = ConstantPad1d,
//   OS_CNN.py:59,70).  In the SWIZZLE_NONE canonical layout a convolution tap is just "+ t rows" on the
//   A-operand descriptor's start address, so all Kmax taps reuse the same shared-memory tile.
// * The packed kernel bank streams through a ring of shared-memory stages with 1-D bulk copies; only
//   live (channel, tap) pairs exist in HBM, and each tap issues MMAs over its live channel suffix only
//   (N = np - n_lo[t], written at TMEM column n_lo[t]) -- the 43 %-dense bank costs 43 % of the FLOPs.
// * Warp roles: warp 0 = copy producer, warp 1 = MMA issuer (one elected thread) + TMEM owner,
//   warps 2-5 = epilogue (TMEM -> registers -> +bias -> coalesced c8 fp32 stores).
// Replaces ConstantPad1d + Conv1d (+ cuDNN dgrad), OS_CNN/OS_CNN.py:70-71; arithmetic SURVEY A1.
#include "tc_common.cuh"

namespace tsc {
namespace tc {

struct ConvTcParams {
    const __nv_bfloat16* w;
    const float* bias;
    float* y;
    int nbias, B, L, ltiles;
    int Rp;            // halo rows in shared memory (multiple of 8)
    int KB;            // input-channel chunks per weight stage (even)
    int NS;            // weight stages
    int stage_bytes;
    int xs_bytes;
    int tmem_cols;
    long long* tl;     // optional phase timeline of CTA 0 (tsc_debug_set_timeline), NULL in production
};

#define TL(i) do { if (p.tl && blockIdx.x == 0) p.tl[i] = clock64(); } while (0)
#define TLS(n, k) do { if (p.tl && blockIdx.x == 0 && (n) < 48) p.tl[8 + 8 * (n) + (k)] = clock64(); } while (0)

static constexpr int TC_THREADS = 192;
static constexpr int SMEM_TAPINFO = 256;  // per-tap issue constants, TSC_MAX_TAPS x 16 B
static constexpr int SMEM_HDR = 256 + TSC_MAX_TAPS * 16;     // barriers + tmem slot + tap info

__global__ void __launch_bounds__(TC_THREADS, 1)
osconv_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ TapTable tt, const ConvTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);            // [8]
    uint64_t* empty = full + 8;                                    // [8]
    uint64_t* x_full = empty + 8;
    uint64_t* acc_full = x_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    uint8_t* xs = smem + SMEM_HDR;
    uint8_t* stages = xs + p.xs_bytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / p.ltiles, l0 = (blockIdx.x % p.ltiles) * 128;
    const int np = tt.np, kc = tt.kc;

    if (warp == 0 && lane == 0) {
        TL(0);
        tma_prefetch_desc(&xmap);
        for (int i = 0; i < p.NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(x_full, 1);
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== copy producer =====
        if (lane == 0) {
            bool dead = false;
            TL(1);
            mbar_arrive_expect_tx(x_full, (uint32_t)(kc * p.Rp * 16));
            for (int kcI = 0; kcI < kc; ++kcI)
                tma_load_4d(xs + (size_t)kcI * p.Rp * 16, &xmap, 0, l0 - tt.pad_left, kcI, b, x_full);
            uint32_t s = 0, ph = 0;
            int sn = 0;
            for (int oi = 0; oi < tt.n_order; ++oi) {
                const int t = tt.order[oi];
                const int nt = np - tt.n_lo[t], kspan = kc - tt.kc_lo[t];
                const __nv_bfloat16* blob = p.w + (size_t)tt.w_off[t] * 8;
                for (int g0 = 0; g0 < kspan; g0 += p.KB) {
                    TLS(sn, 4);
                    mbar_wait(&empty[s], ph ^ 1u, dead, 1);
                    TLS(sn, 5);
                    const int nch = min(p.KB, kspan - g0);
                    const uint32_t bytes = (uint32_t)(nch * nt * 16);
                    mbar_arrive_expect_tx(&full[s], bytes);
                    bulk_load(stages + (size_t)s * p.stage_bytes, blob + (size_t)g0 * nt * 8, bytes, &full[s]);
                    TLS(sn, 6);
                    ++sn;
                    if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // A tcgen05.mma with M = 128 occupies the tensor pipe for N/2 cycles, but one thread cannot issue faster
        // than one every ~54 cycles (measured, tools/mma_bench2.cu), so the issue loop is kept lean: per-tap
        // constants are precomputed by the whole warp into shared memory, descriptors advance by adding to
        // their low word, and the stage ring is tracked without divisions.
        uint32_t* tapinfo = reinterpret_cast<uint32_t*>(smem + SMEM_TAPINFO);     // [n_order][4]
        for (int oi = lane; oi < tt.n_order; oi += 32) {
            const int t = tt.order[oi];
            const int n_lo = tt.n_lo[t], kc_lo = tt.kc_lo[t];
            const int nt = np - n_lo;
            tapinfo[oi * 4 + 0] = (uint32_t)(kc_lo * p.Rp + t);                    // A row offset (16 B units)
            tapinfo[oi * 4 + 1] = (uint32_t)nt | ((uint32_t)(kc - kc_lo) << 16);    // N, k span (chunks)
            tapinfo[oi * 4 + 2] = make_idesc_bf16(128, (uint32_t)nt, false, false, false);
            tapinfo[oi * 4 + 3] = tmem_base + (uint32_t)n_lo;
        }
        __syncwarp();
        if (lane == 0) {
            bool dead = false;
            mbar_wait(x_full, 0, dead, 2);
            tc_fence_after();
            TL(2);
            const uint32_t desc_hi = (128u >> 4) | (1u << 14);                      // SBO = 128 B, descriptor version 1
            const uint32_t a_base16 = (smem_u32(xs) >> 4) | ((uint32_t)p.Rp << 16);  // LBO = Rp * 16 B
            const uint32_t st_base16 = smem_u32(stages) >> 4;
            const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4;
            const uint32_t a_step = 2u * (uint32_t)p.Rp;                             // 16 channels = 2 chunks
            uint32_t s = 0, ph = 0, acc = 0;
            int sn = 0;
            const int n_order = tt.n_order, KB = p.KB;
            for (int oi = 0; oi < n_order; ++oi) {
                const uint4 ti = *reinterpret_cast<const uint4*>(tapinfo + oi * 4);
                const uint32_t nt = ti.y & 0xffffu;
                const int kspan = (int)(ti.y >> 16);
                const uint32_t b_step = 2u * nt;
                uint32_t a_lo = a_base16 + ti.x;
                for (int g0 = 0; g0 < kspan; g0 += KB) {
                    TLS(sn, 0);
                    mbar_wait(&full[s], ph, dead, 3);
                    tc_fence_after();
                    TLS(sn, 1);
                    if (oi == 0 && g0 == 0) TL(3);
                    const int nsteps = min(KB, kspan - g0) >> 1;
                    uint32_t b_lo = (st_base16 + s * stage16) | (nt << 16);          // LBO = nt * 16 B
#pragma unroll 5
                    for (int k = 0; k < nsteps; ++k) {
                        umma_bf16(ti.w, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, ti.z, acc);
                        acc = 1;
                        a_lo += a_step;
                        b_lo += b_step;
                    }
                    TLS(sn, 2);
                    tc_commit(&empty[s]);      // frees the stage when these MMAs have read it
                    TLS(sn, 3);
                    ++sn;
                    if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1u; }
                }
            }
            tc_commit(acc_full);
            TL(4);
        }
    } else {
        // ===== epilogue: TMEM -> registers -> (+bias) -> c8 fp32 =====
        bool dead = false;
        mbar_wait(acc_full, 0, dead, 4);
        tc_fence_after();
        if (warp == 2 && lane == 0) TL(5);
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int l = l0 + row;
        const int npc = np / 8;
        for (int c0 = 0; c0 < np; c0 += 16) {
            float v[16];
            tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (p.bias) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (c0 + i < p.nbias) v[i] += __ldg(p.bias + c0 + i);
            }
            if (l < p.L) {
                float* d0 = p.y + (((size_t)b * npc + (c0 >> 3)) * p.L + l) * 8;
                float* d1 = d0 + (size_t)p.L * 8;
                *reinterpret_cast<float4*>(d0) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(d0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
                *reinterpret_cast<float4*>(d1) = make_float4(v[8], v[9], v[10], v[11]);
                *reinterpret_cast<float4*>(d1 + 4) = make_float4(v[12], v[13], v[14], v[15]);
            }
        }
        if (warp == 2 && lane == 0) TL(6);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    if (warp == 1 && lane == 0) TL(7);
}

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)ptr;
    return fn;
}

// c8 bf16 tensor [B][kc][L][8] as a 4-D tensor map with a (8, rows, 1, 1) box, no swizzle, zero OOB fill.
int make_c8_map(CUtensorMap* map, const void* base, int B, int kc, int L, int box_rows) {
    EncodeTiledFn enc = get_encode_tiled();
    TSC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    TSC_REQUIRE(((uintptr_t)base & 15) == 0, "c8 tensor must be 16-byte aligned");
    TSC_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows %d outside [1,256]", box_rows);
    cuuint64_t dims[4] = {8, (cuuint64_t)L, (cuuint64_t)kc, (cuuint64_t)B};
    cuuint64_t strides[3] = {16, (cuuint64_t)L * 16, (cuuint64_t)kc * L * 16};
    cuuint32_t box[4] = {8, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TSC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

static constexpr int SMEM_CAP = 227 * 1024;
static long long* g_timeline = nullptr;     // debug only: device buffer of >= 8 clock64 samples

}  // namespace tc

void set_conv_timeline(long long* dev) { tc::g_timeline = dev; }

int osconv_tc(int direction, const void* x, int dtype, const void* w, const float* bias, float* y, int B, int L, int Cin,
              int Cout, int Kmax, const int* s_of_tap, cudaStream_t cs) {
    using namespace tc;
    TSC_REQUIRE(dtype == TSC_BF16, "the tcgen05 engine takes bf16 operands");
    TapTable tt;
    if (build_tap_table(direction, Cin, Cout, Kmax, s_of_tap, &tt) != 0) return -1;
    ConvTcParams p;
    p.w = (const __nv_bfloat16*)w;
    p.bias = direction == TSC_DIR_FWD ? bias : nullptr;
    p.nbias = direction == TSC_DIR_FWD ? Cout : 0;
    p.y = y;
    p.B = B; p.L = L; p.ltiles = cdiv(L, 128);
    p.Rp = (128 + Kmax - 1 + 7) & ~7;
    p.xs_bytes = tt.kc * p.Rp * 16;
    int kb = (40 * 1024 / (tt.np * 16)) & ~1;
    if (kb > tt.kc) kb = tt.kc;
    if (kb < 2) kb = 2;
    p.KB = kb;
    p.stage_bytes = kb * tt.np * 16;
    int ns = (SMEM_CAP - SMEM_HDR - p.xs_bytes) / p.stage_bytes;
    if (ns > 8) ns = 8;
    TSC_REQUIRE(ns >= 2, "shape needs %d B of shared memory for the activation tile: unsupported", p.xs_bytes);
    p.NS = ns;
    p.tl = g_timeline;
    p.tmem_cols = tt.np <= 32 ? 32 : tt.np <= 64 ? 64 : tt.np <= 128 ? 128 : 256;
    CUtensorMap xmap;
    if (make_c8_map(&xmap, x, B, tt.kc, L, p.Rp) != 0) return -1;
    const int smem = SMEM_HDR + p.xs_bytes + ns * p.stage_bytes;
    static bool attr_set = false;          // once per process: opt in to the full 227 KB of dynamic shared memory
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(osconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = true;
    }
    osconv_tc_kernel<<<B * p.ltiles, TC_THREADS, smem, cs>>>(xmap, tt, p);
    TSC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tsc

namespace tsc {
int read_clear_watchdog_conv(int* code) {
    int zero = 0;
    cudaError_t e = cudaMemcpyFromSymbol(code, tc::g_watchdog, sizeof(int));
    if (e != cudaSuccess) return (int)e;
    return (int)cudaMemcpyToSymbol(tc::g_watchdog, &zero, sizeof(int));
}
}  // namespace tsc
