// tcgen05 engine, forward / dgrad: the omni-scale convolution as an implicit GEMM on 5th-gen tensor cores.
//
//   positions (128 per CTA)  -> MMA M   (accumulator rows = TMEM lanes)
//   output channels (<= 256) -> MMA N   (the whole channel axis is one accumulator tile in TMEM)
//   (tap, input channel)     -> MMA K   (16 channels per instruction)
//
// * The activation tile with its halo, [kc][128 + Kmax - 1 rows][8 ch] bf16, is staged ONCE per CTA by one
//   TMA load from the c8 tensor (4-D tensor map, out-of-bounds rows zero-filled = ConstantPad1d,
//   OS_CNN.py:59,70).  In the SWIZZLE_NONE canonical layout a convolution tap is just "+ t rows" on the
//   A-operand descriptor's start address, so all Kmax taps reuse the same shared-memory tile.
// * The packed kernel bank streams through a ring of shared-memory stages with 1-D bulk copies; only
//   live (channel, tap) pairs exist in HBM, and each tap issues MMAs over its live channel suffix only
//   (N = np - n_lo[t], written at TMEM column n_lo[t]) -- the 43 %-dense bank costs 43 % of the FLOPs.
// * The whole issue sequence is a precomputed PLAN (tsc_osconv_plan_build, host, once per bank geometry):
//   one 16 B entry per MMA (A offset, B offset + leading-dimension field, instruction descriptor, TMEM
//   column, stage flags) and one 8 B entry per weight stage.  A stage spans several taps (up to 32 KB), so the
//   issuing thread pays one mbarrier wait and one commit per ~10-40 MMAs and otherwise only adds two
//   integers per instruction (measured before the plan: ~745 cycles per tap of which <= 120 were tensor work,
//   profiles/r1_conv_timeline.md).
// * Warp roles: warp 0 = copy producer, warp 1 = MMA issuer (one elected thread) + TMEM owner,
//   warps 2-5 = epilogue: tcgen05.ld (32 columns at a time) -> +bias (staged in shared memory) -> coalesced
//   c8 fp32 stores; optionally the per-CTA BatchNorm partial statistics (forward) or the ReLU mask of the
//   layer below plus the partial sums of its BatchNorm backward (dgrad), reduced over the 32 rows of a warp
//   with a transposing butterfly (31 shuffles per 32 columns).
// Replaces ConstantPad1d + Conv1d (+ cuDNN dgrad), OS_CNN/OS_CNN.py:70-71; arithmetic SURVEY A1/A2.
#include "tc_common.cuh"
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace tsc {
namespace tc {

static constexpr int TC_THREADS = 192;
static constexpr int PLAN_STAGE_BYTES = 32 * 1024;               // weight stage; 16 KB for banks whose activation tile is large
static constexpr int ZERO_BLOCK_BYTES = 512;                      // behind every stage slot: the B operand of padding MMAs
static constexpr int MMA_GROUP = 4;                               // MMAs issued per elect block
static constexpr uint32_t PF_FIRST = 1u << 16, PF_LAST = 1u << 17;      // flags in the w word of a plan entry
static constexpr int PF_STAGE_SHIFT = 18, PF_STAGE_MAX = 1 << 14;         // w >> 18: index of the entry's weight stage
static constexpr int SMEM_HDR = 256;                       // barriers + tmem slot

struct ConvTcParams {
    const __nv_bfloat16* w;
    const uint8_t* plan;
    const float* bias;
    float* y;
    // fused epilogues (all nullable)
    float* stat_partial;       // FWD : [n_cta][np] float2 (mean, M2) over the CTA's valid rows
    const float* mask_y;       // DGRAD: pre-BN output of the layer below, c8 fp32 [B][np/8][L][8]
    const float* mask_scale;   //        z = scale*y + shift; d = dz * [z > 0]   (NULL = no ReLU)
    const float* mask_shift;
    const float* mask_mean;    //        yhat = (y - mean) * invstd
    const float* mask_invstd;
    float* red_partial;        // DGRAD: [n_cta][np] float2 (S1, S2) partial sums over the CTA's valid rows
    // FWD, inference: z = act(aff_scale * (acc + bias) + aff_shift [+ aff_res]) written to aff_out (kind aff_kind) instead of y
    const float* aff_scale;    //        ready-made coefficients, or NULL: derived in the prologue from the BatchNorm tensors below
    const float* aff_shift;
    const float* bn_gamma; const float* bn_beta; const float* bn_mean; const float* bn_var; float bn_eps;
    const float* aff_res;      //        c8 fp32 [B][np/8][L][8] added before the activation (the shortcut branch), or NULL
    void* aff_out;
    int aff_kind, aff_relu;
    int nbias, B, L, ltiles;
    int np;
    int Rp;            // halo rows in shared memory (multiple of 8)
    int kc;
    int pad_left;
    int NS;            // weight stages
    int stage_bytes;   // bytes of one weight stage (the slot is stage_bytes + ZERO_BLOCK_BYTES)
    int plan_bytes;
    int n_stages, n_mma;
    int off_bias, off_wstat, off_plan, off_xs, off_stages;
    int tmem_cols;
    long long* tl;     // optional phase timeline of CTA 0 (tsc_debug_set_timeline), NULL in production
    int debug;         // experiments only (TSC_CONV_DEBUG; garbage results): 2 = the issuer issues no MMAs (pure weight-stream
                       // rate), 4 = the producer signals its stages full without copying (pure issue rate), 8 = keep the
                       // tcgen05 fence after every stage wait (the pre-session-3 behaviour)
};

// DBG is the kernel's template flag: the production instantiation carries no timeline / experiment code at all (the MMA
// issuer is bound by its own instruction count -- ncu source page, profiles/README.md session 4)
#define TL(i) do { if (DBG && p.tl && blockIdx.x == 0) p.tl[i] = clock64(); } while (0)
// per-stage samples of CTA 0 (debug timeline only): slot k of weight stage i, for the first 48 stages
#define TLS(i, k) do { if (DBG && p.tl && blockIdx.x == 0 && (i) < 48) p.tl[8 + (i) * 8 + (k)] = clock64(); } while (0)

// 32 lanes x 32 bit, 32 consecutive columns
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// Column sums over the 32 lanes of a warp for 32 columns held one per register: a transposing butterfly.
// On return lane j holds the sum over all lanes of v[j] (31 shuffles instead of 160).
__device__ __forceinline__ float warp_colsum32(float* v, int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int j = 0; j < o; ++j) {
            const float send = up ? v[j] : v[j + o];
            const float keep = up ? v[j + o] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

// AFF: the inference instantiation (eval-mode BatchNorm folded into the epilogue).  A separate instantiation so that the
// training kernel keeps its register budget (125 per thread: two CTAs of different launches share an SM).
template <bool AFF, bool DBG>
__global__ void __launch_bounds__(TC_THREADS, AFF ? 2 : 1)
osconv_tc_kernel(const __grid_constant__ CUtensorMap xmap, const ConvTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);            // [8]
    uint64_t* empty = full + 8;                                    // [8]
    uint64_t* x_full = empty + 8;
    uint64_t* acc_full = x_full + 1;
    uint64_t* plan_full = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(plan_full + 1);
    float* bias_s = reinterpret_cast<float*>(smem + p.off_bias);
    float2* wstat = reinterpret_cast<float2*>(smem + p.off_wstat);  // [4][np]
    const uint8_t* plan_s = smem + p.off_plan;
    uint8_t* xs = smem + p.off_xs;
    uint8_t* stages = smem + p.off_stages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / p.ltiles, l0 = (blockIdx.x % p.ltiles) * 128;
    const int np = p.np;

    // experiment (TSC_CONV_DEBUG & 16, instrumented instantiation only): every CTA records %globaltimer at entry and exit and
    // its SM id at tl[1024 + 4 * blockIdx.x ..] -- where a launch's wall time goes beyond the lifetime of one CTA
    if (DBG && (p.debug & 16) && p.tl && threadIdx.x == 0) {
        unsigned long long t; unsigned int sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        p.tl[1024 + 4 * blockIdx.x] = (long long)t;
        p.tl[1024 + 4 * blockIdx.x + 2] = (long long)sm;
        p.tl[1024 + 4 * blockIdx.x + 3] = clock64();
    }
    if (warp == 0 && lane == 0) {
        TL(0);
        tma_prefetch_desc(&xmap);
        for (int i = 0; i < p.NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(x_full, 1);
        mbar_init(acc_full, 1);
        mbar_init(plan_full, 1);
        fence_barrier_init();
        pdl_wait();               // everything above overlapped the tail of the previous kernel of the stream
        mbar_arrive_expect_tx(plan_full, (uint32_t)p.plan_bytes);
        bulk_load(smem + p.off_plan, p.plan, (uint32_t)p.plan_bytes, plan_full);
        mbar_arrive_expect_tx(x_full, (uint32_t)(p.kc * p.Rp * 16));
        tma_load_4d(xs, &xmap, 0, l0 - p.pad_left, 0, b, x_full);
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    if (warp >= 2) {
        pdl_wait();
        // zero block at the end of every stage slot (read by the padding MMAs that round a stage up to MMA_GROUP)
        for (int i = threadIdx.x - 64; i < p.NS * 32; i += 128)
            *reinterpret_cast<uint4*>(stages + (size_t)(i >> 5) * (p.stage_bytes + ZERO_BLOCK_BYTES) + p.stage_bytes + (i & 31) * 16) =
                make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();
        // per-channel epilogue constants -> shared memory: [0] bias | mask scale, [1] mask shift, [2] mean, [3] invstd
        for (int c = threadIdx.x - 64; c < np; c += 128) {
            if (p.red_partial) {
                bias_s[c] = p.mask_scale ? __ldg(p.mask_scale + c) : 0.f;        // no ReLU below: z = 0*y + 1 > 0
                bias_s[np + c] = p.mask_scale ? __ldg(p.mask_shift + c) : 1.f;
                bias_s[2 * np + c] = __ldg(p.mask_mean + c);
                bias_s[3 * np + c] = __ldg(p.mask_invstd + c);
            } else {
                bias_s[c] = (p.bias && c < p.nbias) ? __ldg(p.bias + c) : 0.f;
                if (AFF) {
                    if (p.aff_scale) {
                        bias_s[np + c] = __ldg(p.aff_scale + c);
                        bias_s[2 * np + c] = __ldg(p.aff_shift + c);
                    } else if (c < p.nbias) {
                        // eval-mode BatchNorm1d (OS_CNN.py:72 in .eval()): the arithmetic of bn_eval_coeffs_kernel
                        const float sc = __ldg(p.bn_gamma + c) * (1.f / sqrtf(__ldg(p.bn_var + c) + p.bn_eps));
                        bias_s[np + c] = sc;
                        bias_s[2 * np + c] = __ldg(p.bn_beta + c) - __ldg(p.bn_mean + c) * sc;
                    } else {
                        bias_s[np + c] = 0.f;
                        bias_s[2 * np + c] = 0.f;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== copy producer: one bulk copy per plan stage =====
        if (lane == 0) {
            bool dead = false;
            pdl_wait();
            TL(1);
            mbar_wait(plan_full, 0, dead, 8);
            const uint2* st = reinterpret_cast<const uint2*>(plan_s + 16);
            uint32_t s = 0, ph = 0;
            const int n_stages = p.n_stages;
            for (int i = 0; i < n_stages; ++i) {
                const uint2 e = st[i];                        // {source offset in 16 B units, bytes}
                TLS(i, 4);
                mbar_wait(&empty[s], ph ^ 1u, dead, 1);
                TLS(i, 5);
                if (DBG && (p.debug & 4)) { mbar_arrive(&full[s]); if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1u; } continue; }
                mbar_arrive_expect_tx(&full[s], e.y);
                bulk_load(stages + (size_t)s * (p.stage_bytes + ZERO_BLOCK_BYTES), reinterpret_cast<const uint8_t*>(p.w) + (size_t)e.x * 16,
                          e.y, &full[s]);
                TLS(i, 6);
                if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the plan in lock-step (uniform control flow, entries prefetched one
        // iteration ahead); one elected lane issues.  Measured (tools/mma_bench3.cu): 73 cycles per MMA against 105-145
        // for a lane-0-only branch, whose UTCHMMA ptxas wraps in a per-thread ELECT loop. =====
        bool dead = false;
        mbar_wait(plan_full, 0, dead, 9);
        mbar_wait(x_full, 0, dead, 2);
        __syncwarp();             // lanes leave the polling loops at different times: reconverge before every elect
        tc_fence_after();
        if (lane == 0) TL(2);
        uint4* mm = reinterpret_cast<uint4*>(smem + p.off_plan + 16 + ((p.n_stages * 8 + 15) & ~15));
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);                        // SBO = 128 B, descriptor version 1
        const uint32_t a_base16 = (smem_u32(xs) >> 4) | ((uint32_t)p.Rp << 16);    // LBO = Rp * 16 B
        const uint32_t st_base16 = smem_u32(stages) >> 4;
        const uint32_t stage16 = (uint32_t)(p.stage_bytes + ZERO_BLOCK_BYTES) >> 4;
        // Relocate the plan once, one issue group per lane: A / B descriptor words and the TMEM address become final
        // values and the stage flags move to one byte per group, so that the issuing thread only moves four words per
        // MMA to uniform registers (every instruction of its elect block costs ~6 cycles).
        uint8_t* gflag = reinterpret_cast<uint8_t*>(mm + p.n_mma);
        const int n_grp = p.n_mma / MMA_GROUP;
        for (int g = lane; g < n_grp; g += 32) {
            uint32_t fl = 0;
#pragma unroll
            for (int j = 0; j < MMA_GROUP; ++j) {
                uint4 e = mm[g * MMA_GROUP + j];
                const uint32_t slot = (e.w >> PF_STAGE_SHIFT) % (uint32_t)p.NS;
                if (j == 0 && (e.w & PF_FIRST)) fl |= 1u;
                if (j == MMA_GROUP - 1 && (e.w & PF_LAST)) fl |= 2u;
                e.x += a_base16;
                e.y += st_base16 + slot * stage16;
                e.w = tmem_base + (e.w & 0xffffu);
                mm[g * MMA_GROUP + j] = e;
            }
            gflag[g] = (uint8_t)fl;
        }
        __syncwarp();
        uint32_t s = 0, ph = 0, acc = 0;
        int stage_i = 0;
        uint4 e0 = mm[0], e1 = mm[1], e2 = mm[2], e3 = mm[3];
        uint4 f0, f1, f2, f3;
        uint32_t fl = gflag[0], fln;
        // One issue group: prefetch the next group's entries (the read past the last group lands in the flag bytes / the
        // activation tile: in bounds, never issued), wait for the group's weight stage if it opens one, issue, commit.
        // The loop is unrolled by two with the register sets swapped, so the prefetched entries are never copied: the
        // issuing warp is bound by its own instruction count (118 warp instructions per group before, ncu source page).
#define TSC_ISSUE_GROUP(E0, E1, E2, E3, FL, N0, N1, N2, N3, FLN, G)                                                  \
        {                                                                                                            \
            const uint4* nx = mm + (size_t)((G) + 1) * MMA_GROUP;                                                    \
            N0 = nx[0]; N1 = nx[1]; N2 = nx[2]; N3 = nx[3];                                                          \
            FLN = gflag[(G) + 1];                                                                                    \
            if (FL & 1u) {                                                                                           \
                if (lane == 0) TLS(stage_i, 0);                                                                      \
                /* The stage's bytes were written by the async proxy and published through the mbarrier's complete_tx: \
                   tcgen05.mma may read them without a tcgen05 fence (with the fence the issuer paid ~240 cycles per   \
                   already-complete stage). */                                                                       \
                if (!mbar_test_wait(&full[s], ph)) mbar_wait(&full[s], ph, dead, 3);                                 \
                __syncwarp();     /* lanes leave the polling loop at different times: reconverge before the elect */ \
                if (DBG && (p.debug & 8)) tc_fence_after();                                                          \
                if ((G) == 0 && lane == 0) TL(3);                                                                    \
                if (lane == 0) TLS(stage_i, 1);                                                                      \
            }                                                                                                        \
            const bool last = (FL & 2u) != 0;                                                                        \
            if (DBG && (p.debug & 2)) {                                                                              \
                if (last && elect_one()) tc_commit(&empty[s]);                                                       \
            } else if (elect_one()) {                                                                                \
                umma_bf16(E0.w, ((uint64_t)desc_hi << 32) | E0.x, ((uint64_t)desc_hi << 32) | E0.y, E0.z, acc);      \
                umma_bf16(E1.w, ((uint64_t)desc_hi << 32) | E1.x, ((uint64_t)desc_hi << 32) | E1.y, E1.z, 1u);       \
                umma_bf16(E2.w, ((uint64_t)desc_hi << 32) | E2.x, ((uint64_t)desc_hi << 32) | E2.y, E2.z, 1u);       \
                umma_bf16(E3.w, ((uint64_t)desc_hi << 32) | E3.x, ((uint64_t)desc_hi << 32) | E3.y, E3.z, 1u);       \
                if (last) tc_commit(&empty[s]);      /* frees the stage when these MMAs have read it */               \
            }                                                                                                        \
            acc = 1;                                                                                                 \
            if (last) {                                                                                              \
                if (lane == 0) TLS(stage_i, 2);                                                                      \
                if (DBG) ++stage_i;                                                                                  \
                if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1u; }                                                      \
            }                                                                                                        \
        }
        int g = 0;
        for (; g + 1 < n_grp; g += 2) {
            TSC_ISSUE_GROUP(e0, e1, e2, e3, fl, f0, f1, f2, f3, fln, g)
            TSC_ISSUE_GROUP(f0, f1, f2, f3, fln, e0, e1, e2, e3, fl, g + 1)
        }
        if (g < n_grp) TSC_ISSUE_GROUP(e0, e1, e2, e3, fl, f0, f1, f2, f3, fln, g)
#undef TSC_ISSUE_GROUP
        __syncwarp();
        if (elect_one()) tc_commit(acc_full);
        pdl_trigger();            // the next kernel of the stream may start its prologue while the epilogue runs
        if (lane == 0) TL(4);
    } else {
        // ===== epilogue: TMEM -> registers -> (+bias | mask) -> c8 fp32 (+ per-CTA partial reductions) =====
        bool dead = false;
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int l = l0 + row;
        const bool valid = l < p.L;
        const int npc = np / 8;
        const size_t row_off = ((size_t)b * npc * p.L + (size_t)(valid ? l : 0)) * 8;
        const size_t chunk_stride = (size_t)p.L * 8;
        float* ybase = p.y + row_off;
        const bool do_stat = p.stat_partial != nullptr;
        const bool do_red = p.red_partial != nullptr;
        constexpr bool do_aff = AFF;
        const float* mbase = do_red ? p.mask_y + row_off : nullptr;
        // one thread polls the accumulator barrier; the other 127 sleep in a named barrier instead of spinning on
        // mbarrier.try_wait next to the MMA issuer
        if (threadIdx.x == 64) mbar_wait(acc_full, 0, dead, 4);
        asm volatile("bar.sync 2, 128;" ::: "memory");
        tc_fence_after();
        if (warp == 2 && lane == 0) TL(5);
        for (int c0 = 0; c0 < np; c0 += 32) {
            float v[32];
            const bool wide = c0 + 32 <= np;          // np is a multiple of 16: the tail chunk is 16 wide
            if (wide) {
                tmem_ld_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            } else {
                tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
                for (int i = 16; i < 32; ++i) v[i] = 0.f;
            }
            const int ng = wide ? 4 : 2;              // 8-channel groups in this chunk
            if (!do_red) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    if (i < ng * 8) {
                        const float4 bb = *reinterpret_cast<const float4*>(bias_s + c0 + i);
                        v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
                    }
                }
            }
            float yh[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) yh[i] = 0.f;
            if (do_red) {
                // d = dz * [scale*y + shift > 0]; yhat = (y - mean) * invstd of the layer below
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (g < ng) {
                        float yv[8];
                        if (valid) {
                            const float* src = mbase + (size_t)((c0 >> 3) + g) * chunk_stride;
                            const float4 a0 = __ldg(reinterpret_cast<const float4*>(src));
                            const float4 a1 = __ldg(reinterpret_cast<const float4*>(src + 4));
                            yv[0] = a0.x; yv[1] = a0.y; yv[2] = a0.z; yv[3] = a0.w; yv[4] = a1.x; yv[5] = a1.y; yv[6] = a1.z; yv[7] = a1.w;
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) yv[j] = 0.f;
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int c = c0 + g * 8 + j;
                            const float z = fmaf(yv[j], bias_s[c], bias_s[np + c]);
                            if (!(z > 0.f)) v[g * 8 + j] = 0.f;
                            yh[g * 8 + j] = (yv[j] - bias_s[2 * np + c]) * bias_s[3 * np + c];
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) yh[g * 8 + j] = 0.f;
                    }
                }
            }
            if (do_aff) {
                // inference: eval-mode BatchNorm (+ shortcut branch) (+ ReLU) applied to the accumulators; the pre-BN y
                // never reaches HBM and the next layer's operand is written directly
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (g < ng) {
                        float rv[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) rv[j] = 0.f;
                        if (p.aff_res && valid) {
                            const float* src = p.aff_res + row_off + (size_t)((c0 >> 3) + g) * chunk_stride;
                            const float4 a0 = __ldg(reinterpret_cast<const float4*>(src));
                            const float4 a1 = __ldg(reinterpret_cast<const float4*>(src + 4));
                            rv[0] = a0.x; rv[1] = a0.y; rv[2] = a0.z; rv[3] = a0.w; rv[4] = a1.x; rv[5] = a1.y; rv[6] = a1.z; rv[7] = a1.w;
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int c = c0 + g * 8 + j;
                            float z = fmaf(v[g * 8 + j], bias_s[np + c], bias_s[2 * np + c]) + rv[j];
                            if (p.aff_relu) z = fmaxf(z, 0.f);
                            v[g * 8 + j] = z;
                        }
                    }
                }
                if (p.aff_kind == TSC_OUT_C8_BF16) {
                    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(p.aff_out) + row_off;
                    if (valid) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (g < ng) {
                                uint4 raw;
                                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
                                for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
                                *reinterpret_cast<uint4*>(ob + (size_t)((c0 >> 3) + g) * chunk_stride) = raw;
                            }
                        }
                    }
                } else if (p.aff_kind == TSC_OUT_C8_F32) {
                    float* of = reinterpret_cast<float*>(p.aff_out) + row_off;
                    if (valid) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (g < ng) {
                                float* d0 = of + (size_t)((c0 >> 3) + g) * chunk_stride;
                                *reinterpret_cast<float4*>(d0) = make_float4(v[g * 8], v[g * 8 + 1], v[g * 8 + 2], v[g * 8 + 3]);
                                *reinterpret_cast<float4*>(d0 + 4) = make_float4(v[g * 8 + 4], v[g * 8 + 5], v[g * 8 + 6], v[g * 8 + 7]);
                            }
                        }
                    }
                } else if (p.aff_kind == TSC_OUT_NCL_F32) {
                    // [B][Cout][L]: the 32 lanes of a warp are 32 consecutive positions of one channel (128 B per store)
                    float* on = reinterpret_cast<float*>(p.aff_out) + (size_t)b * p.nbias * p.L + (valid ? l : 0);
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < ng * 8 && c0 + i < p.nbias) on[(size_t)(c0 + i) * p.L] = v[i];
                    }
                } else {
                    // TSC_OUT_POOLED (one CTA per sample, L <= 128): column sums over this warp's valid rows
                    float s1[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) s1[i] = valid ? v[i] : 0.f;
                    const float a = warp_colsum32(s1, lane);
                    if (c0 + lane < np) wstat[q * np + c0 + lane] = make_float2(a, 0.f);
                }
            } else if (valid && p.y) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (g < ng) {
                        float* d0 = ybase + (size_t)((c0 >> 3) + g) * chunk_stride;
                        *reinterpret_cast<float4*>(d0) = make_float4(v[g * 8], v[g * 8 + 1], v[g * 8 + 2], v[g * 8 + 3]);
                        *reinterpret_cast<float4*>(d0 + 4) = make_float4(v[g * 8 + 4], v[g * 8 + 5], v[g * 8 + 6], v[g * 8 + 7]);
                    }
                }
            }
            if (do_red) {
                float s1[32], s2[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float x = valid ? v[i] : 0.f;
                    s1[i] = x;
                    s2[i] = x * yh[i];
                }
                const float a = warp_colsum32(s1, lane);
                const float c = warp_colsum32(s2, lane);
                if (c0 + lane < np) wstat[q * np + c0 + lane] = make_float2(a, c);
            } else if (do_stat) {
                // two-pass per warp: column means first, then the centred sums of squares (no cancellation even when
                // |mean| >> std); lane j ends up with (sum, M2) of column c0 + j over this warp's valid rows
                const int nw = max(0, min(32, min(128, p.L - l0) - q * 32));
                float s1[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) s1[i] = valid ? v[i] : 0.f;
                const float a = warp_colsum32(s1, lane);
                const float mean_l = nw > 0 ? a / (float)nw : 0.f;
                float s2[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float m = __shfl_sync(0xffffffffu, mean_l, i);
                    const float dlt = valid ? v[i] - m : 0.f;
                    s2[i] = dlt * dlt;
                }
                const float c = warp_colsum32(s2, lane);
                if (c0 + lane < np) wstat[q * np + c0 + lane] = make_float2(mean_l, c);
            }
        }
        if (do_aff && p.aff_kind == TSC_OUT_POOLED) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const float inv_l = 1.f / (float)p.L;
            for (int c = threadIdx.x - 64; c < p.nbias; c += 128) {
                float a = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) a += wstat[w * np + c].x;
                reinterpret_cast<float*>(p.aff_out)[(size_t)b * p.nbias + c] = a * inv_l;      // AdaptiveAvgPool1d(1)
            }
        } else if (do_stat || do_red) {
            asm volatile("bar.sync 1, 128;" ::: "memory");             // the four epilogue warps
            const int rows_cta = min(128, p.L - l0);
            for (int c = threadIdx.x - 64; c < np; c += 128) {
                if (do_red) {
                    float a = 0.f, d = 0.f;
#pragma unroll
                    for (int w = 0; w < 4; ++w) { const float2 t = wstat[w * np + c]; a += t.x; d += t.y; }
                    reinterpret_cast<float2*>(p.red_partial)[(size_t)blockIdx.x * np + c] = make_float2(a, d);
                } else {
                    // Chan merges of the four warps' (n, mean, M2)
                    float n = 0.f, mean = 0.f, m2 = 0.f;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const int nw = max(0, min(32, rows_cta - w * 32));
                        if (nw > 0) {
                            const float2 t = wstat[w * np + c];
                            welford_merge(n, mean, m2, (float)nw, t.x, t.y);
                        }
                    }
                    reinterpret_cast<float2*>(p.stat_partial)[(size_t)blockIdx.x * np + c] = make_float2(mean, m2);
                }
            }
        }
        if (warp == 2 && lane == 0) TL(6);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    if (warp == 1 && lane == 0) TL(7);
    if (DBG && (p.debug & 16) && p.tl && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.tl[1024 + 4 * blockIdx.x + 1] = (long long)t;
        p.tl[1024 + 4 * blockIdx.x + 3] = clock64() - p.tl[1024 + 4 * blockIdx.x + 3];
    }
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

// c8 bf16 tensor [B][kc][L][8] as a 4-D tensor map with a (8, rows, chunks, 1) box, no swizzle, zero OOB fill.
int make_c8_map(CUtensorMap* map, const void* base, int B, int kc, int L, int box_rows, int box_chunks) {
    EncodeTiledFn enc = get_encode_tiled();
    TSC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    TSC_REQUIRE(((uintptr_t)base & 15) == 0, "c8 tensor must be 16-byte aligned");
    TSC_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows %d outside [1,256]", box_rows);
    TSC_REQUIRE(box_chunks >= 1 && box_chunks <= 256, "TMA box chunks %d outside [1,256]", box_chunks);
    cuuint64_t dims[4] = {8, (cuuint64_t)L, (cuuint64_t)kc, (cuuint64_t)B};
    cuuint64_t strides[3] = {16, (cuuint64_t)L * 16, (cuuint64_t)kc * L * 16};
    cuuint32_t box[4] = {8, (cuuint32_t)box_rows, (cuuint32_t)box_chunks, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TSC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

static constexpr int SMEM_CAP = 227 * 1024;
static constexpr int SMEM_HALF = 113 * 1024;
static long long* g_timeline = nullptr;     // debug only: device buffer of >= 8 clock64 samples

static inline int conv_rp(int Kmax) { return (128 + Kmax - 1 + 7) & ~7; }

// experiment knobs (environment, read once): TSC_CONV_STAGE_KB, TSC_CONV_SMEM_FULL, TSC_CONV_DEBUG
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e && e[0] ? atoi(e) : dflt;
}
static int knob_stage_bytes() { static const int v = env_int("TSC_CONV_STAGE_KB", PLAN_STAGE_BYTES / 1024) * 1024; return v; }
static int knob_smem_full() { static const int v = env_int("TSC_CONV_SMEM_FULL", 0); return v; }
static int knob_debug() { static const int v = env_int("TSC_CONV_DEBUG", 0); return v; }

// ---- the plan: header (16 B) | stage table (8 B each, padded to 16) | MMA table (16 B each) ----------
struct PlanHost {
    std::vector<uint2> stages;
    std::vector<uint4> mmas;
    int stage_bytes = PLAN_STAGE_BYTES;
};

// One MMA of the plan before it is assigned to a stage.
struct PlanUnit { uint32_t a_off, nt, n_lo, bytes, src16; };

// Emit units [i0, i1) as one weight stage: entries carry their in-stage B offset and the stage index (the kernel turns
// both into a shared-memory address once, at start-up); the list is padded to a multiple of MMA_GROUP with instructions
// that add zero (N = 16, B = the zero block behind the stage slot, A = the tile base).
static void emit_stage(PlanHost* ph, const std::vector<PlanUnit>& u, size_t i0, size_t i1) {
    const uint32_t stage_idx = (uint32_t)ph->stages.size();
    uint32_t bytes = 0;
    for (size_t i = i0; i < i1; ++i) {
        uint4 e;
        e.x = u[i].a_off;                                        // A start, 16 B units from the tile base
        e.y = (bytes >> 4) | (u[i].nt << 16);                    // B start within the stage | LBO = nt * 16 B
        // = make_idesc_bf16(128, nt, K-major, K-major): F32 accumulate, BF16 x BF16
        e.z = (1u << 4) | (1u << 7) | (1u << 10) | ((u[i].nt >> 3) << 17) | ((128u >> 4) << 24);
        e.w = u[i].n_lo | (i == i0 ? PF_FIRST : 0u) | (stage_idx << PF_STAGE_SHIFT);
        ph->mmas.push_back(e);
        bytes += u[i].bytes;
    }
    ph->stages.push_back(make_uint2(u[i0].src16, bytes));
    while (ph->mmas.size() % MMA_GROUP != 0) {
        uint4 e;
        e.x = 0;
        e.y = ((uint32_t)ph->stage_bytes >> 4) | (16u << 16);      // LBO = 16 rows * 16 B
        e.z = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        e.w = stage_idx << PF_STAGE_SHIFT;
        ph->mmas.push_back(e);
    }
    ph->mmas.back().w |= PF_LAST;
}

static int build_plan(int direction, int Cin, int Cout, int Kmax, const int* s_of_tap, PlanHost* ph) {
    TapTable tt;
    if (build_tap_table(direction, Cin, Cout, Kmax, s_of_tap, &tt) != 0) return -1;
    const int Rp = conv_rp(Kmax);
    // a large activation tile leaves less room for the weight ring: halve the stage so that two stages still fit the
    // half-SM shared-memory budget (two CTAs of different launches can then share an SM)
    ph->stage_bytes = tt.kc * Rp * 16 > 48 * 1024 ? knob_stage_bytes() / 2 : knob_stage_bytes();
    std::vector<PlanUnit> units;
    for (int oi = 0; oi < tt.n_order; ++oi) {
        const int t = tt.order[oi];
        const int n_lo = tt.n_lo[t], kc_lo = tt.kc_lo[t];
        const int nt = tt.np - n_lo, kspan = tt.kc - kc_lo;
        for (int kp = 0; kp < kspan / 2; ++kp)
            units.push_back({(uint32_t)((kc_lo + 2 * kp) * Rp + t), (uint32_t)nt, (uint32_t)n_lo, (uint32_t)(2 * nt * 16),
                             (uint32_t)tt.w_off[t] + (uint32_t)(2 * kp * nt)});
    }
    TSC_REQUIRE(!units.empty(), "kernel bank has no live tap");
    // Stages hold whole issue groups: the issuing thread is the bottleneck of this kernel (~90 cycles per instruction
    // whatever its N), so an instruction that only pads a group costs as much as a real one.  Units that would leave a
    // partial group at the end of a stage open the next stage instead; only a stage too small for one group, and the
    // last stage, are padded.
    size_t i = 0;
    while (i < units.size()) {
        size_t k = 0;
        uint32_t bytes = 0;
        while (i + k < units.size() && bytes + units[i + k].bytes <= (uint32_t)ph->stage_bytes) { bytes += units[i + k].bytes; ++k; }
        TSC_REQUIRE(k >= 1, "one MMA's weights (%u B) exceed the %d B stage", units[i].bytes, ph->stage_bytes);
        if (i + k < units.size() && k >= (size_t)MMA_GROUP) k -= k % MMA_GROUP;
        TSC_REQUIRE(ph->stages.size() < (size_t)PF_STAGE_MAX, "more than %d weight stages", PF_STAGE_MAX);
        emit_stage(ph, units, i, i + k);
        i += k;
    }
    return 0;
}

static inline size_t plan_bytes_of(const PlanHost& ph) {
    // ... | one flag byte per issue group (filled in by the kernel), padded to 16
    return 16 + ((ph.stages.size() * 8 + 15) & ~(size_t)15) + ph.mmas.size() * 16 + ((ph.mmas.size() / MMA_GROUP + 15) & ~(size_t)15);
}

}  // namespace tc

void set_conv_timeline(long long* dev) { tc::g_timeline = dev; }

size_t osconv_plan_bytes(int direction, int Cin, int Cout, int Kmax, const int* s_of_tap) {
    tc::PlanHost ph;
    if (tc::build_plan(direction, Cin, Cout, Kmax, s_of_tap, &ph) != 0) return 0;
    return tc::plan_bytes_of(ph);
}

int osconv_plan_build(int direction, int Cin, int Cout, int Kmax, const int* s_of_tap, void* host_plan) {
    tc::PlanHost ph;
    if (tc::build_plan(direction, Cin, Cout, Kmax, s_of_tap, &ph) != 0) return -1;
    uint8_t* o = (uint8_t*)host_plan;
    memset(o, 0, tc::plan_bytes_of(ph));
    int hdr[4] = {(int)ph.mmas.size(), (int)ph.stages.size(), ph.stage_bytes, 0x504c414e};
    memcpy(o, hdr, 16);
    memcpy(o + 16, ph.stages.data(), ph.stages.size() * 8);
    memcpy(o + 16 + ((ph.stages.size() * 8 + 15) & ~(size_t)15), ph.mmas.data(), ph.mmas.size() * 16);
    return 0;
}

int osconv_tc(int direction, const void* x, int dtype, const void* w, const void* plan, const float* bias, float* y,
              const tsc_conv_epilogue* epi, int B, int L, int Cin, int Cout, int Kmax, const int* s_of_tap, cudaStream_t cs) {
    using namespace tc;
    TSC_REQUIRE(dtype == TSC_BF16, "the tcgen05 engine takes bf16 operands");
    TSC_REQUIRE(plan != nullptr, "the tcgen05 engine needs the device copy of tsc_osconv_plan_build()'s plan");
    // geometry only (the schedule itself was built once by osconv_plan_build and lives on the device)
    PlanHost ph;
    if (build_plan(direction, Cin, Cout, Kmax, s_of_tap, &ph) != 0) return -1;
    const bool fwd = direction == TSC_DIR_FWD;
    ConvTcParams p;
    memset(&p, 0, sizeof(p));
    p.w = (const __nv_bfloat16*)w;
    p.plan = (const uint8_t*)plan;
    p.bias = fwd ? bias : nullptr;
    p.nbias = fwd ? Cout : 0;
    p.y = y;
    if (epi) {
        if (fwd) {
            p.stat_partial = epi->stat_partial;
            if (epi->affine_out) {
                TSC_REQUIRE((epi->affine_scale && epi->affine_shift) || (epi->bn_gamma && epi->bn_beta && epi->bn_mean && epi->bn_var),
                            "affine epilogue needs (scale, shift) or (gamma, beta, running mean, running var)");
                TSC_REQUIRE(!epi->stat_partial, "affine epilogue (inference) and BatchNorm statistics (training) exclude each other");
                TSC_REQUIRE(epi->affine_out_kind == TSC_OUT_C8_BF16 || epi->affine_out_kind == TSC_OUT_C8_F32 ||
                            epi->affine_out_kind == TSC_OUT_NCL_F32 || epi->affine_out_kind == TSC_OUT_POOLED,
                            "bad affine_out_kind %d", epi->affine_out_kind);
                TSC_REQUIRE(epi->affine_out_kind != TSC_OUT_POOLED || L <= 128,
                            "the pooled inference epilogue needs L <= 128 (one CTA per sample), got %d", L);
                p.aff_scale = epi->affine_scale; p.aff_shift = epi->affine_scale ? epi->affine_shift : nullptr;
                p.bn_gamma = epi->bn_gamma; p.bn_beta = epi->bn_beta; p.bn_mean = epi->bn_mean; p.bn_var = epi->bn_var;
                p.bn_eps = epi->bn_eps;
                p.aff_res = epi->residual;
                p.aff_out = epi->affine_out; p.aff_kind = epi->affine_out_kind; p.aff_relu = epi->affine_relu ? 1 : 0;
            }
        } else if (epi->red_partial) {
            TSC_REQUIRE(epi->mask_y && epi->mask_mean && epi->mask_invstd, "dgrad reduction needs mask_y, mask_mean, mask_invstd");
            TSC_REQUIRE(!epi->mask_scale || epi->mask_shift, "mask_scale needs mask_shift");
            p.mask_y = epi->mask_y; p.mask_scale = epi->mask_scale; p.mask_shift = epi->mask_shift;
            p.mask_mean = epi->mask_mean; p.mask_invstd = epi->mask_invstd; p.red_partial = epi->red_partial;
        }
    }
    p.B = B; p.L = L; p.ltiles = cdiv(L, 128);
    p.np = fwd ? pad16(Cout) : pad16(Cin);
    p.kc = fwd ? pad16(Cin) / 8 : pad16(Cout) / 8;
    p.pad_left = fwd ? (Kmax - 1) / 2 : Kmax / 2;
    p.Rp = conv_rp(Kmax);
    p.n_stages = (int)ph.stages.size();
    p.n_mma = (int)ph.mmas.size();
    p.plan_bytes = (int)plan_bytes_of(ph);
    p.off_bias = SMEM_HDR;
    p.off_wstat = p.off_bias + 4 * p.np * 4;
    p.off_plan = (p.off_wstat + 4 * p.np * 8 + 15) & ~15;
    p.off_xs = (p.off_plan + p.plan_bytes + 127) & ~127;
    p.off_stages = (p.off_xs + p.kc * p.Rp * 16 + 127) & ~127;
    p.stage_bytes = ph.stage_bytes;
    const int slot = p.stage_bytes + ZERO_BLOCK_BYTES;
    // prefer half of an SM's shared memory (113 KB): CTAs of two independent launches (the target and the source branch
    // of a step run on two streams) can then be co-resident and hide each other's load / epilogue latency
    int cap = knob_smem_full() ? SMEM_CAP : SMEM_HALF;
    int ns = (cap - p.off_stages) / slot;
    if (ns < 2 && p.n_stages > 1) { cap = SMEM_CAP; ns = (cap - p.off_stages) / slot; }
    if (ns > 8) ns = 8;
    if (ns > p.n_stages) ns = p.n_stages;
    TSC_REQUIRE(ns >= 1 && (ns >= 2 || p.n_stages == 1), "shape needs %d B of shared memory before the weight stages: unsupported",
                p.off_stages);
    p.NS = ns;
    p.tl = g_timeline;
    p.debug = knob_debug();
    p.tmem_cols = p.np <= 32 ? 32 : p.np <= 64 ? 64 : p.np <= 128 ? 128 : 256;
    CUtensorMap xmap;
    if (make_c8_map(&xmap, x, B, p.kc, L, p.Rp, p.kc) != 0) return -1;
    const int smem = p.off_stages + ns * slot;
    static OnceAttr attr_once;             // once per process: opt in to the full 227 KB of dynamic shared memory
    {
        const cudaError_t e = run_once(attr_once, [] {
            cudaError_t r = cudaFuncSetAttribute(osconv_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(osconv_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(osconv_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(osconv_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            return r;
        });
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    {
        // the instrumented instantiation only when a timeline buffer or an experiment knob is set
        const bool dbg = p.tl != nullptr || p.debug != 0;
        const dim3 grid(B * p.ltiles), block(TC_THREADS);
        cudaError_t le = p.aff_out ? (dbg ? launch_pdl(osconv_tc_kernel<true, true>, grid, block, (size_t)smem, cs, xmap, p)
                                          : launch_pdl(osconv_tc_kernel<true, false>, grid, block, (size_t)smem, cs, xmap, p))
                                   : (dbg ? launch_pdl(osconv_tc_kernel<false, true>, grid, block, (size_t)smem, cs, xmap, p)
                                          : launch_pdl(osconv_tc_kernel<false, false>, grid, block, (size_t)smem, cs, xmap, p));
        if (le != cudaSuccess) { set_error("osconv launch: %s", cudaGetErrorString(le)); return (int)le; }
    }
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
