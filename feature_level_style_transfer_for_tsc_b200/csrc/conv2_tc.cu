// tcgen05 engine, forward / dgrad, second generation: the omni-scale convolution as a PERSISTENT implicit GEMM.
//
//   positions (128 per tile)         -> MMA M      (TMEM lanes; 256 per CTA pair in the cta_group::2 variant)
//   output channels (<= 256)         -> MMA N      (the whole channel axis is one accumulator tile in TMEM)
//   (tap, input channel)             -> MMA K      (16 channels per instruction)
//
// What changed against conv_tc.cu (round 1), and why -- measurements: profiles/r2_conv2_issue_bisection.md, r2_conv2_dual.md,
// r2_conv2_B1024.md:
// * The issue loop lives in the UNIFORM datapath.  The schedule is a table of per-tap RUNS and of weight STAGES in the kernel
//   parameter block (constant bank); per stage the issuing warp does one vote-wait (`__all_sync` of try_wait: the warp's control
//   flow stays provably uniform), ONE elected region that walks the stage's runs, and one commit; a run is 4 LDCU + 3 adds +
//   an inline-PTX K loop of 6 SASS instructions per MMA (descriptors = mad.wide.u32(k, step, base)).  The tensor pipe accepts
//   an MMA only when its predecessor has fetched its operands (~68 cycles at the cfg2 bank's mean N), so what the issuer
//   executes between two MMAs is free up to that length and serial beyond: the per-RUN header is what had to shrink (the
//   number of instructions inside a run, the operand addresses, N and the accumulator column do not matter -- bisection).
// * One CTA walks several position tiles (grid = min(tiles, SMs)); accumulator and activation tiles are multi-buffered, the
//   epilogue of one tile overlaps the MMAs of the next, and -- when the accumulator is <= 128 columns or the bank >= 256 KB --
//   every weight stage serves a PAIR of tiles (`dual`): one L2 -> shared-memory fetch and one barrier round trip per pair.
//   With one tile per CTA (cfg2: 128 tiles) the kernel degenerates to the single-buffered layout that fits half an SM, so
//   that CTAs of two streams stay co-resident.
// * No padding MMAs and no plan relocation pass: a weight stage holds whole runs.
// * cta_group::2 variant (PAIR, TSC_CONV_PAIR=1): the pair's two position tiles share every MMA, each CTA holds its own
//   activation tile and HALF of the tap's live channels (split packed layout of common.cuh); the leader CTA issues, the peer's
//   would-be issuer relays "landed" signals, commits are multicast.  Parity-green but slower than one CTA per tile at every
//   measured shape (the cross-CTA barrier traffic costs more than the halved B fetch saves): an opt-in, not the default.
//
// Unchanged: the c8 activation tile with halo staged by ONE TMA box load (out-of-bounds rows zero-filled = ConstantPad1d,
// OS_CNN.py:59,70), a tap = "+ t rows" on the A descriptor; the packed bank (live (channel, tap) pairs only) streamed through a
// ring of shared-memory stages by 1-D bulk copies; per tap the MMA covers the live channel suffix only; the epilogues (bias,
// BatchNorm partial statistics, dgrad ReLU mask + BatchNorm-backward partial sums, inference affine) of conv_tc.cu.
// Replaces ConstantPad1d + Conv1d (+ cuDNN dgrad), OS_CNN/OS_CNN.py:70-71; arithmetic SURVEY A1/A2.
#include "tc_common.cuh"
#include <stdlib.h>
#include <string.h>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

namespace tsc {
namespace tc {

static constexpr int C2_THREADS = 192;
static constexpr int C2_MAX_RUNS = 320;
static constexpr int C2_MAX_STAGES = 256;
static constexpr int C2_STAGE_BYTES = 16 * 1024;    // per CTA: a stage is 32 KB of the bank, half of it in each CTA of the pair
static constexpr int C2_HDR = 384;
static constexpr uint32_t RF_FIRST = 1u, RF_LAST = 2u;

// One run = consecutive K steps of one tap inside one weight stage: two uint4 (every field ready to use -- the issuing warp
// is bound by its instruction count, so nothing is decoded on the device).
//  [0] x: A start (16 B units from the activation tile base) of the first K step
//      y: B start inside this CTA's half stage (16 B units) | nt / 2 << 16 (= the descriptor's leading-dimension field: each
//         CTA of the pair holds nt / 2 of the tap's nt live channels)
//      z: instruction descriptor (M = 256, N = nt, bf16 x bf16 -> f32, both K-major)
//      w: TMEM column of the tap's first live channel
//  [1] x: K steps   y: B advance per K step (16 B units)   z: accumulate flag (0: the tile's first MMA overwrites)
//      w: RF_FIRST: opens its weight stage | RF_LAST: closes it
struct ConvSched {
    int n_runs, n_stages, stage_bytes;
    int half_rows16;                  // 16 B rows of one stream of the split packed bank
    uint4 runs[2 * C2_MAX_RUNS];
    uint4 stages[C2_MAX_STAGES];      // {source offset in 16 B units, bytes, first run, runs} -- a stage holds whole runs
};

struct Conv2Params {
    const __nv_bfloat16* w;
    const float* bias;
    float* y;
    // fused epilogues (all nullable) -- see tsc_conv_epilogue in include/tsc_b200.h
    float* stat_partial;       // FWD : [tiles][np] float2 (mean, M2) over the tile's valid rows
    const float* mask_y;       // DGRAD: pre-BN output of the layer below, c8 fp32 [B][np/8][L][8]
    const float* mask_scale;   //        z = scale*y + shift; d = dz * [z > 0]   (NULL = no ReLU)
    const float* mask_shift;
    const float* mask_mean;    //        yhat = (y - mean) * invstd
    const float* mask_invstd;
    float* red_partial;        // DGRAD: [tiles][np] float2 (S1, S2) partial sums over the tile's valid rows
    const float* aff_scale;    // FWD, inference: z = act(aff_scale * (acc + bias) + aff_shift [+ aff_res]) -> aff_out
    const float* aff_shift;
    const float* bn_gamma; const float* bn_beta; const float* bn_mean; const float* bn_var; float bn_eps;
    const float* aff_res;
    void* aff_out;
    int aff_kind, aff_relu;
    int nbias, B, L, ltiles, n_tiles;
    int tiles_base, tiles_rem;     // tile PAIRS: n_pairs = tiles_base * (grid / 2) + tiles_rem
    int half_rows16;               // 16 B rows of one stream of the split packed bank (the peer CTA's source offset)
    int np;            // padded output channels of this direction
    int Rp;            // halo rows of the activation tile (multiple of 8)
    int kc;            // input-channel chunks of 8
    int pad_left;
    int NS;            // weight stages in the ring
    int nx, na;        // activation tiles / accumulator tiles (1, 2 or 4)
    int lg_nx, lg_na;  // their base-2 logarithms
    int dual;          // persistent CTAs: two position tiles share every weight stage (one pass over the bank per tile pair)
    int acc_stride;    // TMEM columns per accumulator tile
    int tmem_cols;
    int off_bias, off_wstat, off_xs, off_stages, x_bytes;
    long long* tl;     // optional phase timeline of CTA 0 (tsc_debug_set_timeline), NULL in production
    int debug;         // experiments only (TSC_C2_DEBUG; garbage results): 1 = the producer signals its stages full without copying
};

// DBG is a template flag of the kernel: the production instantiation carries no timeline code (the issuing warp is bound by its
// instruction count)
#define TL2(i) do { if (DBG && p.tl && blockIdx.x == 0) p.tl[i] = clock64(); } while (0)
// per-stage samples of CTA 0 (debug timeline only): slot k of weight stage i (counted over all tiles), first 48 stages
#define TLS2(i, k) do { if (DBG && p.tl && blockIdx.x == 0 && (i) < 48) p.tl[8 + (i) * 8 + (k)] = clock64(); } while (0)

// warp-uniform bounded wait: all 32 lanes poll and the loop branch is a vote, so the warp's control flow stays uniform
// (the issue loop then lives in uniform registers).  false = timed out (watchdog).
__device__ __forceinline__ bool mbar_wait_warp(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t i = 0; i < (1u << 22); ++i) {
        if (__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) return true;
    }
    return false;
}

// ---- CTA-pair primitives (cluster of 2, cta_group::2) ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA's window) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// as mbar_wait_warp, for barriers the PEER CTA arrives on (acquire at cluster scope)
__device__ __forceinline__ bool mbar_wait_warp_cluster(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t i = 0; i < (1u << 22); ++i) {
        if (__all_sync(0xffffffffu, mbar_try_wait_cluster(bar, parity))) return true;
    }
    return false;
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// "all MMAs issued so far by this thread have completed" -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows in each CTA] * B[N/2 rows in each CTA]; one thread of the leader CTA issues for the pair
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
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

// Column sums over the 32 lanes of a warp for 32 columns held one per register: a transposing butterfly.
// On return lane j holds the sum over all lanes of v[j] (31 shuffles instead of 160).
__device__ __forceinline__ float colsum32(float* v, int lane) {
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

// The K steps of one run, issued by the elected lane as ONE inline-PTX loop: per MMA two 32-bit adds on the low descriptor
// words (the high words -- SBO, descriptor version -- never change), the loop counter and the branch.  Written in PTX because
// the issuing warp is bound by its instruction count: from C++ ptxas built the 64-bit descriptors with carry chains and
// register-pair moves (13 SASS instructions per MMA in the tail loop, 26 per MMA over the whole loop -- ncu source page,
// profiles/r2_conv2_B1024.md); this form needs 6.
template <bool PAIR>
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    if (PAIR) tc_commit2(bar);
    else tc_commit(bar);
}
template <bool PAIR>
__device__ __forceinline__ void issue_run(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc,
                                          uint32_t ks, uint32_t a_step, uint32_t b_step) {
    if (PAIR) {
        asm volatile(
            "{\n\t.reg .pred pa, pk;\n\t.reg .b32 k;\n\t.reg .b64 da, db, a0, b0;\n\t"
            "mov.b64 a0, {%1, %3};\n\tmov.b64 b0, {%2, %3};\n\t"
            "mov.u32 k, 0;\n\t"
            "setp.ne.b32 pa, %5, 0;\n"
            "RUN_LOOP:\n\t"
            "mad.wide.u32 da, k, %7, a0;\n\t"
            "mad.wide.u32 db, k, %8, b0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, pa;\n\t"
            "add.u32 k, k, 1;\n\t"
            "setp.ne.u32 pk, k, %6;\n\t"
            "@pk bra.uni RUN_LOOP;\n\t}"
            ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc), "r"(ks), "r"(a_step), "r"(b_step)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred pa, pk;\n\t.reg .b32 k;\n\t.reg .b64 da, db, a0, b0;\n\t"
            "mov.b64 a0, {%1, %3};\n\tmov.b64 b0, {%2, %3};\n\t"
            "mov.u32 k, 0;\n\t"
            "setp.ne.b32 pa, %5, 0;\n"
            "RUN_LOOP:\n\t"
            "mad.wide.u32 da, k, %7, a0;\n\t"
            "mad.wide.u32 db, k, %8, b0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pa;\n\t"
            "add.u32 k, k, 1;\n\t"
            "setp.ne.u32 pk, k, %6;\n\t"
            "@pk bra.uni RUN_LOOP;\n\t}"
            ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc), "r"(ks), "r"(a_step), "r"(b_step)
            : "memory");
    }
}

template <bool AFF, bool DBG, bool PAIR>
__global__ void __launch_bounds__(C2_THREADS, AFF ? 2 : 1)
osconv2_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ ConvSched S, const __grid_constant__ Conv2Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);            // [8]  weight stage filled
    uint64_t* empty = full + 8;                                    // [8]  weight stage consumed
    uint64_t* x_full = empty + 8;                                  // [4]  activation tile landed
    uint64_t* x_empty = x_full + 4;                                // [4]  ... consumed by the tile's MMAs
    uint64_t* acc_full = x_empty + 4;                              // [4]  accumulator tile complete
    uint64_t* acc_empty = acc_full + 4;                            // [4]  ... drained by BOTH epilogues (leader's copy is used)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 4);
    float* bias_s = reinterpret_cast<float*>(smem + p.off_bias);
    float2* wstat = reinterpret_cast<float2*>(smem + p.off_wstat);  // [4][np]
    uint8_t* xs = smem + p.off_xs;
    uint8_t* stages = smem + p.off_stages;

    // the warp index as a shuffle broadcast: ptxas cannot know that threadIdx.x >> 5 is the same in all lanes (it does not know
    // the block shape), and without that knowledge every role branch is a potentially divergent one -- the issuing warp's votes
    // get BRA.DIV guards and its loop falls back to vector registers + R2UR (SASS inspected)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int np = p.np;
    // tiles of this CTA (tile = blockIdx.x + j * gridDim.x); no division here: the issuing warp's loop bounds must stay in the
    // uniform datapath
    // PAIR: cluster of two CTAs = one tcgen05 CTA pair; rank 0 = leader (issues the pair's MMAs), 1 = peer.  A "unit" is what
    // one issuer works on at a time: a tile pair (PAIR) or a tile.
    constexpr int NCTA = PAIR ? 2 : 1;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int pair = (int)blockIdx.x / NCTA, n_pairs_grid = (int)gridDim.x / NCTA;
    const int n_my = p.tiles_base + (pair < p.tiles_rem ? 1 : 0);  // units of this CTA (pair): tile = NCTA * (pair + j * n_pairs_grid) + rank
    const int slot_bytes = S.stage_bytes;

    if (warp == 0 && lane == 0) {
        TL2(0);
        tma_prefetch_desc(&xmap);
        // the leader's full[] / x_full[] collect BOTH CTAs' "landed" signals -- its own copy's expect_tx arrival and the peer
        // relay's remote arrival -- so the issuer waits on one barrier per stage / tile
        const uint32_t n_land = (PAIR && rank == 0) ? 2u : 1u;
        for (int i = 0; i < p.NS; ++i) { mbar_init(&full[i], n_land); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&x_full[i], n_land); mbar_init(&x_empty[i], 1);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], NCTA);   // PAIR: the two CTAs' epilogues
        }
        fence_barrier_init();
    }
    if (warp == 1) { if (PAIR) tmem_alloc2(tmem_slot, (uint32_t)p.tmem_cols); else tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols); }
    if (warp >= 2) {
        pdl_wait();
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
    if (PAIR) cluster_sync_all(); // both CTAs' barriers are initialised before any remote arrive / multicast commit
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== copy producer (ONE thread: a second polling lane of the same warp would stall this one for the length of its
        // mbarrier suspend -- measured): this CTA's half of the packed bank, one bulk copy per stage, and the activation tiles
        // (TMA).  The next tile's activation buffer is requested opportunistically between two weight stages, as soon as the
        // tile that used it has completed, so that waiting for it never holds up the weight stream. =====
        if (lane == 0) {
            bool dead = false;
            pdl_wait();
            TL2(1);
            const int nx = p.nx;
            auto load_x = [&](int k) {
                const int xb = k & (nx - 1);
                const int tile = NCTA * (pair + k * n_pairs_grid) + (int)rank;   // past the last tile (odd count): b == B, all rows
                const int b = tile / p.ltiles, l0 = (tile - b * p.ltiles) * 128;  // out of bounds -> a tile of zeros
                mbar_arrive_expect_tx(&x_full[xb], (uint32_t)p.x_bytes);
                tma_load_4d(xs + (size_t)xb * p.x_bytes, &xmap, 0, l0 - p.pad_left, 0, b, &x_full[xb]);
            };
            // is the buffer of activation tile k free?  (its previous user, tile k - nx, has completed)
            auto x_free = [&](int k, bool block) -> bool {
                const int xb = k & (nx - 1), use = k >> p.lg_nx;
                if (use == 0) return true;
                if (block) { mbar_wait(&x_empty[xb], (uint32_t)((use - 1) & 1), dead, 10); return true; }
                return mbar_try_wait(&x_empty[xb], (uint32_t)((use - 1) & 1));
            };
            load_x(0);
            int x_next = 1;                                   // next activation tile to request
            while (x_next < n_my && x_next < nx) { load_x(x_next); ++x_next; }
            // dual: one pass over the bank serves the tile pair (2u, 2u + 1)
            const int tpp = p.dual ? 2 : 1;                   // tiles per pass
            const int n_pass = (n_my + tpp - 1) / tpp;
            uint32_t s = 0, ph = 0;
            const int n_stages = S.n_stages;
            // this CTA's half of every tap's live channels: the lower / upper stream of the split packed bank
            const uint8_t* w_src = reinterpret_cast<const uint8_t*>(p.w) + (size_t)rank * (size_t)p.half_rows16 * 16;
            for (int u = 0; u < n_pass; ++u) {
                const int j = u * tpp;                        // first tile of this pass
                for (int i = 0; i < n_stages; ++i) {
                    const uint4 e = S.stages[i];
                    if (x_next < n_my && x_next <= j + nx - 1 && x_free(x_next, false)) { load_x(x_next); ++x_next; }
                    TLS2(u * n_stages + i, 4);
                    if (DBG && (p.debug & 512)) mbar_wait_sleep(&empty[s], ph ^ 1u, dead, 1, 200); else mbar_wait(&empty[s], ph ^ 1u, dead, 1);
                    TLS2(u * n_stages + i, 5);
                    if (DBG && (p.debug & 1024)) continue;  // experiment: no weight pipeline at all (the issuer neither waits nor commits)
                    if (DBG && (p.debug & 1)) { mbar_arrive(&full[s]); if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1u; } continue; }
                    mbar_arrive_expect_tx(&full[s], e.y);
                    bulk_load(stages + (size_t)s * slot_bytes, w_src + (size_t)e.x * 16, e.y, &full[s]);
                    if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1u; }
                }
                // whatever the next pass still needs: its buffers are free at the latest when this pass's MMAs have completed
                while (x_next < n_my && x_next < j + 2 * tpp) { x_free(x_next, true); load_x(x_next); ++x_next; }
            }
        }
    } else if (warp == 1) {
        bool ok = true;
        const int n_runs = S.n_runs, n_stages = S.n_stages;
        const int nx = p.nx, na = p.na;
        uint32_t s = 0, ph = 0;
        if (!PAIR || rank == 0) {
            // ===== MMA issuer (leader CTA): the whole warp walks the run table in lock-step, in uniform registers; one elected
            // lane issues for the pair =====
            const uint32_t desc_hi = (128u >> 4) | (1u << 14);                        // SBO = 128 B, descriptor version 1
            const uint32_t xs16 = smem_u32(xs) >> 4;
            const uint32_t a_lbo = (uint32_t)p.Rp << 16;                               // LBO = Rp * 16 B (the two 8-channel K groups)
            const uint32_t a_step = 2u * (uint32_t)p.Rp;                               // one K step = two chunks further
            const uint32_t st16 = smem_u32(stages) >> 4;
            const uint32_t slot16 = (uint32_t)slot_bytes >> 4;
            const uint32_t x16 = (uint32_t)p.x_bytes >> 4;
            uint32_t slot_cur = st16;                                                  // start of weight slot s (16 B units)
            // One pass over the bank serves one tile, or -- persistent CTAs with several tiles (dual) -- the tile pair
            // (2u, 2u + 1): every run is issued for both tiles back to back, so the L2 -> shared-memory weight traffic and the
            // barrier round trips of the weight pipeline are paid once per pair (at B = 1024 they were 31 % of the kernel:
            // profiles/r2_conv2_issue_bisection.md).  The pair's accumulators are two of the `na` TMEM tiles.
            const int tpp = p.dual ? 2 : 1;
            const int n_pass = (n_my + tpp - 1) / tpp;
            for (int u = 0; u < n_pass; ++u) {
                const uint32_t j0 = (uint32_t)(u * tpp);
                const bool has_b = tpp == 2 && (int)j0 + 1 < n_my;
                const uint32_t j1 = has_b ? j0 + 1u : j0;
                const uint32_t xb0 = j0 & (uint32_t)(nx - 1), xuse0 = j0 >> p.lg_nx, xb1 = j1 & (uint32_t)(nx - 1), xuse1 = j1 >> p.lg_nx;
                const uint32_t ab0 = j0 & (uint32_t)(na - 1), ause0 = j0 >> p.lg_na, ab1 = j1 & (uint32_t)(na - 1), ause1 = j1 >> p.lg_na;
                ok = mbar_wait_warp(&x_full[xb0], xuse0 & 1u) && ok;
                if (has_b) ok = mbar_wait_warp(&x_full[xb1], xuse1 & 1u) && ok;
                if (ause0 > 0) ok = mbar_wait_warp(&acc_empty[ab0], (ause0 - 1u) & 1u) && ok;
                if (has_b && ause1 > 0) ok = mbar_wait_warp(&acc_empty[ab1], (ause1 - 1u) & 1u) && ok;
                tc_fence_after();
                if (u == 0) TL2(2);                   // (all lanes store: a lane-dependent branch here would cost the loop its uniformity)
                const uint32_t a_base = (xs16 + xb0 * x16) | a_lbo, a_base1 = (xs16 + xb1 * x16) | a_lbo;
                const uint32_t d_base = tmem_base + ab0 * (uint32_t)p.acc_stride, d_base1 = tmem_base + ab1 * (uint32_t)p.acc_stride;
                if (DBG && (p.debug & 2048)) {
                    // experiment: the micro-benchmark's loop inside this kernel -- 3 * n_runs + 6 identical MMAs (N = 112), no
                    // table, no barriers; what does the issue slot cost here?  (profiles/r2_conv2_issue_bisection.md)
                    const uint32_t idesc_x = (1u << 4) | (1u << 7) | (1u << 10) | ((112u >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
                    const uint32_t b_x = st16 | ((PAIR ? 56u : 112u) << 16);
#pragma unroll 1
                    for (int i = 0; i < 3 * n_runs + 6; i += 4) {
                        if (elect_one()) issue_run<PAIR>(d_base, a_base, b_x, desc_hi, idesc_x, 1u, 4u, 0u, 0u);
                    }
                } else {
                    // Stage by stage: one vote-wait, ONE elected region that walks the stage's runs (a run = the K steps of one
                    // tap inside this stage: a table entry, three adds and the inline-PTX K loop), one commit.  The tensor pipe
                    // accepts an MMA only when its predecessor has fetched its operands (~68 cycles at this bank's mean N): what
                    // the issuer does between two MMAs is free up to that length and serial beyond it -- the per-run header
                    // is what has to stay short (measured: profiles/r2_conv2_issue_bisection.md).
                    const bool no_pipe = DBG && (p.debug & 1024);
#pragma unroll 1
                    for (int i = 0; i < n_stages; ++i) {
                        if (!no_pipe) {
                            TLS2(u * n_stages + i, 0);
                            ok = mbar_wait_warp(&full[s], ph) && ok;
                            TLS2(u * n_stages + i, 1);
                        }
                        if (elect_one()) {
                            const uint32_t r0 = S.stages[i].z, r1 = r0 + S.stages[i].w;     // (read here: a value that lives across the
#pragma unroll 1                                                                            //  vote loop above lands in a vector register)
                            for (uint32_t r = r0; r < r1; ++r) {
                                const uint4 e = S.runs[2 * r], f = S.runs[2 * r + 1];
                                issue_run<PAIR>(d_base + e.w, a_base + e.x, slot_cur + e.y, desc_hi, e.z, f.z, f.x, a_step, f.y);
                                if (has_b) issue_run<PAIR>(d_base1 + e.w, a_base1 + e.x, slot_cur + e.y, desc_hi, e.z, f.z, f.x, a_step, f.y);
                            }
                            if (!no_pipe) mma_commit<PAIR>(&empty[s]);   // frees the stage (in both CTAs) when these MMAs have read it
                        }
                        TLS2(u * n_stages + i, 3);
                        slot_cur += slot16;
                        if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1u; slot_cur = st16; }
                    }
                }
                if (elect_one()) {
                    mma_commit<PAIR>(&acc_full[ab0]);
                    mma_commit<PAIR>(&x_empty[xb0]);
                    if (has_b) {
                        mma_commit<PAIR>(&acc_full[ab1]);
                        mma_commit<PAIR>(&x_empty[xb1]);
                    }
                }
            }
        } else {
            // ===== relay (peer CTA): tells the leader's barriers when this CTA's activation tile / half stage has landed =====
            const uint32_t px_remote = mapa_u32(smem_u32(x_full), 0), pf_remote = mapa_u32(smem_u32(full), 0);
            for (int j = 0; j < n_my; ++j) {
                const uint32_t xb = (uint32_t)j & (uint32_t)(nx - 1), xuse = (uint32_t)j >> p.lg_nx;
                ok = mbar_wait_warp(&x_full[xb], xuse & 1u) && ok;
                if (elect_one()) mbar_arrive_remote(px_remote + xb * 8u);
#pragma unroll 1
                for (int i = 0; i < n_stages; ++i) {
                    ok = mbar_wait_warp(&full[s], ph) && ok;
                    if (elect_one()) mbar_arrive_remote(pf_remote + s * 8u);
                    if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1u; }
                }
            }
        }
        pdl_trigger();            // the next kernel of the stream may start its prologue while the last epilogue runs
        TL2(4);
        if (!ok && lane == 0) atomicCAS(&g_watchdog, 0, (3 << 16) | (int)(blockIdx.x & 0xffff));
    } else {
        // ===== epilogue: TMEM -> registers -> (+bias | mask | affine) -> global (+ per-tile partial reductions) =====
        bool dead = false;
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int npc = np / 8;
        const size_t chunk_stride = (size_t)p.L * 8;
        const bool do_stat = p.stat_partial != nullptr;
        const bool do_red = p.red_partial != nullptr;
        constexpr bool do_aff = AFF;
        const int na = p.na;
        for (int j = 0; j < n_my; ++j) {
            const int tile = NCTA * (pair + j * n_pairs_grid) + (int)rank;
            const bool real = tile < p.n_tiles;                 // an odd tile count leaves the last pair's peer without a tile
            const int b = real ? tile / p.ltiles : 0, l0 = real ? (tile - b * p.ltiles) * 128 : 0;
            const int l = l0 + row;
            const bool valid = real && l < p.L;
            const size_t row_off = ((size_t)b * npc * p.L + (size_t)(valid ? l : 0)) * 8;
            float* ybase = p.y + row_off;
            const float* mbase = do_red ? p.mask_y + row_off : nullptr;
            const int ab = j & (na - 1), ause = j >> p.lg_na;
            const uint32_t t_acc = tmem_base + (uint32_t)(ab * p.acc_stride) + ((uint32_t)(q * 32) << 16);
            // one thread polls the accumulator barrier; the other 127 sleep in a named barrier instead of spinning on
            // mbarrier.try_wait next to the MMA issuer
            if (threadIdx.x == 64) { if (DBG && (p.debug & 256)) mbar_wait_sleep(&acc_full[ab], (uint32_t)(ause & 1), dead, 4, 500); else mbar_wait(&acc_full[ab], (uint32_t)(ause & 1), dead, 4); }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            tc_fence_after();
            if (j == 0 && warp == 2 && lane == 0) TL2(5);
            for (int c0 = 0; c0 < np; c0 += 32) {
                float v[32];
                const bool wide = c0 + 32 <= np;          // np is a multiple of 16: the tail chunk is 16 wide
                if (wide) {
                    tmem_ld32(t_acc + (uint32_t)c0, v);
                } else {
                    tmem_ld_x16(t_acc + (uint32_t)c0, v);
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
                                for (int k = 0; k < 8; ++k) yv[k] = 0.f;
                            }
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const int c = c0 + g * 8 + k;
                                const float z = fmaf(yv[k], bias_s[c], bias_s[np + c]);
                                if (!(z > 0.f)) v[g * 8 + k] = 0.f;
                                yh[g * 8 + k] = (yv[k] - bias_s[2 * np + c]) * bias_s[3 * np + c];
                            }
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
                            for (int k = 0; k < 8; ++k) rv[k] = 0.f;
                            if (p.aff_res && valid) {
                                const float* src = p.aff_res + row_off + (size_t)((c0 >> 3) + g) * chunk_stride;
                                const float4 a0 = __ldg(reinterpret_cast<const float4*>(src));
                                const float4 a1 = __ldg(reinterpret_cast<const float4*>(src + 4));
                                rv[0] = a0.x; rv[1] = a0.y; rv[2] = a0.z; rv[3] = a0.w; rv[4] = a1.x; rv[5] = a1.y; rv[6] = a1.z; rv[7] = a1.w;
                            }
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const int c = c0 + g * 8 + k;
                                float z = fmaf(v[g * 8 + k], bias_s[np + c], bias_s[2 * np + c]) + rv[k];
                                if (p.aff_relu) z = fmaxf(z, 0.f);
                                v[g * 8 + k] = z;
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
                                    for (int k = 0; k < 4; ++k) h2[k] = __floats2bfloat162_rn(v[g * 8 + 2 * k], v[g * 8 + 2 * k + 1]);
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
                        // TSC_OUT_POOLED (one tile per sample, L <= 128): column sums over this warp's valid rows
                        float s1[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) s1[i] = valid ? v[i] : 0.f;
                        const float a = colsum32(s1, lane);
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
                    const float a = colsum32(s1, lane);
                    const float c = colsum32(s2, lane);
                    if (c0 + lane < np) wstat[q * np + c0 + lane] = make_float2(a, c);
                } else if (do_stat) {
                    // two-pass per warp: column means first, then the centred sums of squares (no cancellation even when
                    // |mean| >> std); lane j ends up with (sum, M2) of column c0 + j over this warp's valid rows
                    const int nw = max(0, min(32, min(128, p.L - l0) - q * 32));
                    float s1[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) s1[i] = valid ? v[i] : 0.f;
                    const float a = colsum32(s1, lane);
                    const float mean_l = nw > 0 ? a / (float)nw : 0.f;
                    float s2[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float m = __shfl_sync(0xffffffffu, mean_l, i);
                        const float dlt = valid ? v[i] - m : 0.f;
                        s2[i] = dlt * dlt;
                    }
                    const float c = colsum32(s2, lane);
                    if (c0 + lane < np) wstat[q * np + c0 + lane] = make_float2(mean_l, c);
                }
            }
            // the accumulator tile has been read: hand it back to the issuer, then finish the per-tile reductions
            tc_fence_before();
            asm volatile("bar.sync 1, 128;" ::: "memory");             // the four epilogue warps (also orders wstat)
            if (threadIdx.x == 64) {
                if (!PAIR || rank == 0) mbar_arrive(&acc_empty[ab]);
                else mbar_arrive_remote(mapa_u32(smem_u32(&acc_empty[ab]), 0));
            }
            if (!real) {
                // nothing to write for a missing tile
            } else if (do_aff && p.aff_kind == TSC_OUT_POOLED) {
                const float inv_l = 1.f / (float)p.L;
                for (int c = threadIdx.x - 64; c < p.nbias; c += 128) {
                    float a = 0.f;
#pragma unroll
                    for (int w = 0; w < 4; ++w) a += wstat[w * np + c].x;
                    reinterpret_cast<float*>(p.aff_out)[(size_t)b * p.nbias + c] = a * inv_l;      // AdaptiveAvgPool1d(1)
                }
            } else if (do_stat || do_red) {
                const int rows_cta = min(128, p.L - l0);
                for (int c = threadIdx.x - 64; c < np; c += 128) {
                    if (do_red) {
                        float a = 0.f, d = 0.f;
#pragma unroll
                        for (int w = 0; w < 4; ++w) { const float2 t = wstat[w * np + c]; a += t.x; d += t.y; }
                        reinterpret_cast<float2*>(p.red_partial)[(size_t)tile * np + c] = make_float2(a, d);
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
                        reinterpret_cast<float2*>(p.stat_partial)[(size_t)tile * np + c] = make_float2(mean, m2);
                    }
                }
            }
            if (j + 1 < n_my) asm volatile("bar.sync 1, 128;" ::: "memory");     // wstat is rewritten by the next tile
            if (j == 0 && warp == 2 && lane == 0) TL2(6);
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); // the peer's shared memory and TMEM stay alive until the leader's last MMA has completed
    else __syncthreads();
    if (warp == 1) { if (PAIR) tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
    if (warp == 1 && lane == 0) TL2(7);
}

int make_c8_map(CUtensorMap* map, const void* base, int B, int kc, int L, int box_rows, int box_chunks);   // conv_tc.cu

static inline int conv2_rp(int Kmax) { return (128 + Kmax - 1 + 7) & ~7; }
static int knob2_stage_bytes();
static int knob2_debug();
static int knob2_dual();

// ---- the schedule: runs + stages, built once per bank geometry and direction, cached ----------------------------------------
static int build_sched(bool pair, int direction, int Cin, int Cout, int Kmax, const int* s_of_tap, ConvSched* sc) {
    TapTable tt;
    if (build_tap_table(direction, Cin, Cout, Kmax, s_of_tap, &tt) != 0) return -1;
    const int Rp = conv2_rp(Kmax);
    memset(sc, 0, sizeof(*sc));
    // a large activation tile leaves less room for the weight ring: halve the stage so that two stages still fit the
    // half-SM shared-memory budget (two CTAs of different launches can then share an SM)
    sc->stage_bytes = (tt.kc * Rp * 16 > 48 * 1024 ? knob2_stage_bytes() / 2 : knob2_stage_bytes()) * (pair ? 1 : 2);
    int n_runs = 0, n_stages = 0;
    uint32_t used = 0;            // bytes of the open stage
    uint32_t stage_src = 0;       // its source offset (16 B units)
    bool init_done = false;
    int stage_first_run = 0;
    auto close_stage = [&]() -> int {
        TSC_REQUIRE(n_stages < C2_MAX_STAGES, "kernel bank needs more than %d weight stages of %d B: unsupported", C2_MAX_STAGES,
                    sc->stage_bytes);
        sc->stages[n_stages++] = make_uint4(stage_src, used, (uint32_t)stage_first_run, (uint32_t)(n_runs - stage_first_run));
        sc->runs[2 * (n_runs - 1) + 1].w |= RF_LAST;
        stage_first_run = n_runs;
        used = 0;
        return 0;
    };
    for (int oi = 0; oi < tt.n_order; ++oi) {
        const int t = tt.order[oi];
        const uint32_t n_lo = (uint32_t)tt.n_lo[t], kc_lo = (uint32_t)tt.kc_lo[t];
        const uint32_t nt = (uint32_t)tt.np - n_lo, ksteps = ((uint32_t)tt.kc - kc_lo) / 2;
        const uint32_t rows = pair ? nt / 2 : nt;            // rows of the B operand in one CTA
        const uint32_t step_bytes = 2 * rows * 16;           // per CTA: two 8-channel chunks of `rows` rows
        TSC_REQUIRE(step_bytes <= (uint32_t)sc->stage_bytes, "one MMA's weights (%u B) exceed the %d B stage", step_bytes, sc->stage_bytes);
        uint32_t kp = 0;
        while (kp < ksteps) {
            if (used + step_bytes > (uint32_t)sc->stage_bytes) { if (close_stage() != 0) return -1; }
            uint32_t fit = ((uint32_t)sc->stage_bytes - used) / step_bytes;
            uint32_t n = ksteps - kp < fit ? ksteps - kp : fit;
            if (!init_done) n = 1;                           // the first MMA of a tile overwrites the accumulator: a run of its own
            TSC_REQUIRE(n_runs < C2_MAX_RUNS - 1, "kernel bank needs more than %d issue runs: unsupported", C2_MAX_RUNS - 1);
            if (used == 0) stage_src = pair ? (uint32_t)tt.w_off[t] / 2 + kp * nt       // in this CTA's stream (packed_row())
                                            : (uint32_t)tt.w_off[t] + 2 * kp * nt;
            uint4 e, f;
            e.x = (kc_lo + 2 * kp) * (uint32_t)Rp + (uint32_t)((knob2_debug() & 2) ? (t & ~7) : t);    // (debug 2: taps aligned to 128 B -- garbage results, timing experiment)
            e.y = (used >> 4) | (rows << 16);
            e.z = (1u << 4) | (1u << 7) | (1u << 10) | ((nt >> 3) << 17) | (((pair ? 256u : 128u) >> 4) << 24);
            e.w = n_lo;
            f.x = n;
            f.y = 2 * rows;                                  // two chunks of `rows` rows per K step
            f.z = init_done ? 1u : 0u;
            f.w = used == 0 ? RF_FIRST : 0u;
            // timing experiments only (TSC_C2_DEBUG bits, garbage results): which property of the MMA stream costs issue time?
            //   16: every MMA reads the same A rows (no tap shift, no K walk over the tile's chunks beyond the kernel's own step)
            //   32: every MMA reads the same B rows (start of the weight slot)      64: every MMA has N = np at column 0
            //  128: every MMA has N = 112 at column 0
            {
                const int dbg = knob2_debug();
                if (dbg & 16) e.x = 0;
                if (dbg & (64 | 128)) {
                    const uint32_t nn = (dbg & 64) ? (uint32_t)tt.np : 112u, rr = pair ? nn / 2 : nn;
                    e.y = (e.y & 0xffffu) | (rr << 16);
                    e.z = (1u << 4) | (1u << 7) | (1u << 10) | ((nn >> 3) << 17) | (((pair ? 256u : 128u) >> 4) << 24);
                    e.w = 0;
                    f.y = 2 * rr;
                }
                if (dbg & 32) { e.y &= 0xffff0000u; f.y = 0; }
            }
            sc->runs[2 * n_runs] = e;
            sc->runs[2 * n_runs + 1] = f;
            ++n_runs;
            init_done = true;
            used += n * step_bytes;
            kp += n;
        }
    }
    TSC_REQUIRE(n_runs > 0, "kernel bank has no live tap");
    if (close_stage() != 0) return -1;
    sc->n_runs = n_runs;
    sc->n_stages = n_stages;
    sc->half_rows16 = tt.total_rows / 2;
    return 0;
}

struct SchedKey {
    int pair, direction, Cin, Cout, Kmax;
    std::vector<int> s;
    bool operator<(const SchedKey& o) const {
        if (pair != o.pair) return pair < o.pair;
        if (direction != o.direction) return direction < o.direction;
        if (Cin != o.Cin) return Cin < o.Cin;
        if (Cout != o.Cout) return Cout < o.Cout;
        if (Kmax != o.Kmax) return Kmax < o.Kmax;
        return s < o.s;
    }
};

// Host-side cache of the schedules (the only state of this file besides the attribute opt-in; behind a mutex).
static const ConvSched* get_sched(bool pair, int direction, int Cin, int Cout, int Kmax, const int* s_of_tap) {
    static std::mutex mu;
    static std::map<SchedKey, std::unique_ptr<ConvSched>> cache;
    SchedKey key{pair ? 1 : 0, direction, Cin, Cout, Kmax, std::vector<int>(s_of_tap, s_of_tap + (Kmax > 0 && Kmax <= TSC_MAX_TAPS ? Kmax : 0))};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second.get();
    std::unique_ptr<ConvSched> sc(new ConvSched);
    if (build_sched(pair, direction, Cin, Cout, Kmax, s_of_tap, sc.get()) != 0) return nullptr;
    const ConvSched* out = sc.get();
    cache.emplace(std::move(key), std::move(sc));
    return out;
}

static long long* g_timeline2 = nullptr;     // debug only: device buffer of >= 8 clock64 samples

// experiment knobs (environment, read once): TSC_C2_STAGE_KB, TSC_C2_SMEM_FULL, TSC_C2_GRID, TSC_C2_DUAL (0: one tile per
// pass), TSC_C2_DEBUG (timing experiments with garbage results; selects the instrumented instantiation)
static int env_int2(const char* name, int dflt) {
    const char* e = getenv(name);
    return e && e[0] ? atoi(e) : dflt;
}
static int knob2_stage_bytes() { static const int v = env_int2("TSC_C2_STAGE_KB", C2_STAGE_BYTES / 1024) * 1024; return v; }
static int knob2_smem_full() { static const int v = env_int2("TSC_C2_SMEM_FULL", 0); return v; }
static int knob2_grid() { static const int v = env_int2("TSC_C2_GRID", 0); return v; }
static int knob2_debug() { static const int v = env_int2("TSC_C2_DEBUG", 0); return v; }
static int knob2_dual() { static const int v = env_int2("TSC_C2_DUAL", 1); return v; }

}  // namespace tc

void set_conv2_timeline(long long* dev) { tc::g_timeline2 = dev; }

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
            n = v;
        else
            n = 148;
    }
    return n;
}

int osconv2_tc(int direction, const void* x, int dtype, const void* w, const float* bias, float* y,
               const tsc_conv_epilogue* epi, int B, int L, int Cin, int Cout, int Kmax, const int* s_of_tap, cudaStream_t cs) {
    using namespace tc;
    TSC_REQUIRE(dtype == TSC_BF16, "the tcgen05 engine takes bf16 operands");
    const bool pair = packed_split();       // the CTA-pair variant (and its split packed layout): an opt-in experiment
    const ConvSched* sc = get_sched(pair, direction, Cin, Cout, Kmax, s_of_tap);
    if (!sc) return -1;
    const bool fwd = direction == TSC_DIR_FWD;
    Conv2Params p;
    memset(&p, 0, sizeof(p));
    p.w = (const __nv_bfloat16*)w;
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
                            "the pooled inference epilogue needs L <= 128 (one tile per sample), got %d", L);
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
    p.n_tiles = B * p.ltiles;
    p.np = fwd ? pad16(Cout) : pad16(Cin);
    p.kc = fwd ? pad16(Cin) / 8 : pad16(Cout) / 8;
    p.pad_left = fwd ? (Kmax - 1) / 2 : Kmax / 2;
    p.Rp = conv2_rp(Kmax);
    p.x_bytes = p.kc * p.Rp * 16;
    p.acc_stride = p.np <= 32 ? 32 : p.np <= 64 ? 64 : p.np <= 128 ? 128 : 256;
    const int nsm = sm_count();
    // one tile pair per CTA pair while the pairs fit one wave (single-buffered, half-SM budgets: CTAs of two streams share an
    // SM); otherwise one persistent CTA pair per SM pair with double-buffered accumulator (and, if it fits, activation) tiles
    const int ncta = pair ? 2 : 1;
    const int n_pairs = (p.n_tiles + ncta - 1) / ncta, max_pairs = nsm / ncta;     // units = tiles, or tile pairs
    const bool persistent = n_pairs > max_pairs;
    int grid_pairs = persistent ? max_pairs : n_pairs;
    if (persistent && knob2_grid() > 0 && knob2_grid() < grid_pairs) grid_pairs = knob2_grid();
    const int grid = ncta * grid_pairs;
    // persistent CTAs: two tiles per pass over the bank (not for CTA pairs); four accumulator tiles when they fit TMEM, so
    // that the epilogues of one tile pair overlap the MMAs of the next
    // (measured, profiles/r2_conv2_dual.md: pays when the epilogues still overlap -- four accumulator tiles, np <= 128 -- or
    // when the bank is large, >= 256 KB per pass; a small bank behind a wide accumulator loses the overlap for nothing)
    const size_t bank_bytes = (size_t)sc->half_rows16 * 32;
    p.dual = (persistent && !pair && knob2_dual() && (p.acc_stride <= 128 || bank_bytes >= 256 * 1024)) ? 1 : 0;
    p.na = persistent ? ((p.dual && p.acc_stride <= 128) ? 4 : 2) : 1;
    p.tiles_base = n_pairs / grid_pairs;
    p.tiles_rem = n_pairs % grid_pairs;
    p.half_rows16 = sc->half_rows16;
    p.off_bias = C2_HDR;
    p.off_wstat = p.off_bias + 4 * p.np * 4;
    p.off_xs = (p.off_wstat + 4 * p.np * 8 + 127) & ~127;
    const int slot = sc->stage_bytes;
    const int cap_full = 227 * 1024, cap_half = 113 * 1024;
    auto layout = [&](int nx, int cap, int* ns_out) {
        p.off_stages = (p.off_xs + nx * p.x_bytes + 127) & ~127;
        int ns = (cap - p.off_stages) / slot;
        if (ns > 8) ns = 8;
        if (ns > sc->n_stages) ns = sc->n_stages;
        *ns_out = ns;
        return ns >= 2 || (ns == 1 && sc->n_stages == 1);
    };
    int ns = 0;
    if (persistent) {
        p.nx = 2;
        if (p.dual && layout(4, cap_full, &ns) && ns >= 3) {
            p.nx = 4;                            // the next pair's activation tiles land while this pair is being multiplied
        } else if (!layout(2, cap_full, &ns)) {
            p.nx = 1;
            p.dual = 0;                          // (a tile pair needs two activation buffers)
            if (p.na == 4) p.na = 2;
            TSC_REQUIRE(layout(1, cap_full, &ns), "shape needs %d B of shared memory before the weight stages: unsupported", p.off_stages);
        }
    } else {
        p.nx = 1;
        if (knob2_smem_full() || !layout(1, cap_half, &ns)) TSC_REQUIRE(layout(1, cap_full, &ns), "shape needs %d B of shared memory before the weight stages: unsupported", p.off_stages);
    }
    p.NS = ns;
    p.tmem_cols = p.acc_stride * p.na;
    p.lg_nx = p.nx == 4 ? 2 : p.nx == 2 ? 1 : 0;
    p.lg_na = p.na == 4 ? 2 : p.na == 2 ? 1 : 0;
    p.tl = g_timeline2;
    p.debug = knob2_debug();
    CUtensorMap xmap;
    if (make_c8_map(&xmap, x, B, p.kc, L, p.Rp, p.kc) != 0) return -1;
    const int smem = p.off_stages + ns * slot;
    static OnceAttr attr_once;             // once per process: opt in to the full 227 KB of dynamic shared memory
    {
        const cudaError_t e = run_once(attr_once, [] {
            cudaError_t r = cudaSuccess;
#define TSC_C2_OPTIN(A, D, P) if (r == cudaSuccess) r = cudaFuncSetAttribute(osconv2_kernel<A, D, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
            TSC_C2_OPTIN(false, false, false); TSC_C2_OPTIN(true, false, false); TSC_C2_OPTIN(false, true, false); TSC_C2_OPTIN(true, true, false);
            TSC_C2_OPTIN(false, false, true); TSC_C2_OPTIN(true, false, true); TSC_C2_OPTIN(false, true, true); TSC_C2_OPTIN(true, true, true);
#undef TSC_C2_OPTIN
            return r;
        });
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    // a cluster of two CTAs (the tcgen05 CTA pair) per tile pair
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(C2_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = cs;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)ncta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    const bool dbg = p.tl != nullptr || p.debug != 0;   // the instrumented instantiation only when a timeline buffer or an experiment knob is set
    cudaError_t le;
#define TSC_C2_LAUNCH(A, D, P) le = cudaLaunchKernelEx(&cfg, osconv2_kernel<A, D, P>, xmap, *sc, p)
    if (pair) {
        if (p.aff_out) { if (dbg) TSC_C2_LAUNCH(true, true, true); else TSC_C2_LAUNCH(true, false, true); }
        else           { if (dbg) TSC_C2_LAUNCH(false, true, true); else TSC_C2_LAUNCH(false, false, true); }
    } else {
        if (p.aff_out) { if (dbg) TSC_C2_LAUNCH(true, true, false); else TSC_C2_LAUNCH(true, false, false); }
        else           { if (dbg) TSC_C2_LAUNCH(false, true, false); else TSC_C2_LAUNCH(false, false, false); }
    }
#undef TSC_C2_LAUNCH
    if (le != cudaSuccess) { set_error("osconv2 launch: %s", cudaGetErrorString(le)); return (int)le; }
    TSC_LAUNCH_CHECK();
    return 0;
}

int read_clear_watchdog_conv2(int* code) {
    int zero = 0;
    cudaError_t e = cudaMemcpyFromSymbol(code, tc::g_watchdog, sizeof(int));
    if (e != cudaSuccess) return (int)e;
    return (int)cudaMemcpyToSymbol(tc::g_watchdog, &zero, sizeof(int));
}

}  // namespace tsc
