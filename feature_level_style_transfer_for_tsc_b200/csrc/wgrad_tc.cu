// tcgen05 engine, weight gradient of the omni-scale convolution on live taps only.
//
//   dW[co, ci, t] = sum_{b,l} dY[b, co, l] * X[b, ci, l + t - pad_left]          
//
// One GEMM per tap with  out channels (128 per tile) -> MMA M,  in channels (<= 256) -> MMA N,
// positions -> MMA K (16 per instruction).  In the c8 layout both operands are "MN-major" for this
// contraction (8 channels contiguous, positions 16 B apart), and the tap is a row offset on the
// B-operand descriptor, so ONE staged (dY tile, X tile + halo) pair feeds every tap of a CTA.
//
// Work decomposition: a work item = (128-channel out tile, up to NT consecutive taps) with
// NT = 256 / Cin_p TMEM accumulators of [128 x Cin_p] fp32 each (half of TMEM, so that two CTAs can share an SM); a CTA owns one item and a share of the
// (b, l) position tiles, accumulates in TMEM over all of them, and writes its partial dW once.
// Out tiles are anchored at the END of the channel axis: by nestedness of the kernel bank the outer
// taps are live only on a channel suffix, so they need the last tile only (masked taps cost nothing).
//
// Warp roles: warps 1-2 = MMA issuers (taps alternate between them: with N = Cin_p <= 128 one issuing thread is the
// limiter, ~73 cycles per instruction against N/2 cycles of tensor work -- tools/mma_bench2.cu), warps 3-6 = tile
// producers during the main loop (LDGSTS with zero fill: a c8 row is 16 B, and a TMA box of 16 B rows retires
// about one row per cycle, slower than 128 threads issuing coalesced 16 B copies), then the epilogue.
// Partials are CTA-private blocks [split][item][tap][ci][128 rows] (row-fastest, so every store instruction of a
// warp writes 128 contiguous bytes); wgrad_tc_reduce_kernel adds the splits in a fixed order (deterministic, no
// float atomics), writes the exact zeros of the masked taps and can accumulate into dW (flat gradient bucket).
// Replaces cuDNN/oneDNN wgrad of OS_CNN/OS_CNN.py:71.
#include "tc_common.cuh"
#include <algorithm>

namespace tsc {
namespace tc {

static constexpr int WG_THREADS = 224;
static constexpr int WG_LT = 128;          // positions per stage
static constexpr int WG_MAX_ITEMS = 192;
static constexpr int WG_HDR = 256;
static constexpr int WG_STAGES = 2;
static constexpr int WG_TMEM_COLS = 256;     // half of TMEM: CTAs of two independent launches can be co-resident
static constexpr int WG_SMEM_HALF = 113 * 1024;

// TMEM columns of a CTA = accumulators (taps) per work item x padded in channels.  Half of TMEM when two CTAs can share
// an SM; all of it when the shared-memory ring of the shape rules that out anyway (Cin > 128: the 144 -> 72 bank of the
// classifier then stages one (dY, X) tile pair per THREE taps instead of per tap -- it was the slowest launch of the
// cfg2 step, 115 us, profiles/README.md session 3).
static inline int wg_stage_bytes(int cinp, int NT) { return 16 * WG_LT * 16 + (cinp / 8) * ((WG_LT + NT - 1 + 7) & ~7) * 16; }
static inline int wg_tmem_cols(int cinp) {
    const int nt_half = WG_TMEM_COLS / cinp;
    if (nt_half >= 1 && WG_HDR + WG_STAGES * wg_stage_bytes(cinp, nt_half) <= WG_SMEM_HALF) return WG_TMEM_COLS;
    return 512;
}

struct WgItem { short m0, t0, nt, pad; };
struct WgItems { int n; WgItem it[WG_MAX_ITEMS]; };
// tile (0: channels below m_split, 1: the rest) x tap -> work item (or -1: masked tap, gradient is zero)
struct WgLookup { short item_of[2][TSC_MAX_TAPS]; };

struct WgParams {
    float* part;        // [S][items][NT][cinp][128]
    const __nv_bfloat16* dy;   // c8 [B][np/8][L][8]
    const __nv_bfloat16* x;    // c8 [B][kcx][L][8]
    int B, L, ltiles;
    int taps, pad_left;
    int np;             // padded out channels
    int kcx;            // in-channel chunks
    int RX;             // X rows per chunk in shared memory
    int S;              // position splits
    int NT;             // accumulators (taps) per item
    int stage_bytes;
    int tmem_cols;      // 256 or 512 (wg_tmem_cols)
    long long* tl;      // optional phase timeline of CTA (0,0) (tsc_debug_set_timeline), NULL in production
};

#define WTL(i) do { if (p.tl && blockIdx.x == 0 && blockIdx.y == 0) p.tl[i] = clock64(); } while (0)

__global__ void __launch_bounds__(WG_THREADS, 1)
oswgrad_tc_kernel(const __grid_constant__ WgItems items, const WgParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);        // [2]
    uint64_t* empty = full + WG_STAGES;                        // [2], two arrivals each (one per issuer warp)
    uint64_t* acc_full = empty + WG_STAGES;                    // two arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    uint8_t* stages = smem + WG_HDR;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WgItem item = items.it[blockIdx.x];
    const int sp = blockIdx.y;
    const int cinp = p.kcx * 8;
    const long long ntile = (long long)p.B * p.ltiles;
    const int tile0 = (int)(ntile * sp / p.S), tile1 = (int)(ntile * (sp + 1) / p.S);
    const int dy_bytes = 16 * WG_LT * 16;                       // the A tile always spans 16 chunks of smem

    if (warp == 0 && lane == 0) {
        WTL(0);
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], 128); mbar_init(&empty[i], 2); }
        mbar_init(acc_full, 2);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                   // the prologue above overlapped the tail of the previous kernel of the stream

    if (warp == 0) {
        // (idle: the tiles are staged by the four epilogue warps)
    } else if (warp <= 2) {
        // Each issuer warp walks the tiles in lock-step (uniform control flow) and one elected lane issues the 8 MMAs
        // of a (tile, tap) per elect block: a lane-0-only branch makes ptxas wrap every UTCHMMA in a per-thread ELECT
        // loop (tools/mma_bench3.cu).  Warp 1 takes the even taps of the item, warp 2 the odd ones.
        bool dead = false;
        const int wi = warp - 1;
        const uint32_t idesc = make_idesc_bf16(128, (uint32_t)cinp, true, true, false);
        const uint32_t st16 = smem_u32(stages) >> 4, stage16 = (uint32_t)p.stage_bytes >> 4;
        // MN-major, SWIZZLE_NONE: LBO = 128 B between 8-position groups, SBO = chunk stride
        const uint32_t a_hi = ((uint32_t)(WG_LT * 16) >> 4) | (1u << 14);
        const uint32_t b_hi = ((uint32_t)(p.RX * 16) >> 4) | (1u << 14);
        const uint32_t lbo = (128u >> 4) << 16;
        uint32_t s = 0, ph = 0, acc = 0;
        const int nt = item.nt;
        for (int tile = tile0; tile < tile1; ++tile) {
            mbar_wait(&full[s], ph, dead, 6);
            __syncwarp();             // lanes leave the polling loop at different times: reconverge before the elect
            fence_proxy_async();      // the tile was written by cp.async (generic proxy), the MMA reads it through the async proxy
            tc_fence_after();
            if (wi == 0 && lane == 0 && p.tl && blockIdx.x == 0 && blockIdx.y == 0 && tile - tile0 < 24)
                p.tl[8 + tile - tile0] = clock64();
            const uint32_t a0 = (st16 + s * stage16) | lbo;
            const uint32_t b0 = (st16 + s * stage16 + ((uint32_t)dy_bytes >> 4)) | lbo;
            for (int tl = wi; tl < nt; tl += 2) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(tl * cinp);
                if (elect_one()) {
#pragma unroll
                    for (int k16 = 0; k16 < WG_LT / 16; ++k16) {
                        umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | (a0 + (uint32_t)(k16 * 16)),
                                  ((uint64_t)b_hi << 32) | (b0 + (uint32_t)(tl + k16 * 16)), idesc, acc | (uint32_t)(k16 > 0));
                    }
                }
            }
            __syncwarp();
            if (elect_one()) tc_commit(&empty[s]);      // this warp's MMAs on the stage are done reading it
            acc = 1;
            if (++s == WG_STAGES) { s = 0; ph ^= 1u; }
        }
        __syncwarp();
        if (elect_one()) tc_commit(acc_full);
        pdl_trigger();
        if (wi == 0 && lane == 0) WTL(4);
    } else {
        bool dead = false;
        {
            // ===== producer phase: the four epilogue warps stage (dY tile, X tile + halo) with LDGSTS, zero-filling the
            // rows outside the series (ConstantPad1d) and the chunks past the channel axis =====
            const int ptid = threadIdx.x - 96;
            uint32_t s = 0, ph = 0;
            for (int tile = tile0; tile < tile1; ++tile) {
                const int b = tile / p.ltiles, l0 = (tile % p.ltiles) * WG_LT;
                mbar_wait(&empty[s], ph ^ 1u, dead, 5);
                uint8_t* st = stages + (size_t)s * p.stage_bytes;
                stage_c8_tile(st, p.dy, b, p.np / 8, p.L, item.m0 / 8, 16, l0, WG_LT, ptid, 128);
                stage_c8_tile(st + dy_bytes, p.x, b, p.kcx, p.L, 0, p.kcx, l0 + item.t0 - p.pad_left, p.RX, ptid, 128);
                cp_async_arrive_noinc(&full[s]);
                if (++s == WG_STAGES) { s = 0; ph ^= 1u; }
            }
        }
        if (threadIdx.x == 96) mbar_wait(acc_full, 0, dead, 7);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        tc_fence_after();
        if (threadIdx.x == 96) WTL(5);
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        // CTA-private block [tap][ci][128 rows]: the 32 lanes of a warp store 128 contiguous bytes per instruction
        float* dst = p.part + (((size_t)sp * items.n + blockIdx.x) * p.NT) * (size_t)cinp * 128 + row;
        const bool any = tile1 > tile0;                // a split without a position tile contributes zeros
        for (int tl = 0; tl < item.nt; ++tl) {
            for (int c0 = 0; c0 < cinp; c0 += 16) {
                float v[16];
                if (any) {
                    tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * cinp + c0), v);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
                float* d = dst + ((size_t)tl * cinp + c0) * 128;
#pragma unroll
                for (int i = 0; i < 16; ++i) d[(size_t)i * 128] = v[i];
            }
        }
        if (threadIdx.x == 96) WTL(6);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    if (warp == 1 && lane == 0) WTL(7);
}

// dW[co, ci, t] (+)= sum over splits of the partial of the item covering (tile(co), t); 0 on masked taps.
// One block = one channel tile (128 out channels) x one in channel x 8 consecutive taps; thread = out channel (a row of the
// partial block).  The partials are row-fastest, so a warp's load is 128 contiguous bytes and every byte of the workspace is
// read exactly once in full sectors (the first version read one 32 B sector per 8 threads: 1.3 TB/s out of L2 for the 19 MB of
// the 72->228 bank); all (tap, split) loads of a thread are independent, the sums stay in split order, and a thread writes its
// 8 taps contiguously (one 32 B sector of dW).
__global__ void __launch_bounds__(128) wgrad_tc_reduce_kernel(const float* __restrict__ part, float* __restrict__ dW,
                                                                const __grid_constant__ WgItems items,
                                                                const __grid_constant__ WgLookup lk,
                                                                const __grid_constant__ STable st, int S, int NT, int Cin,
                                                                int Cout, int Kmax, int np, int cinp, int m_split,
                                                                int accumulate) {
    pdl_trigger();
    pdl_wait();
    const int tile = blockIdx.z;
    const int ci = blockIdx.x;
    const int t_base = blockIdx.y * 8;
    const int row = threadIdx.x;
    const int m0 = tile == 1 ? np - 128 : 0;
    const int co = m0 + row;
    // tile 0 owns the channels below m_split (all of them when there is one tile), tile 1 the rest
    const bool mine = co < Cout && (tile == 0 ? (m_split == 0 || co < m_split) : co >= m_split);
    if (!mine) return;
    const size_t split_stride = (size_t)items.n * NT * cinp * 128;
    const float* src[8];
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int t = t_base + j;
        src[j] = nullptr;
        acc[j] = 0.f;
        if (t < Kmax && co >= st.s[t]) {
            const int it = lk.item_of[tile][t];
            if (it >= 0) src[j] = part + (((size_t)it * NT + (t - items.it[it].t0)) * cinp + ci) * 128 + row;
        }
    }
    // four splits x eight taps = 32 independent loads in flight per thread; the sums stay in split order
    int s = 0;
    for (; s + 4 <= S; s += 4) {
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) v[u][j] = src[j] ? __ldg(src[j] + (size_t)(s + u) * split_stride) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v[u][j];
    }
    for (; s < S; ++s) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (src[j]) acc[j] += __ldg(src[j] + (size_t)s * split_stride);
    }
    float* o = dW + ((size_t)co * Cin + ci) * Kmax + t_base;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (t_base + j >= Kmax) continue;
        if (!accumulate) o[j] = acc[j];                 // masked taps: the exact zero
        else if (src[j]) o[j] += acc[j];                // accumulate: a masked tap adds nothing, leave it untouched
    }
}

int make_c8_map(CUtensorMap* map, const void* base, int B, int kc, int L, int box_rows, int box_chunks);   // conv_tc.cu
static long long* g_wg_timeline = nullptr;

}  // namespace tc

static int wgrad_tc_items(int Cin, int Cout, int Kmax, const int* s_of_tap, tc::WgItems* items, tc::WgLookup* lk,
                          int* m_split) {
    using namespace tc;
    const int np = pad16(Cout), cinp = pad16(Cin);
    const int NT = tc::wg_tmem_cols(cinp) / cinp;
    const int MT = np > 128 ? 2 : 1;
    *m_split = MT == 2 ? np - 128 : 0;
    items->n = 0;
    for (int a = 0; a < 2; ++a)
        for (int t = 0; t < TSC_MAX_TAPS; ++t) lk->item_of[a][t] = -1;
    for (int mt = 0; mt < MT; ++mt) {
        const int m0 = (MT == 2 && mt == 1) ? np - 128 : 0;
        // taps that need this tile: the last tile serves every live tap, tile 0 (of two) the taps whose
        // live suffix starts below the last tile
        const int limit = (mt == MT - 1) ? Cout : *m_split;
        int t = 0;
        while (t < Kmax) {
            if (s_of_tap[t] >= limit) { ++t; continue; }
            int n = 0;
            while (t + n < Kmax && n < NT && s_of_tap[t + n] < limit) ++n;
            TSC_REQUIRE(items->n < WG_MAX_ITEMS, "too many wgrad work items");
            for (int j = 0; j < n; ++j) lk->item_of[MT == 2 ? mt : 0][t + j] = (short)items->n;
            items->it[items->n++] = WgItem{(short)m0, (short)t, (short)n, 0};
            t += n;
        }
    }
    TSC_REQUIRE(items->n > 0, "kernel bank has no live tap");
    return 0;
}

// upper bound on the number of work items without the tap table: 2 tiles x ceil(Kmax / NT)
static int wgrad_tc_max_items(int Cin, int Cout, int Kmax) {
    const int cinp = pad16(Cin), NT = tc::wg_tmem_cols(cinp) / cinp, MT = pad16(Cout) > 128 ? 2 : 1;
    return MT * cdiv(Kmax, NT);
}

// position splits for `items` work items: one CTA per SM, a single wave (rounded down), at most 32 splits
static int wgrad_tc_splits_for(int items, int B, int L) {
    const int ntile = B * cdiv(L, tc::WG_LT);
    int s = 148 / items;
    if (s > ntile) s = ntile;
    if (s > 32) s = 32;
    if (s < 1) s = 1;
    return s;
}

int wgrad_tc_splits(int B, int L, int Cin, int Cout, int Kmax) {
    return wgrad_tc_splits_for(wgrad_tc_max_items(Cin, Cout, Kmax), B, L);
}

// The launch sizes its splits from the ACTUAL number of work items (masked taps of the low channel tile drop out: 15
// items instead of the bound of 22 for the 72 -> 228 bank), so the workspace is sized for the most CTA-private blocks
// any launch can write: max(148, items) of them.
size_t wgrad_tc_workspace_bytes(int B, int L, int Cin, int Cout, int Kmax) {
    const int cinp = pad16(Cin), NT = tc::wg_tmem_cols(cinp) / cinp;
    const int blocks = std::max(148, wgrad_tc_max_items(Cin, Cout, Kmax));
    (void)B; (void)L;
    return (size_t)blocks * NT * cinp * 128 * sizeof(float);
}

int oswgrad_tc(const void* dy, const void* x, int dtype, float* dW, void* workspace, int accumulate, int B, int L, int Cin,
               int Cout, int Kmax, const int* s_of_tap, cudaStream_t cs) {
    using namespace tc;
    TSC_REQUIRE(dtype == TSC_BF16, "the tcgen05 engine takes bf16 operands");
    WgItems items;
    WgLookup lk;
    WgParams p;
    int m_split = 0;
    if (wgrad_tc_items(Cin, Cout, Kmax, s_of_tap, &items, &lk, &m_split) != 0) return -1;
    const int np = pad16(Cout), cinp = pad16(Cin);
    p.tmem_cols = wg_tmem_cols(cinp);
    const int NT = p.tmem_cols / cinp;
    p.part = (float*)workspace;
    p.B = B; p.L = L; p.ltiles = cdiv(L, WG_LT);
    p.taps = Kmax; p.pad_left = (Kmax - 1) / 2;
    p.np = np; p.kcx = cinp / 8;
    p.RX = (WG_LT + NT - 1 + 7) & ~7;
    p.S = wgrad_tc_splits_for(items.n, B, L);
    p.NT = NT;
    p.tl = g_wg_timeline;
    p.stage_bytes = wg_stage_bytes(cinp, NT);
    const int smem = WG_HDR + WG_STAGES * p.stage_bytes;
    TSC_REQUIRE(smem <= 227 * 1024, "wgrad shape needs %d B of shared memory: unsupported", smem);
    p.dy = (const __nv_bfloat16*)dy;
    p.x = (const __nv_bfloat16*)x;
    static OnceAttr attr_once;             // once per process: opt in to the full 227 KB of dynamic shared memory
    {
        const cudaError_t e = run_once(attr_once, [] {
            return cudaFuncSetAttribute(oswgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        });
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    { cudaError_t le = launch_pdl(oswgrad_tc_kernel, dim3(items.n, p.S), dim3(WG_THREADS), (size_t)smem, cs, items, p); if (le != cudaSuccess) { set_error("oswgrad launch: %s", cudaGetErrorString(le)); return (int)le; } }
    TSC_LAUNCH_CHECK();
    STable st;
    fill_stable(&st, s_of_tap, Kmax);
    { cudaError_t le = launch_pdl(wgrad_tc_reduce_kernel, dim3(Cin, cdiv(Kmax, 8), m_split > 0 ? 2 : 1), dim3(128), 0, cs, (const float*)p.part,
                                  dW, items, lk, st, p.S, NT, Cin, Cout, Kmax, np, cinp, m_split, accumulate);
      if (le != cudaSuccess) { set_error("wgrad reduce launch: %s", cudaGetErrorString(le)); return (int)le; } }
    TSC_LAUNCH_CHECK();
    return 0;
}

void set_wgrad_timeline(long long* dev) { tc::g_wg_timeline = dev; }

int read_clear_watchdog_wgrad(int* code) {
    int zero = 0;
    cudaError_t e = cudaMemcpyFromSymbol(code, tc::g_watchdog, sizeof(int));
    if (e != cudaSuccess) return (int)e;
    return (int)cudaMemcpyToSymbol(tc::g_watchdog, &zero, sizeof(int));
}

}  // namespace tsc
