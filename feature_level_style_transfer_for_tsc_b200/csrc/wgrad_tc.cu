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
// NT = 512 / Cin_p TMEM accumulators of [128 x Cin_p] fp32 each; a CTA owns one item and a share of the
// (b, l) position tiles, accumulates in TMEM over all of them, and writes its partial dW once.
// Out tiles are anchored at the END of the channel axis: by nestedness of the kernel bank the outer
// taps are live only on a channel suffix, so they need the last tile only (masked taps cost nothing).
// Partials [split][tap][co][ci] are reduced in order by wgrad_reduce_kernel (deterministic, no atomics),
// which also writes the exact zeros of the masked taps.
// Replaces cuDNN/oneDNN wgrad of OS_CNN/OS_CNN.py:71.
#include "tc_common.cuh"

namespace tsc {
namespace tc {

static constexpr int WG_THREADS = 192;
static constexpr int WG_LT = 128;          // positions per stage
static constexpr int WG_MAX_ITEMS = 192;
static constexpr int WG_HDR = 256;
static constexpr int WG_STAGES = 2;

struct WgItem { short m0, t0, nt, pad; };
struct WgItems { int n; WgItem it[WG_MAX_ITEMS]; };

struct WgParams {
    float* part;        // [S][taps][np][kcp]
    int B, L, ltiles;
    int taps, pad_left;
    int np;             // padded out channels
    int kcx;            // in-channel chunks
    int RX;             // X rows per chunk in shared memory
    int S;              // position splits
    int stage_bytes;
    int m_split;        // channels below m_split belong to tile 0 (MT == 2), 0 when MT == 1
};

__global__ void __launch_bounds__(WG_THREADS, 1)
oswgrad_tc_kernel(const __grid_constant__ CUtensorMap dymap, const __grid_constant__ CUtensorMap xmap,
                  const __grid_constant__ WgItems items, const WgParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);        // [2]
    uint64_t* empty = full + WG_STAGES;                        // [2]
    uint64_t* acc_full = empty + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    uint8_t* stages = smem + WG_HDR;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WgItem item = items.it[blockIdx.x];
    const int sp = blockIdx.y;
    const int cinp = p.kcx * 8;
    const long long ntile = (long long)p.B * p.ltiles;
    const int tile0 = (int)(ntile * sp / p.S), tile1 = (int)(ntile * (sp + 1) / p.S);
    const int dy_chunks = min(16, p.np / 8 - item.m0 / 8);
    const int dy_bytes = 16 * WG_LT * 16;                       // the A tile always spans 16 chunks of smem

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&dymap);
        tma_prefetch_desc(&xmap);
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            bool dead = false;
            const uint32_t bytes = (uint32_t)((dy_chunks * WG_LT + p.kcx * p.RX) * 16);
            uint32_t s = 0, ph = 0;
            for (int tile = tile0; tile < tile1; ++tile) {
                const int b = tile / p.ltiles, l0 = (tile % p.ltiles) * WG_LT;
                mbar_wait(&empty[s], ph ^ 1u, dead, 5);
                uint8_t* st = stages + (size_t)s * p.stage_bytes;
                mbar_arrive_expect_tx(&full[s], bytes);
                for (int c = 0; c < dy_chunks; ++c)
                    tma_load_4d(st + (size_t)c * WG_LT * 16, &dymap, 0, l0, item.m0 / 8 + c, b, &full[s]);
                for (int c = 0; c < p.kcx; ++c)
                    tma_load_4d(st + dy_bytes + (size_t)c * p.RX * 16, &xmap, 0, l0 + item.t0 - p.pad_left, c, b, &full[s]);
                if (++s == WG_STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // lean issue loop (one thread issues at most one MMA per ~54 cycles): descriptors advance by adds
            bool dead = false;
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
                tc_fence_after();
                const uint32_t a0 = (st16 + s * stage16) | lbo;
                const uint32_t b0 = (st16 + s * stage16 + ((uint32_t)dy_bytes >> 4)) | lbo;
                uint32_t d_tmem = tmem_base;
                for (int tl = 0; tl < nt; ++tl) {
#pragma unroll
                    for (int k16 = 0; k16 < WG_LT / 16; ++k16) {
                        umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | (a0 + (uint32_t)(k16 * 16)),
                                  ((uint64_t)b_hi << 32) | (b0 + (uint32_t)(tl + k16 * 16)), idesc, acc | (uint32_t)(k16 > 0));
                    }
                    d_tmem += (uint32_t)cinp;
                }
                acc = 1;
                tc_commit(&empty[s]);
                if (++s == WG_STAGES) { s = 0; ph ^= 1u; }
            }
            tc_commit(acc_full);
        }
    } else {
        bool dead = false;
        mbar_wait(acc_full, 0, dead, 7);
        tc_fence_after();
        const int q = warp & 3;
        const int co = item.m0 + q * 32 + lane;
        // tile 0 of a two-tile layer owns co < m_split, the last tile owns the rest
        const bool mine = co < p.np && (item.m0 == 0 ? (p.m_split == 0 || co < p.m_split) : co >= p.m_split);
        const int kcp = cinp;
        if (tile1 > tile0) {
            for (int tl = 0; tl < item.nt; ++tl) {
                const int t = item.t0 + tl;
                float* dst = p.part + (((size_t)sp * p.taps + t) * p.np + co) * kcp;
                for (int c0 = 0; c0 < cinp; c0 += 16) {
                    float v[16];
                    tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * cinp + c0), v);
                    if (mine) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    }
                }
            }
        } else if (mine) {
            // this split had no position tile: its partial is zero
            for (int tl = 0; tl < item.nt; ++tl) {
                float* dst = p.part + (((size_t)sp * p.taps + item.t0 + tl) * p.np + co) * kcp;
                for (int c = 0; c < cinp; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int make_c8_map(CUtensorMap* map, const void* base, int B, int kc, int L, int box_rows, int box_chunks);   // conv_tc.cu

}  // namespace tc

static int wgrad_tc_items(int Cin, int Cout, int Kmax, const int* s_of_tap, tc::WgItems* items, int* m_split) {
    using namespace tc;
    const int np = pad16(Cout), cinp = pad16(Cin);
    const int NT = 512 / cinp;
    const int MT = np > 128 ? 2 : 1;
    *m_split = MT == 2 ? np - 128 : 0;
    items->n = 0;
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
            items->it[items->n++] = WgItem{(short)m0, (short)t, (short)n, 0};
            t += n;
        }
    }
    TSC_REQUIRE(items->n > 0, "kernel bank has no live tap");
    return 0;
}

int wgrad_tc_splits(int B, int L, int Cin, int Cout, int Kmax) {
    // upper bound on the number of work items without the tap table: 2 tiles x ceil(Kmax / NT)
    const int cinp = pad16(Cin), NT = 512 / cinp, MT = pad16(Cout) > 128 ? 2 : 1;
    const int items = MT * cdiv(Kmax, NT);
    const int ntile = B * cdiv(L, tc::WG_LT);
    int s = cdiv(148, items);
    if (s > ntile) s = ntile;
    if (s > 32) s = 32;
    if (s < 1) s = 1;
    return s;
}

int oswgrad_tc(const void* dy, const void* x, int dtype, float* dW, void* workspace, int B, int L, int Cin, int Cout,
               int Kmax, const int* s_of_tap, cudaStream_t cs) {
    using namespace tc;
    TSC_REQUIRE(dtype == TSC_BF16, "the tcgen05 engine takes bf16 operands");
    WgItems items;
    WgParams p;
    if (wgrad_tc_items(Cin, Cout, Kmax, s_of_tap, &items, &p.m_split) != 0) return -1;
    const int np = pad16(Cout), cinp = pad16(Cin);
    const int NT = 512 / cinp;
    p.part = (float*)workspace;
    p.B = B; p.L = L; p.ltiles = cdiv(L, WG_LT);
    p.taps = Kmax; p.pad_left = (Kmax - 1) / 2;
    p.np = np; p.kcx = cinp / 8;
    p.RX = (WG_LT + NT - 1 + 7) & ~7;
    p.S = wgrad_tc_splits(B, L, Cin, Cout, Kmax);
    p.stage_bytes = 16 * WG_LT * 16 + p.kcx * p.RX * 16;
    const int smem = WG_HDR + WG_STAGES * p.stage_bytes;
    TSC_REQUIRE(smem <= 227 * 1024, "wgrad shape needs %d B of shared memory: unsupported", smem);
    CUtensorMap dymap, xmap;
    if (make_c8_map(&dymap, dy, B, np / 8, L, WG_LT, 1) != 0) return -1;
    if (make_c8_map(&xmap, x, B, p.kcx, L, p.RX, 1) != 0) return -1;
    static bool attr_set = false;          // once per process: opt in to the full 227 KB of dynamic shared memory
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(oswgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = true;
    }
    oswgrad_tc_kernel<<<dim3(items.n, p.S), WG_THREADS, smem, cs>>>(dymap, xmap, items, p);
    TSC_LAUNCH_CHECK();
    return launch_wgrad_reduce(p.part, dW, p.S, Cin, Cout, Kmax, np, cinp, s_of_tap, cs);
}

int read_clear_watchdog_wgrad(int* code) {
    int zero = 0;
    cudaError_t e = cudaMemcpyFromSymbol(code, tc::g_watchdog, sizeof(int));
    if (e != cudaSuccess) return (int)e;
    return (int)cudaMemcpyToSymbol(tc::g_watchdog, &zero, sizeof(int));
}

}  // namespace tsc
