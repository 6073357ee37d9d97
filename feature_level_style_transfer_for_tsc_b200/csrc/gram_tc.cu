// tcgen05 engine, Gram-matrix style loss as batched tensor-core contractions on fp32 data (kind::tf32).
//
// Forward, one CTA per sample b:   D_b = (a_b a_b^T - s_b s_b^T) / (C L),   partial_b = sum D_b^2
//   * both Grams accumulate into ONE TMEM accumulator: the s-term is issued with the instruction descriptor's
//     negate-A bit, so the difference never exists in two pieces;
//   * operands are K-major straight from the NCL tensor (the contraction runs over positions, which are contiguous);
//     four producer warps stage [rows][positions] chunks through registers and split every value into
//     hi = top 19 bits (exact in TF32) and lo = x - hi, and the MMA warp issues hi*hi + hi*lo + lo*hi ("3xTF32"):
//     after AdaIN the two Grams agree on their diagonal and nearly agree elsewhere, so D is a difference of almost
//     equal numbers and single-pass TF32 (2^-11) would leave only ~1 digit of the loss; the split keeps ~2^-21;
//   * epilogue: tcgen05.ld -> scale -> D (kept for backward) + the fused reduction sum D^2 (warp shuffles, one
//     partial per CTA, summed in order by sum_partials_kernel: deterministic).
// Backward, one CTA per (sample, 128 positions):  da = k D a,  ds = -k D s,  k = 4 g / (B C^3 L)
//   * M = channel i, N = position, K = channel j; A = D_b (K-major as stored), B = a_b / s_b transposed on the fly
//     by the producers into the K-major canonical layout; single-pass TF32 (tolerance 1e-2, no cancellation here).
// Shared-memory operand layout (SWIZZLE_NONE canonical, fp32): [k-group of 4][row][4 floats]: a core matrix is
// 8 rows x 16 B; LBO = (rows + 1) * 16 B (the +1 keeps the producers' 16 B stores conflict-free), SBO = 128 B.
#include "tc_common.cuh"

namespace tsc {
namespace tc {

static constexpr int GR_THREADS = 192;           // warp 0: spare, warp 1: MMA issuer + TMEM owner, warps 2-5: producers/epilogue
static constexpr int GR_KC = 16;                 // positions (fwd) / channels j (bwd) per stage
static constexpr int GR_HDR = 256;

__device__ __forceinline__ void tmem_alloc_cols(uint32_t* slot, uint32_t n) { tmem_alloc(slot, n); }

struct GramFwdParams {
    const float* a;
    const float* s;
    float* D;
    float* partial;
    int B, C, L;
    int Rp;           // operand rows in smem (C rounded up to 128)
    int Np;           // MMA N (C rounded up to 16)
    int NS;           // stages
    int buf_bytes;    // one operand buffer of one stage: (GR_KC/4) * (Rp + 1) * 16
    int tmem_cols;
    float inv_cl;
    long long* tl;
};
#define GTL(i) do { if (p.tl && blockIdx.x == 0) p.tl[i] = clock64(); } while (0)

template <int FI>
__global__ void __launch_bounds__(GR_THREADS, 1) gram_fwd_tc_kernel(const GramFwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);            // [8]
    uint64_t* empty = full + 8;                                    // [8]
    uint64_t* acc_full = empty + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    float* red = reinterpret_cast<float*>(tmem_slot + 2);          // [4]
    uint8_t* stages = smem + GR_HDR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x;
    const int C = p.C, L = p.L, Rp = p.Rp;
    const int stage_bytes = 4 * p.buf_bytes;                       // a_hi | a_lo | s_hi | s_lo
    const int nchunk = (L + GR_KC - 1) / GR_KC;
    const int mtiles = Rp / 128;

    if (warp == 0 && lane == 0) {
        GTL(0);
        for (int i = 0; i < p.NS; ++i) { mbar_init(&full[i], 128); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_cols(tmem_slot, (uint32_t)p.tmem_cols);
    // rows [C, Rp] of every k-group of every buffer are never written by the producers: zero them once
    {
        const int pad_rows = Rp + 1 - C;
        const int groups = p.NS * 4 * (GR_KC / 4);                  // (stage, buffer, k-group)
        for (int i = threadIdx.x; i < groups * pad_rows; i += GR_THREADS) {
            const int gidx = i / pad_rows, r = C + i % pad_rows;
            *reinterpret_cast<uint4*>(stages + ((size_t)gidx * (Rp + 1) + r) * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lbo_bytes = (uint32_t)(Rp + 1) * 16u;

    if (warp == 1) {
        // ===== MMA issuer (whole warp in lock-step, one elected lane issues) =====
        bool dead = false;
        const uint32_t id_pos = make_idesc_tf32(128, (uint32_t)p.Np, false, false, false);
        const uint32_t id_neg = make_idesc_tf32(128, (uint32_t)p.Np, false, false, true);
        const uint32_t hi32 = (128u >> 4) | (1u << 14);           // SBO = 128 B, descriptor version 1
        uint32_t s = 0, ph = 0;
        for (int ck = 0; ck < nchunk; ++ck) {
            mbar_wait(&full[s], ph, dead, 10);
            __syncwarp();
            tc_fence_after();
            if (lane == 0 && p.tl && blockIdx.x == 0 && ck < 16) p.tl[8 + ck] = clock64();
            const uint32_t base16 = (smem_u32(stages) + s * (uint32_t)stage_bytes) >> 4;
            const uint32_t buf16 = (uint32_t)p.buf_bytes >> 4;
            const uint32_t lbo_f = (lbo_bytes >> 4) << 16;
            if (elect_one()) {
                for (int x = 0; x < 2; ++x) {                      // 0: a (+), 1: s (-)
                    const uint32_t hi_b = base16 + (uint32_t)(2 * x) * buf16, lo_b = hi_b + buf16;
                    const uint32_t idesc = x ? id_neg : id_pos;
                    for (int m = 0; m < mtiles; ++m) {
                        const uint32_t d_t = tmem_base + (uint32_t)(m * p.Np);
                        const uint32_t row16 = (uint32_t)(m * 128);             // 128 rows * 16 B, in 16 B units
#pragma unroll
                        for (int k8 = 0; k8 < GR_KC / 8; ++k8) {
                            const uint32_t koff = (uint32_t)(2 * k8) * (lbo_bytes >> 4);
                            const uint64_t a_hi = ((uint64_t)hi32 << 32) | ((hi_b + koff + row16) | lbo_f);
                            const uint64_t a_lo = ((uint64_t)hi32 << 32) | ((lo_b + koff + row16) | lbo_f);
                            const uint64_t b_hi = ((uint64_t)hi32 << 32) | ((hi_b + koff) | lbo_f);
                            const uint64_t b_lo = ((uint64_t)hi32 << 32) | ((lo_b + koff) | lbo_f);
                            const uint32_t first = (ck == 0 && x == 0 && k8 == 0) ? 0u : 1u;
                            umma_tf32(d_t, a_hi, b_hi, idesc, first);
                            umma_tf32(d_t, a_hi, b_lo, idesc, 1u);
                            umma_tf32(d_t, a_lo, b_hi, idesc, 1u);
                        }
                    }
                }
                tc_commit(&empty[s]);
            }
            __syncwarp();
            if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1u; }
        }
        __syncwarp();
        if (elect_one()) tc_commit(acc_full);
    } else if (warp >= 2) {
        // ===== producers: global fp32 -> registers -> (hi, lo) -> shared, K-major canonical layout =====
        bool dead = false;
        const int ptid = threadIdx.x - 64;
        const float* ab = p.a + (size_t)b * C * L;
        const float* sb = p.s + (size_t)b * C * L;
        const bool vec = (L & 3) == 0;
        // item = (row r, quad q of 4 positions): quads of a row are contiguous in global memory.  All loads of a chunk
        // are issued before their first use, and (FI <= 5, i.e. C <= 160) the loads of chunk ck+1 are in flight while
        // chunk ck is split and stored: one L2 round trip is hidden behind the other chunk's work.
        auto load_chunk = [&](int ck, float4 (&ra)[FI], float4 (&rs)[FI]) {
            const int l0 = ck * GR_KC;
#pragma unroll
            for (int u = 0; u < FI; ++u) {
                const int it = ptid + u * 128;
                const int r = it / (GR_KC / 4), q = it % (GR_KC / 4);
                const int l = l0 + q * 4;
                ra[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                rs[u] = ra[u];
                if (it < C * (GR_KC / 4)) {
                    if (vec && l + 3 < L) {
                        ra[u] = __ldg(reinterpret_cast<const float4*>(ab + (size_t)r * L + l));
                        rs[u] = __ldg(reinterpret_cast<const float4*>(sb + (size_t)r * L + l));
                    } else {
                        float va[4], vs[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            va[j] = l + j < L ? __ldg(ab + (size_t)r * L + l + j) : 0.f;
                            vs[j] = l + j < L ? __ldg(sb + (size_t)r * L + l + j) : 0.f;
                        }
                        ra[u] = make_float4(va[0], va[1], va[2], va[3]);
                        rs[u] = make_float4(vs[0], vs[1], vs[2], vs[3]);
                    }
                }
            }
        };
        auto store_chunk = [&](uint8_t* st, const float4 (&ra)[FI], const float4 (&rs)[FI]) {
#pragma unroll
            for (int u = 0; u < FI; ++u) {
                const int it = ptid + u * 128;
                if (it < C * (GR_KC / 4)) {
                    const int r = it / (GR_KC / 4), q = it % (GR_KC / 4);
                    const float va[4] = {ra[u].x, ra[u].y, ra[u].z, ra[u].w}, vs[4] = {rs[u].x, rs[u].y, rs[u].z, rs[u].w};
                    float ah[4], al[4], sh[4], sl[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        ah[j] = __uint_as_float(__float_as_uint(va[j]) & 0xffffe000u);
                        al[j] = va[j] - ah[j];
                        sh[j] = __uint_as_float(__float_as_uint(vs[j]) & 0xffffe000u);
                        sl[j] = vs[j] - sh[j];
                    }
                    uint8_t* dst = st + ((size_t)q * (Rp + 1) + r) * 16;
                    *reinterpret_cast<float4*>(dst) = make_float4(ah[0], ah[1], ah[2], ah[3]);
                    *reinterpret_cast<float4*>(dst + p.buf_bytes) = make_float4(al[0], al[1], al[2], al[3]);
                    *reinterpret_cast<float4*>(dst + 2 * p.buf_bytes) = make_float4(sh[0], sh[1], sh[2], sh[3]);
                    *reinterpret_cast<float4*>(dst + 3 * p.buf_bytes) = make_float4(sl[0], sl[1], sl[2], sl[3]);
                }
            }
        };
        uint32_t s = 0, ph = 0;
        float4 ra[FI], rs[FI], na[FI], ns_[FI];
        if (FI <= 5) load_chunk(0, ra, rs);
        for (int ck = 0; ck < nchunk; ++ck) {
            if (FI <= 5) {
                if (ck + 1 < nchunk) load_chunk(ck + 1, na, ns_);
            } else {
                load_chunk(ck, ra, rs);
            }
            mbar_wait(&empty[s], ph ^ 1u, dead, 11);
            store_chunk(stages + (size_t)s * stage_bytes, ra, rs);
            fence_proxy_async();          // generic-proxy stores -> visible to the tensor core's async-proxy reads
            mbar_arrive(&full[s]);
            if (threadIdx.x == 64 && p.tl && blockIdx.x == 0 && ck < 16) p.tl[24 + ck] = clock64();
            if (FI <= 5) {
#pragma unroll
                for (int u = 0; u < FI; ++u) { ra[u] = na[u]; rs[u] = ns_[u]; }
            }
            if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1u; }
        }
        // ===== epilogue =====
        if (threadIdx.x == 64) mbar_wait(acc_full, 0, dead, 12);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        tc_fence_after();
        if (threadIdx.x == 64) GTL(5);
        const int q = warp & 3;
        float sq = 0.f;
        for (int m = 0; m < mtiles; ++m) {
            if (m * 128 + q * 32 >= C) continue;              // none of this warp's 32 rows is a real channel (warp-uniform)
            const int i = m * 128 + q * 32 + lane;
            // D is symmetric: thread i stores its element (i, c0+j) at [c0+j][i], so the 32 lanes of a warp write
            // 128 contiguous bytes per instruction instead of 32 rows 576 B apart
            float* dcol = p.D + (size_t)b * C * C + i;
            for (int c0 = 0; c0 < p.Np; c0 += 16) {
                float v[16];
                tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * p.Np + c0), v);
                if (i < C) {
                    float* d0 = dcol + (size_t)c0 * C;
                    if (c0 + 16 <= C) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float d = v[j] * p.inv_cl;
                            d0[j * C] = d;
                            sq = fmaf(d, d, sq);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (c0 + j < C) {
                                const float d = v[j] * p.inv_cl;
                                d0[j * C] = d;
                                sq = fmaf(d, d, sq);
                            }
                        }
                    }
                }
            }
        }
        sq = warp_sum(sq);
        if (lane == 0) red[q] = sq;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) p.partial[b] = (red[0] + red[1]) + (red[2] + red[3]);
        if (threadIdx.x == 64) GTL(6);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

struct GramBwdParams {
    const float* D;
    const float* a;
    const float* s;
    const float* dloss;
    float* da;
    float* ds;
    int B, C, L;
    int Rp;           // rows i of the A operand in smem (C rounded up to 128)
    int NS;
    int dbuf_bytes;   // D chunk: (GR_KC/4) * (Rp + 1) * 16
    int xbuf_bytes;   // x chunk: (GR_KC/4) * (128 + 1) * 16
    int tmem_cols;
};

__global__ void __launch_bounds__(GR_THREADS, 1) gram_bwd_tc_kernel(const GramBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 8;
    uint64_t* acc_full = empty + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    uint8_t* stages = smem + GR_HDR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x, l0 = blockIdx.y * 128;
    const int C = p.C, L = p.L, Rp = p.Rp;
    const int stage_bytes = p.dbuf_bytes + 2 * p.xbuf_bytes;      // D | a^T | s^T
    const int nchunk = (C + GR_KC - 1) / GR_KC;
    const int mtiles = Rp / 128;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.NS; ++i) { mbar_init(&full[i], 128); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_cols(tmem_slot, (uint32_t)p.tmem_cols);
    for (int i = threadIdx.x; i < p.NS * stage_bytes / 16; i += GR_THREADS)
        reinterpret_cast<uint4*>(stages)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lbo_d = (uint32_t)(Rp + 1) * 16u, lbo_x = 129u * 16u;

    if (warp == 1) {
        bool dead = false;
        const uint32_t id_pos = make_idesc_tf32(128, 128, false, false, false);
        const uint32_t id_neg = make_idesc_tf32(128, 128, false, false, true);
        const uint32_t hi32 = (128u >> 4) | (1u << 14);
        uint32_t s = 0, ph = 0;
        for (int ck = 0; ck < nchunk; ++ck) {
            mbar_wait(&full[s], ph, dead, 13);
            __syncwarp();
            tc_fence_after();
            const uint32_t d16 = (smem_u32(stages) + s * (uint32_t)stage_bytes) >> 4;
            const uint32_t a16 = d16 + ((uint32_t)p.dbuf_bytes >> 4), s16 = a16 + ((uint32_t)p.xbuf_bytes >> 4);
            if (elect_one()) {
                for (int m = 0; m < mtiles; ++m) {
#pragma unroll
                    for (int k8 = 0; k8 < GR_KC / 8; ++k8) {
                        const uint64_t ad = ((uint64_t)hi32 << 32) |
                                            ((d16 + (uint32_t)(2 * k8) * (lbo_d >> 4) + (uint32_t)(m * 128)) | ((lbo_d >> 4) << 16));
                        const uint64_t ba = ((uint64_t)hi32 << 32) | ((a16 + (uint32_t)(2 * k8) * (lbo_x >> 4)) | ((lbo_x >> 4) << 16));
                        const uint64_t bs = ((uint64_t)hi32 << 32) | ((s16 + (uint32_t)(2 * k8) * (lbo_x >> 4)) | ((lbo_x >> 4) << 16));
                        const uint32_t acc = (ck == 0 && k8 == 0) ? 0u : 1u;
                        umma_tf32(tmem_base + (uint32_t)(m * 256), ad, ba, id_pos, acc);          // da tile
                        umma_tf32(tmem_base + (uint32_t)(m * 256 + 128), ad, bs, id_neg, acc);    // ds tile (negated)
                    }
                }
                tc_commit(&empty[s]);
            }
            __syncwarp();
            if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1u; }
        }
        __syncwarp();
        if (elect_one()) tc_commit(acc_full);
    } else if (warp >= 2) {
        bool dead = false;
        const int ptid = threadIdx.x - 64;
        const float* Db = p.D + (size_t)b * C * C;
        const float* ab = p.a + (size_t)b * C * L;
        const float* sb = p.s + (size_t)b * C * L;
        const bool vecD = (C & 3) == 0;
        uint32_t s = 0, ph = 0;
        for (int ck = 0; ck < nchunk; ++ck) {
            mbar_wait(&empty[s], ph ^ 1u, dead, 14);
            uint8_t* st = stages + (size_t)s * stage_bytes;
            const int j0 = ck * GR_KC;
            // All loads of the chunk are issued before the first use.
            // A operand: D[i][j0 .. j0+KC): item = (row i, quad q); zero beyond C
            constexpr int DI = (256 * (GR_KC / 4) + 127) / 128;
            float4 rd[DI];
#pragma unroll
            for (int u = 0; u < DI; ++u) {
                const int it = ptid + u * 128;
                const int i = it / (GR_KC / 4), q = it % (GR_KC / 4);
                const int j = j0 + q * 4;
                rd[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (it < C * (GR_KC / 4)) {
                    if (vecD && j + 3 < C) {
                        rd[u] = __ldg(reinterpret_cast<const float4*>(Db + (size_t)i * C + j));
                    } else {
                        rd[u].x = j < C ? __ldg(Db + (size_t)i * C + j) : 0.f;
                        rd[u].y = j + 1 < C ? __ldg(Db + (size_t)i * C + j + 1) : 0.f;
                        rd[u].z = j + 2 < C ? __ldg(Db + (size_t)i * C + j + 2) : 0.f;
                        rd[u].w = j + 3 < C ? __ldg(Db + (size_t)i * C + j + 3) : 0.f;
                    }
                }
            }
            // B operands: x^T[l][j0 .. j0+KC): item = (position l, quad q): four rows j of x at one position
            constexpr int XI = GR_KC / 4;          // 128 positions x (KC/4) quads over 128 threads
            float xa[XI][4], xs_[XI][4];
#pragma unroll
            for (int u = 0; u < XI; ++u) {
                const int l = ptid, q = u;
                const int j = j0 + q * 4, gl = l0 + l;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const bool ok = gl < L && j + v < C;
                    xa[u][v] = ok ? __ldg(ab + (size_t)(j + v) * L + gl) : 0.f;
                    xs_[u][v] = ok ? __ldg(sb + (size_t)(j + v) * L + gl) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < DI; ++u) {
                const int it = ptid + u * 128;
                if (it < C * (GR_KC / 4)) {
                    const int i = it / (GR_KC / 4), q = it % (GR_KC / 4);
                    *reinterpret_cast<float4*>(st + ((size_t)q * (Rp + 1) + i) * 16) = rd[u];
                }
            }
#pragma unroll
            for (int u = 0; u < XI; ++u) {
                uint8_t* dst = st + p.dbuf_bytes + ((size_t)u * 129 + ptid) * 16;
                *reinterpret_cast<float4*>(dst) = make_float4(xa[u][0], xa[u][1], xa[u][2], xa[u][3]);
                *reinterpret_cast<float4*>(dst + p.xbuf_bytes) = make_float4(xs_[u][0], xs_[u][1], xs_[u][2], xs_[u][3]);
            }
            fence_proxy_async();
            mbar_arrive(&full[s]);
            if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1u; }
        }
        if (threadIdx.x == 64) mbar_wait(acc_full, 0, dead, 15);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        tc_fence_after();
        const int q = warp & 3;
        const float k = 4.f * __ldg(p.dloss) / ((float)p.B * (float)C * (float)C * (float)C * (float)L);
        const bool vecL = (L & 3) == 0;
        for (int m = 0; m < mtiles; ++m) {
            const int i = m * 128 + q * 32 + lane;
            for (int w = 0; w < 2; ++w) {
                float* orow = (w ? p.ds : p.da) + ((size_t)b * C + (i < C ? i : 0)) * L + l0;
                for (int c0 = 0; c0 < 128; c0 += 16) {
                    float v[16];
                    tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * 256 + w * 128 + c0), v);
                    if (i < C) {
                        if (vecL && l0 + c0 + 15 < L) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(v[j] * k, v[j + 1] * k, v[j + 2] * k, v[j + 3] * k);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (l0 + c0 + j < L) orow[c0 + j] = v[j] * k;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace tc

int gram_sum_partials(const float* partial, int n, float scale, float* out, cudaStream_t cs);     // style.cu
static long long* g_gram_timeline = nullptr;
void set_gram_timeline(long long* dev) { g_gram_timeline = dev; }

int gram_fwd_tc(const float* a, const float* s, float* D, float* loss, float* ws, int B, int C, int L, cudaStream_t cs) {
    using namespace tc;
    TSC_REQUIRE(C <= 256, "tcgen05 gram supports up to 256 channels, got %d", C);
    GramFwdParams p;
    p.a = a; p.s = s; p.D = D; p.partial = ws;
    p.B = B; p.C = C; p.L = L;
    p.Rp = (C + 127) / 128 * 128;
    p.Np = pad16(C);
    p.buf_bytes = (GR_KC / 4) * (p.Rp + 1) * 16;
    const int stage_bytes = 4 * p.buf_bytes;
    int ns = (227 * 1024 - GR_HDR) / stage_bytes;
    if (ns > 6) ns = 6;
    TSC_REQUIRE(ns >= 1, "gram forward: stage of %d B does not fit", stage_bytes);
    p.NS = ns;
    const int cols = (p.Rp / 128) * p.Np;
    p.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    p.inv_cl = 1.f / ((float)C * (float)L);
    p.tl = g_gram_timeline;
    const int smem = GR_HDR + ns * stage_bytes;
    static OnceAttr attr_once;
    {
        const cudaError_t e = run_once(attr_once, [] {
            cudaError_t r = cudaFuncSetAttribute(gram_fwd_tc_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(gram_fwd_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            return r;
        });
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    if (C * (GR_KC / 4) <= 5 * 128)
        gram_fwd_tc_kernel<5><<<B, GR_THREADS, smem, cs>>>(p);
    else
        gram_fwd_tc_kernel<8><<<B, GR_THREADS, smem, cs>>>(p);
    TSC_LAUNCH_CHECK();
    return gram_sum_partials(ws, B, 1.f / ((float)B * (float)C * (float)C), loss, cs);
}

int gram_bwd_tc(const float* D, const float* a, const float* s, const float* dloss, float* da, float* ds, int B, int C,
                int L, cudaStream_t cs) {
    using namespace tc;
    TSC_REQUIRE(C <= 256, "tcgen05 gram supports up to 256 channels, got %d", C);
    GramBwdParams p;
    p.D = D; p.a = a; p.s = s; p.dloss = dloss; p.da = da; p.ds = ds;
    p.B = B; p.C = C; p.L = L;
    p.Rp = (C + 127) / 128 * 128;
    p.dbuf_bytes = (GR_KC / 4) * (p.Rp + 1) * 16;
    p.xbuf_bytes = (GR_KC / 4) * 129 * 16;
    const int stage_bytes = p.dbuf_bytes + 2 * p.xbuf_bytes;
    int ns = (227 * 1024 - GR_HDR) / stage_bytes;
    if (ns > 6) ns = 6;
    TSC_REQUIRE(ns >= 1, "gram backward: stage of %d B does not fit", stage_bytes);
    p.NS = ns;
    p.tmem_cols = p.Rp / 128 == 1 ? 256 : 512;
    const int smem = GR_HDR + ns * stage_bytes;
    static OnceAttr attr_once;
    {
        const cudaError_t e = run_once(attr_once, [] {
            return cudaFuncSetAttribute(gram_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        });
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    gram_bwd_tc_kernel<<<dim3(B, cdiv(L, 128)), GR_THREADS, smem, cs>>>(p);
    TSC_LAUNCH_CHECK();
    return 0;
}

int read_clear_watchdog_gram(int* code) {
    int zero = 0;
    cudaError_t e = cudaMemcpyFromSymbol(code, tc::g_watchdog, sizeof(int));
    if (e != cudaSuccess) return (int)e;
    return (int)cudaMemcpyToSymbol(tc::g_watchdog, &zero, sizeof(int));
}

}  // namespace tsc
