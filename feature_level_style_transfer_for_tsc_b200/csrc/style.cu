// Feature-level style transfer: single-pass Welford row statistics with the AdaIN re-normalisation,
// its backward, and the fp32 (SIMT) Gram-matrix style loss.  Spec: SURVEY.md 8c, appendix A3/A4.
// Inserted at the site where the reference runs its flow (train_and_test.py:552-561).
//
// Row kernels: a row is one (b, c) series of L fp32 values of an NCL tensor.  A group of G threads
// (G = 32: one warp, or G = 256: one CTA) owns a row, holds it in registers (V float4 per thread),
// reduces with Welford + Chan merges over warp shuffles, and writes the result from registers:
// every tensor is read or written exactly once (AdaIN fwd = 3*4 B per element, bwd = 5*4 B).
#include "common.cuh"

namespace tsc {

struct WState { float n, mean, m2; };

__device__ __forceinline__ void wf_add(WState& s, float x) {
    s.n += 1.f;
    const float d = x - s.mean;
    s.mean += d / s.n;
    s.m2 = fmaf(d, x - s.mean, s.m2);
}

__device__ __forceinline__ WState wf_warp_merge(WState s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float nb = __shfl_xor_sync(0xffffffffu, s.n, o);
        const float mb = __shfl_xor_sync(0xffffffffu, s.mean, o);
        const float qb = __shfl_xor_sync(0xffffffffu, s.m2, o);
        welford_merge(s.n, s.mean, s.m2, nb, mb, qb);
    }
    return s;
}

// Merge across the G threads that own a row.  G == 32: shuffles only.  G == 256: shuffles, then the
// 8 warp results through shared memory (sh must hold 8*3 floats per quantity slot).
template <int G>
__device__ __forceinline__ WState wf_group_merge(WState s, float* sh) {
    s = wf_warp_merge(s);
    if (G == 32) return s;
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { sh[w * 3] = s.n; sh[w * 3 + 1] = s.mean; sh[w * 3 + 2] = s.m2; }
    __syncthreads();
    WState r = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < G / 32; ++i) welford_merge(r.n, r.mean, r.m2, sh[i * 3], sh[i * 3 + 1], sh[i * 3 + 2]);
    return r;
}

template <int G>
__device__ __forceinline__ float sum_group(float v, float* sh) {
    v = warp_sum(v);
    if (G == 32) return v;
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[w] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < G / 32; ++i) r += sh[i];
    return r;
}

// ---- row statistics only (streaming, any L): mean and unbiased variance ---------------------------
__global__ void __launch_bounds__(256) rowstats_kernel(const float* __restrict__ x, float* __restrict__ mean,
                                                       float* __restrict__ var, int R, int L) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= R) return;
    const float* p = x + (long long)row * L;
    WState s = {0.f, 0.f, 0.f};
    if ((L & 3) == 0 && L <= 1024) {
        // the row fits the warp's registers: local two-pass (n, mean, M2), no per-element division
        float4 v[8];
        float sum = 0.f;
        int cnt = 0;
        const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int q = lane + i * 32;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < L / 4) {
                v[i] = __ldg(p4 + q);
                sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
                cnt += 4;
            }
        }
        s.n = (float)cnt;
        s.mean = cnt ? sum / (float)cnt : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (lane + i * 32 < L / 4) {
                const float a = v[i].x - s.mean, b = v[i].y - s.mean, c = v[i].z - s.mean, d = v[i].w - s.mean;
                s.m2 += (a * a + b * b) + (c * c + d * d);
            }
        }
    } else if ((L & 3) == 0) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
        for (int i = lane; i < L / 4; i += 32) {
            const float4 v = __ldg(p4 + i);
            wf_add(s, v.x); wf_add(s, v.y); wf_add(s, v.z); wf_add(s, v.w);
        }
    } else {
        for (int i = lane; i < L; i += 32) wf_add(s, __ldg(p + i));
    }
    s = wf_warp_merge(s);
    if (lane == 0) {
        mean[row] = s.mean;
        var[row] = s.m2 / fmaxf(s.n - 1.f, 1.f);
    }
}

// ---- AdaIN forward: G threads per row, V float4 (VEC=4) or V floats (VEC=1) per thread ------------
template <int G, int V, int VEC>
__global__ void __launch_bounds__(256) adain_fwd_kernel(const float* __restrict__ content, const float* __restrict__ style,
                                                        float* __restrict__ out, float* __restrict__ stats, float eps,
                                                        int R, int L) {
    __shared__ float sh[2][8 * 3];
    const int rows_per_block = 256 / G;
    const int row = blockIdx.x * rows_per_block + threadIdx.x / G;
    const int tg = threadIdx.x % G;
    const bool active = row < R;       // G == 256 -> one row per block, always active
    const long long base = (long long)(active ? row : 0) * L;
    float c[V][VEC], s[V][VEC];
    // The thread's own (n, mean, M2) come from two passes over its REGISTER-resident values (no per-element division);
    // the threads of a row are then combined with Chan's parallel Welford merge.  HBM is still read exactly once.
    float cnt = 0.f, sum_c = 0.f, sum_s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int e = (i * G + tg) * VEC;
#pragma unroll
        for (int k = 0; k < VEC; ++k) { c[i][k] = 0.f; s[i][k] = 0.f; }
        if (active && e < L) {
            if (VEC == 4) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(content + base + e));
                const float4 b = __ldg(reinterpret_cast<const float4*>(style + base + e));
                c[i][0] = a.x; c[i][1 % VEC] = a.y; c[i][2 % VEC] = a.z; c[i][3 % VEC] = a.w;
                s[i][0] = b.x; s[i][1 % VEC] = b.y; s[i][2 % VEC] = b.z; s[i][3 % VEC] = b.w;
            } else {
                c[i][0] = __ldg(content + base + e);
                s[i][0] = __ldg(style + base + e);
            }
            cnt += (float)VEC;
#pragma unroll
            for (int k = 0; k < VEC; ++k) { sum_c += c[i][k]; sum_s += s[i][k]; }
        }
    }
    const float rc = cnt > 0.f ? 1.f / cnt : 0.f;
    WState wc = {cnt, sum_c * rc, 0.f}, ws_ = {cnt, sum_s * rc, 0.f};
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int e = (i * G + tg) * VEC;
        if (active && e < L) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float dc = c[i][k] - wc.mean, ds = s[i][k] - ws_.mean;
                wc.m2 = fmaf(dc, dc, wc.m2);
                ws_.m2 = fmaf(ds, ds, ws_.m2);
            }
        }
    }
    wc = wf_group_merge<G>(wc, sh[0]);
    ws_ = wf_group_merge<G>(ws_, sh[1]);
    const float denom = fmaxf((float)L - 1.f, 1.f);
    const float sig_c = sqrtf(wc.m2 / denom + eps), sig_s = sqrtf(ws_.m2 / denom + eps);
    const float a = sig_s / sig_c;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int e = (i * G + tg) * VEC;
        if (active && e < L) {
            float o[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) o[k] = fmaf(c[i][k] - wc.mean, a, ws_.mean);
            if (VEC == 4)
                *reinterpret_cast<float4*>(out + base + e) = make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]);
            else
                out[base + e] = o[0];
        }
    }
    if (active && tg == 0) {
        float4 st = make_float4(wc.mean, sig_c, ws_.mean, sig_s);
        *reinterpret_cast<float4*>(stats + (long long)row * 4) = st;
    }
}

// ---- AdaIN backward (appendix A3) -------------------------------------------------------------
template <int G, int V, int VEC>
__global__ void __launch_bounds__(256) adain_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ content,
                                                        const float* __restrict__ style, const float* __restrict__ stats,
                                                        float* __restrict__ dcontent, float* __restrict__ dstyle, int R, int L) {
    __shared__ float sh[2][8];
    const int rows_per_block = 256 / G;
    const int row = blockIdx.x * rows_per_block + threadIdx.x / G;
    const int tg = threadIdx.x % G;
    const bool active = row < R;
    const long long base = (long long)(active ? row : 0) * L;
    const float4 st = __ldg(reinterpret_cast<const float4*>(stats + (long long)(active ? row : 0) * 4));
    const float mu_c = st.x, sig_c = st.y, mu_s = st.z, sig_s = st.w;
    const float inv_c = 1.f / sig_c, inv_s = 1.f / sig_s;
    float g[V][VEC], xh[V][VEC], shh[V][VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int e = (i * G + tg) * VEC;
        if (active && e < L) {
            float cc[VEC], ss[VEC];
            if (VEC == 4) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(dy + base + e));
                const float4 b = __ldg(reinterpret_cast<const float4*>(content + base + e));
                const float4 d = __ldg(reinterpret_cast<const float4*>(style + base + e));
                g[i][0] = a.x; g[i][1 % VEC] = a.y; g[i][2 % VEC] = a.z; g[i][3 % VEC] = a.w;
                cc[0] = b.x; cc[1 % VEC] = b.y; cc[2 % VEC] = b.z; cc[3 % VEC] = b.w;
                ss[0] = d.x; ss[1 % VEC] = d.y; ss[2 % VEC] = d.z; ss[3 % VEC] = d.w;
            } else {
                g[i][0] = __ldg(dy + base + e);
                cc[0] = __ldg(content + base + e);
                ss[0] = __ldg(style + base + e);
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                xh[i][k] = (cc[k] - mu_c) * inv_c;
                shh[i][k] = (ss[k] - mu_s) * inv_s;
                s1 += g[i][k];
                s2 = fmaf(g[i][k], xh[i][k], s2);
            }
        }
    }
    s1 = sum_group<G>(s1, sh[0]);
    s2 = sum_group<G>(s2, sh[1]);
    const float a = sig_s / sig_c;
    const float m1 = s1 / (float)L, m2 = s2 / fmaxf((float)L - 1.f, 1.f);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int e = (i * G + tg) * VEC;
        if (active && e < L) {
            float dc[VEC], ds[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                dc[k] = a * (g[i][k] - m1 - xh[i][k] * m2);
                ds[k] = m1 + shh[i][k] * m2;
            }
            if (VEC == 4) {
                *reinterpret_cast<float4*>(dcontent + base + e) = make_float4(dc[0], dc[1 % VEC], dc[2 % VEC], dc[3 % VEC]);
                *reinterpret_cast<float4*>(dstyle + base + e) = make_float4(ds[0], ds[1 % VEC], ds[2 % VEC], ds[3 % VEC]);
            } else {
                dcontent[base + e] = dc[0];
                dstyle[base + e] = ds[0];
            }
        }
    }
}

// ---- fp32 Gram style loss (SIMT engine) -------------------------------------------------------
// D[b][i][j] = (sum_l a[i,l] a[j,l] - s[i,l] s[j,l]) / (C L); one CTA per (b, 32x32 tile), partial sum
// of D^2 per CTA -> workspace; a second 1-block kernel adds the partials in order (deterministic).
__global__ void __launch_bounds__(256) gram_fwd_simt_kernel(const float* __restrict__ a, const float* __restrict__ s,
                                                            float* __restrict__ D, float* __restrict__ partial,
                                                            int C, int L) {
    __shared__ float ai[32][33], aj[32][33], si[32][33], sj[32][33];
    __shared__ float red[8];
    const int b = blockIdx.z, i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // ty in 0..7, 4 rows each
    const float* ab = a + (long long)b * C * L;
    const float* sb = s + (long long)b * C * L;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int l0 = 0; l0 < L; l0 += 32) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int rr = ty * 4 + r;
            const int l = l0 + tx;
            const bool okl = l < L;
            ai[rr][tx] = (okl && i0 + rr < C) ? ab[(long long)(i0 + rr) * L + l] : 0.f;
            si[rr][tx] = (okl && i0 + rr < C) ? sb[(long long)(i0 + rr) * L + l] : 0.f;
            aj[rr][tx] = (okl && j0 + rr < C) ? ab[(long long)(j0 + rr) * L + l] : 0.f;
            sj[rr][tx] = (okl && j0 + rr < C) ? sb[(long long)(j0 + rr) * L + l] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int l = 0; l < 32; ++l) {
            const float vj = aj[tx][l], wj = sj[tx][l];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] += ai[ty * 4 + r][l] * vj - si[ty * 4 + r][l] * wj;
        }
        __syncthreads();
    }
    const float inv = 1.f / ((float)C * (float)L);
    float sq = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty * 4 + r, j = j0 + tx;
        if (i < C && j < C) {
            const float d = acc[r] * inv;
            D[((long long)b * C + i) * C + j] = d;
            sq = fmaf(d, d, sq);
        }
    }
    sq = warp_sum(sq);
    if (tx == 0) red[ty] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ partial, int n, float scale,
                                                           float* __restrict__ out) {
    __shared__ float red[256];
    float t = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) t += partial[i];
    red[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0] * scale;
}

// da[b][i][l] = k * sum_j D[b][i][j] a[b][j][l],  ds = -k * sum_j D s ; k = g * 4 / (B C^3 L).
// blockIdx.z = b*2 + which (0: a, 1: s).  One CTA per 32 (i) x 64 (l) tile.
__global__ void __launch_bounds__(256) gram_bwd_simt_kernel(const float* __restrict__ D, const float* __restrict__ a,
                                                            const float* __restrict__ s, const float* __restrict__ dloss,
                                                            float* __restrict__ da, float* __restrict__ ds,
                                                            int B, int C, int L) {
    __shared__ float dt[32][33];      // D[i0+r][j0+c]
    __shared__ float xt[32][65];      // x[j0+r][l0+c]
    const int which = blockIdx.z & 1, b = blockIdx.z >> 1;
    const float* x = (which ? s : a) + (long long)b * C * L;
    float* o = (which ? ds : da) + (long long)b * C * L;
    const float* Db = D + (long long)b * C * C;
    const int i0 = blockIdx.y * 32, l0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;       // ty in 0..3 -> 8 rows each
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    for (int j0 = 0; j0 < C; j0 += 32) {
        for (int e = threadIdx.x; e < 32 * 32; e += 256) {
            const int r = e >> 5, c = e & 31;
            dt[r][c] = (i0 + r < C && j0 + c < C) ? Db[(long long)(i0 + r) * C + j0 + c] : 0.f;
        }
        for (int e = threadIdx.x; e < 32 * 64; e += 256) {
            const int r = e >> 6, c = e & 63;
            xt[r][c] = (j0 + r < C && l0 + c < L) ? x[(long long)(j0 + r) * L + l0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const float xv = xt[j][tx];
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = fmaf(dt[ty * 8 + r][j], xv, acc[r]);
        }
        __syncthreads();
    }
    const float k = (which ? -4.f : 4.f) * __ldg(dloss) / ((float)B * (float)C * (float)C * (float)C * (float)L);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = i0 + ty * 8 + r, l = l0 + tx;
        if (i < C && l < L) o[(long long)i * L + l] = acc[r] * k;
    }
}

template <int G, int V, int VEC>
static int launch_adain_fwd(const float* c, const float* s, float* o, float* st, float eps, int R, int L, cudaStream_t cs) {
    const int rpb = 256 / G;
    adain_fwd_kernel<G, V, VEC><<<cdiv(R, rpb), 256, 0, cs>>>(c, s, o, st, eps, R, L);
    TSC_LAUNCH_CHECK();
    return 0;
}
template <int G, int V, int VEC>
static int launch_adain_bwd(const float* dy, const float* c, const float* s, const float* st, float* dc, float* ds, int R,
                            int L, cudaStream_t cs) {
    const int rpb = 256 / G;
    adain_bwd_kernel<G, V, VEC><<<cdiv(R, rpb), 256, 0, cs>>>(dy, c, s, st, dc, ds, R, L);
    TSC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tsc

// Dispatch on the row length: elements per thread = V*VEC, threads per row G.
#define TSC_ADAIN_DISPATCH(FN, ...)                                                          \
    do {                                                                                     \
        if ((L & 3) == 0) {                                                                  \
            if (L <= 128) return FN<32, 1, 4>(__VA_ARGS__);                                  \
            if (L <= 256) return FN<32, 2, 4>(__VA_ARGS__);                                  \
            if (L <= 512) return FN<32, 4, 4>(__VA_ARGS__);                                  \
            if (L <= 1024) return FN<32, 8, 4>(__VA_ARGS__);                                 \
            if (L <= 2048) return FN<256, 2, 4>(__VA_ARGS__);                                \
            if (L <= 4096) return FN<256, 4, 4>(__VA_ARGS__);                                \
            if (L <= 8192) return FN<256, 8, 4>(__VA_ARGS__);                                \
        } else {                                                                             \
            if (L <= 128) return FN<32, 4, 1>(__VA_ARGS__);                                  \
            if (L <= 512) return FN<256, 2, 1>(__VA_ARGS__);                                 \
            if (L <= 2048) return FN<256, 8, 1>(__VA_ARGS__);                                \
        }                                                                                    \
        TSC_REQUIRE(false, "row length L=%d not supported by the in-register AdaIN kernel", L); \
    } while (0)

// Same for the backward kernel (dy, content and style are register resident)
// Dispatch on the row length: elements per thread = V*VEC, threads per row G.
#define TSC_ADAIN_BWD_DISPATCH(FN, ...)                                                          \
    do {                                                                                     \
        if ((L & 3) == 0) {                                                                  \
            if (L <= 128) return FN<32, 1, 4>(__VA_ARGS__);                                  \
            if (L <= 256) return FN<32, 2, 4>(__VA_ARGS__);                                  \
            if (L <= 512) return FN<32, 4, 4>(__VA_ARGS__);                                  \
            if (L <= 1024) return FN<256, 1, 4>(__VA_ARGS__);   /* three tensors in registers */ \
            if (L <= 2048) return FN<256, 2, 4>(__VA_ARGS__);                                \
            if (L <= 4096) return FN<256, 4, 4>(__VA_ARGS__);                                \
            if (L <= 8192) return FN<256, 8, 4>(__VA_ARGS__);                                \
        } else {                                                                             \
            if (L <= 128) return FN<32, 4, 1>(__VA_ARGS__);                                  \
            if (L <= 512) return FN<256, 2, 1>(__VA_ARGS__);                                 \
            if (L <= 2048) return FN<256, 8, 1>(__VA_ARGS__);                                \
        }                                                                                    \
        TSC_REQUIRE(false, "row length L=%d not supported by the in-register AdaIN kernel", L); \
    } while (0)

extern "C" {

int tsc_rowstats_welford(const float* x, float* mean, float* var, int R, int L, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(x && mean && var, "NULL tensor");
    TSC_REQUIRE(R > 0 && L > 0, "bad shape [%d,%d]", R, L);
    rowstats_kernel<<<cdiv(R, 8), 256, 0, (cudaStream_t)stream>>>(x, mean, var, R, L);
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_adain_fwd(const float* content, const float* style, float* out, float* stats, float eps, int R, int L,
                  tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(content && style && out && stats, "NULL tensor");
    TSC_REQUIRE(R > 0 && L > 1, "bad shape [%d,%d]", R, L);
    cudaStream_t cs = (cudaStream_t)stream;
    TSC_ADAIN_DISPATCH(launch_adain_fwd, content, style, out, stats, eps, R, L, cs);
}

int tsc_adain_bwd(const float* dy, const float* content, const float* style, const float* stats, float* dcontent,
                  float* dstyle, int R, int L, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(dy && content && style && stats && dcontent && dstyle, "NULL tensor");
    TSC_REQUIRE(R > 0 && L > 1, "bad shape [%d,%d]", R, L);
    cudaStream_t cs = (cudaStream_t)stream;
    TSC_ADAIN_BWD_DISPATCH(launch_adain_bwd, dy, content, style, stats, dcontent, dstyle, R, L, cs);
}

size_t tsc_gram_workspace_bytes(int B, int C, int L) {
    (void)L;
    const int t = tsc::cdiv(C, 32);
    return (size_t)B * t * t * sizeof(float) + 256;
}

}  // extern "C"

namespace tsc {
int gram_fwd_simt(const float* a, const float* s, float* D, float* loss, float* ws, int B, int C, int L, cudaStream_t cs) {
    const int t = cdiv(C, 32);
    gram_fwd_simt_kernel<<<dim3(t, t, B), 256, 0, cs>>>(a, s, D, ws, C, L);
    TSC_LAUNCH_CHECK();
    sum_partials_kernel<<<1, 256, 0, cs>>>(ws, B * t * t, 1.f / ((float)B * (float)C * (float)C), loss);
    TSC_LAUNCH_CHECK();
    return 0;
}
int gram_sum_partials(const float* partial, int n, float scale, float* out, cudaStream_t cs) {
    sum_partials_kernel<<<1, 256, 0, cs>>>(partial, n, scale, out);
    TSC_LAUNCH_CHECK();
    return 0;
}
int gram_bwd_simt(const float* D, const float* a, const float* s, const float* dloss, float* da, float* ds, int B, int C,
                  int L, cudaStream_t cs) {
    gram_bwd_simt_kernel<<<dim3(cdiv(L, 64), cdiv(C, 32), B * 2), 256, 0, cs>>>(D, a, s, dloss, da, ds, B, C, L);
    TSC_LAUNCH_CHECK();
    return 0;
}
}  // namespace tsc

namespace tsc {
// tensor-core Gram (gram_tc.cu)
int gram_fwd_tc(const float* a, const float* s, float* D, float* loss, float* ws, int B, int C, int L, cudaStream_t cs);
int gram_bwd_tc(const float* D, const float* a, const float* s, const float* dloss, float* da, float* ds, int B, int C,
                int L, cudaStream_t cs);
}  // namespace tsc

extern "C" {

int tsc_gram_loss_fwd(int engine, const float* a, const float* s, float* D, float* loss, float* workspace, int B, int C,
                      int L, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(a && s && D && loss && workspace, "NULL tensor");
    TSC_REQUIRE(B > 0 && C > 0 && L > 0, "bad shape [%d,%d,%d]", B, C, L);
    if (engine == TSC_ENGINE_SIMT) return gram_fwd_simt(a, s, D, loss, workspace, B, C, L, (cudaStream_t)stream);
    if (engine == TSC_ENGINE_TCGEN05) return gram_fwd_tc(a, s, D, loss, workspace, B, C, L, (cudaStream_t)stream);
    TSC_REQUIRE(false, "bad engine %d", engine);
}

int tsc_gram_loss_bwd(int engine, const float* D, const float* a, const float* s, const float* dloss, float* da,
                      float* ds, int B, int C, int L, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(D && a && s && dloss && da && ds, "NULL tensor");
    TSC_REQUIRE(B > 0 && C > 0 && L > 0, "bad shape [%d,%d,%d]", B, C, L);
    if (engine == TSC_ENGINE_SIMT) return gram_bwd_simt(D, a, s, dloss, da, ds, B, C, L, (cudaStream_t)stream);
    if (engine == TSC_ENGINE_TCGEN05) return gram_bwd_tc(D, a, s, dloss, da, ds, B, C, L, (cudaStream_t)stream);
    TSC_REQUIRE(false, "bad engine %d", engine);
}

}  // extern "C"
