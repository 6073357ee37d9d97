// BatchNorm1d (+ReLU, + shortcut add) on c8 fp32 tensors: statistics (Welford), apply, backward.
// All HBM-bound: one thread moves one 32 B row of 8 channels; a warp moves 1 KB contiguous.
// Reference semantics: OS_CNN/OS_CNN.py:65,72-74,165,176-180; formulas SURVEY appendix A2.
#include "common.cuh"
#include <stdlib.h>

namespace tsc {

static constexpr int BN_THREADS = 256;

// number of row-splits per channel chunk so that the grid is exactly ONE wave of `per_sm` resident blocks on each of the
// 148 SMs (rounded down: these kernels are latency-bound, and a grid of 600 blocks on 592 slots ran a second wave of
// eight blocks that doubled the kernel's time -- profiles/README.md, session 3)
static int bn_splits(int B, int Cpc, int L, int per_sm = 4) {
    const long long rows = (long long)B * L;
    static const bool legacy = [] { const char* e = getenv("TSC_BN_LEGACY_SPLITS"); return e && e[0] == '1'; }();   // A/B knob
    int s = legacy ? cdiv(148 * 4, Cpc) : (148 * per_sm) / Cpc;
    const int max_s = (int)((rows + BN_THREADS - 1) / BN_THREADS);   // at least one row per thread
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    return s;
}

// ---- pass 1 of the statistics: per (chunk, split) Welford state for 8 channels -------------------
// workspace layout: [Cpc][S][8][3] = (n, mean, M2)
__global__ void __launch_bounds__(BN_THREADS) bn_stats_partial_kernel(const float* __restrict__ y, float* __restrict__ ws,
                                                                      int B, int Cpc, int L, int S) {
    const int ch = blockIdx.x, sp = blockIdx.y;
    const long long rows = (long long)B * L;
    const long long r0 = rows * sp / S, r1 = rows * (sp + 1) / S;
    float n = 0.f, mean[8], m2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { mean[j] = 0.f; m2[j] = 0.f; }
    for (long long r = r0 + threadIdx.x; r < r1; r += BN_THREADS) {
        const int b = (int)(r / L), l = (int)(r % L);
        Row8<float> v;
        v.load(y + (((long long)b * Cpc + ch) * L + l) * 8);
        n += 1.f;
        const float inv = 1.f / n;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = v.v[j] - mean[j];
            mean[j] += d * inv;
            m2[j] += d * (v.v[j] - mean[j]);
        }
    }
    // warp merge (Chan), then across warps through shared memory
    __shared__ float sh[BN_THREADS / 32][8][3];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float nn = n, mm = mean[j], qq = m2[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float nb = __shfl_xor_sync(0xffffffffu, nn, o);
            const float mb = __shfl_xor_sync(0xffffffffu, mm, o);
            const float qb = __shfl_xor_sync(0xffffffffu, qq, o);
            welford_merge(nn, mm, qq, nb, mb, qb);
        }
        if ((threadIdx.x & 31) == 0) {
            sh[threadIdx.x >> 5][j][0] = nn; sh[threadIdx.x >> 5][j][1] = mm; sh[threadIdx.x >> 5][j][2] = qq;
        }
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const int j = threadIdx.x;
        float nn = 0.f, mm = 0.f, qq = 0.f;
        for (int w = 0; w < BN_THREADS / 32; ++w) welford_merge(nn, mm, qq, sh[w][j][0], sh[w][j][1], sh[w][j][2]);
        float* o = ws + (((long long)ch * S + sp) * 8 + j) * 3;
        o[0] = nn; o[1] = mm; o[2] = qq;
    }
}

// ---- pass 2: merge splits per channel, produce coefficients and update the running statistics -----
// One warp per channel: lanes take the splits round-robin and Chan-merge through shuffles.
__global__ void __launch_bounds__(128) bn_stats_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float* __restrict__ mean_o,
                                         float* __restrict__ invstd_o, float* __restrict__ scale_o,
                                         float* __restrict__ shift_o, float* running_mean, float* running_var,
                                         float momentum, float eps, int C, int Cp, int S) {
    const int c = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= Cp) return;
    if (c >= C) {
        if (lane == 0) { mean_o[c] = 0.f; invstd_o[c] = 0.f; scale_o[c] = 0.f; shift_o[c] = 0.f; }
        return;
    }
    const int ch = c >> 3, j = c & 7;
    float n = 0.f, m = 0.f, q = 0.f;
    for (int s = lane; s < S; s += 32) {
        const float* p = ws + (((long long)ch * S + s) * 8 + j) * 3;
        welford_merge(n, m, q, p[0], p[1], p[2]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float nb = __shfl_xor_sync(0xffffffffu, n, o);
        const float mb = __shfl_xor_sync(0xffffffffu, m, o);
        const float qb = __shfl_xor_sync(0xffffffffu, q, o);
        welford_merge(n, m, q, nb, mb, qb);
    }
    if (lane != 0) return;
    const float var_b = q / n;
    const float invstd = 1.f / sqrtf(var_b + eps);
    mean_o[c] = m;
    invstd_o[c] = invstd;
    const float sc = gamma[c] * invstd;
    scale_o[c] = sc;
    shift_o[c] = beta[c] - m * sc;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (q / fmaxf(n - 1.f, 1.f));
}

__global__ void bn_eval_coeffs_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                      float* mean_o, float* invstd_o, float* scale_o, float* shift_o, int C, int Cp) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cp) return;
    if (c >= C) { mean_o[c] = 0.f; invstd_o[c] = 0.f; scale_o[c] = 0.f; shift_o[c] = 0.f; return; }
    const float invstd = 1.f / sqrtf(rv[c] + eps);
    mean_o[c] = rm[c];
    invstd_o[c] = invstd;
    const float sc = gamma[c] * invstd;
    scale_o[c] = sc;
    shift_o[c] = beta[c] - rm[c] * sc;
}

// ---- apply: out = act(scale*y + shift [+ scale2*y2 + shift2]) ----------------------------------
template <int OUT_KIND>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, const float* __restrict__ y2,
                                                        const float* __restrict__ scale2, const float* __restrict__ shift2,
                                                        int relu, void* __restrict__ out, int B, int C, int Cpc, int L) {
    const long long total = (long long)B * Cpc * L;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i % L);
        const long long bc = i / L;
        const int ch = (int)(bc % Cpc);
        const int b = (int)(bc / Cpc);
        Row8<float> v, o;
        v.load(y + i * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = fmaf(v.v[j], __ldg(scale + ch * 8 + j), __ldg(shift + ch * 8 + j));
        if (y2) {
            Row8<float> w;
            w.load(y2 + i * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] += fmaf(w.v[j], __ldg(scale2 + ch * 8 + j), __ldg(shift2 + ch * 8 + j));
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = fmaxf(o.v[j], 0.f);
        }
        if (OUT_KIND == TSC_OUT_C8_F32) {
            o.store(reinterpret_cast<float*>(out) + i * 8);
        } else if (OUT_KIND == TSC_OUT_C8_BF16) {
            Row8<__nv_bfloat16> ob;
#pragma unroll
            for (int j = 0; j < 8; ++j) ob.v[j] = o.v[j];
            ob.store(reinterpret_cast<__nv_bfloat16*>(out) + i * 8);
        } else {
            float* dst = reinterpret_cast<float*>(out);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = ch * 8 + j;
                if (c < C) dst[((long long)b * C + c) * L + l] = o.v[j];
            }
        }
    }
}

// activation derivative mask recomputed from the pre-activation operands
__device__ __forceinline__ void act_mask(float* d, const float* __restrict__ ym, const float* __restrict__ scale,
                                         const float* __restrict__ shift, const float* __restrict__ ym2,
                                         const float* __restrict__ scale2, const float* __restrict__ shift2,
                                         long long i, int ch) {
    if (!ym) return;
    Row8<float> a;
    a.load(ym + i * 8);
    float z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = fmaf(a.v[j], __ldg(scale + ch * 8 + j), __ldg(shift + ch * 8 + j));
    if (ym2) {
        Row8<float> w;
        w.load(ym2 + i * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] += fmaf(w.v[j], __ldg(scale2 + ch * 8 + j), __ldg(shift2 + ch * 8 + j));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = z[j] > 0.f ? d[j] : 0.f;
}

// ---- backward pass 1: partial S1 = sum d, S2 = sum d*yhat per (chunk, split) --------------------
// workspace layout [Cpc][S][8][2]
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(
    const float* __restrict__ dz, const float* __restrict__ y, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ ym, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ ym2, const float* __restrict__ scale2,
    const float* __restrict__ shift2, float* __restrict__ ws, int B, int Cpc, int L, int S) {
    const int ch = blockIdx.x, sp = blockIdx.y;
    const long long rows = (long long)B * L;
    const long long r0 = rows * sp / S, r1 = rows * (sp + 1) / S;
    float s1[8], s2[8], mu[8], is[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; mu[j] = mean[ch * 8 + j]; is[j] = invstd[ch * 8 + j]; }
    for (long long r = r0 + threadIdx.x; r < r1; r += BN_THREADS) {
        const int b = (int)(r / L), l = (int)(r % L);
        const long long i = ((long long)b * Cpc + ch) * L + l;
        Row8<float> d, v;
        d.load(dz + i * 8);
        v.load(y + i * 8);
        act_mask(d.v, ym, scale, shift, ym2, scale2, shift2, i, ch);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s1[j] += d.v[j];
            s2[j] = fmaf(d.v[j], (v.v[j] - mu[j]) * is[j], s2[j]);
        }
    }
    __shared__ float sh[BN_THREADS / 32][8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float a = warp_sum(s1[j]), c = warp_sum(s2[j]);
        if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5][j][0] = a; sh[threadIdx.x >> 5][j][1] = c; }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        const int j = threadIdx.x >> 1, k = threadIdx.x & 1;
        float a = 0.f;
        for (int w = 0; w < BN_THREADS / 32; ++w) a += sh[w][j][k];
        ws[(((long long)ch * S + sp) * 8 + j) * 2 + k] = a;
    }
}

__global__ void __launch_bounds__(128) bn_bwd_finalize_kernel(const float* __restrict__ ws, float* __restrict__ s1,
                                                              float* __restrict__ s2, int Cp, int S) {
    const int c = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= Cp) return;
    const int ch = c >> 3, j = c & 7;
    float a = 0.f, b = 0.f;
    for (int s = lane; s < S; s += 32) {
        const float* p = ws + (((long long)ch * S + s) * 8 + j) * 2;
        a += p[0]; b += p[1];
    }
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { s1[c] = a; s2[c] = b; }
}

// ---- backward pass 2: dy = gamma*invstd*(d - S1/N - yhat*S2/N)  (train)  |  gamma*invstd*d (eval) ---
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(
    const float* __restrict__ dz, const float* __restrict__ y, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ s1,
    const float* __restrict__ s2, int training, const float* __restrict__ ym, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ ym2, const float* __restrict__ scale2,
    const float* __restrict__ shift2, T* __restrict__ dy, int B, int C, int Cpc, int L) {
    const long long total = (long long)B * Cpc * L;
    const float inv_n = 1.f / ((float)B * (float)L);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long bc = i / L;
        const int ch = (int)(bc % Cpc);
        Row8<float> d, v;
        Row8<T> o;
        d.load(dz + i * 8);
        v.load(y + i * 8);
        act_mask(d.v, ym, scale, shift, ym2, scale2, shift2, i, ch);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch * 8 + j;
            float r = 0.f;
            if (c < C) {
                const float is = __ldg(invstd + c);
                const float g = __ldg(gamma + c) * is;
                if (training) {
                    const float yhat = (v.v[j] - __ldg(mean + c)) * is;
                    r = g * (d.v[j] - __ldg(s1 + c) * inv_n - yhat * __ldg(s2 + c) * inv_n);
                } else {
                    r = g * d.v[j];
                }
            }
            o.v[j] = r;
        }
        o.store(dy + i * 8);
    }
}

static int ew_grid(long long rows) {
    long long g = (rows + 255) / 256;
    const long long cap = 148LL * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace tsc

extern "C" {

size_t tsc_bn_workspace_bytes(int B, int C, int L) {
    const int Cpc = tsc::pad16(C) / 8;
    return (size_t)Cpc * tsc::bn_splits(B, Cpc, L) * 8 * 3 * sizeof(float);
}

int tsc_bn_stats(const float* y, const float* gamma, const float* beta, float* ws, float* mean, float* invstd,
                 float* scale, float* shift, float* running_mean, float* running_var, float momentum, float eps,
                 int B, int C, int L, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(y && gamma && beta && ws && mean && invstd && scale && shift, "NULL tensor");
    TSC_REQUIRE(B > 0 && C > 0 && L > 0, "bad shape [%d,%d,%d]", B, C, L);
    const int Cp = pad16(C), Cpc = Cp / 8, S = bn_splits(B, Cpc, L);
    cudaStream_t cs = (cudaStream_t)stream;
    bn_stats_partial_kernel<<<dim3(Cpc, S), BN_THREADS, 0, cs>>>(y, ws, B, Cpc, L, S);
    TSC_LAUNCH_CHECK();
    bn_stats_finalize_kernel<<<cdiv(Cp, 4), 128, 0, cs>>>(ws, gamma, beta, mean, invstd, scale, shift, running_mean,
                                                           running_var, momentum, eps, C, Cp, S);
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_bn_eval_coeffs(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                       float* mean, float* invstd, float* scale, float* shift, int C, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(gamma && beta && rm && rv && mean && invstd && scale && shift, "NULL tensor");
    const int Cp = pad16(C);
    bn_eval_coeffs_kernel<<<cdiv(Cp, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, rm, rv, eps, mean, invstd, scale,
                                                                          shift, C, Cp);
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_bn_apply(const float* y, const float* scale, const float* shift, const float* y2, const float* scale2,
                 const float* shift2, int relu, void* out, int out_kind, int B, int C, int L, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(y && scale && shift && out, "NULL tensor");
    TSC_REQUIRE(!y2 || (scale2 && shift2), "second branch needs scale2/shift2");
    TSC_REQUIRE(B > 0 && C > 0 && L > 0, "bad shape [%d,%d,%d]", B, C, L);
    const int Cpc = pad16(C) / 8;
    const int g = ew_grid((long long)B * Cpc * L);
    cudaStream_t cs = (cudaStream_t)stream;
    switch (out_kind) {
        case TSC_OUT_C8_F32: bn_apply_kernel<TSC_OUT_C8_F32><<<g, 256, 0, cs>>>(y, scale, shift, y2, scale2, shift2, relu, out, B, C, Cpc, L); break;
        case TSC_OUT_C8_BF16: bn_apply_kernel<TSC_OUT_C8_BF16><<<g, 256, 0, cs>>>(y, scale, shift, y2, scale2, shift2, relu, out, B, C, Cpc, L); break;
        case TSC_OUT_NCL_F32: bn_apply_kernel<TSC_OUT_NCL_F32><<<g, 256, 0, cs>>>(y, scale, shift, y2, scale2, shift2, relu, out, B, C, Cpc, L); break;
        default: TSC_REQUIRE(false, "bad out_kind %d", out_kind);
    }
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_bn_bwd_reduce(const float* dz, const float* y, const float* mean, const float* invstd, const float* ym,
                      const float* scale, const float* shift, const float* ym2, const float* scale2,
                      const float* shift2, float* ws, float* s1, float* s2, int B, int C, int L, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(dz && y && mean && invstd && ws && s1 && s2, "NULL tensor");
    TSC_REQUIRE(!ym || (scale && shift), "mask operand needs scale/shift");
    TSC_REQUIRE(!ym2 || (ym && scale2 && shift2), "second mask operand needs the first and scale2/shift2");
    const int Cp = pad16(C), Cpc = Cp / 8, S = bn_splits(B, Cpc, L);
    cudaStream_t cs = (cudaStream_t)stream;
    bn_bwd_reduce_kernel<<<dim3(Cpc, S), BN_THREADS, 0, cs>>>(dz, y, mean, invstd, ym, scale, shift, ym2, scale2, shift2,
                                                             ws, B, Cpc, L, S);
    TSC_LAUNCH_CHECK();
    bn_bwd_finalize_kernel<<<cdiv(Cp, 4), 128, 0, cs>>>(ws, s1, s2, Cp, S);
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_bn_bwd_apply(const float* dz, const float* y, const float* mean, const float* invstd, const float* gamma,
                     const float* s1, const float* s2, int training, const float* ym, const float* scale,
                     const float* shift, const float* ym2, const float* scale2, const float* shift2, void* dy,
                     int dy_dtype, int B, int C, int L, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(dz && y && mean && invstd && gamma && s1 && s2 && dy, "NULL tensor");
    TSC_REQUIRE(!ym || (scale && shift), "mask operand needs scale/shift");
    TSC_REQUIRE(!ym2 || (ym && scale2 && shift2), "second mask operand needs the first and scale2/shift2");
    const int Cpc = pad16(C) / 8;
    const int g = ew_grid((long long)B * Cpc * L);
    cudaStream_t cs = (cudaStream_t)stream;
    if (dy_dtype == TSC_BF16)
        bn_bwd_apply_kernel<__nv_bfloat16><<<g, 256, 0, cs>>>(dz, y, mean, invstd, gamma, s1, s2, training, ym, scale, shift,
                                                            ym2, scale2, shift2, (__nv_bfloat16*)dy, B, C, Cpc, L);
    else if (dy_dtype == TSC_F32)
        bn_bwd_apply_kernel<float><<<g, 256, 0, cs>>>(dz, y, mean, invstd, gamma, s1, s2, training, ym, scale, shift, ym2,
                                                    scale2, shift2, (float*)dy, B, C, Cpc, L);
    else
        TSC_REQUIRE(false, "bad dtype %d", dy_dtype);
    TSC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

// =====================================================================================================
// Fused BatchNorm path of the tcgen05 engine: the statistics arrive as per-CTA partials from the conv
// epilogue (tsc_osconv stat_partial / red_partial), and the merge runs in the prologue of the apply kernels,
// so a training-mode layer costs two launches forward (conv, apply) and needs no separate statistics pass.
// =====================================================================================================
namespace tsc {

static constexpr int BF_THREADS = 256;

// sum of v over the block, result broadcast to all threads; sh: >= BF_THREADS/32 floats per slot (slots: 8 channels)
__device__ __forceinline__ void block_sum8(float (&v)[8], float* sh /*[8][8]*/) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = warp_sum(v[j]);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sh[(threadIdx.x >> 5) * 8 + j] = v[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < BF_THREADS / 32; ++w) a += sh[w * 8 + j];
        v[j] = a;
    }
}

// Coefficients of the 8 channels of chunk `ch` into coef_s[4][8] = (mean, invstd, scale, shift).
// partial != NULL: merge the per-CTA (mean, M2) pairs (n_i = rows of CTA i) -> batch statistics (training);
// partial == NULL: running statistics (eval mode).  `writer` blocks also store them to coef[4][Cp] and update the
// running statistics (momentum, unbiased variance -- torch semantics).
__device__ __forceinline__ void bn_branch_coeffs(const tsc_bn_branch& br, int ch, int C, int Cp, int n_part, int ltiles, int L,
                                                 bool writer, float* coef_s, float* sh) {
    // Warp j of the block owns channel j of the chunk: its lanes stride over the per-CTA (mean, M2) pairs, merge them
    // with Chan's update and finish with a butterfly -- one round of global loads and one barrier per branch (the block
    // used to make two passes over the partials with three block-wide reductions: half of the kernel's time at cfg2
    // size, profiles/README.md session 3).  Lane 0's merge order is fixed, so the result is deterministic.
    (void)sh;
    const int j = threadIdx.x >> 5, lane = threadIdx.x & 31, c = ch * 8 + j;
    float m = 0.f, v = 1.f;
    if (br.stat_partial) {
        const float2* part = reinterpret_cast<const float2*>(br.stat_partial);
        float n = 0.f, mu = 0.f, q = 0.f;
        for (int i = lane; i < n_part; i += 32) {
            const float ni = (float)min(128, L - (i % ltiles) * 128);
            const float2 pr = part[(size_t)i * Cp + c];
            welford_merge(n, mu, q, ni, pr.x, pr.y);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mu, o);
            const float qb = __shfl_xor_sync(0xffffffffu, q, o);
            welford_merge(n, mu, q, nb, mb, qb);
        }
        m = mu;
        v = q / n;
        if (writer && lane == 0 && c < C && br.running_mean && br.momentum > 0.f) {
            br.running_mean[c] = (1.f - br.momentum) * br.running_mean[c] + br.momentum * mu;
            br.running_var[c] = (1.f - br.momentum) * br.running_var[c] + br.momentum * (q / fmaxf(n - 1.f, 1.f));
        }
    } else if (c < C) {
        m = br.running_mean[c];
        v = br.running_var[c];
    }
    if (lane == 0) {
        float invstd = 0.f, sc = 0.f, shf = 0.f;
        if (c < C) {
            invstd = 1.f / sqrtf(v + br.eps);
            sc = br.gamma[c] * invstd;
            shf = br.beta[c] - m * sc;
        } else {
            m = 0.f;
        }
        coef_s[0 * 8 + j] = m; coef_s[1 * 8 + j] = invstd; coef_s[2 * 8 + j] = sc; coef_s[3 * 8 + j] = shf;
        if (writer && br.coef) {
            br.coef[0 * Cp + c] = m; br.coef[1 * Cp + c] = invstd; br.coef[2 * Cp + c] = sc; br.coef[3 * Cp + c] = shf;
        }
    }
    __syncthreads();
}

// grid (Cpc, S): block = one 8-channel chunk x one share of the (b, l) rows.
template <int OUT_KIND>
__global__ void __launch_bounds__(BF_THREADS) bn_apply_fused_kernel(const tsc_bn_branch a, const tsc_bn_branch b2, int two,
                                                                     int n_part, int relu, void* __restrict__ out, int B,
                                                                     int C, int Cpc, int L, int S) {
    __shared__ float coef_a[32], coef_b[32], sh[64];
    pdl_trigger();
    pdl_wait();
    const int ch = blockIdx.x, sp = blockIdx.y;
    const int ltiles = (L + 127) / 128;
    bn_branch_coeffs(a, ch, C, Cpc * 8, n_part, ltiles, L, sp == 0, coef_a, sh);
    if (two) bn_branch_coeffs(b2, ch, C, Cpc * 8, n_part, ltiles, L, sp == 0, coef_b, sh);
    float sc[8], sf[8], sc2[8], sf2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = coef_a[16 + j]; sf[j] = coef_a[24 + j];
        sc2[j] = two ? coef_b[16 + j] : 0.f; sf2[j] = two ? coef_b[24 + j] : 0.f;
    }
    const long long rows = (long long)B * L;
    const long long r0 = rows * sp / S, r1 = rows * (sp + 1) / S;
    for (long long r = r0 + threadIdx.x; r < r1; r += BF_THREADS) {
        const int bb = (int)(r / L), l = (int)(r % L);
        const long long i = ((long long)bb * Cpc + ch) * L + l;
        Row8<float> v, o;
        v.load(a.y_c8 + i * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = fmaf(v.v[j], sc[j], sf[j]);
        if (two) {
            Row8<float> w;
            w.load(b2.y_c8 + i * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] += fmaf(w.v[j], sc2[j], sf2[j]);
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = fmaxf(o.v[j], 0.f);
        }
        if (OUT_KIND == TSC_OUT_C8_F32) {
            o.store(reinterpret_cast<float*>(out) + i * 8);
        } else if (OUT_KIND == TSC_OUT_C8_BF16) {
            Row8<__nv_bfloat16> ob;
#pragma unroll
            for (int j = 0; j < 8; ++j) ob.v[j] = o.v[j];
            ob.store(reinterpret_cast<__nv_bfloat16*>(out) + i * 8);
        } else {
            float* dst = reinterpret_cast<float*>(out);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = ch * 8 + j;
                if (c < C) dst[((long long)bb * C + c) * L + l] = o.v[j];
            }
        }
    }
}

// Pooled output: out[b][c] = mean_l act(scale*y + shift).  grid (Cpc, S): block = one chunk x a range of whole samples.
__global__ void __launch_bounds__(BF_THREADS) bn_apply_pooled_kernel(const tsc_bn_branch a, int n_part, int relu,
                                                                      float* __restrict__ out, int B, int C, int Cpc, int L,
                                                                      int S) {
    __shared__ float coef_a[32], sh[64];
    pdl_trigger();
    pdl_wait();
    const int ch = blockIdx.x, sp = blockIdx.y;
    const int ltiles = (L + 127) / 128;
    bn_branch_coeffs(a, ch, C, Cpc * 8, n_part, ltiles, L, sp == 0, coef_a, sh);
    float sc[8], sf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = coef_a[16 + j]; sf[j] = coef_a[24 + j]; }
    const int b0 = (int)((long long)B * sp / S), b1 = (int)((long long)B * (sp + 1) / S);
    const float inv_l = 1.f / (float)L;
    for (int bb = b0; bb < b1; ++bb) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int l = threadIdx.x; l < L; l += BF_THREADS) {
            Row8<float> v;
            v.load(a.y_c8 + (((long long)bb * Cpc + ch) * L + l) * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float o = fmaf(v.v[j], sc[j], sf[j]);
                if (relu) o = fmaxf(o, 0.f);
                acc[j] += o;
            }
        }
        block_sum8(acc, sh);
        if (threadIdx.x < 8) {
            const int j = threadIdx.x, c = ch * 8 + j;
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { if (k == j) v = acc[k]; }
            if (c < C) out[(long long)bb * C + c] = v * inv_l;
        }
    }
}

// Top of a stack: the incoming gradient is NCL fp32.  d = dout * [z > 0] (z = scale*y + shift [+ second branch]) is
// written as c8 fp32 and the per-block partial sums (S1, S2) of each branch go to red_partial[S][Cp][2].
// TWO is a template parameter: the two-branch form keeps 8 more coefficient vectors live (141 registers, one block per SM,
// four waves at cfg2 size); each form is bounded to two blocks per SM.
template <bool TWO>
__global__ void __launch_bounds__(BF_THREADS, 2) bn_bwd_top_kernel(const float* __restrict__ dout, const tsc_bn_bwd_branch a,
                                                                    const tsc_bn_bwd_branch b2, int two_unused, int relu,
                                                                 float* __restrict__ d_c8, int B, int C, int Cpc, int L,
                                                                 int S, int pooled) {
    __shared__ float sh[64];
    pdl_trigger();
    pdl_wait();
    const int ch = blockIdx.x, sp = blockIdx.y, Cp = Cpc * 8;
    constexpr bool two = TWO;
    (void)two_unused;
    float mean[8], invstd[8], sc[8], sf[8], mean2[8], invstd2[8], sc2[8], sf2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = ch * 8 + j;
        mean[j] = a.coef[c]; invstd[j] = a.coef[Cp + c]; sc[j] = a.coef[2 * Cp + c]; sf[j] = a.coef[3 * Cp + c];
        mean2[j] = two ? b2.coef[c] : 0.f; invstd2[j] = two ? b2.coef[Cp + c] : 0.f;
        sc2[j] = two ? b2.coef[2 * Cp + c] : 0.f; sf2[j] = two ? b2.coef[3 * Cp + c] : 0.f;
    }
    float s1[8], s2[8], s2b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; s2b[j] = 0.f; }
    const long long rows = (long long)B * L;
    const long long r0 = rows * sp / S, r1 = rows * (sp + 1) / S;
    for (long long r = r0 + threadIdx.x; r < r1; r += BF_THREADS) {
        const int bb = (int)(r / L), l = (int)(r % L);
        const long long i = ((long long)bb * Cpc + ch) * L + l;
        Row8<float> y, y2, d;
        y.load(a.y_c8 + i * 8);
        if (two) y2.load(b2.y_c8 + i * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch * 8 + j;
            float g = 0.f;
            if (c < C) g = pooled ? __ldg(dout + (long long)bb * C + c) / (float)L : __ldg(dout + ((long long)bb * C + c) * L + l);
            if (relu) {
                float z = fmaf(y.v[j], sc[j], sf[j]);
                if (two) z += fmaf(y2.v[j], sc2[j], sf2[j]);
                if (!(z > 0.f)) g = 0.f;
            }
            d.v[j] = g;
            s1[j] += g;
            s2[j] = fmaf(g, (y.v[j] - mean[j]) * invstd[j], s2[j]);
            if (two) s2b[j] = fmaf(g, (y2.v[j] - mean2[j]) * invstd2[j], s2b[j]);
        }
        d.store(d_c8 + i * 8);
    }
    block_sum8(s1, sh);
    block_sum8(s2, sh);
    if (two) block_sum8(s2b, sh);
    if (threadIdx.x < 8) {
        const int j = threadIdx.x, c = ch * 8 + j;
        float v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { if (k == j) { v1 = s1[k]; v2 = s2[k]; v3 = s2b[k]; } }
        reinterpret_cast<float2*>(a.red_partial)[(size_t)sp * Cp + c] = make_float2(v1, v2);
        if (two) reinterpret_cast<float2*>(b2.red_partial)[(size_t)sp * Cp + c] = make_float2(v1, v3);
    }
}

// dy = gamma*invstd*(d - S1/N - yhat*S2/N) (training) | gamma*invstd*d (eval), written as c8 bf16/fp32; (S1, S2) are
// summed from red_partial[n_part][Cp][2] in the prologue (fixed order).  Split 0 also writes the parameter
// gradients dgamma = S2, dbeta = S1, dbias = 0 (training) | gamma*invstd*S1 (eval), adding when accumulate != 0.
template <typename T>
__global__ void __launch_bounds__(BF_THREADS) bn_bwd_apply_fused_kernel(const float* __restrict__ d_c8, const tsc_bn_bwd_branch a,
                                                                         int n_part, int accumulate, T* __restrict__ dy,
                                                                         int B, int C, int Cpc, int L, int S) {
    __shared__ float sh[64];
    pdl_trigger();
    pdl_wait();
    const int ch = blockIdx.x, sp = blockIdx.y, Cp = Cpc * 8;
    float s1[8], s2[8];
    {
        // warp w sums the (S1, S2) partials of channel w of the chunk (fixed order), one barrier
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const float2* part = reinterpret_cast<const float2*>(a.red_partial) + ch * 8 + w;
        float a1 = 0.f, a2 = 0.f;
        for (int i = lane; i < n_part; i += 32) {
            const float2 pr = part[(size_t)i * Cp];
            a1 += pr.x; a2 += pr.y;
        }
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (lane == 0) { sh[w] = a1; sh[8 + w] = a2; }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1[j] = sh[j]; s2[j] = sh[8 + j]; }
    }
    float mean[8], invstd[8], g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = ch * 8 + j;
        mean[j] = a.coef[c]; invstd[j] = a.coef[Cp + c];
        g[j] = c < C ? a.gamma[c] * invstd[j] : 0.f;
    }
    if (sp == 0 && threadIdx.x < 8) {
        const int j = threadIdx.x, c = ch * 8 + j;
        if (c < C) {
            float v1 = 0.f, v2 = 0.f, gg = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { if (k == j) { v1 = s1[k]; v2 = s2[k]; gg = g[k]; } }
            const float db = a.training ? 0.f : gg * v1;
            if (a.dgamma) a.dgamma[c] = accumulate ? a.dgamma[c] + v2 : v2;
            if (a.dbeta) a.dbeta[c] = accumulate ? a.dbeta[c] + v1 : v1;
            if (a.dbias) a.dbias[c] = accumulate ? a.dbias[c] + db : db;
        }
    }
    const float inv_n = 1.f / ((float)B * (float)L);
    const long long rows = (long long)B * L;
    const long long r0 = rows * sp / S, r1 = rows * (sp + 1) / S;
    for (long long r = r0 + threadIdx.x; r < r1; r += BF_THREADS) {
        const int bb = (int)(r / L), l = (int)(r % L);
        const long long i = ((long long)bb * Cpc + ch) * L + l;
        Row8<float> d, y;
        Row8<T> o;
        d.load(d_c8 + i * 8);
        if (a.training) y.load(a.y_c8 + i * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float rr;
            if (a.training) {
                const float yhat = (y.v[j] - mean[j]) * invstd[j];
                rr = g[j] * (d.v[j] - s1[j] * inv_n - yhat * s2[j] * inv_n);
            } else {
                rr = g[j] * d.v[j];
            }
            o.v[j] = rr;
        }
        o.store(dy + i * 8);
    }
}

}  // namespace tsc

extern "C" {

static constexpr int BWD_TOP_PER_SM = 2;        // __launch_bounds__ of bn_bwd_top_kernel
static constexpr int BWD_APPLY_PER_SM = 3;      // bn_bwd_apply_fused_kernel: 80 registers x 256 threads

int tsc_bn_fused_splits(int B, int C, int L) { return tsc::bn_splits(B, tsc::pad16(C) / 8, L, BWD_TOP_PER_SM); }

int tsc_bn_apply_fused(const tsc_bn_branch* a, const tsc_bn_branch* b, int n_part, int relu, void* out, int out_kind, int B,
                       int C, int L, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(a && a->y_c8 && a->gamma && a->beta && out, "NULL argument");
    TSC_REQUIRE(a->stat_partial || (a->running_mean && a->running_var), "eval-mode branch needs running statistics");
    TSC_REQUIRE(!b || (b->y_c8 && b->gamma && b->beta && (b->stat_partial || (b->running_mean && b->running_var))),
                "second branch incomplete");
    TSC_REQUIRE(B > 0 && C > 0 && L > 0, "bad shape [%d,%d,%d]", B, C, L);
    TSC_REQUIRE(n_part == B * cdiv(L, 128), "n_part=%d does not match B*ceil(L/128)=%d", n_part, B * cdiv(L, 128));
    const int Cpc = pad16(C) / 8, S = bn_splits(B, Cpc, L);
    const tsc_bn_branch bb = b ? *b : *a;
    cudaStream_t cs = (cudaStream_t)stream;
    dim3 grid(Cpc, S);
    switch (out_kind) {
        case TSC_OUT_C8_F32: launch_pdl(bn_apply_fused_kernel<TSC_OUT_C8_F32>, grid, dim3(BF_THREADS), 0, cs, *a, bb, (int)(b != nullptr), n_part, relu, out, B, C, Cpc, L, S); break;
        case TSC_OUT_C8_BF16: launch_pdl(bn_apply_fused_kernel<TSC_OUT_C8_BF16>, grid, dim3(BF_THREADS), 0, cs, *a, bb, (int)(b != nullptr), n_part, relu, out, B, C, Cpc, L, S); break;
        case TSC_OUT_NCL_F32: launch_pdl(bn_apply_fused_kernel<TSC_OUT_NCL_F32>, grid, dim3(BF_THREADS), 0, cs, *a, bb, (int)(b != nullptr), n_part, relu, out, B, C, Cpc, L, S); break;
        case TSC_OUT_POOLED: {
            TSC_REQUIRE(b == nullptr, "pooled output supports a single branch");
            const int Sb = S < B ? S : B;
            launch_pdl(bn_apply_pooled_kernel, dim3(Cpc, Sb), dim3(BF_THREADS), 0, cs, *a, n_part, relu, (float*)out, B, C, Cpc, L, Sb);
            break;
        }
        default: TSC_REQUIRE(false, "bad out_kind %d", out_kind);
    }
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_bn_bwd_top(const float* dout_ncl, const tsc_bn_bwd_branch* a, const tsc_bn_bwd_branch* b, int relu, float* d_c8,
                   int B, int C, int L, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(dout_ncl && a && a->y_c8 && a->coef && a->red_partial && d_c8, "NULL argument");
    TSC_REQUIRE(!b || (b->y_c8 && b->coef && b->red_partial), "second branch incomplete");
    TSC_REQUIRE(B > 0 && C > 0 && L > 0, "bad shape [%d,%d,%d]", B, C, L);
    const int Cpc = pad16(C) / 8, S = bn_splits(B, Cpc, L, BWD_TOP_PER_SM);
    const tsc_bn_bwd_branch bb = b ? *b : *a;
    launch_pdl(b ? bn_bwd_top_kernel<true> : bn_bwd_top_kernel<false>, dim3(Cpc, S), dim3(BF_THREADS), 0, (cudaStream_t)stream, dout_ncl, *a, bb, (int)(b != nullptr), relu, d_c8,
               B, C, Cpc, L, S, 0);
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_bn_bwd_top_pooled(const float* dpooled, const tsc_bn_bwd_branch* a, int relu, float* d_c8, int B, int C, int L,
                          tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(dpooled && a && a->y_c8 && a->coef && a->red_partial && d_c8, "NULL argument");
    TSC_REQUIRE(B > 0 && C > 0 && L > 0, "bad shape [%d,%d,%d]", B, C, L);
    const int Cpc = pad16(C) / 8, S = bn_splits(B, Cpc, L, BWD_TOP_PER_SM);
    launch_pdl(bn_bwd_top_kernel<false>, dim3(Cpc, S), dim3(BF_THREADS), 0, (cudaStream_t)stream, dpooled, *a, *a, 0, relu, d_c8, B, C, Cpc, L, S, 1);
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_bn_bwd_apply_fused(const float* d_c8, const tsc_bn_bwd_branch* a, int n_part, int accumulate, void* dy_c8,
                           int dy_dtype, int B, int C, int L, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(d_c8 && a && a->coef && a->red_partial && a->gamma && dy_c8, "NULL argument");
    TSC_REQUIRE(!a->training || a->y_c8, "training-mode backward needs y");
    TSC_REQUIRE(B > 0 && C > 0 && L > 0 && n_part > 0, "bad shape [%d,%d,%d] n_part=%d", B, C, L, n_part);
    const int Cpc = pad16(C) / 8, S = bn_splits(B, Cpc, L, BWD_APPLY_PER_SM);
    cudaStream_t cs = (cudaStream_t)stream;
    dim3 grid(Cpc, S);
    if (dy_dtype == TSC_BF16)
        launch_pdl(bn_bwd_apply_fused_kernel<__nv_bfloat16>, grid, dim3(BF_THREADS), 0, cs, d_c8, *a, n_part, accumulate, (__nv_bfloat16*)dy_c8, B, C, Cpc, L, S);
    else if (dy_dtype == TSC_F32)
        launch_pdl(bn_bwd_apply_fused_kernel<float>, grid, dim3(BF_THREADS), 0, cs, d_c8, *a, n_part, accumulate, (float*)dy_c8, B, C, Cpc, L, S);
    else
        TSC_REQUIRE(false, "bad dtype %d", dy_dtype);
    TSC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
