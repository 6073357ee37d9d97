// BatchNorm1d (+ReLU, + shortcut add) on c8 fp32 tensors: statistics (Welford), apply, backward.
// All HBM-bound: one thread moves one 32 B row of 8 channels; a warp moves 1 KB contiguous.
// Reference semantics: OS_CNN/OS_CNN.py:65,72-74,165,176-180; formulas SURVEY appendix A2.
#include "common.cuh"

namespace tsc {

static constexpr int BN_THREADS = 256;

// number of row-splits per channel chunk so that the grid is a couple of waves of 148 SMs
static int bn_splits(int B, int Cpc, int L) {
    const long long rows = (long long)B * L;
    int s = cdiv(148 * 4, Cpc);
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
