// C-DAN head of the transferred features (reference C_DAN.py:49-82, configuration 3 of BASELINE.json).
//
// The reference evaluates, for the target half and the generated (source-to-target) half of a step,
//     p = softmax(logits)                                   C_DAN.py:53-54
//     fusion = (flat(feature) @ R0) / sqrt(1024) * (p @ R1) C_DAN.py:20-25  (RandomLayer; R0 is the big GEMM: cuBLAS)
//     H = -sum p log(p + 1e-5); w = 1 + exp(-H)             C_DAN.py:32-37,66-71 (gradient reversal on H)
//     w /= sum(w).detach(); distance = sum(w[None,:] * critic(fusion))   C_DAN.py:72-79 ([B] x [B,1] broadcast)
//     loss = distance_target - distance_generated            C_DAN.py:81
// as ~45 element-wise / reduction launches forward and as many again in autograd.  Here both halves go through four
// launches: cdan_fuse fwd/bwd (everything between the R0 GEMM / the logits and the critic input, including the
// gradient reversals) and cdan_distance fwd/bwd (everything behind the critic).  HBM-bound, tiny ([2B, 1024] fp32).
#include "common.cuh"
#include <math.h>

namespace tsc {

constexpr int CDAN_THREADS = 256;
constexpr int CDAN_MAX_CLASSES = 32;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// One CTA per row m of the stacked (target rows, then generated rows) batch.
__global__ void __launch_bounds__(CDAN_THREADS)
cdan_fuse_fwd_kernel(const float* __restrict__ y0, const float* __restrict__ logits, const float* __restrict__ r1,
                     float* __restrict__ fusion, float* __restrict__ prob, float* __restrict__ u, int K, int D,
                     float scale_div, float eps) {
    __shared__ float p_s[CDAN_MAX_CLASSES];
    const int m = blockIdx.x, tid = threadIdx.x;
    if (tid < 32) {
        const bool live = tid < K;
        const float z = live ? logits[(size_t)m * K + tid] : -INFINITY;
        const float mx = warp_max(z);
        const float e = live ? expf(z - mx) : 0.f;
        const float p = e / warp_sum(e);
        const float H = warp_sum(live ? -p * logf(p + eps) : 0.f);
        if (live) {
            p_s[tid] = p;
            prob[(size_t)m * K + tid] = p;
        }
        if (tid == 0) u[m] = 1.f + expf(-H);
    }
    __syncthreads();
    const float* y = y0 + (size_t)m * D;
    float* f = fusion + (size_t)m * D;
    for (int j = tid; j < D; j += CDAN_THREADS) {
        float q = 0.f;
        for (int k = 0; k < K; ++k) q = fmaf(p_s[k], r1[(size_t)k * D + j], q);
        f[j] = (y[j] / scale_div) * q;
    }
}

// dfusion is the gradient at the critic's input BEFORE its gradient reversal (widgets.py:121-122): the kernel applies
// -coeff[0] to the target rows (m < B) and -coeff[1] to the generated rows, and coeff[2] to the reversal on the entropy.
__global__ void __launch_bounds__(CDAN_THREADS)
cdan_fuse_bwd_kernel(const float* __restrict__ dfusion, const float* __restrict__ y0, const float* __restrict__ prob,
                     const float* __restrict__ r1, const float* __restrict__ u, const float* __restrict__ du,
                     const float* __restrict__ coeff, float* __restrict__ dy0, float* __restrict__ dlogits, int B, int K,
                     int D, float scale_div, float eps) {
    extern __shared__ float dq_s[];                       // [D] gradient wrt q = p @ R1
    __shared__ float p_s[CDAN_MAX_CLASSES];
    __shared__ float red[CDAN_THREADS / 32][CDAN_MAX_CLASSES];
    const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < K) p_s[tid] = prob[(size_t)m * K + tid];
    __syncthreads();
    const float rev = -coeff[m < B ? 0 : 1];
    const float* y = y0 + (size_t)m * D;
    const float* df = dfusion + (size_t)m * D;
    float* dy = dy0 + (size_t)m * D;
    for (int j = tid; j < D; j += CDAN_THREADS) {
        float q = 0.f;
        for (int k = 0; k < K; ++k) q = fmaf(p_s[k], r1[(size_t)k * D + j], q);
        const float g = rev * df[j];
        dy[j] = g * q / scale_div;
        dq_s[j] = g * (y[j] / scale_div);
    }
    __syncthreads();
    for (int k = 0; k < K; ++k) {
        float acc = 0.f;
        for (int j = tid; j < D; j += CDAN_THREADS) acc = fmaf(dq_s[j], r1[(size_t)k * D + j], acc);
        acc = warp_sum(acc);
        if (lane == 0) red[warp][k] = acc;
    }
    __syncthreads();
    if (tid < 32) {
        const bool live = tid < K;
        const float p = live ? p_s[tid] : 0.f;
        float dp = 0.f;
        if (live) {
#pragma unroll
            for (int w = 0; w < CDAN_THREADS / 32; ++w) dp += red[w][tid];
            if (du != nullptr) {
                // u = 1 + exp(-h), h = reverse(H): dH = -coeff * du * (-(u - 1))
                const float dH = coeff[2] * du[m] * (u[m] - 1.f);
                dp += dH * (-logf(p + eps) - p / (p + eps));
            }
        }
        const float dot = warp_sum(p * dp);
        if (live) dlogits[(size_t)m * K + tid] = p * (dp - dot);
    }
}

__device__ __forceinline__ float block_sum(float v, float* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();                                      // sh may still be read from the previous call
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < CDAN_THREADS / 32; ++w) t += sh[w];
    return t;
}

// Single CTA.  saved = (St, Ss, Ot, Os, sum w_t, sum w_s).
__global__ void __launch_bounds__(CDAN_THREADS)
cdan_distance_fwd_kernel(const float* __restrict__ u, const float* __restrict__ critic, float* __restrict__ loss,
                         float* __restrict__ saved, int B) {
    __shared__ float sh[CDAN_THREADS / 32];
    float ut = 0.f, us = 0.f, ot = 0.f, os = 0.f;
    for (int i = threadIdx.x; i < B; i += CDAN_THREADS) {
        ut += u[i]; us += u[B + i]; ot += critic[i]; os += critic[B + i];
    }
    const float St = block_sum(ut, sh), Ss = block_sum(us, sh), Ot = block_sum(ot, sh), Os = block_sum(os, sh);
    float wt = 0.f, ws = 0.f;
    for (int i = threadIdx.x; i < B; i += CDAN_THREADS) {
        wt += u[i] / St; ws += u[B + i] / Ss;
    }
    const float Wt = block_sum(wt, sh), Ws = block_sum(ws, sh);
    if (threadIdx.x == 0) {
        loss[0] = Wt * Ot - Ws * Os;
        saved[0] = St; saved[1] = Ss; saved[2] = Ot; saved[3] = Os; saved[4] = Wt; saved[5] = Ws;
    }
}

__global__ void cdan_distance_bwd_kernel(const float* __restrict__ dloss, const float* __restrict__ saved,
                                         float* __restrict__ du, float* __restrict__ dcritic, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * B) return;
    const float g = dloss[0];
    const bool tgt = i < B;
    dcritic[i] = tgt ? g * saved[4] : -g * saved[5];
    du[i] = tgt ? g * saved[2] / saved[0] : -g * saved[3] / saved[1];
}

}  // namespace tsc

extern "C" {

int tsc_cdan_fuse_fwd(const float* y0, const float* logits, const float* r1, float* fusion, float* prob, float* u,
                      int M, int K, int D, float scale_div, tsc_stream_t stream) {
    TSC_REQUIRE(y0 && logits && r1 && fusion && prob && u, "tsc_cdan_fuse_fwd: NULL pointer");
    TSC_REQUIRE(M >= 1 && D >= 1, "tsc_cdan_fuse_fwd: bad sizes M=%d D=%d", M, D);
    TSC_REQUIRE(K >= 1 && K <= tsc::CDAN_MAX_CLASSES, "tsc_cdan_fuse_fwd: %d classes outside [1,%d]", K,
                tsc::CDAN_MAX_CLASSES);
    TSC_REQUIRE(scale_div > 0.f, "tsc_cdan_fuse_fwd: scale_div must be positive");
    tsc::cdan_fuse_fwd_kernel<<<M, tsc::CDAN_THREADS, 0, (cudaStream_t)stream>>>(y0, logits, r1, fusion, prob, u, K, D,
                                                                              scale_div, 1e-5f);
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_cdan_fuse_bwd(const float* dfusion, const float* y0, const float* prob, const float* r1, const float* u,
                      const float* du, const float* coeff, float* dy0, float* dlogits, int B, int K, int D,
                      float scale_div, tsc_stream_t stream) {
    TSC_REQUIRE(dfusion && y0 && prob && r1 && u && coeff && dy0 && dlogits, "tsc_cdan_fuse_bwd: NULL pointer");
    TSC_REQUIRE(B >= 1 && D >= 1 && D <= 8192, "tsc_cdan_fuse_bwd: bad sizes B=%d D=%d (D <= 8192)", B, D);
    TSC_REQUIRE(K >= 1 && K <= tsc::CDAN_MAX_CLASSES, "tsc_cdan_fuse_bwd: %d classes outside [1,%d]", K,
                tsc::CDAN_MAX_CLASSES);
    tsc::cdan_fuse_bwd_kernel<<<2 * B, tsc::CDAN_THREADS, (size_t)D * sizeof(float), (cudaStream_t)stream>>>(
        dfusion, y0, prob, r1, u, du, coeff, dy0, dlogits, B, K, D, scale_div, 1e-5f);
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_cdan_distance_fwd(const float* u, const float* critic_out, float* loss, float* saved, int B,
                          tsc_stream_t stream) {
    TSC_REQUIRE(u && critic_out && loss && saved, "tsc_cdan_distance_fwd: NULL pointer");
    TSC_REQUIRE(B >= 1, "tsc_cdan_distance_fwd: B=%d", B);
    tsc::cdan_distance_fwd_kernel<<<1, tsc::CDAN_THREADS, 0, (cudaStream_t)stream>>>(u, critic_out, loss, saved, B);
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_cdan_distance_bwd(const float* dloss, const float* saved, float* du, float* dcritic_out, int B,
                          tsc_stream_t stream) {
    TSC_REQUIRE(dloss && saved && du && dcritic_out, "tsc_cdan_distance_bwd: NULL pointer");
    TSC_REQUIRE(B >= 1, "tsc_cdan_distance_bwd: B=%d", B);
    tsc::cdan_distance_bwd_kernel<<<tsc::cdiv(2 * B, 256), 256, 0, (cudaStream_t)stream>>>(dloss, saved, du, dcritic_out,
                                                                                         B);
    TSC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
