// Callers either side of the hot path: the reductions of the GradNorm joint-stage driver and the
// multi-source entropy vote.  All three are tiny, latency-bound kernels; what they replace is dozens of ATen launches and
// host round trips per call, not HBM traffic.
//
//  * multi_sumsq / multi_norm_finish : L2 norm of every tensor of a list in two launches (fixed summation order, no float
//    atomics).  Replaces the per-parameter torch.norm(...).unsqueeze(0) ... torch.cat(...).sum() of GradNorm
//    (train_and_test.py:683-690: 12 parameter tensors x 5 losses = 60 norm launches + 5 cat + 5 sum per step).
//  * class_precision_kernel : row argmax (first maximum = numpy.argmax) + per-class precision of the predictions
//    (multi_source_voting.py:296-311; also the accuracy of utils.py:27-183, which is sum(correct) / N).
//  * entropy_vote_kernel    : softmax -> entropy -> p * (1 + g exp(-H)) * base^w_m, summed over the models, argmax
//    (multi_source_voting.py:369-407).
#include "common.cuh"
#include <math.h>

namespace tsc {

static constexpr int NORM_SPLITS = 16;       // partial sums per tensor
static constexpr int NORM_THREADS = 256;

// grid (NORM_SPLITS, count): block (s, i) sums the squares of slice s of tensor i (contiguous slices, float4 body when the
// tensor is 16 B aligned) and writes one partial; the order of every addition is fixed by the launch shape.
__global__ void __launch_bounds__(NORM_THREADS) multi_sumsq_kernel(const __grid_constant__ tsc_tensor_list list,
                                                                   float* __restrict__ partial) {
    const int i = blockIdx.y, s = blockIdx.x;
    const float* __restrict__ p = list.p[i];
    const long long n = list.n[i];
    // slices in units of 4 elements so that an aligned tensor keeps aligned slices
    const long long quads = (n + 3) / 4, per = (quads + NORM_SPLITS - 1) / NORM_SPLITS;
    const long long q0 = per * s, q1 = (q0 + per < quads) ? q0 + per : quads;
    float acc = 0.f;
    const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
    for (long long q = q0 + threadIdx.x; q < q1; q += NORM_THREADS) {
        const long long e = q * 4;
        if (aligned && e + 4 <= n) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p + e));
            acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        } else {
            for (long long k = e; k < n && k < e + 4; ++k) { const float v = __ldg(p + k); acc += v * v; }
        }
    }
    __shared__ float wsum[NORM_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < NORM_THREADS / 32; ++w) t += wsum[w];
        partial[i * NORM_SPLITS + s] = t;
    }
}

// one warp: norms[i] = sqrt(sum_s partial[i][s]); norms[count] = sum_i norms[i] (sequential, like torch.cat(...).sum() of
// <= 32 values up to rounding order)
__global__ void multi_norm_finish_kernel(const float* __restrict__ partial, float* __restrict__ norms, int count) {
    const int i = threadIdx.x;
    float v = 0.f;
    if (i < count) {
        float t = 0.f;
#pragma unroll
        for (int s = 0; s < NORM_SPLITS; ++s) t += partial[i * NORM_SPLITS + s];
        v = sqrtf(t);
        norms[i] = v;
    }
    float tot = 0.f;
    for (int j = 0; j < count; ++j) tot += __shfl_sync(0xffffffffu, v, j);
    if (i == 0) norms[count] = tot;
}

// ---- voting ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int row_argmax(const float* __restrict__ r, int K) {
    int best = 0;
    float bv = r[0];
    for (int k = 1; k < K; ++k) {
        const float v = r[k];
        if (v > bv || (v != v && bv == bv)) { bv = v; best = k; }      // first maximum; like numpy.argmax the first NaN wins
    }
    return best;
}

// single CTA: integer counters in shared memory (integer atomics: order-independent results)
__global__ void __launch_bounds__(1024) class_precision_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                               int* __restrict__ pred, int* __restrict__ counts,
                                                               double* __restrict__ precision, int N, int K) {
    __shared__ int c_pred[TSC_MAX_CLASSES], c_ok[TSC_MAX_CLASSES];
    for (int k = threadIdx.x; k < K; k += blockDim.x) { c_pred[k] = 0; c_ok[k] = 0; }
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const int a = row_argmax(logits + (size_t)n * K, K);
        if (pred) pred[n] = a;
        atomicAdd(&c_pred[a], 1);
        if (labels && labels[n] == (long long)a) atomicAdd(&c_ok[a], 1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        counts[k] = c_pred[k];
        counts[K + k] = c_ok[k];
        // multi_source_voting.py:308-311: correct / predicted, 0 for a class that was never predicted
        if (precision) precision[k] = c_pred[k] ? (double)c_ok[k] / (double)c_pred[k] : 0.0;
    }
}

// one thread per test series; K <= TSC_MAX_CLASSES, M models
__global__ void __launch_bounds__(128) entropy_vote_kernel(const float* __restrict__ logits, const double* __restrict__ precision,
                                                           float* __restrict__ score, int* __restrict__ pred, int M, int N, int K,
                                                           float entropy_gain, double weight_base) {
    __shared__ double wpow[TSC_MAX_VOTERS * TSC_MAX_CLASSES];           // base ^ (w_m / mean_m w), NaN -> 0 exponent
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double tot = 0.0;
        for (int m = 0; m < M; ++m) tot += precision[m * K + k];
        const double avg = tot / (double)M;                              // multi_source_voting.py:362
        for (int m = 0; m < M; ++m) {
            double w = precision[m * K + k] / avg;                       // :363-365
            if (w != w) w = 0.0;                                         // np.nan_to_num (0/0 when no model predicts k)
            wpow[m * K + k] = pow(weight_base, w);
        }
    }
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float tot[TSC_MAX_CLASSES];
    for (int k = 0; k < K; ++k) tot[k] = 0.f;
    for (int m = 0; m < M; ++m) {
        const float* r = logits + ((size_t)m * N + n) * K;
        float p[TSC_MAX_CLASSES];
        float se = 0.f;
        for (int k = 0; k < K; ++k) { p[k] = expf(r[k]); se += p[k]; }   // :384 (no max subtraction, as the reference)
        float sp = 0.f;
        for (int k = 0; k < K; ++k) { p[k] = p[k] / se; sp += p[k]; }
        float H = 0.f;                                                    // scipy.stats.entropy: pk / sum(pk), -sum pk log pk
        for (int k = 0; k < K; ++k) {
            const float q = p[k] / sp;
            H += q > 0.f ? -q * logf(q) : 0.f;
        }
        const float gain = 1.f + entropy_gain * expf(-H);                 // :387 float32 scalar
        for (int k = 0; k < K; ++k) {
            const float v = (float)((double)(p[k] * gain) * wpow[m * K + k]);   // float32 row * float64 weights, stored as float32
            tot[k] += v;                                                  // :401 float32 sum in model order
        }
    }
    for (int k = 0; k < K; ++k) score[(size_t)n * K + k] = tot[k];
    pred[n] = row_argmax(tot, K);
}

}  // namespace tsc

extern "C" size_t tsc_multi_l2norm_workspace_bytes(int count) {
    return (size_t)(count > 0 ? count : 0) * tsc::NORM_SPLITS * sizeof(float);
}

extern "C" int tsc_multi_l2norm(const tsc_tensor_list* list, float* norms, float* workspace, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(list && norms && workspace, "NULL argument");
    TSC_REQUIRE(list->count >= 1 && list->count <= TSC_MAX_LIST, "tensor list of %d entries outside [1,%d]", list->count, TSC_MAX_LIST);
    for (int i = 0; i < list->count; ++i) {
        TSC_REQUIRE(list->n[i] >= 0 && (list->p[i] != nullptr || list->n[i] == 0), "tensor %d: NULL with %lld elements", i, list->n[i]);
        TSC_REQUIRE((reinterpret_cast<uintptr_t>(list->p[i]) & 3) == 0, "tensor %d is not 4-byte aligned", i);
    }
    multi_sumsq_kernel<<<dim3(NORM_SPLITS, list->count), NORM_THREADS, 0, (cudaStream_t)stream>>>(*list, workspace);
    TSC_LAUNCH_CHECK();
    multi_norm_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(workspace, norms, list->count);
    TSC_LAUNCH_CHECK();
    return 0;
}

extern "C" int tsc_class_precision(const float* logits, const long long* labels, int* pred, int* counts, double* precision,
                                   int N, int K, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(logits && counts, "NULL argument");
    TSC_REQUIRE(N >= 1, "N=%d: no rows", N);
    TSC_REQUIRE(K >= 1 && K <= TSC_MAX_CLASSES, "K=%d outside [1,%d]", K, TSC_MAX_CLASSES);
    class_precision_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logits, labels, pred, counts, precision, N, K);
    TSC_LAUNCH_CHECK();
    return 0;
}

extern "C" int tsc_entropy_vote(const float* logits, const double* precision, float* score, int* pred, int M, int N, int K,
                                float entropy_gain, float weight_base, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(logits && precision && score && pred, "NULL argument");
    TSC_REQUIRE(M >= 1 && M <= TSC_MAX_VOTERS, "M=%d models outside [1,%d]", M, TSC_MAX_VOTERS);
    TSC_REQUIRE(K >= 1 && K <= TSC_MAX_CLASSES, "K=%d outside [1,%d]", K, TSC_MAX_CLASSES);
    TSC_REQUIRE(N >= 1, "N=%d: no rows", N);
    entropy_vote_kernel<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(logits, precision, score, pred, M, N, K, entropy_gain,
                                                                      (double)weight_base);
    TSC_LAUNCH_CHECK();
    return 0;
}
