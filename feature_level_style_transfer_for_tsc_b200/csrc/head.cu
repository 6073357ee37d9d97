// Classifier head of OS_CNN fused with the training loss: Linear(pooled) -> softmax cross-entropy, forward and backward in one
// launch each.  Replaces `self.hidden(X_f)` (OS_CNN/OS_CNN.py:108-109, cuBLAS addmm) + `nn.CrossEntropyLoss` (train_and_test.py
// :593-603; log_softmax + nll_loss) and their autograd (addmm backward x2, bias sum, log_softmax / nll backward): ~10 library
// launches per head and step.  The sizes are tiny (B <= a few thousand rows, C <= 256 pooled channels, K <= 64 classes): CUDA
// cores, fp32, every reduction in a fixed order (no float atomics).
//
//   logits[b,k] = bias[k] + sum_c pooled[b,c] W[k,c];   p = softmax(logits[b,:]);   loss = -(1/B) sum_b log p[b, y_b]
//   g[b,k]      = dlogits[b,k] + dloss * (p[b,k] - [k == y_b]) / B                  (dlogits: gradient arriving at the logits
//   dpooled = g W;   dW = g^T pooled;   dbias = sum_b g                              from other consumers, e.g. C-DAN; nullable)
#include "common.cuh"
#include <algorithm>

namespace tsc {

static constexpr int HEAD_MAX_K = TSC_MAX_CLASSES;
static constexpr int HEAD_THREADS = 256;
static constexpr int HEAD_GB = 128;           // rows of logit gradients staged per pass of the backward kernel

// one warp per row; lanes stride over the channels with all K partial dot products in registers (K <= 64)
template <int KP>
__global__ void __launch_bounds__(HEAD_THREADS) head_ce_fwd_kernel(const float* __restrict__ pooled, const float* __restrict__ W,
                                                                  const float* __restrict__ bias, const long long* __restrict__ labels,
                                                                  float* __restrict__ logits, float* __restrict__ prob,
                                                                  float* __restrict__ row_loss, float* __restrict__ loss,
                                                                  unsigned int* __restrict__ ticket, int B, int C, int K) {
    extern __shared__ float wsm[];                       // [K][C]
    __shared__ float red[HEAD_THREADS / 32];
    __shared__ bool last;
    pdl_wait();
    for (int i = threadIdx.x; i < K * C; i += HEAD_THREADS) wsm[i] = W[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = blockIdx.x * (HEAD_THREADS / 32) + warp; b < B; b += gridDim.x * (HEAD_THREADS / 32)) {
        float acc[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float x = pooled[(size_t)b * C + c];
#pragma unroll
            for (int k = 0; k < KP; ++k)
                if (k < K) acc[k] = fmaf(x, wsm[k * C + c], acc[k]);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            if (k < K) {
                acc[k] = warp_sum(acc[k]) + bias[k];
                mx = fmaxf(mx, acc[k]);
            }
        }
        float se = 0.f;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) se += expf(acc[k] - mx);
        const float lse = mx + logf(se);
        const int y = labels ? (int)labels[b] : -1;
        float ly = 0.f;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            if (k < K) {
                if (lane == (k & 31)) {
                    logits[(size_t)b * K + k] = acc[k];
                    prob[(size_t)b * K + k] = expf(acc[k] - lse);
                }
                if (k == y) ly = acc[k];
            }
        }
        if (lane == 0 && row_loss) row_loss[b] = (y >= 0 && y < K) ? lse - ly : 0.f;
    }
    if (!loss) return;
    // the last block to finish sums the per-row losses in row order (deterministic)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    float s = 0.f;
    for (int b = threadIdx.x; b < B; b += HEAD_THREADS) s += __ldcg(row_loss + b);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < HEAD_THREADS / 32; ++w) t += red[w];
        *loss = t / (float)B;
        *ticket = 0u;                                    // self-cleaning: the next launch finds the counter at zero
    }
}

__device__ __forceinline__ float head_g(const float* dlogits, const float* prob, const long long* labels, float scale, int b, int k,
                                        int K) {
    float g = dlogits ? dlogits[(size_t)b * K + k] : 0.f;
    if (labels) g += scale * (prob[(size_t)b * K + k] - ((int)labels[b] == k ? 1.f : 0.f));
    return g;
}

// blocks [0, row_blocks): dpooled rows (one warp per row); blocks [row_blocks, ...): one thread per (k, c) of dW (sum over b in
// order), the last of them also dbias
__global__ void __launch_bounds__(HEAD_THREADS) head_ce_bwd_kernel(const float* __restrict__ dloss, const float* __restrict__ dlogits,
                                                                  const float* __restrict__ prob, const long long* __restrict__ labels,
                                                                  const float* __restrict__ pooled, const float* __restrict__ W,
                                                                  float* __restrict__ dpooled, float* __restrict__ dW,
                                                                  float* __restrict__ dbias, int accumulate, int row_blocks, int B,
                                                                  int C, int K) {
    pdl_wait();
    const float scale = (labels ? (dloss ? *dloss : 1.f) : 0.f) / (float)B;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if ((int)blockIdx.x < row_blocks) {
        if (!dpooled) return;
        for (int b = blockIdx.x * (HEAD_THREADS / 32) + warp; b < B; b += row_blocks * (HEAD_THREADS / 32)) {
            // lane k holds g[b, k] and g[b, k + 32] (K <= 64); every channel's dot product reads them by shuffle
            const float g0 = lane < K ? head_g(dlogits, prob, labels, scale, b, lane, K) : 0.f;
            const float g1 = lane + 32 < K ? head_g(dlogits, prob, labels, scale, b, lane + 32, K) : 0.f;
            for (int c0 = 0; c0 < C; c0 += 32) {          // (whole warp in every iteration: the shuffles need all lanes)
                const int c = c0 + lane;
                float a = 0.f;
                for (int k = 0; k < K; ++k) {
                    const float g = __shfl_sync(0xffffffffu, k < 32 ? g0 : g1, k & 31);
                    if (c < C) a = fmaf(g, __ldg(W + k * C + c), a);
                }
                if (c < C) dpooled[(size_t)b * C + c] = a;
            }
        }
        return;
    }
    // dW / dbias: the logit gradients of HEAD_GB rows at a time are staged in shared memory (the first version read dlogits,
    // prob and labels from global memory inside the serial sum over b: 128 dependent L2 round trips, 29 us per launch);
    // the sums stay in row order
    __shared__ float g_s[HEAD_GB * HEAD_MAX_K];
    const int o = ((int)blockIdx.x - row_blocks) * HEAD_THREADS + threadIdx.x;
    const bool is_w = o < K * C, is_b = !is_w && o < K * C + K && dbias;
    const int k = is_w ? o / C : o - K * C, c = is_w ? o - k * C : 0;
    float a = 0.f;
    for (int b0 = 0; b0 < B; b0 += HEAD_GB) {
        const int nb = min(HEAD_GB, B - b0);
        __syncthreads();
        for (int i = threadIdx.x; i < nb * K; i += HEAD_THREADS) g_s[i] = head_g(dlogits, prob, labels, scale, b0 + i / K, i % K, K);
        __syncthreads();
        if (is_w) {
            const float* pc = pooled + (size_t)b0 * C + c;
#pragma unroll 8
            for (int b = 0; b < nb; ++b) a = fmaf(g_s[b * K + k], __ldg(pc + (size_t)b * C), a);
        } else if (is_b) {
            for (int b = 0; b < nb; ++b) a += g_s[b * K + k];
        }
    }
    if (is_w) { if (accumulate) dW[o] += a; else dW[o] = a; }
    else if (is_b) { if (accumulate) dbias[k] += a; else dbias[k] = a; }
}

// out = sum_i w[i] * *x[i]  (device scalars): the step's total loss in one launch instead of a chain of ATen adds / muls
struct ScalarTerms { const float* x[8]; float w[8]; int n; };
__global__ void weighted_scalar_sum_kernel(const __grid_constant__ ScalarTerms t, float* __restrict__ out) {
    pdl_wait();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < t.n; ++i) s = fmaf(t.w[i], *t.x[i], s);
        *out = s;
    }
}

}  // namespace tsc

extern "C" {

size_t tsc_head_ce_workspace_bytes(int B) { return ((size_t)B + 8) * sizeof(float); }

int tsc_head_ce_fwd(const float* pooled, const float* W, const float* bias, const long long* labels, float* logits, float* prob,
                    float* loss, void* workspace, int B, int C, int K, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(pooled && W && bias && logits && prob, "NULL tensor");
    TSC_REQUIRE(!loss || (labels && workspace), "the loss needs labels and a workspace of tsc_head_ce_workspace_bytes(B)");
    TSC_REQUIRE(B >= 1 && C >= 1 && K >= 1 && K <= HEAD_MAX_K, "bad head shape B=%d C=%d K=%d (K <= %d)", B, C, K, HEAD_MAX_K);
    const size_t smem = (size_t)K * C * sizeof(float);
    TSC_REQUIRE(smem <= 96 * 1024, "head weight [%d, %d] does not fit shared memory", K, C);
    const int grid = std::min(cdiv(B, HEAD_THREADS / 32), 296);
    // workspace: [0] ticket (zero before the first use, left at zero by every launch), [8 ...] per-row losses
    unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
    float* row_loss = workspace ? reinterpret_cast<float*>(workspace) + 8 : nullptr;
    cudaStream_t cs = (cudaStream_t)stream;
    static OnceAttr once16, once64;
#define TSC_HEAD_LAUNCH(KP, ONCE)                                                                                                  \
    do {                                                                                                                           \
        if (smem > 48 * 1024) {                                                                                                    \
            const cudaError_t e = run_once(ONCE, [] {                                                                              \
                return cudaFuncSetAttribute(head_ce_fwd_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);       \
            });                                                                                                                    \
            if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }                 \
        }                                                                                                                          \
        const cudaError_t le = launch_pdl(head_ce_fwd_kernel<KP>, dim3(grid), dim3(HEAD_THREADS), smem, cs, pooled, W, bias,       \
                                          labels, logits, prob, row_loss, loss, ticket, B, C, K);                                  \
        if (le != cudaSuccess) { set_error("head_ce_fwd launch: %s", cudaGetErrorString(le)); return (int)le; }                    \
    } while (0)
    if (K <= 16) TSC_HEAD_LAUNCH(16, once16); else TSC_HEAD_LAUNCH(64, once64);
#undef TSC_HEAD_LAUNCH
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_head_ce_bwd(const float* dloss, const float* dlogits, const float* prob, const long long* labels, const float* pooled,
                    const float* W, float* dpooled, float* dW, float* dbias, int accumulate, int B, int C, int K,
                    tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(prob && pooled && W && dW, "NULL tensor");
    TSC_REQUIRE(dlogits || labels, "nothing to differentiate: neither a logits gradient nor a loss");
    TSC_REQUIRE(B >= 1 && C >= 1 && K >= 1 && K <= HEAD_MAX_K, "bad head shape B=%d C=%d K=%d (K <= %d)", B, C, K, HEAD_MAX_K);
    const int row_blocks = dpooled ? std::min(cdiv(B, HEAD_THREADS / 32), 148) : 0;
    const int out_blocks = cdiv(K * C + K, HEAD_THREADS);
    const cudaError_t le = launch_pdl(head_ce_bwd_kernel, dim3(row_blocks + out_blocks), dim3(HEAD_THREADS), 0, (cudaStream_t)stream,
                                      dloss, dlogits, prob, labels, pooled, W, dpooled, dW, dbias, accumulate, row_blocks, B, C, K);
    if (le != cudaSuccess) { set_error("head_ce_bwd launch: %s", cudaGetErrorString(le)); return (int)le; }
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_weighted_scalar_sum(const float* const* terms, const float* weights, int n, float* out, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(terms && weights && out && n >= 1 && n <= 8, "1..8 device scalars");
    ScalarTerms t;
    t.n = n;
    for (int i = 0; i < n; ++i) { TSC_REQUIRE(terms[i] != nullptr, "NULL term %d", i); t.x[i] = terms[i]; t.w[i] = weights[i]; }
    const cudaError_t le = launch_pdl(weighted_scalar_sum_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, t, out);
    if (le != cudaSuccess) { set_error("weighted_scalar_sum launch: %s", cudaGetErrorString(le)); return (int)le; }
    TSC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
