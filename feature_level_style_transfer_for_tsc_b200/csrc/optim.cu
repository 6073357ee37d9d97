// Fused multi-tensor RMSprop over flat parameter / gradient / state buffers (one launch per step).
// Semantics of torch.optim.RMSprop defaults used by the reference (train_and_test.py:97-106):
//   v = alpha v + (1 - alpha) g^2 ;  p -= lr * g / (sqrt(v) + eps),   g = grad_scale * grad
// grad_scale carries the 1/N of the data-parallel gradient average (RMSprop is not linear in g, so the
// scale is applied to the gradient, never folded into the learning rate -- SURVEY 8e).
#include "common.cuh"

namespace tsc {
struct RmsGroups { int n; long long end[TSC_MAX_OPT_GROUPS]; float lr[TSC_MAX_OPT_GROUPS]; float clamp[TSC_MAX_OPT_GROUPS]; };

// WGAN weight clipping of a critic's group after its update (train_and_test.py:763-766); c <= 0: none
__device__ __forceinline__ float clip(float p, float c) { return c > 0.f ? fminf(fmaxf(p, -c), c) : p; }

__global__ void __launch_bounds__(256) rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ v, long long n, float alpha, float eps,
                                                      float grad_scale, const __grid_constant__ RmsGroups grp) {
    for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < n;
         i += (long long)gridDim.x * blockDim.x * 4) {
        int gi = 0;
        while (gi < grp.n - 1 && i >= grp.end[gi]) ++gi;
        if (i + 4 <= n && i + 4 <= grp.end[gi]) {
            const float lr = grp.lr[gi], cl = grp.clamp[gi];
            float4 pp = *reinterpret_cast<float4*>(p + i);
            const float4 gg = *reinterpret_cast<const float4*>(g + i);
            float4 vv = *reinterpret_cast<float4*>(v + i);
            float gx = gg.x * grad_scale, gy = gg.y * grad_scale, gz = gg.z * grad_scale, gw = gg.w * grad_scale;
            vv.x = alpha * vv.x + (1.f - alpha) * gx * gx; pp.x = clip(pp.x - lr * gx / (sqrtf(vv.x) + eps), cl);
            vv.y = alpha * vv.y + (1.f - alpha) * gy * gy; pp.y = clip(pp.y - lr * gy / (sqrtf(vv.y) + eps), cl);
            vv.z = alpha * vv.z + (1.f - alpha) * gz * gz; pp.z = clip(pp.z - lr * gz / (sqrtf(vv.z) + eps), cl);
            vv.w = alpha * vv.w + (1.f - alpha) * gw * gw; pp.w = clip(pp.w - lr * gw / (sqrtf(vv.w) + eps), cl);
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(v + i) = vv;
        } else {
            for (long long k = i; k < n && k < i + 4; ++k) {
                int gk = gi;
                while (gk < grp.n - 1 && k >= grp.end[gk]) ++gk;
                const float gr = g[k] * grad_scale;
                const float vn = alpha * v[k] + (1.f - alpha) * gr * gr;
                v[k] = vn;
                p[k] = clip(p[k] - grp.lr[gk] * gr / (sqrtf(vn) + eps), grp.clamp[gk]);
            }
        }
    }
}
}  // namespace tsc

extern "C" int tsc_rmsprop_step_clamped(float* params, const float* grads, float* square_avg, long long n,
                                        const long long* group_end, const float* group_lr, const float* group_clamp,
                                        int ngroups, float alpha, float eps, float grad_scale, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(params && grads && square_avg && group_end && group_lr, "NULL tensor");
    TSC_REQUIRE(ngroups >= 1 && ngroups <= TSC_MAX_OPT_GROUPS, "ngroups=%d outside [1,%d]", ngroups, TSC_MAX_OPT_GROUPS);
    TSC_REQUIRE(n > 0 && group_end[ngroups - 1] == n, "last group must end at n");
    TSC_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)square_avg) & 15) == 0, "buffers must be 16 B aligned");
    RmsGroups grp;
    grp.n = ngroups;
    for (int i = 0; i < ngroups; ++i) {
        grp.end[i] = group_end[i];
        grp.lr[i] = group_lr[i];
        grp.clamp[i] = group_clamp ? group_clamp[i] : 0.f;
    }
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    rmsprop_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, square_avg, n, alpha, eps, grad_scale, grp);
    TSC_LAUNCH_CHECK();
    return 0;
}

extern "C" int tsc_rmsprop_step(float* params, const float* grads, float* square_avg, long long n,
                                const long long* group_end, const float* group_lr, int ngroups, float alpha, float eps,
                                float grad_scale, tsc_stream_t stream) {
    return tsc_rmsprop_step_clamped(params, grads, square_avg, n, group_end, group_lr, nullptr, ngroups, alpha, eps,
                                    grad_scale, stream);
}
