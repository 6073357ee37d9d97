// Micro-benchmark 2: is the ~140-cycle cost of a small-N tcgen05.mma an issue limit of the single issuing thread?
//   variant 0: one thread, descriptors rebuilt per MMA (as in the conv kernel)
//   variant 1: one thread, precomputed descriptors, 8x unrolled
//   variant 2: two threads (two warps) issuing concurrently into different accumulators
//   variant 3: four threads (four warps)
#include "../feature_level_style_transfer_for_tsc_b200/csrc/tc_common.cuh"
#include <vector>
namespace tsc { void set_error(const char*, ...) {} }
using namespace tsc::tc;

__global__ void __launch_bounds__(128, 1) k(int N, int variant, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nissuers = variant == 2 ? 2 : (variant == 3 ? 4 : 1);
    const int NN = nissuers > 1 ? min(N, 512 / nissuers) : N;
    long long t0 = 0, t1 = 0;
    if (lane == 0 && warp < nissuers) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
        const uint32_t idesc = make_idesc_bf16(128, NN, false, false, false);
        const uint32_t d = tm + warp * (512 / nissuers);
        bool dead = false;
        uint64_t ad[8], bd[8];
        for (int i = 0; i < 8; ++i) { ad[i] = make_smem_desc(a0 + i * 16, 2560, 128); bd[i] = make_smem_desc(b0 + i * 256, NN * 16, 128); }
        t0 = clock64();
        if (variant == 0) {
            for (int r = 0; r < reps; ++r)
                umma_bf16(d, make_smem_desc(a0 + (r % 31) * 16, 2560, 128), make_smem_desc(b0 + (r % 5) * 2 * NN * 16, NN * 16, 128), idesc, true);
        } else {
            for (int r = 0; r < reps; r += 8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) umma_bf16(d, ad[i], bd[i], idesc, true);
            }
        }
        tc_commit(&bar[warp]);
        mbar_wait(&bar[warp], 0, dead, 9);
        t1 = clock64();
        out[blockIdx.x * 4 + warp] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * 4 * sizeof(long long));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int variant = 0; variant < 4; ++variant)
        for (int N : {32, 64, 128, 256}) {
            const int reps = 800;
            cudaMemset(d, 0, 148 * 4 * sizeof(long long));
            k<<<148, 128, 160 * 1024>>>(N, variant, reps, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<long long> h(148 * 4);
            cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (auto v : h) mx = v > mx ? v : mx;
            const int nissuers = variant == 2 ? 2 : (variant == 3 ? 4 : 1);
            const int NN = nissuers > 1 ? std::min(N, 512 / nissuers) : N;
            printf("variant=%d issuers=%d N=%3d : %7.1f cycles per MMA per issuer, %7.1f cycles per MMA overall -> %6.0f MAC/clk/SM\n",
                   variant, nissuers, NN, (double)mx / reps, (double)mx / (reps * nissuers), 128.0 * NN * 16 * reps * nissuers / mx);
        }
    return 0;
}
