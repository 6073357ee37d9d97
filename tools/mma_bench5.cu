// Micro-benchmark 5: is a chain of tcgen05.mma accumulating into ONE TMEM tile limited by the dependent-accumulate
// latency rather than by issue or operand fetch?  The conv kernel's loop (groups of four from a converged warp) with the
// destination alternating between `tiles` independent accumulators, per MMA or per group.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_bench5 tools/mma_bench5.cu
#include "../feature_level_style_transfer_for_tsc_b200/csrc/tc_common.cuh"
#include <vector>
namespace tsc { void set_error(const char*, ...) {} }
using namespace tsc::tc;

__global__ void __launch_bounds__(128, 1) k(int N, int tiles, int per_group, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[1];
    __shared__ uint32_t slot;
    __shared__ uint4 table[256];
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_barrier_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t a0 = smem_u32(smem) >> 4, b0 = (smem_u32(smem) + 32 * 1024) >> 4;
    const int stride = 512 / tiles;                      // TMEM columns between accumulators
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const int t = (i * 7) % 31, kp = i % 5;
        const int tile = per_group ? (i / 4) % tiles : i % tiles;
        table[i] = make_uint4((a0 + (2 * kp) * 160 + t) | (160u << 16), (b0 + ((i * 3) % 8) * 2 * N % 3500) | ((uint32_t)N << 16),
                              make_idesc_bf16(128, N, false, false, false), tm + (uint32_t)(tile * stride));
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t hi = (128u >> 4) | (1u << 14);
    if (warp == 0) {
        bool dead = false;
        long long t0 = clock64();
        uint4 e0 = table[0], e1 = table[1], e2 = table[2], e3 = table[3];
        for (int r = 0; r < reps; r += 4) {
            const int nb = (r + 4) & 255;
            const uint4 f0 = table[nb], f1 = table[nb + 1], f2 = table[nb + 2], f3 = table[nb + 3];
            if (elect_one()) {
                umma_bf16(e0.w, ((uint64_t)hi << 32) | e0.x, ((uint64_t)hi << 32) | e0.y, e0.z, 1u);
                umma_bf16(e1.w, ((uint64_t)hi << 32) | e1.x, ((uint64_t)hi << 32) | e1.y, e1.z, 1u);
                umma_bf16(e2.w, ((uint64_t)hi << 32) | e2.x, ((uint64_t)hi << 32) | e2.y, e2.z, 1u);
                umma_bf16(e3.w, ((uint64_t)hi << 32) | e3.x, ((uint64_t)hi << 32) | e3.y, e3.z, 1u);
            }
            e0 = f0; e1 = f1; e2 = f2; e3 = f3;
        }
        __syncwarp();
        if (elect_one()) tc_commit(&bar[0]);
        __syncwarp();
        mbar_wait(&bar[0], 0, dead, 9);
        long long t1 = clock64();
        if (lane == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int tiles : {1, 2, 4})
        for (int per_group : {0, 1})
            for (int N : {32, 48, 96, 128, 192, 240}) {
                if (N > 512 / tiles || (tiles == 1 && per_group)) continue;
                const int reps = 800;
                k<<<148, 128, 160 * 1024>>>(N, tiles, per_group, reps, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                std::vector<long long> h(148);
                cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
                printf("accumulators=%d alternate per %s  N=%3d : %6.1f cycles per MMA (tensor time %3d)\n", tiles, per_group ? "group" : "MMA  ", N,
                       (double)mx / reps, N / 2);
            }
    return 0;
}
