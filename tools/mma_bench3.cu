// Micro-benchmark 3: cost per tcgen05.mma of different ISSUE-LOOP shapes when every operand comes from a
// per-instruction table in shared memory (the conv kernel's plan).
//   variant 0: lane-0-only branch (divergent), uint4 entry per MMA, all four operands from the table
//   variant 1: whole warp runs the loop (uniform control flow), MMA under elect_one()
//   variant 2: like 1, entries prefetched one iteration ahead
//   variant 3: like 0 but idesc / tmem column loop-invariant (only the two descriptors change)
//   variant 4: like 1 but idesc / tmem column loop-invariant
#include "../feature_level_style_transfer_for_tsc_b200/csrc/tc_common.cuh"
#include <vector>
namespace tsc { void set_error(const char*, ...) {} }
using namespace tsc::tc;

__global__ void __launch_bounds__(128, 1) k(int N, int variant, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    __shared__ uint4 table[256];
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    const uint32_t a0 = smem_u32(smem) >> 4, b0 = (smem_u32(smem) + 64 * 1024) >> 4;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const int n = (variant >= 3) ? N : max(16, N - 16 * (i % 3));
        table[i] = make_uint4((a0 + (i % 31)) | (160u << 16), (b0 + (i % 5) * 2 * n) | ((uint32_t)n << 16),
                              make_idesc_bf16(128, n, false, false, false), (uint32_t)(N - n));
    }
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t hi = (128u >> 4) | (1u << 14);
    if (warp == 0) {
        bool dead = false;
        long long t0 = clock64();
        if (variant == 0 || variant == 3) {
            if (lane == 0) {
                const uint32_t idc = table[0].z;
                for (int r = 0; r < reps; ++r) {
                    const uint4 e = table[r & 255];
                    umma_bf16(variant == 3 ? tm : tm + e.w, ((uint64_t)hi << 32) | e.x, ((uint64_t)hi << 32) | e.y,
                              variant == 3 ? idc : e.z, true);
                }
                tc_commit(&bar[0]);
            }
        } else if (variant == 1 || variant == 4) {
            const uint32_t idc = table[0].z;
            for (int r = 0; r < reps; ++r) {
                const uint4 e = table[r & 255];
                if (elect_one())
                    umma_bf16(variant == 4 ? tm : tm + e.w, ((uint64_t)hi << 32) | e.x, ((uint64_t)hi << 32) | e.y,
                              variant == 4 ? idc : e.z, true);
                __syncwarp();
            }
            if (elect_one()) tc_commit(&bar[0]);
        } else {
            uint4 e = table[0];
            for (int r = 0; r < reps; ++r) {
                const uint4 en = table[(r + 1) & 255];
                if (elect_one())
                    umma_bf16(tm + e.w, ((uint64_t)hi << 32) | e.x, ((uint64_t)hi << 32) | e.y, e.z, true);
                e = en;
            }
            __syncwarp();
            if (elect_one()) tc_commit(&bar[0]);
        }
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
    for (int variant = 0; variant < 5; ++variant)
        for (int N : {64, 128, 240}) {
            const int reps = 800;
            cudaMemset(d, 0, 148 * sizeof(long long));
            k<<<148, 128, 160 * 1024>>>(N, variant, reps, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<long long> h(148);
            cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (auto v : h) mx = v > mx ? v : mx;
            printf("variant=%d N<=%3d : %7.1f cycles per MMA\n", variant, N, (double)mx / reps);
        }
    return 0;
}
