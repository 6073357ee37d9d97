// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, bf16, M=128, cta_group::1) as a function of N and of the
// shared-memory operand layout.  Timing only (operands are zeros).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_bench tools/mma_bench.cu
#include "../feature_level_style_transfer_for_tsc_b200/csrc/tc_common.cuh"
#include <vector>
namespace tsc { void set_error(const char*, ...) {} }
using namespace tsc::tc;

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}
// mode 4: A copied smem -> TMEM by tcgen05.cp (128x256b) before every MMA, MMA reads A from TMEM (double buffered)
// mode 5: A from TMEM without the copy (pure TS-mode MMA rate)
// mode 0: SWIZZLE_NONE K-major A and B (conv kernel)   1: SWIZZLE_128B K-major A and B
// mode 2: SWIZZLE_NONE MN-major A and B (wgrad kernel)  3: SWIZZLE_NONE K-major, A start address shifted by 16 B per MMA
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int N, int mode, int reps, int ksteps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
        uint32_t idesc = make_idesc_bf16(128, N, mode == 2, mode == 2, false);
        bool dead = false;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int k = 0; k < ksteps; ++k) {
                uint64_t ad, bd;
                if (mode >= 4) {
                    ad = make_smem_desc(a0 + k * 2 * 2560 + (r % 31) * 16, 2560, 128);
                    bd = make_smem_desc(b0 + k * 2 * N * 16, N * 16, 128);
                    const uint32_t a_t = tm + 256 + ((r * ksteps + k) & 1) * 8;
                    if (mode == 4) tmem_cp_128x256b(a_t, ad);
                    umma_bf16_ts(tm, a_t, bd, idesc, true);
                    continue;
                }
                if (mode == 1) {
                    ad = make_smem_desc(a0 + k * 32, 16, 1024) | ((uint64_t)2 << 61);
                    bd = make_smem_desc(b0 + k * 32, 16, 1024) | ((uint64_t)2 << 61);
                } else if (mode == 2) {
                    ad = make_smem_desc(a0 + k * 256, 128, 2048);
                    bd = make_smem_desc(b0 + k * 256 + (r % 3) * 16, 128, 2176);
                } else {
                    ad = make_smem_desc(a0 + k * 2 * 2560 + (mode == 3 ? (r % 31) * 16 : 0), 2560, 128);
                    bd = make_smem_desc(b0 + k * 2 * N * 16, N * 16, 128);
                }
                umma_bf16(tm, ad, bd, idesc, true);
            }
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0, dead, 9);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int Ns[] = {32, 64, 96, 128, 160, 192, 240, 256};
    const char* names[] = {"none/K-major", "sw128/K-major", "none/MN-major", "none/K-major+rowshift", "cp+TS-mma", "TS-mma only"};
    for (int grid : {148}) {
        for (int mode = 3; mode < 6; ++mode) {
            for (int N : Ns) {
                const int reps = 200, ksteps = 5;
                mma_bench_kernel<<<grid, 128, 160 * 1024>>>(N, mode, reps, ksteps, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                std::vector<long long> h(grid);
                cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (auto v : h) mx = v > mx ? v : mx;
                const double per = (double)mx / (reps * ksteps);
                printf("grid=%3d mode=%-22s N=%3d : %7.1f cycles/MMA  ideal %5.1f  -> %6.0f MAC/clk/SM\n", grid, names[mode], N, per,
                       128.0 * N / 256.0, 128.0 * N * 16 / per);
            }
        }
    }
    return 0;
}
