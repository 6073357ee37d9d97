// Micro-benchmark 4: does the tap shift of the SWIZZLE_NONE A tile (start address at 16 B granularity, so every
// 128 B core matrix straddles two 128 B lines) slow tcgen05.mma down?  One converged warp issues groups of four MMAs
// (the conv kernel's loop) over a 25 KB A tile and a 32 KB B stage with
//   mode 0: A row shift = multiple of 8 rows (aligned core matrices)      mode 1: arbitrary row shift (the conv case)
#include "../feature_level_style_transfer_for_tsc_b200/csrc/tc_common.cuh"
#include <string>
#include <vector>
namespace tsc { void set_error(const char*, ...) {} }
using namespace tsc::tc;

__global__ void __launch_bounds__(128, 1) k(int N, int mode, int reps, long long* out, const uint8_t* gsrc, int stream_bytes) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint64_t cp_full[2];
    __shared__ volatile int stop_flag;
    __shared__ uint32_t slot;
    __shared__ uint4 table[256];
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    const uint32_t a0 = smem_u32(smem) >> 4, b0 = (smem_u32(smem) + 32 * 1024) >> 4;   // A tile: 32 KB, B stage area: 64 KB
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const int t = mode == 0 ? ((i * 7) % 4) * 8 : (i * 7) % 31;              // tap = row shift
        const int kp = i % 5;                                                    // k-pair: 2 chunks of 160 rows
        table[i] = make_uint4((a0 + (2 * kp) * 160 + t) | (160u << 16), (b0 + ((i * 3) % 8) * 2 * N % 3500) | ((uint32_t)N << 16),
                              make_idesc_bf16(128, N, false, false, false), 0u);
    }
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&cp_full[0], 1); mbar_init(&cp_full[1], 1); stop_flag = 0; fence_barrier_init(); }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc(&slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
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
                umma_bf16(tm, ((uint64_t)hi << 32) | e0.x, ((uint64_t)hi << 32) | e0.y, e0.z, 1u);
                umma_bf16(tm, ((uint64_t)hi << 32) | e1.x, ((uint64_t)hi << 32) | e1.y, e1.z, 1u);
                umma_bf16(tm, ((uint64_t)hi << 32) | e2.x, ((uint64_t)hi << 32) | e2.y, e2.z, 1u);
                umma_bf16(tm, ((uint64_t)hi << 32) | e3.x, ((uint64_t)hi << 32) | e3.y, e3.z, 1u);
            }
            e0 = f0; e1 = f1; e2 = f2; e3 = f3;
        }
        __syncwarp();
        if (elect_one()) tc_commit(&bar[0]);
        __syncwarp();
        mbar_wait(&bar[0], 0, dead, 9);
        long long t1 = clock64();
        if (lane == 0) { out[blockIdx.x] = t1 - t0; stop_flag = 1; }
    } else if (warp == 1 && lane == 0 && stream_bytes > 0) {
        // a weight producer beside the MMAs: a ring of two bulk copies of stream_bytes each into the upper smem
        bool dead = false;
        uint32_t ph = 0;
        uint8_t* dst = smem + 96 * 1024;
        for (int i = 0; i < 2; ++i) { mbar_arrive_expect_tx(&cp_full[i], (uint32_t)stream_bytes); bulk_load(dst + i * stream_bytes, gsrc + i * stream_bytes, (uint32_t)stream_bytes, &cp_full[i]); }
        int s = 0, n = 2;
        while (!stop_flag) {
            mbar_wait(&cp_full[s], ph, dead, 2);
            mbar_arrive_expect_tx(&cp_full[s], (uint32_t)stream_bytes);
            bulk_load(dst + s * stream_bytes, gsrc + (size_t)(n % 8) * stream_bytes, (uint32_t)stream_bytes, &cp_full[s]);
            ++n;
            if (++s == 2) { s = 0; ph ^= 1u; }
        }
        mbar_wait(&cp_full[0], s == 0 ? ph : ph ^ 1u, dead, 3);
        mbar_wait(&cp_full[1], ph, dead, 4);
        out[148 + blockIdx.x] = n;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}

int main() {
    long long* d;
    uint8_t* g;
    cudaMalloc(&d, 2 * 148 * sizeof(long long));
    cudaMalloc(&g, 8 * 32768);
    cudaMemset(g, 0, 8 * 32768);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int mode = 1; mode < 3; ++mode)
        for (int N : {48, 96, 144, 192, 240}) {
            const int reps = 800;
            k<<<148, 128, 160 * 1024>>>(N, 1, reps, d, g, mode == 2 ? 32768 : 0);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<long long> h(296);
            cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("%s  N=%3d : %6.1f cycles per MMA (tensor time %3d)%s\n", mode == 2 ? "with a concurrent 32 KB bulk-copy ring" : "MMAs alone                           ", N,
                   (double)mx / reps, N / 2, mode == 2 ? (std::string("  copies/CTA: ") + std::to_string(h[148])).c_str() : "");
        }
    return 0;
}
