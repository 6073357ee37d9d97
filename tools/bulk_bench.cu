// Micro-benchmark: per-SM throughput of cp.async.bulk (1-D TMA) global -> shared copies from an L2-resident buffer,
// as the conv kernel's weight producer issues them: a ring of NS stages of `stage` bytes, 16 x 32 KB per pass.
#include "../feature_level_style_transfer_for_tsc_b200/csrc/tc_common.cuh"
#include <vector>
namespace tsc { void set_error(const char*, ...) {} }
using namespace tsc::tc;

__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, int total, int stage, int NS, int split, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[8];
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        bool dead = false;
        const int n = total / stage;
        long long t0 = clock64();
        uint32_t ph = 0;
        int issued = 0, done = 0;
        // keep NS copies in flight; a consumed stage is re-armed at once (no MMA in this benchmark)
        for (; issued < NS && issued < n; ++issued) {
            mbar_arrive_expect_tx(&full[issued], (uint32_t)stage);
            for (int j = 0; j < split; ++j)
                bulk_load(smem + (size_t)issued * stage + (size_t)j * (stage / split), src + (size_t)issued * stage + (size_t)j * (stage / split),
                          (uint32_t)(stage / split), &full[issued]);
        }
        int s = 0;
        while (done < n) {
            mbar_wait(&full[s], ph, dead, 1);
            ++done;
            if (issued < n) {
                mbar_arrive_expect_tx(&full[s], (uint32_t)stage);
                for (int j = 0; j < split; ++j)
                    bulk_load(smem + (size_t)s * stage + (size_t)j * (stage / split), src + (size_t)issued * stage + (size_t)j * (stage / split),
                              (uint32_t)(stage / split), &full[s]);
                ++issued;
            }
            if (++s == NS) { s = 0; ph ^= 1u; }
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    const int total = 512 * 1024;
    uint8_t* src; long long* d;
    cudaMalloc(&src, total); cudaMemset(src, 1, total);
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int grid : {1, 8, 128})
        for (int stage : {8192, 32768})
            for (int NS : {2, 5})
                for (int split : {1, 4}) {
                    if (NS * stage > 200 * 1024) continue;
                    for (int rep = 0; rep < 2; ++rep) k<<<grid, 128, NS * stage>>>(src, total, stage, NS, split, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    std::vector<long long> h(148);
                    cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                    printf("grid=%3d stage=%5d NS=%d split=%d : %7lld cycles for 512 KB -> %6.1f B/clk/SM\n", grid, stage, NS, split, mx, (double)total / mx);
                }
    return 0;
}
