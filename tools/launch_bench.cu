// Per-launch overhead of back-to-back dependent kernels inside a CUDA graph, for the launch shapes of the conv kernel:
// how much of a 23 us launch whose CTAs live 14-16 us is launch / drain latency?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/launch_bench tools/launch_bench.cu && tools/launch_bench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__global__ void k_empty(int* sink) { if (sink && threadIdx.x == 9999) *sink = 1; }

// spins for `cycles` SM clocks: a stand-in for a CTA lifetime
__global__ void k_spin(long long cycles, int* sink) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (sink && threadIdx.x == 9999) *sink = 1;
}

__global__ void k_tmem(long long cycles, int* sink) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(slot) : "memory");
    if (sink && threadIdx.x == 9999) *sink = 1;
}

template <typename F>
static float graph_time_us(F launch, int reps) {
    cudaStream_t s;
    cudaStreamCreate(&s);
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ge = nullptr;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < reps; ++i) launch(s);
    cudaError_t ce = cudaStreamEndCapture(s, &g);
    if (ce != cudaSuccess || g == nullptr) { printf("  (capture failed: %s)\n", cudaGetErrorString(ce)); cudaGetLastError(); cudaStreamDestroy(s); return -1.f; }
    ce = cudaGraphInstantiate(&ge, g, 0);
    if (ce != cudaSuccess || ge == nullptr) { printf("  (instantiate failed: %s)\n", cudaGetErrorString(ce)); cudaGetLastError(); cudaGraphDestroy(g); cudaStreamDestroy(s); return -1.f; }
    cudaGraphLaunch(ge, s);
    cudaStreamSynchronize(s);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
    cudaGraphLaunch(ge, s);
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(s);
    return ms * 1e3f / reps;
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    const int reps = 50;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("SM clock (attribute) %d MHz\n", clk_khz / 1000);
    cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(k_spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const int smems[3] = {0, 106 * 1024, 220 * 1024};
    for (int si = 0; si < 3; ++si) {
        const int sm = smems[si];
        for (int grid : {128, 1024}) {
            float t = graph_time_us([&](cudaStream_t s) { k_empty<<<grid, 192, sm, s>>>(nullptr); }, reps);
            printf("empty kernel      grid %4d block 192 smem %3d KB: %6.2f us per launch\n", grid, sm / 1024, t);
        }
    }
    for (long long cyc : {0LL, 10000LL, 20000LL, 30000LL}) {
        float a = graph_time_us([&](cudaStream_t s) { k_spin<<<128, 192, 106 * 1024, s>>>(cyc, nullptr); }, reps);
        float b = graph_time_us([&](cudaStream_t s) { k_tmem<<<128, 192, 106 * 1024, s>>>(cyc, nullptr); }, reps);
        printf("spin %6lld cycles grid 128 smem 106 KB: %6.2f us per launch; with a 256-column TMEM allocation: %6.2f us\n", cyc, a, b);
    }
    // alternating shared-memory carve-outs (conv: 106 KB, BatchNorm kernels: none), as inside the training step
    {
        float t = graph_time_us([&](cudaStream_t s) {
            k_spin<<<128, 192, 106 * 1024, s>>>(10000, nullptr);
            k_spin<<<592, 256, 0, s>>>(10000, nullptr);
        }, reps);
        float u = graph_time_us([&](cudaStream_t s) {
            k_spin<<<128, 192, 106 * 1024, s>>>(10000, nullptr);
            k_spin<<<128, 192, 106 * 1024, s>>>(10000, nullptr);
        }, reps);
        printf("pair of 10 k-cycle kernels: alternating carve-out (106 KB / 0 KB) %6.2f us per pair, same carve-out %6.2f us per pair\n", t, u);
    }
    // programmatic dependent launch: the next kernel's launch overlaps the tail of this one
    {
        float t = graph_time_us([&](cudaStream_t s) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(128); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = 106 * 1024; cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, k_spin, 20000LL, (int*)nullptr);
        }, reps);
        printf("spin  20000 cycles with the programmatic-serialization attribute (no griddepcontrol in the kernel): %6.2f us per launch\n", t);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
