// Shared helpers for libtsc_b200 (sm_100a).  See include/tsc_b200.h for the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <mutex>
#include "../../include/tsc_b200.h"

namespace tsc {

// ---- error plumbing -----------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define TSC_REQUIRE(cond, ...)                       \
    do {                                             \
        if (!(cond)) {                               \
            ::tsc::set_error(__VA_ARGS__);           \
            return -1;                               \
        }                                            \
    } while (0)
#define TSC_LAUNCH_CHECK()                                            \
    do {                                                              \
        cudaError_t e__ = cudaGetLastError();                         \
        if (e__ != cudaSuccess) {                                     \
            ::tsc::set_error("%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return (int)e__;                                          \
        }                                                             \
    } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------
// The step is a chain of ~100 short dependent kernels.  Kernels launched through launch_pdl carry the
// programmatic-stream-serialization attribute: the next kernel of the stream may be scheduled, and run its
// prologue (barrier init, TMEM allocation, shared-memory clearing), while this one drains; it blocks in
// pdl_wait() -- which every such kernel executes before its first global-memory access -- until the whole
// preceding grid has completed and flushed.  pdl_trigger() marks the point from which dependents may be launched.
// The attribute is opt-in (TSC_PDL=1): without it the device-side calls are no-ops.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// One-time, thread-safe opt-in of a kernel to the full 227 KB of dynamic shared memory (function attributes are process
// state: this is the only host-side state the launch wrappers keep besides the geometry caches, both behind a lock).
struct OnceAttr {
    std::once_flag flag;
    cudaError_t err = cudaSuccess;
};
template <typename F>
static inline cudaError_t run_once(OnceAttr& o, F&& f) {
    std::call_once(o.flag, [&] { o.err = f(); });
    return o.err;
}

static inline int pad16(int c) { return (c + 15) & ~15; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- per-bank tap table (host side; passed to kernels by value) -----------------------------------
// Generic implicit-GEMM conv:  y[b,n,l] = bias[n] + sum_t sum_{kc>=kc_lo[t]} sum_j
//      x[b, kc*8+j, l + t - pad_left] * blob_t[kc-kc_lo[t]][n-n_lo[t]][j]      for n >= n_lo[t]
// blob_t is tap t's [kc - kc_lo][np - n_lo][8] block; w_off[t] = rows of the blobs before it in issue order.  In memory
// the rows are split into two streams, see packed_row().
struct TapTable {
    int taps;        // Kmax
    int pad_left;
    int kc;          // input-channel chunks of 8 (padded input channels / 8)
    int np;          // padded output channels (multiple of 16)
    int n_order;     // number of live taps
    int total_rows;  // total 8-element rows of the packed buffer
    short order[TSC_MAX_TAPS];   // live taps; the first one has n_lo == 0 && kc_lo == 0 coverage of all n
    short n_lo[TSC_MAX_TAPS];
    short kc_lo[TSC_MAX_TAPS];
    int w_off[TSC_MAX_TAPS];
};

// Row (16 B = 8 elements) of the packed kernel bank that holds (tap t, input-channel chunk `chunk_rel` past the tap's first
// stored chunk, output row `row_rel` past the tap's first stored row); nt = stored rows of the tap (a multiple of 16),
// w_off_t = TapTable::w_off[t], total_rows = TapTable::total_rows.
//   classic layout (split == 0): [tap in issue order][chunk][nt rows][8] -- the B operand of one CTA's MMA is contiguous;
//   split layout   (split != 0): two streams -- the lower and the upper half of every tap's stored rows -- each
//     [tap in issue order][chunk][nt / 2 rows][8]: the CTA-pair variant of the tcgen05 convolution (cta_group::2, M = 256)
//     holds N/2 rows of the B operand in each CTA, so each CTA streams ONE contiguous half of the bank.
// Which one a process uses is fixed at start-up (packed_split(): TSC_CONV_PAIR=1 selects the pair kernel and the split layout);
// every kernel that writes or reads a packed bank goes through this function.
__host__ __device__ __forceinline__ long long packed_row(int total_rows, int w_off_t, int nt, int chunk_rel, int row_rel, int split) {
    if (!split) return (long long)w_off_t + (long long)chunk_rel * nt + row_rel;
    const int half = nt >> 1;
    const int h = row_rel >= half ? 1 : 0;
    return (long long)h * (total_rows >> 1) + (w_off_t >> 1) + (long long)chunk_rel * half + (row_rel - h * half);
}
bool packed_split();      // api.cu

// Build the table for one direction.  Returns 0 or -1 (error text set).
int build_tap_table(int direction, int Cin, int Cout, int Kmax, const int* s_of_tap, TapTable* tt);

// s(t) by value (kernel parameter): first live out channel per weight tap.
struct STable { int s[TSC_MAX_TAPS]; };
static inline void fill_stable(STable* st, const int* s_of_tap, int Kmax) {
    for (int t = 0; t < TSC_MAX_TAPS; ++t) st->s[t] = t < Kmax ? s_of_tap[t] : 0x7fffffff;
}

// Ordered reduction of wgrad partials [S][Kmax][np][kcp] -> dW [Cout][Cin][Kmax] (layout.cu).
int launch_wgrad_reduce(const float* part, float* dW, int S, int Cin, int Cout, int Kmax, int np, int kcp,
                        const int* s_of_tap, int accumulate, cudaStream_t stream);

// ---- dtype helpers ----------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive channels of one c8 row
template <typename T> struct Row8;
template <> struct Row8<float> {
    float v[8];
    __device__ __forceinline__ void load(const float* p) {
        float4 a = *reinterpret_cast<const float4*>(p);
        float4 b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct Row8<__nv_bfloat16> {
    float v[8];
    __device__ __forceinline__ void load(const __nv_bfloat16* p) {
        uint4 raw = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const {
        uint4 raw;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = raw;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Chan/Welford merge of (n, mean, M2) pairs
__device__ __forceinline__ void welford_merge(float& n, float& mean, float& m2, float nb, float meanb, float m2b) {
    if (nb == 0.f) return;
    float nt = n + nb;
    float delta = meanb - mean;
    float f = nb / nt;
    mean += delta * f;
    m2 += m2b + delta * delta * n * f;
    n = nt;
}

}  // namespace tsc
