// Layout conversion (NCL fp32 <-> c8) and kernel-bank packing.  HBM-bound helpers.
#include "common.cuh"

namespace tsc {

// ---- NCL fp32 -> c8 (T).  One thread per (b, chunk, l) row of 8 channels.  Reads are coalesced
// along l for each of the 8 channels (8 x 128 B per warp), the write is one 16/32 B row per thread
// (contiguous across the warp). ------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) ncl_to_c8_kernel(const float* __restrict__ src, T* __restrict__ dst,
                                                         int B, int C, int Cpc, int L) {
    pdl_trigger();
    pdl_wait();
    const long long total = (long long)B * Cpc * L;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i % L);
        const long long bc = i / L;
        const int ch = (int)(bc % Cpc);
        const int b = (int)(bc / Cpc);
        Row8<T> r;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch * 8 + j;
            r.v[j] = c < C ? __ldg(src + ((long long)b * C + c) * L + l) : 0.f;
        }
        r.store(dst + i * 8);
    }
}

__global__ void __launch_bounds__(256) c8_to_ncl_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                         int B, int C, int Cpc, int L) {
    pdl_trigger();
    pdl_wait();
    const long long total = (long long)B * Cpc * L;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i % L);
        const long long bc = i / L;
        const int ch = (int)(bc % Cpc);
        const int b = (int)(bc / Cpc);
        Row8<float> r;
        r.load(src + i * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch * 8 + j;
            if (c < C) dst[((long long)b * C + c) * L + l] = r.v[j];
        }
    }
}

// ---- kernel-bank packing -------------------------------------------------------------------------
// s(t) travels to the device inside the kernel parameter block (no allocation, graph safe).
// blockIdx.y = index into tt.order; threads stride over the tap's blob [kc-kc_lo][np-n_lo][8].
static int grid_for(long long n, int block = 256) {
    long long g = (n + block - 1) / block;
    const long long cap = 148LL * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <typename T>
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ W, T* __restrict__ packed,
                                                             int direction, int Cin, int Cout, int Kmax,
                                                             const __grid_constant__ TapTable tt,
                                                             const __grid_constant__ STable st, int split) {
    const int t = tt.order[blockIdx.y];
    const int n_lo = tt.n_lo[t], kc_lo = tt.kc_lo[t];
    const int nrows = tt.np - n_lo;
    const int elems = (tt.kc - kc_lo) * nrows * 8;
    const int wt = direction == TSC_DIR_FWD ? t : Kmax - 1 - t;
    const int s = st.s[wt];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < elems; e += gridDim.x * blockDim.x) {
        const int j = e & 7;
        const int n = n_lo + (e >> 3) % nrows;
        const int kch = kc_lo + (e >> 3) / nrows;
        const int k = kch * 8 + j;
        const int co = direction == TSC_DIR_FWD ? n : k;
        const int ci = direction == TSC_DIR_FWD ? k : n;
        float v = 0.f;
        if (co < Cout && ci < Cin && co >= s) v = W[((long long)co * Cin + ci) * Kmax + wt];
        packed[packed_row(tt.total_rows, tt.w_off[t], nrows, kch - kc_lo, n - n_lo, split) * 8 + j] = from_f32<T>(v);
    }
}

// One launch per layer: forward pack + dgrad pack + in-place zeroing of the masked taps.
// blockIdx.y in [0, nf): forward taps; [nf, nf+nd): dgrad taps; nf+nd: the zeroing slab.
// Packing reads only live (co >= s(t)) weights and the zeroing writes only masked ones: no race.
template <typename T>
__global__ void __launch_bounds__(256) pack_pair_kernel(float* __restrict__ W, T* __restrict__ pf, T* __restrict__ pd,
                                                         int Cin, int Cout, int Kmax,
                                                         const __grid_constant__ TapTable tf,
                                                         const __grid_constant__ TapTable td,
                                                         const __grid_constant__ STable st, int nd, int zero_masked, int split) {
    const int nf = tf.n_order;
    int y = blockIdx.y;
    if (y >= nf + nd) {
        if (!zero_masked) return;
        const int total = Cout * Cin * Kmax;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
            const int t = i % Kmax;
            const int co = i / (Kmax * Cin);
            if (co < st.s[t]) W[i] = 0.f;
        }
        return;
    }
    const bool fwd = y < nf;
    const TapTable& tt = fwd ? tf : td;
    if (!fwd) y -= nf;
    const int t = tt.order[y];
    const int n_lo = tt.n_lo[t], kc_lo = tt.kc_lo[t];
    const int nrows = tt.np - n_lo;
    const int elems = (tt.kc - kc_lo) * nrows * 8;
    const int wt = fwd ? t : Kmax - 1 - t;
    const int s = st.s[wt];
    T* out = fwd ? pf : pd;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < elems; e += gridDim.x * blockDim.x) {
        const int j = e & 7;
        const int n = n_lo + (e >> 3) % nrows;
        const int kch = kc_lo + (e >> 3) / nrows;
        const int k = kch * 8 + j;
        const int co = fwd ? n : k;
        const int ci = fwd ? k : n;
        float v = 0.f;
        if (co < Cout && ci < Cin && co >= s) v = W[((long long)co * Cin + ci) * Kmax + wt];
        out[packed_row(tt.total_rows, tt.w_off[t], nrows, kch - kc_lo, n - n_lo, split) * 8 + j] = from_f32<T>(v);
    }
}

__global__ void __launch_bounds__(256) zero_masked_kernel(float* __restrict__ W, int Cin, int Cout, int Kmax,
                                                            const __grid_constant__ STable st) {
    const int total = Cout * Cin * Kmax;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int t = i % Kmax;
        const int co = i / (Kmax * Cin);
        if (co < st.s[t]) W[i] = 0.f;
    }
}

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dW,
                                                             int S, int Cin, int Cout, int Kmax, int np, int kcp,
                                                             const __grid_constant__ STable st, int accumulate) {
    const int total = Cout * Cin * Kmax;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int t = i % Kmax;
        const int ci = (i / Kmax) % Cin;
        const int co = i / (Kmax * Cin);
        float acc = 0.f;
        if (co >= st.s[t]) {
            const long long stride = (long long)Kmax * np * kcp;
            const float* p = part + ((long long)t * np + co) * kcp + ci;
            for (int s = 0; s < S; ++s) acc += p[s * stride];
        }
        dW[i] = accumulate ? dW[i] + acc : acc;
    }
}

int launch_wgrad_reduce(const float* part, float* dW, int S, int Cin, int Cout, int Kmax, int np, int kcp,
                        const int* s_of_tap, int accumulate, cudaStream_t stream) {
    STable st;
    for (int t = 0; t < TSC_MAX_TAPS; ++t) st.s[t] = t < Kmax ? s_of_tap[t] : 0x7fffffff;
    wgrad_reduce_kernel<<<grid_for((long long)Cout * Cin * Kmax), 256, 0, stream>>>(part, dW, S, Cin, Cout, Kmax,
                                                                                    np, kcp, st, accumulate);
    TSC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tsc

extern "C" {

int tsc_ncl_to_c8(const float* src, void* dst, int dst_dtype, int B, int C, int L, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(src && dst, "NULL tensor");
    TSC_REQUIRE(B > 0 && C > 0 && L > 0, "bad shape [%d,%d,%d]", B, C, L);
    const int Cpc = pad16(C) / 8;
    const long long rows = (long long)B * Cpc * L;
    cudaStream_t st = (cudaStream_t)stream;
    if (dst_dtype == TSC_BF16)
        launch_pdl(ncl_to_c8_kernel<__nv_bfloat16>, dim3(grid_for(rows)), dim3(256), 0, st, src, (__nv_bfloat16*)dst, B, C, Cpc, L);
    else if (dst_dtype == TSC_F32)
        launch_pdl(ncl_to_c8_kernel<float>, dim3(grid_for(rows)), dim3(256), 0, st, src, (float*)dst, B, C, Cpc, L);
    else
        TSC_REQUIRE(false, "bad dtype %d", dst_dtype);
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_c8_to_ncl(const float* src, float* dst, int B, int C, int L, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(src && dst, "NULL tensor");
    TSC_REQUIRE(B > 0 && C > 0 && L > 0, "bad shape [%d,%d,%d]", B, C, L);
    const int Cpc = pad16(C) / 8;
    launch_pdl(c8_to_ncl_kernel, dim3(grid_for((long long)B * Cpc * L)), dim3(256), 0, (cudaStream_t)stream, src, dst, B, C, Cpc, L);
    TSC_LAUNCH_CHECK();
    return 0;
}

int tsc_pack_weights(int direction, int dtype, float* W, void* packed, int Cin, int Cout, int Kmax,
                     const int* s_of_tap, int zero_masked, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(W && packed, "NULL tensor");
    TapTable tt;
    if (build_tap_table(direction, Cin, Cout, Kmax, s_of_tap, &tt) != 0) return -1;
    STable st;
    for (int t = 0; t < TSC_MAX_TAPS; ++t) st.s[t] = t < Kmax ? s_of_tap[t] : 0x7fffffff;
    cudaStream_t cs = (cudaStream_t)stream;
    if (zero_masked) {
        zero_masked_kernel<<<grid_for((long long)Cout * Cin * Kmax), 256, 0, cs>>>(W, Cin, Cout, Kmax, st);
        TSC_LAUNCH_CHECK();
    }
    const int max_elems = tt.kc * tt.np * 8;
    dim3 grid(cdiv(max_elems, 256 * 4), tt.n_order);
    if (dtype == TSC_BF16)
        pack_weights_kernel<__nv_bfloat16><<<grid, 256, 0, cs>>>(W, (__nv_bfloat16*)packed, direction, Cin, Cout, Kmax, tt, st, packed_split() ? 1 : 0);
    else if (dtype == TSC_F32)
        pack_weights_kernel<float><<<grid, 256, 0, cs>>>(W, (float*)packed, direction, Cin, Cout, Kmax, tt, st, packed_split() ? 1 : 0);
    else
        TSC_REQUIRE(false, "bad dtype %d", dtype);
    TSC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

extern "C" int tsc_pack_weights_pair(int dtype, float* W, void* packed_fwd, void* packed_dgrad, int Cin, int Cout,
                                     int Kmax, const int* s_of_tap, int zero_masked, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(W && packed_fwd, "NULL tensor");
    TapTable tf, td;
    if (build_tap_table(TSC_DIR_FWD, Cin, Cout, Kmax, s_of_tap, &tf) != 0) return -1;
    if (build_tap_table(TSC_DIR_DGRAD, Cin, Cout, Kmax, s_of_tap, &td) != 0) return -1;
    STable st;
    fill_stable(&st, s_of_tap, Kmax);
    const int nd = packed_dgrad ? td.n_order : 0;
    const int max_elems = max(tf.kc * tf.np, td.kc * td.np) * 8;
    dim3 grid(cdiv(max_elems, 256 * 4), tf.n_order + nd + (zero_masked ? 1 : 0));
    cudaStream_t cs = (cudaStream_t)stream;
    if (dtype == TSC_BF16)
        pack_pair_kernel<__nv_bfloat16><<<grid, 256, 0, cs>>>(W, (__nv_bfloat16*)packed_fwd, (__nv_bfloat16*)packed_dgrad, Cin,
                                                            Cout, Kmax, tf, td, st, nd, zero_masked, packed_split() ? 1 : 0);
    else if (dtype == TSC_F32)
        pack_pair_kernel<float><<<grid, 256, 0, cs>>>(W, (float*)packed_fwd, (float*)packed_dgrad, Cin, Cout, Kmax, tf, td,
                                                    st, nd, zero_masked, packed_split() ? 1 : 0);
    else
        TSC_REQUIRE(false, "bad dtype %d", dtype);
    TSC_LAUNCH_CHECK();
    return 0;
}

// =====================================================================================================
// Kernel-bank packing for a whole stack of layers in ONE launch (tsc_pack_weights_multi).
// Block = (layer, chunk of 8 out channels, pair of in-channel chunks = 16 in channels): it reads its
// [8][16][Kmax] slab of W with coalesced rows, masks W in place, and emits the 16 B rows of both packed
// layouts: forward blob_t[kc][n][8 ci] and dgrad blob_t'[kc = co chunk][n = ci][8 co].
// The per-tap geometry (live suffix, blob offsets) is recomputed per block from s(t) exactly as
// build_tap_table does on the host (stable order by suffix start, widest tap first).
// =====================================================================================================
namespace tsc {

struct PackTaps {
    short n_lo_f[TSC_MAX_TAPS];   // forward: first stored out channel of conv tap t (multiple of 16), -1 = dead
    short kc_lo_d[TSC_MAX_TAPS];  // dgrad  : first stored out-channel chunk of conv tap t' (even), -1 = dead
    int off_f[TSC_MAX_TAPS];      // blob offsets in 16 B rows
    int off_d[TSC_MAX_TAPS];
    int tot_f, tot_d;             // total rows of the forward / dgrad bank (packed_row's total_rows)
};

__device__ void pack_build_taps(const tsc_pack_layer& ly, PackTaps* pt, int* scratch /* [2 * TSC_MAX_TAPS + 1] */) {
    const int Kmax = ly.Kmax, Cout = ly.Cout;
    const int np_f = (ly.Cout + 15) & ~15, kc_f = ((ly.Cin + 15) & ~15) / 8;     // forward: N = out, K = in
    const int np_d = (ly.Cin + 15) & ~15, kc_d = ((ly.Cout + 15) & ~15) / 8;     // dgrad  : N = in,  K = out
    const int t = threadIdx.x;
    int* key_f = scratch;                      // sort key of tap t (n_lo; the widened first tap gets -1), dead = INT_MAX
    int* key_d = scratch + TSC_MAX_TAPS;
    int* first_key = scratch + 2 * TSC_MAX_TAPS;
    if (t == 0) *first_key = 0x7fffffff;
    __syncthreads();
    if (t < Kmax) {
        const int sf = ly.s_of_tap[t];
        const int nf = sf >= Cout ? -1 : (sf / 16) * 16;
        pt->n_lo_f[t] = (short)nf;
        if (nf >= 0) atomicMin(first_key, nf * 256 + t);        // widest tap, lowest index on ties = first in issue order
        const int sd = ly.s_of_tap[Kmax - 1 - t];
        const int kd = sd >= Cout ? -1 : (sd / 16) * 2;
        pt->kc_lo_d[t] = (short)kd;
        key_d[t] = kd < 0 ? 0x7fffffff : kd;
    }
    __syncthreads();
    const int first = *first_key & 255;
    if (t < Kmax) {
        const int nf = pt->n_lo_f[t];
        key_f[t] = nf < 0 ? 0x7fffffff : (t == first ? -1 : nf);
    }
    __syncthreads();
    if (t < Kmax) {
        // offsets = total size of the blobs that precede tap t in (key, index) order; iterations are independent
        const int kf = key_f[t], kd = key_d[t];
        int off = 0, offd = 0;
#pragma unroll 8
        for (int u = 0; u < Kmax; ++u) {
            const int ku = key_f[u], kud = key_d[u];
            const bool before_f = ku != 0x7fffffff && u != t && (ku < kf || (ku == kf && u < t));
            const bool before_d = kud != 0x7fffffff && u != t && (kud < kd || (kud == kd && u < t));
            off += before_f ? kc_f * (np_f - (ku < 0 ? 0 : ku)) : 0;
            offd += before_d ? (kc_d - kud) * np_d : 0;
        }
        pt->off_f[t] = kf == 0x7fffffff ? 0 : off;
        pt->off_d[t] = kd == 0x7fffffff ? 0 : offd;
        if (t == first) pt->n_lo_f[t] = 0;
    }
    if (t == 0) {
        // bank totals (the split layout's second stream starts at half of them)
        int tf = 0, td = 0;
        for (int u = 0; u < Kmax; ++u) {
            const int ku = key_f[u], kud = key_d[u];
            if (ku != 0x7fffffff) tf += kc_f * (np_f - (ku < 0 ? 0 : ku));
            if (kud != 0x7fffffff) td += (kc_d - kud) * np_d;
        }
        pt->tot_f = tf;
        pt->tot_d = td;
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(256) pack_multi_kernel(const __grid_constant__ tsc_pack_batch batch, int split) {
    extern __shared__ float wsm[];                    // [8][16][Kmax]
    __shared__ PackTaps pt;
    __shared__ int scratch[2 * TSC_MAX_TAPS + 1];
    __shared__ short s_tap[TSC_MAX_TAPS];             // s(t): the load loop indexes it per element -- from the constant bank
                                                      // a warp's 31 distinct taps serialise (the kernel's hot spot in round 1)
    const tsc_pack_layer& ly = batch.layer[blockIdx.y];
    const int Cin = ly.Cin, Cout = ly.Cout, Kmax = ly.Kmax;
    const int np_f = (Cout + 15) & ~15, cin_p = (Cin + 15) & ~15;
    const int n_cc = np_f / 8, n_kp = cin_p / 16;
    pdl_trigger();
    if ((int)blockIdx.x >= n_cc * n_kp) return;
    const int cc = blockIdx.x / n_kp, kp = blockIdx.x % n_kp;
    const int co0 = cc * 8, ci0 = kp * 16;
    if ((int)threadIdx.x < Kmax) s_tap[threadIdx.x] = ly.s_of_tap[threadIdx.x];
    pack_build_taps(ly, &pt, scratch);       // geometry only: overlaps the tail of the previous kernel (ends with a barrier)
    pdl_wait();
    // ---- load (and mask in place) ----
    float* W = ly.W;
    const int slab = 16 * Kmax;
    for (int e = threadIdx.x; e < 8 * slab; e += 256) {
        const int r = e / slab, q = e % slab;
        const int ci = ci0 + q / Kmax, t = q % Kmax, co = co0 + r;
        float v = 0.f;
        if (co < Cout && ci < Cin) {
            float* p = W + ((size_t)co * Cin + ci) * Kmax + t;
            const float w = *p;                          // the whole slab row is read: coalesced, no divergent loads
            if (co >= s_tap[t]) v = w;
            else if (ly.zero_masked && w != 0.f) *p = 0.f;   // masked taps stay zero once masked (their gradient is an
                                                             // exact zero): after the first step nothing is stored
        }
        wsm[e] = v;
    }
    __syncthreads();
    T* pf = reinterpret_cast<T*>(ly.packed_fwd);
    T* pd = reinterpret_cast<T*>(ly.packed_dgrad);
    const int kc_f = cin_p / 8, np_d = cin_p;
    // ---- forward rows: (t, k = 0..1 chunk of this pair, r = out channel) -> 8 in channels ----
    for (int e = threadIdx.x; e < Kmax * 16; e += 256) {
        const int t = e / 16, k = (e >> 3) & 1, r = e & 7;
        const int n_lo = pt.n_lo_f[t];
        const int co = co0 + r;
        if (n_lo < 0 || co < n_lo) continue;
        const int nt = np_f - n_lo;
        Row8<T> row;
#pragma unroll
        for (int j = 0; j < 8; ++j) row.v[j] = wsm[(r * 16 + k * 8 + j) * Kmax + t];
        row.store(pf + packed_row(pt.tot_f, pt.off_f[t], nt, kp * 2 + k, co - n_lo, split) * 8);
    }
    (void)kc_f;
    // ---- dgrad rows: (t', n = in channel of this pair) -> the 8 out channels of this chunk ----
    if (pd) {
        for (int e = threadIdx.x; e < Kmax * 16; e += 256) {
            const int t = e / 16, q = e & 15;
            const int kc_lo = pt.kc_lo_d[t];
            if (kc_lo < 0 || cc < kc_lo) continue;
            const int wt = Kmax - 1 - t;
            Row8<T> row;
#pragma unroll
            for (int j = 0; j < 8; ++j) row.v[j] = wsm[(j * 16 + q) * Kmax + wt];
            row.store(pd + packed_row(pt.tot_d, pt.off_d[t], np_d, cc - kc_lo, ci0 + q, split) * 8);
        }
    }
}

}  // namespace tsc

extern "C" int tsc_pack_weights_multi(int dtype, const tsc_pack_batch* batch, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(batch && batch->n >= 1 && batch->n <= TSC_PACK_MAX_LAYERS, "bad layer count");
    int max_blocks = 0, max_k = 0;
    for (int i = 0; i < batch->n; ++i) {
        const tsc_pack_layer& ly = batch->layer[i];
        TSC_REQUIRE(ly.W && ly.packed_fwd, "layer %d: NULL tensor", i);
        TSC_REQUIRE(ly.Kmax >= 1 && ly.Kmax <= TSC_MAX_TAPS && ly.Cin >= 1 && ly.Cin <= TSC_MAX_CHANNELS_WIDE && ly.Cout >= 1 &&
                        ly.Cout <= TSC_MAX_CHANNELS_WIDE, "layer %d: bad geometry", i);
        max_blocks = max(max_blocks, (pad16(ly.Cout) / 8) * (pad16(ly.Cin) / 16));
        max_k = max(max_k, ly.Kmax);
    }
    const int smem = 8 * 16 * max_k * (int)sizeof(float);
    dim3 grid(max_blocks, batch->n);
    cudaStream_t cs = (cudaStream_t)stream;
    if (dtype == TSC_BF16) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(pack_multi_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        launch_pdl(pack_multi_kernel<__nv_bfloat16>, grid, dim3(256), (size_t)smem, cs, *batch, packed_split() ? 1 : 0);
    } else if (dtype == TSC_F32) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(pack_multi_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        launch_pdl(pack_multi_kernel<float>, grid, dim3(256), (size_t)smem, cs, *batch, packed_split() ? 1 : 0);
    } else {
        TSC_REQUIRE(false, "bad dtype %d", dtype);
    }
    TSC_LAUNCH_CHECK();
    return 0;
}
