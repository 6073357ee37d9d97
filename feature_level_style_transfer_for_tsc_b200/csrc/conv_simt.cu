// SIMT (CUDA-core, fp32 arithmetic) engine for the omni-scale convolution: forward / dgrad as one
// generic implicit GEMM over the c8 layout, and the live-tap weight gradient.
// This is the "fp32, <= 1e-5" precision mode; with bf16 operands it performs exactly the arithmetic
// of the tcgen05 engine (bf16 products, fp32 accumulation) and serves as its bit-faithful checker.
// Replaces ConstantPad1d + Conv1d and their backward (OS_CNN/OS_CNN.py:70-71); formulas SURVEY A1.
#include "common.cuh"

namespace tsc {

static constexpr int CT_POS = 64;      // positions per CTA tile
static constexpr int CT_N = 64;        // output channels per CTA tile
static constexpr int CT_TG = 4;        // taps staged per shared-memory round

// y[b,n,l] = bias[n] + sum_t sum_kc sum_j x[b, kc*8+j, l+t-pad_left] * blob_t[kc-kc_lo][n-n_lo][j]
template <typename T>
__global__ void __launch_bounds__(256) osconv_simt_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                          const float* __restrict__ bias, int nbias,
                                                          float* __restrict__ y, int B, int L,
                                                          const __grid_constant__ TapTable tt, int split) {
    __shared__ float xs[8][CT_POS + TSC_MAX_TAPS];
    __shared__ __align__(16) float ws[CT_TG][8][CT_N];
    const int ltiles = (L + CT_POS - 1) / CT_POS;
    const int b = blockIdx.x / ltiles, l0 = (blockIdx.x % ltiles) * CT_POS, n0 = blockIdx.y * CT_N;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int np = tt.np, kc = tt.kc;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;

    const int nrows = CT_POS + tt.taps - 1;
    for (int kcI = 0; kcI < kc; ++kcI) {
        __syncthreads();
        for (int r = tid; r < nrows; r += 256) {
            const int l = l0 + r - tt.pad_left;
            Row8<T> v;
            if (l >= 0 && l < L) {
                v.load(x + (((long long)b * kc + kcI) * L + l) * 8);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v.v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) xs[j][r] = v.v[j];
        }
        for (int g0 = 0; g0 < tt.n_order; g0 += CT_TG) {
            if (g0 > 0) __syncthreads();
            {
                const int tgi = tid >> 6, n = tid & 63, oi = g0 + tgi;
                Row8<T> v;
#pragma unroll
                for (int j = 0; j < 8; ++j) v.v[j] = 0.f;
                if (oi < tt.n_order) {
                    const int t = tt.order[oi];
                    const int n_lo = tt.n_lo[t], kc_lo = tt.kc_lo[t], nn = n0 + n;
                    if (kcI >= kc_lo && nn >= n_lo && nn < np)
                        v.load(w + packed_row(tt.total_rows, tt.w_off[t], np - n_lo, kcI - kc_lo, nn - n_lo, split) * 8);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) ws[tgi][j][n] = v.v[j];
            }
            __syncthreads();
            const int ng = min(CT_TG, tt.n_order - g0);
            for (int tgi = 0; tgi < ng; ++tgi) {
                const int t = tt.order[g0 + tgi];
                if (kcI < tt.kc_lo[t] || n0 + CT_N <= tt.n_lo[t]) continue;     // CTA-uniform skip
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float a[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = xs[j][tx + 16 * i + t];
                    const float4 b4 = *reinterpret_cast<const float4*>(&ws[tgi][j][ty * 4]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[i][0] = fmaf(a[i], b4.x, acc[i][0]);
                        acc[i][1] = fmaf(a[i], b4.y, acc[i][1]);
                        acc[i][2] = fmaf(a[i], b4.z, acc[i][2]);
                        acc[i][3] = fmaf(a[i], b4.w, acc[i][3]);
                    }
                }
            }
        }
    }
    const int n = n0 + ty * 4;
    if (n < np) {
        float bv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) bv[k] = (bias && n + k < nbias) ? __ldg(bias + n + k) : 0.f;
        const int npc = np / 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int l = l0 + tx + 16 * i;
            if (l < L) {
                float* dst = y + (((long long)b * npc + (n >> 3)) * L + l) * 8 + (n & 7);
                *reinterpret_cast<float4*>(dst) =
                    make_float4(acc[i][0] + bv[0], acc[i][1] + bv[1], acc[i][2] + bv[2], acc[i][3] + bv[3]);
            }
        }
    }
}

// ---- weight gradient ----------------------------------------------------------------------------
// CTA: 32 out channels x 8 in channels (one chunk) x 8 taps, over its share of the (b, l) positions.
// part[S][taps][np][kcp]
static constexpr int WG_CO = 32;
static constexpr int WG_TAPS = 8;

template <typename T>
__global__ void __launch_bounds__(256) oswgrad_simt_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                           float* __restrict__ part, int B, int L, int kc_x, int np,
                                                           int taps, int pad_left, int S,
                                                           const __grid_constant__ STable st) {
    __shared__ float ds[WG_CO][CT_POS + 1];
    __shared__ float xs[8][CT_POS + WG_TAPS + 1];          // 73 columns: odd stride, conflict free
    const int co0 = (blockIdx.x / kc_x) * WG_CO, cix = blockIdx.x % kc_x;
    const int t0 = blockIdx.y * WG_TAPS, sp = blockIdx.z;
    // CTA-uniform skip: no (co, t) of this tile is live (live <=> co >= s(t))
    {
        int smin = 0x7fffffff;
        for (int tt = 0; tt < WG_TAPS && t0 + tt < taps; ++tt) smin = min(smin, st.s[t0 + tt]);
        if (co0 + WG_CO - 1 < smin) return;
    }
    const int tid = threadIdx.x, co_i = tid >> 3, ci_j = tid & 7;
    const int npc = np / 8, lt = (L + CT_POS - 1) / CT_POS;
    const long long nblk = (long long)B * lt;
    const long long blk0 = nblk * sp / S, blk1 = nblk * (sp + 1) / S;
    float acc[WG_TAPS];
#pragma unroll
    for (int k = 0; k < WG_TAPS; ++k) acc[k] = 0.f;
    for (long long blk = blk0; blk < blk1; ++blk) {
        const int b = (int)(blk / lt), l0 = (int)(blk % lt) * CT_POS;
        __syncthreads();
        {
            const int c4 = tid >> 6, l = tid & 63;
            Row8<T> v;
            if (l0 + l < L && co0 / 8 + c4 < npc) {
                v.load(dy + (((long long)b * npc + co0 / 8 + c4) * L + l0 + l) * 8);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v.v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) ds[c4 * 8 + j][l] = v.v[j];
        }
        if (tid < CT_POS + WG_TAPS - 1) {
            const int l = l0 + tid + t0 - pad_left;
            Row8<T> v;
            if (l >= 0 && l < L) {
                v.load(x + (((long long)b * kc_x + cix) * L + l) * 8);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v.v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) xs[j][tid] = v.v[j];
        }
        __syncthreads();
#pragma unroll 2
        for (int lb = 0; lb < CT_POS; lb += 8) {
            float wv[16];
#pragma unroll
            for (int q = 0; q < 15; ++q) wv[q] = xs[ci_j][lb + q];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float d = ds[co_i][lb + u];
#pragma unroll
                for (int k = 0; k < WG_TAPS; ++k) acc[k] = fmaf(d, wv[u + k], acc[k]);
            }
        }
    }
    const int kcp = kc_x * 8;
#pragma unroll
    for (int k = 0; k < WG_TAPS; ++k) {
        const int t = t0 + k;
        if (t < taps && co0 + co_i < np)
            part[(((long long)sp * taps + t) * np + co0 + co_i) * kcp + cix * 8 + ci_j] = acc[k];
    }
}

int wgrad_simt_splits(int B, int L, int Cin, int Cout, int Kmax) {
    const int base = cdiv(pad16(Cout), WG_CO) * (pad16(Cin) / 8) * cdiv(Kmax, WG_TAPS);
    const int nblk = B * cdiv(L, CT_POS);
    int s = cdiv(148 * 6, base);
    if (s > nblk) s = nblk;
    if (s > 32) s = 32;
    if (s < 1) s = 1;
    return s;
}

int osconv_simt(int direction, const void* x, int dtype, const void* w, const float* bias, float* y, int B, int L, int Cin,
                int Cout, int Kmax, const int* s_of_tap, cudaStream_t cs) {
    TapTable tt;
    if (build_tap_table(direction, Cin, Cout, Kmax, s_of_tap, &tt) != 0) return -1;
    dim3 grid(B * cdiv(L, CT_POS), cdiv(tt.np, CT_N));
    const int nbias = direction == TSC_DIR_FWD ? Cout : 0;
    const float* bp = direction == TSC_DIR_FWD ? bias : nullptr;
    if (dtype == TSC_BF16)
        osconv_simt_kernel<__nv_bfloat16><<<grid, 256, 0, cs>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bp, nbias, y, B, L, tt, packed_split() ? 1 : 0);
    else
        osconv_simt_kernel<float><<<grid, 256, 0, cs>>>((const float*)x, (const float*)w, bp, nbias, y, B, L, tt, packed_split() ? 1 : 0);
    TSC_LAUNCH_CHECK();
    return 0;
}

int oswgrad_simt(const void* dy, const void* x, int dtype, float* dW, void* workspace, int accumulate, int B, int L, int Cin,
                 int Cout, int Kmax, const int* s_of_tap, cudaStream_t cs) {
    const int np = pad16(Cout), kc_x = pad16(Cin) / 8;
    const int S = wgrad_simt_splits(B, L, Cin, Cout, Kmax);
    STable st;
    fill_stable(&st, s_of_tap, Kmax);
    dim3 grid(cdiv(np, WG_CO) * kc_x, cdiv(Kmax, WG_TAPS), S);
    const int pad_left = (Kmax - 1) / 2;
    float* part = (float*)workspace;
    if (dtype == TSC_BF16)
        oswgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, cs>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, part, B, L, kc_x, np, Kmax, pad_left, S, st);
    else
        oswgrad_simt_kernel<float><<<grid, 256, 0, cs>>>((const float*)dy, (const float*)x, part, B, L, kc_x, np, Kmax, pad_left, S, st);
    TSC_LAUNCH_CHECK();
    return launch_wgrad_reduce(part, dW, S, Cin, Cout, Kmax, np, kc_x * 8, s_of_tap, accumulate, cs);
}

}  // namespace tsc
