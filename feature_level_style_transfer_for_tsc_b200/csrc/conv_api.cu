// C-ABI entry points of the convolution family: engine dispatch (SIMT fp32 / tcgen05 bf16).
#include "common.cuh"
#include <stdlib.h>

namespace tsc {
int osconv_simt(int direction, const void* x, int dtype, const void* w, const float* bias, float* y, int B, int L, int Cin,
                int Cout, int Kmax, const int* s_of_tap, cudaStream_t cs);
int osconv_tc(int direction, const void* x, int dtype, const void* w, const void* plan, const float* bias, float* y,
              const tsc_conv_epilogue* epi, int B, int L, int Cin, int Cout, int Kmax, const int* s_of_tap, cudaStream_t cs);
int osconv2_tc(int direction, const void* x, int dtype, const void* w, const float* bias, float* y,
               const tsc_conv_epilogue* epi, int B, int L, int Cin, int Cout, int Kmax, const int* s_of_tap, cudaStream_t cs);
int read_clear_watchdog_conv2(int* code);
void set_conv2_timeline(long long* dev);
size_t osconv_plan_bytes(int direction, int Cin, int Cout, int Kmax, const int* s_of_tap);
int osconv_plan_build(int direction, int Cin, int Cout, int Kmax, const int* s_of_tap, void* host_plan);
int oswgrad_simt(const void* dy, const void* x, int dtype, float* dW, void* workspace, int accumulate, int B, int L, int Cin,
                 int Cout, int Kmax, const int* s_of_tap, cudaStream_t cs);
int oswgrad_tc(const void* dy, const void* x, int dtype, float* dW, void* workspace, int accumulate, int B, int L, int Cin,
               int Cout, int Kmax, const int* s_of_tap, cudaStream_t cs);
size_t wgrad_tc_workspace_bytes(int B, int L, int Cin, int Cout, int Kmax);
int wgrad_simt_splits(int B, int L, int Cin, int Cout, int Kmax);
int wgrad_tc_splits(int B, int L, int Cin, int Cout, int Kmax);
int read_clear_watchdog_conv(int* code);
int read_clear_watchdog_wgrad(int* code);
int read_clear_watchdog_gram(int* code);
void set_conv_timeline(long long* dev);
void set_wgrad_timeline(long long* dev);
void set_gram_timeline(long long* dev);
}  // namespace tsc

extern "C" {

size_t tsc_osconv_plan_bytes(int direction, int Cin, int Cout, int Kmax, const int* s_of_tap) {
    return tsc::osconv_plan_bytes(direction, Cin, Cout, Kmax, s_of_tap);
}

int tsc_osconv_plan_build(int direction, int Cin, int Cout, int Kmax, const int* s_of_tap, void* host_plan) {
    using namespace tsc;
    TSC_REQUIRE(host_plan && s_of_tap, "NULL argument");
    return osconv_plan_build(direction, Cin, Cout, Kmax, s_of_tap, host_plan);
}

int tsc_osconv(int engine, int direction, const void* x, int dtype, const void* w, const void* plan, const float* bias,
               float* y, const tsc_conv_epilogue* epilogue, int B, int L, int Cin, int Cout, int Kmax,
               const int* s_of_tap, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(x && w && (y || (epilogue && epilogue->affine_out && direction == TSC_DIR_FWD)), "NULL tensor");
    TSC_REQUIRE(B > 0 && L > 0, "bad shape B=%d L=%d", B, L);
    TSC_REQUIRE(dtype == TSC_F32 || dtype == TSC_BF16, "bad dtype %d", dtype);
    if (engine == TSC_ENGINE_SIMT) {
        TSC_REQUIRE(!epilogue || (!epilogue->stat_partial && !epilogue->red_partial && !epilogue->affine_out),
                    "fused epilogues exist on the tcgen05 engine only");
        return osconv_simt(direction, x, dtype, w, bias, y, B, L, Cin, Cout, Kmax, s_of_tap, (cudaStream_t)stream);
    }
    if (engine == TSC_ENGINE_TCGEN05) {
        (void)plan;               // (round-1 interface: the schedule now travels in the kernel parameter block)
        TSC_REQUIRE(Cin <= TSC_MAX_CHANNELS && Cout <= TSC_MAX_CHANNELS,
                    "tcgen05 engine: channel counts (%d,%d) exceed one TMEM accumulator tile (%d); use TSC_ENGINE_SIMT", Cin, Cout,
                    TSC_MAX_CHANNELS);
        return osconv2_tc(direction, x, dtype, w, bias, y, epilogue, B, L, Cin, Cout, Kmax, s_of_tap, (cudaStream_t)stream);
    }
    TSC_REQUIRE(false, "bad engine %d", engine);
}

size_t tsc_oswgrad_workspace_bytes(int engine, int B, int L, int Cin, int Cout, int Kmax) {
    using namespace tsc;
    if (engine == TSC_ENGINE_TCGEN05) return wgrad_tc_workspace_bytes(B, L, Cin, Cout, Kmax);
    return (size_t)wgrad_simt_splits(B, L, Cin, Cout, Kmax) * Kmax * pad16(Cout) * pad16(Cin) * sizeof(float);
}

int tsc_oswgrad(int engine, const void* dy, const void* x, int dtype, float* dW, void* workspace, int accumulate, int B,
                int L, int Cin, int Cout, int Kmax, const int* s_of_tap, tsc_stream_t stream) {
    using namespace tsc;
    TSC_REQUIRE(dy && x && dW && workspace, "NULL tensor");
    TSC_REQUIRE(B > 0 && L > 0, "bad shape B=%d L=%d", B, L);
    TSC_REQUIRE(dtype == TSC_F32 || dtype == TSC_BF16, "bad dtype %d", dtype);
    TSC_REQUIRE(Kmax >= 1 && Kmax <= TSC_MAX_TAPS && s_of_tap, "bad kernel bank");
    const int cmax = engine == TSC_ENGINE_SIMT ? TSC_MAX_CHANNELS_WIDE : TSC_MAX_CHANNELS;
    TSC_REQUIRE(Cin >= 1 && Cin <= cmax && Cout >= 1 && Cout <= cmax, "channel counts (%d,%d) outside [1,%d] of this engine", Cin, Cout, cmax);
    if (engine == TSC_ENGINE_SIMT)
        return oswgrad_simt(dy, x, dtype, dW, workspace, accumulate, B, L, Cin, Cout, Kmax, s_of_tap, (cudaStream_t)stream);
    if (engine == TSC_ENGINE_TCGEN05)
        return oswgrad_tc(dy, x, dtype, dW, workspace, accumulate, B, L, Cin, Cout, Kmax, s_of_tap, (cudaStream_t)stream);
    TSC_REQUIRE(false, "bad engine %d", engine);
}

int tsc_debug_set_timeline(void* dev_buf) {
    tsc::set_conv_timeline((long long*)dev_buf);
    tsc::set_conv2_timeline((long long*)dev_buf);
    tsc::set_wgrad_timeline((long long*)dev_buf);
    tsc::set_gram_timeline((long long*)dev_buf);
    return 0;
}

int tsc_debug_read_and_clear_watchdog(int* host_code) {
    using namespace tsc;
    int a = 0, b = 0, c = 0, a2 = 0, r;
    if ((r = read_clear_watchdog_conv(&a)) != 0) return r;
    if ((r = read_clear_watchdog_conv2(&a2)) != 0) return r;
    if (!a) a = a2;
    if ((r = read_clear_watchdog_wgrad(&b)) != 0) return r;
    if ((r = read_clear_watchdog_gram(&c)) != 0) return r;
    *host_code = a ? a : (b ? b : c);
    return 0;
}

}  // extern "C"
