// libtsc_b200: error plumbing, tap tables, version.  See include/tsc_b200.h.
#include "common.cuh"
#include <string.h>
#include <algorithm>
#include <stdlib.h>

namespace tsc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool packed_split() {
    // the CTA-pair (cta_group::2) convolution and its split packed layout: measured SLOWER than the one-CTA kernel on this
    // workload's narrow MMAs (profiles/README.md, round 2), so it is an opt-in experiment: TSC_CONV_PAIR=1
    static const bool on = [] { const char* e = getenv("TSC_CONV_PAIR"); return e && e[0] == '1'; }();
    return on;
}

bool pdl_enabled() {
    // off by default: measured on the cfg2 step it costs 2-3 % (early-resident CTAs of the next kernel take the SM slot
    // that the other branch's stream would have used) -- profiles/README.md; TSC_PDL=1 switches it on
    static const bool on = [] { const char* e = getenv("TSC_PDL"); return e && e[0] == '1'; }();
    return on;
}

// Per-tap geometry of the implicit GEMM (see TapTable in common.cuh).
// FWD  : contraction over input channels, all of them for every live tap; output channels are the
//        live suffix [s(t), Cout), rounded down to the MMA N granularity (16).
// DGRAD: the transposed convolution; conv tap t' uses weight tap Kmax-1-t'; the contraction runs over
//        the live *output* channels (a suffix), rounded down to a whole MMA K step (16 channels).
int build_tap_table(int direction, int Cin, int Cout, int Kmax, const int* s_of_tap, TapTable* tt) {
    TSC_REQUIRE(direction == TSC_DIR_FWD || direction == TSC_DIR_DGRAD, "bad direction %d", direction);
    TSC_REQUIRE(Kmax >= 1 && Kmax <= TSC_MAX_TAPS, "Kmax=%d outside [1,%d]", Kmax, TSC_MAX_TAPS);
    TSC_REQUIRE(Cin >= 1 && Cin <= TSC_MAX_CHANNELS_WIDE && Cout >= 1 && Cout <= TSC_MAX_CHANNELS_WIDE,
                "channel counts (%d,%d) outside [1,%d]", Cin, Cout, TSC_MAX_CHANNELS_WIDE);
    TSC_REQUIRE(s_of_tap != nullptr, "s_of_tap is NULL");
    memset(tt, 0, sizeof(*tt));
    tt->taps = Kmax;
    const bool fwd = direction == TSC_DIR_FWD;
    tt->pad_left = fwd ? (Kmax - 1) / 2 : Kmax / 2;
    tt->kc = fwd ? pad16(Cin) / 8 : pad16(Cout) / 8;
    tt->np = fwd ? pad16(Cout) : pad16(Cin);
    int n = 0;
    for (int t = 0; t < Kmax; ++t) {
        const int wt = fwd ? t : Kmax - 1 - t;     // weight tap used by conv tap t
        const int s = s_of_tap[wt];
        TSC_REQUIRE(s >= 0, "s_of_tap[%d]=%d negative", wt, s);
        if (s >= Cout) continue;                   // dead tap
        tt->order[n] = (short)t;
        tt->n_lo[t] = fwd ? (short)((s / 16) * 16) : 0;
        tt->kc_lo[t] = fwd ? 0 : (short)((s / 16) * 2);
        ++n;
    }
    TSC_REQUIRE(n > 0, "kernel bank has no live tap");
    tt->n_order = n;
    // widest tap first: the first MMA of a tile overwrites (does not accumulate into) the accumulator
    // and must therefore cover every output column.
    std::stable_sort(tt->order, tt->order + n, [&](short a, short b) {
        return fwd ? tt->n_lo[a] < tt->n_lo[b] : tt->kc_lo[a] < tt->kc_lo[b];
    });
    if (fwd) tt->n_lo[tt->order[0]] = 0;
    int rows = 0;
    for (int i = 0; i < n; ++i) {
        const int t = tt->order[i];
        tt->w_off[t] = rows;
        rows += (tt->kc - tt->kc_lo[t]) * (tt->np - tt->n_lo[t]);
    }
    tt->total_rows = rows;
    return 0;
}

}  // namespace tsc

extern "C" {

int tsc_version(void) { return TSC_VERSION; }
const char* tsc_last_error(void) { return tsc::g_err; }
int tsc_pad_channels(int C) { return tsc::pad16(C); }

int tsc_device_supports_tcgen05(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10;
}

size_t tsc_packed_weight_bytes(int direction, int dtype, int Cin, int Cout, int Kmax, const int* s_of_tap) {
    tsc::TapTable tt;
    if (tsc::build_tap_table(direction, Cin, Cout, Kmax, s_of_tap, &tt) != 0) return 0;
    return (size_t)tt.total_rows * 8 * (dtype == TSC_BF16 ? 2 : 4);
}

}  // extern "C"
