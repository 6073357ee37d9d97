// tcgen05 engine, weight gradient (placeholder until the kernel lands: reports "not implemented").
#include "tc_common.cuh"
namespace tsc {
int wgrad_tc_splits(int, int, int, int, int) { return 1; }
int oswgrad_tc(const void*, const void*, int, float*, void*, int, int, int, int, int, const int*, cudaStream_t) {
    set_error("tcgen05 wgrad not implemented");
    return -1;
}
int read_clear_watchdog_wgrad(int* code) { *code = 0; return 0; }
}  // namespace tsc
