// tcgen05 engine, Gram style loss (placeholder until the kernel lands: reports "not implemented").
#include "tc_common.cuh"
namespace tsc {
int gram_fwd_tc(const float*, const float*, float*, float*, float*, int, int, int, cudaStream_t) {
    set_error("tcgen05 gram forward not implemented");
    return -1;
}
int gram_bwd_tc(const float*, const float*, const float*, const float*, float*, float*, int, int, int, cudaStream_t) {
    set_error("tcgen05 gram backward not implemented");
    return -1;
}
int read_clear_watchdog_gram(int* code) { *code = 0; return 0; }
}  // namespace tsc
