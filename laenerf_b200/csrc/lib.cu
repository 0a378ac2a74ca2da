// lib.cu -- error plumbing, version and launch accounting of liblaenerf_b200.so.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace lnrf {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("LNRF_PDL"); return e && e[0] == '1'; }();  // measured: no gain, see common.cuh
    return on;
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return LNRF_ERR_CUDA;
}

}  // namespace lnrf

extern "C" {
const char* lnrf_last_error(void) { return lnrf::t_error; }
int lnrf_version(void) { return 100; }  // 0.1.0
int lnrf_compiled_arch(void) { return 100; }  // sm_100a
uint64_t lnrf_launch_count(void) { return lnrf::g_launch_count.load(std::memory_order_relaxed); }
// struct sizes of this build: a binding checks its own mirror of the two descriptor structs against these before the first call
size_t lnrf_sizeof_render_desc(void) { return sizeof(lnrf_render_desc); }
size_t lnrf_sizeof_opt_tensor(void) { return sizeof(lnrf_opt_tensor); }
}
