// abi.cu -- process-wide pieces of the C ABI: error string, version, launch counter.
#include "common.cuh"
#include <cstdarg>
#include <cstdio>

namespace rn {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace rn

extern "C" const char* rn_last_error_string(void) { return rn::g_err; }
extern "C" int rn_abi_version(void) { return 1; }
extern "C" uint64_t rn_launch_count(void) { return rn::g_launch_count.load(); }
// a replayed CUDA graph re-launches the kernels that were counted once at capture time
extern "C" void rn_note_graph_replay(uint64_t kernels_in_graph) { rn::g_launch_count.fetch_add(kernels_in_graph); }
