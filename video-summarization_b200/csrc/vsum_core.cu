// Error text, ABI version and launch accounting for libvsum_b200.
#include "vsum_common.cuh"

#include <atomic>
#include <cstring>

namespace vsum {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

char *error_buffer() { return g_err; }

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace vsum

extern "C" int vsum_abi_version(void) { return 1; }
extern "C" const char *vsum_last_error(void) { return vsum::error_buffer(); }
extern "C" int64_t vsum_launch_count(void) { return vsum::g_launches.load(std::memory_order_relaxed); }
