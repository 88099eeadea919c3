// Error string, ABI version and launch counter of the rgbd_b200 C-ABI.
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

extern "C" void rgbd_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" void rgbd_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" const char *rgbd_last_error(void) { return g_err; }
extern "C" int rgbd_abi_version(void) { return 1; }
extern "C" int64_t rgbd_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}
extern "C" int rgbd_conv_desc_size(void) { return (int)sizeof(rgbd_conv_desc); }
