// Library-level entry points: version, thread-local error string, launch counter.
#include "common.cuh"
#include "../../include/ubs_gnn.h"
#include <stdarg.h>

namespace ubs {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace ubs

extern "C" UBS_API int ubs_version(void) { return UBS_GNN_VERSION; }
extern "C" UBS_API const char* ubs_last_error(void) { return ubs::g_err; }
extern "C" UBS_API int64_t ubs_launch_count(void) { return (int64_t)ubs::g_launches.load(); }
extern "C" UBS_API void ubs_reset_launch_count(void) { ubs::g_launches.store(0); }
extern "C" UBS_API void ubs_add_launch_count(int64_t n) { ubs::g_launches.fetch_add((long long)n); }
