// Library-wide state of libmzb200.so: last-error string (per thread) and the launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace mzb {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int pdl_level()
{
    static const int lvl = [] { const char *e = getenv("MZB_PDL"); return e ? atoi(e) : 0; }();
    return lvl;
}

}  // namespace mzb

extern "C" {
int mzb_version(void) { return MZB_VERSION; }
const char *mzb_last_error(void) { return mzb::g_err; }
uint64_t mzb_launch_count(void) { return mzb::g_launches.load(std::memory_order_relaxed); }
size_t mzb_sizeof(int which) { return which == 0 ? sizeof(mz_tree_args) : (which == 1 ? sizeof(mz_op) : (which == 2 ? sizeof(rb_ring) : 0)); }
}
