// Error string, version and launch accounting of libasac_b200.so.
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace asac {
static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
// Programmatic dependent launch of the step's kernels.  On by default since round 2: with griddepcontrol.wait moved
// behind the predecessor-independent part of the critic backward, the policy backward and the post pass the step is
// 7-10 us shorter; ASAC_PDL=0 in the environment (or asac_set_pdl(0)) launches every kernel in plain
// stream order.  Read at launch (and capture) time: a captured CUDA graph keeps the mode it was captured with.
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = getenv("ASAC_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
        g_pdl.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}
int set_pdl(int on) {
    const int before = pdl_enabled() ? 1 : 0;
    g_pdl.store(on ? 1 : 0, std::memory_order_relaxed);
    return before;
}
}  // namespace asac

extern "C" const char *asac_last_error(void) { return asac::g_error; }
extern "C" int asac_version(void) { return 100; }
extern "C" int asac_set_pdl(int on) { return asac::set_pdl(on); }
extern "C" int64_t asac_launch_count(void) { return asac::g_launches.load(std::memory_order_relaxed); }
extern "C" void asac_reset_launch_count(void) { asac::g_launches.store(0, std::memory_order_relaxed); }
