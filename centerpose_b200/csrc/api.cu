// Library-wide state of the centerpose_b200 C ABI (error text, launch counter, version).
#include "common.cuh"

namespace cpb {
thread_local char g_err[512] = {0};
std::atomic<unsigned long long> g_launches{0};
}  // namespace cpb

extern "C" int cpb200_version(void) { return 100; }
extern "C" const char *cpb200_last_error(void) { return cpb::g_err; }
extern "C" unsigned long long cpb200_launch_count(void) { return cpb::g_launches.load(); }
