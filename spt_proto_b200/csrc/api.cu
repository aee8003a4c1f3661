// Library-wide state of libspt_b200: error string, launch counter, ABI version.
#include "common.cuh"

namespace spt {
thread_local char g_last_error[512] = "";
std::atomic<uint64_t> g_launch_count{0};
}  // namespace spt

extern "C" int spt_abi_version(void) { return SPT_ABI_VERSION; }
extern "C" const char *spt_last_error(void) { return spt::g_last_error; }
extern "C" uint64_t spt_launch_count(void) { return spt::g_launch_count.load(std::memory_order_relaxed); }
