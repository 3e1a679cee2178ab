// Option table of the library (gd_options.cuh): environment read ONCE, then plain loads on the launch path.
#include "gd_common.cuh"
#include "gd_options.cuh"
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>

namespace gd {

static const char* const kOptNames[OPT_COUNT] = {
    "GD_CPS", "GD_NO_PWL", "GD_NO_CTAB", "GD_NO_RTAB", "GD_NO_VTAB", "GD_RTAB_N", "GD_CTAB_N", "GD_VTAB_N", "GD_VTAB_K",
    "GD_TILE", "GD_R", "GD_EB", "GD_FORCE_STREAMED", "GD_NPOLY", "GD_NO_VSKIP", "GD_NO_DIRECT", "GD_NO_LIGHT",
    "GD_FORCE_LIGHT", "GD_LTILE", "GD_LR", "GD_NO_GATED_HOST", "GD_GATE_CHUNKS", "GD_GATE_COPIES_FIRST", "GD_STILE",
    "GD_STHREADS", "GD_STREAM_LEGACY", "GD_SROWS", "GD_SSTAGES", "GD_SWARPS", "GD_NO_BWD_CTAB", "GD_NO_LEAN", "GD_LEAN_R",
    "GD_LEAN_G", "GD_LEAN_VTAB_N", "GD_LEAN_CTAB_N", "GD_LEAN_RTAB_N", "GD_LEAN_PARTS", "GD_NO_PDL", "GD_LAUNCH_BLOCKING"};

static std::atomic<long long> g_opt[OPT_COUNT];
static std::once_flag g_opt_once;
static std::atomic<long long> g_opt_epoch{0};

long long opt_epoch() { return g_opt_epoch.load(std::memory_order_relaxed); }

static void load_env() {
    for (int i = 0; i < OPT_COUNT; ++i) {
        const char* e = getenv(kOptNames[i]);
        g_opt[i].store(e ? atoll(e) : kOptUnset, std::memory_order_relaxed);
    }
    // tools that make kernel launches block the host (Nsight Compute marks its target with NV_NSIGHT_INJECTION_PORT_BASE /
    // NV_COMPUTE_PROFILER_PERFWORKS_DIR; other injectors use CUDA_INJECTION64_PATH): the gated host pipeline then queues
    // its copies before the kernel (gd_host.cu)
    if (getenv("CUDA_INJECTION64_PATH") || getenv("NV_NSIGHT_INJECTION_PORT_BASE") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") ||
        getenv("CUDA_LAUNCH_BLOCKING"))
        g_opt[OPT_LAUNCH_BLOCKING].store(1, std::memory_order_relaxed);
}

long long opt_get(Opt o) {
    std::call_once(g_opt_once, load_env);
    return g_opt[o].load(std::memory_order_relaxed);
}

}  // namespace gd

extern "C" int gd_set_option(const char* name, int64_t value, int32_t unset) {
    GD_CHECK_ARG(name != nullptr, "gd_set_option: name is NULL");
    std::call_once(gd::g_opt_once, gd::load_env);
    for (int i = 0; i < gd::OPT_COUNT; ++i)
        if (strcmp(name, gd::kOptNames[i]) == 0) {
            gd::g_opt[i].store(unset ? gd::kOptUnset : (long long)value, std::memory_order_relaxed);
            gd::g_opt_epoch.fetch_add(1, std::memory_order_relaxed);
            return GD_OK;
        }
    gd::set_error("gd_set_option: unknown option '%s'", name);
    return GD_ERR_INVALID;
}
