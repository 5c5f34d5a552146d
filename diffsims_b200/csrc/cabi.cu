// C-ABI housekeeping: version, the per-thread error string, and the process-wide tuning options.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace ds {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Options: the environment is read ONCE (first use, under std::call_once); afterwards a launch path only does
// relaxed atomic loads, so concurrent launches on different streams never race on getenv or on a static.
// -1 means "not set: use the measured default".
struct OptDef {
    const char *name, *env;
};
static const OptDef g_defs[OPT_COUNT] = {
    {"render_pipe", "DS_RENDER_PIPE"},
    {"render_group", "DS_RENDER_GROUP"},
    {"render_fronts", "DS_RENDER_FRONTS"},
    {"render_pipe_maxcap", "DS_RENDER_PIPE_MAXCAP"},
    {"render_nostage", "DS_RENDER_NOSTAGE"},
    {"render_mma", "DS_RENDER_MMA"},
    {"render_mma_min", "DS_RENDER_MMA_MIN"},
    {"render_mma_tmpl_min", "DS_RENDER_MMA_TMPL_MIN"},
    {"render_umma", "DS_RENDER_UMMA"},
    {"render_umma_window", "DS_RENDER_UMMA_WINDOW"},
    {"render_zero_tma", "DS_RENDER_ZERO_TMA"},
    {"render_umma_team", "DS_RENDER_UMMA_TEAM"},
    {"render_rows", "DS_RENDER_ROWS"},
    {"render_rows_stages", "DS_RENDER_ROWS_STAGES"},
    {"sim_lines", "DS_SIM_LINES"},
    {"sim_split", "DS_SIM_SPLIT"},
    {"sim_cta", "DS_SIM_CTA"},
    {"sim_stash", "DS_SIM_STASH"},
};
static std::atomic<int> g_opt[OPT_COUNT];
static std::once_flag g_opt_once;
static void load_env() {
    for (int i = 0; i < OPT_COUNT; ++i) {
        const char *e = getenv(g_defs[i].env);
        g_opt[i].store(e ? atoi(e) : -1, std::memory_order_relaxed);
    }
}
int option(Opt o) {
    std::call_once(g_opt_once, load_env);
    return g_opt[o].load(std::memory_order_relaxed);
}

int num_sms() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}
}  // namespace ds

extern "C" int ds_abi_version(void) { return DS_ABI_VERSION; }
extern "C" const char *ds_last_error(void) { return ds::g_err; }

extern "C" int ds_set_option(const char *name, int32_t value) {
    using namespace ds;
    DS_REQUIRE(name != nullptr, "ds_set_option: null name");
    std::call_once(g_opt_once, load_env);
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(name, g_defs[i].name) == 0) {
            g_opt[i].store(value, std::memory_order_relaxed);
            return 0;
        }
    set_error("ds_set_option: unknown option '%s'", name);
    return -1;
}

extern "C" int ds_get_option(const char *name, int32_t *value) {
    using namespace ds;
    DS_REQUIRE(name != nullptr && value != nullptr, "ds_get_option: null argument");
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(name, g_defs[i].name) == 0) {
            *value = option((Opt)i);
            return 0;
        }
    set_error("ds_get_option: unknown option '%s'", name);
    return -1;
}
