// Shared context of the C++ drop-in wrappers; see lsdb_host.h.
#include "lsdb_host.h"

#include <mutex>
#include <stdio.h>
#include <stdlib.h>

namespace lsdb_host {
static lsdb_ctx* g_ctx = 0;
static std::mutex g_mu;

[[noreturn]] void die(const char* what, int rc) {
    fprintf(stderr, "lsdb200: %s failed (code %d): %s\nlsdb200: there is no CPU fallback; an sm_100 (B200) device is required\n",
            what, rc, g_ctx ? lsdb_last_error(g_ctx) : "no context");
    abort();
}

lsdb_ctx* context() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx) {
        const char* d = getenv("LSDB_DEVICE");
        const int rc = lsdb_create(&g_ctx, d ? atoi(d) : 0, 0);
        if (rc != LSDB_OK) { g_ctx = 0; die("lsdb_create", rc); }
    }
    return g_ctx;
}

void shutdown() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ctx) lsdb_destroy(g_ctx);
    g_ctx = 0;
}
}  // namespace lsdb_host
