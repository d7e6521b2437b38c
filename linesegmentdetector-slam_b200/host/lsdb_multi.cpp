// One process, several GPUs (SURVEY.md §8b `lsdb_create(ctx, device_ids[])`, §8e "one host thread / stream per GPU"): a batch
// of maps is cut into contiguous, balanced shards (the first n % k devices get one extra map — the split shard.py uses for
// the one-process-per-GPU launch), every shard runs on its own device through its own context from its own host thread, and
// the tables land in the caller's arrays in batch order.  Maps are independent units: no collective, no exchange.
// Part of liblsdb200.so; calls the C ABI only.
#include "../../include/lsdb200.h"

#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>

struct lsdb_multi {
    std::vector<int> devices;
    std::vector<lsdb_ctx*> ctx;
    std::string err;
};

extern "C" int lsdb_multi_create(lsdb_multi** out, const int* device_ids, int n_devices) {
    if (!out || !device_ids || n_devices <= 0) return LSDB_ERR_ARG;
    *out = 0;
    lsdb_multi* m = new lsdb_multi();
    for (int d = 0; d < n_devices; d++) {
        lsdb_ctx* c = 0;
        const int rc = lsdb_create(&c, device_ids[d], 0);   // a private stream per context: the same device may be listed twice
        if (rc != LSDB_OK) {
            for (size_t k = 0; k < m->ctx.size(); k++) lsdb_destroy(m->ctx[k]);
            delete m;
            return rc;
        }
        m->devices.push_back(device_ids[d]);
        m->ctx.push_back(c);
    }
    *out = m;
    return LSDB_OK;
}

extern "C" void lsdb_multi_destroy(lsdb_multi* m) {
    if (!m) return;
    for (size_t k = 0; k < m->ctx.size(); k++) lsdb_destroy(m->ctx[k]);
    delete m;
}

extern "C" int lsdb_multi_devices(const lsdb_multi* m) { return m ? (int)m->ctx.size() : 0; }
extern "C" const char* lsdb_multi_last_error(const lsdb_multi* m) { return m ? m->err.c_str() : "null handle"; }

// the shard of device d: [first, first + count)
extern "C" void lsdb_multi_shard(int n_items, int d, int n_devices, int* first, int* count) {
    const int base = n_items / n_devices, extra = n_items % n_devices;
    *first = d * base + (d < extra ? d : extra);
    *count = base + (d < extra ? 1 : 0);
}

extern "C" int lsdb_multi_lsd(lsdb_multi* m, int n_maps, const uint8_t* const* maps, const int* cols, const int* rows,
                              const lsdb_lsd_params* prm, int max_lines, int* counts, lsdb_line* lines, lsdb_rect* rects) {
    if (!m || n_maps < 0 || (n_maps > 0 && (!maps || !cols || !rows)) || !prm || max_lines <= 0 || !counts) return LSDB_ERR_ARG;
    const int k = (int)m->ctx.size();
    std::vector<int> rc(k, LSDB_OK);
    std::vector<std::string> msg(k);
    std::vector<std::thread> th;
    for (int d = 0; d < k; d++) {
        th.emplace_back([&, d]() {
            int first, cnt;
            lsdb_multi_shard(n_maps, d, k, &first, &cnt);
            if (cnt == 0) return;
            lsdb_ctx* c = m->ctx[d];
            lsdb_batch* b = 0;
            int r = lsdb_batch_create(c, cnt, cols + first, rows + first, prm, max_lines, &b);
            if (r == LSDB_OK) r = lsdb_batch_upload(b, maps + first);
            if (r == LSDB_OK) r = lsdb_batch_run(b);
            if (r == LSDB_OK) r = lsdb_batch_download(b, counts + first, lines ? lines + (size_t)first * max_lines : 0,
                                                     rects ? rects + (size_t)first * max_lines : 0);
            if (r != LSDB_OK) msg[d] = lsdb_last_error(c);
            if (b) lsdb_batch_destroy(b);
            rc[d] = r;
        });
    }
    for (int d = 0; d < k; d++) th[d].join();
    for (int d = 0; d < k; d++)
        if (rc[d] != LSDB_OK) { m->err = "device " + std::to_string(m->devices[d]) + ": " + msg[d]; return rc[d]; }
    return LSDB_OK;
}
