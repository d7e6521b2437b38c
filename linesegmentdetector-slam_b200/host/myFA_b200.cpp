// Drop-in body for myfa::FeatureAssociation (reference LSD/myFA.h:83, LSD/myFA.cpp:13-184).
//
// The scoring half — the pair filter :29-41, the pthread pool dispatch :22-63 and, per task,
// thread_ScanToMapMatch :186-272 (NormalizedLineDirection :274-305, rotateScanIm :307-355, CalcScore
// :357-396) — and the reduction that follows it (keep score < 3, order by score, best hypothesis /
// 1/score^2 weighted mean, :65-171) run on the device through lsdb_fa_estimate_frames.  The branch logic
// of the hidden-Markov chain (first frame / tracking / lost) stays here, and the reference's own
// myfa::ukf (:404-536) is called, not re-implemented.
//
// Build: compile against the reference's myFA.h; the reference's myFA.cpp stays in the build for ukf()
// with its FeatureAssociation renamed away (-DFeatureAssociation=FeatureAssociation_cpu on that TU, see
// INTEGRATION.md).  Differences a caller can observe:
//   * hypotheses are ordered (scan line, map line, pairing) instead of by thread arrival, so equal-score
//     ties sort deterministically (the reference's order is timing-dependent);
//   * structScore::rotateScanImPoint is NULL (the reference stores a pointer it has already freed, :256-258);
//   * the device copy of mapCache / mapLinesInfo is cached across frames, keyed by the Mat's data pointer,
//     its size and a hash of the map lines (the reference's drivers build both once per map).
#include <myFA.h>

#include <algorithm>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "lsdb_host.h"

static_assert(sizeof(structLinesInfo) == sizeof(lsdb_line), "structLinesInfo layout (LSD/baseFunc.h:33-44)");

namespace {

struct MapKey {
    const void* data; int rows, cols, nLines; unsigned long long hash, sample;
    bool operator==(const MapKey& o) const {
        return data == o.data && rows == o.rows && cols == o.cols && nLines == o.nLines && hash == o.hash && sample == o.sample;
    }
};
MapKey g_key = {0, 0, 0, 0, 0, 0};
lsdb_fa_map* g_map = 0;

unsigned long long fnv1a(const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

// a fingerprint of the cache's CONTENT (4096 cells spread over the matrix): a map rewritten in place, or a new Mat that got the
// old one's address, must not be served from the stale device copy
unsigned long long sample_cells(const cv::Mat& m) {
    unsigned long long h = 1469598103934665603ull;
    const size_t n = (size_t)m.rows * (size_t)m.cols, step = n / 4096 + 1;
    for (size_t i = 0; i < n; i += step) {
        const double v = m.ptr<double>((int)(i / m.cols))[i % m.cols];
        unsigned long long b; memcpy(&b, &v, 8);
        h ^= b; h *= 1099511628211ull;
    }
    return h;
}

lsdb_fa_map* device_map(lsdb_ctx* ctx, myfa::structFAInput* in) {
    const int nLines = (int)in->mapLinesInfo.size();
    MapKey k = {in->mapCache.data, in->mapCache.size[0], in->mapCache.size[1], nLines,
                nLines ? fnv1a(&in->mapLinesInfo[0], sizeof(structLinesInfo) * (size_t)nLines) : 0ull, sample_cells(in->mapCache)};
    if (g_map && k == g_key && !getenv("LSDB_FA_NO_CACHE")) return g_map;
    if (g_map) lsdb_fa_map_destroy(g_map);
    g_map = 0;
    const int rows = k.rows, cols = k.cols;
    std::vector<double> packed;
    const double* cache = in->mapCache.ptr<double>(0);
    if (in->mapCache.step != (size_t)cols * sizeof(double)) {   // padded rows: pack
        packed.resize((size_t)rows * cols);
        for (int y = 0; y < rows; y++) memcpy(&packed[(size_t)y * cols], in->mapCache.ptr<double>(y), sizeof(double) * (size_t)cols);
        cache = packed.data();
    }
    const int rc = lsdb_fa_map_create(ctx, cache, cols, rows, nLines ? (const lsdb_line*)&in->mapLinesInfo[0] : 0, nLines, &g_map);
    if (rc != LSDB_OK) lsdb_host::die("lsdb_fa_map_create", rc);
    g_key = k;
    return g_map;
}

myfa::structFAOutput lost_track() {   // "no match: start a new chain", LSD/myFA.cpp:70-90 and :131-151
    myfa::structFAOutput o;
    const double px[9] = {-1, -1, 0, 0, 0, 0, 0, 0, 0};
    const double pd[9] = {100, 100, 100, 1, 1, 1, 0.1, 0.1, 0.1};
    for (int i = 0; i < 9; i++) {
        o.kalman_x(i) = px[i];
        for (int j = 0; j < 9; j++) o.kalman_P(i, j) = i == j ? pd[i] : 0.0;
    }
    return o;
}

}  // namespace

namespace myfa {

structFAOutput FeatureAssociation(structFAInput* FAInput) {
    lsdb_ctx* ctx = lsdb_host::context();
    lsdb_fa_map* m = device_map(ctx, FAInput);

    const int nScan = (int)FAInput->scanLinesInfo.size(), nPts = (int)FAInput->scanImPoint.size();
    const int nMap = (int)FAInput->mapLinesInfo.size();
    const int lineOff[2] = {0, nScan}, ptOff[2] = {0, nPts};
    std::vector<double> pts(2 * (size_t)(nPts > 0 ? nPts : 1));
    for (int i = 0; i < nPts; i++) { pts[2 * i] = FAInput->scanImPoint[i].x; pts[2 * i + 1] = FAInput->scanImPoint[i].y; }
    const double lidar[2] = {FAInput->lidarPose.x, FAInput->lidarPose.y};
    const double last[3] = {FAInput->lastPose.x, FAInput->lastPose.y, FAInput->lastPose.ang};
    (void)nMap;
    // scoring AND the reduction that follows it (keep score < 3, order by score, best / weighted mean) run on the device;
    // one record comes back
    lsdb_fa_estimate est;
    const int rc = lsdb_fa_estimate_frames(ctx, m, 1, nScan ? (const lsdb_line*)&FAInput->scanLinesInfo[0] : 0, lineOff, pts.data(), ptOff,
                                           lidar, last, &est);
    if (rc != LSDB_OK) lsdb_host::die("lsdb_fa_estimate_frames", rc);
    if (est.n_kept == 0) return lost_track();

    structFAOutput out;
    if (fabs(FAInput->lastPose.x + 1) < 0.0001) {   // first frame of a chain: take the best hypothesis (:100-110)
        out.kalman_x = FAInput->kalman_x;
        out.kalman_P = FAInput->kalman_P;
        out.kalman_x(0) = est.best_x;
        out.kalman_x(1) = est.best_y;
        out.kalman_x(2) = est.best_ang;
        printf("Score:%lf\n", est.best_score);
        return out;
    }

    // later frames: 1/score^2 weighted mean of every kept hypothesis (:160-171), then the reference's UKF
    structScore e;
    e.pos.x = est.mean_x;
    e.pos.y = est.mean_y;
    e.pos.ang = est.mean_ang;
    e.rotateScanImPoint = 0;
    e.score = est.mean_score;
    printf("Score:%lf\n", e.score);
    return ukf(FAInput, e);
}

}  // namespace myfa
