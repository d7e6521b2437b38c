// Drop-in body for the catkin snapshot's myfa::FeatureAssociation (reference ROS/lsd/include/FeatureAssociation.h:46-60,
// ROS/lsd/src/FeatureAssociation.cpp:36-130; called by LSD/main_on_linux.cpp:132 and ROS/lsd/src/main_on_linux.cpp) — SURVEY.md §8 f4.
//
// Every hypothesis of the frame (the length filter :63-70, the four pairings of ScanToMapMatch :132-200, RotateScanIm :254-299 and
// the ray re-projection score ScanToMapMatchScore :202-252) is evaluated by lsdb_fa_legacy on the device; poseAll comes back as the
// reference's 15 x T matrix and the two estimates as the reference computes them (:117-127).  ScanlineIm, MapValue and the pixel
// values of MaplineIm are not read by the reference either (only MaplineIm's size is).
//
// Build: compile against the snapshot's headers (-I ROS/lsd/include) in place of ROS/lsd/src/FeatureAssociation.cpp, or next to it with
// its FeatureAssociation renamed away when the helpers (samplePos, ...) are wanted.  Differences a caller can observe:
//   * no candidate pair at all: the reference reads column 0 of an empty poseAll (undefined); here poseAll is 15 x 0 and the
//     estimates are left untouched;
//   * the device copy of MapCache / MaplinesInfo is cached across frames, keyed by the Mat's data pointer, its size and a hash of
//     the map lines (the node builds both once per map, LSD/main_on_linux.cpp:78-86).
#include <FeatureAssociation.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lsdb_host.h"

static_assert(sizeof(structLinesInfo) == sizeof(lsdb_line), "structLinesInfo layout (ROS/lsd/include/baseFunc.h:25-36)");

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
}  // namespace

namespace myfa {

void FeatureAssociation(const Mat& ScanlineIm, const vector<structLinesInfo>& ScanlinesInfo, const vector<structLinesInfo>& MaplinesInfo,
                        const structMapParam& MapParam, const int* LidarPos, const Mat& MaplineIm, const Mat& MapCache, const Mat& MapValue,
                        const vector<double>& ScanRanges, const vector<double>& ScanAngles, double* estimatePose_realworld,
                        double* estimatePose, Mat& poseAll) {
    (void)ScanlineIm; (void)MapValue;
    lsdb_ctx* ctx = lsdb_host::context();
    const int nMap = (int)MaplinesInfo.size(), nScan = (int)ScanlinesInfo.size(), nRays = (int)ScanRanges.size();
    if (MapCache.rows != MaplineIm.rows || MapCache.cols != MaplineIm.cols) {
        fprintf(stderr, "lsdb200: FeatureAssociation: MapCache (%d x %d) and MaplineIm (%d x %d) differ in size\n", MapCache.cols, MapCache.rows,
                MaplineIm.cols, MaplineIm.rows);
        abort();
    }
    MapKey k = {MapCache.data, MapCache.rows, MapCache.cols, nMap, nMap ? fnv1a(&MaplinesInfo[0], sizeof(structLinesInfo) * (size_t)nMap) : 0ull,
                sample_cells(MapCache)};
    if (!g_map || !(k == g_key)) {
        if (g_map) lsdb_fa_map_destroy(g_map);
        g_map = 0;
        const int rc = lsdb_fa_map_create(ctx, MapCache.ptr<double>(0), MapCache.cols, MapCache.rows,
                                          nMap ? reinterpret_cast<const lsdb_line*>(&MaplinesInfo[0]) : 0, nMap, &g_map);
        if (rc) lsdb_host::die("lsdb_fa_map_create", rc);
        g_key = k;
    }
    // every scan line can pair with every map line, four pairings each
    const size_t cap = 4 * (size_t)nScan * (size_t)nMap;
    std::vector<double> rec(cap * 15 + 1);
    int T = 0;
    const int rc = lsdb_fa_legacy(ctx, g_map, nScan ? reinterpret_cast<const lsdb_line*>(&ScanlinesInfo[0]) : 0, nScan, MapParam.mapResol,
                                  MapParam.mapOriX, MapParam.mapOriY, LidarPos, nRays ? &ScanRanges[0] : 0, nRays ? &ScanAngles[0] : 0, nRays,
                                  &rec[0], (int)cap, &T, estimatePose, estimatePose_realworld);
    if (rc) lsdb_host::die("lsdb_fa_legacy", rc);
    poseAll.create(15, T, CV_64F);
    for (int c = 0; c < T; c++)
        for (int r = 0; r < 15; r++) *poseAll.ptr<double>(r, c) = rec[(size_t)c * 15 + r];
}

}  // namespace myfa
