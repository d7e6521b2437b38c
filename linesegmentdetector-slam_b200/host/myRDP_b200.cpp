// Drop-in body for myrdp::FeatureScan (reference LSD/myRDP.h:63, LSD/myRDP.cpp:9-185).
//
// Compiled against the reference's own myRDP.h (so structFeatureScan, structLidarPointPolar and structLinesInfo are
// the reference's) and linked with liblsdb200.so; the reference body is renamed out of the way on its translation unit
// (-DFeatureScan=FeatureScan_cpu on myRDP.cpp).  One frame per call, as the reference's callers
// use it (LSD/main_on_windows.cpp:127, LSD/main_on_linux.cpp:68); code that owns many frames should call
// lsdb_feature_scan_frames directly — one launch for all of them.
//
// Contract kept from the reference:
//   * linesInfo is malloc'd here (at least 360 records, :44) and owned by the caller; lineIm is a fresh CV_8UC1 Mat of
//     oriYLim x oriXLim; scanImPoint holds the raster samples in the reference's order; lidarPos as :38-40.
//   * beams must be finite (the callers drop Inf ranges, LSD/main_on_windows.cpp:110-123); len_lp >= 1.
//   * no error channel: failures abort with the library's message.
// Observable difference: the `split` flags of the caller's beams are not written (the callers reset them every frame
// and never read them, LSD/main_on_windows.cpp:121, LSD/main_on_linux.cpp:60).
#include <myRDP.h>

#include <stdlib.h>
#include <string.h>
#include <vector>

#include "lsdb_host.h"

static_assert(sizeof(structLinesInfo) == sizeof(lsdb_line), "structLinesInfo layout (LSD/baseFunc.h:33-44)");

namespace myrdp {

structFeatureScan FeatureScan(structMapParam mapParam, structLidarPointPolar* lidarPointPolar, int len_lp, int RegionPointLimitNumber,
                              double threLine, double lineDistThreM) {
    lsdb_ctx* ctx = lsdb_host::context();
    std::vector<double> ranges((size_t)(len_lp > 0 ? len_lp : 0)), angles(ranges.size());
    for (int i = 0; i < len_lp; i++) { ranges[i] = lidarPointPolar[i].range; angles[i] = lidarPointPolar[i].angle; }
    const int beamOff[2] = {0, len_lp};
    lsdb_rdp_params prm;
    prm.least_point = RegionPointLimitNumber; prm.thre_line = threLine; prm.least_dist_m = lineDistThreM;
    lsdb_scan_info info;
    int lineOff[2], ptOff[2];
    long long imOff[2];
    int rc = lsdb_feature_scan_frames(ctx, mapParam.mapResol, mapParam.mapOriX, mapParam.mapOriY, &prm, 1, ranges.data(), angles.data(),
                                      beamOff, &info, 0, 0, lineOff, 0, 0, ptOff, 0, 0, imOff);
    if (rc != LSDB_OK) lsdb_host::die("lsdb_feature_scan_frames", rc);
    std::vector<lsdb_line> lines((size_t)(info.n_lines > 0 ? info.n_lines : 1));
    std::vector<double> pts(2 * (size_t)(info.n_pts > 0 ? info.n_pts : 1));
    std::vector<uint8_t> im((size_t)(imOff[1] > 0 ? imOff[1] : 1));
    rc = lsdb_feature_scan_frames(ctx, mapParam.mapResol, mapParam.mapOriX, mapParam.mapOriY, &prm, 1, ranges.data(), angles.data(),
                                  beamOff, &info, lines.data(), (int)lines.size(), lineOff, pts.data(), (int)(pts.size() / 2), ptOff,
                                  im.data(), (long long)im.size(), imOff);
    if (rc != LSDB_OK) lsdb_host::die("lsdb_feature_scan_frames", rc);

    structFeatureScan FS;
    const int W = info.im_cols > 0 ? info.im_cols : 0, H = info.im_rows > 0 ? info.im_rows : 0;
    FS.lineIm = Mat::zeros(H, W, CV_8UC1);
    for (int y = 0; y < H && W > 0; y++) memcpy(FS.lineIm.ptr<uint8_t>(y), &im[(size_t)y * W], (size_t)W);
    const int n = info.n_lines;
    FS.linesInfo = (structLinesInfo*)malloc(sizeof(structLinesInfo) * (size_t)(n > 360 ? n : 360));
    memcpy(FS.linesInfo, lines.data(), sizeof(structLinesInfo) * (size_t)n);
    FS.len_linesInfo = n;
    FS.lidarPos.x = info.lidar_x; FS.lidarPos.y = info.lidar_y; FS.lidarPos.num = 0;
    FS.scanImPoint.resize((size_t)info.n_pts);
    for (int i = 0; i < info.n_pts; i++) { FS.scanImPoint[i].x = pts[2 * (size_t)i]; FS.scanImPoint[i].y = pts[2 * (size_t)i + 1]; FS.scanImPoint[i].ang = 0; }
    return FS;
}

}  // namespace myrdp
