// TEST INFRASTRUCTURE: a C entry point over the catkin snapshot's association (ROS/lsd/include/FeatureAssociation.h:46-60, the
// 13-argument myfa::FeatureAssociation), linked either with the UNMODIFIED ROS/lsd/src/FeatureAssociation.cpp (libref_rosfa.so with
// stock libm, libref_rosfa_lsdm.so with every libm call bound to lsd_math.h) or with this repo's drop-in body
// (host/FeatureAssociation_ros_b200.cpp -> liblsdb200.so, libdropin_rosfa.so).  Lines travel as n x 10 doubles in oracle column order
// (k b dx dy x1 y1 x2 y2 len orient); poseAll comes back as its 15 x T row-major matrix.
#include <FeatureAssociation.h>
#include <stdint.h>
#include <string.h>

static void to_lines(const double* a, int n, std::vector<structLinesInfo>& v) {
    v.resize(n);
    for (int i = 0; i < n; i++) {
        const double* o = a + 10 * i;
        structLinesInfo& L = v[i];
        L.k = o[0]; L.b = o[1]; L.dx = o[2]; L.dy = o[3]; L.x1 = o[4]; L.y1 = o[5]; L.x2 = o[6]; L.y2 = o[7]; L.len = o[8]; L.orient = (int)o[9];
    }
}

extern "C" int ros_fa(const double* scan_lines, int n_scan, const double* map_lines, int n_map, double resol, double ori_x, double ori_y,
                      const int* lidar_pos, int map_cols, int map_rows, const double* map_cache, const double* ranges, const double* angles,
                      int n_rays, double* pose_all, int max_cols, double* est, double* est_real) {
    std::vector<structLinesInfo> sl, ml;
    to_lines(scan_lines, n_scan, sl); to_lines(map_lines, n_map, ml);
    structMapParam mp; memset(&mp, 0, sizeof(mp));
    mp.oriMapCol = map_cols; mp.oriMapRow = map_rows; mp.mapResol = resol; mp.mapOriX = ori_x; mp.mapOriY = ori_y;
    cv::Mat scanIm(8, 8, CV_8UC1);                 // only handed through (RotateScanIm's use of it is commented out upstream)
    cv::Mat mapIm(map_rows, map_cols, CV_8UC1);    // MaplineIm: its size is the map's size; the "== 1" count it feeds is never used
    cv::Mat mapValue(1, 1, CV_8UC1);               // unused by the callee
    cv::Mat cache(map_rows, map_cols, CV_64FC1);
    memcpy(cache.data, map_cache, sizeof(double) * (size_t)map_rows * map_cols);
    std::vector<double> r(ranges, ranges + n_rays), a(angles, angles + n_rays);
    cv::Mat poseAll;
    // no candidate pair at all: the callee would read column 0 of an empty matrix (ROS/lsd/src/FeatureAssociation.cpp:119) — not called
    {
        const double d = 0.3 / resol;
        bool any = false;
        for (int i = 0; i < n_scan && !any; i++)
            for (int j = 0; j < n_map && !any; j++) any = ml[j].len >= sl[i].len - d && ml[j].len <= sl[i].len + d;
        if (!any) return 0;
    }
    myfa::FeatureAssociation(scanIm, sl, ml, mp, lidar_pos, mapIm, cache, mapValue, r, a, est_real, est, poseAll);
    const int T = poseAll.cols;
    for (int k = 0; k < 15 && pose_all; k++)
        for (int c = 0; c < T && c < max_cols; c++) pose_all[(size_t)k * max_cols + c] = *poseAll.ptr<double>(k, c);
    return T;
}
