// TEST INFRASTRUCTURE: C entry points over the catkin snapshot's API (ROS/lsd/include/myLSD.h:131-132) — the three-argument
// createMapCache and myLineSegmentDetector with `double pseBin` — linked either with the UNMODIFIED ROS/lsd/src/myLSD.cpp
// (libref_ros.so, LSD only: its createMapCache reads past the map) or with this repo's drop-in bodies built -DLSDB_ROS_FLAVOUR
// (libdropin_ros.so).  Same record layout as ref_harness.cpp:ref_lsd.
#include <myLSD.h>
#include <stdint.h>
#include <string.h>

extern "C" {

int ros_lsd(const uint8_t* map, int cols, int rows, double sca, double sig, double angThre, double denThre, double pseBinArg,
            double* lines, int max_lines, uint8_t* line_im, uint8_t* map_out) {
    cv::Mat m(rows, cols, CV_8UC1);
    memcpy(m.data, map, (size_t)rows * cols);
    mylsd::structLSD r = mylsd::myLineSegmentDetector(m, cols, rows, sca, sig, angThre, denThre, pseBinArg);
    const int n = r.len_linesInfo;
    for (int i = 0; i < n && i < max_lines && lines; i++) {
        const structLinesInfo& L = r.linesInfo[i];
        double* o = lines + 10 * i;
        o[0] = L.k; o[1] = L.b; o[2] = L.dx; o[3] = L.dy; o[4] = L.x1; o[5] = L.y1; o[6] = L.x2; o[7] = L.y2; o[8] = L.len; o[9] = (double)L.orient;
    }
    if (line_im) for (int y = 0; y < rows; y++) memcpy(line_im + (size_t)y * cols, r.lineIm.ptr<uint8_t>(y), (size_t)cols);
    if (map_out) memcpy(map_out, m.data, (size_t)rows * cols);
    return n;
}

#ifndef ROS_HARNESS_NO_MAP_CACHE
void ros_map_cache(const uint8_t* map, int cols, int rows, double res, double zmax, double* out) {
    cv::Mat m(rows, cols, CV_8UC1);
    memcpy(m.data, map, (size_t)rows * cols);
    cv::Mat c = mylsd::createMapCache(m, res, zmax);
    for (int y = 0; y < rows; y++) memcpy(out + (size_t)y * cols, c.ptr<double>(y), sizeof(double) * (size_t)cols);
}
#endif

}
