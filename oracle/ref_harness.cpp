// TEST INFRASTRUCTURE — not product code.
//
// C-ABI harness around the UNMODIFIED reference sources
// (/root/reference/LSD/{myLSD,myFA,myRDP,baseFunc}.cpp, threadpool.c), compiled
// where they lie by oracle/Makefile into oracle/_ref/libref_*.so.  It only
// marshals plain buffers into the reference's own entry points:
//   mylsd::myLineSegmentDetector   (LSD/myLSD.h:132, LSD/myLSD.cpp:129)
//   mylsd::createMapCache          (LSD/myLSD.h:131, LSD/myLSD.cpp:11)
//   myfa::NormalizedLineDirection / rotateScanIm / CalcScore
//                                   (LSD/myFA.h:85-87, LSD/myFA.cpp:274,307,357)
//   myfa::FeatureAssociation       (LSD/myFA.h:83,  LSD/myFA.cpp:13)
//   myrdp::FeatureScan             (LSD/myRDP.h:63, LSD/myRDP.cpp:9)
// Internal Mats of myLineSegmentDetector (usedMap, degMap, magMap, regIdx, the
// Gaussian image) are recovered through the allocation registry of
// shim/opencv.hpp; the sorted seed list through a qsort interposer below.
#include <myLSD.h>
#include <myFA.h>
#include <myRDP.h>
#include <dlfcn.h>
#include <string.h>
#include <vector>

// ---------------------------------------------------------------- qsort hook
// The reference sorts its nodeBinCell list with libc qsort (LSD/myLSD.cpp:204).
// The library is linked -Bsymbolic, so that call binds here; we forward to the
// real qsort and keep a copy of the result when armed.
static thread_local int g_seed_capture_armed = 0;
static thread_local std::vector<int>* g_seed_capture = 0;

extern "C" void qsort(void* base, size_t nmemb, size_t size, int (*compar)(const void*, const void*)) {
    typedef void (*qsort_fn)(void*, size_t, size_t, int (*)(const void*, const void*));
    static qsort_fn real = 0;
    if (!real) real = (qsort_fn)dlsym(RTLD_NEXT, "qsort");
    real(base, nmemb, size, compar);
    if (g_seed_capture_armed && g_seed_capture && size == sizeof(mylsd::nodeBinCell)) {
        const mylsd::nodeBinCell* c = (const mylsd::nodeBinCell*)base;
        g_seed_capture->resize(nmemb * 3);
        for (size_t i = 0; i < nmemb; i++) {
            (*g_seed_capture)[3 * i + 0] = c[i].value;
            (*g_seed_capture)[3 * i + 1] = c[i].x;
            (*g_seed_capture)[3 * i + 2] = c[i].y;
        }
        g_seed_capture_armed = 0;
    }
}

static void copy_mat(const cv::Mat& m, void* dst) {
    if (!dst || m.empty()) return;
    memcpy(dst, m.data, (size_t)m.rows * m.step);
}

extern "C" {

const char* ref_variant(void) {
#if defined(REF_VARIANT_DROPIN)
    return "dropin"; // myLineSegmentDetector / FeatureAssociation are this repo's B200 drop-in bodies
#elif defined(REF_VARIANT_LSDM)
    return "lsdm";   // libm calls interposed by the repo's portable math (oracle ii)
#else
    return "glibc";  // stock libm of this box (oracle i)
#endif
}

// Runs the reference LSD on a copy of `map` (rows x cols, u8).  Every output
// pointer may be NULL.  Returns the segment count (structLSD::len_linesInfo).
//   lines    : [max_lines][10] = k b dx dy x1 y1 x2 y2 len orient  (LSD/baseFunc.h:33-44)
//   line_im  : rows*cols u8      map_out : rows*cols u8 (the in-place remap, LSD/myLSD.cpp:135-142)
//   gauss/mag/deg : H'*W' f64    used/reg_idx : H'*W' u8
//   seeds    : [max_seeds][3] = bin,x,y in sorted order; *n_seeds = list length
int ref_lsd(const uint8_t* map, int cols, int rows, double sca, double sig, double angThre,
            double denThre, int pseBinArg, double* lines, int max_lines, uint8_t* line_im,
            uint8_t* map_out, double* gauss, double* mag, double* deg, uint8_t* used,
            uint8_t* reg_idx, int* seeds, int max_seeds, int* n_seeds) {
    cv::Mat m(rows, cols, CV_8UC1);
    memcpy(m.data, map, (size_t)rows * cols);

    std::vector<cv::Mat> kept;
    std::vector<int> seedv;
    cv::MatRegistry& reg = cv::mat_registry();
    // allocation order inside one call (LSD/myLSD.cpp:387-388,145-147,178,214-215):
    // 0 auxImage, 1 newImage(Gauss), 2 usedMap, 3 degMap, 4 magMap, 5 pseIdx, 6 regIdx, 7 lineIm
    reg.kept = &kept;
    reg.budget = 8;
    g_seed_capture = &seedv;
    g_seed_capture_armed = 1;

    mylsd::structLSD r = mylsd::myLineSegmentDetector(m, cols, rows, sca, sig, angThre, denThre, pseBinArg);

    reg.budget = 0;
    reg.kept = 0;
    g_seed_capture_armed = 0;
    g_seed_capture = 0;

    int n = r.len_linesInfo;
    if (lines) {
        for (int i = 0; i < n && i < max_lines; i++) {
            const structLinesInfo& L = r.linesInfo[i];
            double* o = lines + 10 * i;
            o[0] = L.k; o[1] = L.b; o[2] = L.dx; o[3] = L.dy; o[4] = L.x1; o[5] = L.y1;
            o[6] = L.x2; o[7] = L.y2; o[8] = L.len; o[9] = (double)L.orient;
        }
    }
    copy_mat(r.lineIm, line_im);
    copy_mat(m, map_out);
    if (kept.size() >= 8) {
        copy_mat(kept[1], gauss);
        copy_mat(kept[2], used);
        copy_mat(kept[3], deg);
        copy_mat(kept[4], mag);
        copy_mat(kept[6], reg_idx);
    }
    int ns = (int)(seedv.size() / 3);
    if (n_seeds) *n_seeds = ns;
    if (seeds) memcpy(seeds, seedv.data(), sizeof(int) * 3 * (size_t)(ns < max_seeds ? ns : max_seeds));
    free(r.linesInfo);
    return n;
}

// mylsd::createMapCache on a copy of `map`; out = rows*cols f64 (metres).
void ref_create_map_cache(const uint8_t* map, int cols, int rows, double res, double* out) {
    cv::Mat m(rows, cols, CV_8UC1);
    memcpy(m.data, map, (size_t)rows * cols);
    cv::Mat c = mylsd::createMapCache(m, res);
    copy_mat(c, out);
}

static void fill_fa_input(myfa::structFAInput& in, const double* scan_lines, int n_scan,
                          const double* map_lines, int n_map, const double* pts, int n_pts,
                          const double* map_cache, int cols, int rows, const double* lidar_pose,
                          const double* last_pose) {
    in.scanLinesInfo.resize(n_scan);
    in.mapLinesInfo.resize(n_map);
    for (int i = 0; i < n_scan; i++) {
        const double* s = scan_lines + 10 * i;
        structLinesInfo L = {s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], (int)s[9]};
        in.scanLinesInfo[i] = L;
    }
    for (int i = 0; i < n_map; i++) {
        const double* s = map_lines + 10 * i;
        structLinesInfo L = {s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], (int)s[9]};
        in.mapLinesInfo[i] = L;
    }
    in.scanImPoint.resize(n_pts);
    for (int i = 0; i < n_pts; i++) {
        in.scanImPoint[i].x = pts[2 * i];
        in.scanImPoint[i].y = pts[2 * i + 1];
        in.scanImPoint[i].ang = 0;
    }
    in.mapCache = cv::Mat(rows, cols, CV_64FC1);
    memcpy(in.mapCache.data, map_cache, sizeof(double) * (size_t)rows * cols);
    in.lidarPose.x = lidar_pose[0]; in.lidarPose.y = lidar_pose[1]; in.lidarPose.ang = 0;
    in.lastPose.x = last_pose[0]; in.lastPose.y = last_pose[1]; in.lastPose.ang = last_pose[2];
    in.ScanPose.x = in.ScanPose.y = in.ScanPose.ang = 0;
}

// Serial, deterministic replay of the scoring part of FeatureAssociation
// (pair filter LSD/myFA.cpp:29-41, the four pairings :194-235, pose build
// :238-244) calling the reference's own NormalizedLineDirection / rotateScanIm /
// CalcScore.  One record per (pair, pairing): out_idx[3] = iScan,iMap,iPair(1..4),
// out_val[4] = x,y,ang,score (score = +inf when gated out or < 70 % in bounds).
// Returns the number of records (pairs*4), at most max_rec are written.
int ref_fa_scores(const double* scan_lines, int n_scan, const double* map_lines, int n_map,
                  const double* pts, int n_pts, const double* map_cache, int cols, int rows,
                  const double* lidar_pose, const double* last_pose, int* out_idx, double* out_val,
                  int max_rec) {
    myfa::structFAInput in;
    fill_fa_input(in, scan_lines, n_scan, map_lines, n_map, pts, n_pts, map_cache, cols, rows, lidar_pose, last_pose);
    int nrec = 0;
    for (int is = 0; is < n_scan; is++) {
        double lenS = in.scanLinesInfo[is].len;
        if (lenS < ignoreScanLength) continue;
        double lenDiff = in.scanLinesInfo[is].len * scanToMapDiff;
        for (int im = 0; im < n_map; im++) {
            double lenM = in.mapLinesInfo[im].len;
            if (lenM < lenS - lenDiff || lenM > lenS + lenDiff) continue;
            const structLinesInfo& M = in.mapLinesInfo[im];
            const structLinesInfo& S = in.scanLinesInfo[is];
            for (int i = 1; i <= 4; i++) {
                myfa::structStaEnd ms, ss;
                if (i <= 2) { ms.staX = M.x1; ms.staY = M.y1; ms.endX = M.x2; ms.endY = M.y2; }
                else        { ms.staX = M.x2; ms.staY = M.y2; ms.endX = M.x1; ms.endY = M.y1; }
                if (i == 1 || i == 3) { ss.staX = S.x1; ss.staY = S.y1; ss.endX = S.x2; ss.endY = S.y2; }
                else                  { ss.staX = S.x2; ss.staY = S.y2; ss.endX = S.x1; ss.endY = S.y1; }
                structPosition mapPose, scanPose;
                mapPose.x = ms.staX; mapPose.y = ms.staY; mapPose.ang = myfa::NormalizedLineDirection(ms);
                scanPose.x = ss.staX; scanPose.y = ss.staY; scanPose.ang = myfa::NormalizedLineDirection(ss);
                myfa::structRotateScanIm RSI = myfa::rotateScanIm(&in, mapPose, scanPose, in.lastPose);
                double score = INFINITY, px = 0, py = 0, pa = 0;
                if (RSI.numScanImPoint != 0) {
                    px = RSI.rotateLidarPos.x; py = RSI.rotateLidarPos.y; pa = RSI.rotateLidarPos.ang;
                    score = myfa::CalcScore(&in, RSI);
                    free(RSI.rotateScanImPoint);
                }
                if (nrec < max_rec) {
                    out_idx[3 * nrec] = is; out_idx[3 * nrec + 1] = im; out_idx[3 * nrec + 2] = i;
                    out_val[4 * nrec] = px; out_val[4 * nrec + 1] = py; out_val[4 * nrec + 2] = pa; out_val[4 * nrec + 3] = score;
                }
                nrec++;
            }
        }
    }
    return nrec;
}

// The reference FeatureAssociation as shipped (30-thread pool + HMM/UKF), for
// CPU-baseline timing.  kalman_x9/kalman_P81 are in/out (row-major).
void ref_feature_association(const double* scan_lines, int n_scan, const double* map_lines, int n_map,
                             const double* pts, int n_pts, const double* map_cache, int cols, int rows,
                             const double* lidar_pose, const double* last_pose, const double* scan_pose,
                             double* kalman_x9, double* kalman_P81) {
    myfa::structFAInput in;
    fill_fa_input(in, scan_lines, n_scan, map_lines, n_map, pts, n_pts, map_cache, cols, rows, lidar_pose, last_pose);
    in.ScanPose.x = scan_pose[0]; in.ScanPose.y = scan_pose[1]; in.ScanPose.ang = scan_pose[2];
    for (int i = 0; i < 9; i++) in.kalman_x(i) = kalman_x9[i];
    for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) in.kalman_P(i, j) = kalman_P81[9 * i + j];
    myfa::structFAOutput out = myfa::FeatureAssociation(&in);
    for (int i = 0; i < 9; i++) kalman_x9[i] = out.kalman_x(i);
    for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) kalman_P81[9 * i + j] = out.kalman_P(i, j);
}

// myfa::ukf (LSD/myFA.cpp:404) alone: kalman_x9 / kalman_P81 in/out (row-major), scan_pose = odometry increment,
// est = x, y, ang of the pose estimate.  Deterministic (no thread pool involved).
void ref_ukf(double* kalman_x9, double* kalman_P81, const double* scan_pose, const double* est) {
    myfa::structFAInput in;
    in.ScanPose.x = scan_pose[0]; in.ScanPose.y = scan_pose[1]; in.ScanPose.ang = scan_pose[2];
    for (int i = 0; i < 9; i++) in.kalman_x(i) = kalman_x9[i];
    for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) in.kalman_P(i, j) = kalman_P81[9 * i + j];
    myfa::structScore e;
    e.pos.x = est[0]; e.pos.y = est[1]; e.pos.ang = est[2]; e.score = 0; e.rotateScanImPoint = 0;
    myfa::structFAOutput out = myfa::ukf(&in, e);
    for (int i = 0; i < 9; i++) kalman_x9[i] = out.kalman_x(i);
    for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) kalman_P81[9 * i + j] = out.kalman_P(i, j);
}

// myrdp::FeatureScan on one frame (finite ranges only, as LSD/main_on_windows.cpp:110-123
// filters them).  map_param = cols rows resol oriX oriY.  Outputs: lines [max_lines][10],
// pts [max_pts][2], lidar_pos[2], im_size[2] = cols,rows of the scan raster.
// Returns number of lines; *n_pts = number of raster points.
int ref_feature_scan(const double* map_param, const double* ranges, const double* angles, int n,
                     double* lines, int max_lines, double* pts, int max_pts, int* n_pts,
                     double* lidar_pos, int* im_size, uint8_t* line_im, int line_im_cap) {
    structMapParam mp;
    mp.oriMapCol = (int)map_param[0]; mp.oriMapRow = (int)map_param[1];
    mp.mapResol = map_param[2]; mp.mapOriX = map_param[3]; mp.mapOriY = map_param[4];
    std::vector<myrdp::structLidarPointPolar> lp(n > 0 ? n : 1);
    for (int i = 0; i < n; i++) { lp[i].range = ranges[i]; lp[i].angle = angles[i]; lp[i].split = false; }
    myrdp::structFeatureScan FS = myrdp::FeatureScan(mp, lp.data(), n, rdp_leastPoint, rdp_threLine, rdp_leastDist);
    for (int i = 0; i < FS.len_linesInfo && i < max_lines; i++) {
        const structLinesInfo& L = FS.linesInfo[i];
        double* o = lines + 10 * i;
        o[0] = L.k; o[1] = L.b; o[2] = L.dx; o[3] = L.dy; o[4] = L.x1; o[5] = L.y1;
        o[6] = L.x2; o[7] = L.y2; o[8] = L.len; o[9] = L.orient;
    }
    int np = (int)FS.scanImPoint.size();
    if (n_pts) *n_pts = np;
    for (int i = 0; i < np && i < max_pts; i++) { pts[2 * i] = FS.scanImPoint[i].x; pts[2 * i + 1] = FS.scanImPoint[i].y; }
    if (lidar_pos) { lidar_pos[0] = FS.lidarPos.x; lidar_pos[1] = FS.lidarPos.y; }
    if (im_size) { im_size[0] = FS.lineIm.cols; im_size[1] = FS.lineIm.rows; }
    if (line_im && (size_t)FS.lineIm.rows * FS.lineIm.cols <= (size_t)line_im_cap) copy_mat(FS.lineIm, line_im);
    int nl = FS.len_linesInfo;
    free(FS.linesInfo);
    return nl;
}

// FeatureScan over many frames with nothing copied out (the CPU timing leg of bench.py): returns the total line count,
// *n_pts_total = total raster samples.
long long ref_feature_scan_many(const double* map_param, const double* ranges, const double* angles, const int* beam_off,
                                int n_frames, long long* n_pts_total) {
    structMapParam mp;
    mp.oriMapCol = (int)map_param[0]; mp.oriMapRow = (int)map_param[1];
    mp.mapResol = map_param[2]; mp.mapOriX = map_param[3]; mp.mapOriY = map_param[4];
    long long nl = 0, np = 0;
    std::vector<myrdp::structLidarPointPolar> lp;
    for (int f = 0; f < n_frames; f++) {
        const int n = beam_off[f + 1] - beam_off[f];
        lp.resize(n > 0 ? n : 1);
        for (int i = 0; i < n; i++) { lp[i].range = ranges[beam_off[f] + i]; lp[i].angle = angles[beam_off[f] + i]; lp[i].split = false; }
        myrdp::structFeatureScan FS = myrdp::FeatureScan(mp, lp.data(), n, rdp_leastPoint, rdp_threLine, rdp_leastDist);
        nl += FS.len_linesInfo; np += (long long)FS.scanImPoint.size();
        free(FS.linesInfo);
    }
    if (n_pts_total) *n_pts_total = np;
    return nl;
}

}  // extern "C"
