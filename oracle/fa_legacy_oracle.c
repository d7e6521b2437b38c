/* TEST INFRASTRUCTURE — plain-C restatement of the catkin snapshot's association
 * (/root/reference/ROS/lsd/src/FeatureAssociation.cpp: :5-34 NormalizedLineDirection, :36-130 FeatureAssociation,
 * :132-200 ScanToMapMatch, :202-252 ScanToMapMatchScore, :254-299 RotateScanIm), serial, in the reference's pair
 * order and summation order, libm calls through lsd_math.h ("oracle (ii)" arithmetic).  Pinned against the unmodified
 * source (oracle/_ref/libref_rosfa_lsdm.so bit for bit, libref_rosfa.so to rounding) by tests/test_fa_legacy.py. */
#include "lsd_oracle.h"
#include "lsd_math.h"
#include <stdlib.h>

#define PI_ 3.14159265358979323846 /* M_PI */

static double nld(double x1, double y1, double x2, double y2) { /* :5-34 */
    double ang = 0;
    double dy = y2 - y1, dx = x2 - x1;
    if (dy && !dx) ang = dy > 0 ? 90 : -90;
    else if (!dy && dx) ang = dx > 0 ? 0 : 180;
    else ang = lsdm_atan(dy / dx) * 180 / PI_;
    if (dx < 0) {
        if (ang < 0) ang += 180;
        else if (ang > 0) ang -= 180;
    }
    return ang;
}

static double match_score(const double* pose, const double* cache, int cols, int rows, const double* ranges, const double* angles,
                          int n, double resol) { /* :202-252 */
    if (pose[0] > cols || pose[0] < 1 || pose[1] > rows || pose[1] < 1) return INFINITY;
    double dist = 0, dist_count = 0, max_count = 0, scanlen = 0;
    const unsigned sizeX = (unsigned)cols, sizeY = (unsigned)rows;
    for (int i = 0; i < n; i++) {
        const double gx = floor(ranges[i] * lsdm_cos(angles[i] + pose[2] * PI_ / 180) / resol) + pose[0] - 1;
        const double gy = floor(ranges[i] * lsdm_sin(angles[i] + pose[2] * PI_ / 180) / resol) + pose[1] - 1;
        if (gx > 1 && gx < sizeX && gy > 1 && gy < sizeY) {
            scanlen++;
            const double v = cache[(size_t)(int)gy * cols + (int)gx];
            if (v == 2) max_count++;
            else { dist += v; dist_count++; }
        }
    }
    if (scanlen < (size_t)n * 0.75) return INFINITY;
    return (dist + 7 * max_count) / (dist_count + max_count) + 10 * (n - scanlen) / n;
}

/* lines: n x 10 doubles (k b dx dy x1 y1 x2 y2 len orient).  pose_all: T records of 15 doubles (the COLUMNS of the reference's
 * 15 x T poseAll), at most max_cols of them written; returns T.  est / est_real (3 doubles each) are set when T > 0. */
int lsdo_fa_legacy(const double* scan_lines, int n_scan, const double* map_lines, int n_map, double resol, double ori_x, double ori_y,
                   const int* lidar_pos, int cols, int rows, const double* map_cache, const double* ranges, const double* angles,
                   int n_rays, double* pose_all, int max_cols, double* est, double* est_real) {
    const double len_diff = 0.3 / resol; /* :61-62 */
    int T = 0, best = -1;
    double best_rec[15];
    for (int i = 0; i < n_scan; i++) {
        const double* S = scan_lines + 10 * i;
        const double target = S[8];
        for (int j = 0; j < n_map; j++) {
            const double* M = map_lines + 10 * j;
            if (!(M[8] >= target - len_diff && M[8] <= target + len_diff)) continue; /* :69 */
            for (int k = 0; k < 4; k++) { /* :157-197 */
                double mp[4], sp[4], rec[15];
                if (k < 2) { mp[0] = M[4]; mp[1] = M[5]; mp[2] = M[6]; mp[3] = M[7]; }
                else { mp[0] = M[6]; mp[1] = M[7]; mp[2] = M[4]; mp[3] = M[5]; }
                if ((k & 1) == 0) { sp[0] = S[4]; sp[1] = S[5]; sp[2] = S[6]; sp[3] = S[7]; }
                else { sp[0] = S[6]; sp[1] = S[7]; sp[2] = S[4]; sp[3] = S[5]; }
                const double mdir = nld(mp[0], mp[1], mp[2], mp[3]);
                const double sdir = nld(sp[0], sp[1], sp[2], sp[3]);
                /* RotateScanIm :254-299 */
                const double ang_diff = mdir - sdir;
                const double c = lsdm_cos(ang_diff / 180 * PI_), s = lsdm_sin(ang_diff / 180 * PI_);
                rec[0] = floor((lidar_pos[0] - sp[0]) * c - (lidar_pos[1] - sp[1]) * s + mp[0]);
                rec[1] = floor((lidar_pos[0] - sp[0]) * s + (lidar_pos[1] - sp[1]) * c + mp[1]);
                rec[2] = sdir + ang_diff;
                rec[3] = match_score(rec, map_cache, cols, rows, ranges, angles, n_rays, resol);
                for (int q = 0; q < 4; q++) { rec[4 + q] = mp[q]; rec[8 + q] = sp[q]; }
                rec[12] = (unsigned)i; rec[13] = (unsigned)j; rec[14] = k;
                if (pose_all && T < max_cols) for (int q = 0; q < 15; q++) pose_all[(size_t)T * 15 + q] = rec[q];
                if (best < 0 || rec[3] < best_rec[3]) { /* :117-119: starts at column 0, moves on a strict < only */
                    best = T;
                    for (int q = 0; q < 15; q++) best_rec[q] = rec[q];
                }
                T++;
            }
        }
    }
    if (T > 0) { /* :120-127 */
        est[0] = best_rec[0]; est[1] = best_rec[1]; est[2] = best_rec[2] / 180 * PI_;
        est_real[0] = est[0] * resol + ori_x; est_real[1] = est[1] * resol + ori_y; est_real[2] = est[2];
    }
    return T;
}
