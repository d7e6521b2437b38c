/* TEST INFRASTRUCTURE — plain-C restatement of the scoring part of myfa::FeatureAssociation
 * (/root/reference/LSD/myFA.cpp:27-59 pair filter, :186-272 four pairings, :274-305
 * NormalizedLineDirection, :307-355 rotateScanIm, :357-396 CalcScore), serial and in the
 * reference's summation order.  Pinned against oracle/_ref by tests/test_oracle.py. */
#include "lsd_oracle.h"
#include "lsd_math.h"
#include <stdlib.h>

static const double kIgnoreScanLength = 40;  /* LSD/baseFunc.h:80 */
static const double kScanToMapDiff = 0.35;   /* :82 */
static const double kMaxEstiDist = 60;       /* :86 */
static const double kZOccMaxDis = 1;         /* :60 */

static double sind_(double x, double pi) { return lsdm_sin(x / 180.0 * pi); }  /* LSD/baseFunc.cpp:6-12 */
static double cosd_(double x, double pi) { return lsdm_cos(x / 180.0 * pi); }
static double atand_(double x, double pi) { return lsdm_atan(x) * 180.0 / pi; }

static double norm_line_dir(double staX, double staY, double endX, double endY, double pi) { /* :274-305 */
    double angle;
    if (staX == endX && staY != endY) angle = staY < endY ? 90 : -90;
    else if (staX != endX && staY == endY) angle = staX < endX ? 0 : 180;
    else angle = atand_((endY - staY) / (endX - staX), pi);
    if (angle < 0 && staX > endX) return angle + 180;
    if (angle > 0 && staX > endX) return angle - 180;
    return angle;
}

int lsdo_fa_scores(const double* scan_lines, int n_scan, const double* map_lines, int n_map,
                   const double* pts, int n_pts, const double* map_cache, int cols, int rows,
                   const double* lidar_pose, const double* last_pose, int32_t* out_idx,
                   double* out_val, int max_rec) {
    const double pi = 4.0 * lsdm_atan(1.0);
    int nrec = 0;
    for (int is = 0; is < n_scan; is++) {
        const double* S = scan_lines + 10 * is;
        double lenS = S[8];
        if (lenS < kIgnoreScanLength) continue;
        double lenDiff = lenS * kScanToMapDiff;
        for (int im = 0; im < n_map; im++) {
            const double* M = map_lines + 10 * im;
            double lenM = M[8];
            if (lenM < lenS - lenDiff || lenM > lenS + lenDiff) continue;
            for (int i = 1; i <= 4; i++) {
                double msx, msy, mex, mey, ssx, ssy, sex, sey;
                if (i <= 2) { msx = M[4]; msy = M[5]; mex = M[6]; mey = M[7]; }
                else { msx = M[6]; msy = M[7]; mex = M[4]; mey = M[5]; }
                if (i == 1 || i == 3) { ssx = S[4]; ssy = S[5]; sex = S[6]; sey = S[7]; }
                else { ssx = S[6]; ssy = S[7]; sex = S[4]; sey = S[5]; }
                double mapAng = norm_line_dir(msx, msy, mex, mey, pi);
                double scanAng = norm_line_dir(ssx, ssy, sex, sey, pi);
                /* rotateScanIm :307-355 */
                double angDiff = mapAng - scanAng;
                double cs = cosd_(angDiff, pi), sn = sind_(angDiff, pi);
                double lx = (lidar_pose[0] - ssx) * cs - (lidar_pose[1] - ssy) * sn + msx;
                double ly = (lidar_pose[0] - ssx) * sn + (lidar_pose[1] - ssy) * cs + msy;
                double ddx = lx - last_pose[0], ddy = ly - last_pose[1];
                double score = INFINITY, pa = 0, px = 0, py = 0;
                if (sqrt(ddx * ddx + ddy * ddy) < kMaxEstiDist || last_pose[0] == -1) {
                    /* CalcScore :357-396 */
                    double sumValid = 0, sumMax = 0, numValid = 0;
                    for (int k = 0; k < n_pts; k++) {
                        double ox = pts[2 * k] - ssx, oy = pts[2 * k + 1] - ssy;
                        double rx = ox * cs - oy * sn + msx;
                        double ry = ox * sn + oy * cs + msy;
                        int x = (int)round(rx), y = (int)round(ry);
                        if (y >= 0 && y < rows && x >= 0 && x < cols) {
                            numValid += 1;
                            double v = map_cache[(size_t)y * cols + x];
                            if (v >= kZOccMaxDis) sumMax += 10;
                            else sumValid += v;
                        }
                    }
                    double a = angDiff;
                    while (a <= -180) a += 360;
                    while (a > 180) a -= 360;
                    px = lx; py = ly; pa = a;
                    double numAll = n_pts;
                    if (n_pts == 0) score = INFINITY; /* RSI.numScanImPoint == 0 -> INFINITY (:256-258) */
                    else if (numValid < 0.7 * numAll) score = INFINITY;
                    else score = (sumValid + sumMax) / numValid + 10 * (numAll - numValid) / numAll;
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
