/* TEST INFRASTRUCTURE — plain-C restatement of myrdp::FeatureScan, the scan front-end that
 * feeds the association path (/root/reference/LSD/myRDP.cpp:9-185 FeatureScan, :187-216
 * SplitMerge, :218-272 SplitMergeAssistant, :274-352 RegionSegmentation, :354-375
 * getThresholdDeltaDist).  Serial, in the reference's evaluation order; the quirks are kept
 * (maxX/maxY start at 0, the head-tail join overwrites cell 0, samples on row/column 0 are
 * dropped, the last raster column/row is out of bounds).  Pinned against oracle/_ref
 * (ref_feature_scan) on every bundled Lidar.txt frame by tests/test_oracle.py. */
#include "lsd_oracle.h"
#include "lsd_math.h"
#include <math.h>
#include <stdlib.h>

static double delta_thre(double r) { /* LSD/myRDP.cpp:354-375 */
    if (r <= 0.3) return 0.02;
    if (r <= 0.5) return 0.05;
    if (r <= 0.8) return 0.11;
    if (r <= 1) return 0.17;
    if (r <= 2) return 0.6;
    if (r <= 3) return 0.7;
    if (r <= 4) return 0.85;
    if (r <= 5) return 0.9;
    if (r <= 6) return 1;
    return 1.1;
}

/* x86 cvttsd2si: out-of-range and NaN give INT_MIN */
static int d2i(double v) { return (v >= -2147483648.0 && v < 2147483648.0) ? (int)v : (int)0x80000000; }

/* SplitMergeAssistant (:218-272).  The recursion only ever SETS split flags and each call
 * depends on (start, end) alone, so an explicit stack gives the same flags. */
static void split_merge(const double* range, const double* px, const double* py, int n, int start, int end,
                        double threLine, uint8_t* split, int* stack) {
    int sp = 0;
    stack[sp++] = start; stack[sp++] = end;
    while (sp) {
        int e = stack[--sp], s = stack[--sp];
        int len = e > s ? e - s + 1 : n + e - s + 1;
        if (len <= 2) continue;
        double k = (py[e] - py[s]) / (px[e] - px[s]);
        double d = py[e] - k * px[e];
        double dist_max = 0;
        int i_max = 0;
        for (int i = 1; i < len - 1; i++) {
            int a = s + i; if (a >= n) a -= n;
            double dist = fabs(k * px[a] - py[a] + d) / sqrt(k * k + 1);
            if (dist > dist_max) { dist_max = dist; i_max = a; }
        }
        double thre = range[i_max] > 9 ? range[i_max] * threLine : threLine;
        if (dist_max > thre) {
            split[i_max] = 1;
            stack[sp++] = s; stack[sp++] = i_max;
            stack[sp++] = i_max; stack[sp++] = e;
        }
    }
}

int lsdo_feature_scan(const double* map_param, const double* range, const double* angle, int n,
                      int leastPoint, double threLine, double leastDistM, double* lines, int max_lines,
                      double* pts, int max_pts, int* n_pts, double* lidar_pos, int* im_size,
                      uint8_t* line_im, int line_im_cap) {
    const double resol = map_param[2], oriX = map_param[3], oriY = map_param[4];
    const double pose[3] = {0, 0, 0};
    double* px = (double*)malloc(sizeof(double) * (n + 1) * 4);
    double *py = px + (n + 1), *gx = py + (n + 1), *gy = gx + (n + 1);
    int* cell = (int*)malloc(sizeof(int) * (4 * n + 8));
    int* stack = cell + 2 * n + 2;
    uint8_t* split = (uint8_t*)calloc(n + 1, 1);
    for (int i = 0; i < n; i++) { /* :289-293 */
        px[i] = range[i] * lsdm_cos(angle[i] + pose[2]) + pose[0];
        py[i] = range[i] * lsdm_sin(angle[i] + pose[2]) + pose[1];
    }
    /* RegionSegmentation :294-352.  cell[2c], cell[2c+1] = start, end point numbers */
    int startNum = 0, nc = 0;
    for (int i = 0; i < n; i++) {
        int j = i == n - 1 ? 0 : i + 1;
        double dx = px[i] - px[j], dy = py[i] - py[j];
        double dd = sqrt(dx * dx + dy * dy);
        double thre = delta_thre(range[i]);
        if (dd > thre) {
            cell[2 * nc] = startNum; cell[2 * nc + 1] = i;
            if (abs(i - startNum) >= leastPoint) nc++;
            startNum = i + 1;
        }
        if (dd <= thre && i == n - 1) cell[0] = startNum;   /* head-tail join */
    }
    for (int c = 0; c < nc; c++) split_merge(range, px, py, n, cell[2 * c], cell[2 * c + 1], threLine, split, stack);

    /* FeatureScan :17-41 */
    double minX = INFINITY, minY = INFINITY, maxX = 0, maxY = 0;
    for (int i = 0; i < n; i++) {
        gx[i] = floor((range[i] * lsdm_cos(angle[i] + pose[2]) + pose[0] - oriX) / resol);
        gy[i] = floor((range[i] * lsdm_sin(angle[i] + pose[2]) + pose[1] - oriY) / resol);
        if (gx[i] < minX) minX = gx[i];
        if (gx[i] > maxX) maxX = gx[i];
        if (gy[i] < minY) minY = gy[i];
        if (gy[i] > maxY) maxY = gy[i];
    }
    int W = d2i(ceil(maxX - minX)), H = d2i(ceil(maxY - minY));
    if (lidar_pos) {
        lidar_pos[0] = floor((pose[0] - oriX) / resol - minX);
        lidar_pos[1] = floor((pose[1] - oriY) / resol - minY);
    }
    if (im_size) { im_size[0] = W; im_size[1] = H; }
    int raster = line_im && W > 0 && H > 0 && (long long)W * H <= line_im_cap;
    if (raster) for (long long i = 0; i < (long long)W * H; i++) line_im[i] = 0;
    const double pi = 4.0 * lsdm_atan(1.0);
    double distThre = leastDistM / resol;
    int nl = 0, np = 0;
    for (int c = 0; c < nc; c++) { /* :49-176 */
        int s = cell[2 * c], e = cell[2 * c + 1];
        int len = e > s ? e - s + 1 : n + e - s + 1;
        int a = s;                                          /* axis[j] */
        for (int j = 0; j <= len; j++) {                    /* j == len closes with the end point */
            int b;
            if (j < len) { b = s + j; if (b >= n) b -= n; if (!split[b]) continue; }
            else b = e;
            double ax = gx[a], ay = gy[a], bx = gx[b], by = gy[b];
            a = b;
            double ddx = ax - bx, ddy = ay - by;
            double dist = sqrt(ddx * ddx + ddy * ddy);
            if (!(dist >= distThre)) continue;
            double x1 = ax - minX, y1 = ay - minY, x2 = bx - minX, y2 = by - minY;
            double k = (y2 - y1) / (x2 - x1);
            double ang = lsdm_atan(k) * 180.0 / pi;
            int orient = 1;
            if (ang < 0) { ang += 180; orient = -1; }
            int xLow = d2i(floor(x1 > x2 ? x2 : x1)), xHigh = d2i(ceil(x1 > x2 ? x1 : x2));
            int yLow = d2i(floor(y1 > y2 ? y2 : y1)), yHigh = d2i(ceil(y1 > y2 ? y1 : y2));
            double xr = fabs(x2 - x1), yr = fabs(y2 - y1);
            int xl = xHigh - xLow + 1, yl = yHigh - yLow + 1;
            int cnt = xl > yl ? xl : yl;                    /* emission count :132-153 */
            for (int m = 0; m < cnt; m++) {
                int xx, yy;
                if (xr > yr) { xx = m + xLow; yy = d2i(round((xx - x1) * k + y1)); }
                else { yy = m + yLow; xx = d2i(round((yy - y1) / k + x1)); }
                if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
                if (xx == 0 || yy == 0) continue;
                if (raster) line_im[(long long)yy * W + xx] = 255;
                if (np < max_pts) { pts[2 * np] = xx; pts[2 * np + 1] = yy; }
                np++;
            }
            if (nl < max_lines) {
                double* o = lines + 10 * nl;
                o[0] = k; o[1] = (y1 + y2) / 2.0 - k * (x1 + x2) / 2.0;
                o[2] = lsdm_cos(ang / 180.0 * pi); o[3] = lsdm_sin(ang / 180.0 * pi);
                o[4] = x1; o[5] = y1; o[6] = x2; o[7] = y2;
                double ey = y2 - y1, ex = x2 - x1;
                o[8] = sqrt(ey * ey + ex * ex); o[9] = orient;
            }
            nl++;
        }
    }
    if (n_pts) *n_pts = np;
    free(px); free(cell); free(split);
    return nl;
}

/* the same over many frames, nothing copied out (CPU timing leg of bench.py when oracle/_ref is absent) */
long long lsdo_feature_scan_many(const double* map_param, const double* range, const double* angle, const int* beam_off,
                                 int n_frames, long long* n_pts_total) {
    long long nl = 0, np = 0;
    uint8_t* im = (uint8_t*)malloc(1 << 22);
    for (int f = 0; f < n_frames; f++) {
        int n = beam_off[f + 1] - beam_off[f], npf = 0, sz[2];
        nl += lsdo_feature_scan(map_param, range + beam_off[f], angle + beam_off[f], n, 3, 0.08, 0.5, 0, 0, 0, 0, &npf, 0, sz, im, 1 << 22);
        np += npf;
    }
    free(im);
    if (n_pts_total) *n_pts_total = np;
    return nl;
}
