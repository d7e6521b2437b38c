// The catkin snapshot's association on sm_100a: every (scan line, map line of similar length, end-point
// pairing) hypothesis of one lidar frame gets its pose and its ray re-projection score in ONE launch — the replacement for the
// serial loops of /root/reference/ROS/lsd/src/FeatureAssociation.cpp (:36-130 FeatureAssociation, :132-200 ScanToMapMatch,
// :202-252 ScanToMapMatchScore, :254-299 RotateScanIm, :5-34 NormalizedLineDirection).
//
// One WARP per hypothesis.  Lane 0's pose algebra is what every lane computes (same inputs, same FMA-free double code as the
// oracle: lsd_math.h), then the lanes take the frame's rays 32 at a time: each re-projects its ray (correctly rounded cos / sin),
// looks its cell up in mapCache and the warp adds the distances IN RAY ORDER — a serial shuffle chain over the in-range rays —
// so the score has the bits of the reference's sequential sum, not just its value.  Counts are integers and order-free.
// The result record is the column the reference stores in its 15 x T `poseAll` (pose, score, the two end-point quadruples, the
// scan-line / map-line / pairing indices).  mapCache (8 B per map pixel) stays L2-resident; the stage is FP64-issue-bound.
#include "lsdb_common.cuh"

#define FAL_THREADS 128
#define FAL_PI 3.14159265358979323846   // M_PI, the snapshot's constant (:2-3)

__device__ double fal_norm_line_dir(double x1, double y1, double x2, double y2) {   // :5-34
    double ang = 0;
    const double dy = y2 - y1, dx = x2 - x1;
    if (dy != 0 && !(dx != 0)) ang = dy > 0 ? 90 : -90;
    else if (!(dy != 0) && dx != 0) ang = dx > 0 ? 0 : 180;
    else ang = lsdm_atan(dy / dx) * 180 / FAL_PI;
    if (dx < 0) {
        if (ang < 0) ang += 180;
        else if (ang > 0) ang -= 180;
    }
    return ang;
}

__global__ void __launch_bounds__(FAL_THREADS) lsdb_fa_legacy_kernel(int nHyp, const int2* __restrict__ pairs,
                                                                     const LsdbFaLine* __restrict__ scanLines,
                                                                     const LsdbFaLine* __restrict__ mapLines, int lidarX, int lidarY,
                                                                     const double* __restrict__ mapCache, int cols, int rows, double resol,
                                                                     const double* __restrict__ ranges, const double* __restrict__ angles,
                                                                     int nRays, double* __restrict__ out) {
    const int hyp = (blockIdx.x * FAL_THREADS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (hyp >= nHyp) return;
    const int2 pr = pairs[hyp >> 2];
    const int k = hyp & 3;
    const LsdbFaLine S = scanLines[pr.x], M = mapLines[pr.y];
    double mp[4], sp[4];   // the four pairings, :157-182
    if (k < 2) { mp[0] = M.x1; mp[1] = M.y1; mp[2] = M.x2; mp[3] = M.y2; } else { mp[0] = M.x2; mp[1] = M.y2; mp[2] = M.x1; mp[3] = M.y1; }
    if ((k & 1) == 0) { sp[0] = S.x1; sp[1] = S.y1; sp[2] = S.x2; sp[3] = S.y2; } else { sp[0] = S.x2; sp[1] = S.y2; sp[2] = S.x1; sp[3] = S.y1; }
    const double mdir = fal_norm_line_dir(mp[0], mp[1], mp[2], mp[3]);
    const double sdir = fal_norm_line_dir(sp[0], sp[1], sp[2], sp[3]);
    const double angDiff = mdir - sdir;                                                  // RotateScanIm :264
    const double cs = lsdm_cos(angDiff / 180 * FAL_PI), sn = lsdm_sin(angDiff / 180 * FAL_PI);   // :282-283
    const double px = floor((lidarX - sp[0]) * cs - (lidarY - sp[1]) * sn + mp[0]);      // :295-297
    const double py = floor((lidarX - sp[0]) * sn + (lidarY - sp[1]) * cs + mp[1]);
    const double pang = sdir + angDiff;

    double score;
    if (px > cols || px < 1 || py > rows || py < 1) score = INFINITY;                    // :211-212
    else {
        const double th = pang * FAL_PI / 180;
        const double sizeX = (double)(unsigned)cols, sizeY = (double)(unsigned)rows;
        double dist = 0.0;
        int distCount = 0, maxCount = 0, scanLen = 0;
        for (int base = 0; base < nRays; base += 32) {
            const int i = base + lane;
            bool in = false, isMax = false;
            double v = 0.0;
            if (i < nRays) {
                const double r = ranges[i], a = angles[i] + th;
                const double gx = floor(r * lsdm_cos(a) / resol) + px - 1;               // :226-227
                const double gy = floor(r * lsdm_sin(a) / resol) + py - 1;
                if (gx > 1 && gx < sizeX && gy > 1 && gy < sizeY) {                      // :233-234
                    in = true;
                    v = mapCache[(size_t)(int)gy * cols + (int)gx];
                    isMax = v == 2;                                                      // :239
                }
            }
            const unsigned int inM = __ballot_sync(0xffffffffu, in), maxM = __ballot_sync(0xffffffffu, isMax);
            scanLen += __popc(inM); maxCount += __popc(maxM);
            unsigned int dm = inM & ~maxM;
            distCount += __popc(dm);
            while (dm) {                                                                 // :243, in ray order
                const int f = __ffs(dm) - 1;
                dm &= dm - 1;
                dist += __shfl_sync(0xffffffffu, v, f);
            }
        }
        if ((double)scanLen < (double)(size_t)nRays * 0.75) score = INFINITY;            // :248-249
        else score = (dist + 7 * (double)maxCount) / ((double)distCount + (double)maxCount) + 10 * ((double)nRays - (double)scanLen) / (double)nRays;
    }
    if (lane == 0) {
        double* o = out + (size_t)hyp * 15;
        o[0] = px; o[1] = py; o[2] = pang; o[3] = score;
        for (int q = 0; q < 4; q++) { o[4 + q] = mp[q]; o[8 + q] = sp[q]; }
        o[12] = (double)(unsigned)pr.x; o[13] = (double)(unsigned)pr.y; o[14] = (double)k;
    }
}

void lsdb_launch_fa_legacy(cudaStream_t s, int nPairs, const int2* pairs, const LsdbFaLine* scanLines, const LsdbFaLine* mapLines,
                           int lidarX, int lidarY, const double* mapCache, int cols, int rows, double resol, const double* ranges,
                           const double* angles, int nRays, double* out) {
    const int nHyp = 4 * nPairs;
    if (nHyp <= 0) return;
    const int warpsPerCta = FAL_THREADS / 32;
    lsdb_fa_legacy_kernel<<<(nHyp + warpsPerCta - 1) / warpsPerCta, FAL_THREADS, 0, s>>>(nHyp, pairs, scanLines, mapLines, lidarX, lidarY, mapCache,
                                                                                        cols, rows, resol, ranges, angles, nRays, out);
}
