// Association scoring on sm_100a: every (scan line, map line, endpoint pairing) hypothesis of a batch
// of scan frames is scored against mapCache in ONE launch — the replacement for myFA's pthread
// pool (reference /root/reference/LSD/myFA.cpp:27-63 dispatch, :186-272 thread_ScanToMapMatch,
// :274-305 NormalizedLineDirection, :307-355 rotateScanIm, :357-396 CalcScore).
//
// One CTA per hypothesis.  The pose algebra is evaluated redundantly by every thread from the same
// FMA-free double code as the host (lsd_math.h), so poses are bit-identical to the oracle; the
// per-point sum of mapCache values is a tree reduction, i.e. equal to the reference's sequential
// sum only up to reassociation (~1e-13 relative; the contract is 1e-6).  mapCache (rows*cols*8 B,
// 4.7 MB for the bundled map) stays L2-resident; the stage is gather/issue-bound, not HBM-bound.
#include "lsdb_common.cuh"

#define FA_THREADS 128

__device__ double fa_norm_line_dir(double staX, double staY, double endX, double endY, double pi) {  // :274-305
    double angle;
    if (staX == endX && staY != endY) angle = staY < endY ? 90 : -90;
    else if (staX != endX && staY == endY) angle = staX < endX ? 0 : 180;
    else angle = lsdm_atan((endY - staY) / (endX - staX)) * 180.0 / pi;  // atand, LSD/baseFunc.cpp:14-16
    if (angle < 0 && staX > endX) return angle + 180;
    if (angle > 0 && staX > endX) return angle - 180;
    return angle;
}

__global__ void __launch_bounds__(FA_THREADS) lsdb_fa_kernel(const LsdbFaTask* __restrict__ tasks, const LsdbFaLine* __restrict__ scanLines,
                                                             const int* __restrict__ scanLineOff, const double* __restrict__ pts,
                                                             const int* __restrict__ ptOff, const double* __restrict__ lidarPose,
                                                             const double* __restrict__ lastPose, const LsdbFaLine* __restrict__ mapLines,
                                                             const double* __restrict__ mapCache, int cols, int rows, double pi,
                                                             LsdbFaHyp* __restrict__ out) {
    __shared__ double sSum[FA_THREADS / 32];
    __shared__ int sValid[FA_THREADS / 32], sMax[FA_THREADS / 32];
    const int hyp = blockIdx.x, t = hyp >> 2, i = (hyp & 3) + 1;
    const LsdbFaTask task = tasks[t];
    const LsdbFaLine S = scanLines[scanLineOff[task.frame] + task.iScan];
    const LsdbFaLine M = mapLines[task.iMap];
    double msx, msy, mex, mey, ssx, ssy, sex, sey;  // the four pairings, :194-235
    if (i <= 2) { msx = M.x1; msy = M.y1; mex = M.x2; mey = M.y2; } else { msx = M.x2; msy = M.y2; mex = M.x1; mey = M.y1; }
    if (i == 1 || i == 3) { ssx = S.x1; ssy = S.y1; sex = S.x2; sey = S.y2; } else { ssx = S.x2; ssy = S.y2; sex = S.x1; sey = S.y1; }
    const double mapAng = fa_norm_line_dir(msx, msy, mex, mey, pi);
    const double scanAng = fa_norm_line_dir(ssx, ssy, sex, sey, pi);
    const double angDiff = mapAng - scanAng;                       // :310
    const double cs = lsdm_cos(angDiff / 180.0 * pi), sn = lsdm_sin(angDiff / 180.0 * pi);  // cosd/sind
    const double lpx = lidarPose[2 * task.frame], lpy = lidarPose[2 * task.frame + 1];
    const double lax = lastPose[3 * task.frame], lay = lastPose[3 * task.frame + 1];
    const double lx = (lpx - ssx) * cs - (lpy - ssy) * sn + msx;   // :324-325
    const double ly = (lpx - ssx) * sn + (lpy - ssy) * cs + msy;
    const double ddx = lx - lax, ddy = ly - lay;
    const bool gate = sqrt(ddx * ddx + ddy * ddy) < 60 || lax == -1;  // maxEstiDist, :330
    const int p0 = ptOff[task.frame], nPts = ptOff[task.frame + 1] - p0;

    double sum = 0.0;
    int nValid = 0, nMax = 0;
    if (gate) {
        for (int k = threadIdx.x; k < nPts; k += FA_THREADS) {
            const double ox = pts[2 * (size_t)(p0 + k)] - ssx, oy = pts[2 * (size_t)(p0 + k) + 1] - ssy;  // :318-319
            const double rx = ox * cs - oy * sn + msx;                                                   // :334-335
            const double ry = ox * sn + oy * cs + msy;
            const int x = lsdb_x86_d2i(round(rx)), y = lsdb_x86_d2i(round(ry));                          // :367-368
            if (y >= 0 && y < rows && x >= 0 && x < cols) {
                nValid++;
                const double v = mapCache[(size_t)y * cols + x];
                if (v >= 1.0) nMax++; else sum += v;                                                     // z_occ_max_dis, :373-382
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        nValid += __shfl_xor_sync(0xffffffffu, nValid, o);
        nMax += __shfl_xor_sync(0xffffffffu, nMax, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sSum[w] = sum; sValid[w] = nValid; sMax[w] = nMax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < FA_THREADS / 32; k++) { sum += sSum[k]; nValid += sValid[k]; nMax += sMax[k]; }
        LsdbFaHyp h;
        h.frame = task.frame; h.iScan = task.iScan; h.iMap = task.iMap; h.iPair = i;
        h.x = 0; h.y = 0; h.ang = 0; h.score = INFINITY;
        if (gate && nPts != 0) {
            double a = angDiff;                                     // :339-343
            while (a <= -180) a += 360;
            while (a > 180) a -= 360;
            h.x = lx; h.y = ly; h.ang = a;
            const double numAll = nPts, numValid = nValid;
            const double sumMax = 10.0 * nMax;                      // sum of nMax tens is exact
            if (!(numValid < 0.7 * numAll))                         // :389-392
                h.score = (sum + sumMax) / numValid + 10 * (numAll - numValid) / numAll;
        }
        out[hyp] = h;
    }
}

void lsdb_launch_fa(cudaStream_t s, int nTasks, const LsdbFaTask* tasks, const LsdbFaLine* scanLines, const int* scanLineOff,
                    const double* pts, const int* ptOff, const double* lidarPose, const double* lastPose,
                    const LsdbFaLine* mapLines, const double* mapCache, int cols, int rows, double pi, LsdbFaHyp* out) {
    if (nTasks > 0)
        lsdb_fa_kernel<<<nTasks * 4, FA_THREADS, 0, s>>>(tasks, scanLines, scanLineOff, pts, ptOff, lidarPose, lastPose, mapLines,
                                                         mapCache, cols, rows, pi, out);
}
