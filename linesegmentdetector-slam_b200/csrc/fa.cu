// Association scoring on sm_100a: every (scan line, map line, endpoint pairing) hypothesis of a batch
// of scan frames is scored against mapCache in ONE launch — the replacement for myFA's pthread
// pool (reference /root/reference/LSD/myFA.cpp:27-63 dispatch, :186-272 thread_ScanToMapMatch,
// :274-305 NormalizedLineDirection, :307-355 rotateScanIm, :357-396 CalcScore).
//
// Two kernels.  lsdb_fa_pose_kernel, one THREAD per hypothesis: the pose algebra (two atand, cosd, sind, the lidar
// pose, the 60 px gate) from the same FMA-free double code as the host (lsd_math.h), so poses are bit-identical to
// the oracle; 32 hypotheses share every instruction of the ~2.5 k-instruction transcendental chain instead of one CTA
// repeating it in all of its threads.  lsdb_fa_score_kernel, one WARP per hypothesis: rotates the ~1.1 k scan points and
// gathers mapCache; the per-point sum is a tree reduction, i.e. equal to the reference's sequential sum only up to
// reassociation (~1e-13 relative; the contract is 1e-6).  mapCache (rows*cols*8 B, 4.7 MB for the bundled map) stays
// L2-resident; the stage is FP64-issue-bound, not HBM-bound.
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

struct FaPose { double cs, sn; int gate, pad; };

__device__ __forceinline__ void fa_endpoints(const LsdbFaLine& S, const LsdbFaLine& M, int i, double& msx, double& msy, double& mex,
                                             double& mey, double& ssx, double& ssy, double& sex, double& sey) {  // the four pairings, :194-235
    if (i <= 2) { msx = M.x1; msy = M.y1; mex = M.x2; mey = M.y2; } else { msx = M.x2; msy = M.y2; mex = M.x1; mey = M.y1; }
    if (i == 1 || i == 3) { ssx = S.x1; ssy = S.y1; sex = S.x2; sey = S.y2; } else { ssx = S.x2; ssy = S.y2; sex = S.x1; sey = S.y1; }
}

__global__ void __launch_bounds__(FA_THREADS) lsdb_fa_pose_kernel(int nHyp, const LsdbFaTask* __restrict__ tasks,
                                                                  const LsdbFaLine* __restrict__ scanLines, const int* __restrict__ scanLineOff,
                                                                  const int* __restrict__ ptOff, const double* __restrict__ lidarPose,
                                                                  const double* __restrict__ lastPose, const LsdbFaLine* __restrict__ mapLines,
                                                                  double pi, FaPose* __restrict__ pose, LsdbFaHyp* __restrict__ out) {
    const int hyp = blockIdx.x * FA_THREADS + threadIdx.x;
    if (hyp >= nHyp) return;
    const int t = hyp >> 2, i = (hyp & 3) + 1;
    const LsdbFaTask task = tasks[t];
    const LsdbFaLine S = scanLines[scanLineOff[task.frame] + task.iScan];
    const LsdbFaLine M = mapLines[task.iMap];
    double msx, msy, mex, mey, ssx, ssy, sex, sey;
    fa_endpoints(S, M, i, msx, msy, mex, mey, ssx, ssy, sex, sey);
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
    const int nPts = ptOff[task.frame + 1] - ptOff[task.frame];
    FaPose P; P.cs = cs; P.sn = sn; P.gate = gate ? 1 : 0; P.pad = 0;
    pose[hyp] = P;
    LsdbFaHyp h;
    h.frame = task.frame; h.iScan = task.iScan; h.iMap = task.iMap; h.iPair = i;
    h.x = 0; h.y = 0; h.ang = 0; h.score = INFINITY;
    if (gate && nPts != 0) {
        double a = angDiff;                                     // :339-343
        while (a <= -180) a += 360;
        while (a > 180) a -= 360;
        h.x = lx; h.y = ly; h.ang = a;
    }
    out[hyp] = h;
}

__global__ void __launch_bounds__(FA_THREADS) lsdb_fa_score_kernel(int nHyp, const LsdbFaTask* __restrict__ tasks,
                                                                   const LsdbFaLine* __restrict__ scanLines, const int* __restrict__ scanLineOff,
                                                                   const double* __restrict__ pts, const int* __restrict__ ptOff,
                                                                   const LsdbFaLine* __restrict__ mapLines, const double* __restrict__ mapCache,
                                                                   int cols, int rows, const FaPose* __restrict__ pose, LsdbFaHyp* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int hyp = blockIdx.x * (FA_THREADS / 32) + (threadIdx.x >> 5);
    if (hyp >= nHyp) return;
    const FaPose P = pose[hyp];
    const int t = hyp >> 2, i = (hyp & 3) + 1;
    const LsdbFaTask task = tasks[t];
    const int p0 = ptOff[task.frame], nPts = ptOff[task.frame + 1] - p0;
    if (!P.gate || nPts == 0) return;                               // score stays +inf
    const LsdbFaLine S = scanLines[scanLineOff[task.frame] + task.iScan];
    const LsdbFaLine M = mapLines[task.iMap];
    double msx, msy, mex, mey, ssx, ssy, sex, sey;
    fa_endpoints(S, M, i, msx, msy, mex, mey, ssx, ssy, sex, sey);
    const double cs = P.cs, sn = P.sn;
    double sum = 0.0;
    int nMax = 0, nInvalid = 0;                                  // nInvalid: warp-wide count, the same in every lane
    const double numAll = nPts, thr = 0.7 * numAll;             // :389-392: fewer than 70 % of the points inside the map -> +inf
    for (int base = 0; base < nPts; base += 32) {
        const int k = base + lane;
        bool outside = false;
        if (k < nPts) {
            const double2 pt = *reinterpret_cast<const double2*>(&pts[2 * (size_t)(p0 + k)]);
            const double ox = pt.x - ssx, oy = pt.y - ssy;                                               // :318-319
            const double rx = ox * cs - oy * sn + msx;                                                   // :334-335
            const double ry = ox * sn + oy * cs + msy;
            const int x = lsdb_x86_d2i(round(rx)), y = lsdb_x86_d2i(round(ry));                          // :367-368
            if (y >= 0 && y < rows && x >= 0 && x < cols) {
                const double v = mapCache[(size_t)y * cols + x];
                if (v >= 1.0) nMax++; else sum += v;                                                     // z_occ_max_dis, :373-382
            } else outside = true;
        }
        nInvalid += __popc(__ballot_sync(0xffffffffu, outside));
        // even if every remaining point fell inside, too few would: the score is +inf whatever the rest says (about half
        // of the hypotheses of an un-gated frame end here, after 30-60 % of their points)
        if ((double)(nPts - nInvalid) < thr) return;
    }
    const int nValid = nPts - nInvalid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        nMax += __shfl_xor_sync(0xffffffffu, nMax, o);
    }
    if (lane == 0) {
        const double numValid = nValid;
        const double sumMax = 10.0 * nMax;                      // sum of nMax tens is exact
        if (!(numValid < thr))                                  // :389-392
            out[hyp].score = (sum + sumMax) / numValid + 10 * (numAll - numValid) / numAll;
    }
}

// ------------------------------------------------------------------ per-frame reduction (LSD/myFA.cpp:65-171, the part before ukf)
// One CTA per frame: keep the hypotheses with score < 3 (:261), order them by score (CompScore :398-402; ties by
// (scan line, map line, pairing), i.e. a stable sort of the launch order), then the best one (first frame of a chain,
// :100-110) and the 1/score^2 weighted mean accumulated IN THAT ORDER by one thread (:160-171) — the same operations
// in the same order as the host tail of host/myFA_b200.cpp, so the estimate is bit-identical to it.
#define FA_RED_THREADS 256
#define FA_RED_CAP 2048
__global__ void __launch_bounds__(FA_RED_THREADS) lsdb_fa_reduce_kernel(const LsdbFaHyp* __restrict__ hyp, const int* __restrict__ hypOff,
                                                                        LsdbFaEst* __restrict__ est) {
    __shared__ double key[FA_RED_CAP];
    __shared__ int idx[FA_RED_CAP];
    __shared__ int nKept;
    const int f = blockIdx.x, a = hypOff[f], b = hypOff[f + 1];
    if (threadIdx.x == 0) nKept = 0;
    __syncthreads();
    for (int h = a + threadIdx.x; h < b; h += FA_RED_THREADS) {
        const double sc = hyp[h].score;
        if (sc < 3) {
            const int k = atomicAdd(&nKept, 1);
            if (k < FA_RED_CAP) { key[k] = sc; idx[k] = h; }
        }
    }
    __syncthreads();
    const int n = nKept;
    LsdbFaEst E;
    E.nHyp = b - a; E.nKept = n;
    E.bx = E.by = E.bang = 0; E.bscore = INFINITY; E.mx = E.my = E.mang = 0; E.mscore = INFINITY;
    if (n == 0 || n > FA_RED_CAP) {
        if (n > FA_RED_CAP) E.nKept = -n;   // too many for the shared-memory sort: the caller reduces on the host
        if (threadIdx.x == 0) est[f] = E;
        return;
    }
    int m = 1;
    while (m < n) m <<= 1;
    for (int k = n + threadIdx.x; k < m; k += FA_RED_THREADS) { key[k] = INFINITY; idx[k] = 0x7fffffff; }
    __syncthreads();
    for (int size = 2; size <= m; size <<= 1)          // bitonic sort on (score, launch index)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < m; t += FA_RED_THREADS) {
                const int p = t ^ stride;
                if (p > t) {
                    const bool up = (t & size) == 0;
                    const bool gt = key[t] > key[p] || (key[t] == key[p] && idx[t] > idx[p]);
                    if (gt == up) {
                        const double kk = key[t]; key[t] = key[p]; key[p] = kk;
                        const int ii = idx[t]; idx[t] = idx[p]; idx[p] = ii;
                    }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) {
        const LsdbFaHyp best = hyp[idx[0]];
        E.bx = best.x; E.by = best.y; E.bang = best.ang; E.bscore = best.score;
        double sumX = 0, sumY = 0, sumAng = 0, sumW = 0;
        for (int k = 0; k < n; k++) {
            const LsdbFaHyp h = hyp[idx[k]];
            const double w = 1 / (h.score * h.score);
            sumX += h.x * w; sumY += h.y * w; sumAng += h.ang * w; sumW += w;
        }
        E.mx = sumX / sumW; E.my = sumY / sumW; E.mang = sumAng / sumW;
        E.mscore = 1 / sqrt(sumW / n);
        est[f] = E;
    }
}

void lsdb_launch_fa_reduce(cudaStream_t s, int nFrames, const LsdbFaHyp* hyp, const int* hypOff, LsdbFaEst* est) {
    if (nFrames > 0) lsdb_fa_reduce_kernel<<<nFrames, FA_RED_THREADS, 0, s>>>(hyp, hypOff, est);
}

size_t lsdb_fa_pose_bytes(int nTasks) { return sizeof(FaPose) * (size_t)nTasks * 4; }

void lsdb_launch_fa(cudaStream_t s, int nTasks, const LsdbFaTask* tasks, const LsdbFaLine* scanLines, const int* scanLineOff,
                    const double* pts, const int* ptOff, const double* lidarPose, const double* lastPose,
                    const LsdbFaLine* mapLines, const double* mapCache, int cols, int rows, double pi, LsdbFaHyp* out, void* poseBuf) {
    if (nTasks <= 0) return;
    const int nHyp = nTasks * 4;
    FaPose* pose = reinterpret_cast<FaPose*>(poseBuf);
    lsdb_fa_pose_kernel<<<(nHyp + FA_THREADS - 1) / FA_THREADS, FA_THREADS, 0, s>>>(nHyp, tasks, scanLines, scanLineOff, ptOff, lidarPose, lastPose,
                                                                                   mapLines, pi, pose, out);
    const int perCta = FA_THREADS / 32;
    lsdb_fa_score_kernel<<<(nHyp + perCta - 1) / perCta, FA_THREADS, 0, s>>>(nHyp, tasks, scanLines, scanLineOff, pts, ptOff, mapLines, mapCache,
                                                                            cols, rows, pose, out);
}

// ---- the pair filter of LSD/myFA.cpp:29-41 on the device (used when the scan lines are already there) ----
// count: one thread per scan line; lines shorter than ignoreScanLength = 40 pair with nothing, the others with every map
// line whose length is within scanToMapDiff = 0.35 of theirs (LSD/baseFunc.h:80-82)
__device__ __forceinline__ bool fa_pair_ok(double lenS, double lenDiff, double lenM) { return !(lenM < lenS - lenDiff || lenM > lenS + lenDiff); }

__global__ void lsdb_fa_pair_count_kernel(int nL, const LsdbFaLine* __restrict__ scanLines, const LsdbFaLine* __restrict__ mapLines, int nMap,
                                          int* __restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nL) return;
    const double lenS = scanLines[i].len;
    int c = 0;
    if (!(lenS < 40)) {
        const double lenDiff = lenS * 0.35;
        for (int im = 0; im < nMap; im++) c += fa_pair_ok(lenS, lenDiff, mapLines[im].len) ? 1 : 0;
    }
    cnt[i] = c;
}

// exclusive prefix sum of n ints in three small launches (tiles of 1024, their sums, the add-back); out[n] = total
#define FA_SCAN_TILE 1024
__global__ void __launch_bounds__(FA_SCAN_TILE) lsdb_scan_tiles_kernel(int n, const int* in, int* out, int* __restrict__ tileSum) {   // in may alias out
    __shared__ int wsum[32];
    const int i = blockIdx.x * FA_SCAN_TILE + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = i < n ? in[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        wsum[lane] = w;
    }
    __syncthreads();
    const int incl = x + (warp ? wsum[warp - 1] : 0);
    if (i < n) out[i] = incl - v;
    if (threadIdx.x == FA_SCAN_TILE - 1) tileSum[blockIdx.x] = incl;
}
__global__ void __launch_bounds__(FA_SCAN_TILE) lsdb_scan_sums_kernel(int nTiles, int* __restrict__ tileSum, int* __restrict__ total) {
    __shared__ int wsum[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nTiles; base += FA_SCAN_TILE) {
        const int i = base + threadIdx.x;
        const int v = i < nTiles ? tileSum[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            wsum[lane] = w;
        }
        __syncthreads();
        const int incl = x + (warp ? wsum[warp - 1] : 0) + carry;
        if (i < nTiles) tileSum[i] = incl - v;      // exclusive offset of the tile
        __syncthreads();
        if (threadIdx.x == FA_SCAN_TILE - 1) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void lsdb_scan_add_kernel(int n, int* __restrict__ out, const int* __restrict__ tileOff, const int* __restrict__ total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tileOff[i / FA_SCAN_TILE];
    if (i == 0) out[n] = *total;
}

// ---- ordered compaction of the hypotheses the reference keeps (score < 3, LSD/myFA.cpp:261-265) ----
__global__ void lsdb_fa_keep_flag_kernel(int nHyp, const LsdbFaHyp* __restrict__ hyp, double below, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nHyp) flag[i] = hyp[i].score < below ? 1 : 0;
}
__global__ void lsdb_fa_keep_write_kernel(int nHyp, const LsdbFaHyp* __restrict__ hyp, double below, const int* __restrict__ off,
                                          LsdbFaHyp* __restrict__ out, int cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nHyp && hyp[i].score < below && off[i] < cap) out[off[i]] = hyp[i];
}

// tasks in (frame, scan line, map line) order at their offsets; hypOff[f] = 4 * tasks before frame f
__global__ void lsdb_fa_pair_write_kernel(int nL, int nFrames, const LsdbFaLine* __restrict__ scanLines, const int* __restrict__ lineOff,
                                          const LsdbFaLine* __restrict__ mapLines, int nMap, const int* __restrict__ taskOff,
                                          LsdbFaTask* __restrict__ tasks, int* __restrict__ hypOff) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nFrames) hypOff[i] = 4 * taskOff[lineOff[i]];
    if (i >= nL) return;
    const int n = taskOff[i + 1] - taskOff[i];
    if (n == 0) return;
    int lo = 0, hi = nFrames;                                // frame of scan line i: lineOff[f] <= i < lineOff[f+1]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (lineOff[mid] <= i) lo = mid; else hi = mid; }
    const double lenS = scanLines[i].len, lenDiff = lenS * 0.35;
    LsdbFaTask* t = tasks + taskOff[i];
    int k = 0;
    for (int im = 0; im < nMap; im++)
        if (fa_pair_ok(lenS, lenDiff, mapLines[im].len)) { LsdbFaTask o = {lo, i - lineOff[lo], im, 0}; t[k++] = o; }
}

// scratch ints needed by lsdb_launch_fa_pairs_count: cnt[nL] is reused as taskOff[nL+1], then tile sums and the total
size_t lsdb_fa_pairs_scratch_ints(int nL) { return (size_t)nL + 1 + (size_t)(nL + FA_SCAN_TILE - 1) / FA_SCAN_TILE + 2; }

// phase 1: taskOff[0..nL] (exclusive prefix of the pair counts; taskOff[nL] = number of tasks)
void lsdb_launch_fa_pairs_count(cudaStream_t s, int nL, const LsdbFaLine* scanLines, const LsdbFaLine* mapLines, int nMap, int* scratch) {
    if (nL <= 0) return;
    const int nTiles = (nL + FA_SCAN_TILE - 1) / FA_SCAN_TILE;
    int* taskOff = scratch; int* tileSum = scratch + nL + 1; int* total = tileSum + nTiles;
    lsdb_fa_pair_count_kernel<<<(nL + 255) / 256, 256, 0, s>>>(nL, scanLines, mapLines, nMap, taskOff);
    lsdb_scan_tiles_kernel<<<nTiles, FA_SCAN_TILE, 0, s>>>(nL, taskOff, taskOff, tileSum);
    lsdb_scan_sums_kernel<<<1, FA_SCAN_TILE, 0, s>>>(nTiles, tileSum, total);
    lsdb_scan_add_kernel<<<(nL + 255) / 256, 256, 0, s>>>(nL, taskOff, tileSum, total);
}
// scratch ints of lsdb_launch_fa_keep: off[nHyp + 1], tile sums, total
size_t lsdb_fa_keep_scratch_ints(int nHyp) { return (size_t)nHyp + 1 + (size_t)(nHyp + FA_SCAN_TILE - 1) / FA_SCAN_TILE + 2; }
// kept hypotheses (score < below) in launch order -> out[0..cap); scratch[nHyp] = their number
void lsdb_launch_fa_keep(cudaStream_t s, int nHyp, const LsdbFaHyp* hyp, double below, int* scratch, LsdbFaHyp* out, int cap) {
    if (nHyp <= 0) return;
    const int nTiles = (nHyp + FA_SCAN_TILE - 1) / FA_SCAN_TILE;
    int* off = scratch; int* tileSum = scratch + nHyp + 1; int* total = tileSum + nTiles;
    lsdb_fa_keep_flag_kernel<<<(nHyp + 255) / 256, 256, 0, s>>>(nHyp, hyp, below, off);
    lsdb_scan_tiles_kernel<<<nTiles, FA_SCAN_TILE, 0, s>>>(nHyp, off, off, tileSum);
    lsdb_scan_sums_kernel<<<1, FA_SCAN_TILE, 0, s>>>(nTiles, tileSum, total);
    lsdb_scan_add_kernel<<<(nHyp + 255) / 256, 256, 0, s>>>(nHyp, off, tileSum, total);
    lsdb_fa_keep_write_kernel<<<(nHyp + 255) / 256, 256, 0, s>>>(nHyp, hyp, below, off, out, cap);
}

// phase 2: the task list and the per-frame hypothesis offsets
void lsdb_launch_fa_pairs_write(cudaStream_t s, int nL, int nFrames, const LsdbFaLine* scanLines, const int* lineOff, const LsdbFaLine* mapLines,
                                int nMap, const int* scratch, LsdbFaTask* tasks, int* hypOff) {
    const int n = (nL > nFrames + 1 ? nL : nFrames + 1);
    lsdb_fa_pair_write_kernel<<<(n + 255) / 256, 256, 0, s>>>(nL, nFrames, scanLines, lineOff, mapLines, nMap, scratch, tasks, hypOff);
}
