// Scan front-end of the association path: myrdp::FeatureScan for a batch of lidar frames
// (LSD/myRDP.cpp:9-185; RegionSegmentation :274-352, SplitMerge/SplitMergeAssistant :187-272,
// getThresholdDeltaDist :354-375).  One CTA per frame.  The frame's beams live in shared memory; the
// steps whose result depends on order (cluster boundaries, the list of line pieces, the order of the
// raster samples) are kept in the reference's order, everything else is spread over the threads:
//   beams -> metric / grid coordinates (correctly rounded cos/sin, computed once instead of three times)
//   break flags -> clusters (one thread; <= n steps of a 1-byte read)
//   Ramer-Douglas-Peucker per cluster: warp-wide arg-max with the reference's first-maximum rule and an
//     explicit stack (the recursion only sets split flags, so the visiting order is free)
//   line pieces -> ordered compaction by ballot -> one warp per line: sample count, then (second pass)
//     records, raster samples and the 0/255 scan raster at their final offsets.
// The kernel runs twice: COUNT fills lsdb_scan_info so that the host can lay out the ragged outputs,
// WRITE recomputes the frame (a few microseconds) and stores them.  -fmad=false like the rest of the library.
#include "lsdb_common.cuh"
#include <limits.h>

#define FS_NT 128
#define FS_NW (FS_NT / 32)

struct FsShared {
    double red[4][FS_NW];
    double minX, minY, maxX, maxY;
    int nc, nSeg, nLines, nPts, overflow;
};

__device__ __forceinline__ double fs_delta_thre(double r) {  // LSD/myRDP.cpp:354-375
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

size_t lsdb_fscan_smem(int maxBeams) {
    const size_t cap = (size_t)maxBeams + 1;
    return sizeof(FsShared) + 8 * 4 * cap + 4 * (2 * (cap + 1) + 2 * (cap + 2) + 2 * (2 * cap + 2) + (2 * cap + 2)) + 2 * cap + 16;
}

// geometry of one kept line piece (grid coordinates relative to the frame's minimum), LSD/myRDP.cpp:75-131
struct FsLine {
    double x1, y1, x2, y2, k;
    int xLow, yLow, cnt, alongX;
};
__device__ __forceinline__ FsLine fs_line(double ax, double ay, double bx, double by, double minX, double minY) {
    FsLine L;
    L.x1 = ax - minX; L.y1 = ay - minY; L.x2 = bx - minX; L.y2 = by - minY;
    L.k = (L.y2 - L.y1) / (L.x2 - L.x1);
    const int xLow = lsdb_x86_d2i(floor(L.x1 > L.x2 ? L.x2 : L.x1)), xHigh = lsdb_x86_d2i(ceil(L.x1 > L.x2 ? L.x1 : L.x2));
    const int yLow = lsdb_x86_d2i(floor(L.y1 > L.y2 ? L.y2 : L.y1)), yHigh = lsdb_x86_d2i(ceil(L.y1 > L.y2 ? L.y1 : L.y2));
    const int xl = xHigh - xLow + 1, yl = yHigh - yLow + 1;
    L.xLow = xLow; L.yLow = yLow;
    L.cnt = xl > yl ? xl : yl;                               // emission loop bound, :132-153
    L.alongX = fabs(L.x2 - L.x1) > fabs(L.y2 - L.y1);        // which axis is walked, :104
    return L;
}
// sample m of a line: true when it is stored (inside the raster and off row/column 0), :107-139
__device__ __forceinline__ bool fs_sample(const FsLine& L, int m, int W, int H, int& xx, int& yy) {
    if (L.alongX) { xx = m + L.xLow; yy = lsdb_x86_d2i(round(((double)xx - L.x1) * L.k + L.y1)); }
    else { yy = m + L.yLow; xx = lsdb_x86_d2i(round(((double)yy - L.y1) / L.k + L.x1)); }
    return !(xx < 0 || xx >= W || yy < 0 || yy >= H) && xx != 0 && yy != 0;
}

template <bool WRITE>
__global__ void __launch_bounds__(FS_NT) lsdb_fscan_kernel(
    int maxBeams, const double* __restrict__ ranges, const double* __restrict__ angles, const int* __restrict__ beamOff,
    double resol, double oriX, double oriY, int leastPoint, double threLine, double leastDistM, double pi,
    LsdbFsInfo* __restrict__ info, const int* __restrict__ lineOff, const int* __restrict__ ptOff,
    const long long* __restrict__ imOff, LsdbFaLine* __restrict__ lines, double* __restrict__ pts, uint8_t* __restrict__ lineIm) {
    extern __shared__ __align__(16) unsigned char fsRaw[];
    FsShared& sh = *(FsShared*)fsRaw;
    const int cap = maxBeams + 1;
    double* px = (double*)(fsRaw + sizeof(FsShared));
    double* py = px + cap; double* gx = py + cap; double* gy = gx + cap;
    int* cell = (int*)(gy + cap);            // 2*(cap+1): start, end beam of each cluster
    int* stack = cell + 2 * (cap + 1);       // 2*(cap+2)
    int* seg = stack + 2 * (cap + 2);        // 2*(2*cap+2): beam pairs of the line pieces, compacted in place
    int* cnt = seg + 2 * (2 * cap + 2);      // 2*cap+2: samples per kept line, then their exclusive prefix
    uint8_t* brk = (uint8_t*)(cnt + 2 * cap + 2);
    uint8_t* split = brk + cap;
    const int segCap = 2 * cap + 2;

    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = beamOff[f + 1] - beamOff[f];
    const double* R = ranges + beamOff[f];
    const double* A = angles + beamOff[f];
    const double pose0 = 0.0, pose1 = 0.0, pose2 = 0.0;      // scanPose, :11

    // ---- beams -> coordinates (:21-34, :196-201, :289-293) and the raster extent ----
    double mnX = INFINITY, mnY = INFINITY, mxX = 0, mxY = 0;
    for (int i = tid; i < n; i += FS_NT) {
        const double r = R[i], a = A[i] + pose2;
        const double x = r * lsdm_cos(a) + pose0, y = r * lsdm_sin(a) + pose1;
        px[i] = x; py[i] = y;
        const double X = floor((x - oriX) / resol), Y = floor((y - oriY) / resol);
        gx[i] = X; gy[i] = Y;
        if (X < mnX) mnX = X;
        if (X > mxX) mxX = X;
        if (Y < mnY) mnY = Y;
        if (Y > mxY) mxY = Y;
        split[i] = 0;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        double t;
        t = __shfl_xor_sync(0xffffffffu, mnX, o); if (t < mnX) mnX = t;
        t = __shfl_xor_sync(0xffffffffu, mxX, o); if (t > mxX) mxX = t;
        t = __shfl_xor_sync(0xffffffffu, mnY, o); if (t < mnY) mnY = t;
        t = __shfl_xor_sync(0xffffffffu, mxY, o); if (t > mxY) mxY = t;
    }
    if (lane == 0) { sh.red[0][warp] = mnX; sh.red[1][warp] = mxX; sh.red[2][warp] = mnY; sh.red[3][warp] = mxY; }
    __syncthreads();
    if (tid == 0) {
        double a = sh.red[0][0], b = sh.red[1][0], c = sh.red[2][0], d = sh.red[3][0];
        for (int w = 1; w < FS_NW; w++) {
            if (sh.red[0][w] < a) a = sh.red[0][w];
            if (sh.red[1][w] > b) b = sh.red[1][w];
            if (sh.red[2][w] < c) c = sh.red[2][w];
            if (sh.red[3][w] > d) d = sh.red[3][w];
        }
        sh.minX = a; sh.maxX = b; sh.minY = c; sh.maxY = d; sh.overflow = 0;
    }
    // ---- break flags (:304-337): 1 = gap after beam i, 2 = last beam joins the first ----
    for (int i = tid; i < n; i += FS_NT) {
        const int j = i == n - 1 ? 0 : i + 1;
        const double dx = px[i] - px[j], dy = py[i] - py[j];
        const double dd = sqrt(dx * dx + dy * dy);
        const double thre = fs_delta_thre(R[i]);
        brk[i] = dd > thre ? 1 : (dd <= thre && i == n - 1 ? 2 : 0);
    }
    __syncthreads();
    // ---- clusters, in beam order (:304-346) ----
    if (tid == 0) {
        int startNum = 0, nc = 0;
        for (int i = 0; i < n; i++) {
            const int b = brk[i];
            if (b == 1) {
                cell[2 * nc] = startNum; cell[2 * nc + 1] = i;
                if (abs(i - startNum) >= leastPoint) nc++;
                startNum = i + 1;
            } else if (b == 2) cell[0] = startNum;
        }
        sh.nc = nc;
    }
    __syncthreads();
    const int nc = sh.nc;
    const double minX = sh.minX, minY = sh.minY;

    if (warp == 0) {
        // ---- Ramer-Douglas-Peucker (:218-272) ----
        for (int c = 0; c < nc; c++) {
            int sp = 0;
            if (lane == 0) { stack[0] = cell[2 * c]; stack[1] = cell[2 * c + 1]; }
            sp = 2;
            __syncwarp();
            while (sp) {
                const int e = stack[sp - 1], s = stack[sp - 2];
                sp -= 2;
                __syncwarp();
                const int len = e > s ? e - s + 1 : n + e - s + 1;
                if (len <= 2) continue;
                const double k = (py[e] - py[s]) / (px[e] - px[s]);
                const double d = py[e] - k * px[e];
                const double den = sqrt(k * k + 1);
                double best = 0; int bestPos = INT_MAX;
                for (int i = 1 + lane; i < len - 1; i += 32) {
                    int a = s + i; if (a >= n) a -= n;
                    const double dist = fabs(k * px[a] - py[a] + d) / den;
                    if (dist > best) { best = dist; bestPos = i; }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int op = __shfl_xor_sync(0xffffffffu, bestPos, o);
                    if (ob > best || (ob == best && op < bestPos)) { best = ob; bestPos = op; }
                }
                int iMax = 0;
                if (bestPos != INT_MAX) { iMax = s + bestPos; if (iMax >= n) iMax -= n; }
                const double rr = R[iMax];
                const double thre = rr > 9 ? rr * threLine : threLine;
                if (best > thre) {
                    if (lane == 0) {
                        split[iMax] = 1;
                        stack[sp] = s; stack[sp + 1] = iMax; stack[sp + 2] = iMax; stack[sp + 3] = e;
                    }
                    sp += 4;
                }
                __syncwarp();
            }
        }
        __syncwarp();
        // ---- line pieces in the reference's order (:49-74) ----
        if (lane == 0) {
            int ns = 0;
            for (int c = 0; c < nc; c++) {
                const int s = cell[2 * c], e = cell[2 * c + 1];
                const int len = e > s ? e - s + 1 : n + e - s + 1;
                int a = s;
                for (int j = 0; j <= len; j++) {
                    int b;
                    if (j < len) { b = s + j; if (b >= n) b -= n; if (!split[b]) continue; }
                    else b = e;
                    if (ns < segCap) { seg[2 * ns] = a; seg[2 * ns + 1] = b; }
                    ns++;
                    a = b;
                }
            }
            if (ns > segCap) { sh.overflow = 1; ns = segCap; }
            sh.nSeg = ns;
        }
        __syncwarp();
        // ---- keep pieces of at least lineDistThre (:80-81), order preserved ----
        const int nSeg = sh.nSeg;
        const double distThre = leastDistM / resol;
        int nl = 0;
        for (int base = 0; base < nSeg; base += 32) {
            const int i = base + lane;
            int a = 0, b = 0; bool keep = false;
            if (i < nSeg) {
                a = seg[2 * i]; b = seg[2 * i + 1];
                const double ddx = gx[a] - gx[b], ddy = gy[a] - gy[b];
                keep = sqrt(ddx * ddx + ddy * ddy) >= distThre;
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            __syncwarp();
            if (keep) { const int o = nl + __popc(m & ((1u << lane) - 1)); seg[2 * o] = a; seg[2 * o + 1] = b; }
            nl += __popc(m);
            __syncwarp();
        }
        if (lane == 0) sh.nLines = nl;
    }
    __syncthreads();
    const int nLines = sh.nLines;
    const int W = lsdb_x86_d2i(ceil(sh.maxX - minX)), H = lsdb_x86_d2i(ceil(sh.maxY - minY));

    // ---- samples per line ----
    for (int L = warp; L < nLines; L += FS_NW) {
        const int a = seg[2 * L], b = seg[2 * L + 1];
        const FsLine g = fs_line(gx[a], gy[a], gx[b], gy[b], minX, minY);
        int c = 0;
        for (int m0 = 0; m0 < g.cnt; m0 += 32) {
            int xx, yy;
            const bool ok = m0 + lane < g.cnt && fs_sample(g, m0 + lane, W, H, xx, yy);
            c += __popc(__ballot_sync(0xffffffffu, ok));
        }
        if (lane == 0) cnt[L] = c;
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int L = 0; L < nLines; L++) { const int c = cnt[L]; cnt[L] = run; run += c; }
        sh.nPts = run;
    }
    __syncthreads();

    if (!WRITE) {
        if (tid == 0) {
            LsdbFsInfo o;
            o.nLines = sh.overflow ? -1 : nLines; o.nPts = sh.nPts; o.W = W; o.H = H;
            o.lidarX = floor((pose0 - oriX) / resol - minX);   // :38-40
            o.lidarY = floor((pose1 - oriY) / resol - minY);
            info[f] = o;
        }
        return;
    }
    // ---- records, samples and raster at their final places (:82-174) ----
    LsdbFaLine* outL = lines + lineOff[f];
    double* outP = pts + 2 * (size_t)ptOff[f];
    uint8_t* im = lineIm ? lineIm + imOff[f] : 0;
    for (int L = warp; L < nLines; L += FS_NW) {
        const int a = seg[2 * L], b = seg[2 * L + 1];
        const FsLine g = fs_line(gx[a], gy[a], gx[b], gy[b], minX, minY);
        if (lane == 0) {
            double ang = lsdm_atan(g.k) * 180.0 / pi;      // atand, LSD/baseFunc.cpp:14-16
            int orient = 1;
            if (ang < 0) { ang += 180; orient = -1; }
            LsdbFaLine o;
            o.k = g.k;
            o.b = (g.y1 + g.y2) / 2.0 - g.k * (g.x1 + g.x2) / 2.0;
            o.dx = lsdm_cos(ang / 180.0 * pi); o.dy = lsdm_sin(ang / 180.0 * pi);   // cosd / sind, :6-12
            o.x1 = g.x1; o.y1 = g.y1; o.x2 = g.x2; o.y2 = g.y2;
            const double ey = g.y2 - g.y1, ex = g.x2 - g.x1;
            o.len = sqrt(ey * ey + ex * ex);
            o.orient = orient; o.pad = 0;
            outL[L] = o;
        }
        int run = cnt[L];
        for (int m0 = 0; m0 < g.cnt; m0 += 32) {
            int xx = 0, yy = 0;
            const bool ok = m0 + lane < g.cnt && fs_sample(g, m0 + lane, W, H, xx, yy);
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const size_t o = (size_t)(run + __popc(m & ((1u << lane) - 1)));
                outP[2 * o] = (double)xx; outP[2 * o + 1] = (double)yy;
                if (im) im[(size_t)yy * W + xx] = 255;
            }
            run += __popc(m);
        }
    }
}

int lsdb_launch_fscan(cudaStream_t s, int pass, int nFrames, int maxBeams, const double* ranges, const double* angles, const int* beamOff,
                      double resol, double oriX, double oriY, int leastPoint, double threLine, double leastDistM, double pi,
                      LsdbFsInfo* info, const int* lineOff, const int* ptOff, const long long* imOff, LsdbFaLine* lines, double* pts,
                      uint8_t* lineIm) {
    const size_t smem = lsdb_fscan_smem(maxBeams);
    cudaError_t e;
    if (pass == 0) {
        if ((e = cudaFuncSetAttribute(lsdb_fscan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        lsdb_fscan_kernel<false><<<nFrames, FS_NT, smem, s>>>(maxBeams, ranges, angles, beamOff, resol, oriX, oriY, leastPoint, threLine,
                                                               leastDistM, pi, info, lineOff, ptOff, imOff, lines, pts, lineIm);
    } else {
        if ((e = cudaFuncSetAttribute(lsdb_fscan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        lsdb_fscan_kernel<true><<<nFrames, FS_NT, smem, s>>>(maxBeams, ranges, angles, beamOff, resol, oriX, oriY, leastPoint, threLine,
                                                              leastDistM, pi, info, lineOff, ptOff, imOff, lines, pts, lineIm);
    }
    return (int)cudaGetLastError();
}
