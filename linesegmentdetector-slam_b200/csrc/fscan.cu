// Scan front-end of the association path: myrdp::FeatureScan for a batch of lidar frames
// (LSD/myRDP.cpp:9-185; RegionSegmentation :274-352, SplitMerge/SplitMergeAssistant :187-272,
// getThresholdDeltaDist :354-375).  Two kernels:
//   lsdb_fscan_frames_kernel — one CTA per frame, the frame's beams in shared memory.  The steps whose result
//     depends on order (cluster boundaries, the list of line pieces) keep the reference's order, everything else
//     is spread over the threads:
//       beams -> metric coordinates (correctly rounded cos/sin, computed once instead of three times)
//       break flags -> clusters (one thread; <= n steps of a 1-byte read)
//       Ramer-Douglas-Peucker, one warp per cluster: warp-wide arg-max with the reference's first-maximum rule and
//         an explicit stack (the recursion only sets split flags, so the visiting order is free).  The distance is a
//         numerator over a per-call constant; division by a positive constant is monotone, so the maximum quotient
//         is the quotient of the maximum numerator and only numerators within 2^-48 of it can tie with it: one
//         division per call instead of one per point, same winner.
//       grid coordinates, raster extent, line pieces -> ordered compaction by ballot -> samples per kept line
//     Output: lsdb_scan_info and, per kept line, its end points and the offset of its samples inside the frame.
//   lsdb_fscan_lines_kernel — one warp per kept line of the whole batch, after the host has laid out the ragged
//     outputs: the structLinesInfo record, the raster samples in order, the 0/255 raster.
// -fmad=false like the rest of the library.
#include "lsdb_common.cuh"
#include <limits.h>

#define FS_NT 64
#define FS_NW (FS_NT / 32)
#define FS_LW 8                     // warps (= lines) per CTA of the lines kernel

struct FsShared {
    double red[4][FS_NW];
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

// shared memory of one frame: coordinates 2 x f64, clusters 2 x i32, one RDP stack per warp (the first doubles as
// the line-piece list afterwards) 2 x i32 each, samples per line i32, break + split flags 2 x u8
size_t lsdb_fscan_smem(int maxBeams) {
    const size_t cap = (size_t)maxBeams + 1;
    return sizeof(FsShared) + 8 * 2 * cap + 4 * (2 * (cap + 1) + FS_NW * 2 * (cap + 2) + (cap + 2)) + 2 * cap + 16;
}

// geometry of one kept line piece (grid coordinates relative to the frame's minimum), LSD/myRDP.cpp:75-131
struct FsLine {
    double x1, y1, x2, y2, k;
    int xLow, yLow, cnt, alongX;
};
__device__ __forceinline__ FsLine fs_line(double x1, double y1, double x2, double y2) {
    FsLine L;
    L.x1 = x1; L.y1 = y1; L.x2 = x2; L.y2 = y2;
    L.k = (y2 - y1) / (x2 - x1);
    const int xLow = lsdb_x86_d2i(floor(x1 > x2 ? x2 : x1)), xHigh = lsdb_x86_d2i(ceil(x1 > x2 ? x1 : x2));
    const int yLow = lsdb_x86_d2i(floor(y1 > y2 ? y2 : y1)), yHigh = lsdb_x86_d2i(ceil(y1 > y2 ? y1 : y2));
    const int xl = xHigh - xLow + 1, yl = yHigh - yLow + 1;
    L.xLow = xLow; L.yLow = yLow;
    L.cnt = xl > yl ? xl : yl;                               // emission loop bound, :132-153
    L.alongX = fabs(x2 - x1) > fabs(y2 - y1);                // which axis is walked, :104
    return L;
}
// sample m of a line: true when it is stored (inside the raster and off row/column 0), :107-139
__device__ __forceinline__ bool fs_sample(const FsLine& L, int m, int W, int H, int& xx, int& yy) {
    if (L.alongX) { xx = m + L.xLow; yy = lsdb_x86_d2i(round(((double)xx - L.x1) * L.k + L.y1)); }
    else { yy = m + L.yLow; xx = lsdb_x86_d2i(round(((double)yy - L.y1) / L.k + L.x1)); }
    return !(xx < 0 || xx >= W || yy < 0 || yy >= H) && xx != 0 && yy != 0;
}

__global__ void __launch_bounds__(FS_NT, 12) lsdb_fscan_frames_kernel(
    int maxBeams, const double* __restrict__ ranges, const double* __restrict__ angles, const int* __restrict__ beamOff,
    double resol, double oriX, double oriY, int leastPoint, double threLine, double leastDistM,
    LsdbFsInfo* __restrict__ info, LsdbFsPiece* __restrict__ pieces) {
    extern __shared__ __align__(16) unsigned char fsRaw[];
    FsShared& sh = *(FsShared*)fsRaw;
    const int cap = maxBeams + 1;
    double* px = (double*)(fsRaw + sizeof(FsShared));   // metric x, later grid x
    double* py = px + cap;
    int* cell = (int*)(py + cap);            // 2*(cap+1): start, end beam of each cluster
    int* stacks = cell + 2 * (cap + 1);      // FS_NW x 2*(cap+2)
    int* seg = stacks;                       // beam pairs of the line pieces (after RDP), compacted in place
    int* cnt = stacks + FS_NW * 2 * (cap + 2);   // cap+2: samples per kept line, then their exclusive prefix
    uint8_t* brk = (uint8_t*)(cnt + cap + 2);
    uint8_t* split = brk + cap;

    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = beamOff[f + 1] - beamOff[f];
    const double* R = ranges + beamOff[f];
    const double* A = angles + beamOff[f];
    const double pose0 = 0.0, pose1 = 0.0, pose2 = 0.0;      // scanPose, :11
    const int segCap = n + 2;                                // disjoint clusters give at most n + 1 pieces

    // ---- beams -> metric coordinates (:196-201, :289-293) ----
    for (int i = tid; i < n; i += FS_NT) {
        const double r = R[i], a = A[i] + pose2;
        px[i] = r * lsdm_cos(a) + pose0; py[i] = r * lsdm_sin(a) + pose1;
        split[i] = 0;
    }
    if (tid == 0) sh.overflow = 0;
    __syncthreads();
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

    // ---- Ramer-Douglas-Peucker (:218-272), clusters dealt to the warps ----
    {
        int* stack = stacks + warp * 2 * (cap + 2);
        for (int c = warp; c < nc; c += FS_NW) {
            if (lane == 0) { stack[0] = cell[2 * c]; stack[1] = cell[2 * c + 1]; }
            int sp = 2;
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
                double bestN = 0;                                   // largest numerator (NaN never wins, as in `dist > dist_max`)
                for (int i = 1 + lane; i < len - 1; i += 32) {
                    int a = s + i; if (a >= n) a -= n;
                    const double num = fabs(k * px[a] - py[a] + d);
                    if (num > bestN) bestN = num;
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, bestN, o); if (t > bestN) bestN = t; }
                double best = 0; int iMax = 0;
                const double q = bestN / den;                       // the maximum distance
                if (q > 0) {
                    const double lo = bestN < 1e-280 ? 0.0 : bestN * 0.99999999999999645;   // 1 - 2^-48
                    int pos = INT_MAX;                              // first point whose distance equals the maximum
                    for (int i = 1 + lane; i < len - 1; i += 32) {
                        int a = s + i; if (a >= n) a -= n;
                        const double num = fabs(k * px[a] - py[a] + d);
                        if (num >= lo && num / den == q) { pos = i; break; }
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, pos, o); if (t < pos) pos = t; }
                    if (pos != INT_MAX) { best = q; iMax = s + pos; if (iMax >= n) iMax -= n; }
                }
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
    }
    __syncthreads();

    // ---- grid coordinates in place (:21-34) and the raster extent ----
    double mnX = INFINITY, mnY = INFINITY, mxX = 0, mxY = 0;
    for (int i = tid; i < n; i += FS_NT) {
        const double X = floor((px[i] - oriX) / resol), Y = floor((py[i] - oriY) / resol);
        px[i] = X; py[i] = Y;
        if (X < mnX) mnX = X;
        if (X > mxX) mxX = X;
        if (Y < mnY) mnY = Y;
        if (Y > mxY) mxY = Y;
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
    // ---- line pieces in the reference's order (:49-74); the stacks are free now ----
    __syncthreads();
    if (tid == 0) {
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
    __syncthreads();
    double minX = sh.red[0][0], maxX = sh.red[1][0], minY = sh.red[2][0], maxY = sh.red[3][0];
#pragma unroll
    for (int w = 1; w < FS_NW; w++) {
        if (sh.red[0][w] < minX) minX = sh.red[0][w];
        if (sh.red[1][w] > maxX) maxX = sh.red[1][w];
        if (sh.red[2][w] < minY) minY = sh.red[2][w];
        if (sh.red[3][w] > maxY) maxY = sh.red[3][w];
    }
    const int W = lsdb_x86_d2i(ceil(maxX - minX)), H = lsdb_x86_d2i(ceil(maxY - minY));
    // ---- keep pieces of at least lineDistThre (:80-81), order preserved ----
    if (warp == 0) {
        const int nSeg = sh.nSeg;
        const double distThre = leastDistM / resol;
        int nl = 0;
        for (int base = 0; base < nSeg; base += 32) {
            const int i = base + lane;
            int a = 0, b = 0; bool keep = false;
            if (i < nSeg) {
                a = seg[2 * i]; b = seg[2 * i + 1];
                const double ddx = px[a] - px[b], ddy = py[a] - py[b];
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

    // ---- samples per line ----
    for (int L = warp; L < nLines; L += FS_NW) {
        const int a = seg[2 * L], b = seg[2 * L + 1];
        const FsLine g = fs_line(px[a] - minX, py[a] - minY, px[b] - minX, py[b] - minY);
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
        LsdbFsInfo o;
        o.nLines = sh.overflow ? -1 : nLines; o.nPts = run; o.W = W; o.H = H;
        o.lidarX = floor((pose0 - oriX) / resol - minX);   // :38-40
        o.lidarY = floor((pose1 - oriY) / resol - minY);
        info[f] = o;
    }
    __syncthreads();
    LsdbFsPiece* P = pieces + beamOff[f] + 2 * f;           // room for n + 2 pieces per frame
    for (int L = tid; L < nLines; L += FS_NT) {
        const int a = seg[2 * L], b = seg[2 * L + 1];
        LsdbFsPiece o;
        o.x1 = px[a] - minX; o.y1 = py[a] - minY; o.x2 = px[b] - minX; o.y2 = py[b] - minY; o.off = cnt[L]; o.pad = 0;
        P[L] = o;
    }
}

// one warp per kept line of the batch: record (:82-103, :154-174), samples and raster (:104-153)
__global__ void __launch_bounds__(FS_LW * 32) lsdb_fscan_lines_kernel(
    int nFrames, int nLinesTotal, const int* __restrict__ beamOff, const LsdbFsInfo* __restrict__ info,
    const LsdbFsPiece* __restrict__ pieces, const int* __restrict__ lineOff, const int* __restrict__ ptOff,
    const long long* __restrict__ imOff, const int* __restrict__ imPitch, int imVal, double pi, LsdbFaLine* __restrict__ lines,
    double* __restrict__ pts, uint8_t* __restrict__ lineIm) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * FS_LW + (threadIdx.x >> 5);
    if (g >= nLinesTotal) return;
    int lo = 0, hi = nFrames;                                // frame of line g: lineOff[f] <= g < lineOff[f+1]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (lineOff[mid] <= g) lo = mid; else hi = mid; }
    const int f = lo, L = g - lineOff[f];
    const LsdbFsPiece pc = pieces[beamOff[f] + 2 * f + L];
    const FsLine ln = fs_line(pc.x1, pc.y1, pc.x2, pc.y2);
    const int W = info[f].W, H = info[f].H;
    {
        double ang = 0; int orient = 1;
        if (lane == 0) {
            ang = lsdm_atan(ln.k) * 180.0 / pi;            // atand, LSD/baseFunc.cpp:14-16
            if (ang < 0) { ang += 180; orient = -1; }
        }
        ang = __shfl_sync(0xffffffffu, ang, 0);
        double cs = 0;
        if (lane < 2) cs = lsdm_sincos_eval(ang / 180.0 * pi, lane == 0);   // cosd / sind, :6-12 — one code path for both lanes
        const double sn = __shfl_sync(0xffffffffu, cs, 1);
        if (lane == 0) {
            LsdbFaLine o;
            o.k = ln.k;
            o.b = (ln.y1 + ln.y2) / 2.0 - ln.k * (ln.x1 + ln.x2) / 2.0;
            o.dx = cs; o.dy = sn;
            o.x1 = ln.x1; o.y1 = ln.y1; o.x2 = ln.x2; o.y2 = ln.y2;
            const double ey = ln.y2 - ln.y1, ex = ln.x2 - ln.x1;
            o.len = sqrt(ey * ey + ex * ex);
            o.orient = orient; o.pad = 0;
            lines[g] = o;
        }
    }
    double* outP = pts + 2 * ((size_t)ptOff[f] + (size_t)pc.off);
    uint8_t* im = lineIm ? lineIm + imOff[f] : 0;
    const int pitch = imPitch ? imPitch[f] : W;              // rasters packed (FS.lineIm) or written into a batch's padded source rows
    int run = 0;
    for (int m0 = 0; m0 < ln.cnt; m0 += 32) {
        int xx = 0, yy = 0;
        const bool ok = m0 + lane < ln.cnt && fs_sample(ln, m0 + lane, W, H, xx, yy);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const size_t o = (size_t)(run + __popc(m & ((1u << lane) - 1)));
            outP[2 * o] = (double)xx; outP[2 * o + 1] = (double)yy;
            if (im) im[(size_t)yy * pitch + xx] = (uint8_t)imVal;
        }
        run += __popc(m);
    }
}

int lsdb_launch_fscan_frames(cudaStream_t s, int nFrames, int maxBeams, const double* ranges, const double* angles, const int* beamOff,
                             double resol, double oriX, double oriY, int leastPoint, double threLine, double leastDistM,
                             LsdbFsInfo* info, LsdbFsPiece* pieces) {
    const size_t smem = lsdb_fscan_smem(maxBeams);
    cudaError_t e = cudaFuncSetAttribute(lsdb_fscan_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lsdb_fscan_frames_kernel<<<nFrames, FS_NT, smem, s>>>(maxBeams, ranges, angles, beamOff, resol, oriX, oriY, leastPoint, threLine,
                                                           leastDistM, info, pieces);
    return (int)cudaGetLastError();
}

int lsdb_launch_fscan_lines(cudaStream_t s, int nFrames, int nLinesTotal, const int* beamOff, const LsdbFsInfo* info,
                            const LsdbFsPiece* pieces, const int* lineOff, const int* ptOff, const long long* imOff, const int* imPitch,
                            int imVal, double pi, LsdbFaLine* lines, double* pts, uint8_t* lineIm) {
    if (nLinesTotal <= 0) return 0;
    lsdb_fscan_lines_kernel<<<(nLinesTotal + FS_LW - 1) / FS_LW, FS_LW * 32, 0, s>>>(nFrames, nLinesTotal, beamOff, info, pieces, lineOff,
                                                                                     ptOff, imOff, imPitch, imVal, pi, lines, pts, lineIm);
    return (int)cudaGetLastError();
}
