// Stage C of the LSD hot path on sm_100a: the seed loop — region growing, rectangle fit, density
// refinement, NFA validation — as ONE persistent kernel with speculative, order-preserving commit.
//
// Replaces (reference, /root/reference/LSD/myLSD.cpp): the sequential seed loop :218-272,
// RegionGrower :491-590, CenterGetter/OrientationGetter/RectangleConverter :592-734,
// RegionRadiusReducer :736-802, Refiner :804-880, LogGammaCalculator :882-924,
// RectangleNFACalculator :926-1059 and RectangleImprover :1061-1158.
//
// Exactness.  The reference's result depends on the order seeds are visited (usedMap evolves) and
// on the order pixels join a region (regDeg is re-estimated after every accepted pixel).  Both are
// kept: a warp replays one region in the reference's exact candidate order (lanes only evaluate
// the 32 next neighbour candidates in parallel; accepts are applied one at a time), and regions
// are RETIRED strictly in seed order:
//   * the sorted seed list is cut into chunks of 32 cells; warps take chunks by ticket;
//   * a warp evaluates the live seeds of its chunk speculatively against the current state;
//   * it then waits until every earlier chunk has retired (frontier == its ticket) and validates
//     each evaluation: it stands iff no region ACCEPTED since the evaluation started overlaps the
//     bounding box of the pixels the evaluation examined (regions that are too small, fail
//     refinement or are NFA-rejected never change what growth sees — SURVEY.md §0 fact 4);
//     otherwise the seed is simply re-evaluated at the frontier, where the state is final;
//   * commits (usedMap 1 / 2, labels, rectangle record) happen only at the frontier.
// So usedMap, labels and the segment list are exactly the sequential result.
//
// One CTA (16 warps) per map, CTAs pull maps from a counter.  All floating-point sums the
// reference accumulates sequentially are accumulated sequentially here too (lane-parallel loads,
// serial adds); counts and min/max are order-free and reduced in parallel.  Stage is
// latency-bound (dependent gathers + double-double trig chains), not bandwidth-bound.
#include "lsdb_common.cuh"
#include "../../include/lsdb200.h"

#define NW LSDB_GROW_WARPS
#define LOG_CAP 1024
#define FULL 0xffffffffu

enum { OC_NONE = 0, OC_NOCHANGE = 1, OC_REJECT = 2, OC_ACCEPT = 3, OC_DEFER = 4 };
enum { ST_CELLS = 0, ST_LIVE, ST_GROWS, ST_GROWNPX, ST_SMALL, ST_REGROWS, ST_RRR, ST_NFACALLS, ST_NFAPX, ST_REJECTS,
       ST_ACCEPTS, ST_SPEC, ST_RESPEC, ST_CHUNKS, ST_N,
       // cycle counters (lane 0 of every warp, summed): only kept in LSDB_TIMING builds, reported through stat[] slots 14..19
       TM_GROW = ST_N, TM_RECT, TM_NFA, TM_WAIT, TM_RETIRE, TM_SPEC, TM_N };

struct GrowShared {
    volatile int frontier;
    int nextChunk;
    volatile int logCount;
    volatile int nSeg;
    volatile int abortFlag;
    int img;
    int nChunks;
    int nCells;
    unsigned int logBox[LOG_CAP][2];
    unsigned long long stats[TM_N];
};

struct Rect { double x1, y1, x2, y2, wid, cX, cY, deg, dx, dy, p, prec; };

struct BBox { int x0, y0, x1, y1; };

#if 1  /* cycle counters are cheap (one clock64 + one smem atomic per measured call) and feed bench.py */
#define TIC long long t0_ = clock64()
#define TOC(c, idx) do { if ((c).lane == 0) atomicAdd(&(c).sh->stats[idx], (unsigned long long)(clock64() - t0_)); } while (0)
#else
#define TIC
#define TOC(c, idx)
#endif
#define STAT(c, idx, v) do { if ((c).lane == 0) atomicAdd(&(c).sh->stats[idx], (unsigned long long)(v)); } while (0)

struct WarpCtx {
    int W, H, lane, w;
    unsigned int mybit;
    unsigned int* state;
    const double* deg;
    const double* mag;
    unsigned int* list;   // working point list (packed y<<16|x)
    unsigned int* tlist;  // second list: RRR backup / commit stash
    int listCap;
    const LsdbLsdConst* kc;
    const double* lgammaTab;
    int lgammaN;
    GrowShared* sh;
    double logNT, regThre;
};

// noinline wrappers keep one copy of each math routine in the kernel
__device__ __noinline__ double d_sin(double x) { return lsdm_sin(x); }
__device__ __noinline__ double d_cos(double x) { return lsdm_cos(x); }
__device__ __noinline__ double d_atan2(double y, double x) { return lsdm_atan2(y, x); }
__device__ __noinline__ double d_log(double x) { return lsdm_log(x); }
__device__ __noinline__ double d_log10(double x) { return lsdm_log10(x); }
__device__ __noinline__ double d_exp(double x) { return lsdm_exp(x); }
__device__ __noinline__ double d_pow(double x, double y) { return lsdm_pow(x, y); }

__device__ __forceinline__ unsigned int pack_xy(int x, int y) { return ((unsigned int)y << 16) | (unsigned int)x; }
__device__ __forceinline__ int px_of(unsigned int v) { return (int)(v & 0xffffu); }
__device__ __forceinline__ int py_of(unsigned int v) { return (int)(v >> 16); }

__device__ __forceinline__ void bbox_add(BBox& b, int x, int y) {
    b.x0 = min(b.x0, x); b.y0 = min(b.y0, y); b.x1 = max(b.x1, x); b.y1 = max(b.y1, y);
}

// clear this warp's curMap bit on list[0..num)
__device__ void clear_bits(const WarpCtx& c, const unsigned int* lst, int num) {
    for (int k = c.lane; k < num; k += 32) {
        unsigned int v = lst[k];
        atomicAnd(&c.state[(size_t)py_of(v) * c.W + px_of(v)], ~c.mybit);
    }
    __syncwarp();
}

// ------------------------------------------------------------------ RegionGrower (:491-590)
// Returns the region size (points in c.list), -1 on list overflow.  regDeg in/out.
__device__ int grow_region(WarpCtx& c, int sx, int sy, double& regDeg, double degThre, BBox& bb) {
    const int W = c.W, H = c.H;
    const double pi = c.kc->pi;
    const double pi32 = pi * 3 / 2.0, pi2 = 2.0 * pi;
    TIC;
    double sinDeg = d_sin(regDeg), cosDeg = d_cos(regDeg);
    if (c.lane == 0) {
        c.list[0] = pack_xy(sx, sy);
        atomicOr(&c.state[(size_t)sy * W + sx], c.mybit);
    }
    __syncwarp();
    bbox_add(bb, sx, sy);
    int num = 1, exNum = 0;
    while (exNum != num) {
        exNum = num;
        int cur = 0;
        while (cur < 9 * num) {
            const int lim = 9 * num;
            const int myc = cur + c.lane;
            const bool valid = myc < lim;
            const int pt = myc / 9, nb = myc - pt * 9;
            const unsigned int pv = valid ? c.list[pt] : 0u;
            const int m = py_of(pv) + nb / 3 - 1, n = px_of(pv) + (nb - (nb / 3) * 3) - 1;
            const bool inb = valid && m >= 0 && n >= 0 && m < H && n < W;
            const size_t p = inb ? (size_t)m * W + n : 0;
            const unsigned int st = inb ? lsdb_ld_state(&c.state[p]) : LSDB_ST_BAN;
            const double dg = inb ? c.deg[p] : 0.0;
            bool cand = inb && !(st & (LSDB_ST_BAN | c.mybit));
            int start = 0;
            while (true) {
                double degDif = fabs(regDeg - dg);
                if (degDif > pi32) degDif = fabs(degDif - pi2);
                const bool pass = cand && c.lane >= start && degDif < degThre;
                const unsigned int b = __ballot_sync(FULL, pass);
                if (!b) break;
                const int f = __ffs(b) - 1;
                const double curDeg = __shfl_sync(FULL, dg, f);
                const int fn = __shfl_sync(FULL, n, f), fm = __shfl_sync(FULL, m, f);
                cosDeg += d_cos(curDeg);
                sinDeg += d_sin(curDeg);
                regDeg = d_atan2(sinDeg, cosDeg);
                if (num >= c.listCap - 1) return -1;
                if (c.lane == f) {
                    atomicOr(&c.state[p], c.mybit);
                    c.list[num] = pack_xy(n, m);
                }
                if (inb && n == fn && m == fm) cand = false;
                bbox_add(bb, fn, fm);
                num++;
                start = f + 1;
            }
            __syncwarp();
            cur = min(cur + 32, lim);
        }
    }
    STAT(c, ST_GROWS, 1); STAT(c, ST_GROWNPX, num);
    TOC(c, TM_GROW);
    return num;
}

// ------------------------------------------------------------------ RectangleConverter (:592-734)
__device__ Rect rect_from_region(const WarpCtx& c, const unsigned int* lst, int num, double regDeg, double aliPro,
                                 double degThre) {
    const int W = c.W;
    const double pi = c.kc->pi;
    TIC;
    double cenX = 0, cenY = 0, weiSum = 0;
    for (int base = 0; base < num; base += 32) {  // CenterGetter :608-613, sums in list order
        const int k = base + c.lane;
        const unsigned int v = k < num ? lst[k] : 0u;
        const double wv = k < num ? c.mag[(size_t)py_of(v) * W + px_of(v)] : 0.0;
        const int cnt = min(32, num - base);
        for (int j = 0; j < cnt; j++) {
            const double wj = __shfl_sync(FULL, wv, j);
            const unsigned int vj = __shfl_sync(FULL, v, j);
            cenX += wj * px_of(vj);
            cenY += wj * py_of(vj);
            weiSum += wj;
        }
    }
    cenX = cenX / weiSum;
    cenY = cenY / weiSum;
    double Ixx = 0, Iyy = 0, Ixy = 0;
    weiSum = 0;
    for (int base = 0; base < num; base += 32) {  // OrientationGetter :637-643
        const int k = base + c.lane;
        const unsigned int v = k < num ? lst[k] : 0u;
        const double wv = k < num ? c.mag[(size_t)py_of(v) * W + px_of(v)] : 0.0;
        const int cnt = min(32, num - base);
        for (int j = 0; j < cnt; j++) {
            const double wj = __shfl_sync(FULL, wv, j);
            const unsigned int vj = __shfl_sync(FULL, v, j);
            const double ey = py_of(vj) - cenY, ex = px_of(vj) - cenX;
            Ixx += wj * (ey * ey);
            Iyy += wj * (ex * ex);
            Ixy -= wj * ex * ey;
            weiSum += wj;
        }
    }
    Ixx /= weiSum; Iyy /= weiSum; Ixy /= weiSum;
    const double dI = Ixx - Iyy;
    const double lamb = (Ixx + Iyy - sqrt(dI * dI + 4 * Ixy * Ixy)) / 2.0;
    double inertiaDeg;
    if (fabs(Ixx) > fabs(Iyy)) inertiaDeg = d_atan2(lamb - Ixx, Ixy);
    else inertiaDeg = d_atan2(Ixy, lamb - Iyy);
    double regDif = inertiaDeg - regDeg;
    while (regDif <= -pi) regDif += 2 * pi;
    while (regDif > pi) regDif -= 2 * pi;
    if (regDif < 0) regDif = -regDif;
    if (regDif > degThre) inertiaDeg += pi;

    const double dx = d_cos(inertiaDeg), dy = d_sin(inertiaDeg);
    double lenMin = 0, lenMax = 0, widMin = 0, widMax = 0;  // :701-714, order-free
    for (int k = c.lane; k < num; k += 32) {
        const unsigned int v = lst[k];
        const double len = (px_of(v) - cenX) * dx + (py_of(v) - cenY) * dy;
        const double wid = -(px_of(v) - cenX) * dy + (py_of(v) - cenY) * dx;
        if (len < lenMin) lenMin = len;
        if (len > lenMax) lenMax = len;
        if (wid < widMin) widMin = wid;
        if (wid > widMax) widMax = wid;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double t;
        t = __shfl_xor_sync(FULL, lenMin, o); if (t < lenMin) lenMin = t;
        t = __shfl_xor_sync(FULL, lenMax, o); if (t > lenMax) lenMax = t;
        t = __shfl_xor_sync(FULL, widMin, o); if (t < widMin) widMin = t;
        t = __shfl_xor_sync(FULL, widMax, o); if (t > widMax) widMax = t;
    }
    Rect r;
    r.x1 = cenX + lenMin * dx; r.y1 = cenY + lenMin * dy;
    r.x2 = cenX + lenMax * dx; r.y2 = cenY + lenMax * dy;
    r.wid = widMax - widMin;
    r.cX = cenX; r.cY = cenY; r.deg = inertiaDeg; r.dx = dx; r.dy = dy;
    r.p = aliPro; r.prec = degThre;
    if (r.wid < 1) r.wid = 1;
    TOC(c, TM_RECT);
    return r;
}

__device__ __forceinline__ double rect_density(int num, const Rect& r) {  // :757-758,:827
    const double ax = r.x1 - r.x2, ay = r.y1 - r.y2;
    return num / (sqrt(ax * ax + ay * ay) * r.wid);
}
__device__ __forceinline__ double dist_xy(int ox, int oy, double x, double y) {
    const double a = ox - x, b = oy - y;
    return sqrt(a * a + b * b);
}

// ------------------------------------------------------------------ LogGammaCalculator (:882-924)
__device__ double log_gamma_calc(int x) {
    double val;
    if (x > 15) {
        const double xd = x;
        val = 0.918938533204673 + (xd - 0.5) * d_log(xd) - xd +
              0.5 * xd * d_log(xd * lsdm_sinh(1.0 / xd) + 1.0 / (810 * d_pow(xd, 6)));
    } else {
        const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705,
                             1168.92649479, 83.8676043424, 2.50662827511};
        double a = (x + 0.5) * d_log(x + 5.5) - (x + 5.5);
        double b = 0;
        for (int i = 0; i < 7; i++) {
            a -= d_log(x + i);
            b += q[i] * d_pow(x, i);
        }
        val = a + d_log(b);
    }
    return val;
}
__device__ __forceinline__ double log_gamma(const WarpCtx& c, int x) {
    if (x >= 0 && x < c.lgammaN) return c.lgammaTab[x];
    return log_gamma_calc(x);
}

__global__ void lsdb_lgamma_table_kernel(double* tab, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tab[i] = i >= 1 ? log_gamma_calc(i) : 0.0;
}

// ------------------------------------------------------------------ RectangleNFACalculator (:926-1059)
__device__ double rect_nfa_impl(WarpCtx& c, const Rect& rec, double logNT) {
    const int xLim = c.W, yLim = c.H;
    const double pi = c.kc->pi;
    const double pi32 = pi * 3 / 2.0, pi2 = 2 * pi;
    double verX[4], verY[4];
    verX[0] = rec.x1 - rec.dy * rec.wid / 2.0;
    verX[1] = rec.x2 - rec.dy * rec.wid / 2.0;
    verX[2] = rec.x2 + rec.dy * rec.wid / 2.0;
    verX[3] = rec.x1 + rec.dy * rec.wid / 2.0;
    verY[0] = rec.y1 + rec.dx * rec.wid / 2.0;
    verY[1] = rec.y2 + rec.dx * rec.wid / 2.0;
    verY[2] = rec.y2 - rec.dx * rec.wid / 2.0;
    verY[3] = rec.y1 - rec.dx * rec.wid / 2.0;
    int offset;
    if ((rec.x1 < rec.x2) && (rec.y1 <= rec.y2)) offset = 0;
    else if ((rec.x1 >= rec.x2) && (rec.y1 < rec.y2)) offset = 1;
    else if ((rec.x1 > rec.x2) && (rec.y1 >= rec.y2)) offset = 2;
    else offset = 3;
    const double vX0 = verX[offset & 3], vX1 = verX[(offset + 1) & 3], vX2 = verX[(offset + 2) & 3], vX3 = verX[(offset + 3) & 3];
    const double vY0 = verY[offset & 3], vY1 = verY[(offset + 1) & 3], vY2 = verY[(offset + 2) & 3], vY3 = verY[(offset + 3) & 3];

    int allPixNum = 0, aliPixNum = 0;
    const int xr = lsdb_x86_d2i(ceil(vX0) - floor(vX2));
    const int xRang_len = (xr == (int)0x80000000 ? xr : abs(xr)) + 1;
    if (xRang_len > 0 && xRang_len < 100000000) {
        const double x0c = ceil(vX0);
        const double k0 = (vY1 - vY0) / (vX1 - vX0);
        const double k1 = (vY2 - vY1) / (vX2 - vX1);
        const double k2 = (vY2 - vY3) / (vX2 - vX3);
        const double k3 = (vY3 - vY0) / (vX3 - vX0);
        for (int i = c.lane; i < xRang_len; i += 32) {
            const int xi = lsdb_x86_d2i(i + x0c);
            // the reference fills yLow/yHigh with two partition passes (:987-1004); xi is
            // increasing, so entry i takes the first branch iff xi < vertex (NaN vertex: slot stays 0)
            int yl = 0, yh = 0;
            if (xi < vX3) yl = lsdb_x86_d2i(ceil(vY0 + (xi - vX0) * k3));
            else if (xi >= vX3) yl = lsdb_x86_d2i(ceil(vY3 + (xi - vX3) * k2));
            if (xi < vX1) yh = lsdb_x86_d2i(floor(vY0 + (xi - vX0) * k0));
            else if (xi >= vX1) yh = lsdb_x86_d2i(floor(vY1 + (xi - vX1) * k1));
            if (xi < 0 || xi >= xLim) continue;
            const int j0 = yl < 0 ? 0 : yl, j1 = yh > yLim - 1 ? yLim - 1 : yh;
            for (int j = j0; j <= j1; j++) {
                allPixNum++;
                double degDif = fabs(rec.deg - c.deg[(size_t)j * xLim + xi]);
                if (degDif > pi32) degDif = fabs(degDif - pi2);
                if (degDif < rec.prec) aliPixNum++;
            }
        }
        __syncwarp();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            allPixNum += __shfl_xor_sync(FULL, allPixNum, o);
            aliPixNum += __shfl_xor_sync(FULL, aliPixNum, o);
        }
    }
    STAT(c, ST_NFACALLS, 1); STAT(c, ST_NFAPX, allPixNum);

    if (allPixNum == 0 || aliPixNum == 0) return -logNT;
    if (allPixNum == aliPixNum) return -logNT - allPixNum * d_log10(rec.p);
    const double proTerm = rec.p / (1.0 - rec.p);
    const double log1Coef = log_gamma(c, allPixNum + 1) - log_gamma(c, aliPixNum + 1) - log_gamma(c, allPixNum - aliPixNum + 1);
    const double log1Term = log1Coef + aliPixNum * d_log(rec.p) + (allPixNum - aliPixNum) * d_log(1 - rec.p);
    double term = d_exp(log1Term);
    const double eps = 2.2204e-16;
    if (fabs(term) < 100 * eps) {
        if (aliPixNum > allPixNum * rec.p) return -d_log10(term) - logNT;
        return -logNT;
    }
    double binTail = term;
    const double tole = 0.1;
    for (int i = aliPixNum + 1; i <= allPixNum; i++) {
        const double binTerm = (allPixNum - i + 1) / (i * 1.0);
        const double multTerm = binTerm * proTerm;
        term *= multTerm;
        binTail += term;
        if (binTerm < 1) {
            const double err = term * ((1 - d_pow(multTerm, allPixNum - i + 1)) / (1.0 - multTerm) - 1);
            if (err < tole * fabs(-d_log10(binTail) - logNT) * binTail) break;
        }
    }
    return -d_log10(binTail) - logNT;
}

__device__ double rect_nfa(WarpCtx& c, const Rect& rec, double logNT) {
    TIC;
    const double v = rect_nfa_impl(c, rec, logNT);
    TOC(c, TM_NFA);
    return v;
}

// ------------------------------------------------------------------ RectangleImprover (:1061-1158)
__device__ double rectangle_improver(WarpCtx& c, Rect& rec, double logNT) {
    const double pi = c.kc->pi;
    const double delt = 0.5, delt2 = delt / 2.0;
    double best = rect_nfa(c, rec, logNT);
    Rect bestRec = rec;
    if (best > 0) return best;
    Rect r = bestRec;
    double v;
    for (int i = 0; i < 5; i++) {
        r.p /= 2.0; r.prec = r.p * pi;
        v = rect_nfa(c, r, logNT);
        if (v > best) { best = v; bestRec = r; }
    }
    if (best > 0) { rec = bestRec; return best; }
    for (int side = 0; side < 3; side++) {  // 0: width, 1: side one, 2: side two  (:1096-1143)
        r = bestRec;
        for (int i = 0; i < 5; i++) {
            if (r.wid - delt >= 0.5) {
                if (side == 1) { r.x1 -= r.dy * delt2; r.y1 += r.dx * delt2; r.x2 -= r.dy * delt2; r.y2 += r.dx * delt2; }
                if (side == 2) { r.x1 += r.dy * delt2; r.y1 -= r.dx * delt2; r.x2 += r.dy * delt2; r.y2 -= r.dx * delt2; }
                r.wid -= delt;
                v = rect_nfa(c, r, logNT);
                if (v > best) { best = v; bestRec = r; }
            }
        }
        if (best > 0) { rec = bestRec; return best; }
    }
    r = bestRec;
    for (int i = 0; i < 5; i++) {
        r.p /= 2.0; r.prec = r.p * pi;
        v = rect_nfa(c, r, logNT);
        if (v > best) { best = v; bestRec = r; }
    }
    rec = bestRec;
    return best;
}

// ------------------------------------------------------------------ one seed: grow -> rect -> refine -> NFA
// On OC_REJECT / OC_ACCEPT the pixels with curMap==1 are compacted into c.tlist[0..*nCommit) and the
// warp's curMap bits are cleared.  allowTlist=false forbids touching c.tlist (it holds a stash):
// the evaluation returns OC_DEFER as soon as it would need it.
__device__ int eval_seed(WarpCtx& c, int p0, bool allowTlist, Rect& rec, double& logNFA, BBox& bb, int& nCommit,
                         bool& touchedT) {
    const LsdbLsdConst* kc = c.kc;
    const int W = c.W;
    const int sx = p0 % W, sy = p0 / W;
    bb.x0 = bb.y0 = 0x7fffffff; bb.x1 = bb.y1 = -1;
    touchedT = false;
    double regDeg = c.deg[p0];
    int num = grow_region(c, sx, sy, regDeg, kc->degThre, bb);
    if (num < 0) { c.sh->abortFlag = LSDB_ERR_CAPACITY; return OC_NOCHANGE; }
    if (num < c.regThre) {  // :228
        clear_bits(c, c.list, num);
        STAT(c, ST_SMALL, 1);
        return OC_NOCHANGE;
    }
    if (!allowTlist) { clear_bits(c, c.list, num); return OC_DEFER; }  // would need the stash buffer
    touchedT = true;
    rec = rect_from_region(c, c.list, num, regDeg, kc->aliPro, kc->degThre);
    bool usedT = false;
    int tnum = 0;
    // Refiner :804-880
    double den = rect_density(num, rec);
    if (!(den >= kc->denThre)) {
        const double pi = kc->pi;
        const double cenDeg = c.deg[p0];
        double difSum = 0, squSum = 0;
        int ptNum = 0;
        for (int base = 0; base < num; base += 32) {  // :839-853, sums in list order
            const int k = base + c.lane;
            const unsigned int v = k < num ? c.list[k] : 0u;
            bool in = false;
            double dd = 0;
            if (k < num && dist_xy(sx, sy, (double)px_of(v), (double)py_of(v)) < rec.wid) {
                in = true;
                dd = c.deg[(size_t)py_of(v) * W + px_of(v)] - cenDeg;
                while (dd <= -pi) dd += 2 * pi;
                while (dd > pi) dd -= 2 * pi;
            }
            unsigned int mask = __ballot_sync(FULL, in);
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                const double dj = __shfl_sync(FULL, dd, j);
                difSum += dj;
                squSum += dj * dj;
                ptNum++;
            }
        }
        const double meanDif = difSum / (ptNum * 1.0);
        const double degThre2 = 2.0 * sqrt((squSum - 2 * meanDif * difSum) / (ptNum * 1.0) + meanDif * meanDif);
        clear_bits(c, c.list, num);
        regDeg = cenDeg;
        num = grow_region(c, sx, sy, regDeg, degThre2, bb);
        STAT(c, ST_REGROWS, 1);
        if (num < 0) { c.sh->abortFlag = LSDB_ERR_CAPACITY; return OC_NOCHANGE; }
        if (num < 2) { clear_bits(c, c.list, num); return OC_NOCHANGE; }
        rec = rect_from_region(c, c.list, num, regDeg, rec.p, rec.prec);
        den = rect_density(num, rec);
        if (den < kc->denThre) {
            for (int k = c.lane; k < num; k += 32) c.tlist[k] = c.list[k];  // every pixel the bit was set on
            __syncwarp();
            usedT = true; tnum = num;
            // RegionRadiusReducer :736-802 (rect_from_region inside needs reg.deg = regDeg)
            bool ok = true;
            {
                double d2 = rect_density(num, rec);
                if (!(d2 > kc->denThre)) {
                    const double rad1 = dist_xy(sx, sy, rec.x1, rec.y1), rad2 = dist_xy(sx, sy, rec.x2, rec.y2);
                    double rad = rad1 > rad2 ? rad1 : rad2;
                    while (d2 < kc->denThre) {
                        rad *= 0.75;
                        if (c.lane == 0) {
                            int i = 0, nn = num;
                            c.list[nn] = 0u;  // slot [num] reads as (0,0)  
                            while (i <= nn) {
                                const unsigned int v = c.list[i];
                                if (dist_xy(sx, sy, (double)px_of(v), (double)py_of(v)) > rad) {
                                    atomicAnd(&c.state[(size_t)py_of(v) * W + px_of(v)], ~c.mybit);
                                    c.list[i] = c.list[nn - 1];
                                    c.list[nn - 1] = 0u;
                                    i--;
                                    nn--;
                                }
                                i++;
                            }
                            num = nn;
                            atomicAdd(&c.sh->stats[ST_RRR], 1ull);
                        }
                        __syncwarp();
                        num = __shfl_sync(FULL, num, 0);
                        if (num < 2) { ok = false; break; }
                        rec = rect_from_region(c, c.list, num, regDeg, rec.p, rec.prec);
                        d2 = rect_density(num, rec);
                    }
                }
            }
            if (!ok) { clear_bits(c, c.tlist, tnum); return OC_NOCHANGE; }
        }
    }
    logNFA = rectangle_improver(c, rec, c.logNT);
    // finalize: commit list = pixels whose curMap bit is still set
    int outN = 0;
    if (!usedT) {
        for (int k = c.lane; k < num; k += 32) {
            const unsigned int v = c.list[k];
            c.tlist[k] = v;
            atomicAnd(&c.state[(size_t)py_of(v) * W + px_of(v)], ~c.mybit);
        }
        outN = num;
    } else {
        for (int base = 0; base < tnum; base += 32) {
            const int k = base + c.lane;
            unsigned int v = 0;
            bool keep = false;
            if (k < tnum) {
                v = c.tlist[k];
                const unsigned int old = atomicAnd(&c.state[(size_t)py_of(v) * W + px_of(v)], ~c.mybit);
                keep = (old & c.mybit) != 0;
            }
            const unsigned int mk = __ballot_sync(FULL, keep);
            __syncwarp();
            if (keep) c.tlist[outN + __popc(mk & ((1u << c.lane) - 1u))] = v;
            outN += __popc(mk);
            __syncwarp();
        }
    }
    __syncwarp();
    nCommit = outN;
    return logNFA <= 0 ? OC_REJECT : OC_ACCEPT;
}

// any ACCEPT logged in [L0, logCount) overlapping the examined box (region bbox dilated by 1)?
__device__ bool has_conflict(const WarpCtx& c, int L0, unsigned int bb0, unsigned int bb1) {
    const int L1 = c.sh->logCount;
    if (L1 - L0 > LOG_CAP) return true;
    const int ax0 = (int)(bb0 & 0xffff) - 1, ay0 = (int)(bb0 >> 16) - 1, ax1 = (int)(bb1 & 0xffff) + 1, ay1 = (int)(bb1 >> 16) + 1;
    bool hit = false;
    for (int e = L0 + c.lane; e < L1; e += 32) {
        const unsigned int b0 = c.sh->logBox[e & (LOG_CAP - 1)][0], b1 = c.sh->logBox[e & (LOG_CAP - 1)][1];
        const int bx0 = b0 & 0xffff, by0 = b0 >> 16, bx1 = b1 & 0xffff, by1 = b1 >> 16;
        if (bx0 <= ax1 && bx1 >= ax0 && by0 <= ay1 && by1 >= ay0) hit = true;
    }
    return __any_sync(FULL, hit);
}

// commit c.tlist[0..nCommit) at the frontier (:242-271)
__device__ void commit_region(WarpCtx& c, int outcome, int nCommit, const Rect& rec, double logNFA, int* labels,
                              LsdbRect* rects, int maxSeg) {
    GrowShared* sh = c.sh;
    const int W = c.W;
    if (outcome == OC_REJECT) {
        for (int k = c.lane; k < nCommit; k += 32) {
            const unsigned int v = c.tlist[k];
            atomicOr(&c.state[(size_t)py_of(v) * W + px_of(v)], LSDB_ST_REJ);
        }
        STAT(c, ST_REJECTS, 1);
        __syncwarp();
        return;
    }
    const int idx = sh->nSeg;
    int x0 = 0x7fffffff, y0 = 0x7fffffff, x1 = -1, y1 = -1;
    for (int k = c.lane; k < nCommit; k += 32) {
        const unsigned int v = c.tlist[k];
        const size_t p = (size_t)py_of(v) * W + px_of(v);
        atomicOr(&c.state[p], LSDB_ST_BAN);
        labels[p] += idx + 1;  // regIdx += curMap*(regCnt+1), :261 (int32 here, u8 there)
        x0 = min(x0, px_of(v)); y0 = min(y0, py_of(v)); x1 = max(x1, px_of(v)); y1 = max(y1, py_of(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x0 = min(x0, __shfl_xor_sync(FULL, x0, o)); y0 = min(y0, __shfl_xor_sync(FULL, y0, o));
        x1 = max(x1, __shfl_xor_sync(FULL, x1, o)); y1 = max(y1, __shfl_xor_sync(FULL, y1, o));
    }
    if (c.lane == 0) {
        if (idx < maxSeg) {
            const double sca = c.kc->sca;
            LsdbRect& R = rects[idx];
            double rx1 = rec.x1, ry1 = rec.y1, rx2 = rec.x2, ry2 = rec.y2, rw = rec.wid;
            if (sca != 1) {  // :252-258
                rx1 = (rx1 - 1.0) / sca + 1; ry1 = (ry1 - 1.0) / sca + 1;
                rx2 = (rx2 - 1.0) / sca + 1; ry2 = (ry2 - 1.0) / sca + 1;
                rw = (rw - 1.0) / sca + 1;
            }
            R.v[0] = rx1; R.v[1] = ry1; R.v[2] = rx2; R.v[3] = ry2; R.v[4] = rw; R.v[5] = rec.cX; R.v[6] = rec.cY;
            R.v[7] = rec.deg; R.v[8] = rec.dx; R.v[9] = rec.dy; R.v[10] = rec.p; R.v[11] = rec.prec; R.v[12] = logNFA;
        } else {
            sh->abortFlag = LSDB_ERR_CAPACITY;
        }
        const int L = sh->logCount;
        sh->logBox[L & (LOG_CAP - 1)][0] = pack_xy(x0, y0);
        sh->logBox[L & (LOG_CAP - 1)][1] = pack_xy(x1, y1);
        __threadfence_block();
        sh->logCount = L + 1;
        sh->nSeg = idx + 1;
        atomicAdd(&sh->stats[ST_ACCEPTS], 1ull);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(NW * 32, 1) lsdb_grow_kernel(int nImgs, const LsdbImg* __restrict__ imgs, LsdbImgDyn* __restrict__ dyn,
                                                               const LsdbLsdConst* __restrict__ kc, const double* __restrict__ mag,
                                                               const double* __restrict__ deg, unsigned int* __restrict__ state,
                                                               const unsigned int* __restrict__ cells, int* __restrict__ labels,
                                                               LsdbRect* __restrict__ rects, int maxSeg, unsigned int* __restrict__ lists,
                                                               int listCap, const double* __restrict__ lgammaTab, int lgammaN,
                                                               int* __restrict__ imgCounter) {
    __shared__ GrowShared sh;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    WarpCtx c;
    c.lane = lane; c.w = w; c.mybit = 1u << (LSDB_ST_WARP_SHIFT + w);
    c.kc = kc; c.lgammaTab = lgammaTab; c.lgammaN = lgammaN; c.sh = &sh;
    c.listCap = listCap;
    c.list = lists + ((size_t)blockIdx.x * NW + w) * 2 * (size_t)listCap;
    c.tlist = c.list + listCap;

    while (true) {
        __syncthreads();
        if (tid == 0) {
            const int img = atomicAdd(imgCounter, 1);
            sh.img = img;
            if (img < nImgs) {
                sh.frontier = 0; sh.nextChunk = 0; sh.logCount = 0; sh.nSeg = 0; sh.abortFlag = 0;
                sh.nCells = dyn[img].nCells;
                sh.nChunks = (sh.nCells + LSDB_CHUNK - 1) / LSDB_CHUNK;
            }
        }
        if (tid < TM_N) sh.stats[tid] = 0;
        __syncthreads();
        const int img = sh.img;
        if (img >= nImgs) break;
        const LsdbImg im = imgs[img];
        c.W = im.W; c.H = im.H; c.logNT = im.logNT; c.regThre = im.regThre;
        c.state = state + im.nOff; c.deg = deg + im.nOff; c.mag = mag + im.nOff;
        int* lab = labels + im.nOff;
        const unsigned int* cl = cells + im.nOff;
        LsdbRect* rc = rects + im.segOff;
        const int nCells = sh.nCells, nChunks = sh.nChunks;

        while (true) {
            int chunk = 0;
            if (lane == 0) chunk = atomicAdd(&sh.nextChunk, 1);
            chunk = __shfl_sync(FULL, chunk, 0);
            if (chunk >= nChunks || sh.abortFlag) break;
            const int ci = chunk * LSDB_CHUNK + lane;
            const int myp = ci < nCells ? (int)cl[ci] : -1;
            const bool live = myp >= 0 && (lsdb_ld_state(&c.state[myp]) & 3u) == 0;
            const unsigned int liveMask = __ballot_sync(FULL, live);
            // per-lane record of the speculative evaluation of "my" cell
            int recOc = OC_NONE, recL0 = 0;
            unsigned int recB0 = 0, recB1 = 0;
            // stash (at most one non-NOCHANGE speculative result per chunk)
            int stashLane = -1, stashOc = OC_NONE, stashN = 0;
            Rect stashRec; double stashNfa = 0;
            Rect rec; double nfa = 0; BBox bb; int nCommit = 0; bool touchedT = false;

            // ---------------- speculative phase
            unsigned int rem = liveMask;
            long long tSpec = clock64();
            while (rem && sh.frontier != chunk && !sh.abortFlag) {
                const int k = __ffs(rem) - 1;
                rem &= rem - 1;
                const int p = __shfl_sync(FULL, myp, k);
                if (lsdb_ld_state(&c.state[p]) & 3u) continue;
                const int L0 = sh.logCount;
                __threadfence_block();
                const int oc = eval_seed(c, p, stashLane < 0, rec, nfa, bb, nCommit, touchedT);
                STAT(c, ST_SPEC, 1);
                if (oc == OC_DEFER) break;
                if (lane == k) { recOc = oc; recL0 = L0; recB0 = pack_xy(bb.x0, bb.y0); recB1 = pack_xy(bb.x1, bb.y1); }
                if (oc != OC_NOCHANGE) { stashLane = k; stashOc = oc; stashN = nCommit; stashRec = rec; stashNfa = nfa; }
            }

            // ---------------- wait for every earlier chunk to retire
            long long tWait = clock64();
            if (lane == 0) atomicAdd(&sh.stats[TM_SPEC], (unsigned long long)(tWait - tSpec));
            if (lane == 0) {
                unsigned int spins = 0;
                while (sh.frontier != chunk && !sh.abortFlag) {
                    __nanosleep(64);
                    if (++spins > (1u << 26)) sh.abortFlag = LSDB_ERR_TIMEOUT;
                }
            }
            __syncwarp();
            __threadfence_block();
            if (sh.abortFlag) break;

            // ---------------- retire phase: in seed order, validate or re-evaluate, commit
            long long tRet = clock64();
            if (lane == 0) atomicAdd(&sh.stats[TM_WAIT], (unsigned long long)(tRet - tWait));
            rem = liveMask;
            while (rem) {
                const int k = __ffs(rem) - 1;
                rem &= rem - 1;
                const int p = __shfl_sync(FULL, myp, k);
                if (lsdb_ld_state(&c.state[p]) & 3u) continue;  // :222
                STAT(c, ST_LIVE, 1);
                const int oc = __shfl_sync(FULL, recOc, k);
                const int L0 = __shfl_sync(FULL, recL0, k);
                const unsigned int b0 = __shfl_sync(FULL, recB0, k), b1 = __shfl_sync(FULL, recB1, k);
                if (oc != OC_NONE && !has_conflict(c, L0, b0, b1)) {
                    if (oc == OC_NOCHANGE) continue;
                    if (k == stashLane) {
                        commit_region(c, stashOc, stashN, stashRec, stashNfa, lab, rc, maxSeg);
                        stashLane = -1;
                        continue;
                    }
                }
                if (k == stashLane) stashLane = -1;  // stale stash: bits are already cleared
                const int oc2 = eval_seed(c, p, true, rec, nfa, bb, nCommit, touchedT);
                if (touchedT && stashLane >= 0) stashLane = -2;  // the buffer of a later stash was overwritten
                STAT(c, ST_RESPEC, 1);
                if (oc2 == OC_REJECT || oc2 == OC_ACCEPT) commit_region(c, oc2, nCommit, rec, nfa, lab, rc, maxSeg);
                if (sh.abortFlag) break;
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) { atomicAdd(&sh.stats[ST_CHUNKS], 1ull); sh.frontier = chunk + 1; }
            if (lane == 0) atomicAdd(&sh.stats[TM_RETIRE], (unsigned long long)(clock64() - tRet));
        }
        __syncthreads();
        if (tid == 0) {
            dyn[img].nSeg = sh.nSeg;
            if (sh.abortFlag) dyn[img].err = sh.abortFlag;
            sh.stats[ST_CELLS] = nCells;
        }
        __syncthreads();
        if (tid < TM_N) dyn[img].stat[tid] = (long long)sh.stats[tid];
    }
}

__global__ void lsdb_used_plane_kernel(const unsigned int* __restrict__ state, uint8_t* __restrict__ used, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        unsigned int s = state[i];
        used[i] = (s & LSDB_ST_BAN) ? 1 : ((s & LSDB_ST_REJ) ? 2 : 0);
    }
}

void lsdb_launch_grow(cudaStream_t s, int nImgs, int nCtas, const LsdbImg* imgs, LsdbImgDyn* dyn,
                      const LsdbLsdConst* kc, const double* mag, const double* deg, unsigned int* state,
                      const unsigned int* cells, int* labels, LsdbRect* rects, int maxSeg,
                      unsigned int* lists, int listCap, const double* lgammaTab, int lgammaN, int* imgCounter) {
    if (nImgs > 0)
        lsdb_grow_kernel<<<nCtas, NW * 32, 0, s>>>(nImgs, imgs, dyn, kc, mag, deg, state, cells, labels, rects, maxSeg,
                                                   lists, listCap, lgammaTab, lgammaN, imgCounter);
}

void lsdb_launch_lgamma_table(cudaStream_t s, double* tab, int n) {
    lsdb_lgamma_table_kernel<<<(n + 127) / 128, 128, 0, s>>>(tab, n);
}

void lsdb_launch_used_plane(cudaStream_t s, const unsigned int* state, uint8_t* used, int n) {
    lsdb_used_plane_kernel<<<(n + 255) / 256, 256, 0, s>>>(state, used, n);
}

int lsdb_grow_max_ctas(int device) {
    int sms = 0, per = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, lsdb_grow_kernel, NW * 32, 0);
    if (per < 1) per = 1;
    return sms * per;
}
