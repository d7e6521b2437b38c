// Lane-serial core of the region stage (seed -> RegionGrower -> rectangle -> Refiner -> NFA), sm_100a.
//
// One LANE evaluates one seed, start to finish, as plain sequential code in the reference's order:
//   RegionGrower :491-590, CenterGetter/OrientationGetter/RectangleConverter :592-734, RegionRadiusReducer :736-802,
//   Refiner :804-880, LogGammaCalculator :882-924, RectangleNFACalculator :926-1059, RectangleImprover :1061-1158
// (all /root/reference/LSD/myLSD.cpp).  32 independent chains per warp replace the one warp-cooperative chain of round 1:
// the stage is bound by the latency of dependent gathers and of the correctly rounded double-double math, so the
// lever is the number of chains in flight, not the speed of one.
//
// Everything here is __host__ __device__ and touches memory through plain pointers only: tests/test_region_core.py
// compiles it with the host compiler and runs the reference's sequential seed loop on it against the oracle, so the
// arithmetic and control flow of the lanes are pinned without a GPU.  The concurrent part (speculation, parking,
// ordered retire) lives in region.cu.
#pragma once
#include "lsdb_common.cuh"

#if defined(__CUDA_ARCH__)
#define RG_HD __device__ __forceinline__
#define RG_HDN __device__ __noinline__
#define RG_FFS(x) __ffs((int)(x))
#define RG_LD_STATE(p) lsdb_ld_state(p)
#define RG_MARK_GROWING(M, spec, p) do { if ((spec) >= 0) atomicOr(&(M).state[p], LSDB_ST_GROWING); } while (0)
#define RG_LD_STATE_NB(p) __ldcg(p)   /* L2 load the compiler may schedule freely (no asm memory clobber): neighbour gathers */
#define RG_LD_BM(p) (*(const volatile unsigned int*)(p))
#define RG_FUNNEL_R(lo, hi, sh) __funnelshift_r((lo), (hi), (sh))
#define RG_CLOCK() clock64()
#define RG_ACTIVEMASK() __activemask()
#define RG_ANY(mask, p) __any_sync((mask), (p))
#define RG_POPC(x) __popc(x)
#else
#define RG_CLOCK() 0ll
#define RG_ACTIVEMASK() 1u
#define RG_ANY(mask, p) (p)
#define RG_POPC(x) __builtin_popcount(x)
#define RG_HD static inline
#define RG_HDN static
#define RG_FFS(x) __builtin_ffs((int)(x))
#define RG_LD_STATE(p) (*(p))
#define RG_LD_STATE_NB(p) (*(p))
#define RG_MARK_GROWING(M, spec, p) do { } while (0)
#define RG_LD_BM(p) (*(p))
#define RG_FUNNEL_R(lo, hi, sh) ((unsigned int)(((((unsigned long long)(hi)) << 32) | (unsigned long long)(lo)) >> (sh)))
#endif

enum { RG_OC_NONE = 0, RG_OC_NOCHANGE = 1, RG_OC_REJECT = 2, RG_OC_ACCEPT = 3, RG_OC_DEFER = 4 };

struct RgRect { double x1, y1, x2, y2, wid, cX, cY, deg, dx, dy, p, prec; };

// the map a lane works on (one copy per CTA in shared memory on the device)
struct RgMap {
    int W, H, pw, n;
    unsigned int* state;        // per-pixel state words (lsdb_common.cuh)
    const double* deg;
    const double* mag;
    const double* cs;           // (cos deg, sin deg) per pixel, interleaved
    unsigned int* bm;           // ban plane (usedMap == 1), one bit per pixel, pw words per row
    const LsdbLsdConst* kc;
    const double* lgammaTab;
    int lgammaN;
    int T;                      // ceil(regThre)
    double logNT, regThre;
};

// work buffers of one evaluating lane
struct RgLane {
    unsigned int* L0;           // first grow G1 (cap words each)
    unsigned int* L1;           // Refiner re-grow G2 (working copy: RegionRadiusReducer edits it), finally the commit list
    unsigned int* L2;           // G2 as grown
    unsigned short* rej;        // per listed point: the neighbours that failed the angle test (9-bit mask)
    unsigned int* pnd;          // pixels skipped because an earlier seed's parked accept covers them
    unsigned int* vis;          // private curMap: one bit per pixel, pw words per row; all zero between evaluations
    int cap, pndCap;
};

struct RgEval {
    int oc;                     // RG_OC_*
    int nG1, nG2, nCommit;      // accepted by the first grow / by the re-grow (0 = none) / pixels of the commit list
    int usedT;                  // the Refiner re-grew: commit list in L1, else L0
    int npnd;                   // entries of pnd (-1: overflow, the evaluation cannot be re-validated)
    int x0, y0, x1, y1;         // bounding box of every accepted pixel
    RgRect rec;
    double logNFA;
    // counters
    int nGrows, nGrownPx, nRegrow, nRrr, nNfa, nNfaPx;
    long long cycGrow, cycRect, cycNfa;   // lane clock64 deltas (device only)
    int nSteps, nStepLanes;               // grower steps of this lane, and the lanes that ran them side by side (summed)
#ifdef RG_COUNT_WORK
    long long nPass, nVisit, nLoadPts, nCand;
#endif
};

RG_HD unsigned int rg_pack(int x, int y) { return ((unsigned int)y << 16) | (unsigned int)x; }
RG_HD int rg_px(unsigned int v) { return (int)(v & 0xffffu); }
RG_HD int rg_py(unsigned int v) { return (int)(v >> 16); }

RG_HD int rg_x86_d2i(double v) {   // x86-64 cvttsd2si: NaN / out of range -> INT_MIN
    if (!(v > -2147483649.0 && v < 2147483648.0)) return (int)0x80000000;
    return (int)v;
}

// bits (x-1, x, x+1) of one row of the lane's private curMap plane, bit 0 = x-1 (columns outside the image read as 0)
RG_HD unsigned int rg_row3(const unsigned int* row, int x, int W, int pw) {
    unsigned int out;
    if (x == 0) out = (row[0] << 1) & 7u;
    else {
        const int xl = x - 1, wi = xl >> 5, sh = xl & 31;
        const unsigned int lo = row[wi];
        const unsigned int hi = (sh > 29 && wi + 1 < pw) ? row[wi + 1] : 0u;
        out = RG_FUNNEL_R(lo, hi, sh) & 7u;
    }
    if (x + 1 >= W) out &= 3u;
    return out;
}
// 3x3 neighbourhood in the reference's scan order (:533-535): bit (dy+1)*3 + (dx+1)
RG_HD unsigned int rg_vis9(const RgMap& M, const unsigned int* vis, int x, int y) {
    unsigned int r = 0;
    if (y - 1 >= 0) r |= rg_row3(vis + (size_t)(y - 1) * M.pw, x, M.W, M.pw);
    r |= rg_row3(vis + (size_t)y * M.pw, x, M.W, M.pw) << 3;
    if (y + 1 < M.H) r |= rg_row3(vis + (size_t)(y + 1) * M.pw, x, M.W, M.pw) << 6;
    return r;
}
RG_HD void rg_vis_set(const RgMap& M, unsigned int* vis, int x, int y) { vis[(size_t)y * M.pw + (x >> 5)] |= 1u << (x & 31); }
RG_HD void rg_vis_clr(const RgMap& M, unsigned int* vis, int x, int y) { vis[(size_t)y * M.pw + (x >> 5)] &= ~(1u << (x & 31)); }
RG_HD bool rg_vis_get(const RgMap& M, const unsigned int* vis, int x, int y) { return (vis[(size_t)y * M.pw + (x >> 5)] >> (x & 31)) & 1u; }
RG_HD void rg_vis_clear_list(const RgMap& M, unsigned int* vis, const unsigned int* lst, int n) {
    for (int k = 0; k < n; k++) rg_vis_clr(M, vis, rg_px(lst[k]), rg_py(lst[k]));
}

// does the parked mark in state word `st` come from a seed at or before chunk `myChunk` (window order mod 4096)?
RG_HD bool rg_pend_applies(unsigned int st, unsigned int kinds, int myChunk) {
    return (st & kinds) && myChunk >= 0 && (((unsigned int)myChunk - (st >> LSDB_ST_TAG_SHIFT)) & 4095u) < 2048u;
}

// noinline wrappers: one copy of each routine in the kernel
RG_HDN double rg_atan2(double y, double x) { return lsdm_atan2(y, x); }
RG_HDN double rg_cos(double x) { return lsdm_cos(x); }
RG_HDN double rg_sin(double x) { return lsdm_sin(x); }
RG_HDN double rg_log(double x) { return lsdm_log(x); }
RG_HDN double rg_log10(double x) { return lsdm_log10(x); }
RG_HDN double rg_exp(double x) { return lsdm_exp(x); }
RG_HDN double rg_pow(double x, double y) { return lsdm_pow(x, y); }

// ------------------------------------------------------------------ RegionGrower (:491-590), one lane
// Literal replay of the reference: FIFO list, 3x3 scan in row-major order, regDeg re-estimated after every accept,
// whole-list passes until one adds nothing.  Three exact short cuts:
//  * the angle test |regDeg - deg| < tol is decided on the running sums (dot^2 > cos^2(tol) |S|^2, no atan2 on the chain)
//    and falls back to the literal test when the margin is below 4e-13 relative (the reference's own rounding moves the
//    decision by < 2e-15) or when tol > 1.5 rad;
//  * a neighbour that is outside, banned or already in the region stays so, hence later passes re-test only the
//    neighbours that failed the ANGLE test (9-bit mask per listed point);
//  * the neighbours of a point are filtered by the ban plane and the lane's private curMap before their angle data
//    is fetched, and the fetches of up to four candidates are issued together (one memory round trip per point).
// Speculative evaluations (specChunk >= 0) count a pixel an EARLIER seed has parked for acceptance as banned and list it
// in pnd (re-checked when the evaluation retires).
// Returns the region size, or -1 when it outgrew B.cap (the lane's curMap bits of list[0..cap-1) are then still set).
// SMALL = the scout: no curMap plane (membership = search of the lane's own short list) and the growth stops as soon as
// the region reaches `stopAt` points — enough to know that it is not one of the ~97 % the reference drops at :228.
// The growth as a step machine: rg_grow_init, then rg_grow_step (ONE listed point per call) while G.running, then
// rg_grow_finish.  rg_lane_grow below runs it to completion; region.cu steps many of them side by side, one per lane.
struct RgGrow {
    unsigned int* list;
    int num, exNum, startNum, i, stop, np, cap;
    unsigned int v, rmask;           // the listed point being scanned; its neighbours to look at
    double cosS, sinS, n2, c2n2, m2, c2, regExact, regDeg0, degThre;
    int bx0, by0, bx1, by1;
    int specChunk, stopAt;
    bool nrmOK, haveExact, tauSmall, running;
};

template <bool SMALL>
RG_HD void rg_grow_init(const RgMap& M, const RgLane& B, RgGrow& G, unsigned int* list, int sx, int sy, double regDeg0, double degThre,
                        int specChunk, int stopAt, int npnd, const RgEval& ev) {
    G.list = list; G.regDeg0 = regDeg0; G.degThre = degThre; G.specChunk = specChunk; G.stopAt = stopAt; G.np = npnd; G.cap = B.cap;
    G.tauSmall = degThre <= 1.5;
    G.c2 = 0.0;
    if (G.tauSmall) { const double cTau = degThre == M.kc->degThre ? M.kc->cosDegThre : rg_cos(degThre); G.c2 = cTau * cTau; }
    const size_t sp = (size_t)sy * M.W + sx;
    G.cosS = M.cs[2 * sp]; G.sinS = M.cs[2 * sp + 1];   // cos(regDeg), sin(regDeg) with regDeg = deg[seed]  (:515-516)
    G.n2 = G.cosS * G.cosS + G.sinS * G.sinS; G.c2n2 = G.c2 * G.n2; G.m2 = 4e-13 * G.n2;
    G.nrmOK = G.n2 > 1e-18;
    G.haveExact = true;
    G.regExact = regDeg0;
    list[0] = rg_pack(sx, sy);
    B.rej[0] = 0;
    if (!SMALL) rg_vis_set(M, B.vis, sx, sy);
    G.num = 1; G.startNum = 0; G.stop = 0;
    G.exNum = 1;   // list length when the current pass started
    G.bx0 = ev.x0; G.by0 = ev.y0; G.bx1 = ev.x1; G.by1 = ev.y1;
    if (sx < G.bx0) G.bx0 = sx; if (sx > G.bx1) G.bx1 = sx; if (sy < G.by0) G.by0 = sy; if (sy > G.by1) G.by1 = sy;
    G.v = list[0]; G.rmask = 0x1efu; G.i = 0;
    G.running = true;
}

// One step = one listed point.  The two loops of the reference (passes over the list until one adds nothing, :525; the
// points of the list, :527) are one flat sequence of steps.  Structured control flow only: the lanes of a warp step side
// by side.  G.stop ends the growth: 1 = list overflow, 2 = the scout's region reached stopAt points.
// v = the listed point being scanned, rmask = its neighbours to look at: all eight on its first scan, in later passes
// the ones that failed the angle test.  Both are carried from one point to the next (the next point is either listed
// already — fetched a point ahead — or the first pixel accepted at this point).
template <bool SMALL>
RG_HD void rg_grow_step(const RgMap& M, const RgLane& B, RgGrow& G) {
    const int W = M.W, H = M.H;
    unsigned int* list = G.list;
    const int i = G.i;
    int num = G.num;
    const int startNum = G.startNum;
    const int x = rg_px(G.v), y = rg_py(G.v);
    const int numAt = num;
    unsigned int nv = 0, nmask = 0x1efu;
    if (i + 1 < numAt) { nv = list[i + 1]; if (i + 1 < startNum) nmask = B.rej[i + 1]; }
    unsigned int firstAcc = 0;
    unsigned int cm = G.rmask;
    if (i >= startNum) {   // neighbours outside the image (:536)
        if (y == 0) cm &= ~0x007u;
        if (y == H - 1) cm &= ~0x1c0u;
        if (x == 0) cm &= ~0x049u;
        if (x == W - 1) cm &= ~0x124u;
    }
    unsigned int nr = 0;
    int stop = 0;
    if (cm) {
        // round trip 1: the state words of the neighbours (usedMap, parked marks) and the lane's curMap rows
        unsigned int st8[8];
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int nb = t < 4 ? t : t + 1;
            st8[t] = LSDB_ST_BAN;
            if ((cm >> nb) & 1u) st8[t] = RG_LD_STATE_NB(&M.state[(size_t)(y + nb / 3 - 1) * W + (x + nb % 3 - 1)]);
        }
        if (!SMALL) cm &= ~rg_vis9(M, B.vis, x, y);   // in the region already
        else {
            unsigned int own = 0;
            for (int k = 0; k < num; k++) {
                const int ddx = rg_px(list[k]) - x + 1, ddy = rg_py(list[k]) - y + 1;
                if ((unsigned int)ddx < 3u && (unsigned int)ddy < 3u) own |= 1u << (ddy * 3 + ddx);
            }
            cm &= ~own;
        }
        // open = candidates: not in the region, usedMap != 1 (:537), not parked for acceptance by an earlier seed
        unsigned int open = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int nb = t < 4 ? t : t + 1;
            const unsigned int st = st8[t];
            if (((cm >> nb) & 1u) && !(st & LSDB_ST_BAN)) {
                if (rg_pend_applies(st, LSDB_ST_PACC, G.specChunk)) {
                    if (G.np >= 0 && G.np < B.pndCap) B.pnd[G.np++] = rg_pack(x + nb % 3 - 1, y + nb / 3 - 1); else G.np = -1;
                } else open |= 1u << nb;
            }
        }
        // round trip 2 (mostly L1 hits: the angle planes are immutable): angle data of the open candidates, four at a
        // time, then the decisions, in scan order
        const double pi = M.kc->pi, pi32 = pi * 3 / 2.0, pi2 = 2.0 * pi;
        while (open && !stop) {
            int nb4[4]; double cd4[4], sd4[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                nb4[t] = -1; cd4[t] = 0; sd4[t] = 0;
                if (open) {
                    const int nb = RG_FFS(open) - 1;
                    open &= open - 1;
                    const size_t p = (size_t)(y + nb / 3 - 1) * W + (x + nb % 3 - 1);
                    nb4[t] = nb; cd4[t] = M.cs[2 * p]; sd4[t] = M.cs[2 * p + 1];
                }
            }
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int nb = nb4[t];
                if (nb >= 0 && !stop) {
                    const int m = y + nb / 3 - 1, n = x + nb % 3 - 1;
                    const double cd = cd4[t], sd = sd4[t];
                    bool pass = false, unc = true;
                    if (G.tauSmall) {
                        const double dot = G.cosS * cd + G.sinS * sd;
                        const double d2 = dot * dot - G.c2n2;
                        pass = dot > 0 && d2 > 0;
                        unc = (dot > 0 && !(fabs(d2) > G.m2)) || !G.nrmOK;
                    }
                    if (unc) {   // the literal test, :540-543
                        if (!G.haveExact) { G.regExact = rg_atan2(G.sinS, G.cosS); G.haveExact = true; }
                        double degDif = fabs(G.regExact - M.deg[(size_t)m * W + n]);
                        if (degDif > pi32) degDif = fabs(degDif - pi2);
                        pass = degDif < G.degThre;
                    }
                    if (pass && num >= G.cap - 1) { stop = 1; pass = false; }
                    if (pass) {
                        const unsigned int pk = rg_pack(n, m);
                        if (num == numAt) firstAcc = pk;
                        list[num] = pk;
                        if (!SMALL) { rg_vis_set(M, B.vis, n, m); RG_MARK_GROWING(M, G.specChunk, (size_t)m * W + n); }
                        num++;
                        G.cosS += cd;   // :545-546
                        G.sinS += sd;
                        G.haveExact = false;
                        G.n2 = G.cosS * G.cosS + G.sinS * G.sinS;
                        G.c2n2 = G.c2 * G.n2; G.m2 = 4e-13 * G.n2; G.nrmOK = G.n2 > 1e-18;
                        if (n < G.bx0) G.bx0 = n; if (n > G.bx1) G.bx1 = n; if (m < G.by0) G.by0 = m; if (m > G.by1) G.by1 = m;
                        if (SMALL && num >= G.stopAt) stop = 2;
                    } else if (!stop) {
                        nr |= 1u << nb;
                    }
                }
            }
        }
    }
    B.rej[i] = (unsigned short)nr;
    if (i + 1 < numAt) { G.v = nv; G.rmask = nmask; }
    else { G.v = firstAcc; G.rmask = 0x1efu; }   // i + 1 == numAt: the first pixel accepted at this point, if any
    G.num = num;
    G.i = i + 1;
    if (stop) { G.stop = stop; G.running = false; }
    else if (G.i >= num) {   // end of a pass (:525): another one if this one added something
        G.startNum = num;
        if (num == G.exNum) G.running = false;
        else { G.exNum = num; G.i = 0; G.v = list[0]; G.rmask = B.rej[0]; }
    }
}

// region size (-1: it outgrew the list; the curMap bits of list[0..cap-1) are then still set); bounding box, pending count
RG_HD int rg_grow_finish(const RgGrow& G, int& npnd, double& regDegOut, RgEval& ev) {
    ev.x0 = G.bx0; ev.y0 = G.by0; ev.x1 = G.bx1; ev.y1 = G.by1;
    npnd = G.np;
    if (G.stop == 1) return -1;
    if (G.stop == 2) return G.num;
    regDegOut = G.num > 1 ? (G.haveExact ? G.regExact : rg_atan2(G.sinS, G.cosS)) : G.regDeg0;   // :547 after the last accept
    ev.nGrows++; ev.nGrownPx += G.num;
    return G.num;
}

// SMALL = the scout: no curMap plane (membership = search of the lane's own short list) and the growth stops as soon as
// the region reaches `stopAt` points — enough to know that it is not one of the ~97 % the reference drops at :228.
// The loop head is an explicit warp vote: the lanes that entered together meet there after every step, so that their
// gathers are in flight at the same time.
template <bool SMALL>
RG_HDN int rg_lane_grow(const RgMap& M, const RgLane& B, unsigned int* list, int sx, int sy, double regDeg0, double degThre,
                        int specChunk, int stopAt, int& npnd, double& regDegOut, RgEval& ev) {
    RgGrow G;
    rg_grow_init<SMALL>(M, B, G, list, sx, sy, regDeg0, degThre, specChunk, stopAt, npnd, ev);
    const unsigned int team = RG_ACTIVEMASK();
    while (RG_ANY(team, G.running)) {
        if (G.running) {
            ev.nSteps++; ev.nStepLanes += RG_POPC(RG_ACTIVEMASK());
            rg_grow_step<SMALL>(M, B, G);
        }
    }
    return rg_grow_finish(G, npnd, regDegOut, ev);
}

// ------------------------------------------------------------------ RectangleConverter (:592-734), sums in list order
RG_HDN RgRect rg_rect(const RgMap& M, const unsigned int* lst, int num, double regDeg, double aliPro, double degThre) {
    const int W = M.W;
    const double pi = M.kc->pi;
    double cenX = 0, cenY = 0, weiSum = 0;
    // the addends depend on gathers (list entry -> magnitude): fetched eight points at a time so that the round trips
    // overlap; the sums themselves are accumulated one by one, in list order, like the reference
    for (int k0 = 0; k0 < num; k0 += 8) {   // CenterGetter :608-613
        unsigned int v[8]; double w[8];
#pragma unroll
        for (int t = 0; t < 8; t++) v[t] = k0 + t < num ? lst[k0 + t] : 0u;
#pragma unroll
        for (int t = 0; t < 8; t++) w[t] = k0 + t < num ? M.mag[(size_t)rg_py(v[t]) * W + rg_px(v[t])] : 0.0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            if (k0 + t < num) {
                cenX += w[t] * rg_px(v[t]);
                cenY += w[t] * rg_py(v[t]);
                weiSum += w[t];
            }
        }
    }
    cenX = cenX / weiSum; cenY = cenY / weiSum;
    double Ixx = 0, Iyy = 0, Ixy = 0;
    weiSum = 0;
    for (int k0 = 0; k0 < num; k0 += 8) {   // OrientationGetter :637-643
        unsigned int v[8]; double w[8];
#pragma unroll
        for (int t = 0; t < 8; t++) v[t] = k0 + t < num ? lst[k0 + t] : 0u;
#pragma unroll
        for (int t = 0; t < 8; t++) w[t] = k0 + t < num ? M.mag[(size_t)rg_py(v[t]) * W + rg_px(v[t])] : 0.0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            if (k0 + t < num) {
                const double ey = rg_py(v[t]) - cenY, ex = rg_px(v[t]) - cenX;
                Ixx += w[t] * (ey * ey);
                Iyy += w[t] * (ex * ex);
                Ixy -= w[t] * ex * ey;
                weiSum += w[t];
            }
        }
    }
    Ixx /= weiSum; Iyy /= weiSum; Ixy /= weiSum;
    const double dI = Ixx - Iyy;
    const double lamb = (Ixx + Iyy - sqrt(dI * dI + 4 * Ixy * Ixy)) / 2.0;
    double inertiaDeg;
    if (fabs(Ixx) > fabs(Iyy)) inertiaDeg = rg_atan2(lamb - Ixx, Ixy);
    else inertiaDeg = rg_atan2(Ixy, lamb - Iyy);
    double regDif = inertiaDeg - regDeg;
    while (regDif <= -pi) regDif += 2 * pi;
    while (regDif > pi) regDif -= 2 * pi;
    if (regDif < 0) regDif = -regDif;
    if (regDif > degThre) inertiaDeg += pi;
    const double dx = rg_cos(inertiaDeg), dy = rg_sin(inertiaDeg);
    double lenMin = 0, lenMax = 0, widMin = 0, widMax = 0;   // :701-714
    for (int k = 0; k < num; k++) {
        const unsigned int v = lst[k];
        const double len = (rg_px(v) - cenX) * dx + (rg_py(v) - cenY) * dy;
        const double wid = -(rg_px(v) - cenX) * dy + (rg_py(v) - cenY) * dx;
        if (len < lenMin) lenMin = len;
        if (len > lenMax) lenMax = len;
        if (wid < widMin) widMin = wid;
        if (wid > widMax) widMax = wid;
    }
    RgRect r;
    r.x1 = cenX + lenMin * dx; r.y1 = cenY + lenMin * dy;
    r.x2 = cenX + lenMax * dx; r.y2 = cenY + lenMax * dy;
    r.wid = widMax - widMin;
    r.cX = cenX; r.cY = cenY; r.deg = inertiaDeg; r.dx = dx; r.dy = dy;
    r.p = aliPro; r.prec = degThre;
    if (r.wid < 1) r.wid = 1;
    return r;
}

RG_HD double rg_density(int num, const RgRect& r) {   // :757-758, :827
    const double ax = r.x1 - r.x2, ay = r.y1 - r.y2;
    return num / (sqrt(ax * ax + ay * ay) * r.wid);
}
RG_HD double rg_dist(int ox, int oy, double x, double y) {
    const double a = ox - x, b = oy - y;
    return sqrt(a * a + b * b);
}

// ------------------------------------------------------------------ LogGammaCalculator (:882-924)
RG_HDN double rg_log_gamma_calc(int x) {
    double val;
    if (x > 15) {
        const double xd = x;
        val = 0.918938533204673 + (xd - 0.5) * rg_log(xd) - xd +
              0.5 * xd * rg_log(xd * lsdm_sinh(1.0 / xd) + 1.0 / (810 * rg_pow(xd, 6)));
    } else {
        const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705,
                             1168.92649479, 83.8676043424, 2.50662827511};
        double a = (x + 0.5) * rg_log(x + 5.5) - (x + 5.5);
        double b = 0;
        for (int i = 0; i < 7; i++) {
            a -= rg_log(x + i);
            b += q[i] * rg_pow(x, i);
        }
        val = a + rg_log(b);
    }
    return val;
}
RG_HD double rg_log_gamma(const RgMap& M, int x) {
    if (x >= 0 && x < M.lgammaN) return M.lgammaTab[x];
    return rg_log_gamma_calc(x);
}

// ------------------------------------------------------------------ RectangleNFACalculator (:926-1059)
RG_HDN double rg_nfa(const RgMap& M, const RgRect& rec, RgEval& ev) {
    const int xLim = M.W, yLim = M.H;
    const double logNT = M.logNT;
    const double pi = M.kc->pi;
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
    double vX0, vX1, vX2, vX3, vY0, vY1, vY2, vY3;
    // rotate without dynamic indexing (keeps the vertices in registers)
    if (offset == 0) { vX0 = verX[0]; vX1 = verX[1]; vX2 = verX[2]; vX3 = verX[3]; vY0 = verY[0]; vY1 = verY[1]; vY2 = verY[2]; vY3 = verY[3]; }
    else if (offset == 1) { vX0 = verX[1]; vX1 = verX[2]; vX2 = verX[3]; vX3 = verX[0]; vY0 = verY[1]; vY1 = verY[2]; vY2 = verY[3]; vY3 = verY[0]; }
    else if (offset == 2) { vX0 = verX[2]; vX1 = verX[3]; vX2 = verX[0]; vX3 = verX[1]; vY0 = verY[2]; vY1 = verY[3]; vY2 = verY[0]; vY3 = verY[1]; }
    else { vX0 = verX[3]; vX1 = verX[0]; vX2 = verX[1]; vX3 = verX[2]; vY0 = verY[3]; vY1 = verY[0]; vY2 = verY[1]; vY3 = verY[2]; }

    int allPixNum = 0, aliPixNum = 0;
    const int xr = rg_x86_d2i(ceil(vX0) - floor(vX2));
    const int xRang_len = (xr == (int)0x80000000 ? xr : (xr < 0 ? -xr : xr)) + 1;
    if (xRang_len > 0 && xRang_len < 100000000) {
        const double x0c = ceil(vX0);
        const double k0 = (vY1 - vY0) / (vX1 - vX0);
        const double k1 = (vY2 - vY1) / (vX2 - vX1);
        const double k2 = (vY2 - vY3) / (vX2 - vX3);
        const double k3 = (vY3 - vY0) / (vX3 - vX0);
        for (int i = 0; i < xRang_len; i++) {
            const int xi = rg_x86_d2i(i + x0c);
            // the reference fills yLow/yHigh with two partition passes (:987-1004); xi is increasing, so entry i
            // takes the first branch iff xi < vertex (NaN vertex: the slot stays 0)
            int yl = 0, yh = 0;
            if (xi < vX3) yl = rg_x86_d2i(ceil(vY0 + (xi - vX0) * k3));
            else if (xi >= vX3) yl = rg_x86_d2i(ceil(vY3 + (xi - vX3) * k2));
            if (xi < vX1) yh = rg_x86_d2i(floor(vY0 + (xi - vX0) * k0));
            else if (xi >= vX1) yh = rg_x86_d2i(floor(vY1 + (xi - vX1) * k1));
            if (xi < 0 || xi >= xLim) continue;
            const int j0 = yl < 0 ? 0 : yl, j1 = yh > yLim - 1 ? yLim - 1 : yh;
            for (int j = j0; j <= j1; j += 4) {   // counts are order-free: four gathers in flight
                double dv[4];
#pragma unroll
                for (int t = 0; t < 4; t++) dv[t] = j + t <= j1 ? M.deg[(size_t)(j + t) * xLim + xi] : 0.0;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    if (j + t <= j1) {
                        allPixNum++;
                        double degDif = fabs(rec.deg - dv[t]);
                        if (degDif > pi32) degDif = fabs(degDif - pi2);
                        if (degDif < rec.prec) aliPixNum++;
                    }
                }
            }
        }
    }
    ev.nNfa++; ev.nNfaPx += allPixNum;

    if (allPixNum == 0 || aliPixNum == 0) return -logNT;
    // log(p), log(1-p), log10(p): p only takes the values aliPro/2^k (:1085,:1149) — host-made table
    double logP, log1mP, log10P;
    {
        int k = -1;
        for (int i = 0; i < LSDB_NP; i++) if (rec.p == M.kc->pTab[i]) k = i;
        if (k >= 0) { logP = M.kc->logP[k]; log1mP = M.kc->log1mP[k]; log10P = M.kc->log10P[k]; }
        else { logP = rg_log(rec.p); log1mP = rg_log(1 - rec.p); log10P = rg_log10(rec.p); }
    }
    if (allPixNum == aliPixNum) return -logNT - allPixNum * log10P;
    const double proTerm = rec.p / (1.0 - rec.p);
    const double log1Coef = rg_log_gamma(M, allPixNum + 1) - rg_log_gamma(M, aliPixNum + 1) - rg_log_gamma(M, allPixNum - aliPixNum + 1);
    const double log1Term = log1Coef + aliPixNum * logP + (allPixNum - aliPixNum) * log1mP;
    double term = rg_exp(log1Term);
    const double eps = 2.2204e-16;
    if (fabs(term) < 100 * eps) {
        if (aliPixNum > allPixNum * rec.p) return -rg_log10(term) - logNT;
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
            // break test of :1052-1054.  It is a comparison: first decided with the platform pow/log10 (<= 2 ulp) and a
            // 1e-9 safety margin; only a knife-edge falls back to the correctly rounded pow/log10 the reference's
            // arithmetic is defined by.
            const double nn = (double)(allPixNum - i + 1);
            const double X = (1 - pow(multTerm, nn)) / (1.0 - multTerm);
            const double errA = term * (X - 1);
            const double l10 = log10(binTail);
            const double rhsA = tole * fabs(-l10 - logNT) * binTail;
            const double scale = fabs(term) * (fabs(X) + 1) + tole * fabs(binTail) * (fabs(l10) + fabs(logNT));
            bool brk;
            if (fabs(errA - rhsA) > 1e-9 * scale) brk = errA < rhsA;
            else {
                const double err = term * ((1 - rg_pow(multTerm, nn)) / (1.0 - multTerm) - 1);
                brk = err < tole * fabs(-rg_log10(binTail) - logNT) * binTail;
            }
            if (brk) break;
        }
    }
    return -rg_log10(binTail) - logNT;
}

// ------------------------------------------------------------------ RectangleImprover (:1061-1158)
RG_HDN double rg_improve(const RgMap& M, RgRect& rec, RgEval& ev) {
    const double pi = M.kc->pi;
    const double delt = 0.5, delt2 = delt / 2.0;
    double best = rg_nfa(M, rec, ev);
    RgRect bestRec = rec;
    if (best > 0) return best;
    RgRect r = bestRec;
    double v;
    for (int i = 0; i < 5; i++) {
        r.p /= 2.0; r.prec = r.p * pi;
        v = rg_nfa(M, r, ev);
        if (v > best) { best = v; bestRec = r; }
    }
    if (best > 0) { rec = bestRec; return best; }
    for (int side = 0; side < 3; side++) {   // 0: width, 1: side one, 2: side two  (:1096-1143)
        r = bestRec;
        for (int i = 0; i < 5; i++) {
            if (r.wid - delt >= 0.5) {
                if (side == 1) { r.x1 -= r.dy * delt2; r.y1 += r.dx * delt2; r.x2 -= r.dy * delt2; r.y2 += r.dx * delt2; }
                if (side == 2) { r.x1 += r.dy * delt2; r.y1 -= r.dx * delt2; r.x2 += r.dy * delt2; r.y2 -= r.dx * delt2; }
                r.wid -= delt;
                v = rg_nfa(M, r, ev);
                if (v > best) { best = v; bestRec = r; }
            }
        }
        if (best > 0) { rec = bestRec; return best; }
    }
    r = bestRec;
    for (int i = 0; i < 5; i++) {
        r.p /= 2.0; r.prec = r.p * pi;
        v = rg_nfa(M, r, ev);
        if (v > best) { best = v; bestRec = r; }
    }
    rec = bestRec;
    return best;
}

// ------------------------------------------------------------------ one seed: grow -> rectangle -> Refiner -> NFA  (:225-250)
// As a job that alternates between growth steps and the sequential work between two growths:
//   rg_job_begin;  while (J.g.running) rg_grow_step<false>;  rg_job_after_grow -> true: a re-grow was started, step again
// The lane's private curMap is all zero again when the job has finished.
// Lists when finished:  G1 = B.L0[0..nG1)   G2 = B.L2[0..nG2) (if usedT)   commit list = (usedT ? B.L1 : B.L0)[0..nCommit)
struct RgJob {
    RgGrow g;
    RgEval ev;
    RgRect rec;
    double regDeg, degThre2;
    int p0, specChunk, stage, npnd;   // stage: 1 = first grow, 2 = the Refiner's re-grow
};

RG_HD void rg_job_begin(const RgMap& M, const RgLane& B, RgJob& J, int p0, int specChunk) {
    RgEval& ev = J.ev;
    ev.oc = RG_OC_NOCHANGE; ev.nG1 = 0; ev.nG2 = 0; ev.nCommit = 0; ev.usedT = 0; ev.npnd = 0; ev.logNFA = 0;
    ev.x0 = ev.y0 = 0x7fffffff; ev.x1 = ev.y1 = -1;
    ev.nGrows = ev.nGrownPx = ev.nRegrow = ev.nRrr = ev.nNfa = ev.nNfaPx = 0;
    ev.cycGrow = ev.cycRect = ev.cycNfa = 0; ev.nSteps = ev.nStepLanes = 0;
    J.p0 = p0; J.specChunk = specChunk; J.stage = 1; J.npnd = 0;
    J.regDeg = M.deg[p0];
    rg_grow_init<false>(M, B, J.g, B.L0, p0 % M.W, p0 / M.W, J.regDeg, M.kc->degThre, specChunk, 0, 0, ev);
}

// Written as a sequence of phases guarded by flags instead of early returns: the lanes of a warp run this side by side, and
// a lane that skips a phase must meet the others again right after it.
RG_HDN bool rg_job_after_grow(const RgMap& M, const RgLane& B, RgJob& J) {
    const LsdbLsdConst* kc = M.kc;
    const int W = M.W;
    const int sx = J.p0 % W, sy = J.p0 / W;
    RgEval& ev = J.ev;
    long long tc = RG_CLOCK();
    double regDeg = J.regDeg;
    int npnd = J.npnd;
    int num = rg_grow_finish(J.g, npnd, regDeg, ev);
    J.npnd = npnd; ev.npnd = npnd;
    bool go = true;        // the evaluation is still running
    bool again = false;    // a re-grow has been started
    RgRect rec = J.rec;
    if (J.stage == 1) {
        if (num < 0) { rg_vis_clear_list(M, B.vis, B.L0, B.cap - 1); ev.oc = RG_OC_DEFER; go = false; }
        else {
            ev.nG1 = num;
            if (num < M.regThre) { rg_vis_clear_list(M, B.vis, B.L0, num); go = false; }   // :228
        }
        double den = 0;
        if (go) {
            rec = rg_rect(M, B.L0, num, regDeg, kc->aliPro, kc->degThre);
            den = rg_density(num, rec);
        }
        if (go && !(den >= kc->denThre)) {   // Refiner :804-880: new tolerance, re-grow from the seed
            const double pi = kc->pi;
            const double cenDeg = M.deg[J.p0];
            double difSum = 0, squSum = 0;
            int ptNum = 0;
            for (int k = 0; k < num; k++) {   // :839-853
                const unsigned int v = B.L0[k];
                if (rg_dist(sx, sy, (double)rg_px(v), (double)rg_py(v)) < rec.wid) {
                    double dd = M.deg[(size_t)rg_py(v) * W + rg_px(v)] - cenDeg;
                    while (dd <= -pi) dd += 2 * pi;
                    while (dd > pi) dd -= 2 * pi;
                    difSum += dd;
                    squSum += dd * dd;
                    ptNum++;
                }
            }
            const double meanDif = difSum / (ptNum * 1.0);
            J.degThre2 = 2.0 * sqrt((squSum - 2 * meanDif * difSum) / (ptNum * 1.0) + meanDif * meanDif);
            rg_vis_clear_list(M, B.vis, B.L0, num);
            J.regDeg = cenDeg; J.rec = rec; J.stage = 2;
            ev.nRegrow++;
            rg_grow_init<false>(M, B, J.g, B.L1, sx, sy, cenDeg, J.degThre2, J.specChunk, 0, npnd, ev);
            again = true; go = false;
        }
    } else {
        if (num < 0) { rg_vis_clear_list(M, B.vis, B.L1, B.cap - 1); ev.oc = RG_OC_DEFER; go = false; }
        else {
            for (int k = 0; k < num; k++) B.L2[k] = B.L1[k];
            ev.usedT = 1; ev.nG2 = num;
            if (num < 2) { rg_vis_clear_list(M, B.vis, B.L2, ev.nG2); go = false; }   // :861-864
        }
        if (go) {
            rec = rg_rect(M, B.L1, num, regDeg, rec.p, rec.prec);
            double d2 = rg_density(num, rec);
            if (d2 < kc->denThre) {
                // RegionRadiusReducer :736-802, in place, with the `i <= num` quirk
                bool ok = true;
                if (!(d2 > kc->denThre)) {
                    const double rad1 = rg_dist(sx, sy, rec.x1, rec.y1), rad2 = rg_dist(sx, sy, rec.x2, rec.y2);
                    double rad = rad1 > rad2 ? rad1 : rad2;
                    while (ok && d2 < kc->denThre) {
                        rad *= 0.75;
                        int i = 0, nn = num;
                        B.L1[nn] = 0u;   // slot [num] reads as (0,0)
                        while (i <= nn && nn > 0) {   // nn <= 0: the reference would index [-1] here (heap underflow, UB)
                            const unsigned int v = B.L1[i];
                            if (rg_dist(sx, sy, (double)rg_px(v), (double)rg_py(v)) > rad) {
                                rg_vis_clr(M, B.vis, rg_px(v), rg_py(v));
                                B.L1[i] = B.L1[nn - 1];
                                B.L1[nn - 1] = 0u;
                                i--;
                                nn--;
                            }
                            i++;
                        }
                        num = nn;
                        ev.nRrr++;
                        if (num < 2) ok = false;
                        else {
                            rec = rg_rect(M, B.L1, num, regDeg, rec.p, rec.prec);
                            d2 = rg_density(num, rec);
                        }
                    }
                }
                if (!ok) { rg_vis_clear_list(M, B.vis, B.L2, ev.nG2); go = false; }
            }
        }
    }
    ev.cycRect += RG_CLOCK() - tc; tc = RG_CLOCK();
    if (go) {
        const double logNFA = rg_improve(M, rec, ev);
        // commit list = pixels whose curMap bit is still set (what :242-248 / :259-265 visit); clear the bits
        if (!ev.usedT) {
            rg_vis_clear_list(M, B.vis, B.L0, ev.nG1);
            ev.nCommit = ev.nG1;
        } else {
            int outN = 0;
            for (int k = 0; k < ev.nG2; k++) {
                const unsigned int v = B.L2[k];
                if (rg_vis_get(M, B.vis, rg_px(v), rg_py(v))) { B.L1[outN++] = v; rg_vis_clr(M, B.vis, rg_px(v), rg_py(v)); }
            }
            ev.nCommit = outN;
        }
        ev.rec = rec; ev.logNFA = logNFA;
        ev.oc = logNFA <= 0 ? RG_OC_REJECT : RG_OC_ACCEPT;
        ev.cycNfa += RG_CLOCK() - tc;
    }
    return again;
}

// the whole evaluation of one seed by one lane (the lanes of a warp that enter together step side by side)
RG_HDN void rg_eval_lane(const RgMap& M, const RgLane& B, int p0, int specChunk, RgEval& evOut) {
    RgJob J;
    rg_job_begin(M, B, J, p0, specChunk);
    const unsigned int team = RG_ACTIVEMASK();
    bool active = true;
    while (RG_ANY(team, active)) {
        if (active) {
            if (J.g.running) { J.ev.nSteps++; J.ev.nStepLanes += RG_POPC(RG_ACTIVEMASK()); rg_grow_step<false>(M, B, J.g); }
            else active = rg_job_after_grow(M, B, J);
        }
    }
    evOut = J.ev;
}
