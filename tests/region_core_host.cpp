// TEST INFRASTRUCTURE: the lane-serial core of the region stage (csrc/region_core.h) compiled for the HOST and driven
// by the reference's sequential seed loop (LSD/myLSD.cpp:218-272).  tests/test_region_core.py compares the result with
// the oracle, so the lanes' arithmetic and control flow are pinned without a GPU.
#include "region_core.h"
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <vector>

extern "C" int rgcore_lsd(int W, int H, const double* mag, const double* deg, const int* seeds /* pixel index, seed order */, int nSeeds,
                          double sca, double angThre, double denThre, int laneCap, int useScout, unsigned char* usedOut, int* labelsOut,
                          double* rectsOut /* [max][13] */, int maxRects, long long* statsOut /* live, grows, grownpx, small, regrows, rrr, nfa, nfapx, rejects, accepts */) {
    LsdbLsdConst kc;
    memset(&kc, 0, sizeof kc);
    const double pi = 4.0 * lsdm_atan(1.0);
    kc.sca = sca; kc.pi = pi;
    kc.degThre = angThre / 180.0 * pi;
    kc.gradThre = 2.0 / lsdm_sin(kc.degThre);
    kc.aliPro = angThre / 180.0;
    kc.denThre = denThre;
    kc.cosDegThre = lsdm_cos(kc.degThre);
    { double p = kc.aliPro; for (int k = 0; k < LSDB_NP; k++, p /= 2.0) { kc.pTab[k] = p; kc.logP[k] = lsdm_log(p); kc.log1mP[k] = lsdm_log(1 - p); kc.log10P[k] = lsdm_log10(p); } }
    const size_t n = (size_t)W * H;
    const int pw = (W + 31) / 32;
    std::vector<unsigned int> state(n, 0u), bm((size_t)H * pw, 0u), vis((size_t)H * pw, 0u);
    std::vector<double> cs(2 * n, 0.0);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const size_t p = (size_t)y * W + x;
            const bool ban = y >= 1 && x >= 1 && mag[p] < kc.gradThre;   // :165-166 (row 0 / col 0 stay 0)
            if (ban) { state[p] = LSDB_ST_BAN; bm[(size_t)y * pw + (x >> 5)] |= 1u << (x & 31); }
            else { cs[2 * p] = lsdm_cos(deg[p]); cs[2 * p + 1] = lsdm_sin(deg[p]); }
        }
    std::vector<double> lg(1 << 12);
    for (size_t i = 0; i < lg.size(); i++) lg[i] = i >= 1 ? rg_log_gamma_calc((int)i) : 0.0;
    RgMap M;
    M.W = W; M.H = H; M.pw = pw; M.n = (int)n; M.state = state.data(); M.deg = deg; M.mag = mag; M.cs = cs.data(); M.bm = bm.data();
    M.kc = &kc; M.lgammaTab = lg.data(); M.lgammaN = (int)lg.size();
    M.logNT = 5 * (lsdm_log10(H) + lsdm_log10(W)) / 2.0;
    M.regThre = -M.logNT / lsdm_log10(angThre / 180.0);
    M.T = (int)ceil(M.regThre);
    const int cap = laneCap > 0 ? laneCap : (int)n + 2;
    std::vector<unsigned int> L0(cap + 1), L1(cap + 1), L2(cap + 1), pnd(64);
    std::vector<unsigned short> rej(cap + 1);
    RgLane B;
    B.L0 = L0.data(); B.L1 = L1.data(); B.L2 = L2.data(); B.rej = rej.data(); B.pnd = pnd.data(); B.vis = vis.data(); B.cap = cap; B.pndCap = 64;
    long long st[10] = {0};
    int nSeg = 0;
    if (labelsOut) memset(labelsOut, 0, n * sizeof(int));
    for (int i = 0; i < nSeeds; i++) {
        const int p0 = seeds[i];
        if (state[p0] & 3u) continue;   // :222
        st[0]++;
        RgEval ev;
        if (useScout) {   // the scout the GPU runs first: same growth, private short list, stops at T points
            unsigned int slst[40], spnd[8]; unsigned short srej[40];
            RgLane S; S.L0 = slst; S.L1 = 0; S.L2 = 0; S.rej = srej; S.pnd = spnd; S.vis = 0; S.cap = 40; S.pndCap = 8;
            RgEval sev; sev.x0 = sev.y0 = 0x7fffffff; sev.x1 = sev.y1 = -1; sev.nGrows = sev.nGrownPx = 0;
            int snp = 0; double rd;
            const int sn = M.T <= 32 ? rg_lane_grow<true>(M, S, slst, p0 % W, p0 / W, deg[p0], kc.degThre, -1, M.T, snp, rd, sev) : M.T;
            if (sn < M.T) { st[1]++; st[2] += sn; st[3]++; continue; }
        }
        rg_eval_lane(M, B, p0, -1, ev);
#ifdef RG_COUNT_WORK
        if (ev.nG1 >= M.regThre) { static long long tp=0,tv=0,tl=0,tc=0,te=0,tpx=0; tp+=ev.nPass; tv+=ev.nVisit; tl+=ev.nLoadPts; tc+=ev.nCand; te++; tpx+=ev.nGrownPx; if (getenv("RG_PRINT") && (te%200==0)) fprintf(stderr,"large evals %lld: px %lld passes %lld visits %lld loadpts %lld cands %lld\n",te,tpx,tp,tv,tl,tc); }
#endif
        st[1] += ev.nGrows; st[2] += ev.nGrownPx; st[4] += ev.nRegrow; st[5] += ev.nRrr; st[6] += ev.nNfa; st[7] += ev.nNfaPx;
        if (ev.oc == RG_OC_DEFER) return -1;
        if (ev.oc == RG_OC_NOCHANGE) { if (ev.nG1 < M.regThre) st[3]++; continue; }
        const unsigned int* px = ev.usedT ? B.L1 : B.L0;
        if (ev.oc == RG_OC_REJECT) {
            st[8]++;
            for (int k = 0; k < ev.nCommit; k++) state[(size_t)rg_py(px[k]) * W + rg_px(px[k])] |= LSDB_ST_REJ;
            continue;
        }
        st[9]++;
        for (int k = 0; k < ev.nCommit; k++) {
            const int x = rg_px(px[k]), y = rg_py(px[k]);
            state[(size_t)y * W + x] |= LSDB_ST_BAN;
            bm[(size_t)y * pw + (x >> 5)] |= 1u << (x & 31);
            if (labelsOut) labelsOut[(size_t)y * W + x] += nSeg + 1;
        }
        if (nSeg < maxRects && rectsOut) {
            double* R = rectsOut + (size_t)nSeg * 13;
            double rx1 = ev.rec.x1, ry1 = ev.rec.y1, rx2 = ev.rec.x2, ry2 = ev.rec.y2, rw = ev.rec.wid;
            if (sca != 1) {   // :252-258
                rx1 = (rx1 - 1.0) / sca + 1; ry1 = (ry1 - 1.0) / sca + 1;
                rx2 = (rx2 - 1.0) / sca + 1; ry2 = (ry2 - 1.0) / sca + 1;
                rw = (rw - 1.0) / sca + 1;
            }
            R[0] = rx1; R[1] = ry1; R[2] = rx2; R[3] = ry2; R[4] = rw; R[5] = ev.rec.cX; R[6] = ev.rec.cY; R[7] = ev.rec.deg;
            R[8] = ev.rec.dx; R[9] = ev.rec.dy; R[10] = ev.rec.p; R[11] = ev.rec.prec; R[12] = ev.logNFA;
        }
        nSeg++;
    }
    // the private curMap must be all zero between evaluations
    for (size_t i = 0; i < vis.size(); i++) if (vis[i]) return -2;
    if (usedOut) for (size_t p = 0; p < n; p++) usedOut[p] = (state[p] & LSDB_ST_BAN) ? 1 : ((state[p] & LSDB_ST_REJ) ? 2 : 0);
    if (statsOut) memcpy(statsOut, st, sizeof st);
    return nSeg;
}
