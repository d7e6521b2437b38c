// Stage C of the LSD hot path on sm_100a: the seed loop — region growing, rectangle fit, density refinement, NFA
// validation — as ONE persistent kernel with speculative, order-preserving commit.
//
// Replaces (reference, /root/reference/LSD/myLSD.cpp): the sequential seed loop :218-272 and everything it calls
// (:491-1158).  The arithmetic of one seed is region_core.h: plain sequential code run by ONE LANE, so a warp carries
// up to 32 independent evaluations instead of one (round 1 ran one warp-cooperative chain per warp and was bound by
// the latency of that chain: 10 % of the warp slots busy, 22 % of the stall samples waiting for instructions of a
// 390 KB kernel).
//
// Exactness.  The reference's result depends on the order seeds are visited (usedMap evolves).  Kept by
// speculate -> park -> retire in seed order:
//   * one CTA per map; the sorted seed list is cut into chunks of 32 cells; a WORKER warp claims a super-chunk of up
//     to 8 chunks by ticket and may run a bounded window ahead of the commit frontier;
//   * scouting, one seed per lane (rg_lane_grow<true>): decides the ~97 % of live seeds whose region stays below
//     regThre pixels — the reference drops those without touching any state (:228) — and flags the rest as large;
//   * large seeds, one seed per lane, RG_P lanes at a time (rg_eval_lane): grow, rectangle, Refiner / radius
//     reducer, NFA.  Each evaluating lane owns three lists and a private curMap bit plane;
//   * every evaluation is parked in the record arena of its super-chunk: the outcome, the pixels it accepted, the
//     pixels it skipped because an EARLIER seed's parked accept covers them, the rectangle.  A parked accept / reject
//     marks its pixels in the state words so that later seeds speculate as if it had been committed already;
//   * warp 0 is the FRONTIER warp: it retires READY chunks in order.  A parked evaluation stands iff (V1) every
//     pixel it accepted is still un-banned — bans only grow, and a candidate rejected by angle stays out whether or
//     not it is banned later, so the evaluation then replays identically — and (V2) every pixel it took for banned
//     because of a parked accept is banned by now.  V1 is pre-filtered by a coarse "last accept" grid.  Seeds whose
//     evaluation is missing or invalid are evaluated again at the frontier, several lanes at a time (no-change
//     outcomes commute; the first outcome that changes the state ends the batch);
//   * commits (usedMap 1 / 2, labels, rectangle record) happen only at the frontier.
// So usedMap, labels and the segment list are exactly the sequential result, whatever the timing.
#include "region_core.h"
#include "../../include/lsdb200.h"
#include <stdlib.h>

#define NW_MAX LSDB_GROW_WARPS
#define ARENA_HDR 32       // words: 13 doubles (rect + logNFA), [26] nCommit, [27] outcome, [28] offset of the commit list
#define GRID 32            // coarse cells per axis of the accept grid
#define FULL 0xffffffffu
#define RING 1024          // chunks a CTA may run ahead of the commit frontier
#define NSLOTS (RING / LSDB_SUPER)
#define RG_P 8             // lanes of a warp that evaluate large seeds at once
#define RG_LANE_CAP 8192   // their list capacity (points); a region that outgrows it is evaluated on the CTA's big lane set
#define RG_PND_CAP 1024
#define SG_CAP 32          // the scout handles regThre <= SG_CAP
#define SG_PND 24
#define QWORDS (2 * 32 * LSDB_SUPER)

enum { ST_CELLS = 0, ST_LIVE, ST_GROWS, ST_GROWNPX, ST_SMALL, ST_REGROWS, ST_RRR, ST_NFACALLS, ST_NFAPX, ST_REJECTS,
       ST_ACCEPTS, ST_SPEC, ST_RESPEC, ST_CHUNKS, ST_N,
       TM_GROW = ST_N, TM_RECT, TM_NFA, TM_WAIT, TM_RETIRE, TM_SPEC, TM_RESPEC, TM_MAPCYC, TM_MAPNS, ST_ROUNDS, RS_NONE, RS_CONFLICT, RS_COMMIT, RS_LOST, RS_REQUEUE, RS_DROPPED, RS_BIG, TM_LANEGROW, TM_N };
// TM_GROW: warp cycles in the large-seed rounds; TM_LANEGROW / TM_RECT / TM_NFA: LANE cycles inside RegionGrower / rectangle+Refiner / NFA

struct RgShared {
    RgMap M;
    volatile int frontier;     // first chunk not yet retired
    int nextChunk;             // ticket counter
    int runAhead;              // chunks the claims may lead the commit frontier
    int supShift;              // a claim covers 1 << supShift chunks (<= LSDB_SUPER)
    volatile int nSeg;
    volatile int abortFlag;
    int img, nChunks, nCells, cellShift;
    volatile int chunkFlag[RING];            // 1 = evaluated, records parked
    unsigned int slotHead[NSLOTS];           // bump pointer of the record arena of every super-chunk in flight
    unsigned int grid[GRID * GRID];          // per coarse cell: 1 + index of the LAST accepted region that touched the cell
    unsigned long long stats[TM_N];
};

// per-seed record of a parked evaluation (one per lane of a chunk), SoA in global memory: [RING][32]
struct ChunkRecs {
    int* oc; int* L0; unsigned int* b0; unsigned int* b1; unsigned int* off; int* chk; unsigned int* chkOff; unsigned int* pndOff; int* pnd;
};
#define REC_BYTES_PER_CELL 40

struct LsdbRegionParams {
    int nImgs;
    const LsdbImg* imgs; LsdbImgDyn* dyn; const LsdbLsdConst* kc;
    const double* mag; const double* deg; const double* cs;
    unsigned int* state; const unsigned int* cells; int* labels; LsdbRect* rects; int maxSeg;
    unsigned int* work; size_t workWordsPerCta; int listCap, arenaCap, runAhead;
    unsigned int* vis; size_t planeWords;
    unsigned char* recBuf; const double* lgammaTab; int lgammaN; int* imgCounter; unsigned int* banBits; int bmCapWords; int flags;
};

__host__ __device__ inline size_t rg_lane_words(int cap) { return 3 * (size_t)(cap + 2) + (size_t)(cap + 2) / 2 + 1 + RG_PND_CAP; }
__host__ __device__ inline size_t rg_warp_words() { return (1 + RG_P) * (size_t)QWORDS + RG_P * ((rg_lane_words(RG_LANE_CAP) + 1) & ~(size_t)1); }
__host__ __device__ inline size_t rg_cta_words(int listCap, int arenaCap, int nw) {
    return (size_t)NSLOTS * arenaCap + (size_t)nw * rg_warp_words() + ((rg_lane_words(listCap) + 1) & ~(size_t)1);
}

struct WarpCtx {
    int lane, w;
    unsigned int* arenas; int arenaCap;
    unsigned int* q;        // live cells of the super-chunk being scouted
    unsigned int* lq;       // large seeds found by the scouts: RG_P lists (one per claimed super-chunk) of QWORDS words: pixel, cell
    RgLane L;               // evaluation buffers of this lane (lanes < RG_P)
    RgLane big;             // the CTA's one full-size lane set (frontier warp, lane 0)
    int* labels; LsdbRect* rects; int maxSeg;
    const unsigned int* cells;
    ChunkRecs R;
};

#define STAT(sh, idx, v) atomicAdd(&(sh).stats[idx], (unsigned long long)(v))

__device__ __forceinline__ RgLane rg_lane_set(unsigned int* base, int cap, unsigned int* vis) {
    RgLane B;
    B.L0 = base; B.L1 = base + (cap + 2); B.L2 = base + 2 * (size_t)(cap + 2);
    B.rej = reinterpret_cast<unsigned short*>(base + 3 * (size_t)(cap + 2));
    B.pnd = base + 3 * (size_t)(cap + 2) + (size_t)(cap + 2) / 2 + 1;
    B.vis = vis; B.cap = cap; B.pndCap = RG_PND_CAP;
    return B;
}

__device__ __forceinline__ bool ban_at(const RgMap& M, int x, int y) {
    return (RG_LD_BM(M.bm + (size_t)y * M.pw + (x >> 5)) >> (x & 31)) & 1u;
}
__device__ __forceinline__ bool seed_taken(unsigned int st, int chunk) {
    return (st & 3u) != 0 || rg_pend_applies(st, LSDB_ST_PACC | LSDB_ST_PREJ, chunk);
}
__device__ __forceinline__ int slot_of_chunk(const RgShared& sh, int chunk) { return (chunk >> sh.supShift) & (NSLOTS - 1); }

// mark the pixels of a parked accept / reject candidate (kind = LSDB_ST_PACC / LSDB_ST_PREJ); the earliest chunk tag wins
__device__ void park_pixels(const RgMap& M, int lane, const unsigned int* px, int n, unsigned int kind, int chunk) {
    const unsigned int tag = (unsigned int)chunk & 4095u;
    for (int k = lane; k < n; k += 32) {
        unsigned int* w = &M.state[(size_t)rg_py(px[k]) * M.W + rg_px(px[k])];
        unsigned int old = lsdb_ld_state(w);
        while (true) {
            unsigned int neu;
            if ((old & (LSDB_ST_PACC | LSDB_ST_PREJ)) && ((tag - (old >> LSDB_ST_TAG_SHIFT)) & 4095u) < 2048u) neu = old | kind;   // an earlier seed marked it
            else neu = (old & ((1u << LSDB_ST_TAG_SHIFT) - 1u)) | kind | (tag << LSDB_ST_TAG_SHIFT);
            if (neu == old) break;
            const unsigned int seen = atomicCAS(w, old, neu);
            if (seen == old) break;
            old = seen;
        }
    }
    __syncwarp();
}
// a parked candidate was dropped (its seed died, or the evaluation was invalidated): take its marks back
__device__ void unpark_pixels(const RgMap& M, int lane, const unsigned int* px, int n, int chunk) {
    const unsigned int tag = (unsigned int)chunk & 4095u;
    for (int k = lane; k < n; k += 32) {
        unsigned int* w = &M.state[(size_t)rg_py(px[k]) * M.W + rg_px(px[k])];
        const unsigned int old = lsdb_ld_state(w);
        if ((old & (LSDB_ST_PACC | LSDB_ST_PREJ)) && (old >> LSDB_ST_TAG_SHIFT) == tag) atomicAnd(w, ~(LSDB_ST_PACC | LSDB_ST_PREJ));
    }
    __syncwarp();
}

// `need` words in the record arena of super-chunk slot `slot` (warp-uniform call); -1 when the arena is full
__device__ __forceinline__ int slot_alloc(RgShared& sh, const WarpCtx& c, int slot, int need) {
    int off = 0;
    if (c.lane == 0) off = (int)atomicAdd(&sh.slotHead[slot], (unsigned int)((need + 1) & ~1));
    off = __shfl_sync(FULL, off, 0);
    return off + need <= c.arenaCap ? off : -1;
}

__device__ bool any_banned(const RgMap& M, int lane, const unsigned int* px, int n) {
    bool hit = false;
    for (int k = lane; k < n; k += 32) if (ban_at(M, rg_px(px[k]), rg_py(px[k]))) hit = true;
    return __any_sync(FULL, hit);
}
__device__ bool any_unbanned(const RgMap& M, int lane, const unsigned int* px, int n) {
    bool hit = false;
    for (int k = lane; k < n; k += 32) if (!ban_at(M, rg_px(px[k]), rg_py(px[k]))) hit = true;
    return __any_sync(FULL, hit);
}
// does a pixel of the list carry the parked-accept mark of a seed at or before chunk `chunk`?
__device__ bool any_parked(const RgMap& M, int lane, const unsigned int* px, int n, int chunk) {
    bool hit = false;
    for (int k = lane; k < n; k += 32)
        if (rg_pend_applies(lsdb_ld_state(&M.state[(size_t)rg_py(px[k]) * M.W + rg_px(px[k])]), LSDB_ST_PACC, chunk)) hit = true;
    return __any_sync(FULL, hit);
}

// Has a region been accepted, since `L0` regions had been accepted, anywhere near the box [b0,b1]?  Conservative
// (never misses an overlap): decided per cell of the coarse accept grid.
__device__ __forceinline__ bool grid_hit(const RgShared& sh, unsigned int b0, unsigned int b1, int L0) {
    const int s = sh.cellShift;
    const int cx0 = rg_px(b0) >> s, cy0 = rg_py(b0) >> s, cx1 = rg_px(b1) >> s, cy1 = rg_py(b1) >> s;
    for (int cy = cy0; cy <= cy1; cy++)
        for (int cx = cx0; cx <= cx1; cx++)
            if (*(volatile const unsigned int*)&sh.grid[cy * GRID + cx] > (unsigned int)L0) return true;
    return false;
}

// ------------------------------------------------------------------ commit at the frontier (:242-271)
// hd = 13 doubles (rectangle + logNFA) readable by lane `hdLane` (its registers' copy in memory, or the arena record)
__device__ void commit_region(RgShared& sh, const WarpCtx& c, const double* hd, const unsigned int* px, int nCommit, int outcome) {
    const RgMap& M = sh.M;
    const int W = M.W;
    if (outcome == RG_OC_REJECT) {
        for (int k = c.lane; k < nCommit; k += 32) atomicOr(&M.state[(size_t)rg_py(px[k]) * W + rg_px(px[k])], LSDB_ST_REJ);
        if (c.lane == 0) STAT(sh, ST_REJECTS, 1);
        __syncwarp();
        return;
    }
    const int idx = __shfl_sync(FULL, (int)sh.nSeg, 0);
    for (int k = c.lane; k < nCommit; k += 32) {
        const unsigned int v = px[k];
        const size_t p = (size_t)rg_py(v) * W + rg_px(v);
        atomicOr(&M.state[p], LSDB_ST_BAN);
        atomicOr(M.bm + (size_t)rg_py(v) * M.pw + (rg_px(v) >> 5), 1u << (rg_px(v) & 31));
        c.labels[p] += idx + 1;   // regIdx += curMap*(regCnt+1), :261 (int32 here, u8 there)
        atomicMax(&sh.grid[(rg_py(v) >> sh.cellShift) * GRID + (rg_px(v) >> sh.cellShift)], (unsigned int)idx + 1u);
    }
    __syncwarp();
    __threadfence_block();
    if (c.lane == 0) {
        if (idx < c.maxSeg) {
            const double sca = M.kc->sca;
            LsdbRect& R = c.rects[idx];
            double rx1 = hd[0], ry1 = hd[1], rx2 = hd[2], ry2 = hd[3], rw = hd[4];
            if (sca != 1) {   // :252-258
                rx1 = (rx1 - 1.0) / sca + 1; ry1 = (ry1 - 1.0) / sca + 1;
                rx2 = (rx2 - 1.0) / sca + 1; ry2 = (ry2 - 1.0) / sca + 1;
                rw = (rw - 1.0) / sca + 1;
            }
            R.v[0] = rx1; R.v[1] = ry1; R.v[2] = rx2; R.v[3] = ry2; R.v[4] = rw;
            for (int k = 5; k < 13; k++) R.v[k] = hd[k];
        } else {
            sh.abortFlag = LSDB_ERR_CAPACITY;
        }
        sh.nSeg = idx + 1;
        STAT(sh, ST_ACCEPTS, 1);
    }
    __syncwarp();
}

__device__ __forceinline__ void eval_header(const RgEval& ev, double* hd) {
    hd[0] = ev.rec.x1; hd[1] = ev.rec.y1; hd[2] = ev.rec.x2; hd[3] = ev.rec.y2; hd[4] = ev.rec.wid; hd[5] = ev.rec.cX; hd[6] = ev.rec.cY;
    hd[7] = ev.rec.deg; hd[8] = ev.rec.dx; hd[9] = ev.rec.dy; hd[10] = ev.rec.p; hd[11] = ev.rec.prec; hd[12] = ev.logNFA;
}
__device__ __forceinline__ void eval_stats(RgShared& sh, const RgEval& ev) {
    STAT(sh, ST_GROWS, ev.nGrows); STAT(sh, ST_GROWNPX, ev.nGrownPx);
    if (ev.nRegrow) STAT(sh, ST_REGROWS, ev.nRegrow);
#ifdef RG_PROF_STEPS   /* development: grower steps / lanes side by side, in the rrr_passes / nfa_px slots */
    STAT(sh, ST_RRR, ev.nSteps); STAT(sh, ST_NFAPX, ev.nStepLanes);
    if (ev.nNfa) STAT(sh, ST_NFACALLS, ev.nNfa);
#else
    if (ev.nRrr) STAT(sh, ST_RRR, ev.nRrr);
    if (ev.nNfa) { STAT(sh, ST_NFACALLS, ev.nNfa); STAT(sh, ST_NFAPX, ev.nNfaPx); }
#endif
    STAT(sh, TM_LANEGROW, ev.cycGrow); STAT(sh, TM_RECT, ev.cycRect); STAT(sh, TM_NFA, ev.cycNfa);
    if (ev.oc == RG_OC_NOCHANGE && ev.nG1 < sh.M.regThre) STAT(sh, ST_SMALL, 1);
}

template <typename T>
__device__ __forceinline__ T* shfl_ptr(T* p, int src) {
    return reinterpret_cast<T*>(__shfl_sync(FULL, (unsigned long long)p, src));
}

// ------------------------------------------------------------------ worker: park the evaluation lane `l` holds
// Returns 0 = parked or dropped (the seed is then decided at the frontier), 1 = evaluate it again: an earlier seed
// parked an accept over pixels this evaluation took (it would fail V1 when it retires).
__device__ int park_large(RgShared& sh, const WarpCtx& c, int l, int p0, int ci, const RgEval& ev, int L1c, bool force) {
    const RgMap& M = sh.M;
    const int lane = c.lane;
    const int chunkJ = ci >> 5;
    const int oc = __shfl_sync(FULL, ev.oc, l), nG1 = __shfl_sync(FULL, ev.nG1, l), nG2 = __shfl_sync(FULL, ev.nG2, l);
    const int nCommit = __shfl_sync(FULL, ev.nCommit, l), usedT = __shfl_sync(FULL, ev.usedT, l), npnd = __shfl_sync(FULL, ev.npnd, l);
    const unsigned int* L0 = shfl_ptr(c.L.L0, l); const unsigned int* L1 = shfl_ptr(c.L.L1, l); const unsigned int* L2 = shfl_ptr(c.L.L2, l);
    const unsigned int* PN = shfl_ptr(c.L.pnd, l);
    int taken = 0;
    if (lane == 0) taken = seed_taken(lsdb_ld_state(&M.state[p0]), chunkJ);
    if (__shfl_sync(FULL, taken, 0)) return 0;   // swallowed by a region parked a moment ago: stays OC_NONE
    if (oc == RG_OC_DEFER || npnd < 0) { if (lane == 0) STAT(sh, RS_DROPPED, 1); return 0; }
    if (!force && (any_parked(M, lane, L0, nG1, chunkJ) || (usedT && any_parked(M, lane, L2, nG2, chunkJ)))) {
        if (lane == 0) STAT(sh, RS_REQUEUE, 1);
        return 1;
    }
    const bool heavy = oc == RG_OC_ACCEPT || oc == RG_OC_REJECT;
    // body: [G1 | G2 |] commit list, then the pending dependencies
    const int nA = usedT ? nG1 : (heavy ? 0 : nG1);     // G1 kept apart from the commit list
    const int nB = usedT ? nG2 : 0;
    const int nC = heavy ? nCommit : 0;
    const int chk = nA + nB + nC;
    const int used = (ARENA_HDR + chk + npnd + 1) & ~1;
    const int slot = slot_of_chunk(sh, chunkJ);
    const int off = slot_alloc(sh, c, slot, used);
    if (off < 0) { if (lane == 0) STAT(sh, RS_DROPPED, 1); return 0; }   // arena full: decided at the frontier
    unsigned int* dst = c.arenas + (size_t)slot * c.arenaCap + off;
    unsigned int* body = dst + ARENA_HDR;
    for (int k = lane; k < nA; k += 32) body[k] = L0[k];
    for (int k = lane; k < nB; k += 32) body[nA + k] = L2[k];
    const unsigned int* CL = usedT ? L1 : L0;
    for (int k = lane; k < nC; k += 32) body[nA + nB + k] = CL[k];
    for (int k = lane; k < npnd; k += 32) body[chk + k] = PN[k];
    if (lane == l) {
        eval_header(ev, reinterpret_cast<double*>(dst));
        dst[26] = (unsigned int)nC; dst[27] = (unsigned int)oc; dst[28] = (unsigned int)(ARENA_HDR + nA + nB);
        const size_t ri = (size_t)(chunkJ & (RING - 1)) * 32 + (ci & 31);
        c.R.L0[ri] = L1c; c.R.b0[ri] = rg_pack(ev.x0, ev.y0); c.R.b1[ri] = rg_pack(ev.x1, ev.y1);
        c.R.off[ri] = (unsigned int)off; c.R.chk[ri] = chk; c.R.chkOff[ri] = (unsigned int)off + ARENA_HDR;
        c.R.pnd[ri] = npnd; c.R.pndOff[ri] = (unsigned int)(off + ARENA_HDR + chk);
        c.R.oc[ri] = oc;
    }
    __syncwarp();
    // later seeds speculate as if these pixels already were usedMap 1 (accept) or 2 (reject); every evaluation that relied
    // on the marks is re-checked when it retires
    if (heavy) park_pixels(M, lane, body + nA + nB, nC, oc == RG_OC_ACCEPT ? LSDB_ST_PACC : LSDB_ST_PREJ, chunkJ);
    return 0;
}

// ------------------------------------------------------------------ worker: scout one super-chunk
// A  collect the live cells (seed order) into a queue; then, 32 queued seeds at a time:
// B  scout, one seed per lane; "no change" results are parked (accepted pixels kept for re-validation), the seeds
//    whose region reaches T pixels are listed in `lq` (pixel, cell), in seed order.  Returns their number.
__device__ int scout_super(RgShared& sh, WarpCtx& c, int chunk0, int nSub, unsigned int* lq) {
    const RgMap& M = sh.M;
    const int lane = c.lane;
    const unsigned int lt = (1u << lane) - 1u;
    const int nCells = sh.nCells;
    const int T = M.T;
    const int slot = slot_of_chunk(sh, chunk0);
    unsigned int* q = c.q;   // [2k] pixel index, [2k+1] (sub << 5 | lane)
    int qn = 0, nl = 0;
    for (int s = 0; s < nSub; s++) {   // ---- A
        const int ci = (chunk0 + s) * LSDB_CHUNK + lane;
        const int p = ci < nCells ? (int)c.cells[ci] : -1;
        const bool live = p >= 0 && !seed_taken(lsdb_ld_state(&M.state[p]), chunk0 + s);
        const unsigned int bal = __ballot_sync(FULL, live);
        if (live) {
            const int k = qn + __popc(bal & lt);
            q[2 * k] = (unsigned int)p;
            q[2 * k + 1] = (unsigned int)((s << 5) | lane);
        }
        qn += __popc(bal);
        c.R.oc[(size_t)((chunk0 + s) & (RING - 1)) * 32 + lane] = RG_OC_NONE;
    }
    __syncwarp();
    for (int base = 0; base < qn; base += 32) {   // ---- B
        const int k = base + lane;
        bool act = k < qn;
        const unsigned int rel = act ? q[2 * k + 1] : 0u;
        const int myChunk = chunk0 + (int)(rel >> 5);
        const int myp = act ? (int)q[2 * k] : 0;
        unsigned int lst[SG_CAP + 2];
        unsigned short rej[SG_CAP + 2];
        unsigned int pnd[SG_PND];
        int num = 0, npnd = 0;
        bool large = act;
        const int L0c = sh.nSeg;
        __threadfence_block();
        RgEval sev;
        sev.x0 = sev.y0 = 0x7fffffff; sev.x1 = sev.y1 = -1; sev.nGrows = 0; sev.nGrownPx = 0;
        if (act && T <= SG_CAP) {
            RgLane S;
            S.L0 = lst; S.L1 = 0; S.L2 = 0; S.rej = rej; S.pnd = pnd; S.vis = 0; S.cap = SG_CAP + 2; S.pndCap = SG_PND;
            double rd;
            num = rg_lane_grow<true>(M, S, lst, myp % M.W, myp / M.W, M.deg[myp], M.kc->degThre, myChunk, T, npnd, rd, sev);
            large = num >= T;
        }
        __syncwarp();
        const bool small = act && !large && npnd >= 0;
        // park the accepted pixels (re-validated at retire time if a region was accepted nearby) and the pending dependencies
        const int need = small ? num + npnd : 0;
        int incl = need;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        const int tot = __shfl_sync(FULL, incl, 31);
        const int off = tot > 0 ? slot_alloc(sh, c, slot, tot) : -1;
        if (small && off >= 0) {
            const size_t ri = (size_t)(myChunk & (RING - 1)) * 32 + (rel & 31u);
            const int mine = off + (incl - need);
            unsigned int* dst = c.arenas + (size_t)slot * c.arenaCap + mine;
            for (int j = 0; j < num; j++) dst[j] = lst[j];
            for (int j = 0; j < npnd; j++) dst[num + j] = pnd[j];
            c.R.L0[ri] = L0c; c.R.b0[ri] = rg_pack(sev.x0, sev.y0); c.R.b1[ri] = rg_pack(sev.x1, sev.y1);
            c.R.off[ri] = 0; c.R.chk[ri] = num; c.R.chkOff[ri] = (unsigned int)mine;
            c.R.pnd[ri] = npnd; c.R.pndOff[ri] = (unsigned int)(mine + num);
            c.R.oc[ri] = RG_OC_NOCHANGE;
        }   // else: arena full / too many dependencies — this seed is decided at the frontier
        const unsigned int largeMask = __ballot_sync(FULL, act && large);
        if (act && large) {
            const int e = nl + __popc(largeMask & lt);
            lq[2 * e] = (unsigned int)myp;
            lq[2 * e + 1] = (unsigned int)(myChunk * LSDB_CHUNK + (int)(rel & 31u));
        }
        nl += __popc(largeMask);
        {
            const unsigned int nSmall = __popc(__ballot_sync(FULL, small));
            int pxs = small ? num : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pxs += __shfl_xor_sync(FULL, pxs, o);
            if (lane == 0) { STAT(sh, ST_SPEC, nSmall); STAT(sh, ST_GROWS, nSmall); STAT(sh, ST_SMALL, nSmall); STAT(sh, ST_GROWNPX, pxs); }
        }
    }
    __syncwarp();
    return nl;
}

// ------------------------------------------------------------------ worker: one pass
// Claims up to RG_P super-chunks, scouts them, then evaluates their large seeds in rounds: lane j takes the next
// large seed of super-chunk j that is still live (grow, rectangle, Refiner, NFA — rg_eval_lane), and the results are
// parked in lane order — their pixels marked — before the next round.  Consecutive seeds of the sorted list mostly sit
// on the same wall (the first one swallows the rest), so the seeds of DIFFERENT super-chunks are what can be
// evaluated side by side.  A super-chunk is flagged READY as soon as its last large seed is settled.
__device__ bool worker_pass(RgShared& sh, WarpCtx& c) {
    const RgMap& M = sh.M;
    const int lane = c.lane;
    const int nChunks = sh.nChunks;
    const long long tSpec = clock64();
    int myChunk0 = -1, myNSub = 0, nClaimed = 0;   // lane j: super-chunk j of this pass
    for (int j = 0; j < RG_P; j++) {
        int chunk = -1;
        if (lane == 0) {
            // claim a ticket only while the ring has room
            if (sh.nextChunk < nChunks && sh.nextChunk - sh.frontier < sh.runAhead) {
                chunk = atomicAdd(&sh.nextChunk, 1 << sh.supShift);
                if (chunk >= nChunks) chunk = -1;
                else sh.slotHead[slot_of_chunk(sh, chunk)] = 0;   // the slot's previous super-chunk has retired
            }
        }
        chunk = __shfl_sync(FULL, chunk, 0);
        if (chunk < 0) break;
        if (lane == j) { myChunk0 = chunk; myNSub = min(1 << sh.supShift, nChunks - chunk); }
        nClaimed++;
    }
    if (nClaimed == 0) return false;
    int nl = 0;   // lane j: large seeds of super-chunk j
    for (int j = 0; j < nClaimed; j++) {
        const int n = scout_super(sh, c, __shfl_sync(FULL, myChunk0, j), __shfl_sync(FULL, myNSub, j), c.lq + (size_t)j * QWORDS);
        if (lane == j) nl = n;
    }
    const unsigned int* myq = c.lq + (size_t)(lane < RG_P ? lane : 0) * QWORDS;
    int pos = 0, retries = 0;
    bool flagged = false;
    while (true) {
        if (__shfl_sync(FULL, (int)sh.abortFlag, 0)) break;
        // the next large seed of my super-chunk that nobody has swallowed meanwhile
        int p0 = 0, ci = 0;
        bool mine = false;
        if (lane < nClaimed) {
            while (pos < nl) {
                p0 = (int)myq[2 * pos]; ci = (int)myq[2 * pos + 1];
                if (!seed_taken(lsdb_ld_state(&M.state[p0]), ci >> 5)) break;
                pos++; retries = 0;
            }
            mine = pos < nl;
        }
        // super-chunks without large seeds left are READY
        __threadfence();
        if (lane < nClaimed && !mine && !flagged) {
            for (int t = 0; t < myNSub; t++) sh.chunkFlag[(myChunk0 + t) & (RING - 1)] = 1;
            flagged = true;
        }
        const unsigned int mineMask = __ballot_sync(FULL, mine);
        if (!mineMask) break;
        const long long tG = clock64();
        const int L1c = sh.nSeg;
        __threadfence_block();
        RgEval ev;
        ev.oc = RG_OC_NONE; ev.nG1 = ev.nG2 = ev.nCommit = ev.usedT = ev.npnd = 0;
        if (mine) { rg_eval_lane(M, c.L, p0, ci >> 5, ev); eval_stats(sh, ev); STAT(sh, ST_SPEC, 1); }
        __syncwarp();
        __threadfence_block();
        for (int l = 0; l < nClaimed; l++) {
            if (!((mineMask >> l) & 1u)) continue;
            const int pl = __shfl_sync(FULL, p0, l), cl = __shfl_sync(FULL, ci, l);
            const bool force = __shfl_sync(FULL, retries, l) >= 2;
            const int again = park_large(sh, c, l, pl, cl, ev, L1c, force);
            if (lane == l) { if (again) retries++; else { pos++; retries = 0; } }
        }
        if (lane == 0) { STAT(sh, TM_GROW, clock64() - tG); STAT(sh, ST_ROUNDS, 1); }
    }
    if (lane == 0) STAT(sh, TM_SPEC, clock64() - tSpec);
    __syncwarp();
    return true;
}

// ------------------------------------------------------------------ frontier: retire one READY chunk
// Lane-level validation of a parked "no change" result whose pixel lists are short (the scouts' results): every pixel it
// accepted must still be un-banned (looked at only when a region was accepted nearby since), and every pixel it skipped
// because of a parked accept must be banned by now.
__device__ __forceinline__ bool lane_valid(const RgMap& M, const unsigned int* earena, bool hit, int nchk, unsigned int chkOff,
                                           int npnd, unsigned int pndOff) {
    unsigned int bad = 0;   // no short-circuit: the loads of a list are independent and must overlap
    if (hit) {
        const unsigned int* px = earena + chkOff;
        for (int j = 0; j < nchk; j++) bad |= ban_at(M, rg_px(px[j]), rg_py(px[j])) ? 1u : 0u;
    }
    const unsigned int* pp = earena + pndOff;
    for (int j = 0; j < npnd; j++) bad |= ban_at(M, rg_px(pp[j]), rg_py(pp[j])) ? 0u : 1u;
    return bad == 0;
}

__device__ void retire_chunk(RgShared& sh, WarpCtx& c, int chunk) {
    const RgMap& M = sh.M;
    const int lane = c.lane;
#ifdef RG_PROF_RETIRE
    const long long tR0 = clock64();
#endif
    const int slot = chunk & (RING - 1);
    const int ci = chunk * LSDB_CHUNK + lane;
    const int myp = ci < sh.nCells ? (int)c.cells[ci] : -1;
    const size_t ri = (size_t)slot * 32 + lane;
    const int recOc = c.R.oc[ri], recL0 = c.R.L0[ri], recChk = c.R.chk[ri];
    const unsigned int recB0 = c.R.b0[ri], recB1 = c.R.b1[ri], recOff = c.R.off[ri], recChkOff = c.R.chkOff[ri];
    const int recPnd = recOc != RG_OC_NONE ? c.R.pnd[ri] : 0;
    const unsigned int recPndOff = c.R.pndOff[ri];
    bool committedParked = false;   // this lane's parked accept / reject was committed as parked
    const unsigned int* earena = c.arenas + (size_t)slot_of_chunk(sh, chunk) * c.arenaCap;
    bool live = myp >= 0 && (lsdb_ld_state(&M.state[myp]) & 3u) == 0;   // :222
    // short "no change" records are validated by their own lane, in parallel; the rest is walked in seed order
    const bool laneCheck = recOc == RG_OC_NOCHANGE && recChk >= 0 && recChk + recPnd <= 64;
    bool laneOK = false;
    if (live && laneCheck) laneOK = lane_valid(M, earena, grid_hit(sh, recB0, recB1, recL0), recChk, recChkOff, recPnd, recPndOff);
    bool doneL = false;             // this lane's seed has had its turn
    unsigned int liveAtTurn = __ballot_sync(FULL, live);
    unsigned int work = __ballot_sync(FULL, live && !(laneCheck && laneOK));
#ifdef RG_PROF_RETIRE
    if (lane == 0) STAT(sh, ST_RRR, clock64() - tR0);
    const long long tR1 = clock64();
#endif
    while (work) {
        if (__shfl_sync(FULL, (int)sh.abortFlag, 0)) break;
        const int k = __ffs(work) - 1;
        const int oc = __shfl_sync(FULL, recOc, k);
        const bool lc = __shfl_sync(FULL, (int)laneCheck, k) != 0;
        bool valid = oc != RG_OC_NONE && !lc;   // a lane-checked record that got here has failed
        if (valid) {
            const int nchk = __shfl_sync(FULL, recChk, k);
            const bool hit = grid_hit(sh, __shfl_sync(FULL, recB0, k), __shfl_sync(FULL, recB1, k), __shfl_sync(FULL, recL0, k));
            if (hit) valid = nchk >= 0 && !any_banned(M, lane, earena + __shfl_sync(FULL, recChkOff, k), nchk);
        }
        if (valid) {   // every pixel the evaluation took for banned because of a parked accept must be banned by now
            const int npn = __shfl_sync(FULL, recPnd, k);
            if (npn > 0) valid = !any_unbanned(M, lane, earena + __shfl_sync(FULL, recPndOff, k), npn);
        }
        int changedAt = -1;   // lane index of the seed whose commit changed usedMap in this round
        if (valid) {
            work &= ~(1u << k);
            if (lane == k) doneL = true;
            if (oc != RG_OC_NOCHANGE) {
                const unsigned int* recp = earena + __shfl_sync(FULL, recOff, k);
                commit_region(sh, c, reinterpret_cast<const double*>(recp), recp + recp[28], (int)recp[26], (int)recp[27]);
                changedAt = k;
                if (lane == k) committedParked = true;
            }
        } else {
            // evaluate at the frontier, where the state is final: seed k and the seeds after it that are known to need
            // it (no record, or a lane-checked record that failed), up to RG_P, one per lane; stop at the first seed
            // whose parked large record has not been looked at yet
            const long long t0 = clock64();
            const bool needEval = ((work >> lane) & 1u) && (recOc == RG_OC_NONE || laneCheck);
            const unsigned int needMask = __ballot_sync(FULL, needEval) | (1u << k);
            const unsigned int stopMask = work & ~needMask;                    // parked large records still to validate
            const unsigned int upto = stopMask ? ((1u << (__ffs(stopMask) - 1)) - 1u) : FULL;
            unsigned int batch = needMask & upto & ~((1u << k) - 1u);
            const int cnt = min(__popc(batch), RG_P);
            const int src = lane < cnt ? __fns(batch, 0, lane + 1) : 0;
            const int p0 = __shfl_sync(FULL, myp, src);
            const int ocSrc = __shfl_sync(FULL, recOc, src);
            RgEval ev;
            ev.oc = RG_OC_NONE; ev.nG1 = ev.nG2 = ev.nCommit = ev.usedT = ev.npnd = 0;
            if (lane < cnt) { rg_eval_lane(M, c.L, p0, -1, ev); eval_stats(sh, ev); }
            __syncwarp();
            bool bigUsed = false;
            for (int l = 0; l < cnt; l++) {
                const int kk = __shfl_sync(FULL, src, l);       // chunk lane of the l-th seed of the batch
                int oc2 = __shfl_sync(FULL, ev.oc, l);
                const int ocOld = __shfl_sync(FULL, ocSrc, l);
                if (oc2 == RG_OC_DEFER) {
                    // the region outgrew the lane's lists: once more with the CTA's full-size set (lane 0's plane is clean)
                    const int pb = __shfl_sync(FULL, p0, l);
                    RgEval evb;
                    evb.oc = RG_OC_NONE; evb.nG1 = evb.nG2 = evb.nCommit = evb.usedT = evb.npnd = 0;
                    if (lane == 0) { rg_eval_lane(M, c.big, pb, -1, evb); eval_stats(sh, evb); STAT(sh, RS_BIG, 1); }
                    __syncwarp();
                    oc2 = __shfl_sync(FULL, evb.oc, 0);
                    if (oc2 == RG_OC_DEFER) { if (lane == 0) sh.abortFlag = LSDB_ERR_CAPACITY; break; }
                    if (oc2 == RG_OC_REJECT || oc2 == RG_OC_ACCEPT) {
                        __shared__ double hdBig[16];
                        if (lane == 0) eval_header(evb, hdBig);
                        __syncwarp();
                        const int usedT = __shfl_sync(FULL, evb.usedT, 0);
                        commit_region(sh, c, hdBig, usedT ? c.big.L1 : c.big.L0, __shfl_sync(FULL, evb.nCommit, 0), oc2);
                        changedAt = kk;
                    }
                    bigUsed = true;
                } else if (oc2 == RG_OC_REJECT || oc2 == RG_OC_ACCEPT) {
                    __shared__ double hdL[16];
                    if (lane == l) eval_header(ev, hdL);
                    __syncwarp();
                    const int usedT = __shfl_sync(FULL, ev.usedT, l);
                    const unsigned int* px = usedT ? shfl_ptr(c.L.L1, l) : shfl_ptr(c.L.L0, l);
                    commit_region(sh, c, hdL, px, __shfl_sync(FULL, ev.nCommit, l), oc2);
                    changedAt = kk;
                }
                work &= ~(1u << kk);
                if (lane == kk) doneL = true;
                if (lane == 0) {
                    STAT(sh, ST_RESPEC, 1);
                    STAT(sh, ocOld == RG_OC_NONE ? RS_NONE : RS_CONFLICT, 1);
                    if (ocOld == RG_OC_ACCEPT || ocOld == RG_OC_REJECT) STAT(sh, RS_LOST, 1);
                    if (changedAt == kk) STAT(sh, RS_COMMIT, 1);
                }
                if (changedAt >= 0) break;   // usedMap changed: the evaluations after this one are void
            }
            (void)bigUsed;
            if (lane == 0) STAT(sh, TM_RESPEC, clock64() - t0);
        }
        if (changedAt >= 0) {
            // usedMap changed: refresh the cells that come after the committing seed in this chunk
            __threadfence_block();
            if (lane > changedAt && live && !doneL) {
                live = (lsdb_ld_state(&M.state[myp]) & 3u) == 0;
                if (live && laneCheck) laneOK = lane_valid(M, earena, grid_hit(sh, recB0, recB1, recL0), recChk, recChkOff, recPnd, recPndOff);
            }
            const unsigned int later = ~((2u << changedAt) - 1u);
            liveAtTurn = (liveAtTurn & ~later) | (__ballot_sync(FULL, live) & later);
            work = __ballot_sync(FULL, lane > changedAt && live && !doneL && !(laneCheck && laneOK)) | (work & ~later & ~(1u << changedAt));
        }
    }
    // parked accepts / rejects that were not committed as parked (seed dead at its turn, or evaluation invalidated):
    // take their marks back so that later speculation stops counting on them
#ifdef RG_PROF_RETIRE
    if (lane == 0) STAT(sh, ST_NFAPX, clock64() - tR1);
#endif
    unsigned int drop = __ballot_sync(FULL, (recOc == RG_OC_ACCEPT || recOc == RG_OC_REJECT) && !committedParked);
    while (drop) {
        const int k = __ffs(drop) - 1;
        drop &= drop - 1;
        const unsigned int* recp = earena + __shfl_sync(FULL, recOff, k);
        unpark_pixels(M, lane, recp + recp[28], (int)recp[26], chunk);
    }
    if (lane == 0) STAT(sh, ST_LIVE, __popc(liveAtTurn));
    __syncwarp();
}

// drain the READY prefix at the frontier; returns the number of chunks retired
__device__ int retire_ready(RgShared& sh, WarpCtx& c, int maxChunks) {
    int n = 0;
    const long long t0 = clock64();
    while (n < maxChunks) {
        if (__shfl_sync(FULL, (int)sh.abortFlag, 0)) break;
        int f = 0, ready = 0;
        if (c.lane == 0) { f = sh.frontier; ready = f < sh.nChunks && sh.chunkFlag[f & (RING - 1)] == 1; }
        f = __shfl_sync(FULL, f, 0);
        if (!__shfl_sync(FULL, ready, 0)) break;
        __threadfence();
        retire_chunk(sh, c, f);
        __threadfence_block();
        if (c.lane == 0) {
            sh.chunkFlag[f & (RING - 1)] = 0;
            STAT(sh, ST_CHUNKS, 1);
            __threadfence_block();
            sh.frontier = f + 1;
        }
        __syncwarp();
        n++;
    }
    if (n && c.lane == 0) STAT(sh, TM_RETIRE, clock64() - t0);
    return n;
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) lsdb_region_kernel(const LsdbRegionParams P) {
    __shared__ RgShared sh;
    extern __shared__ unsigned int bmShared[];   // the map's ban plane, one bit per pixel, when it fits (bmCapWords words)
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
    WarpCtx c;
    c.lane = lane; c.w = w;
    c.arenaCap = P.arenaCap; c.maxSeg = P.maxSeg;
    unsigned int* ctaBase = P.work + (size_t)blockIdx.x * P.workWordsPerCta;
    c.arenas = ctaBase;
    unsigned int* warpBase = ctaBase + (size_t)NSLOTS * P.arenaCap + (size_t)w * rg_warp_words();
    c.q = warpBase;
    c.lq = warpBase + QWORDS;
    {
        unsigned int* visCta = P.vis + (size_t)blockIdx.x * nw * RG_P * P.planeWords;
        const int l = lane < RG_P ? lane : 0;
        c.L = rg_lane_set(warpBase + (1 + RG_P) * (size_t)QWORDS + (size_t)l * ((rg_lane_words(RG_LANE_CAP) + 1) & ~(size_t)1), RG_LANE_CAP,
                          visCta + ((size_t)w * RG_P + l) * P.planeWords);
        c.big = rg_lane_set(ctaBase + (size_t)NSLOTS * P.arenaCap + (size_t)nw * rg_warp_words(), P.listCap, visCta);
    }
    {
        unsigned char* base = P.recBuf + (size_t)blockIdx.x * (RING * 32 * REC_BYTES_PER_CELL);
        c.R.oc = reinterpret_cast<int*>(base);
        c.R.L0 = c.R.oc + RING * 32; c.R.b0 = reinterpret_cast<unsigned int*>(c.R.L0 + RING * 32); c.R.b1 = c.R.b0 + RING * 32;
        c.R.off = c.R.b1 + RING * 32; c.R.chk = reinterpret_cast<int*>(c.R.off + RING * 32); c.R.chkOff = reinterpret_cast<unsigned int*>(c.R.chk + RING * 32);
        c.R.pndOff = c.R.chkOff + RING * 32; c.R.pnd = reinterpret_cast<int*>(c.R.pndOff + RING * 32);
    }

    while (true) {
        __syncthreads();
        if (tid == 0) {
            const int img = atomicAdd(P.imgCounter, 1);
            sh.img = img;
            if (img < P.nImgs) {
                sh.frontier = 0; sh.nextChunk = 0; sh.nSeg = 0; sh.abortFlag = 0;
                sh.nCells = P.dyn[img].nCells;
                sh.nChunks = (sh.nCells + LSDB_CHUNK - 1) / LSDB_CHUNK;
                // chunks per claim: LSDB_SUPER for long seed lists; a short list (a small map alone on the device) is cut
                // finer so that every worker gets several claims
                int shift = 3;
                static_assert(LSDB_SUPER == 8, "supShift starts at log2(LSDB_SUPER)");
                if ((P.flags >> 8) & 15) shift = ((P.flags >> 8) & 15) - 1;
                else while (shift > 0 && (sh.nChunks >> shift) < 6 * nw) shift--;
                sh.supShift = shift;
                int ra = P.runAhead;
                if (ra <= 0 || ra > RING - (NW_MAX + 1) * LSDB_SUPER) ra = RING - (NW_MAX + 1) * LSDB_SUPER;
                sh.runAhead = min(ra, (NSLOTS - nw - 1) << shift);   // a slot is not reused while its claim is in flight
            }
        }
        if (tid < TM_N) sh.stats[tid] = 0;
        if (tid < NSLOTS) sh.slotHead[tid] = 0;
        for (int i = tid; i < GRID * GRID; i += blockDim.x) sh.grid[i] = 0;
        for (int i = tid; i < RING; i += blockDim.x) sh.chunkFlag[i] = 0;
        __syncthreads();
        const int img = sh.img;
        if (img >= P.nImgs) break;
        long long mapC0 = 0; unsigned long long mapT0 = 0;
        if (tid == 0) { mapC0 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(mapT0)); }
        const LsdbImg im = P.imgs[img];
        if (tid == 0) {
            RgMap& M = sh.M;
            M.W = im.W; M.H = im.H; M.pw = im.pw; M.n = im.n;
            M.state = P.state + im.nOff; M.deg = P.deg + im.nOff; M.mag = P.mag + im.nOff; M.cs = P.cs + 2 * im.nOff;
            M.kc = P.kc; M.lgammaTab = P.lgammaTab; M.lgammaN = P.lgammaN;
            M.logNT = im.logNT; M.regThre = im.regThre; M.T = (int)ceil(im.regThre);
            int k = 0;
            while (((max(im.W, im.H) - 1) >> k) > GRID - 1) k++;
            sh.cellShift = k;
            // the ban plane written by the stencil stage moves into shared memory when it fits
            M.bm = (im.H * im.pw <= P.bmCapWords) ? bmShared : P.banBits + im.banOff;
        }
        {
            const int words = im.H * im.pw;
            if (words <= P.bmCapWords) {
                const unsigned int* srcB = P.banBits + im.banOff;
                for (int i = tid; i < words; i += blockDim.x) bmShared[i] = srcB[i];
            }
        }
        __syncthreads();
        c.labels = P.labels + im.nOff;
        c.cells = P.cells + im.nOff;
        c.rects = P.rects + im.segOff;
        const int nChunks = sh.nChunks;
        unsigned int idle = 0;

        while (!__shfl_sync(FULL, (int)sh.abortFlag, 0)) {
            if (__shfl_sync(FULL, (int)sh.frontier, 0) >= nChunks) break;
            bool didWork = false;
            if (w == 0) {   // the frontier warp
                didWork = retire_ready(sh, c, nw > 1 ? RING : 4) > 0;
            }
            if (w != 0 || nw == 1) {   // workers (a one-warp team does both jobs in turn)
                if (worker_pass(sh, c)) didWork = true;
            }
            if (didWork) { idle = 0; continue; }
            const long long tw = clock64();
            __nanosleep(w == 0 ? 100 : 400);
            if (lane == 0) {
                STAT(sh, TM_WAIT, clock64() - tw);
                if (++idle > (1u << 24)) sh.abortFlag = LSDB_ERR_TIMEOUT;
            }
            __syncwarp();
        }
        __syncthreads();
        if (tid == 0) {
            P.dyn[img].nSeg = sh.nSeg;
            if (sh.abortFlag) P.dyn[img].err = sh.abortFlag;
            sh.stats[ST_CELLS] = sh.nCells;
            unsigned long long mapT1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(mapT1));
            sh.stats[TM_MAPCYC] = (unsigned long long)(clock64() - mapC0); sh.stats[TM_MAPNS] = mapT1 - mapT0;
        }
        __syncthreads();
        if (tid < TM_N) P.dyn[img].stat[tid] = (long long)sh.stats[tid];
    }
}

__global__ void lsdb_lgamma_table_kernel(double* tab, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tab[i] = i >= 1 ? rg_log_gamma_calc(i) : 0.0;
}

__global__ void lsdb_used_plane_kernel(const unsigned int* __restrict__ state, uint8_t* __restrict__ used, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        unsigned int s = state[i];
        used[i] = (s & LSDB_ST_BAN) ? 1 : ((s & LSDB_ST_REJ) ? 2 : 0);
    }
}

static size_t region_dyn_smem(int bmCapWords) { return (size_t)((bmCapWords + 1) & ~1) * 4; }

void lsdb_launch_grow(cudaStream_t s, int nImgs, int nCtas, int warpsPerCta, const LsdbImg* imgs, LsdbImgDyn* dyn,
                      const LsdbLsdConst* kc, const double* mag, const double* deg, const double* cosm, const double* sinm,
                      unsigned int* state, const unsigned int* cells, int* labels, LsdbRect* rects, int maxSeg,
                      unsigned int* lists, int listCap, int arenaCap, int runAhead, unsigned char* recBuf, const double* lgammaTab, int lgammaN,
                      int* imgCounter, unsigned int* banBits, int bmCapWords, int steal, unsigned int* vis, size_t planeWords) {
    (void)sinm;
    cudaFuncSetAttribute(lsdb_region_kernel<NW_MAX * 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)region_dyn_smem(bmCapWords));
    if (nImgs <= 0) return;
    LsdbRegionParams P;
    P.nImgs = nImgs; P.imgs = imgs; P.dyn = dyn; P.kc = kc; P.mag = mag; P.deg = deg; P.cs = cosm; P.state = state; P.cells = cells;
    P.labels = labels; P.rects = rects; P.maxSeg = maxSeg; P.work = lists; P.workWordsPerCta = rg_cta_words(listCap, arenaCap, warpsPerCta);
    P.listCap = listCap; P.arenaCap = arenaCap; P.runAhead = runAhead; P.vis = vis; P.planeWords = planeWords; P.recBuf = recBuf;
    P.lgammaTab = lgammaTab; P.lgammaN = lgammaN; P.imgCounter = imgCounter; P.banBits = banBits; P.bmCapWords = bmCapWords; P.flags = steal;
    lsdb_region_kernel<NW_MAX * 32><<<nCtas, warpsPerCta * 32, region_dyn_smem(bmCapWords), s>>>(P);
}

size_t lsdb_grow_words_per_cta(int listCap, int arenaCap, int warpsPerCta) { return rg_cta_words(listCap, arenaCap, warpsPerCta); }
size_t lsdb_grow_rec_bytes_per_cta(void) { return (size_t)RING * 32 * REC_BYTES_PER_CELL; }
size_t lsdb_grow_vis_planes_per_cta(int warpsPerCta) { return (size_t)warpsPerCta * RG_P; }

void lsdb_launch_lgamma_table(cudaStream_t s, double* tab, int n) {
    lsdb_lgamma_table_kernel<<<(n + 127) / 128, 128, 0, s>>>(tab, n);
}

void lsdb_launch_used_plane(cudaStream_t s, const unsigned int* state, uint8_t* used, int n) {
    lsdb_used_plane_kernel<<<(n + 255) / 256, 256, 0, s>>>(state, used, n);
}

// how many CTAs of `warpsPerCta` warps fit on the device at once (the kernel is persistent)
int lsdb_grow_max_ctas(int device, int warpsPerCta, int bmCapWords) {
    int sms = 0, per = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaFuncSetAttribute(lsdb_region_kernel<NW_MAX * 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)region_dyn_smem(bmCapWords));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, lsdb_region_kernel<NW_MAX * 32>, warpsPerCta * 32, region_dyn_smem(bmCapWords));
    if (per < 1) per = 1;
    return sms * per;
}

// largest ban plane (in 32-bit words) that fits in shared memory next to the kernel's static data
int lsdb_grow_max_bitmap_words(int device) {
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, lsdb_region_kernel<NW_MAX * 32>) != cudaSuccess) return 0;
    const long long room = (long long)optin - (long long)fa.sharedSizeBytes - 256;
    return room > 0 ? (int)(room / 4) : 0;
}
