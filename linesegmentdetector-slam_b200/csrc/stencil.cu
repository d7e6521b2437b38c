// Stage A of the LSD hot path on sm_100a: value remap + separable decimating Gaussian + gradient /
// level-line angle / threshold mask, fused in one shared-memory-staged stencil kernel.
//
// Replaces (reference, /root/reference/LSD/myLSD.cpp): the remap loop :135-142, GaussianSampler
// :378-484 (X pass :420-448, Y pass :452-482) and the gradient loop :151-174 incl. maxGrad.
// Bit-exactness rules: taps are accumulated i = 0..16 in order with separate mul and add
// (-fmad=false), the source byte is converted exactly, atan2/cos/sin come from lsd_math.h.
//
// Occupancy grids are sparse: after the remap ~1 % of the source pixels are non-zero.  Adding a
// zero tap is exact (v + (+0*k) = v, and every non-zero tap is positive, so no -0 can arise), so
// the passes only touch non-zero taps:
//   * staging builds one bit per source pixel (non-zero after the remap); the bytes themselves stay in global memory;
//   * X pass: the 17-tap window of an output is a 17-bit slice of its row's bit vector; an empty
//     slice costs one funnel shift, a non-empty one accumulates exactly its set taps in ascending
//     tap order and flags the element in a per-column bit vector over the rows;
//   * Y pass: same, on the column bit vectors;  a tile whose whole window is zero skips both;
//   * the pixels with a non-zero gradient (~15 %) are compacted before the double-double
//     atan2 / cos / sin calls, so those run with full lanes instead of under divergence;
//   * results are assembled in shared-memory tiles and written with fully coalesced rows.
//
// One CTA per 32x32 tile of the scaled image; source window <= 136 x 160 bytes.
// HBM traffic per source pixel: 1 B read + 0.09*(8+8+4 [+16 for growable pixels]) B written.
#include <stdlib.h>
#include "lsdb_common.cuh"

#define SRC_PITCH LSDB_SRC_PITCH
#define ROW_WORDS (SRC_PITCH / 32)
#define GW 33
#define NT 256

struct StencilSmem {
    union {
        double aux[LSDB_SRC_MAX * GW];                       // X-pass output (only flagged elements are written/read)
        struct { double magT[1024], degT[1024], cosT[1024], sinT[1024]; } out;  // output tiles (after the Y pass)
    } u;
    double g[GW * GW];
    double taps[3 * 17];
    double wmax[8];
    unsigned int rowBits[LSDB_SRC_MAX * ROW_WORDS];          // bit x of row r: src[r][x] != 0
    unsigned int colBits[GW * ROW_WORDS];                    // bit r of column c: aux[r][c] != 0
    short idxX[GW * 17];
    short idxY[GW * 17];
    unsigned short queue[1024];                              // pixels that need atan2: general angles from the front,
                                                             // axis-aligned gradients (exact special angles) from the back
    unsigned char stT[1024];
    unsigned char neRows[LSDB_SRC_MAX];                      // source rows of the window that hold a non-zero pixel
    int qn, qt;
    int geo[6];
    int nNe;
    int anySrc;
};

__device__ __forceinline__ int lsdb_reflect(int j, int lim) {  // LSD/myLSD.cpp:435-443
    int dou = 2 * lim;
    while (j < 0) j += dou;
    while (j >= dou) j -= dou;
    if (j >= lim) j = dou - j - 1;
    return j;
}

// window [lo,hi] of source indices (after reflection) needed by centres c0..c1 with half-width h
__device__ __forceinline__ void lsdb_window(int c0, int c1, int h, int lim, int* lo, int* hi) {
    int lo_raw = c0 - h, hi_raw = c1 + h;
    if (lo_raw >= 0 && hi_raw < lim) { *lo = lo_raw; *hi = hi_raw; return; }
    if (lim <= LSDB_SRC_MAX - 16) { *lo = 0; *hi = lim - 1; return; }
    if (lo_raw < 0) {
        int m = -1 - lo_raw;
        *lo = 0; *hi = hi_raw > m ? hi_raw : m;
        if (*hi > lim - 1) *hi = lim - 1;
    } else {
        int m = 2 * lim - 1 - hi_raw;
        *lo = lo_raw < m ? lo_raw : m; *hi = lim - 1;
        if (*lo < 0) *lo = 0;
    }
}

// bit k = byte k of w is non-zero
__device__ __forceinline__ unsigned int nz4(unsigned int w) {
    const unsigned int t = __vcmpne4(w, 0u) & 0x01010101u;     // 1 in the low bit of every non-zero byte
    return ((t * 0x01020408u) >> 24) & 0xfu;                    // gather: byte 0 -> bit 0 ... byte 3 -> bit 3
}

// 17 bits starting at bit `s` of a ROW_WORDS-word bit vector
__device__ __forceinline__ unsigned int bits17(const unsigned int* v, int s) {
    const int wi = s >> 5;
    const unsigned int lo = v[wi];
    const unsigned int hi = wi + 1 < ROW_WORDS ? v[wi + 1] : 0u;
    return __funnelshift_r(lo, hi, s & 31) & 0x1ffffu;
}

__global__ void __launch_bounds__(NT, 4) lsdb_stencil_kernel(const LsdbImg* __restrict__ imgs, const int* __restrict__ tileImg,
                                                          LsdbImgDyn* __restrict__ dyn, const LsdbLsdConst* __restrict__ kc,
                                                          const uint8_t* __restrict__ src, double* __restrict__ mag,
                                                          double* __restrict__ deg, double* __restrict__ cosm,
                                                          double* __restrict__ sinm, unsigned int* __restrict__ state,
                                                          unsigned int* __restrict__ banBits, unsigned int* __restrict__ nzBits,
                                                          double* __restrict__ gaussOut, int tileBase) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StencilSmem& S = *reinterpret_cast<StencilSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tileIdx = blockIdx.x + tileBase;   // a launch may cover a range of tiles only (one map tiled over several GPUs)
    const int imgIdx = tileImg[tileIdx];
    const LsdbImg im = imgs[imgIdx];
    const int lt = tileIdx - im.tile0;
    const int tx = lt % im.tilesX, ty = lt / im.tilesX;
    const int x0 = tx * LSDB_TILE, y0 = ty * LSDB_TILE;
    const int x1 = min(x0 + LSDB_TILE, im.W), y1 = min(y0 + LSDB_TILE, im.H);
    const int gxs = max(x0 - 1, 0), gys = max(y0 - 1, 0);
    const int gw = x1 - gxs, gh = y1 - gys;
    const double sca = kc->sca;
    const int h = kc->h;  // 8

    // centres of the first / last Gaussian column and row: xc = floor(x/sca + 0.5)  (:428,:460) and the source window —
    // the same for every thread: one thread does the double-precision divisions, the rest read the result
    if (tid == 0) {
        const int xcA_ = lsdb_x86_d2i(floor(gxs / sca + 0.5)), xcB_ = lsdb_x86_d2i(floor((x1 - 1) / sca + 0.5));
        const int ycA_ = lsdb_x86_d2i(floor(gys / sca + 0.5)), ycB_ = lsdb_x86_d2i(floor((y1 - 1) / sca + 0.5));
        int a0, a1, b0, b1;
        lsdb_window(xcA_, xcB_, h, im.cols, &a0, &a1);
        lsdb_window(ycA_, ycB_, h, im.rows, &b0, &b1);
        S.geo[0] = a0; S.geo[1] = a1; S.geo[2] = b0; S.geo[3] = b1;
        S.geo[4] = (xcA_ - h >= 0 && xcB_ + h < im.cols) ? 1 : 0;   // taps are consecutive source pixels (no reflection at an
        S.geo[5] = (ycA_ - h >= 0 && ycB_ + h < im.rows) ? 1 : 0;   // image border) in x / in y
    }
    __syncthreads();
    const int sx0 = S.geo[0], sx1 = S.geo[1], sy0 = S.geo[2], sy1 = S.geo[3];
    const bool contigX = S.geo[4] != 0, contigY = S.geo[5] != 0;
    const int ax0 = sx0 & ~15;                         // 16-byte aligned window start
    const int nVec = (sx1 + 1 - ax0 + 15) >> 4;        // uint4 per row
    const int nRows = sy1 - sy0 + 1;

    if (tid < 51) S.taps[tid] = kc->taps[tid];
    if (tid == 0) { S.qn = 0; S.qt = 0; S.nNe = 0; S.anySrc = 0; }
    if (tid < gw) {   // tap positions of column tid (window-relative); consecutive unless reflected at an image border
        const int xc = lsdb_x86_d2i(floor((gxs + tid) / sca + 0.5));
        for (int i = 0; i < 17; i++) S.idxX[tid * 17 + i] = (short)((contigX ? xc - h + i : lsdb_reflect(xc - h + i, im.cols)) - ax0);
    } else if (tid >= 64 && tid < 64 + gh) {
        const int r = tid - 64;
        const int yc = lsdb_x86_d2i(floor((gys + r) / sca + 0.5));
        for (int i = 0; i < 17; i++) S.idxY[r * 17 + i] = (short)((contigY ? yc - h + i : lsdb_reflect(yc - h + i, im.rows)) - sy0);
    }
    for (int o = tid; o < nRows * ROW_WORDS; o += NT) S.rowBits[o] = 0u;
    for (int o = tid; o < GW * ROW_WORDS; o += NT) S.colBits[o] = 0u;
    __syncthreads();

    // ---- stage the source window, applying the remap 1->255, 255->0 for y>=1, x>=1 (:135-142)
    {
        const uint8_t* base = src + im.srcOff;
        unsigned int any = 0;
        for (int o = tid; o < nRows * nVec; o += NT) {
            int r = o / nVec, v = o - r * nVec;
            int gy = sy0 + r;
            const uint4 q = *reinterpret_cast<const uint4*>(base + (size_t)gy * im.srcPitch + ax0 + 16 * v);
            unsigned int w[4] = {q.x, q.y, q.z, q.w};
            if ((q.x | q.y | q.z | q.w) == 0u) continue;   // free space: stays 0, nothing to flag (the common case)
            if (gy >= 1) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    unsigned int e1 = __vcmpeq4(w[k], 0x01010101u), e255 = __vcmpeq4(w[k], 0xffffffffu);
                    unsigned int rmp = (w[k] & ~(e1 | e255)) | e1;
                    if (k == 0 && v == 0 && ax0 == 0) rmp = (rmp & 0xffffff00u) | (w[k] & 0xffu);  // column 0 untouched
                    w[k] = rmp;
                }
            }
            const unsigned int nz = nz4(w[0]) | (nz4(w[1]) << 4) | (nz4(w[2]) << 8) | (nz4(w[3]) << 12);
            if (nz) {
                atomicOr(&S.rowBits[r * ROW_WORDS + (v >> 1)], nz << ((v & 1) * 16));
                any = 1;
            }
        }
        if (any) S.anySrc = 1;
    }
    __syncthreads();
    const bool tileHasData = S.anySrc != 0;   // uniform

    if (tileHasData) {
        // ---- X pass: aux[r][c] = sum_i src[r][idxX[c][i]] * ker_phase(c)[i]   (:420-448), non-zero taps only.
        // The rows that hold a non-zero pixel are listed first; their outputs are spread over all threads.
        for (int r = tid; r < nRows; r += NT) {
            const unsigned int* rb = &S.rowBits[r * ROW_WORDS];
            if (rb[0] | rb[1] | rb[2] | rb[3] | rb[4]) S.neRows[atomicAdd(&S.nNe, 1)] = (unsigned char)r;
        }
        __syncthreads();
        // element (row k of the non-empty list, column c): a warp takes a row for columns 0..31, the 33rd column of all
        // rows is swept afterwards, one row per thread — no integer division by the (variable) tile width
        const int nNe = S.nNe;
        for (int pass = 0; pass < 2; pass++) {
            const int kBeg = pass == 0 ? warp : tid, kStep = pass == 0 ? NT / 32 : NT;
            const int c = pass == 0 ? lane : 32;
            if (c >= gw) continue;
            for (int k = kBeg; k < nNe; k += kStep) {
                const int r = S.neRows[k];
                const unsigned int* rb = &S.rowBits[r * ROW_WORDS];
                const short* ix = &S.idxX[c * 17];
                unsigned int m;
                if (contigX) m = bits17(rb, ix[0]);
                else {
                    m = 0;
#pragma unroll
                    for (int i = 0; i < 17; i++) { const int p = ix[i]; m |= ((rb[p >> 5] >> (p & 31)) & 1u) << i; }
                }
                if (m) {
                    // the few non-zero taps are re-read from global memory (L1/L2 hits: the window was just staged) and remapped
                    // on the fly; only the bit vectors live in shared memory, which buys a fourth resident CTA per SM
                    const int gy = sy0 + r;
                    const uint8_t* row = src + im.srcOff + (size_t)gy * im.srcPitch + ax0;
                    const double* ker = &S.taps[((gxs + c) % 3) * 17];
                    double v = 0.0;
                    while (m) {
                        const int i = __ffs(m) - 1;
                        m &= m - 1;
                        const int px = ix[i];
                        unsigned int b = row[px];
                        if (gy >= 1 && ax0 + px >= 1) b = b == 1u ? 255u : b;   // :135-142 (255 -> 0 never has its bit set)
                        v += (double)b * ker[i];
                    }
                    S.u.aux[r * GW + c] = v;
                    atomicOr(&S.colBits[c * ROW_WORDS + (r >> 5)], 1u << (r & 31));
                }
            }
        }
        __syncthreads();

        // ---- Y pass: g[r][c] = sum_i aux[idxY[r][i]][c] * ker_phase(r)[i]   (:452-482), non-zero taps only
        for (int pass = 0; pass < 2; pass++) {
            const int rBeg = pass == 0 ? warp : tid, rStep = pass == 0 ? NT / 32 : NT;
            const int c = pass == 0 ? lane : 32;
            if (c >= gw) continue;
            for (int r = rBeg; r < gh; r += rStep) {
                const short* iy = &S.idxY[r * 17];
                const unsigned int* cb = &S.colBits[c * ROW_WORDS];
                unsigned int m;
                if (contigY) m = bits17(cb, iy[0]);
                else {
                    m = 0;
#pragma unroll
                    for (int i = 0; i < 17; i++) { const int p = iy[i]; m |= ((cb[p >> 5] >> (p & 31)) & 1u) << i; }
                }
                double v = 0.0;
                if (m) {
                    const double* ker = &S.taps[((gys + r) % 3) * 17];
                    while (m) {
                        const int i = __ffs(m) - 1;
                        m &= m - 1;
                        v += S.u.aux[iy[i] * GW + c] * ker[i];
                    }
                }
                S.g[r * GW + c] = v;
                if (gaussOut) {
                    int gx = gxs + c, gy = gys + r;
                    if (gx >= x0 && gy >= y0) gaussOut[im.nOff + (size_t)gy * im.W + gx] = v;
                }
            }
        }
    } else {
        for (int o = tid; o < gh * gw; o += NT) {
            S.g[o / gw * GW + o % gw] = 0.0;
            if (gaussOut) {
                int gx = gxs + o % gw, gy = gys + o / gw;
                if (gx >= x0 && gy >= y0) gaussOut[im.nOff + (size_t)gy * im.W + gx] = 0.0;
            }
        }
    }
    __syncthreads();   // aux is dead from here on: its storage becomes the output tiles

    // ---- gradient, threshold mask, maxGrad (:151-174); pixels that need atan2 / cos / sin are queued
    const double gradThre = kc->gradThre, pi = kc->pi;
    double tmax = 0.0;
    for (int ly = warp; ly < LSDB_TILE; ly += NT / 32) {
        const int x = x0 + lane, y = y0 + ly;
        const int t = ly * 32 + lane;
        bool banned = true;   // pixels beyond the row end read as banned in the bit plane
        bool nonzero = false; // ... and as zero in the "mag != 0" plane of the ordering stage
        if (x < x1 && y < y1) {
            double m = 0.0;
            unsigned int st = 0;
            bool need = false, axis = false;
            if (x >= 1 && y >= 1) {
                const int gr = y - gys, gc = x - gxs;
                const double A = S.g[gr * GW + gc], B = S.g[gr * GW + gc - 1];
                const double C = S.g[(gr - 1) * GW + gc], D = S.g[(gr - 1) * GW + gc - 1];
                const double gradX = (B + D - A - C) / 2.0;
                const double gradY = (C + D - A - B) / 2.0;
                if (__double_as_longlong(gradX) == 0 && __double_as_longlong(gradY) == 0) {
                    st = LSDB_ST_BAN;  // mag = 0 < gradThre; atan2(+0,-0) = pi -> reset to 0 (:169-171)
                } else {
                    m = sqrt(gradX * gradX + gradY * gradY);
                    if (m < gradThre) st = LSDB_ST_BAN;
                    tmax = fmax(tmax, m);
                    need = true;
                    axis = gradX == 0.0 || gradY == 0.0;   // atan2 is then exactly 0, pi or +-pi/2
                }
            } else {
                need = true; axis = true;   // row 0 / column 0: mag = deg = 0, growable — cos/sin of 0 for RegionGrower's sums
            }
            S.u.out.magT[t] = m;
            nonzero = m != 0.0;
            S.u.out.degT[t] = 0.0;
            S.stT[t] = (unsigned char)st;
            banned = st != 0;
            if (need) {
                if (axis) S.queue[1023 - atomicAdd(&S.qt, 1)] = (unsigned short)t;
                else S.queue[atomicAdd(&S.qn, 1)] = (unsigned short)t;
            }
        }
        // usedMap==1 as one bit per pixel, row-pitched: the region pipeline keeps this plane in shared memory
        const unsigned int bal = __ballot_sync(0xffffffffu, banned);
        const unsigned int nzb = __ballot_sync(0xffffffffu, nonzero);
        if (lane == 0 && y < y1) { banBits[im.banOff + (size_t)y * im.pw + (x0 >> 5)] = bal; nzBits[im.banOff + (size_t)y * im.pw + (x0 >> 5)] = nzb; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    if (lane == 0) S.wmax[warp] = tmax;
    __syncthreads();
    if (tid == 0) {
        double m = S.wmax[0];
        for (int k = 1; k < 8; k++) m = fmax(m, S.wmax[k]);
        if (m > 0.0) atomicMax(&dyn[imgIdx].maxGradBits, (unsigned long long)__double_as_longlong(m));
    }

    // ---- level-line angle (:169-171) and the addends of RegionGrower's running sums (:515-516,:545-546), compacted:
    // first the general gradients (double-double atan2 / cos / sin, full lanes) ...
    const int qn = S.qn, qt = S.qt;
    for (int k = tid; k < qn; k += NT) {
        const int t = S.queue[k];
        const int gr = y0 + (t >> 5) - gys, gc = x0 + (t & 31) - gxs;
        const double A = S.g[gr * GW + gc], B = S.g[gr * GW + gc - 1];
        const double C = S.g[(gr - 1) * GW + gc], D = S.g[(gr - 1) * GW + gc - 1];
        const double gradX = (B + D - A - C) / 2.0;
        const double gradY = (C + D - A - B) / 2.0;
        double d = lsdm_atan2(gradX, -gradY);
        if (fabs(d - pi) < 0.000001) d = 0.0;
        S.u.out.degT[t] = d;
        if (S.stT[t] == 0) {
            S.u.out.cosT[t] = lsdm_cos(d);
            S.u.out.sinT[t] = lsdm_sin(d);
        }
    }
    // ... then the axis-aligned ones: atan2 returns exactly 0, pi (reset to 0) or +-pi/2; cos/sin of those come from
    // kc->axisCS, which the host filled with the same lsdm_cos / lsdm_sin
    for (int k = tid; k < qt; k += NT) {
        const int t = S.queue[1023 - k];
        const int x = x0 + (t & 31), y = y0 + (t >> 5);
        int sel = 0;   // 0: d = 0, 1: d = +pi/2, 2: d = -pi/2
        if (x >= 1 && y >= 1) {
            const int gr = y - gys, gc = x - gxs;
            const double A = S.g[gr * GW + gc], B = S.g[gr * GW + gc - 1];
            const double C = S.g[(gr - 1) * GW + gc], D = S.g[(gr - 1) * GW + gc - 1];
            const double gradX = (B + D - A - C) / 2.0;
            // atan2(gradX, -gradY): gradX == 0 -> 0 or pi (-> 0);  gradY == 0 (gradX != 0) -> +-pi/2 by the sign of gradX
            if (gradX != 0.0) sel = gradX > 0.0 ? 1 : 2;
        }
        S.u.out.degT[t] = kc->axisDeg[sel];
        if (S.stT[t] == 0) {
            S.u.out.cosT[t] = kc->axisCS[2 * sel];
            S.u.out.sinT[t] = kc->axisCS[2 * sel + 1];
        }
    }
    __syncthreads();

    // ---- coalesced write-out
    for (int ly = warp; ly < LSDB_TILE; ly += NT / 32) {
        const int x = x0 + lane, y = y0 + ly;
        if (x < x1 && y < y1) {
            const int t = ly * 32 + lane;
            const size_t p = im.nOff + (size_t)y * im.W + x;
            const unsigned int st = S.stT[t];
            mag[p] = S.u.out.magT[t];
            deg[p] = S.u.out.degT[t];
            state[p] = st;
            if (st == 0) {
                cosm[2 * p] = S.u.out.cosT[t];   // one interleaved (cos, sin) plane: sinm == cosm + 1
                sinm[2 * p] = S.u.out.sinT[t];
            }
        }
    }
}

int lsdb_stencil_mode(void) {
    // LSDB_STENCIL=1: this file's kernel (the first cut); anything else: stencil2.cu.  LSDB_STENCIL_G: tiles per CTA of stencil2.cu
    // (1, 2 or 4; default 1).  LSDB_STENCIL_DEFER=0 keeps the deferred pixels of stencil2.cu inside their tiles.
    const char* e = getenv("LSDB_STENCIL");
    const int version = e && e[0] == '1' ? 1 : 2;
    e = getenv("LSDB_STENCIL_G");
    int g = e ? atoi(e) : 1;
    if (g != 2 && g != 4) g = 1;
    e = getenv("LSDB_STENCIL_DEFER");   // 0: off; n >= 2: on, the list capped at n records (exercises the overflow path); else: on
    const int defer = !(e && e[0] == '0' && e[1] == 0);
    int cap = e ? atoi(e) : 0;
    if (cap < 2 || cap > (1 << 20)) cap = 0;
    return version | (g << 2) | (defer << 5) | (cap << 8);
}

int lsdb_launch_stencil(cudaStream_t s, int nTiles, const LsdbImg* imgs, const int* tileImg, LsdbImgDyn* dyn,
                         const LsdbLsdConst* kc, const uint8_t* src, double* mag, double* deg, double* cosm, double* sinm,
                         unsigned int* state, unsigned int* banBits, unsigned int* nzBits, double* gaussOut, int tileBase,
                         void* deferBuf, size_t deferBytes, int* deferCount, int mode) {
    const size_t capBytes = (size_t)(mode >> 8) * 24;   // 24 = bytes per deferred record
    if (capBytes && capBytes < deferBytes) deferBytes = capBytes;
    if ((mode & 3) != 1)
        return lsdb_launch_stencil_v2(s, nTiles, imgs, tileImg, dyn, kc, src, mag, deg, cosm, sinm, state, banBits, nzBits, gaussOut, tileBase,
                                      (mode >> 5) & 1 ? deferBuf : nullptr, deferBytes, deferCount, (mode >> 2) & 7);
    cudaFuncSetAttribute(lsdb_stencil_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StencilSmem));   // per device, cheap
    if (nTiles > 0)
        lsdb_stencil_kernel<<<nTiles, NT, sizeof(StencilSmem), s>>>(imgs, tileImg, dyn, kc, src, mag, deg, cosm, sinm, state, banBits, nzBits, gaussOut, tileBase);
    return nTiles > 0 ? 1 : 0;
}
