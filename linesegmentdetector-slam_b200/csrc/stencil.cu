// Stage A of the LSD hot path on sm_100a: value remap + separable decimating Gaussian + gradient /
// level-line angle / threshold mask, fused in one shared-memory-staged stencil kernel.
//
// Replaces (reference, /root/reference/LSD/myLSD.cpp): the remap loop :135-142, GaussianSampler
// :378-484 (X pass :420-448, Y pass :452-482) and the gradient loop :151-174 incl. maxGrad.
// Bit-exactness rules: taps are accumulated i = 0..16 in order with separate mul and add
// (-fmad=false), the source byte is converted exactly, atan2 comes from lsd_math.h.  Skipping a
// window whose 17 inputs are all zero is exact (+0*k = +0, v + +0 = v).
//
// One CTA per 32x32 tile of the scaled image.  The CTA stages the <=125x125 source window (16-byte
// vector loads, remap applied on the fly) in shared memory, runs the X pass into a shared
// [rows][33] f64 strip, the Y pass into a shared 33x33 f64 tile (one halo row/column for the 2x2
// gradient stencil) and writes mag/deg/state with fully coalesced 256-byte rows.
// HBM traffic per source pixel: 1 B read + 0.09*(8+8+4) B written.
#include "lsdb_common.cuh"

#define SRC_PITCH 160
#define GW 33

struct StencilSmem {
    double aux[LSDB_SRC_MAX * GW];
    double g[GW * GW];
    double taps[3 * 17];
    unsigned char src[LSDB_SRC_MAX * SRC_PITCH];
    short idxX[GW * 17];
    short idxY[GW * 17];
    double wmax[8];
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

__global__ void __launch_bounds__(256) lsdb_stencil_kernel(const LsdbImg* __restrict__ imgs, const int* __restrict__ tileImg,
                                                           LsdbImgDyn* __restrict__ dyn, const LsdbLsdConst* __restrict__ kc,
                                                           const uint8_t* __restrict__ src, double* __restrict__ mag,
                                                           double* __restrict__ deg, double* __restrict__ cosm,
                                                           double* __restrict__ sinm, unsigned int* __restrict__ state,
                                                           unsigned int* __restrict__ banBits, double* __restrict__ gaussOut) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StencilSmem& S = *reinterpret_cast<StencilSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int imgIdx = tileImg[blockIdx.x];
    const LsdbImg im = imgs[imgIdx];
    const int lt = blockIdx.x - im.tile0;
    const int tx = lt % im.tilesX, ty = lt / im.tilesX;
    const int x0 = tx * LSDB_TILE, y0 = ty * LSDB_TILE;
    const int x1 = min(x0 + LSDB_TILE, im.W), y1 = min(y0 + LSDB_TILE, im.H);
    const int gxs = max(x0 - 1, 0), gys = max(y0 - 1, 0);
    const int gw = x1 - gxs, gh = y1 - gys;
    const double sca = kc->sca;
    const int h = kc->h;  // 8

    // centres of the first / last Gaussian column and row: xc = floor(x/sca + 0.5)  (:428,:460)
    const int xcA = lsdb_x86_d2i(floor(gxs / sca + 0.5)), xcB = lsdb_x86_d2i(floor((x1 - 1) / sca + 0.5));
    const int ycA = lsdb_x86_d2i(floor(gys / sca + 0.5)), ycB = lsdb_x86_d2i(floor((y1 - 1) / sca + 0.5));
    int sx0, sx1, sy0, sy1;
    lsdb_window(xcA, xcB, h, im.cols, &sx0, &sx1);
    lsdb_window(ycA, ycB, h, im.rows, &sy0, &sy1);
    const int ax0 = sx0 & ~15;                         // 16-byte aligned window start
    const int nVec = (sx1 + 1 - ax0 + 15) >> 4;        // uint4 per row
    const int nRows = sy1 - sy0 + 1;

    if (tid < 51) S.taps[tid] = kc->taps[tid];
    for (int o = tid; o < gw * 17; o += 256) {
        int c = o / 17, i = o - c * 17;
        int xc = lsdb_x86_d2i(floor((gxs + c) / sca + 0.5));
        S.idxX[o] = (short)(lsdb_reflect(xc - h + i, im.cols) - ax0);
    }
    for (int o = tid; o < gh * 17; o += 256) {
        int r = o / 17, i = o - r * 17;
        int yc = lsdb_x86_d2i(floor((gys + r) / sca + 0.5));
        S.idxY[o] = (short)(lsdb_reflect(yc - h + i, im.rows) - sy0);
    }
    // ---- stage the source window, applying the remap 1->255, 255->0 for y>=1, x>=1 (:135-142)
    {
        const uint8_t* base = src + im.srcOff;
        for (int o = tid; o < nRows * nVec; o += 256) {
            int r = o / nVec, v = o - r * nVec;
            int gy = sy0 + r;
            const uint4 q = *reinterpret_cast<const uint4*>(base + (size_t)gy * im.srcPitch + ax0 + 16 * v);
            unsigned int w[4] = {q.x, q.y, q.z, q.w};
            if (gy >= 1) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    unsigned int e1 = __vcmpeq4(w[k], 0x01010101u), e255 = __vcmpeq4(w[k], 0xffffffffu);
                    unsigned int rmp = (w[k] & ~(e1 | e255)) | e1;
                    if (k == 0 && v == 0 && ax0 == 0) rmp = (rmp & 0xffffff00u) | (w[k] & 0xffu);  // column 0 untouched
                    w[k] = rmp;
                }
            }
            *reinterpret_cast<uint4*>(&S.src[r * SRC_PITCH + 16 * v]) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    __syncthreads();

    // ---- X pass: aux[r][c] = sum_i src[r][idxX[c][i]] * ker_phase(c)[i]   (:420-448)
    for (int o = tid; o < nRows * gw; o += 256) {
        int r = o / gw, c = o - r * gw;
        const unsigned char* row = &S.src[r * SRC_PITCH];
        const short* ix = &S.idxX[c * 17];
        unsigned int b[17];
        unsigned int any = 0;
#pragma unroll
        for (int i = 0; i < 17; i++) { b[i] = row[ix[i]]; any |= b[i]; }
        double v = 0.0;
        if (any) {
            const double* ker = &S.taps[((gxs + c) % 3) * 17];
#pragma unroll
            for (int i = 0; i < 17; i++) v += (double)b[i] * ker[i];
        }
        S.aux[r * GW + c] = v;
    }
    __syncthreads();

    // ---- Y pass: g[r][c] = sum_i aux[idxY[r][i]][c] * ker_phase(r)[i]   (:452-482)
    for (int o = tid; o < gh * gw; o += 256) {
        int r = o / gw, c = o - r * gw;
        const short* iy = &S.idxY[r * 17];
        double a[17];
        bool any = false;
#pragma unroll
        for (int i = 0; i < 17; i++) { a[i] = S.aux[iy[i] * GW + c]; any |= (a[i] != 0.0); }
        double v = 0.0;
        if (any) {
            const double* ker = &S.taps[((gys + r) % 3) * 17];
#pragma unroll
            for (int i = 0; i < 17; i++) v += a[i] * ker[i];
        }
        S.g[r * GW + c] = v;
        if (gaussOut) {
            int gx = gxs + c, gy = gys + r;
            if (gx >= x0 && gy >= y0) gaussOut[im.nOff + (size_t)gy * im.W + gx] = v;
        }
    }
    __syncthreads();

    // ---- gradient, level-line angle, threshold mask, maxGrad   (:151-174)
    const double gradThre = kc->gradThre, pi = kc->pi;
    double tmax = 0.0;
    const int lx = tid & 31;
    for (int ly = tid >> 5; ly < LSDB_TILE; ly += 8) {
        int x = x0 + lx, y = y0 + ly;
        bool banned = true;   // pixels beyond the row end read as banned in the bitmap
        if (x < x1 && y < y1) {
            double m = 0.0, d = 0.0;
            unsigned int st = 0;
            if (x >= 1 && y >= 1) {
                int gr = y - gys, gc = x - gxs;
                double A = S.g[gr * GW + gc], B = S.g[gr * GW + gc - 1];
                double C = S.g[(gr - 1) * GW + gc], D = S.g[(gr - 1) * GW + gc - 1];
                double gradX = (B + D - A - C) / 2.0;
                double gradY = (C + D - A - B) / 2.0;
                if (__double_as_longlong(gradX) == 0 && __double_as_longlong(gradY) == 0) {
                    st = LSDB_ST_BAN;  // mag = 0 < gradThre; atan2(+0,-0) = pi -> reset to 0 (:169-171)
                } else {
                    m = sqrt(gradX * gradX + gradY * gradY);
                    if (m < gradThre) st = LSDB_ST_BAN;
                    d = lsdm_atan2(gradX, -gradY);
                    if (fabs(d - pi) < 0.000001) d = 0.0;
                    tmax = fmax(tmax, m);
                }
            }
            size_t p = im.nOff + (size_t)y * im.W + x;
            mag[p] = m;
            deg[p] = d;
            state[p] = st;
            banned = st != 0;
            if (st == 0) {  // growable pixel: the addends of RegionGrower's running sums (:515-516,:545-546)
                cosm[p] = lsdm_cos(d);
                sinm[p] = lsdm_sin(d);
            }
        }
        // usedMap==1 as one bit per pixel, row-pitched: the region pipeline keeps this plane in shared memory
        const unsigned int bal = __ballot_sync(0xffffffffu, banned);
        if (lx == 0 && y < y1) banBits[im.banOff + (size_t)y * im.pw + (x0 >> 5)] = bal;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    if (lx == 0) S.wmax[tid >> 5] = tmax;
    __syncthreads();
    if (tid == 0) {
        double m = S.wmax[0];
        for (int k = 1; k < 8; k++) m = fmax(m, S.wmax[k]);
        if (m > 0.0) atomicMax(&dyn[imgIdx].maxGradBits, (unsigned long long)__double_as_longlong(m));
    }
}

void lsdb_launch_stencil(cudaStream_t s, int nTiles, const LsdbImg* imgs, const int* tileImg, LsdbImgDyn* dyn,
                         const LsdbLsdConst* kc, const uint8_t* src, double* mag, double* deg, double* cosm, double* sinm,
                         unsigned int* state, unsigned int* banBits, double* gaussOut) {
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(lsdb_stencil_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StencilSmem));
        attr = true;
    }
    if (nTiles > 0)
        lsdb_stencil_kernel<<<nTiles, 256, sizeof(StencilSmem), s>>>(imgs, tileImg, dyn, kc, src, mag, deg, cosm, sinm, state, banBits, gaussOut);
}
