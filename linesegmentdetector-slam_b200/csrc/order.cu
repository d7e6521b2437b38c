// Stage B of the LSD hot path on sm_100a: the pseudo-ordering of seed pixels.
//
// Replaces LSD/myLSD.cpp:176-204 (bin quantisation, compaction, qsort with Comp :486-489).  The
// reference's comparator never returns 0, and glibc's qsort (a merge sort for these sizes) then
// yields exactly (bin descending, raster index ascending) — SURVEY.md §0 fact 3 — so a STABLE
// counting sort on the 1024 bins reproduces the reference seed order bit-for-bit, with a
// deterministic tie-break instead of quicksort's unspecified one.
//
// A map is cut into K bands of rows, one 32-warp CTA per band (K grows as the batch shrinks, so a lone
// 4096^2 or 16384^2 map still covers the device); warp w of band k owns a contiguous run of rows, i.e.
// a contiguous raster segment.  Three launches:
//   count    bin = min(floor(mag * (pseBin/maxGrad)), pseBin) per pixel; per-(warp,bin) counts -> global
//   scan     one CTA per map: exclusive scan over bins (descending) and, inside a bin, over the warps of
//            all bands in raster order; the counts become write offsets
//   scatter  the same walk again: stable scatter, rank inside a warp step by __match_any_sync
// Only the pixels with a non-zero gradient are touched: the stencil stage leaves one bit per pixel
// ("mag != 0", ~15 % set on occupancy grids), a warp reads a row's bit words, skips the empty ones and
// gathers the 8-byte magnitudes of the set bits only — twice (no bin plane is written and re-read).
// Algorithmic bytes: 8n (mag) read + 4c (seed list) written; DRAM traffic here is below
// that: 2 x (n/8 + the touched sectors of mag) + 4c + the count tables.
#include "lsdb_common.cuh"

#define ORDER_WARPS 32
#define ORDER_BINS 1025

// rows [r0, r1) of warp w in band k of a map with H rows cut into K bands
__device__ __forceinline__ void order_rows(int H, int K, int k, int w, int* r0, int* r1) {
    const int per = (H + K * ORDER_WARPS - 1) / (K * ORDER_WARPS);
    const int a = (k * ORDER_WARPS + w) * per;
    *r0 = a < H ? a : H;
    *r1 = a + per < H ? a + per : H;
}

// WRITE = false: count;  WRITE = true: scatter with the offsets in tab
template <bool WRITE>
__device__ __forceinline__ void order_walk(const LsdbImg& im, const double* __restrict__ m, const unsigned int* __restrict__ nz, double zoom,
                                           int pseBin, int r0, int r1, unsigned int* myTab, unsigned int* __restrict__ out, int lane) {
    const int pw = im.pw, W = im.W;
    for (int y = r0; y < r1; y++) {
        const unsigned int* row = nz + (size_t)y * pw;
        for (int wb = 0; wb < pw; wb += 32) {
            const unsigned int word = wb + lane < pw ? row[wb + lane] : 0u;
            unsigned int busy = __ballot_sync(0xffffffffu, word != 0u);
            while (busy) {
                // up to four non-empty words per step: their gathers are in flight together
                int xs[4]; double v[4]; bool on[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    on[k] = false; v[k] = 0.0; xs[k] = 0;
                    if (busy) {
                        const int j = __ffs(busy) - 1;
                        busy &= busy - 1;
                        const unsigned int wv = __shfl_sync(0xffffffffu, word, j);
                        const int x = (wb + j) * 32 + lane;
                        xs[k] = x;
                        on[k] = (wv >> lane) & 1u;
                        if (on[k]) v[k] = m[(size_t)y * W + x];   // bits beyond the row end are never set
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    int t = 0;
                    if (on[k]) {
                        t = lsdb_x86_d2i(floor(v[k] * zoom));   // :182-184
                        if (t > pseBin) t = pseBin;
                        t &= 0xffff;                            // pseIdx is CV_16UC1 (:178,187)
                    }
                    const unsigned int grp = __match_any_sync(0xffffffffu, t);
                    if (WRITE) {
                        unsigned int pos = 0;
                        if (t != 0) pos = myTab[t] + __popc(grp & ((1u << lane) - 1u));
                        __syncwarp();
                        if (t != 0) {
                            out[pos] = (unsigned int)y * (unsigned int)W + (unsigned int)xs[k];
                            if (lane == __ffs(grp) - 1) myTab[t] += __popc(grp);
                        }
                    } else {
                        if (t != 0 && lane == __ffs(grp) - 1) myTab[t] += __popc(grp);
                    }
                    __syncwarp();
                }
            }
        }
    }
}

__global__ void __launch_bounds__(1024) lsdb_order_count_kernel(const LsdbImg* __restrict__ imgs, const LsdbImgDyn* __restrict__ dyn,
                                                                const LsdbLsdConst* __restrict__ kc, const double* __restrict__ mag,
                                                                const unsigned int* __restrict__ nzBits, const int2* __restrict__ bandOf,
                                                                const int2* __restrict__ bandsOfImg, unsigned int* __restrict__ tabs) {
    extern __shared__ unsigned int cnt[];  // [ORDER_WARPS][ORDER_BINS]
    const int2 bo = bandOf[blockIdx.x];    // (map, band)
    const LsdbImg im = imgs[bo.x];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const double maxGrad = __longlong_as_double((long long)dyn[bo.x].maxGradBits);
    unsigned int* gt = tabs + (size_t)blockIdx.x * ORDER_WARPS * ORDER_BINS;
    if (!(maxGrad > 0.0)) {   // blank map: no seeds (the scan kernel sets nCells = 0)
        for (int i = tid; i < ORDER_WARPS * ORDER_BINS; i += 1024) gt[i] = 0;
        return;
    }
    const int pseBin = kc->pseBin;
    const double zoom = 1.0 * pseBin / maxGrad;  // :179
    for (int i = tid; i < ORDER_WARPS * ORDER_BINS; i += 1024) cnt[i] = 0;
    __syncthreads();
    int r0, r1;
    order_rows(im.H, bandsOfImg[bo.x].y, bo.y, w, &r0, &r1);
    order_walk<false>(im, mag + im.nOff, nzBits + im.banOff, zoom, pseBin, r0, r1, cnt + w * ORDER_BINS, 0, lane);
    __syncthreads();
    for (int i = tid; i < ORDER_WARPS * ORDER_BINS; i += 1024) gt[i] = cnt[i];
}

// one CTA per map: counts -> write offsets.  thread tid <-> bin 1024 - tid (descending bins first)
__global__ void __launch_bounds__(1024) lsdb_order_scan_kernel(LsdbImgDyn* __restrict__ dyn, const int2* __restrict__ bandsOfImg,
                                                               unsigned int* __restrict__ tabs) {
    __shared__ unsigned int warpSum[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int2 bi = bandsOfImg[blockIdx.x];   // (first band CTA, number of bands)
    unsigned int* gt = tabs + (size_t)bi.x * ORDER_WARPS * ORDER_BINS;
    const int nSeg = bi.y * ORDER_WARPS;      // raster segments of the map, in order
    const int b = 1024 - tid;
    unsigned int total = 0;
    for (int k = 0; k < nSeg; k++) total += gt[(size_t)k * ORDER_BINS + b];
    unsigned int incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warpSum[w] = incl;
    __syncthreads();
    if (w == 0) {
        unsigned int s = warpSum[lane], si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        warpSum[lane] = si - s;
        if (lane == 31) dyn[blockIdx.x].nCells = (int)si;
    }
    __syncthreads();
    unsigned int running = warpSum[w] + incl - total;
    for (int k = 0; k < nSeg; k++) {
        const unsigned int c = gt[(size_t)k * ORDER_BINS + b];
        gt[(size_t)k * ORDER_BINS + b] = running;
        running += c;
    }
}

__global__ void __launch_bounds__(1024) lsdb_order_scatter_kernel(const LsdbImg* __restrict__ imgs, const LsdbImgDyn* __restrict__ dyn,
                                                                  const LsdbLsdConst* __restrict__ kc, const double* __restrict__ mag,
                                                                  const unsigned int* __restrict__ nzBits, const int2* __restrict__ bandOf,
                                                                  const int2* __restrict__ bandsOfImg, const unsigned int* __restrict__ tabs,
                                                                  unsigned int* __restrict__ cells) {
    extern __shared__ unsigned int cnt[];  // [ORDER_WARPS][ORDER_BINS]: write offsets
    const int2 bo = bandOf[blockIdx.x];
    const LsdbImg im = imgs[bo.x];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const double maxGrad = __longlong_as_double((long long)dyn[bo.x].maxGradBits);
    if (!(maxGrad > 0.0)) return;
    const int pseBin = kc->pseBin;
    const double zoom = 1.0 * pseBin / maxGrad;
    const unsigned int* gt = tabs + (size_t)blockIdx.x * ORDER_WARPS * ORDER_BINS;
    for (int i = tid; i < ORDER_WARPS * ORDER_BINS; i += 1024) cnt[i] = gt[i];
    __syncthreads();
    int r0, r1;
    order_rows(im.H, bandsOfImg[bo.x].y, bo.y, w, &r0, &r1);
    order_walk<true>(im, mag + im.nOff, nzBits + im.banOff, zoom, pseBin, r0, r1, cnt + w * ORDER_BINS, cells + im.nOff, lane);
}

size_t lsdb_order_tab_words_per_band(void) { return (size_t)ORDER_WARPS * ORDER_BINS; }

void lsdb_launch_order(cudaStream_t s, int nImgs, int nBands, const LsdbImg* imgs, LsdbImgDyn* dyn, const LsdbLsdConst* kc,
                       const double* mag, const unsigned int* nzBits, const int2* bandOf, const int2* bandsOfImg, unsigned int* tabs,
                       unsigned int* cells) {
    const int smem = ORDER_WARPS * ORDER_BINS * sizeof(unsigned int);
    cudaFuncSetAttribute(lsdb_order_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per device, cheap
    cudaFuncSetAttribute(lsdb_order_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (nImgs <= 0) return;
    lsdb_order_count_kernel<<<nBands, 1024, smem, s>>>(imgs, dyn, kc, mag, nzBits, bandOf, bandsOfImg, tabs);
    lsdb_order_scan_kernel<<<nImgs, 1024, 0, s>>>(dyn, bandsOfImg, tabs);
    lsdb_order_scatter_kernel<<<nBands, 1024, smem, s>>>(imgs, dyn, kc, mag, nzBits, bandOf, bandsOfImg, tabs, cells);
}
