// Stage B of the LSD hot path on sm_100a: the pseudo-ordering of seed pixels.
//
// Replaces LSD/myLSD.cpp:176-204 (bin quantisation, compaction, qsort with Comp :486-489).  The
// reference's comparator never returns 0, and glibc's qsort (a merge sort for these sizes) then
// yields exactly (bin descending, raster index ascending) — SURVEY.md §0 fact 3 — so a STABLE
// counting sort on the 1024 bins reproduces the reference seed order bit-for-bit, with a
// deterministic tie-break instead of quicksort's unspecified one.
//
// A map is cut into K bands of rows, one 16-warp CTA per band (K grows as the batch shrinks, so a lone
// 4096^2 or 16384^2 map still covers the device); warp w of band k owns a contiguous run of rows, i.e.
// a contiguous raster segment.  Three launches:
//   count    bin = min(floor(mag * (pseBin/maxGrad)), pseBin) per pixel; per-(warp,bin) counts -> global
//   scan     one CTA per map: exclusive scan over bins (descending) and, inside a bin, over the warps of
//            all bands in raster order; the counts become write offsets
//   scatter  the same walk again: stable scatter, rank inside a warp step by __match_any_sync
// Only the pixels with a non-zero gradient are touched: the stencil stage leaves one bit per pixel
// ("mag != 0", ~15 % set on occupancy grids), a warp reads a row's bit words, skips the empty ones and
// gathers the 8-byte magnitudes of the set bits only; their bins are left packed (2 bytes per NON-ZERO pixel) for the
// scatter pass, which therefore reads bit words and packed bins only.
// Algorithmic bytes: 8n (mag) read + 4c (seed list) written; DRAM traffic here is below
// that: 2 x n/8 + the touched sectors of mag + 2 x 2c' + 4c + the count tables (c' = non-zero pixels).
#include "lsdb_common.cuh"

#define ORDER_WARPS 16          // warps per band CTA: 64 KB of per-(warp,bin) counters, three CTAs per SM
#define ORDER_NT (ORDER_WARPS * 32)
#define ORDER_BINS 1025

// rows [r0, r1) of warp w in band k of a map with H rows cut into K bands
__device__ __forceinline__ void order_rows(int H, int K, int k, int w, int* r0, int* r1) {
    const int per = (H + K * ORDER_WARPS - 1) / (K * ORDER_WARPS);
    const int a = (k * ORDER_WARPS + w) * per;
    *r0 = a < H ? a : H;
    *r1 = a + per < H ? a + per : H;
}

// WRITE = false: count;  WRITE = true: scatter with the offsets in tab.
// The rows [r0, r1) of a warp are one contiguous run of words of the bit plane: walked 32 words at a time (one per lane),
// the next 32 fetched while the current ones are worked on.
// The count pass leaves the bins of the non-zero pixels of a warp's rows packed (u16, in raster order) at the start of the
// warp's own stretch of the bin plane; the scatter pass reads them back in the same order instead of gathering and
// quantising the magnitudes a second time.
template <bool WRITE>
__device__ __forceinline__ void order_walk(const LsdbImg& im, const double* __restrict__ m, const unsigned int* __restrict__ nz, double zoom,
                                           int pseBin, int r0, int r1, unsigned int* myTab, unsigned int* __restrict__ out,
                                           unsigned short* __restrict__ packed, int lane) {
    const int pw = im.pw, W = im.W;
    const int w0 = r0 * pw, w1 = r1 * pw;
    packed += (size_t)r0 * W;
    unsigned int nPacked = 0;   // non-zero pixels of this warp so far (warp-uniform)
    const unsigned int ltm = (1u << lane) - 1u;
    unsigned int next = w0 + lane < w1 ? nz[w0 + lane] : 0u;
    for (int wb = w0; wb < w1; wb += 32) {
        const unsigned int word = next;
        next = wb + 32 + lane < w1 ? nz[wb + 32 + lane] : 0u;
        const int wi = wb + lane, y = wi / pw;
        const unsigned int pbase = (unsigned int)y * (unsigned int)W + (unsigned int)(wi - y * pw) * 32u;   // raster index of the word's first pixel
        unsigned int busy = __ballot_sync(0xffffffffu, word != 0u);
        while (busy) {
            // up to four non-empty words per step: their gathers are in flight together
            unsigned int ps[4], slot[4]; double v[4]; bool on[4]; int tq[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                on[k] = false; v[k] = 0.0; ps[k] = 0; slot[k] = 0; tq[k] = 0;
                if (busy) {
                    const int j = __ffs(busy) - 1;
                    busy &= busy - 1;
                    const unsigned int wv = __shfl_sync(0xffffffffu, word, j);
                    ps[k] = __shfl_sync(0xffffffffu, pbase, j) + (unsigned int)lane;
                    on[k] = (wv >> lane) & 1u;                 // bits beyond the row end are never set
                    slot[k] = nPacked + __popc(wv & ltm);
                    nPacked += __popc(wv);
                    if (on[k]) { if (WRITE) tq[k] = packed[slot[k]]; else v[k] = m[ps[k]]; }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int t = tq[k];
                if (!WRITE && on[k]) {
                    t = lsdb_x86_d2i(floor(v[k] * zoom));   // :182-184
                    if (t > pseBin) t = pseBin;
                    t &= 0xffff;                            // pseIdx is CV_16UC1 (:178,187)
                    packed[slot[k]] = (unsigned short)t;
                }
                if (WRITE) {
                    // stable: the lanes of one bin keep their raster order (rank inside the group), the group's run follows
                    // the bin's earlier pixels of this warp
                    const unsigned int grp = __match_any_sync(0xffffffffu, t);
                    unsigned int pos = 0;
                    if (t != 0) pos = myTab[t] + __popc(grp & ((1u << lane) - 1u));
                    __syncwarp();
                    if (t != 0) {
                        out[pos] = ps[k];
                        if (lane == __ffs(grp) - 1) myTab[t] += __popc(grp);
                    }
                    __syncwarp();
                } else if (t != 0) {
                    atomicAdd(&myTab[t], 1u);   // counting needs no order: the warp's private row of the table, conflicts resolved by the hardware
                }
            }
        }
    }
}

__global__ void __launch_bounds__(ORDER_NT, 3) lsdb_order_count_kernel(const LsdbImg* __restrict__ imgs, const LsdbImgDyn* __restrict__ dyn,
                                                                const LsdbLsdConst* __restrict__ kc, const double* __restrict__ mag,
                                                                const unsigned int* __restrict__ nzBits, const int2* __restrict__ bandOf,
                                                                const int2* __restrict__ bandsOfImg, unsigned int* __restrict__ tabs,
                                                                unsigned short* __restrict__ bins) {
    extern __shared__ unsigned int cnt[];  // [ORDER_WARPS][ORDER_BINS]
    const int2 bo = bandOf[blockIdx.x];    // (map, band)
    const LsdbImg im = imgs[bo.x];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const double maxGrad = __longlong_as_double((long long)dyn[bo.x].maxGradBits);
    unsigned int* gt = tabs + (size_t)blockIdx.x * ORDER_WARPS * ORDER_BINS;
    if (!(maxGrad > 0.0)) {   // blank map: no seeds (the scan kernel sets nCells = 0)
        for (int i = tid; i < ORDER_WARPS * ORDER_BINS; i += ORDER_NT) gt[i] = 0;
        return;
    }
    const int pseBin = kc->pseBin;
    const double zoom = 1.0 * pseBin / maxGrad;  // :179
    for (int i = tid; i < ORDER_WARPS * ORDER_BINS; i += ORDER_NT) cnt[i] = 0;
    __syncthreads();
    int r0, r1;
    order_rows(im.H, bandsOfImg[bo.x].y, bo.y, w, &r0, &r1);
    order_walk<false>(im, mag + im.nOff, nzBits + im.banOff, zoom, pseBin, r0, r1, cnt + w * ORDER_BINS, 0, bins + im.nOff, lane);
    __syncthreads();
    for (int i = tid; i < ORDER_WARPS * ORDER_BINS; i += ORDER_NT) gt[i] = cnt[i];
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

__global__ void __launch_bounds__(ORDER_NT, 3) lsdb_order_scatter_kernel(const LsdbImg* __restrict__ imgs, const LsdbImgDyn* __restrict__ dyn,
                                                                  const LsdbLsdConst* __restrict__ kc, const double* __restrict__ mag,
                                                                  const unsigned int* __restrict__ nzBits, const int2* __restrict__ bandOf,
                                                                  const int2* __restrict__ bandsOfImg, const unsigned int* __restrict__ tabs,
                                                                  unsigned short* __restrict__ bins, unsigned int* __restrict__ cells) {
    extern __shared__ unsigned int cnt[];  // [ORDER_WARPS][ORDER_BINS]: write offsets
    const int2 bo = bandOf[blockIdx.x];
    const LsdbImg im = imgs[bo.x];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const double maxGrad = __longlong_as_double((long long)dyn[bo.x].maxGradBits);
    if (!(maxGrad > 0.0)) return;
    const int pseBin = kc->pseBin;
    const double zoom = 1.0 * pseBin / maxGrad;
    const unsigned int* gt = tabs + (size_t)blockIdx.x * ORDER_WARPS * ORDER_BINS;
    for (int i = tid; i < ORDER_WARPS * ORDER_BINS; i += ORDER_NT) cnt[i] = gt[i];
    __syncthreads();
    int r0, r1;
    order_rows(im.H, bandsOfImg[bo.x].y, bo.y, w, &r0, &r1);
    order_walk<true>(im, mag + im.nOff, nzBits + im.banOff, zoom, pseBin, r0, r1, cnt + w * ORDER_BINS, cells + im.nOff, bins + im.nOff, lane);
}

size_t lsdb_order_tab_words_per_band(void) { return (size_t)ORDER_WARPS * ORDER_BINS; }

void lsdb_launch_order(cudaStream_t s, int nImgs, int nBands, const LsdbImg* imgs, LsdbImgDyn* dyn, const LsdbLsdConst* kc,
                       const double* mag, const unsigned int* nzBits, const int2* bandOf, const int2* bandsOfImg, unsigned int* tabs,
                       unsigned short* bins, unsigned int* cells) {
    const int smem = ORDER_WARPS * ORDER_BINS * sizeof(unsigned int);
    cudaFuncSetAttribute(lsdb_order_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per device, cheap
    cudaFuncSetAttribute(lsdb_order_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (nImgs <= 0) return;
    lsdb_order_count_kernel<<<nBands, ORDER_NT, smem, s>>>(imgs, dyn, kc, mag, nzBits, bandOf, bandsOfImg, tabs, bins);
    lsdb_order_scan_kernel<<<nImgs, 1024, 0, s>>>(dyn, bandsOfImg, tabs);
    lsdb_order_scatter_kernel<<<nBands, ORDER_NT, smem, s>>>(imgs, dyn, kc, mag, nzBits, bandOf, bandsOfImg, tabs, bins, cells);
}
