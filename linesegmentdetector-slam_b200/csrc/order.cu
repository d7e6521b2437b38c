// Stage B of the LSD hot path on sm_100a: the pseudo-ordering of seed pixels.
//
// Replaces LSD/myLSD.cpp:176-204 (bin quantisation, compaction, qsort with Comp :486-489).  The
// reference's comparator never returns 0, and glibc's qsort (a merge sort for these sizes) then
// yields exactly (bin descending, raster index ascending) — SURVEY.md §0 fact 3 — so a STABLE
// counting sort on the 1024 bins reproduces the reference seed order bit-for-bit, with a
// deterministic tie-break instead of quicksort's unspecified one.
//
// One CTA (32 warps) per map.  Warp w owns the w-th contiguous raster segment:
//   pass 1  bin = min(floor(mag * (pseBin/maxGrad)), pseBin) -> u16 plane, per-(warp,bin) counts in smem
//   scan    block-wide exclusive scan over bins (descending) and over warps inside each bin
//   pass 2  stable scatter of the non-zero-bin pixels: rank inside a warp step by __match_any_sync
// Algorithmic bytes: 8n (mag) read + 4c (seed list) written; this version also writes and re-reads
// the 2n-byte bin plane.
#include "lsdb_common.cuh"

#define ORDER_WARPS 32
#define ORDER_BINS 1025

__global__ void __launch_bounds__(1024) lsdb_order_kernel(const LsdbImg* __restrict__ imgs, LsdbImgDyn* __restrict__ dyn,
                                                          const LsdbLsdConst* __restrict__ kc, const double* __restrict__ mag,
                                                          unsigned short* __restrict__ bins, unsigned int* __restrict__ cells) {
    extern __shared__ unsigned int cnt[];  // [ORDER_WARPS][ORDER_BINS] + 32 scan slots
    unsigned int* warpSum = cnt + ORDER_WARPS * ORDER_BINS;
    const LsdbImg im = imgs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const double maxGrad = __longlong_as_double((long long)dyn[blockIdx.x].maxGradBits);
    if (!(maxGrad > 0.0)) {  // blank map: the reference walks an uninitialised list here (UB) — we emit no seeds
        if (tid == 0) dyn[blockIdx.x].nCells = 0;
        return;
    }
    const int pseBin = kc->pseBin;
    const double zoom = 1.0 * pseBin / maxGrad;  // :179
    for (int i = tid; i < ORDER_WARPS * ORDER_BINS; i += 1024) cnt[i] = 0;
    __syncthreads();

    const int n = im.n;
    const int seg = ((n + ORDER_WARPS - 1) / ORDER_WARPS + 31) & ~31;
    const int p0 = w * seg, p1 = min(p0 + seg, n);
    const double* m = mag + im.nOff;
    unsigned short* bp = bins + im.nOff;
    unsigned int* myCnt = cnt + w * ORDER_BINS;

    for (int base = p0; base < p1; base += 128) {
        double v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int p = base + 32 * k + lane;
            v[k] = p < p1 ? m[p] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int p = base + 32 * k + lane;
            int t = lsdb_x86_d2i(floor(v[k] * zoom));  // :182-184
            if (t > pseBin) t = pseBin;
            t &= 0xffff;                                // pseIdx is CV_16UC1 (:178,187)
            if (p < p1) bp[p] = (unsigned short)t;
            unsigned int grp = __match_any_sync(0xffffffffu, t);
            if (t != 0 && p < p1 && lane == __ffs(grp) - 1) myCnt[t] += __popc(grp);
            __syncwarp();
        }
    }
    __syncthreads();

    // exclusive scan over bins in descending order; thread tid <-> bin 1024 - tid
    {
        const int b = 1024 - tid;
        unsigned int total = 0;
        for (int k = 0; k < ORDER_WARPS; k++) total += cnt[k * ORDER_BINS + b];
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
        for (int k = 0; k < ORDER_WARPS; k++) {
            unsigned int c = cnt[k * ORDER_BINS + b];
            cnt[k * ORDER_BINS + b] = running;
            running += c;
        }
    }
    __syncthreads();

    unsigned int* out = cells + im.nOff;
    for (int base = p0; base < p1; base += 128) {
        int t4[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int p = base + 32 * k + lane;
            t4[k] = p < p1 ? (int)bp[p] : 0;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int p = base + 32 * k + lane;
            int t = t4[k];
            unsigned int grp = __match_any_sync(0xffffffffu, t);
            unsigned int pos = 0;
            if (t != 0) pos = myCnt[t] + __popc(grp & ((1u << lane) - 1u));
            __syncwarp();
            if (t != 0) {
                out[pos] = (unsigned int)p;
                if (lane == __ffs(grp) - 1) myCnt[t] += __popc(grp);
            }
            __syncwarp();
        }
    }
}

void lsdb_launch_order(cudaStream_t s, int nImgs, const LsdbImg* imgs, LsdbImgDyn* dyn, const LsdbLsdConst* kc,
                       const double* mag, unsigned short* bins, unsigned int* cells) {
    const int smem = (ORDER_WARPS * ORDER_BINS + 32) * sizeof(unsigned int);
    cudaFuncSetAttribute(lsdb_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per device, cheap
    if (nImgs > 0) lsdb_order_kernel<<<nImgs, 1024, smem, s>>>(imgs, dyn, kc, mag, bins, cells);
}
