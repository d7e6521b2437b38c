// createMapCache on the device: the truncated brush-fire distance map that the association scoring gathers from.
//
// Replaces mylsd::createMapCache (reference LSD/myLSD.cpp:11-127).  The reference runs ONE FIFO queue over the whole
// map: every occupied cell (value 1, raster order) is a source; a dequeued entry (src, cur) claims its four
// neighbours in the order up, left, down, right (:48,:67,:86,:105) if they are still unclaimed and if
// dist(cur, src) <= cell_radius, and the claimed cell stores dist(cur, src) * res — the distance of its PARENT, not
// its own (:54,:73,:92,:111).  Which source reaches a cell first is decided by queue order, so the result depends on
// that order and it is reproduced exactly:
//   * a FIFO queue visits entries level by level (level = number of steps from the source);
//   * inside a level a cell goes to the EARLIEST queue entry that may claim it -> atomicMin over the key
//     (queue position of the parent) * 4 + direction;
//   * the next level's queue order is the order of the successful claims = ascending key -> an ordered (scan-based)
//     compaction of the winning claims.
// One propose / flag+reduce / scan / scatter round per level (<= ~1.5 * cell_radius levels); the host reads back the
// size of the next level.  Not on the per-frame path: it runs once per map.
#include "../../include/lsdb200.h"
#include "lsdb_common.cuh"
#include <stdio.h>

#define MC_BLOCK 256
#define MC_ITEMS 4          // items per thread in the compaction kernels
#define MC_TILE (MC_BLOCK * MC_ITEMS)

__device__ __forceinline__ unsigned int mc_pack(int i, int j) { return ((unsigned int)i << 16) | (unsigned int)j; }
__device__ __forceinline__ int mc_i(unsigned int v) { return (int)(v >> 16); }
__device__ __forceinline__ int mc_j(unsigned int v) { return (int)(v & 0xffffu); }

// mapCache = 0 / z_occ_max_dis, flag = occupied (:25-39)
__global__ void mc_init_kernel(const uint8_t* __restrict__ map, int n, double maxDist, double* __restrict__ cache,
                               uint8_t* __restrict__ flag, unsigned int* __restrict__ claim) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const bool occ = map[p] == 1;
    cache[p] = occ ? 0.0 : maxDist;
    flag[p] = occ ? 1 : 0;
    claim[p] = 0xffffffffu;
}

// block-wide exclusive scan of one value per thread; returns the thread's offset, *total = block sum
__device__ __forceinline__ unsigned int mc_block_scan(unsigned int v, unsigned int* total) {
    __shared__ unsigned int warpSum[MC_BLOCK / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warpSum[w] = incl;
    __syncthreads();
    unsigned int base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < MC_BLOCK / 32; k++) {
        const unsigned int s = warpSum[k];
        if (k < w) base += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

// sources in raster order: count per tile / scatter
__global__ void mc_src_count_kernel(const uint8_t* __restrict__ flag, int n, unsigned int* __restrict__ tileSum) {
    unsigned int c = 0;
    const int base = blockIdx.x * MC_TILE + threadIdx.x * MC_ITEMS;
#pragma unroll
    for (int k = 0; k < MC_ITEMS; k++) if (base + k < n && flag[base + k]) c++;
    unsigned int tot;
    mc_block_scan(c, &tot);
    if (threadIdx.x == 0) tileSum[blockIdx.x] = tot;
}
__global__ void mc_src_scatter_kernel(const uint8_t* __restrict__ flag, int n, int cols, const unsigned int* __restrict__ tileOff,
                                      unsigned int* __restrict__ fSrc, unsigned int* __restrict__ fCur) {
    unsigned int c = 0;
    const int base = blockIdx.x * MC_TILE + threadIdx.x * MC_ITEMS;
#pragma unroll
    for (int k = 0; k < MC_ITEMS; k++) if (base + k < n && flag[base + k]) c++;
    unsigned int tot;
    unsigned int pos = tileOff[blockIdx.x] + mc_block_scan(c, &tot);
#pragma unroll
    for (int k = 0; k < MC_ITEMS; k++)
        if (base + k < n && flag[base + k]) {
            const int p = base + k;
            const unsigned int v = mc_pack(p / cols, p % cols);
            fSrc[pos] = v; fCur[pos] = v;
            pos++;
        }
}

// exclusive scan of the tile sums in place (single block), total -> *count
__global__ void mc_scan_tiles_kernel(unsigned int* __restrict__ tileSum, int nTiles, unsigned int* __restrict__ count) {
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nTiles; base += MC_BLOCK) {
        const int t = base + threadIdx.x;
        const unsigned int v = t < nTiles ? tileSum[t] : 0u;
        unsigned int tot;
        const unsigned int off = mc_block_scan(v, &tot);
        const unsigned int c0 = carry;
        if (t < nTiles) tileSum[t] = c0 + off;
        __syncthreads();
        if (threadIdx.x == 0) carry = c0 + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry;
}

__device__ __forceinline__ bool mc_neighbour(int ci, int cj, int dir, int rows, int cols, int* ni, int* nj) {
    // order of the reference: up (:48), left (:67), down (:86), right (:105)
    if (dir == 0) { if (ci < 1) return false; *ni = ci - 1; *nj = cj; }
    else if (dir == 1) { if (cj < 1) return false; *ni = ci; *nj = cj - 1; }
    else if (dir == 2) { if (ci >= rows - 1) return false; *ni = ci + 1; *nj = cj; }
    else { if (cj >= cols - 1) return false; *ni = ci; *nj = cj + 1; }
    return true;
}
__device__ __forceinline__ double mc_dist(unsigned int src, unsigned int cur) {
    const double di = abs(mc_i(cur) - mc_i(src)), dj = abs(mc_j(cur) - mc_j(src));
    return sqrt(di * di + dj * dj);
}

// every entry of the level proposes itself to its unclaimed neighbours; the smallest key wins
__global__ void mc_propose_kernel(const unsigned int* __restrict__ fSrc, const unsigned int* __restrict__ fCur, int nF, int rows,
                                  int cols, int cellRadius, const uint8_t* __restrict__ flag, unsigned int* __restrict__ claim) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nF) return;
    const unsigned int src = fSrc[k], cur = fCur[k];
    if (!(mc_dist(src, cur) <= cellRadius)) return;
    const int ci = mc_i(cur), cj = mc_j(cur);
#pragma unroll
    for (int d = 0; d < 4; d++) {
        int ni, nj;
        if (!mc_neighbour(ci, cj, d, rows, cols, &ni, &nj)) continue;
        const size_t q = (size_t)ni * cols + nj;
        if (flag[q] == 0) atomicMin(&claim[q], (unsigned int)k * 4u + (unsigned int)d);
    }
}

__device__ __forceinline__ bool mc_won(const unsigned int* fSrc, const unsigned int* fCur, int nF, int key, int rows, int cols,
                                       int cellRadius, const uint8_t* flag, const unsigned int* claim, size_t* q, unsigned int* src, double* dist) {
    const int k = key >> 2, d = key & 3;
    if (k >= nF) return false;
    const unsigned int s = fSrc[k], c = fCur[k];
    const double dd = mc_dist(s, c);
    if (!(dd <= cellRadius)) return false;
    int ni, nj;
    if (!mc_neighbour(mc_i(c), mc_j(c), d, rows, cols, &ni, &nj)) return false;
    const size_t p = (size_t)ni * cols + nj;
    if (flag[p] != 0 || claim[p] != (unsigned int)key) return false;
    *q = p; *src = s; *dist = dd;
    return true;
}

__global__ void mc_win_count_kernel(const unsigned int* __restrict__ fSrc, const unsigned int* __restrict__ fCur, int nF, int rows,
                                    int cols, int cellRadius, const uint8_t* __restrict__ flag, const unsigned int* __restrict__ claim,
                                    unsigned int* __restrict__ tileSum) {
    unsigned int c = 0;
    const int base = blockIdx.x * MC_TILE + threadIdx.x * MC_ITEMS;
    size_t q; unsigned int s; double dd;
#pragma unroll
    for (int i = 0; i < MC_ITEMS; i++) if (mc_won(fSrc, fCur, nF, base + i, rows, cols, cellRadius, flag, claim, &q, &s, &dd)) c++;
    unsigned int tot;
    mc_block_scan(c, &tot);
    if (threadIdx.x == 0) tileSum[blockIdx.x] = tot;
}

// winners in key order: store the parent's distance (:54), and queue (src, cell) for the next level.  The flags are set
// by a separate kernel afterwards: mc_won must see the flags of the level's start.
__global__ void mc_win_scatter_kernel(const unsigned int* __restrict__ fSrc, const unsigned int* __restrict__ fCur, int nF, int rows,
                                      int cols, int cellRadius, const uint8_t* __restrict__ flag, const unsigned int* __restrict__ claim,
                                      const unsigned int* __restrict__ tileOff, double res, double* __restrict__ cache,
                                      unsigned int* __restrict__ nSrc, unsigned int* __restrict__ nCur) {
    const int base = blockIdx.x * MC_TILE + threadIdx.x * MC_ITEMS;
    size_t q[MC_ITEMS]; unsigned int s[MC_ITEMS]; double dd[MC_ITEMS]; bool w[MC_ITEMS];
    unsigned int c = 0;
#pragma unroll
    for (int i = 0; i < MC_ITEMS; i++) {
        w[i] = mc_won(fSrc, fCur, nF, base + i, rows, cols, cellRadius, flag, claim, &q[i], &s[i], &dd[i]);
        if (w[i]) c++;
    }
    unsigned int tot;
    unsigned int pos = tileOff[blockIdx.x] + mc_block_scan(c, &tot);
#pragma unroll
    for (int i = 0; i < MC_ITEMS; i++)
        if (w[i]) {
            cache[q[i]] = dd[i] * res;
            nSrc[pos] = s[i];
            nCur[pos] = mc_pack((int)(q[i] / cols), (int)(q[i] % cols));
            pos++;
        }
}
__global__ void mc_mark_kernel(const unsigned int* __restrict__ nCur, int n, int cols, uint8_t* __restrict__ flag) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) flag[(size_t)mc_i(nCur[k]) * cols + mc_j(nCur[k])] = 1;
}

struct lsdb_ctx_view { int device; cudaStream_t stream; };   // leading members of lsdb_ctx (api.cu)

extern "C" int lsdb_map_cache_device(int device, void* streamV, const uint8_t* map, int cols, int rows, double res, double maxDist, double unreached,
                                     double* out, char* err, int errLen) {
    cudaStream_t s = (cudaStream_t)streamV;
#define MCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(err, errLen, "lsdb_map_cache: %s (line %d)", cudaGetErrorString(e_), __LINE__); goto fail; } } while (0)
    const size_t n = (size_t)cols * rows;
    const int cellRadius = (int)floor(maxDist / res);   // :13
    uint8_t *mapD = 0, *flag = 0;
    double* cache = 0;
    unsigned int *claim = 0, *f[4] = {0, 0, 0, 0}, *tileSum = 0, *count = 0;
    unsigned int nF = 0;
    int rc = LSDB_ERR_CUDA;
    const int maxTiles = (int)((4 * n + MC_TILE - 1) / MC_TILE) + 1;
    MCK(cudaSetDevice(device));
    MCK(cudaMalloc(&mapD, n)); MCK(cudaMalloc(&flag, n)); MCK(cudaMalloc(&cache, n * 8)); MCK(cudaMalloc(&claim, n * 4));
    for (int k = 0; k < 4; k++) MCK(cudaMalloc(&f[k], n * 4));
    MCK(cudaMalloc(&tileSum, (size_t)maxTiles * 4)); MCK(cudaMalloc(&count, 4));
    MCK(cudaMemcpyAsync(mapD, map, n, cudaMemcpyHostToDevice, s));
    mc_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mapD, (int)n, unreached, cache, flag, claim);
    {   // level 0: the occupied cells in raster order (:21-38)
        const int nt = (int)((n + MC_TILE - 1) / MC_TILE);
        mc_src_count_kernel<<<nt, MC_BLOCK, 0, s>>>(flag, (int)n, tileSum);
        mc_scan_tiles_kernel<<<1, MC_BLOCK, 0, s>>>(tileSum, nt, count);
        mc_src_scatter_kernel<<<nt, MC_BLOCK, 0, s>>>(flag, (int)n, cols, tileSum, f[0], f[1]);
        MCK(cudaMemcpyAsync(&nF, count, 4, cudaMemcpyDeviceToHost, s));
        MCK(cudaStreamSynchronize(s));
    }
    for (int cur = 0; nF > 0; cur ^= 1) {
        unsigned int *fSrc = f[2 * cur], *fCur = f[2 * cur + 1], *nSrc = f[2 * (cur ^ 1)], *nCur = f[2 * (cur ^ 1) + 1];
        const int nt = (int)(((size_t)4 * nF + MC_TILE - 1) / MC_TILE);
        mc_propose_kernel<<<(nF + 255) / 256, 256, 0, s>>>(fSrc, fCur, (int)nF, rows, cols, cellRadius, flag, claim);
        mc_win_count_kernel<<<nt, MC_BLOCK, 0, s>>>(fSrc, fCur, (int)nF, rows, cols, cellRadius, flag, claim, tileSum);
        mc_scan_tiles_kernel<<<1, MC_BLOCK, 0, s>>>(tileSum, nt, count);
        mc_win_scatter_kernel<<<nt, MC_BLOCK, 0, s>>>(fSrc, fCur, (int)nF, rows, cols, cellRadius, flag, claim, tileSum, res, cache, nSrc, nCur);
        unsigned int nNext = 0;
        MCK(cudaMemcpyAsync(&nNext, count, 4, cudaMemcpyDeviceToHost, s));
        MCK(cudaStreamSynchronize(s));
        if (nNext) mc_mark_kernel<<<(nNext + 255) / 256, 256, 0, s>>>(nCur, (int)nNext, cols, flag);
        nF = nNext;
    }
    MCK(cudaMemcpyAsync(out, cache, n * 8, cudaMemcpyDeviceToHost, s));
    MCK(cudaStreamSynchronize(s));
    MCK(cudaGetLastError());
    rc = LSDB_OK;
fail:
    cudaFree(mapD); cudaFree(flag); cudaFree(cache); cudaFree(claim);
    for (int k = 0; k < 4; k++) cudaFree(f[k]);
    cudaFree(tileSum); cudaFree(count);
    return rc;
#undef MCK
}
