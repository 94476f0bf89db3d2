// distinctive.cu -- batched MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:444-535).
//
// For every map point: all-pairs Hamming distances among the descriptors of its observations (:489-499, Distances[i][i] = 0),
// per row the median vDists[0.5 * (N - 1)] of the sorted row (:507-509), and the row with the smallest median, first row
// winning ties (`median < BestMedian`, :510-514).  The observation lists arrive as a CSR over one descriptor array.
// One CTA per map point: the descriptors are staged in shared memory; a warp owns a row, its lanes take the columns, the
// distances go into the warp's 257-bin histogram and a warp scan finds the k-th smallest; the rows are then min-reduced on
// the key median << 16 | row.  Points with more observations than the staging holds are walked from global memory.
#include "internal.cuh"

namespace {

constexpr int DD_THREADS = 256, DD_WARPS = DD_THREADS / 32, DD_STAGE = 1024; // descriptors staged per CTA (32 KB)

__global__ void __launch_bounds__(DD_THREADS)
distinctive_kernel(int n_mp, const int32_t *__restrict__ offsets, const uint4 *__restrict__ desc, int32_t *__restrict__ best_idx,
                   int32_t *__restrict__ best_median, unsigned long long *__restrict__ counters)
{
    __shared__ uint4 sD[2 * DD_STAGE];
    __shared__ int hist[DD_WARPS][264];
    __shared__ unsigned s_best;
    const int mp = blockIdx.x;
    if (mp >= n_mp) return;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int s = offsets[mp], N = offsets[mp + 1] - s;
    if (N <= 0) { // :455-456, :481-482 return without touching the descriptor
        if (t == 0) { best_idx[mp] = -1; best_median[mp] = -1; }
        return;
    }
    const uint4 *g = desc + 2 * (size_t)s;
    const bool staged = N <= DD_STAGE;
    if (staged)
        for (int i = t; i < 2 * N; i += DD_THREADS) sD[i] = g[i];
    if (t == 0) s_best = 0xFFFFFFFFu;
    __syncthreads();
    const uint4 *D = staged ? sD : g;
    const int k = (int)(0.5 * (double)(N - 1)); // index into the sorted row (:509)
    unsigned mine = 0xFFFFFFFFu;
    for (int i = warp; i < N; i += DD_WARPS) {
        for (int b = lane; b < 264; b += 32) hist[warp][b] = 0;
        __syncwarp();
        const uint4 a0 = D[2 * i], a1 = D[2 * i + 1];
        for (int j = lane; j < N; j += 32) {
            const int d = (j == i) ? 0 : ham256(a0, a1, D[2 * j], D[2 * j + 1]); // :492
            atomicAdd(&hist[warp][d], 1);
        }
        __syncwarp();
        // smallest v with #{d <= v} >= k + 1: warp scan over the 257 bins, 9 bins per lane
        int c[9], sum = 0;
        for (int u = 0; u < 9; u++) {
            const int b = lane * 9 + u;
            c[u] = b < 257 ? hist[warp][b] : 0;
            sum += c[u];
        }
        int incl = sum;
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += up;
        }
        int run = incl - sum, med = 0x7FFFFFFF;
        for (int u = 0; u < 9; u++) {
            run += c[u];
            if (run >= k + 1 && med == 0x7FFFFFFF) med = lane * 9 + u;
        }
        med = __reduce_min_sync(FULL_MASK, med);
        mine = min(mine, ((unsigned)med << 16) | (unsigned)min(i, 0xFFFF)); // rows ascend per warp: first row wins ties
        __syncwarp();
    }
    if (lane == 0 && mine != 0xFFFFFFFFu) atomicMin(&s_best, mine);
    __syncthreads();
    if (t == 0) {
        // rows beyond 65535 cannot be encoded in the key; such lists do not occur (a map point has one observation per key frame)
        best_idx[mp] = (int)(s_best & 0xFFFFu);
        best_median[mp] = (int)(s_best >> 16);
        atomicAdd(&counters[0], (unsigned long long)N * (unsigned long long)(N - 1) / 2ull); // DescriptorDistance calls (:491)
    }
}

} // namespace

extern "C" int orbgpu_compute_distinctive_descriptors(orbgpu_ctx *ctx, int32_t n_mp, const int32_t *offsets, const uint8_t *desc,
                                                      int32_t *best_idx, int32_t *best_median)
{
    ARG_TRY(ctx && n_mp >= 0 && (n_mp == 0 || (offsets && best_idx)));
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    if (n_mp == 0) return ORBGPU_OK;
    const int64_t total = offsets[n_mp];
    ARG_TRY(total >= 0 && (total == 0 || desc));
    for (int i = 0; i < n_mp; i++) ARG_TRY(offsets[i] <= offsets[i + 1] && offsets[i + 1] - offsets[i] <= 65535);
    const size_t ob = align256((size_t)(n_mp + 1) * 4), db = align256((size_t)(total > 0 ? total : 1) * 32), rb = align256((size_t)n_mp * 4);
    rc = arena_reserve(ctx, ob + db + 2 * rb + 256);
    if (rc) return rc;
    int32_t *d_off = (int32_t *)arena_take(ctx, ob);
    uint4 *d_desc = (uint4 *)arena_take(ctx, db);
    int32_t *d_bi = (int32_t *)arena_take(ctx, rb), *d_bm = (int32_t *)arena_take(ctx, rb);
    if (!d_off || !d_desc || !d_bi || !d_bm) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    CU_TRY(cudaMemcpyAsync(d_off, offsets, (size_t)(n_mp + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (total) CU_TRY(cudaMemcpyAsync(d_desc, desc, (size_t)total * 32, cudaMemcpyHostToDevice, ctx->stream));
    distinctive_kernel<<<n_mp, DD_THREADS, 0, ctx->stream>>>(n_mp, d_off, d_desc, d_bi, d_bm, ctx->d_counters);
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    const OutPiece out[2] = {{best_idx, d_bi, (size_t)n_mp * 4}, {best_median, d_bm, (size_t)n_mp * 4}};
    return ctx_download(ctx, out, 2);
}
