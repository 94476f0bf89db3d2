// stereo_coarse.cu -- coarse stage of Frame::ComputeStereoMatches (src/Frame.cc:1117-1250): for every
// left keypoint the best right keypoint of its image row band by descriptor distance.
//   row index      :1143-1156  every right keypoint is listed in the rows [floor(y - r), ceil(y + r)], r = 2 * scaleFactor[octave]
//   candidates     :1177-1181  the list of row (int)vL of the left keypoint, ascending right index
//   filters        :1194-1200  octave within +-1 of the left keypoint's, uR in [uL - mbf/mb, uL]
//   selection      :1183, :1203-1209  bestDist starts at TH_HIGH, strict <, first candidate wins ties
//   acceptance     :1214  bestDist < (TH_HIGH + TH_LOW) / 2
// The sub-pixel SAD refinement that follows (:1216-1290) reads the image pyramids and stays with the caller.
// The row index becomes a device CSR (one thread per row, two passes over the right keypoints keep each list in ascending
// right index); then one warp per left keypoint scans its row list with coalesced loads and a (dist, position) warp argmin.
#include <cstring>

#include "internal.cuh"

namespace {

__global__ void stereo_rows_kernel(int n_right, const float2 *__restrict__ xy_r, const int32_t *__restrict__ oct_r,
                                   const float *__restrict__ scale_factors, int n_rows, int32_t *__restrict__ row_start,
                                   int32_t *__restrict__ row_items, int fill)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    int cnt = 0;
    const int base = fill ? row_start[row] : 0;
    for (int i = 0; i < n_right; i++) {
        const float y = xy_r[i].y;
        const float r = __fmul_rn(2.0f, scale_factors[oct_r[i]]);
        // rows outside the image are undefined behaviour in the reference (vRowIndices[yi] out of range); they are clamped away
        const int maxr = (int)ceilf(__fadd_rn(y, r)), minr = (int)floorf(__fsub_rn(y, r));
        if (row >= minr && row <= maxr) {
            if (fill) row_items[base + cnt] = i;
            cnt++;
        }
    }
    if (!fill) row_start[row] = cnt;
}

__global__ void stereo_scan_kernel(int n_rows, int32_t *__restrict__ row_start)
{
    // exclusive scan of n_rows counts by one warp (n_rows is an image height)
    const int lane = threadIdx.x;
    int carry = 0;
    for (int b = 0; b < n_rows; b += 32) {
        const int i = b + lane;
        const int v = i < n_rows ? row_start[i] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += u;
        }
        if (i < n_rows) row_start[i] = carry + incl - v;
        carry += __shfl_sync(FULL_MASK, incl, 31);
    }
    if (lane == 0) row_start[n_rows] = carry;
}

__global__ void stereo_match_kernel(int n_left, const uint4 *__restrict__ desc_l, const float2 *__restrict__ xy_l,
                                    const int32_t *__restrict__ oct_l, const uint4 *__restrict__ desc_r,
                                    const float2 *__restrict__ xy_r, const int32_t *__restrict__ oct_r, int n_rows,
                                    const int32_t *__restrict__ row_start, const int32_t *__restrict__ row_items, float max_d,
                                    int32_t *__restrict__ best_idx, int32_t *__restrict__ best_dist, unsigned long long *__restrict__ counters)
{
    const int iL = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (iL >= n_left) return;
    const int lane = lane_id();
    const float2 p = xy_l[iL];
    const int levelL = oct_l[iL];
    int bi = -1, bd = ORBGPU_TH_HIGH;
    const int row = (int)p.y; // vRowIndices[vL]: float -> index truncation (:1177)
    const float minU = __fsub_rn(p.x, max_d), maxU = p.x; // minD = 0
    unsigned ncmp = 0;
    if (p.y >= 0.f && row < n_rows && !(maxU < 0.f)) {
        const int s = row_start[row], e = row_start[row + 1];
        const uint4 a0 = desc_l[2 * iL], a1 = desc_l[2 * iL + 1];
        uint32_t best = 0xFFFFFFFFu;
        for (int c = s + lane; c < e; c += 32) {
            const int iR = row_items[c];
            const int o = oct_r[iR];
            if (o < levelL - 1 || o > levelL + 1) continue;
            const float uR = xy_r[iR].x;
            if (uR >= minU && uR <= maxU) {
                const int d = ham256(a0, a1, desc_r[2 * iR], desc_r[2 * iR + 1]);
                ncmp++;
                best = min(best, ((uint32_t)d << 20) | (uint32_t)(c - s)); // strict <: the first candidate wins ties
            }
        }
        best = __reduce_min_sync(FULL_MASK, best);
        if (best != 0xFFFFFFFFu && (int)(best >> 20) < bd) {
            bd = (int)(best >> 20);
            bi = row_items[s + (int)(best & 0xFFFFF)];
        }
    }
    ncmp = __reduce_add_sync(FULL_MASK, ncmp);
    if (lane == 0) {
        const int th = (ORBGPU_TH_HIGH + ORBGPU_TH_LOW) / 2;
        best_idx[iL] = bd < th ? bi : -1;
        best_dist[iL] = bd;
        if (ncmp) atomicAdd(&counters[0], (unsigned long long)ncmp);
    }
}

} // namespace

extern "C" int orbgpu_stereo_coarse_match(orbgpu_ctx *ctx, int32_t n_left, const uint8_t *desc_l, const float *kp_xy_l,
                                          const int32_t *octave_l, int32_t n_right, const uint8_t *desc_r, const float *kp_xy_r,
                                          const int32_t *octave_r, const float *scale_factors, int32_t n_levels, int32_t n_rows,
                                          float mb, float mbf, int32_t *best_idx_r, int32_t *best_dist)
{
    ARG_TRY(ctx && n_left >= 0 && n_right >= 0 && n_rows > 0 && n_levels > 0 && n_levels <= 64 && scale_factors);
    ARG_TRY(n_left == 0 || (desc_l && kp_xy_l && octave_l && best_idx_r && best_dist));
    ARG_TRY(n_right == 0 || (desc_r && kp_xy_r && octave_r));
    for (int i = 0; i < n_right; i++) ARG_TRY(octave_r[i] >= 0 && octave_r[i] < n_levels);
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    if (n_left == 0) return ORBGPU_OK;
    if (n_right == 0) {
        for (int i = 0; i < n_left; i++) { best_idx_r[i] = -1; best_dist[i] = ORBGPU_TH_HIGH; }
        return ORBGPU_OK;
    }
    // a right keypoint is listed in at most ceil(2r)+2 rows, r = 2 * the largest scale factor
    float sf_max = 0.f;
    for (int i = 0; i < n_levels; i++) sf_max = scale_factors[i] > sf_max ? scale_factors[i] : sf_max;
    const size_t per_kp = (size_t)(4.0f * sf_max) + 4;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o_dl = take((size_t)n_left * 32), o_xl = take((size_t)n_left * 8), o_ol = take((size_t)n_left * 4),
                 o_dr = take((size_t)n_right * 32), o_xr = take((size_t)n_right * 8), o_or = take((size_t)n_right * 4), o_sf = take(64 * 4);
    const size_t up = off;
    rc = stage_reserve(ctx, up);
    if (rc) return rc;
    rc = arena_reserve(ctx, up + align256((size_t)(n_rows + 1) * 4) + align256((size_t)n_right * per_kp * 4) + 2 * align256((size_t)n_left * 4));
    if (rc) return rc;
    char *H = ctx->h_stage;
    memcpy(H + o_dl, desc_l, (size_t)n_left * 32); memcpy(H + o_xl, kp_xy_l, (size_t)n_left * 8); memcpy(H + o_ol, octave_l, (size_t)n_left * 4);
    memcpy(H + o_dr, desc_r, (size_t)n_right * 32); memcpy(H + o_xr, kp_xy_r, (size_t)n_right * 8); memcpy(H + o_or, octave_r, (size_t)n_right * 4);
    memcpy(H + o_sf, scale_factors, (size_t)n_levels * 4);
    char *D = (char *)arena_take(ctx, up);
    int32_t *row_start = (int32_t *)arena_take(ctx, (size_t)(n_rows + 1) * 4);
    int32_t *row_items = (int32_t *)arena_take(ctx, (size_t)n_right * per_kp * 4);
    int32_t *d_bi = (int32_t *)arena_take(ctx, (size_t)n_left * 4), *d_bd = (int32_t *)arena_take(ctx, (size_t)n_left * 4);
    if (!D || !row_start || !row_items || !d_bi || !d_bd) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    CU_TRY(cudaMemcpyAsync(D, H, up, cudaMemcpyHostToDevice, ctx->stream));
    const float2 *xr = (const float2 *)(D + o_xr);
    const int32_t *orr = (const int32_t *)(D + o_or);
    const float *sf = (const float *)(D + o_sf);
    stereo_rows_kernel<<<(n_rows + 127) / 128, 128, 0, ctx->stream>>>(n_right, xr, orr, sf, n_rows, row_start, row_items, 0);
    stereo_scan_kernel<<<1, 32, 0, ctx->stream>>>(n_rows, row_start);
    stereo_rows_kernel<<<(n_rows + 127) / 128, 128, 0, ctx->stream>>>(n_right, xr, orr, sf, n_rows, row_start, row_items, 1);
    const float max_d = mbf / mb; // :1160-1163 (minZ = mb, minD = 0)
    stereo_match_kernel<<<(n_left * 32 + 255) / 256, 256, 0, ctx->stream>>>(n_left, (const uint4 *)(D + o_dl), (const float2 *)(D + o_xl),
                                                                           (const int32_t *)(D + o_ol), (const uint4 *)(D + o_dr), xr, orr,
                                                                           n_rows, row_start, row_items, max_d, d_bi, d_bd, ctx->d_counters);
    ctx->launches += 4;
    CU_TRY(cudaGetLastError());
    const OutPiece out[2] = {{best_idx_r, d_bi, (size_t)n_left * 4}, {best_dist, d_bd, (size_t)n_left * 4}};
    return ctx_download(ctx, out, 2);
}
