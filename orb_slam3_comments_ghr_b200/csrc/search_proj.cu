// search_proj.cu -- ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints,
// thFarPoints) (ORBmatcher.cc:44-242, Nleft == -1 path) + RadiusByViewingCos (:245-252), config C2.
//
// Same two-phase shape as search_init.cu: the window candidates of every map point and their
// Hamming distances are computed in parallel (one warp per map point); an ordered pass then
// replays the reference's sequential rule "skip a keypoint whose current map point has
// Observations() > 0" (:102-104) -- a later map point sees the assignments of earlier ones.
#include <cstring>

#include "internal.cuh"

namespace {

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;

struct MapPointsView {
    int n;
    const uint4 *desc;
    const float2 *proj_xy;
    const float *proj_xr;
    const int32_t *scale_level;
    const float *view_cos;
    const float *depth;
    const uint8_t *in_view;
    const uint8_t *bad;
    const int32_t *n_obs;
};

__global__ void proj_candidates_kernel(FrameView f, MapPointsView mp, float th, int far_points, float th_far,
                                       uint32_t *__restrict__ lists, int stride, int32_t *__restrict__ counts,
                                       unsigned long long *__restrict__ counters)
{
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= mp.n) return;
    const int lane = lane_id();
    int cnt = 0;
    const bool active = mp.in_view[m] && !(far_points && mp.depth[m] > th_far) && !mp.bad[m]; // :55-62
    if (active) {
        const int level = mp.scale_level[m];
        float r = ((double)mp.view_cos[m] > 0.998) ? 2.5f : 4.0f; // :245-252
        if (th != 1.0f) r = __fmul_rn(r, th);                      // :49,:74
        const float radius = __fmul_rn(r, f.scale_factors[level]); // :80
        const float2 p = mp.proj_xy[m];
        const uint4 qa = mp.desc[2 * m], qb = mp.desc[2 * m + 1];
        const float xr = (f.u_right && mp.proj_xr) ? mp.proj_xr[m] : 0.f;
        uint32_t *out = lists + (size_t)m * stride;
        cnt = window_scan_if(
            f, p.x, p.y, radius, level - 1, level,
            [&](const int4 &it) {
                if (f.u_right) { // :107-117 stereo gate (static per candidate)
                    const float ur = f.u_right[it.w];
                    if (ur > 0.f) {
                        const float er = fabsf(__fsub_rn(xr, ur));
                        if (er > radius) return false;
                    }
                }
                return true;
            },
            [&](bool ok, int pos, int slot, int4 it) {
                if (ok) {
                    const int dist = ham256(qa, qb, f.desc_sorted[2 * slot], f.desc_sorted[2 * slot + 1]);
                    out[pos] = ((uint32_t)dist << 20) | (uint32_t)it.w;
                }
            });
    }
    if (lane == 0) counts[m] = cnt;
}

// one warp replays the map points in vpMapPoints order
__global__ void __launch_bounds__(256)
proj_resolve_kernel(FrameView f, MapPointsView mp, const uint32_t *__restrict__ lists, int stride,
                    const int32_t *__restrict__ counts, float nnratio, int32_t *__restrict__ cur_obs,
                    int32_t *__restrict__ kp_mp, int32_t *__restrict__ nmatches_out, unsigned long long *__restrict__ counters)
{
    const int lane = threadIdx.x & 31;
    if (threadIdx.x >= 32) return;
    int nmatches = 0;
    unsigned long long ncmp = 0;
    for (int m = 0; m < mp.n; m++) {
        const int cnt = counts[m];
        if (cnt == 0) continue; // filtered (:55-62) or empty window (:84)
        const uint32_t *lst = lists + (size_t)m * stride;
        uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
        int nvalid = 0;
        for (int base = 0; base < cnt; base += 32) {
            const int p = base + lane;
            if (p < cnt) {
                const uint32_t e = lst[p];
                const int idx = (int)(e & 0xFFFFF);
                if (!(cur_obs[idx] > 0)) { // :102-104
                    top2_push(b1, b2, (e & 0xFFF00000u) | (uint32_t)p);
                    nvalid++;
                }
            }
        }
        ncmp += nvalid; // DescriptorDistance is only called for candidates that pass the skip rule
        uint32_t m1, m2;
        warp_top2(b1, b2, m1, m2);
        if (m1 != KEY_NONE && lane == 0) {
            const int bestDist = (int)(m1 >> 20);
            if (bestDist <= ORBGPU_TH_HIGH) { // :147
                const int bestIdx = (int)(lst[m1 & 0xFFFFF] & 0xFFFFF);
                const int bestLevel = f.octave[bestIdx];
                int bestDist2 = 256, bestLevel2 = -1;
                if (m2 != KEY_NONE) {
                    bestDist2 = (int)(m2 >> 20);
                    bestLevel2 = f.octave[lst[m2 & 0xFFFFF] & 0xFFFFF];
                }
                const float lim = __fmul_rn(nnratio, (float)bestDist2);
                const bool reject = (bestLevel == bestLevel2) && ((float)bestDist > lim);  // :151
                if (!reject && (bestLevel != bestLevel2 || (float)bestDist <= lim)) {       // :154
                    kp_mp[bestIdx] = m;                                                    // :156
                    cur_obs[bestIdx] = mp.n_obs[m];
                    nmatches++;
                }
            }
        }
        __syncwarp();
    }
    for (int off = 16; off; off >>= 1) ncmp += __shfl_xor_sync(FULL_MASK, ncmp, off);
    if (lane == 0) {
        *nmatches_out = nmatches;
        counters[0] = ncmp;
    }
}

} // namespace

extern "C" int orbgpu_search_by_projection_local(orbgpu_ctx *ctx, const orbgpu_frame *f, const orbgpu_mappoints_host *mps,
                                                 float th, int32_t far_points, float th_far_points, float nnratio,
                                                 const int32_t *kp_prior_obs, int32_t *kp_mp, int32_t *nmatches)
{
    ARG_TRY(ctx && f && mps && nmatches);
    ARG_TRY(f->n == 0 || (kp_prior_obs && kp_mp));
    ARG_TRY(mps->n >= 0 && (mps->n == 0 || (mps->desc && mps->proj_xy && mps->scale_level && mps->view_cos && mps->depth &&
                                            mps->in_view && mps->bad && mps->n_obs)));
    ARG_TRY(f->n < (1 << 20));
    ARG_TRY(!(f->u_right && !mps->proj_xr));
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    *nmatches = 0;
    const int n = f->n, M = mps->n;
    if (M == 0 || n == 0) return ORBGPU_OK;
    for (int i = 0; i < M; i++) // the window radius is scaled by mvScaleFactors[level]
        ARG_TRY(!mps->in_view[i] || (mps->scale_level[i] >= 0 && mps->scale_level[i] < f->n_levels));
    // pack the map-point arrays into pinned staging -> one H2D copy
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o_desc = take((size_t)M * 32), o_xy = take((size_t)M * 8), o_xr = take((size_t)M * 4), o_lvl = take((size_t)M * 4),
                 o_cos = take((size_t)M * 4), o_dep = take((size_t)M * 4), o_nobs = take((size_t)M * 4), o_inv = take(M),
                 o_bad = take(M), o_prior = take((size_t)n * 4), o_kpmp = take((size_t)n * 4);
    const size_t up_bytes = off;
    const int stride = n;
    rc = stage_reserve(ctx, up_bytes);
    if (rc) return rc;
    rc = arena_reserve(ctx, up_bytes + align256((size_t)M * stride * 4) + align256((size_t)M * 4) + 512);
    if (rc) return rc;
    char *H = ctx->h_stage;
    memcpy(H + o_desc, mps->desc, (size_t)M * 32);
    memcpy(H + o_xy, mps->proj_xy, (size_t)M * 8);
    if (mps->proj_xr) memcpy(H + o_xr, mps->proj_xr, (size_t)M * 4);
    memcpy(H + o_lvl, mps->scale_level, (size_t)M * 4);
    memcpy(H + o_cos, mps->view_cos, (size_t)M * 4);
    memcpy(H + o_dep, mps->depth, (size_t)M * 4);
    memcpy(H + o_nobs, mps->n_obs, (size_t)M * 4);
    memcpy(H + o_inv, mps->in_view, M);
    memcpy(H + o_bad, mps->bad, M);
    memcpy(H + o_prior, kp_prior_obs, (size_t)n * 4);
    memcpy(H + o_kpmp, kp_mp, (size_t)n * 4);
    char *D = (char *)arena_take(ctx, up_bytes);
    uint32_t *lists = (uint32_t *)arena_take(ctx, (size_t)M * stride * 4);
    int32_t *counts = (int32_t *)arena_take(ctx, (size_t)M * 4);
    int32_t *d_nm = (int32_t *)arena_take(ctx, 256);
    if (!D || !lists || !counts || !d_nm) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    CU_TRY(cudaMemcpyAsync(D, H, up_bytes, cudaMemcpyHostToDevice, ctx->stream));
    MapPointsView mv;
    mv.n = M;
    mv.desc = (const uint4 *)(D + o_desc); mv.proj_xy = (const float2 *)(D + o_xy);
    mv.proj_xr = mps->proj_xr ? (const float *)(D + o_xr) : nullptr;
    mv.scale_level = (const int32_t *)(D + o_lvl); mv.view_cos = (const float *)(D + o_cos); mv.depth = (const float *)(D + o_dep);
    mv.in_view = (const uint8_t *)(D + o_inv); mv.bad = (const uint8_t *)(D + o_bad); mv.n_obs = (const int32_t *)(D + o_nobs);
    int32_t *cur_obs = (int32_t *)(D + o_prior), *d_kpmp = (int32_t *)(D + o_kpmp);
    const FrameView v = frame_view(f);
    proj_candidates_kernel<<<(M * 32 + 255) / 256, 256, 0, ctx->stream>>>(v, mv, th, far_points, th_far_points, lists, stride, counts,
                                                                         ctx->d_counters);
    proj_resolve_kernel<<<1, 32, 0, ctx->stream>>>(v, mv, lists, stride, counts, nnratio, cur_obs, d_kpmp, d_nm, ctx->d_counters);
    ctx->launches += 2;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(kp_mp, d_kpmp, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaMemcpyAsync(nmatches, d_nm, 4, cudaMemcpyDeviceToHost, ctx->stream));
    return ctx_fetch_comparisons(ctx);
}
