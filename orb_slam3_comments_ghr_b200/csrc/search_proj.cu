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

// Ordered pass.  The reference walks vpMapPoints in order and skips a keypoint whose current map point has
// Observations() > 0 (:102-104), so map point m sees the assignments of the map points before it.  A keypoint k is
// therefore unavailable to m iff it was locked on entry (prior Observations() > 0) or some m' < m with
// Observations() > 0 chose it: lock[k] = min such m' (-1 when locked on entry).  The choices and the lock times
// define each other through a triangular system (m only depends on m' < m); it is solved by fixed-point iteration --
// every round re-decides all map points in parallel against the previous round's lock times; map point 0 is right
// after round 0, and by induction one more map point prefix is right after every round, so the fixed point is the
// reference's sequential result (typically reached in 2-3 rounds: windows rarely chain).
// One CTA: lock times live in shared memory, a thread owns map points m = t, t+T, ...
constexpr int RESOLVE_THREADS = 1024;
__global__ void __launch_bounds__(RESOLVE_THREADS)
proj_resolve_kernel(FrameView f, MapPointsView mp, const uint32_t *__restrict__ lists, int stride,
                    const int32_t *__restrict__ counts, float nnratio, const int32_t *__restrict__ prior_obs,
                    int32_t *__restrict__ choice, int32_t *__restrict__ kp_mp, int32_t *__restrict__ nmatches_out,
                    unsigned long long *__restrict__ counters)
{
    extern __shared__ int lock_time[]; // [f.n]
    __shared__ int s_changed, s_nmatches;
    __shared__ unsigned long long s_ncmp;
    const int t = threadIdx.x;
    for (int k = t; k < f.n; k += RESOLVE_THREADS) lock_time[k] = prior_obs[k] > 0 ? -1 : 0x7FFFFFFF;
    for (int m = t; m < mp.n; m += RESOLVE_THREADS) choice[m] = -2; // "not decided yet"
    if (t == 0) s_changed = 0;
    __syncthreads();
    for (;;) {
        int nm = 0;
        unsigned long long ncmp = 0;
        bool changed = false;
        for (int m = t; m < mp.n; m += RESOLVE_THREADS) {
            const int cnt = counts[m];
            int pick = -1;
            if (cnt > 0) { // else filtered (:55-62) or empty window (:84)
                const uint32_t *lst = lists + (size_t)m * stride;
                int bestDist = 256, bestDist2 = 256, bestIdx = -1, idx2 = -1; // :89-93
                for (int p = 0; p < cnt; p++) {
                    const uint32_t e = lst[p];
                    const int idx = (int)(e & 0xFFFFF), dist = (int)(e >> 20);
                    if (lock_time[idx] < m) continue; // :102-104
                    ncmp++;                           // DescriptorDistance is only called past the skip rule
                    if (dist < bestDist) {            // :125-141
                        bestDist2 = bestDist; idx2 = bestIdx;
                        bestDist = dist; bestIdx = idx;
                    } else if (dist < bestDist2) {
                        bestDist2 = dist; idx2 = idx;
                    }
                }
                if (bestIdx >= 0 && bestDist <= ORBGPU_TH_HIGH) { // :147
                    const int bestLevel = f.octave[bestIdx], bestLevel2 = idx2 >= 0 ? f.octave[idx2] : -1;
                    const float lim = __fmul_rn(nnratio, (float)bestDist2);
                    if (!(bestLevel == bestLevel2 && (float)bestDist > lim)) pick = bestIdx; // :151-154
                }
            }
            if (pick != choice[m]) {
                choice[m] = pick;
                changed = true;
            }
            nm += pick >= 0;
        }
        if (changed) s_changed = 1;
        __syncthreads();
        const bool again = s_changed != 0;
        __syncthreads();
        if (!again) { // fixed point: publish
            if (t == 0) { s_nmatches = 0; s_ncmp = 0; }
            __syncthreads();
            for (int o = 16; o; o >>= 1) {
                nm += __shfl_xor_sync(FULL_MASK, nm, o);
                ncmp += __shfl_xor_sync(FULL_MASK, ncmp, o);
            }
            if ((t & 31) == 0) {
                if (nm) atomicAdd(&s_nmatches, nm);
                if (ncmp) atomicAdd(&s_ncmp, ncmp);
            }
            // F.mvpMapPoints[bestIdx] = pMP (:156): the last map point that chose a keypoint keeps it
            for (int m = t; m < mp.n; m += RESOLVE_THREADS)
                if (choice[m] >= 0) kp_mp[choice[m]] = -1; // whatever the keypoint held on entry is replaced
            __syncthreads();
            for (int m = t; m < mp.n; m += RESOLVE_THREADS)
                if (choice[m] >= 0) atomicMax(&kp_mp[choice[m]], m);
            __syncthreads();
            if (t == 0) {
                *nmatches_out = s_nmatches;
                counters[0] = s_ncmp;
            }
            return;
        }
        // lock times of this round's choices
        if (t == 0) s_changed = 0;
        for (int k = t; k < f.n; k += RESOLVE_THREADS) lock_time[k] = prior_obs[k] > 0 ? -1 : 0x7FFFFFFF;
        __syncthreads();
        for (int m = t; m < mp.n; m += RESOLVE_THREADS) {
            const int c = choice[m];
            if (c >= 0 && mp.n_obs[m] > 0) atomicMin(&lock_time[c], m);
        }
        __syncthreads();
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
    rc = arena_reserve(ctx, up_bytes + align256((size_t)M * stride * 4) + 2 * align256((size_t)M * 4) + 512);
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
    int32_t *choice = (int32_t *)arena_take(ctx, (size_t)M * 4);
    int32_t *d_nm = (int32_t *)arena_take(ctx, 256);
    if (!D || !lists || !counts || !choice || !d_nm) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
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
    const size_t lock_bytes = (size_t)n * sizeof(int);
    if (lock_bytes > 200 * 1024) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "frame too large for the shared-memory lock table");
    proj_resolve_kernel<<<1, RESOLVE_THREADS, lock_bytes, ctx->stream>>>(v, mv, lists, stride, counts, nnratio, cur_obs, choice, d_kpmp,
                                                                        d_nm, ctx->d_counters);
    ctx->launches += 2;
    CU_TRY(cudaGetLastError());
    const OutPiece out[2] = {{kp_mp, d_kpmp, (size_t)n * 4}, {nmatches, d_nm, 4}};
    return ctx_download(ctx, out, 2);
}

int search_proj_device_init() { return set_max_dyn_smem(proj_resolve_kernel); }
