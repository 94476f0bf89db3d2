// search_proj.cu -- ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints,
// thFarPoints) (ORBmatcher.cc:44-242, Nleft == -1 path) + RadiusByViewingCos (:245-252), config C2.
//
// Same two-phase shape as search_init.cu: the window candidates of every map point and their
// Hamming distances are computed in parallel (one warp per map point); an ordered pass then
// replays the reference's sequential rule "skip a keypoint whose current map point has
// Observations() > 0" (:102-104) -- a later map point sees the assignments of earlier ones.
#include <cstring>
#include <vector>

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

// Candidate lists live in ONE pool sized by the candidates that actually exist (a window holds a few dozen key points; a dense
// M x n matrix would be 160-400 MB for a loop-closing Fuse): every warp first counts its window, reserves a contiguous range with
// one atomic on counters[1] and then fills it.  A pool that turns out too small drops nothing silently: the cursor keeps counting,
// the host sees it exceed the capacity and repeats the call with a pool of exactly that size.
__global__ void proj_candidates_kernel(FrameView f, MapPointsView mp, float th, int far_points, float th_far,
                                       uint32_t *__restrict__ lists, unsigned long long pool_cap, uint32_t *__restrict__ offs,
                                       int32_t *__restrict__ counts, unsigned long long *__restrict__ counters)
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
        auto gate = [&](const int4 &it) {
            if (f.u_right) { // :107-117 stereo gate (static per candidate)
                const float ur = f.u_right[it.w];
                if (ur > 0.f) {
                    const float er = fabsf(__fsub_rn(xr, ur));
                    if (er > radius) return false;
                }
            }
            return true;
        };
        cnt = window_scan_if(f, p.x, p.y, radius, level - 1, level, gate, [](bool, int, int, int4) {});
        if (cnt) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&counters[1], (unsigned long long)cnt);
            base = __shfl_sync(FULL_MASK, base, 0);
            if (base + cnt <= pool_cap) {
                uint32_t *out = lists + base;
                window_scan_if(f, p.x, p.y, radius, level - 1, level, gate, [&](bool ok, int pos, int slot, int4 it) {
                    if (ok) {
                        const int dist = ham256(qa, qb, f.desc_sorted[2 * slot], f.desc_sorted[2 * slot + 1]);
                        out[pos] = ((uint32_t)dist << 20) | (uint32_t)it.w;
                    }
                });
                if (lane == 0) offs[m] = (uint32_t)base;
            } else {
                cnt = 0; // pool too small: the host repeats the call (see above)
            }
        }
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
proj_resolve_kernel(FrameView f, MapPointsView mp, const uint32_t *__restrict__ lists, const uint32_t *__restrict__ offs,
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
                const uint32_t *lst = lists + offs[m];
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

// Frame::isInFrustum, Nleft == -1 (Frame.cc:676-782), one thread per map point, the reference's fp32 operation order
// (no contraction: every product and sum is its own rounding).  Writes the members the function sets in the MapPoint;
// a point rejected after the image-bounds test keeps its projection (:712-713).
struct FrustumView {
    orbgpu_frustum_host fr;
    int n;
    const float *world_pos, *normal, *min_distance, *max_distance;
    const uint8_t *skip;
    uint8_t *in_view;
    float2 *proj_xy;
    float *proj_xr, *depth, *view_cos;
    int32_t *scale_level;
};
__global__ void frustum_kernel(const FrustumView v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    const orbgpu_frustum_host &fr = v.fr;
    uint8_t inview = 0;
    float2 uvout = make_float2(-1.f, -1.f); // :682-684
    float xr = 0.f, dep = 0.f, vc = 0.f;
    int lvl = 0;
    if (!(v.skip && v.skip[i])) {
        const float P0 = v.world_pos[3 * i], P1 = v.world_pos[3 * i + 1], P2 = v.world_pos[3 * i + 2];
        float Pc[3];
#pragma unroll
        for (int r = 0; r < 3; r++) // :695 mRcw * P + mtcw
            Pc[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(fr.Rcw[3 * r], P0), __fmul_rn(fr.Rcw[3 * r + 1], P1)), __fmul_rn(fr.Rcw[3 * r + 2], P2)),
                              fr.tcw[r]);
        const float Pc_dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(Pc[0], Pc[0]), __fmul_rn(Pc[1], Pc[1])), __fmul_rn(Pc[2], Pc[2])));
        const float PcZ = Pc[2];
        const float invz = __fdiv_rn(1.0f, PcZ);
        if (!(PcZ < 0.0f)) { // :701
            // Pinhole::project (Pinhole.cpp:64-71): fx * X / Z + cx
            const float u = __fadd_rn(__fdiv_rn(__fmul_rn(fr.K[0], Pc[0]), PcZ), fr.K[2]);
            const float w = __fadd_rn(__fdiv_rn(__fmul_rn(fr.K[1], Pc[1]), PcZ), fr.K[3]);
            if (!(u < fr.min_x || u > fr.max_x) && !(w < fr.min_y || w > fr.max_y)) { // :707-710
                uvout = make_float2(u, w); // :712-713
                const float maxDistance = __fmul_rn(1.2f, v.max_distance[i]), minDistance = __fmul_rn(0.8f, v.min_distance[i]);
                const float PO0 = __fsub_rn(P0, fr.Ow[0]), PO1 = __fsub_rn(P1, fr.Ow[1]), PO2 = __fsub_rn(P2, fr.Ow[2]);
                const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(PO0, PO0), __fmul_rn(PO1, PO1)), __fmul_rn(PO2, PO2)));
                if (!(dist < minDistance || dist > maxDistance)) { // :723
                    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(PO0, v.normal[3 * i]), __fmul_rn(PO1, v.normal[3 * i + 1])),
                                                __fmul_rn(PO2, v.normal[3 * i + 2]));
                    const float viewCos = __fdiv_rn(dot, dist); // :730
                    if (!(viewCos < fr.viewing_cos_limit)) {
                        // MapPoint::PredictScale (MapPoint.cc:722-738): log in double, rounded to float (see the header)
                        const float ratio = __fdiv_rn(v.max_distance[i], dist);
                        int nScale = (int)ceilf(__fdiv_rn((float)log((double)ratio), fr.log_scale_factor));
                        if (nScale < 0) nScale = 0;
                        else if (nScale >= fr.n_levels) nScale = fr.n_levels - 1;
                        inview = 1;
                        xr = __fsub_rn(u, __fmul_rn(fr.mbf, invz)); // :743
                        dep = Pc_dist;
                        lvl = nScale;
                        vc = viewCos;
                    }
                }
            }
        }
    }
    v.in_view[i] = inview;
    v.proj_xy[i] = uvout;
    v.proj_xr[i] = xr;
    v.depth[i] = dep;
    v.scale_level[i] = lvl;
    v.view_cos[i] = vc;
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
    const size_t pool_cap = list_pool_entries(ctx, M, n);
    rc = stage_reserve(ctx, up_bytes);
    if (rc) return rc;
    rc = arena_reserve(ctx, up_bytes + align256(pool_cap * 4) + 3 * align256((size_t)M * 4) + 512);
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
    uint32_t *lists = (uint32_t *)arena_take(ctx, pool_cap * 4), *offs = (uint32_t *)arena_take(ctx, (size_t)M * 4);
    int32_t *counts = (int32_t *)arena_take(ctx, (size_t)M * 4);
    int32_t *choice = (int32_t *)arena_take(ctx, (size_t)M * 4);
    int32_t *d_nm = (int32_t *)arena_take(ctx, 256);
    if (!D || !lists || !offs || !counts || !choice || !d_nm) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    CU_TRY(cudaMemcpyAsync(D, H, up_bytes, cudaMemcpyHostToDevice, ctx->stream));
    MapPointsView mv;
    mv.n = M;
    mv.desc = (const uint4 *)(D + o_desc); mv.proj_xy = (const float2 *)(D + o_xy);
    mv.proj_xr = mps->proj_xr ? (const float *)(D + o_xr) : nullptr;
    mv.scale_level = (const int32_t *)(D + o_lvl); mv.view_cos = (const float *)(D + o_cos); mv.depth = (const float *)(D + o_dep);
    mv.in_view = (const uint8_t *)(D + o_inv); mv.bad = (const uint8_t *)(D + o_bad); mv.n_obs = (const int32_t *)(D + o_nobs);
    int32_t *cur_obs = (int32_t *)(D + o_prior), *d_kpmp = (int32_t *)(D + o_kpmp);
    const FrameView v = frame_view(f);
    proj_candidates_kernel<<<(M * 32 + 255) / 256, 256, 0, ctx->stream>>>(v, mv, th, far_points, th_far_points, lists, pool_cap, offs, counts,
                                                                         ctx->d_counters);
    const size_t lock_bytes = (size_t)n * sizeof(int);
    if (lock_bytes > 200 * 1024) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "frame too large for the shared-memory lock table");
    proj_resolve_kernel<<<1, RESOLVE_THREADS, lock_bytes, ctx->stream>>>(v, mv, lists, offs, counts, nnratio, cur_obs, choice, d_kpmp,
                                                                        d_nm, ctx->d_counters);
    ctx->launches += 2;
    CU_TRY(cudaGetLastError());
    std::vector<int32_t> kp_in(kp_mp, kp_mp + n); // in/out argument: kept for the (rare) repeat with a larger pool
    const OutPiece out[2] = {{kp_mp, d_kpmp, (size_t)n * 4}, {nmatches, d_nm, 4}};
    rc = ctx_download(ctx, out, 2);
    if (rc) return rc;
    if (list_pool_overflowed(ctx, pool_cap)) {
        memcpy(kp_mp, kp_in.data(), (size_t)n * 4);
        return orbgpu_search_by_projection_local(ctx, f, mps, th, far_points, th_far_points, nnratio, kp_prior_obs, kp_mp, nmatches);
    }
    return ORBGPU_OK;
}

extern "C" int orbgpu_is_in_frustum(orbgpu_ctx *ctx, const orbgpu_frustum_host *fr, int32_t n, const float *world_pos, const float *normal,
                                    const float *min_distance, const float *max_distance, uint8_t *in_view, float *proj_xy, float *proj_xr,
                                    float *depth, int32_t *scale_level, float *view_cos)
{
    ARG_TRY(ctx && fr && n >= 0 && (n == 0 || (world_pos && normal && min_distance && max_distance)));
    ARG_TRY(fr->n_levels > 0 && fr->n_levels <= 64);
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    if (n == 0) return ORBGPU_OK;
    const size_t N = (size_t)n;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o_wp = take(N * 12), o_nm = take(N * 12), o_mn = take(N * 4), o_mx = take(N * 4);
    const size_t up_bytes = off;
    const size_t o_iv = take(N), o_xy = take(N * 8), o_xr = take(N * 4), o_dp = take(N * 4), o_lv = take(N * 4), o_vc = take(N * 4);
    rc = stage_reserve(ctx, up_bytes);
    if (rc) return rc;
    rc = arena_reserve(ctx, off + 256);
    if (rc) return rc;
    char *H = ctx->h_stage;
    memcpy(H + o_wp, world_pos, N * 12);
    memcpy(H + o_nm, normal, N * 12);
    memcpy(H + o_mn, min_distance, N * 4);
    memcpy(H + o_mx, max_distance, N * 4);
    char *D = (char *)arena_take(ctx, off);
    if (!D) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    CU_TRY(cudaMemcpyAsync(D, H, up_bytes, cudaMemcpyHostToDevice, ctx->stream));
    FrustumView v;
    v.fr = *fr; v.n = n;
    v.world_pos = (const float *)(D + o_wp); v.normal = (const float *)(D + o_nm);
    v.min_distance = (const float *)(D + o_mn); v.max_distance = (const float *)(D + o_mx); v.skip = nullptr;
    v.in_view = (uint8_t *)(D + o_iv); v.proj_xy = (float2 *)(D + o_xy); v.proj_xr = (float *)(D + o_xr); v.depth = (float *)(D + o_dp);
    v.scale_level = (int32_t *)(D + o_lv); v.view_cos = (float *)(D + o_vc);
    frustum_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(v);
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    const OutPiece out[6] = {{in_view, v.in_view, N}, {proj_xy, v.proj_xy, N * 8}, {proj_xr, v.proj_xr, N * 4},
                             {depth, v.depth, N * 4}, {scale_level, v.scale_level, N * 4}, {view_cos, v.view_cos, N * 4}};
    return ctx_download(ctx, out, 6);
}

// Tracking::SearchLocalPoints on the device: isInFrustum of every local map point, then SearchByProjection(Frame&, vector<MapPoint*>&)
// on the projections it left in HBM -- no host loop (5000 transforms / projections / gates on one host thread) between the two.
extern "C" int orbgpu_search_local_points(orbgpu_ctx *ctx, const orbgpu_frame *f, const orbgpu_frustum_host *fr,
                                          const orbgpu_localpoints_host *pts, float th, int32_t far_points, float th_far_points,
                                          float nnratio, const int32_t *kp_prior_obs, int32_t *kp_mp, uint8_t *in_view, int32_t *nmatches)
{
    ARG_TRY(ctx && f && fr && pts && nmatches);
    ARG_TRY(f->n == 0 || (kp_prior_obs && kp_mp));
    ARG_TRY(pts->n >= 0 && (pts->n == 0 || (pts->desc && pts->world_pos && pts->normal && pts->min_distance && pts->max_distance &&
                                            pts->bad && pts->n_obs)));
    ARG_TRY(f->n < (1 << 20) && fr->n_levels > 0 && fr->n_levels <= f->n_levels);
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    *nmatches = 0;
    const int n = f->n, M = pts->n;
    if (M == 0) return ORBGPU_OK;
    const size_t Mz = (size_t)M, nz = (size_t)(n > 0 ? n : 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o_desc = take(Mz * 32), o_wp = take(Mz * 12), o_nm = take(Mz * 12), o_mn = take(Mz * 4), o_mx = take(Mz * 4),
                 o_skip = take(Mz), o_bad = take(Mz), o_nobs = take(Mz * 4), o_prior = take(nz * 4), o_kpmp = take(nz * 4);
    const size_t up_bytes = off;
    const size_t o_iv = take(Mz), o_xy = take(Mz * 8), o_xr = take(Mz * 4), o_dp = take(Mz * 4), o_lv = take(Mz * 4), o_vc = take(Mz * 4);
    const size_t pool_cap = list_pool_entries(ctx, M, n > 0 ? n : 1);
    rc = stage_reserve(ctx, up_bytes);
    if (rc) return rc;
    rc = arena_reserve(ctx, off + align256(pool_cap * 4) + 3 * align256(Mz * 4) + 512);
    if (rc) return rc;
    char *H = ctx->h_stage;
    memcpy(H + o_desc, pts->desc, Mz * 32);
    memcpy(H + o_wp, pts->world_pos, Mz * 12);
    memcpy(H + o_nm, pts->normal, Mz * 12);
    memcpy(H + o_mn, pts->min_distance, Mz * 4);
    memcpy(H + o_mx, pts->max_distance, Mz * 4);
    if (pts->skip) memcpy(H + o_skip, pts->skip, Mz); else memset(H + o_skip, 0, Mz);
    memcpy(H + o_bad, pts->bad, Mz);
    memcpy(H + o_nobs, pts->n_obs, Mz * 4);
    if (n > 0) {
        memcpy(H + o_prior, kp_prior_obs, (size_t)n * 4);
        memcpy(H + o_kpmp, kp_mp, (size_t)n * 4);
    }
    char *D = (char *)arena_take(ctx, off);
    uint32_t *lists = (uint32_t *)arena_take(ctx, pool_cap * 4), *offs = (uint32_t *)arena_take(ctx, Mz * 4);
    int32_t *counts = (int32_t *)arena_take(ctx, Mz * 4), *choice = (int32_t *)arena_take(ctx, Mz * 4);
    int32_t *d_nm = (int32_t *)arena_take(ctx, 256);
    if (!D || !lists || !offs || !counts || !choice || !d_nm) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    CU_TRY(cudaMemcpyAsync(D, H, up_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemsetAsync(d_nm, 0, 4, ctx->stream));
    FrustumView v;
    v.fr = *fr; v.n = M;
    v.world_pos = (const float *)(D + o_wp); v.normal = (const float *)(D + o_nm);
    v.min_distance = (const float *)(D + o_mn); v.max_distance = (const float *)(D + o_mx); v.skip = (const uint8_t *)(D + o_skip);
    v.in_view = (uint8_t *)(D + o_iv); v.proj_xy = (float2 *)(D + o_xy); v.proj_xr = (float *)(D + o_xr); v.depth = (float *)(D + o_dp);
    v.scale_level = (int32_t *)(D + o_lv); v.view_cos = (float *)(D + o_vc);
    frustum_kernel<<<(M + 255) / 256, 256, 0, ctx->stream>>>(v);
    LAUNCH_COUNT(ctx);
    if (n > 0) {
        MapPointsView mv;
        mv.n = M;
        mv.desc = (const uint4 *)(D + o_desc); mv.proj_xy = v.proj_xy; mv.proj_xr = v.proj_xr; mv.scale_level = v.scale_level;
        mv.view_cos = v.view_cos; mv.depth = v.depth; mv.in_view = v.in_view; mv.bad = (const uint8_t *)(D + o_bad);
        mv.n_obs = (const int32_t *)(D + o_nobs);
        const FrameView fv = frame_view(f);
        proj_candidates_kernel<<<(M * 32 + 255) / 256, 256, 0, ctx->stream>>>(fv, mv, th, far_points, th_far_points, lists, pool_cap, offs, counts,
                                                                             ctx->d_counters);
        const size_t lock_bytes = (size_t)n * sizeof(int);
        if (lock_bytes > 200 * 1024) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "frame too large for the shared-memory lock table");
        proj_resolve_kernel<<<1, RESOLVE_THREADS, lock_bytes, ctx->stream>>>(fv, mv, lists, offs, counts, nnratio, (int32_t *)(D + o_prior),
                                                                            choice, (int32_t *)(D + o_kpmp), d_nm, ctx->d_counters);
        ctx->launches += 2;
    }
    CU_TRY(cudaGetLastError());
    std::vector<int32_t> kp_in(kp_mp, kp_mp + n);
    const OutPiece out[3] = {{kp_mp, D + o_kpmp, (size_t)n * 4}, {nmatches, d_nm, 4}, {in_view, v.in_view, Mz}};
    rc = ctx_download(ctx, out, 3);
    if (rc) return rc;
    if (list_pool_overflowed(ctx, pool_cap)) {
        if (n > 0) memcpy(kp_mp, kp_in.data(), (size_t)n * 4);
        return orbgpu_search_local_points(ctx, f, fr, pts, th, far_points, th_far_points, nnratio, kp_prior_obs, kp_mp, in_view, nmatches);
    }
    return ORBGPU_OK;
}

int search_proj_device_init() { return set_max_dyn_smem(proj_resolve_kernel); }
