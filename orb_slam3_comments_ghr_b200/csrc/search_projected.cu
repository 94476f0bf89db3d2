// search_projected.cu -- the common skeleton of the projection-gated searches that project on their own (SURVEY.md
// row a6): SearchByProjection(Cur, Last) (ORBmatcher.cc:1957-2191), SearchByProjection(Cur, KF, sAlreadyFound)
// (:2203-2330), SearchByProjection(KF, Sim3, ...) x2 (:498-733), Fuse x2 (:1330-1682), SearchBySim3 (:1684-1955).
//
// After the caller's per-point prologue (pose transform, projection, frustum / distance / viewing-angle gates,
// PredictScale -- evaluated with the caller's own Sophus/Eigen, see include/orbmatch_b200/ORBmatcher.hpp) all of
// them run: GetFeaturesInArea(u, v, radius[, levels]) -> per-candidate gates -> DescriptorDistance -> best-only
// (strict <, first candidate wins) -> threshold -> assignment, optionally with a "keypoint already taken" skip rule
// that couples the points in order, optionally followed by the rotation-histogram cull.
//
//   phase 1 (one warp per point): contiguous window scan over the cell-ordered copy of the frame, static gates
//           (octave range, stereo gate :2052-2059, Fuse chi2 gate :1463-1492), XOR+POPC; ordered (dist<<20 | id) lists
//   phase 2 (one CTA): the ordered skip rule is solved as a fixed point over lock times (see search_proj.cu),
//           independent points take one round; then histogram, three maxima, cull and owner table.
#include <cstring>

#include "internal.cuh"

namespace {

struct ProjPointsView {
    int n;
    const uint4 *desc;
    const float2 *uv;
    const float *radius;
    const int32_t *min_level, *max_level;
    const float *ur;
    const uint8_t *active, *locks;
    const float *angle;
};

__global__ void projected_candidates_kernel(FrameView f, ProjPointsView pt, int stereo_gate, int chi2_gate,
                                            const float *__restrict__ inv_sigma2, uint32_t *__restrict__ lists, unsigned long long pool_cap,
                                            uint32_t *__restrict__ offs, unsigned long long *__restrict__ counters,
                                            int32_t *__restrict__ counts)
{
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= pt.n) return;
    const int lane = lane_id();
    int cnt = 0;
    if (pt.active[m]) {
        const float2 p = pt.uv[m];
        const float radius = pt.radius[m];
        const float ur = pt.ur ? pt.ur[m] : 0.f;
        const uint4 qa = pt.desc[2 * m], qb = pt.desc[2 * m + 1];
        auto gate = [&](const int4 &it) {
            if (stereo_gate && f.u_right) { // :2052-2059
                const float kr = f.u_right[it.w];
                if (kr > 0.f && fabsf(__fsub_rn(ur, kr)) > radius) return false;
            }
            if (chi2_gate) { // :1463-1492
                const float ex = __fsub_rn(p.x, __int_as_float(it.x)), ey = __fsub_rn(p.y, __int_as_float(it.y));
                const float inv = inv_sigma2[it.z & 0xffff];
                const float kr = f.u_right ? f.u_right[it.w] : -1.f;
                float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                if (kr >= 0.f) {
                    const float er = __fsub_rn(ur, kr);
                    e2 = __fadd_rn(e2, __fmul_rn(er, er));
                    if ((double)__fmul_rn(e2, inv) > 7.8) return false;
                } else if ((double)__fmul_rn(e2, inv) > 5.99) return false;
            }
            return true;
        };
        // count, reserve a range of the pool with one atomic, fill (see proj_candidates_kernel in search_proj.cu)
        cnt = window_scan_if(f, p.x, p.y, radius, pt.min_level[m], pt.max_level[m], gate, [](bool, int, int, int4) {});
        if (cnt) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&counters[1], (unsigned long long)cnt);
            base = __shfl_sync(FULL_MASK, base, 0);
            if (base + cnt <= pool_cap) {
                uint32_t *out = lists + base;
                window_scan_if(f, p.x, p.y, radius, pt.min_level[m], pt.max_level[m], gate, [&](bool ok, int pos, int slot, int4 it) {
                    if (ok) {
                        const int dist = ham256(qa, qb, f.desc_sorted[2 * slot], f.desc_sorted[2 * slot + 1]);
                        out[pos] = ((uint32_t)dist << 20) | (uint32_t)it.w;
                    }
                });
                if (lane == 0) offs[m] = (uint32_t)base;
            } else {
                cnt = 0; // pool too small: the host repeats the call with the size the cursor reports
            }
        }
    }
    if (lane == 0) counts[m] = cnt;
}

constexpr int PR_THREADS = 1024;
__global__ void __launch_bounds__(PR_THREADS)
projected_resolve_kernel(FrameView f, ProjPointsView pt, const uint32_t *__restrict__ lists, const uint32_t *__restrict__ offs,
                         const int32_t *__restrict__ counts, float max_dist, int ordered, int check_ori,
                         const uint8_t *__restrict__ kp_locked, int32_t *__restrict__ choice, int32_t *__restrict__ best_dist,
                         int32_t *__restrict__ kp_owner, int32_t *__restrict__ nmatches_out, unsigned long long *__restrict__ counters)
{
    extern __shared__ int lock_time[]; // [f.n]
    __shared__ int s_changed, s_nmatches, s_removed;
    __shared__ unsigned long long s_ncmp;
    __shared__ int hist[ORBGPU_HISTO_LENGTH];
    __shared__ int ind[3];
    const int t = threadIdx.x;
    auto reset_locks = [&]() {
        for (int k = t; k < f.n; k += PR_THREADS) lock_time[k] = (kp_locked && kp_locked[k]) ? -1 : 0x7FFFFFFF;
    };
    reset_locks();
    for (int m = t; m < pt.n; m += PR_THREADS) choice[m] = -2; // "not decided yet"
    if (t == 0) { s_changed = 0; s_nmatches = 0; s_removed = 0; s_ncmp = 0; }
    if (t < ORBGPU_HISTO_LENGTH) hist[t] = 0;
    __syncthreads();
    int nm = 0;
    unsigned long long ncmp = 0;
    for (;;) {
        nm = 0;
        ncmp = 0;
        bool changed = false;
        for (int m = t; m < pt.n; m += PR_THREADS) {
            const int cnt = counts[m];
            int pick = -1, bestDist = 256;
            if (cnt > 0) {
                const uint32_t *lst = lists + offs[m];
                int bestIdx = -1;
                for (int p = 0; p < cnt; p++) {
                    const uint32_t e = lst[p];
                    const int idx = (int)(e & 0xFFFFF), dist = (int)(e >> 20);
                    if (ordered && lock_time[idx] < m) continue; // keypoint already taken
                    ncmp++;                                      // DescriptorDistance is only called past the skip rule
                    if (dist < bestDist) {
                        bestDist = dist;
                        bestIdx = idx;
                    }
                }
                if (bestIdx >= 0 && (float)bestDist <= max_dist) pick = bestIdx;
            }
            if (pick != choice[m]) {
                choice[m] = pick;
                changed = true;
            }
            best_dist[m] = pick >= 0 ? bestDist : 256;
            nm += pick >= 0;
        }
        if (!ordered) break; // independent points: one round decides
        if (changed) s_changed = 1;
        __syncthreads();
        const bool again = s_changed != 0;
        __syncthreads();
        if (!again) break;
        if (t == 0) s_changed = 0;
        reset_locks();
        __syncthreads();
        for (int m = t; m < pt.n; m += PR_THREADS) {
            const int c = choice[m];
            if (c >= 0 && (!pt.locks || pt.locks[m])) atomicMin(&lock_time[c], m);
        }
        __syncthreads();
    }
    __syncthreads();
    // ---- rotation histogram over the accepted matches (:2068-2088), three maxima, cull (:2163-2186)
    if (check_ori && pt.angle) {
        for (int m = t; m < pt.n; m += PR_THREADS) {
            const int c = choice[m];
            if (c < 0) continue;
            const int bin = rot_bin(pt.angle[m], f.angle[c]);
            if (bin >= 0 && bin < ORBGPU_HISTO_LENGTH) atomicAdd(&hist[bin], 1);
        }
        __syncthreads();
        if (t == 0) three_maxima(hist, ORBGPU_HISTO_LENGTH, ind[0], ind[1], ind[2]);
        __syncthreads();
    }
    // ---- owner table: the last point that took a keypoint keeps it; a culled entry clears the keypoint
    if (kp_owner) {
        for (int k = t; k < f.n; k += PR_THREADS) kp_owner[k] = -1;
        __syncthreads();
        for (int m = t; m < pt.n; m += PR_THREADS)
            if (choice[m] >= 0) atomicMax(&kp_owner[choice[m]], m);
        __syncthreads();
    }
    int removed = 0;
    if (check_ori && pt.angle) {
        for (int m = t; m < pt.n; m += PR_THREADS) {
            const int c = choice[m];
            if (c < 0) continue;
            const int bin = rot_bin(pt.angle[m], f.angle[c]);
            if (bin >= 0 && bin < ORBGPU_HISTO_LENGTH && bin != ind[0] && bin != ind[1] && bin != ind[2]) {
                removed++;                      // nmatches-- per culled entry
                if (kp_owner) kp_owner[c] = -1; // all writers store the same value
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        nm += __shfl_xor_sync(FULL_MASK, nm, o);
        removed += __shfl_xor_sync(FULL_MASK, removed, o);
        ncmp += __shfl_xor_sync(FULL_MASK, ncmp, o);
    }
    if ((t & 31) == 0) {
        if (nm) atomicAdd(&s_nmatches, nm);
        if (removed) atomicAdd(&s_removed, removed);
        if (ncmp) atomicAdd(&s_ncmp, ncmp);
    }
    __syncthreads();
    if (t == 0) {
        *nmatches_out = s_nmatches - s_removed;
        counters[0] = s_ncmp;
    }
}

} // namespace

extern "C" int orbgpu_search_projected(orbgpu_ctx *ctx, const orbgpu_frame *f, const orbgpu_projpoints_host *pts,
                                       const orbgpu_projsearch_params *prm, const uint8_t *kp_locked, int32_t *best_idx,
                                       int32_t *best_dist, int32_t *kp_owner, int32_t *nmatches)
{
    ARG_TRY(ctx && f && pts && prm && nmatches);
    ARG_TRY(pts->n >= 0 && (pts->n == 0 || (pts->desc && pts->uv && pts->radius && pts->min_level && pts->max_level && pts->active &&
                                            best_idx && best_dist)));
    ARG_TRY(f->n < (1 << 20));
    ARG_TRY(!(prm->chi2_gate && !prm->inv_level_sigma2));
    ARG_TRY(!((prm->stereo_gate || prm->chi2_gate) && f->u_right && !pts->ur));
    ARG_TRY(!(prm->check_ori && !pts->angle));
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    *nmatches = 0;
    const int n = f->n, M = pts->n;
    if (kp_owner)
        for (int i = 0; i < n; i++) kp_owner[i] = -1;
    if (M == 0) return ORBGPU_OK;
    if (n == 0) {
        for (int i = 0; i < M; i++) { best_idx[i] = -1; best_dist[i] = 256; }
        return ORBGPU_OK;
    }
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o_desc = take((size_t)M * 32), o_uv = take((size_t)M * 8), o_rad = take((size_t)M * 4), o_min = take((size_t)M * 4),
                 o_max = take((size_t)M * 4), o_ur = take((size_t)M * 4), o_ang = take((size_t)M * 4), o_act = take(M), o_lck = take(M),
                 o_kl = take(n), o_sig = take(64 * 4);
    const size_t up_bytes = off;
    const size_t pool_cap = list_pool_entries(ctx, M, n);
    rc = stage_reserve(ctx, up_bytes);
    if (rc) return rc;
    rc = arena_reserve(ctx, up_bytes + align256(pool_cap * 4) + 4 * align256((size_t)M * 4) + align256((size_t)n * 4) + 512);
    if (rc) return rc;
    char *H = ctx->h_stage;
    memcpy(H + o_desc, pts->desc, (size_t)M * 32);
    memcpy(H + o_uv, pts->uv, (size_t)M * 8);
    memcpy(H + o_rad, pts->radius, (size_t)M * 4);
    memcpy(H + o_min, pts->min_level, (size_t)M * 4);
    memcpy(H + o_max, pts->max_level, (size_t)M * 4);
    if (pts->ur) memcpy(H + o_ur, pts->ur, (size_t)M * 4);
    if (pts->angle) memcpy(H + o_ang, pts->angle, (size_t)M * 4);
    memcpy(H + o_act, pts->active, M);
    if (pts->locks) memcpy(H + o_lck, pts->locks, M);
    if (kp_locked) memcpy(H + o_kl, kp_locked, n);
    if (prm->chi2_gate) memcpy(H + o_sig, prm->inv_level_sigma2, (size_t)(f->n_levels < 64 ? f->n_levels : 64) * 4);
    char *D = (char *)arena_take(ctx, up_bytes);
    uint32_t *lists = (uint32_t *)arena_take(ctx, pool_cap * 4), *offs = (uint32_t *)arena_take(ctx, (size_t)M * 4);
    int32_t *counts = (int32_t *)arena_take(ctx, (size_t)M * 4), *choice = (int32_t *)arena_take(ctx, (size_t)M * 4),
            *d_bd = (int32_t *)arena_take(ctx, (size_t)M * 4), *d_owner = (int32_t *)arena_take(ctx, (size_t)n * 4),
            *d_nm = (int32_t *)arena_take(ctx, 256);
    if (!D || !lists || !offs || !counts || !choice || !d_bd || !d_owner || !d_nm) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    CU_TRY(cudaMemcpyAsync(D, H, up_bytes, cudaMemcpyHostToDevice, ctx->stream));
    ProjPointsView pv;
    pv.n = M;
    pv.desc = (const uint4 *)(D + o_desc); pv.uv = (const float2 *)(D + o_uv); pv.radius = (const float *)(D + o_rad);
    pv.min_level = (const int32_t *)(D + o_min); pv.max_level = (const int32_t *)(D + o_max);
    pv.ur = pts->ur ? (const float *)(D + o_ur) : nullptr;
    pv.angle = pts->angle ? (const float *)(D + o_ang) : nullptr;
    pv.active = (const uint8_t *)(D + o_act);
    pv.locks = pts->locks ? (const uint8_t *)(D + o_lck) : nullptr;
    const FrameView v = frame_view(f);
    projected_candidates_kernel<<<(M * 32 + 255) / 256, 256, 0, ctx->stream>>>(v, pv, prm->stereo_gate, prm->chi2_gate,
                                                                              (const float *)(D + o_sig), lists, pool_cap, offs, ctx->d_counters, counts);
    const size_t lock_bytes = (size_t)n * sizeof(int);
    if (lock_bytes > 200 * 1024) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "frame too large for the shared-memory lock table");
    projected_resolve_kernel<<<1, PR_THREADS, lock_bytes, ctx->stream>>>(v, pv, lists, offs, counts, prm->max_dist, prm->ordered,
                                                                        prm->check_ori, kp_locked ? (const uint8_t *)(D + o_kl) : nullptr,
                                                                        choice, d_bd, kp_owner ? d_owner : nullptr, d_nm, ctx->d_counters);
    ctx->launches += 2;
    CU_TRY(cudaGetLastError());
    const OutPiece out[4] = {{best_idx, choice, (size_t)M * 4}, {best_dist, d_bd, (size_t)M * 4},
                             {kp_owner, kp_owner ? d_owner : nullptr, (size_t)n * 4}, {nmatches, d_nm, 4}};
    rc = ctx_download(ctx, out, 4);
    if (rc) return rc;
    if (list_pool_overflowed(ctx, pool_cap)) // the pool was too small for these windows: once more with the size the kernels counted
        return orbgpu_search_projected(ctx, f, pts, prm, kp_locked, best_idx, best_dist, kp_owner, nmatches);
    return ORBGPU_OK;
}

int search_projected_device_init() { return set_max_dyn_smem(projected_resolve_kernel); }
