// search_bow.cu -- ORBmatcher::SearchByBoW, KeyFrame<->Frame (ORBmatcher.cc:262-496, Nleft == -1
// path) and KeyFrame<->KeyFrame (ORBmatcher.cc:890-1043), config C3.
//
// The reference merge-joins the two FeatureVector maps on NodeId and, inside a shared node, walks
// the keyframe features in order; a frame feature matched by an earlier keyframe feature is
// skipped by later ones (:335 / vbMatched2 :962,:990).  A feature belongs to exactly one node, so
// that dependency is node-local: one CTA per keyframe node (binary search for the partner node in
// the other FeatureVector), keyframe features replayed in order inside the CTA, threads over the
// partner's features with a block-wide lexicographic (distance, position) top-2.  The rotation
// histogram is accumulated with atomics (bin sizes are order independent) and culled afterwards.
#include "internal.cuh"

namespace {

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr int BOW_MAX_WARPS = 32;

// MODE 0: KF <-> F   (match indexed by F feature, value = KF feature)
// MODE 1: KF1 <-> KF2 (match indexed by KF1 feature, value = KF2 feature)
template <int MODE>
__global__ void bow_match_kernel(FrameView kf, FrameView f, const uint8_t *__restrict__ kf_valid,
                                 const uint8_t *__restrict__ f_valid, float nnratio, int check_ori, int32_t *match,
                                 uint8_t *matched2, int32_t *__restrict__ bin_of, int *__restrict__ hist, int *__restrict__ nmatches,
                                 unsigned long long *__restrict__ counters, int stage_cap)
{
    // optional staging (dynamic shared memory, stage_cap descriptors): when the partner node fits, its descriptors and
    // "already matched" flags are copied once and every keyframe feature of the node is replayed against shared memory --
    // a large node (e.g. the single root bucket of levelsup >= L) is otherwise one global round trip per keyframe feature
    extern __shared__ uint4 bow_smem[];
    __shared__ uint32_t wm1[BOW_MAX_WARPS], wm2[BOW_MAX_WARPS];
    __shared__ int s_b;
    const int a = blockIdx.x;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
    if (t == 0) { // lower_bound merge-join (:292-467) == look the node id up in the other sorted list
        const uint32_t nid = kf.fv_node_ids[a];
        int lo = 0, hi = f.fv_n_nodes;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (f.fv_node_ids[mid] < nid) lo = mid + 1; else hi = mid;
        }
        s_b = (lo < f.fv_n_nodes && f.fv_node_ids[lo] == nid) ? lo : -1;
    }
    __syncthreads();
    const int b = s_b;
    if (b < 0) return;
    const int s1 = kf.fv_offsets[a], e1 = kf.fv_offsets[a + 1];
    const int s2 = f.fv_offsets[b], n2 = f.fv_offsets[b + 1] - s2;
    const bool staged = n2 <= stage_cap && (e1 - s1) >= 4; // worth it only when several keyframe features share the copy
    int *sIdx2 = (int *)(bow_smem + 2 * (size_t)stage_cap);   // [stage_cap] feature id of the partner's entry
    float *sAng2 = (float *)(sIdx2 + stage_cap);               // [stage_cap] its angle
    uint8_t *sTaken = (uint8_t *)(sAng2 + stage_cap);          // [stage_cap] already matched / not eligible
    if (staged) {
        for (int p = t; p < n2; p += blockDim.x) {
            const int idx2 = (int)f.fv_features[s2 + p];
            bow_smem[2 * p] = f.desc[2 * idx2];
            bow_smem[2 * p + 1] = f.desc[2 * idx2 + 1];
            sIdx2[p] = idx2;
            sAng2[p] = f.angle[idx2];
            sTaken[p] = (MODE == 0) ? (uint8_t)(match[idx2] >= 0) : (uint8_t)(matched2[idx2] || !f_valid[idx2]);
        }
        __syncthreads();
    }
    int my_matches = 0;
    unsigned long long ncmp = 0;
    // the keyframe feature of the NEXT iteration is fetched while the current one is scanned (three dependent global loads)
    int n_idx1 = (int)kf.fv_features[s1];
    bool n_valid = kf_valid[n_idx1] != 0;
    uint4 n_qa = kf.desc[2 * n_idx1], n_qb = kf.desc[2 * n_idx1 + 1];
    float n_ang = check_ori ? kf.angle[n_idx1] : 0.f;
    for (int iKF = s1; iKF < e1; iKF++) {
        const int idx1 = n_idx1;
        const bool valid1 = n_valid;
        const uint4 qa = n_qa, qb = n_qb;
        const float ang1 = n_ang;
        if (iKF + 1 < e1) {
            n_idx1 = (int)kf.fv_features[iKF + 1];
            n_valid = kf_valid[n_idx1] != 0;
            n_qa = kf.desc[2 * n_idx1]; n_qb = kf.desc[2 * n_idx1 + 1];
            if (check_ori) n_ang = kf.angle[n_idx1];
        }
        if (!valid1) continue; // :311-315 / :937-941 (block-uniform)
        uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
        if (staged) {
            for (int p = t; p < n2; p += blockDim.x) {
                if (sTaken[p]) continue; // :335 / :962-966
                const uint32_t d = (uint32_t)ham256(qa, qb, bow_smem[2 * p], bow_smem[2 * p + 1]);
                top2_push(b1, b2, (d << 20) | (uint32_t)p);
                ncmp++;
            }
        } else {
            for (int p = t; p < n2; p += blockDim.x) {
                const int idx2 = (int)f.fv_features[s2 + p];
                if (MODE == 0) {
                    if (match[idx2] >= 0) continue; // :335
                } else {
                    if (matched2[idx2] || !f_valid[idx2]) continue; // :962-966
                }
                const uint32_t d = (uint32_t)ham256(qa, qb, f.desc[2 * idx2], f.desc[2 * idx2 + 1]);
                top2_push(b1, b2, (d << 20) | (uint32_t)p);
                ncmp++;
            }
        }
        uint32_t m1, m2;
        warp_top2(b1, b2, m1, m2);
        if (nwarps > 1) {
            if (lane == 0) { wm1[warp] = m1; wm2[warp] = m2; }
            __syncthreads();
            if (warp == 0) { // the per-warp pairs are merged by one warp-wide top-2 (keys are unique or KEY_NONE)
                const uint32_t c1 = lane < nwarps ? wm1[lane] : KEY_NONE, c2 = lane < nwarps ? wm2[lane] : KEY_NONE;
                warp_top2(c1, c2, m1, m2);
            }
        }
        if (t == 0 && m1 != KEY_NONE) {
            const int bestDist1 = (int)(m1 >> 20);
            const int bestDist2 = (m2 == KEY_NONE) ? 256 : (int)(m2 >> 20);
            const bool th_ok = (MODE == 0) ? (bestDist1 <= ORBGPU_TH_LOW) : (bestDist1 < ORBGPU_TH_LOW); // :392 / :985
            if (th_ok && (float)bestDist1 < __fmul_rn(nnratio, (float)bestDist2)) {                     // :395 / :987
                const int best2 = staged ? sIdx2[m1 & 0xFFFFF] : (int)f.fv_features[s2 + (m1 & 0xFFFFF)];
                const int slot = (MODE == 0) ? best2 : idx1;
                if (MODE == 0) {
                    match[best2] = idx1; // vpMapPointMatches[bestIdxF] = pMP (of KF feature idx1)
                } else {
                    match[idx1] = best2; // vpMatches12[idx1] = vpMapPoints2[bestIdx2]
                    matched2[best2] = 1;
                }
                if (staged) sTaken[m1 & 0xFFFFF] = 1;
                if (check_ori) { // :405-419 / :992-1002
                    const int bin = rot_bin(ang1, staged ? sAng2[m1 & 0xFFFFF] : f.angle[best2]);
                    if (bin >= 0 && bin < ORBGPU_HISTO_LENGTH) {
                        atomicAdd(&hist[bin], 1);
                        bin_of[slot] = bin;
                    }
                }
                my_matches++;
            }
        }
        __syncthreads(); // the new match must be visible before the next keyframe feature scans
    }
    for (int off = 16; off; off >>= 1) ncmp += __shfl_xor_sync(FULL_MASK, ncmp, off);
    if (lane == 0 && ncmp) atomicAdd(&counters[0], ncmp);
    if (t == 0 && my_matches) atomicAdd(nmatches, my_matches);
}

__global__ void bow_cull_kernel(int n, int check_ori, int32_t *__restrict__ match, const int32_t *__restrict__ bin_of,
                                const int *__restrict__ hist, int *__restrict__ nmatches)
{
    __shared__ int ind[3];
    __shared__ int removed;
    if (threadIdx.x == 0) {
        removed = 0;
        three_maxima(hist, ORBGPU_HISTO_LENGTH, ind[0], ind[1], ind[2]);
    }
    __syncthreads();
    if (!check_ori) return;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { // :470-493 / :1022-1040
        const int b = bin_of[i];
        if (b >= 0 && b != ind[0] && b != ind[1] && b != ind[2]) {
            match[i] = -1;
            atomicAdd(&removed, 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *nmatches -= removed;
}

int run_bow(orbgpu_ctx *ctx, int mode, const orbgpu_frame *kf, const orbgpu_frame *f, const uint8_t *kf_valid, const uint8_t *f_valid,
            float nnratio, int check_ori, int32_t *match_out, int32_t *nmatches)
{
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    *nmatches = 0;
    const int n_out = (mode == 0) ? f->n : kf->n; // size of the match vector
    if (n_out == 0) return ORBGPU_OK;
    const size_t ob = align256((size_t)n_out * 4);
    rc = arena_reserve(ctx, 2 * ob + align256(kf->n + 1) + 2 * align256(f->n + 1) + 1024);
    if (rc) return rc;
    int32_t *d_match = (int32_t *)arena_take(ctx, (size_t)n_out * 4), *d_bin = (int32_t *)arena_take(ctx, (size_t)n_out * 4);
    uint8_t *d_kfv = (uint8_t *)arena_take(ctx, kf->n + 1), *d_fv = (uint8_t *)arena_take(ctx, f->n + 1),
            *d_m2 = (uint8_t *)arena_take(ctx, f->n + 1);
    int *d_hist = (int *)arena_take(ctx, 256); // [0..29] histogram, [32] nmatches
    int *d_nm = d_hist + 32;
    CU_TRY(cudaMemsetAsync(d_match, 0xFF, (size_t)n_out * 4, ctx->stream));
    CU_TRY(cudaMemsetAsync(d_bin, 0xFF, (size_t)n_out * 4, ctx->stream));
    CU_TRY(cudaMemsetAsync(d_hist, 0, 256, ctx->stream));
    CU_TRY(cudaMemsetAsync(d_m2, 0, f->n + 1, ctx->stream));
    if (kf->n) CU_TRY(cudaMemcpyAsync(d_kfv, kf_valid, kf->n, cudaMemcpyHostToDevice, ctx->stream));
    if (mode == 1 && f->n) CU_TRY(cudaMemcpyAsync(d_fv, f_valid, f->n, cudaMemcpyHostToDevice, ctx->stream));
    if (kf->fv_n_nodes > 0 && f->fv_n_nodes > 0) {
        const int threads = f->fv_max_node <= 32 ? 32 : (f->fv_max_node <= 512 ? 128 : (f->fv_max_node <= 1024 ? 256 : 1024));
        // staging capacity: the largest partner node, as far as shared memory goes (33 B per descriptor)
        int stage_cap = f->fv_max_node > 64 ? f->fv_max_node : 0;
        if (stage_cap > 4800) stage_cap = 4800;
        stage_cap = (stage_cap + 15) & ~15;
        const size_t smem = (size_t)stage_cap * 41; // 32 B descriptor + feature id + angle + flag
        const FrameView vk = frame_view(kf), vf = frame_view(f);
        if (mode == 0) {
            CU_TRY(cudaFuncSetAttribute(bow_match_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 1024 ? smem : 1024)));
            bow_match_kernel<0><<<kf->fv_n_nodes, threads, smem, ctx->stream>>>(vk, vf, d_kfv, d_fv, nnratio, check_ori, d_match, d_m2, d_bin,
                                                                               d_hist, d_nm, ctx->d_counters, stage_cap);
        } else {
            CU_TRY(cudaFuncSetAttribute(bow_match_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 1024 ? smem : 1024)));
            bow_match_kernel<1><<<kf->fv_n_nodes, threads, smem, ctx->stream>>>(vk, vf, d_kfv, d_fv, nnratio, check_ori, d_match, d_m2, d_bin,
                                                                               d_hist, d_nm, ctx->d_counters, stage_cap);
        }
        LAUNCH_COUNT(ctx);
        bow_cull_kernel<<<1, 256, 0, ctx->stream>>>(n_out, check_ori, d_match, d_bin, d_hist, d_nm);
        LAUNCH_COUNT(ctx);
        CU_TRY(cudaGetLastError());
    }
    CU_TRY(cudaMemcpyAsync(match_out, d_match, (size_t)n_out * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaMemcpyAsync(nmatches, d_nm, 4, cudaMemcpyDeviceToHost, ctx->stream));
    return ctx_fetch_comparisons(ctx);
}

} // namespace

__global__ void three_maxima_kernel(const int *histo, int L, int *ind)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) three_maxima(histo, L, ind[0], ind[1], ind[2]);
}

// ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2341-2383) on the device, exposed for parity tests
extern "C" int orbgpu_compute_three_maxima(orbgpu_ctx *ctx, const int32_t *histo, int32_t L, int32_t *ind)
{
    ARG_TRY(ctx && histo && ind && L > 0 && L <= 4096);
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    rc = arena_reserve(ctx, align256((size_t)L * 4) + 256);
    if (rc) return rc;
    int *d_h = (int *)arena_take(ctx, (size_t)L * 4), *d_i = (int *)arena_take(ctx, 256);
    CU_TRY(cudaMemcpyAsync(d_h, histo, (size_t)L * 4, cudaMemcpyHostToDevice, ctx->stream));
    three_maxima_kernel<<<1, 32, 0, ctx->stream>>>(d_h, L, d_i);
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(ind, d_i, 12, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return ORBGPU_OK;
}

extern "C" int orbgpu_search_by_bow_kf_f(orbgpu_ctx *ctx, const orbgpu_frame *kf, const orbgpu_frame *f, const uint8_t *kf_mp_valid,
                                         float nnratio, int32_t check_ori, int32_t *match_f2kf, int32_t *nmatches)
{
    ARG_TRY(ctx && kf && f && nmatches && (f->n == 0 || match_f2kf) && (kf->n == 0 || kf_mp_valid));
    ARG_TRY(f->n < (1 << 20));
    return run_bow(ctx, 0, kf, f, kf_mp_valid, nullptr, nnratio, check_ori, match_f2kf, nmatches);
}

extern "C" int orbgpu_search_by_bow_kf_kf(orbgpu_ctx *ctx, const orbgpu_frame *kf1, const orbgpu_frame *kf2,
                                          const uint8_t *kf1_mp_valid, const uint8_t *kf2_mp_valid, float nnratio, int32_t check_ori,
                                          int32_t *match_12, int32_t *nmatches)
{
    ARG_TRY(ctx && kf1 && kf2 && nmatches && (kf1->n == 0 || (match_12 && kf1_mp_valid)) && (kf2->n == 0 || kf2_mp_valid));
    ARG_TRY(kf2->n < (1 << 20));
    return run_bow(ctx, 1, kf1, kf2, kf1_mp_valid, kf2_mp_valid, nnratio, check_ori, match_12, nmatches);
}
