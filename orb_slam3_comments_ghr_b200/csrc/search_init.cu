// search_init.cu -- ORBmatcher::SearchForInitialization (ORBmatcher.cc:735-878), config C1.
//
// The reference loop carries state from one F1 keypoint to the next (vMatchedDistance,
// vnMatches21), so the work is split in two phases:
//   phase 1 (parallel, one warp per F1 keypoint): enumerate the F2 window candidates in the
//           reference's iteration order and compute their Hamming distances; store
//           (dist<<20 | i2) lists.  This is where every DescriptorDistance call happens.
//   phase 2 (one CTA): the ordered part -- the `vMatchedDistance[i2] <= dist` filter (:790), the
//           lexicographic (dist, position) top-2, thresholds, match displacement (:813-817) and the
//           rotation histogram.  Solved as a fixed point over the keypoints' ACCEPTS: the matched distance
//           keypoint k sees at partner i2 is the smallest distance among the accepts onto i2 by keypoints
//           before k, so with an inverse index (partner -> keypoints that list it) every keypoint is
//           re-decided in parallel against the current accepts until a sweep changes nothing; keypoint k
//           only depends on keypoints before it, so the fixed point is the sequential result.  Cases whose
//           lists do not fit the shared-memory budget are replayed in order by one warp.
#include <climits>

#include "internal.cuh"

namespace {

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;

__global__ void init_candidates_kernel(FrameView f1, FrameView f2, const float2 *__restrict__ prev, float window,
                                       uint32_t *__restrict__ lists, int stride, int32_t *__restrict__ counts,
                                       unsigned long long *__restrict__ counters)
{
    const int i1 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i1 >= f1.n) return;
    const int lane = lane_id();
    const int level1 = f1.octave[i1];
    int cnt = 0;
    if (level1 <= 0) { // :762 only level-0 keypoints
        const float2 p = prev[i1];
        const uint4 qa = f1.desc[2 * i1], qb = f1.desc[2 * i1 + 1];
        uint32_t *out = lists + (size_t)i1 * stride;
        cnt = window_scan(f2, p.x, p.y, window, level1, level1, [&](bool ok, int pos, int slot, int4 it) {
            if (ok) {
                const int dist = ham256(qa, qb, f2.desc_sorted[2 * slot], f2.desc_sorted[2 * slot + 1]);
                out[pos] = ((uint32_t)dist << 20) | (uint32_t)it.w;
            }
        });
    }
    if (lane == 0) {
        counts[i1] = cnt;
        if (cnt) atomicAdd(&counters[0], (unsigned long long)cnt);
    }
}

// Ordered pass: one CTA.  All threads first compact the non-empty candidate lists into shared memory (block scan of
// the counts) together with vMatchedDistance / vnMatches21, so that the single warp that replays the reference's
// loop order only touches shared memory: ~100 cycles per level-0 keypoint instead of several L2 round trips.
constexpr int INIT_THREADS = 1024;
constexpr int INIT_LIST_CAP = 24 * 1024; // staged list entries (96 KB); larger cases replay from global memory
constexpr int INIT_FP_CAP = INIT_LIST_CAP / 2; // fixed-point path: lists in the first half, the inverse index in the second
__global__ void __launch_bounds__(INIT_THREADS)
init_resolve_kernel(FrameView f1, FrameView f2, float2 *__restrict__ prev, const uint32_t *__restrict__ lists, int stride,
                    const int32_t *__restrict__ counts, float nnratio, int check_ori, int32_t *__restrict__ bin_of,
                    int32_t *__restrict__ matches12, int32_t *__restrict__ nmatches_out)
{
    extern __shared__ int init_smem[];
    int *vMatchedDistance = init_smem;           // [n2]
    int *vnMatches21 = vMatchedDistance + f2.n;  // [n2]
    int *sAct = vnMatches21 + f2.n;              // [n1] active (non-empty) F1 keypoints, ascending
    int *sOff = sAct + f1.n;                     // [n1] start of their lists
    int *choice = sOff + f1.n;                   // [n1] fixed-point path: accept of active keypoint k (i2 << 9 | dist), -1 none
    uint32_t *sLists = (uint32_t *)(choice + f1.n); // [INIT_LIST_CAP]
    __shared__ int s_changed;
    __shared__ int hist[ORBGPU_HISTO_LENGTH];
    __shared__ int ind[3];
    __shared__ int s_nmatches, s_removed, s_nact, s_total;
    __shared__ int warp_cnt[32], warp_act[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < f2.n; i += INIT_THREADS) {
        vMatchedDistance[i] = INT_MAX; // :752
        vnMatches21[i] = -1;           // :754
    }
    for (int i = t; i < f1.n; i += INIT_THREADS) {
        matches12[i] = -1; // :739
        bin_of[i] = -1;
    }
    if (t < ORBGPU_HISTO_LENGTH) hist[t] = 0;
    if (t == 0) { s_nmatches = 0; s_removed = 0; }
    // ---- block scan over contiguous chunks of F1 keypoints: list offsets and the compact index of the active ones
    const int per = (f1.n + INIT_THREADS - 1) / INIT_THREADS;
    const int lo = min(f1.n, t * per), hi = min(f1.n, lo + per);
    int my_cnt = 0, my_act = 0;
    for (int i = lo; i < hi; i++) {
        const int c = counts[i];
        my_cnt += c;
        my_act += c > 0;
    }
    int inc_cnt = my_cnt, inc_act = my_act;
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(FULL_MASK, inc_cnt, o), b = __shfl_up_sync(FULL_MASK, inc_act, o);
        if (lane >= o) { inc_cnt += a; inc_act += b; }
    }
    if (lane == 31) { warp_cnt[warp] = inc_cnt; warp_act[warp] = inc_act; }
    __syncthreads();
    if (warp == 0) {
        int a = warp_cnt[lane], b = warp_act[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, a, o), v = __shfl_up_sync(FULL_MASK, b, o);
            if (lane >= o) { a += u; b += v; }
        }
        warp_cnt[lane] = a; warp_act[lane] = b;
        if (lane == 31) { s_total = a; s_nact = b; }
    }
    __syncthreads();
    const bool staged = s_total <= INIT_LIST_CAP;
    {
        int off = (warp ? warp_cnt[warp - 1] : 0) + inc_cnt - my_cnt;
        int k = (warp ? warp_act[warp - 1] : 0) + inc_act - my_act;
        for (int i = lo; i < hi; i++) {
            const int c = counts[i];
            if (c > 0) {
                sAct[k] = i;
                sOff[k] = staged ? off : i * stride;
                k++;
                off += c;
            }
        }
    }
    __syncthreads();
    const int nact = s_nact;
    if (staged) { // one warp per active keypoint copies its list
        for (int k = warp; k < nact; k += INIT_THREADS / 32) {
            const int i1 = sAct[k], c = counts[i1];
            const uint32_t *src = lists + (size_t)i1 * stride;
            for (int p = lane; p < c; p += 32) sLists[sOff[k] + p] = src[p];
        }
    }
    __syncthreads();
    const uint32_t *L = staged ? sLists : lists;
    const bool fixed_point = staged && s_total <= INIT_FP_CAP;
    if (fixed_point) {
        const int total = s_total, n2 = f2.n;
        // accepts grouped by partner: partner i2 owns acc[accOff[i2] ...] with room for every keypoint that lists it
        uint32_t *acc = sLists + INIT_FP_CAP; // [total] (k << 9 | dist)
        int *accOff = vMatchedDistance, *accCnt = vnMatches21; // the replay state is not needed on this path
        for (int i = t; i < n2; i += INIT_THREADS) accCnt[i] = 0;
        for (int k = t; k < nact; k += INIT_THREADS) choice[k] = -1;
        __syncthreads();
        for (int e = t; e < total; e += INIT_THREADS) atomicAdd(&accCnt[sLists[e] & 0xFFFFF], 1);
        __syncthreads();
        { // exclusive scan of the partners' capacities over contiguous chunks
            const int per2 = (n2 + INIT_THREADS - 1) / INIT_THREADS;
            const int lo2 = min(n2, t * per2), hi2 = min(n2, lo2 + per2);
            int mine = 0;
            for (int i = lo2; i < hi2; i++) mine += accCnt[i];
            int incl = mine;
            for (int o = 1; o < 32; o <<= 1) {
                const int a = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o) incl += a;
            }
            if (lane == 31) warp_cnt[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                int a = warp_cnt[lane];
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(FULL_MASK, a, o);
                    if (lane >= o) a += u;
                }
                warp_cnt[lane] = a;
            }
            __syncthreads();
            int run = (warp ? warp_cnt[warp - 1] : 0) + incl - mine;
            for (int i = lo2; i < hi2; i++) {
                accOff[i] = run;
                run += accCnt[i];
            }
        }
        __syncthreads();
        for (int iter = 0; iter <= nact + 1; iter++) {
            // the accepts of the previous sweep, by partner
            for (int i = t; i < n2; i += INIT_THREADS) accCnt[i] = 0;
            __syncthreads();
            if (t == 0) s_changed = 0; // every thread has read the previous sweep's flag before the barrier above
            for (int k = t; k < nact; k += INIT_THREADS) {
                const int c = choice[k];
                if (c >= 0) acc[accOff[c >> 9] + atomicAdd(&accCnt[c >> 9], 1)] = ((uint32_t)k << 9) | (uint32_t)(c & 0x1FF);
            }
            __syncthreads();
            for (int k = warp; k < nact; k += INIT_THREADS / 32) {
                const int s0 = sOff[k], cnt = (k + 1 < nact ? sOff[k + 1] : total) - s0;
                uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
                for (int p = lane; p < cnt; p += 32) {
                    const uint32_t e = sLists[s0 + p];
                    const int dist = (int)(e >> 20), i2 = (int)(e & 0xFFFFF);
                    int md = INT_MAX; // vMatchedDistance[i2] as keypoint k finds it (:752, :822): accepts by keypoints before k
                    for (int x = accOff[i2], xe = x + accCnt[i2]; x < xe; x++) {
                        const uint32_t v = acc[x];
                        if ((int)(v >> 9) < k) md = min(md, (int)(v & 0x1FF));
                    }
                    if (!(md <= dist)) top2_push(b1, b2, ((uint32_t)dist << 20) | (uint32_t)p); // :790
                }
                uint32_t m1, m2;
                warp_top2(b1, b2, m1, m2);
                int c = -1;
                if (m1 != KEY_NONE) {
                    const int bestDist = (int)(m1 >> 20);
                    const float second = (m2 == KEY_NONE) ? (float)INT_MAX : (float)(int)(m2 >> 20);
                    if (bestDist <= ORBGPU_TH_LOW && (float)bestDist < __fmul_rn(second, nnratio)) // :807, :810
                        c = (int)((sLists[s0 + (m1 & 0xFFFFF)] & 0xFFFFF) << 9) | bestDist;
                }
                if (lane == 0 && c != choice[k]) { // the sweep reads acc[], not choice[]
                    choice[k] = c;
                    s_changed = 1;
                }
            }
            __syncthreads();
            if (!s_changed) break; // acc[] holds exactly the final accepts
        }
        int kept = 0;
        for (int k = t; k < nact; k += INIT_THREADS) {
            const int c = choice[k];
            if (c < 0) continue;
            const int i2 = c >> 9, i1 = sAct[k];
            bin_of[i1] = i2; // partner at accept time (for the rotation histogram), displaced or not
            bool displaced = false; // :813-817: a later accept onto the same partner takes the match away
            for (int x = accOff[i2], xe = x + accCnt[i2]; x < xe; x++)
                if ((int)(acc[x] >> 9) > k) displaced = true;
            if (!displaced) {
                matches12[i1] = i2;
                kept++;
            }
        }
        for (int o = 16; o; o >>= 1) kept += __shfl_xor_sync(FULL_MASK, kept, o);
        if (lane == 0 && kept) atomicAdd(&s_nmatches, kept);
    } else if (t < 32) {
        int nmatches = 0;
        for (int k = 0; k < nact; k++) { // keypoints with level > 0 (:762) or an empty window (:771) are not listed
            const int i1 = sAct[k];
            const int cnt = (k + 1 < nact && staged) ? sOff[k + 1] - sOff[k] : counts[i1];
            const uint32_t *lst = L + sOff[k];
            uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
            for (int base = 0; base < cnt; base += 32) {
                const int p = base + lane;
                if (p < cnt) {
                    const uint32_t e = lst[p];
                    const int dist = (int)(e >> 20), i2 = (int)(e & 0xFFFFF);
                    if (!(vMatchedDistance[i2] <= dist)) // :790
                        top2_push(b1, b2, ((uint32_t)dist << 20) | (uint32_t)p);
                }
            }
            uint32_t m1, m2;
            warp_top2(b1, b2, m1, m2);
            if (m1 != KEY_NONE) {
                const int bestDist = (int)(m1 >> 20);
                if (bestDist <= ORBGPU_TH_LOW) { // :807
                    const float second = (m2 == KEY_NONE) ? (float)INT_MAX : (float)(int)(m2 >> 20);
                    if ((float)bestDist < __fmul_rn(second, nnratio)) { // :810
                        if (lane == 0) {
                            const int bestIdx2 = (int)(lst[m1 & 0xFFFFF] & 0xFFFFF);
                            const int prev_owner = vnMatches21[bestIdx2];
                            if (prev_owner >= 0) { // :813-817
                                matches12[prev_owner] = -1;
                                nmatches--;
                            }
                            matches12[i1] = bestIdx2;
                            vnMatches21[bestIdx2] = i1;
                            vMatchedDistance[bestIdx2] = bestDist;
                            bin_of[i1] = bestIdx2; // partner at accept time (for the rotation histogram)
                            nmatches++;
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (lane == 0) s_nmatches = nmatches;
    }
    __syncthreads();
    if (check_ori) {
        // :826-840 -- every keypoint that was accepted goes into its bin, also when it is displaced later (:813-817 do
        // not touch rotHist).  The replay left the accepted partner in bin_of; the angles are fetched here, in parallel.
        for (int i1 = t; i1 < f1.n; i1 += INIT_THREADS) {
            const int i2 = bin_of[i1];
            int bin = -1;
            if (i2 >= 0) {
                bin = rot_bin(f1.angle[i1], f2.angle[i2]);
                if (bin >= 0 && bin < ORBGPU_HISTO_LENGTH) atomicAdd(&hist[bin], 1); else bin = -1;
            }
            bin_of[i1] = bin;
        }
        __syncthreads();
        if (t == 0) three_maxima(hist, ORBGPU_HISTO_LENGTH, ind[0], ind[1], ind[2]);
        __syncthreads();
        for (int i1 = t; i1 < f1.n; i1 += INIT_THREADS) { // :846-869
            const int b = bin_of[i1];
            if (b >= 0 && b != ind[0] && b != ind[1] && b != ind[2] && matches12[i1] >= 0) {
                matches12[i1] = -1;
                atomicAdd(&s_removed, 1);
            }
        }
    }
    __syncthreads();
    for (int i1 = t; i1 < f1.n; i1 += INIT_THREADS) { // :873-875
        const int m = matches12[i1];
        if (m >= 0) prev[i1] = f2.xy[m];
    }
    if (t == 0) *nmatches_out = s_nmatches - s_removed;
}

} // namespace

extern "C" int orbgpu_search_for_initialization(orbgpu_ctx *ctx, const orbgpu_frame *f1, const orbgpu_frame *f2,
                                                float *prev_matched_xy, int32_t window_size, float nnratio, int32_t check_ori,
                                                int32_t *matches12, int32_t *nmatches)
{
    ARG_TRY(ctx && f1 && f2 && matches12 && nmatches);
    ARG_TRY(f1->n == 0 || prev_matched_xy);
    ARG_TRY(f2->n < (1 << 20));
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    *nmatches = 0;
    const int n1 = f1->n, n2 = f2->n;
    if (n1 == 0) return ORBGPU_OK;
    const int stride = n2 > 0 ? n2 : 1;
    const size_t b1 = align256((size_t)n1 * 4), b2 = align256((size_t)(n2 + 1) * 4);
    rc = arena_reserve(ctx, align256((size_t)n1 * 8) + align256((size_t)n1 * stride * 4) + 3 * b1 + 256);
    (void)b2;
    if (rc) return rc;
    float2 *d_prev = (float2 *)arena_take(ctx, (size_t)n1 * 8);
    uint32_t *lists = (uint32_t *)arena_take(ctx, (size_t)n1 * stride * 4);
    int32_t *counts = (int32_t *)arena_take(ctx, n1 * 4), *bin_of = (int32_t *)arena_take(ctx, n1 * 4),
            *d_m12 = (int32_t *)arena_take(ctx, n1 * 4);
    int32_t *d_nm = (int32_t *)arena_take(ctx, 256);
    CU_TRY(cudaMemcpyAsync(d_prev, prev_matched_xy, (size_t)n1 * 8, cudaMemcpyHostToDevice, ctx->stream));
    const FrameView v1 = frame_view(f1), v2 = frame_view(f2);
    init_candidates_kernel<<<(n1 * 32 + 255) / 256, 256, 0, ctx->stream>>>(v1, v2, d_prev, (float)window_size, lists, stride, counts,
                                                                          ctx->d_counters);
    const size_t smem = ((size_t)2 * n2 + (size_t)3 * n1 + INIT_LIST_CAP) * 4;
    if (smem > 220 * 1024) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "frames too large for the shared-memory replay state");
    CU_TRY(cudaFuncSetAttribute(init_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    init_resolve_kernel<<<1, INIT_THREADS, smem, ctx->stream>>>(v1, v2, d_prev, lists, stride, counts, nnratio, check_ori, bin_of, d_m12,
                                                                d_nm);
    ctx->launches += 2;
    CU_TRY(cudaGetLastError());
    const OutPiece out[3] = {{matches12, d_m12, (size_t)n1 * 4}, {prev_matched_xy, d_prev, (size_t)n1 * 8}, {nmatches, d_nm, 4}};
    return ctx_download(ctx, out, 3);
}
