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
//           only depends on keypoints before it, so the fixed point is the sequential result.  Lists and state that
//           do not fit the shared-memory budget stay in global memory (same algorithm).
#include <climits>

#include "internal.cuh"

namespace {

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;

// lists: one row per LEVEL-0 keypoint of F1 (row = rank0[i1], the keypoint's rank among the level-0 keypoints, built at upload),
// stride = number of level-0 keypoints of F2 -- the window only admits octave 0 (:768), so a list cannot be longer
__global__ void init_candidates_kernel(FrameView f1, FrameView f2, const int32_t *__restrict__ rank0, const float2 *__restrict__ prev,
                                       float window, uint32_t *__restrict__ lists, int stride, int32_t *__restrict__ counts,
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
        uint32_t *out = lists + (size_t)rank0[i1] * stride;
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
// the counts), then iterate the accept fixed point described above.
constexpr int INIT_THREADS = 1024;
constexpr int INIT_LIST_CAP = 24 * 1024; // staged list entries (96 KB); longer lists are read from global memory
constexpr int INIT_FP_CAP = INIT_LIST_CAP / 2; // fixed-point path: lists in the first half, the inverse index in the second
// Shared-memory layout: [sLists INIT_LIST_CAP][per-partner state 2 x n2, when g_state == null][per-active-keypoint state
// 3 x nact_cap, when g_act == null].  Frames too large for that (the monocular initialisation extracts 5 x nFeatures key points,
// Tracking.cc:667) keep the state arrays -- and lists beyond INIT_LIST_CAP, and their inverse index g_acc -- in global memory
// (L2 resident): same algorithm, same results.
__global__ void __launch_bounds__(INIT_THREADS)
init_resolve_kernel(FrameView f1, FrameView f2, const int32_t *__restrict__ rank0, float2 *__restrict__ prev,
                    const uint32_t *__restrict__ lists, int stride, const int32_t *__restrict__ counts, float nnratio, int check_ori,
                    int nact_cap, int *__restrict__ g_state, int *__restrict__ g_act, uint32_t *__restrict__ g_acc,
                    int32_t *__restrict__ bin_of, int32_t *__restrict__ matches12, int32_t *__restrict__ nmatches_out)
{
    extern __shared__ int init_smem[];
    uint32_t *sLists = (uint32_t *)init_smem;    // [INIT_LIST_CAP]
    int *vMatchedDistance = g_state ? g_state : init_smem + INIT_LIST_CAP; // [n2]
    int *vnMatches21 = vMatchedDistance + f2.n;  // [n2]
    int *sAct = g_act ? g_act : (g_state ? init_smem + INIT_LIST_CAP : vnMatches21 + f2.n); // [nact_cap] active (non-empty) F1 keypoints, ascending
    int *sOff = sAct + nact_cap;                 // [nact_cap] start of their lists
    int *choice = sOff + nact_cap;               // [nact_cap] accept of active keypoint k (i2 << 9 | dist), -1 none
    __shared__ int s_changed;
    __shared__ int hist[ORBGPU_HISTO_LENGTH];
    __shared__ int ind[3];
    __shared__ int s_nmatches, s_removed, s_nact, s_total;
    __shared__ int warp_cnt[32], warp_act[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
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
                sOff[k] = staged ? off : rank0[i] * stride;
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
            const uint32_t *src = lists + (size_t)rank0[i1] * stride;
            for (int p = lane; p < c; p += 32) sLists[sOff[k] + p] = src[p];
        }
    }
    __syncthreads();
    const uint32_t *L = staged ? sLists : lists;
    {
        const int total = s_total, n2 = f2.n;
        // accepts grouped by partner: partner i2 owns acc[accOff[i2] ...] with room for every keypoint that lists it.  Lists of
        // up to INIT_FP_CAP entries leave the second half of the staging area to the inverse index; longer ones keep it in global memory
        uint32_t *acc = (staged && total <= INIT_FP_CAP) ? sLists + INIT_FP_CAP : g_acc; // [total] (k << 9 | dist)
        int *accOff = vMatchedDistance, *accCnt = vnMatches21;
        for (int i = t; i < n2; i += INIT_THREADS) accCnt[i] = 0;
        for (int k = t; k < nact; k += INIT_THREADS) choice[k] = -1;
        __syncthreads();
        if (staged) {
            for (int e = t; e < total; e += INIT_THREADS) atomicAdd(&accCnt[sLists[e] & 0xFFFFF], 1);
        } else {
            for (int k = warp; k < nact; k += INIT_THREADS / 32) {
                const int c = counts[sAct[k]];
                for (int p2 = lane; p2 < c; p2 += 32) atomicAdd(&accCnt[L[sOff[k] + p2] & 0xFFFFF], 1);
            }
        }
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
                const int s0 = sOff[k], cnt = staged ? (k + 1 < nact ? sOff[k + 1] : total) - s0 : counts[sAct[k]];
                uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
                for (int p = lane; p < cnt; p += 32) {
                    const uint32_t e = L[s0 + p];
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
                        c = (int)((L[s0 + (m1 & 0xFFFFF)] & 0xFFFFF) << 9) | bestDist;
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
    // one list row per level-0 key point of F1 (:762), as long as F2 has level-0 key points (:768 searches octave 0 only)
    const int r1 = f1->n_level0 > 0 ? f1->n_level0 : 1, stride = f2->n_level0 > 0 ? f2->n_level0 : 1;
    const size_t list_entries = (size_t)r1 * stride;
    // shared memory: staged lists always; per-partner and per-key-point state when they fit beside them (checked BEFORE any launch)
    const size_t lists_b = (size_t)INIT_LIST_CAP * 4, state_b = (size_t)2 * n2 * 4, act_b = (size_t)3 * r1 * 4;
    const size_t budget = ORBGPU_SMEM_OPTIN - 2048 - lists_b;
    const bool state_smem = state_b + act_b <= budget, act_smem = state_smem || act_b <= budget;
    const size_t smem = lists_b + (state_smem ? state_b : 0) + (act_smem ? act_b : 0);
    const size_t b1 = align256((size_t)n1 * 4);
    rc = arena_reserve(ctx, align256((size_t)n1 * 8) + 2 * align256(list_entries * 4) + 3 * b1 + align256(state_b) + align256(act_b) + 1024);
    if (rc) return rc;
    float2 *d_prev = (float2 *)arena_take(ctx, (size_t)n1 * 8);
    uint32_t *lists = (uint32_t *)arena_take(ctx, list_entries * 4), *g_acc = (uint32_t *)arena_take(ctx, list_entries * 4);
    int32_t *counts = (int32_t *)arena_take(ctx, n1 * 4), *bin_of = (int32_t *)arena_take(ctx, n1 * 4),
            *d_m12 = (int32_t *)arena_take(ctx, n1 * 4);
    int *g_state = state_smem ? nullptr : (int *)arena_take(ctx, state_b), *g_act = act_smem ? nullptr : (int *)arena_take(ctx, act_b);
    int32_t *d_nm = (int32_t *)arena_take(ctx, 256);
    if (!d_prev || !lists || !g_acc || !counts || !bin_of || !d_m12 || !d_nm || (!state_smem && !g_state) || (!act_smem && !g_act))
        return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    CU_TRY(cudaMemcpyAsync(d_prev, prev_matched_xy, (size_t)n1 * 8, cudaMemcpyHostToDevice, ctx->stream));
    const FrameView v1 = frame_view(f1), v2 = frame_view(f2);
    init_candidates_kernel<<<(n1 * 32 + 255) / 256, 256, 0, ctx->stream>>>(v1, v2, f1->rank0, d_prev, (float)window_size, lists, stride, counts,
                                                                          ctx->d_counters);
    init_resolve_kernel<<<1, INIT_THREADS, smem, ctx->stream>>>(v1, v2, f1->rank0, d_prev, lists, stride, counts, nnratio, check_ori, r1, g_state,
                                                                g_act, g_acc, bin_of, d_m12, d_nm);
    ctx->launches += 2;
    CU_TRY(cudaGetLastError());
    const OutPiece out[3] = {{matches12, d_m12, (size_t)n1 * 4}, {prev_matched_xy, d_prev, (size_t)n1 * 8}, {nmatches, d_nm, 4}};
    return ctx_download(ctx, out, 3);
}

int search_init_device_init() { return set_max_dyn_smem(init_resolve_kernel); }
