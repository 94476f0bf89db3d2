// search_init.cu -- ORBmatcher::SearchForInitialization (ORBmatcher.cc:735-878), config C1.
//
// The reference loop carries state from one F1 keypoint to the next (vMatchedDistance,
// vnMatches21), so the work is split in two phases:
//   phase 1 (parallel, one warp per F1 keypoint): enumerate the F2 window candidates in the
//           reference's iteration order and compute their Hamming distances; store
//           (dist<<20 | i2) lists.  This is where every DescriptorDistance call happens.
//   phase 2 (ordered, one warp): walk the F1 keypoints in index order over the stored lists,
//           applying the `vMatchedDistance[i2] <= dist` filter, the lexicographic (dist,
//           position) top-2, thresholds, match displacement and the rotation histogram exactly
//           as written.  Only integer compares and two fp32 multiplies per keypoint.
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

__global__ void __launch_bounds__(256)
init_resolve_kernel(FrameView f1, FrameView f2, float2 *__restrict__ prev, const uint32_t *__restrict__ lists, int stride,
                    const int32_t *__restrict__ counts, float nnratio, int check_ori, int32_t *__restrict__ vMatchedDistance,
                    int32_t *__restrict__ vnMatches21, int32_t *__restrict__ bin_of, int32_t *__restrict__ matches12,
                    int32_t *__restrict__ nmatches_out)
{
    __shared__ int hist[ORBGPU_HISTO_LENGTH];
    __shared__ int ind[3];
    __shared__ int s_nmatches, s_removed;
    const int t = threadIdx.x, lane = t & 31;
    for (int i = t; i < f2.n; i += blockDim.x) {
        vMatchedDistance[i] = INT_MAX; // :752
        vnMatches21[i] = -1;           // :754
    }
    for (int i = t; i < f1.n; i += blockDim.x) {
        matches12[i] = -1; // :739
        bin_of[i] = -1;
    }
    if (t < ORBGPU_HISTO_LENGTH) hist[t] = 0;
    if (t == 0) { s_nmatches = 0; s_removed = 0; }
    __syncthreads();
    if (t < 32) {
        int nmatches = 0;
        for (int i1 = 0; i1 < f1.n; i1++) {
            const int cnt = counts[i1];
            if (cnt == 0) continue; // level > 0 (:762) or empty window (:771)
            const uint32_t *lst = lists + (size_t)i1 * stride;
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
                            nmatches++;
                            if (check_ori) { // :826-840
                                const int bin = rot_bin(f1.angle[i1], f2.angle[bestIdx2]);
                                if (bin >= 0 && bin < ORBGPU_HISTO_LENGTH) {
                                    hist[bin]++;
                                    bin_of[i1] = bin;
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (lane == 0) s_nmatches = nmatches;
    }
    __syncthreads();
    if (check_ori) { // :846-869
        if (t == 0) three_maxima(hist, ORBGPU_HISTO_LENGTH, ind[0], ind[1], ind[2]);
        __syncthreads();
        for (int i1 = t; i1 < f1.n; i1 += blockDim.x) {
            const int b = bin_of[i1];
            if (b >= 0 && b != ind[0] && b != ind[1] && b != ind[2] && matches12[i1] >= 0) {
                matches12[i1] = -1;
                atomicAdd(&s_removed, 1);
            }
        }
        __syncthreads();
    }
    for (int i1 = t; i1 < f1.n; i1 += blockDim.x) { // :873-875
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
    rc = arena_reserve(ctx, align256((size_t)n1 * 8) + align256((size_t)n1 * stride * 4) + 3 * b1 + 2 * b2 + 256);
    if (rc) return rc;
    float2 *d_prev = (float2 *)arena_take(ctx, (size_t)n1 * 8);
    uint32_t *lists = (uint32_t *)arena_take(ctx, (size_t)n1 * stride * 4);
    int32_t *counts = (int32_t *)arena_take(ctx, n1 * 4), *bin_of = (int32_t *)arena_take(ctx, n1 * 4),
            *d_m12 = (int32_t *)arena_take(ctx, n1 * 4);
    int32_t *vmd = (int32_t *)arena_take(ctx, (n2 + 1) * 4), *vn21 = (int32_t *)arena_take(ctx, (n2 + 1) * 4);
    int32_t *d_nm = (int32_t *)arena_take(ctx, 256);
    CU_TRY(cudaMemcpyAsync(d_prev, prev_matched_xy, (size_t)n1 * 8, cudaMemcpyHostToDevice, ctx->stream));
    const FrameView v1 = frame_view(f1), v2 = frame_view(f2);
    init_candidates_kernel<<<(n1 * 32 + 255) / 256, 256, 0, ctx->stream>>>(v1, v2, d_prev, (float)window_size, lists, stride, counts,
                                                                          ctx->d_counters);
    init_resolve_kernel<<<1, 256, 0, ctx->stream>>>(v1, v2, d_prev, lists, stride, counts, nnratio, check_ori, vmd, vn21, bin_of,
                                                    d_m12, d_nm);
    ctx->launches += 2;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(matches12, d_m12, (size_t)n1 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaMemcpyAsync(prev_matched_xy, d_prev, (size_t)n1 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaMemcpyAsync(nmatches, d_nm, 4, cudaMemcpyDeviceToHost, ctx->stream));
    return ctx_fetch_comparisons(ctx);
}
