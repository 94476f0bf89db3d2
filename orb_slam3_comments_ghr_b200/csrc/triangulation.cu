// triangulation.cu -- batched ORBmatcher::SearchForTriangulation (ORBmatcher.cc:1045-1328, monocular
// pinhole path) with Pinhole::epipolarConstrain (Pinhole.cpp:189-219), config C4.
//
// In this fork vbMatched2 is never set (ORBmatcher.cc:1261-1262 are commented out), so there is no
// state carried between keyframe-1 features: for every feature idx1 without a map point the result
// is the candidate idx2 (same vocabulary node, no map point) that passes dist <= TH_LOW, the epipole
// gate and the epipolar test with the smallest distance; the reference's `dist > bestDist -> skip`
// (:1180) makes the LAST such candidate win ties, and node lists are ascending in feature id, so the
// winner is the lexicographic minimum of (dist, -idx2).
//
// HBM layout: all keyframes of a batch live in one kfset (descriptors as packed uint4 pairs,
// float2 keypoints, octave, angle, map-point mask, node id per feature) plus a per-keyframe CSR of
// the map-point-free features grouped by node (built once at upload by a per-keyframe bitonic sort).
// One CTA per pair; one warp per shared node; the n1 x n2 candidate pairs of a node are flattened
// over the lanes; survivors (rare: planted matches) reduce into shared memory with atomicMin.
#include <algorithm>
#include <cstring>

#include "internal.cuh"

namespace {

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr uint32_t NODE_NONE = 0xFFFFFFFFu;
constexpr int TRI_THREADS = 256;

// one block per keyframe: CSR (by node id) of the features WITHOUT a map point
__global__ void __launch_bounds__(1024)
kfset_csr_kernel(int n_feat, int cap, const uint32_t *__restrict__ node_id, const uint8_t *__restrict__ has_mp,
                 int32_t *__restrict__ kf_n_nodes, uint32_t *__restrict__ kf_node_ids, int32_t *__restrict__ kf_node_off,
                 int32_t *__restrict__ kf_feat)
{
    extern __shared__ unsigned long long keys[]; // [cap]
    __shared__ int s_m, s_groups;
    const int kf = blockIdx.x, t = threadIdx.x;
    const uint32_t *nid = node_id + (size_t)kf * n_feat;
    const uint8_t *mp = has_mp + (size_t)kf * n_feat;
    if (t == 0) { s_m = 0; s_groups = 0; }
    __syncthreads();
    int valid = 0;
    for (int i = t; i < cap; i += blockDim.x) {
        unsigned long long k = ~0ull;
        if (i < n_feat && !mp[i] && nid[i] != NODE_NONE) {
            k = ((unsigned long long)nid[i] << 32) | (unsigned)i;
            valid++;
        }
        keys[i] = k;
    }
    if (valid) atomicAdd(&s_m, valid);
    __syncthreads();
    for (int k = 2; k <= cap; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < cap; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    const int m = s_m;
    uint32_t *ids = kf_node_ids + (size_t)kf * n_feat;
    int32_t *off = kf_node_off + (size_t)kf * (n_feat + 1);
    int32_t *feat = kf_feat + (size_t)kf * n_feat;
    // group heads -> ordinal by counting heads before i (serial per thread chunk + block scan via atomics on ordered chunks)
    // simple two-pass: thread t owns a contiguous chunk
    const int per = (m + blockDim.x - 1) / blockDim.x;
    const int s = min(m, t * per), e = min(m, s + per);
    int heads = 0;
    for (int i = s; i < e; i++) {
        feat[i] = (int32_t)(keys[i] & 0xFFFFFFFFull);
        if (i == 0 || (uint32_t)(keys[i] >> 32) != (uint32_t)(keys[i - 1] >> 32)) heads++;
    }
    // exclusive scan of heads over threads
    __shared__ int warp_sums[32];
    int incl = heads;
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(FULL_MASK, incl, o);
        if ((t & 31) >= o) incl += u;
    }
    if ((t & 31) == 31) warp_sums[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        int w = (t < (int)(blockDim.x >> 5)) ? warp_sums[t] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, w, o);
            if (t >= o) w += u;
        }
        warp_sums[t] = w;
    }
    __syncthreads();
    int g = ((t >> 5) > 0 ? warp_sums[(t >> 5) - 1] : 0) + incl - heads;
    const int total = warp_sums[31];
    for (int i = s; i < e; i++)
        if (i == 0 || (uint32_t)(keys[i] >> 32) != (uint32_t)(keys[i - 1] >> 32)) {
            ids[g] = (uint32_t)(keys[i] >> 32);
            off[g] = i;
            g++;
        }
    if (t == 0) {
        off[total] = m;
        kf_n_nodes[kf] = total;
    }
}

struct KfSetView {
    int n_kf, n_feat;
    const uint4 *desc;
    const float2 *xy;
    const int32_t *octave;
    const float *angle;
    const float *u_right;
    const float *scale_factors;
    const float *level_sigma2;
    const int32_t *kf_n_nodes;
    const uint32_t *kf_node_ids;
    const int32_t *kf_node_off;
    const int32_t *kf_feat;
};

// Pinhole::epipolarConstrain (Pinhole.cpp:203-218) with the pair's F12 (row-major)
__device__ __forceinline__ bool epipolar_ok(const float *F, float x1, float y1, float x2, float y2, float unc)
{
    const float a = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[0]), __fmul_rn(y1, F[3])), F[6]);
    const float b = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[1]), __fmul_rn(y1, F[4])), F[7]);
    const float c = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[2]), __fmul_rn(y1, F[5])), F[8]);
    const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, x2), __fmul_rn(b, y2)), c);
    const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    if (den == 0.f) return false;
    const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
    return (double)dsqr < __dmul_rn(3.84, (double)unc); // double compare (:218)
}

__global__ void __launch_bounds__(TRI_THREADS)
triangulation_pairs_kernel(KfSetView s, int n_pairs, const int32_t *__restrict__ kf1, const int32_t *__restrict__ kf2,
                           const float *__restrict__ ep, const float *__restrict__ f12, int only_stereo, int coarse, int check_ori,
                           int32_t *__restrict__ matches12, int32_t *__restrict__ nmatches, unsigned long long *__restrict__ counters)
{
    extern __shared__ uint32_t sm_best[]; // [n_feat] keys, then [n_feat] bytes of histogram bins
    __shared__ int hist[ORBGPU_HISTO_LENGTH];
    __shared__ int ind[3];
    __shared__ float sF[9];
    __shared__ float sEp[2];
    __shared__ int s_count;
    const int p = blockIdx.x;
    if (p >= n_pairs) return;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = TRI_THREADS / 32;
    const int n = s.n_feat;
    uint8_t *sm_bin = (uint8_t *)(sm_best + n);
    const int k1 = kf1[p], k2 = kf2[p];
    for (int i = t; i < n; i += TRI_THREADS) sm_best[i] = KEY_NONE;
    if (t < ORBGPU_HISTO_LENGTH) hist[t] = 0;
    if (t < 9) sF[t] = f12[9 * (size_t)p + t];
    if (t < 2) sEp[t] = ep[2 * (size_t)p + t];
    if (t == 0) s_count = 0;
    __syncthreads();

    const uint4 *desc1 = s.desc + (size_t)k1 * n * 2, *desc2 = s.desc + (size_t)k2 * n * 2;
    const float2 *xy1 = s.xy + (size_t)k1 * n, *xy2 = s.xy + (size_t)k2 * n;
    const int32_t *oct2 = s.octave + (size_t)k2 * n;
    const float *ur1 = s.u_right ? s.u_right + (size_t)k1 * n : nullptr, *ur2 = s.u_right ? s.u_right + (size_t)k2 * n : nullptr;
    const int nn1 = s.kf_n_nodes[k1], nn2 = s.kf_n_nodes[k2];
    const uint32_t *ids1 = s.kf_node_ids + (size_t)k1 * n, *ids2 = s.kf_node_ids + (size_t)k2 * n;
    const int32_t *off1 = s.kf_node_off + (size_t)k1 * (n + 1), *off2 = s.kf_node_off + (size_t)k2 * (n + 1);
    const int32_t *feat1 = s.kf_feat + (size_t)k1 * n, *feat2 = s.kf_feat + (size_t)k2 * n;

    unsigned long long ncmp = 0;
    for (int a = warp; a < nn1; a += nwarps) {
        const uint32_t nid = ids1[a];
        int lo = 0, hi = nn2; // merge-join (:1113-1292) == lookup in the other sorted node list
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (ids2[mid] < nid) lo = mid + 1; else hi = mid;
        }
        if (lo >= nn2 || ids2[lo] != nid) continue;
        const int s1 = off1[a], n1f = off1[a + 1] - s1;
        const int s2 = off2[lo], n2f = off2[lo + 1] - s2;
        const int total = n1f * n2f;
        for (int tp = lane; tp < total; tp += 32) {
            const int i1 = tp / n2f, i2 = tp - i1 * n2f;
            const int idx1 = feat1[s1 + i1], idx2 = feat2[s2 + i2];
            const bool st1 = ur1 ? (ur1[idx1] >= 0.f) : false; // :1134
            const bool st2 = ur2 ? (ur2[idx2] >= 0.f) : false; // :1168
            if (only_stereo && (!st1 || !st2)) continue;       // :1136-1138, :1170-1172
            const int dist = ham256(desc1[2 * idx1], desc1[2 * idx1 + 1], desc2[2 * idx2], desc2[2 * idx2 + 1]);
            ncmp++;
            if (dist > ORBGPU_TH_LOW) continue; // :1180 (bestDist starts at TH_LOW)
            const float2 p2 = xy2[idx2];
            const int o2 = oct2[idx2];
            if (!st1 && !st2) { // :1191-1203 epipole gate
                const float dx = __fsub_rn(sEp[0], p2.x), dy = __fsub_rn(sEp[1], p2.y);
                if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, s.scale_factors[o2])) continue;
            }
            if (!coarse) { // :1246
                const float2 p1 = xy1[idx1];
                if (!epipolar_ok(sF, p1.x, p1.y, p2.x, p2.y, s.level_sigma2[o2])) continue;
            }
            atomicMin(&sm_best[idx1], ((uint32_t)dist << 20) | (0xFFFFFu - (uint32_t)idx2));
        }
    }
    __syncthreads();
    const float *ang1 = s.angle + (size_t)k1 * n, *ang2 = s.angle + (size_t)k2 * n;
    int mine = 0;
    if (check_ori) { // :1266-1277
        for (int i = t; i < n; i += TRI_THREADS) {
            const uint32_t key = sm_best[i];
            int bin = 255;
            if (key != KEY_NONE) {
                const int idx2 = (int)(0xFFFFFu - (key & 0xFFFFFu));
                bin = rot_bin(ang1[i], ang2[idx2]);
                if (bin >= 0 && bin < ORBGPU_HISTO_LENGTH) atomicAdd(&hist[bin], 1); else bin = 254;
            }
            sm_bin[i] = (uint8_t)bin;
        }
        __syncthreads();
        if (t == 0) three_maxima(hist, ORBGPU_HISTO_LENGTH, ind[0], ind[1], ind[2]);
        __syncthreads();
    }
    int32_t *row = matches12 + (size_t)p * n;
    for (int i = t; i < n; i += TRI_THREADS) {
        const uint32_t key = sm_best[i];
        int m = -1;
        if (key != KEY_NONE) {
            m = (int)(0xFFFFFu - (key & 0xFFFFFu));
            if (check_ori) { // :1295-1314
                const int b = sm_bin[i];
                if (b < ORBGPU_HISTO_LENGTH && b != ind[0] && b != ind[1] && b != ind[2]) m = -1;
            }
        }
        if (m >= 0) mine++;
        row[i] = m;
    }
    for (int o = 16; o; o >>= 1) {
        mine += __shfl_xor_sync(FULL_MASK, mine, o);
        ncmp += __shfl_xor_sync(FULL_MASK, ncmp, o);
    }
    if (lane == 0) {
        if (mine) atomicAdd(&s_count, mine);
        if (ncmp) atomicAdd(&counters[0], ncmp);
    }
    __syncthreads();
    if (t == 0) nmatches[p] = s_count;
}

KfSetView kfset_view(const orbgpu_kfset *s)
{
    KfSetView v;
    v.n_kf = s->n_kf; v.n_feat = s->n_feat;
    v.desc = s->desc; v.xy = s->xy; v.octave = s->octave; v.angle = s->angle; v.u_right = s->u_right;
    v.scale_factors = s->scale_factors; v.level_sigma2 = s->level_sigma2;
    v.kf_n_nodes = s->kf_n_nodes; v.kf_node_ids = s->kf_node_ids; v.kf_node_off = s->kf_node_off; v.kf_feat = s->kf_feat;
    return v;
}

} // namespace

extern "C" void orbgpu_kfset_destroy(orbgpu_kfset *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    cudaFree(s->desc); cudaFree(s->xy); cudaFree(s->octave); cudaFree(s->angle); cudaFree(s->has_mp); cudaFree(s->u_right);
    cudaFree(s->node_id); cudaFree(s->scale_factors); cudaFree(s->level_sigma2);
    cudaFree(s->kf_n_nodes); cudaFree(s->kf_node_ids); cudaFree(s->kf_node_off); cudaFree(s->kf_feat);
    delete s;
}

extern "C" int orbgpu_kfset_upload(orbgpu_ctx *ctx, const orbgpu_kfset_host *h, orbgpu_kfset **out)
{
    ARG_TRY(ctx && h && out);
    ARG_TRY(h->n_kf > 0 && h->n_feat > 0 && h->n_feat <= 8192 && h->n_feat < (1 << 20));
    ARG_TRY(h->desc && h->kp_xy && h->octave && h->angle && h->has_mp && h->node_id && h->scale_factors && h->level_sigma2);
    ARG_TRY(h->n_levels > 0 && h->n_levels <= 64);
    CU_TRY(cudaSetDevice(ctx->device));
    orbgpu_kfset *s = new orbgpu_kfset();
    s->device = ctx->device;
    s->n_kf = h->n_kf; s->n_feat = h->n_feat; s->n_levels = h->n_levels;
    const size_t T = (size_t)h->n_kf * h->n_feat;
#define UP(dst, src, bytes)                                                                         \
    do {                                                                                            \
        CU_TRY(cudaMalloc((void **)&(dst), (bytes)));                                               \
        CU_TRY(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, ctx->stream));        \
    } while (0)
    UP(s->desc, h->desc, T * 32);
    UP(s->xy, h->kp_xy, T * 8);
    UP(s->octave, h->octave, T * 4);
    UP(s->angle, h->angle, T * 4);
    UP(s->has_mp, h->has_mp, T);
    if (h->u_right) UP(s->u_right, h->u_right, T * 4);
    UP(s->node_id, h->node_id, T * 4);
    UP(s->scale_factors, h->scale_factors, (size_t)h->n_levels * 4);
    UP(s->level_sigma2, h->level_sigma2, (size_t)h->n_levels * 4);
#undef UP
    CU_TRY(cudaMalloc((void **)&s->kf_n_nodes, (size_t)h->n_kf * 4));
    CU_TRY(cudaMalloc((void **)&s->kf_node_ids, T * 4));
    CU_TRY(cudaMalloc((void **)&s->kf_node_off, (size_t)h->n_kf * (h->n_feat + 1) * 4));
    CU_TRY(cudaMalloc((void **)&s->kf_feat, T * 4));
    int cap = 1;
    while (cap < h->n_feat) cap <<= 1;
    const size_t smem = (size_t)cap * 8;
    CU_TRY(cudaFuncSetAttribute(kfset_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kfset_csr_kernel<<<h->n_kf, 1024, smem, ctx->stream>>>(h->n_feat, cap, s->node_id, s->has_mp, s->kf_n_nodes, s->kf_node_ids,
                                                          s->kf_node_off, s->kf_feat);
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    *out = s;
    return ORBGPU_OK;
}

extern "C" int orbgpu_search_for_triangulation_batch_dev(orbgpu_ctx *ctx, const orbgpu_kfset *s, int32_t n_pairs, const int32_t *kf1_dev,
                                                         const int32_t *kf2_dev, const float *ep_dev, const float *f12_dev,
                                                         int32_t only_stereo, int32_t coarse, int32_t check_ori, int32_t *matches12_dev,
                                                         int32_t *nmatches_dev)
{
    ARG_TRY(ctx && s && n_pairs >= 0);
    ARG_TRY(n_pairs == 0 || (kf1_dev && kf2_dev && ep_dev && f12_dev && matches12_dev && nmatches_dev));
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    if (n_pairs == 0) return ORBGPU_OK;
    const size_t smem = (size_t)s->n_feat * 5 + 16;
    if (smem > 48 * 1024) CU_TRY(cudaFuncSetAttribute(triangulation_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    triangulation_pairs_kernel<<<n_pairs, TRI_THREADS, smem, ctx->stream>>>(kfset_view(s), n_pairs, kf1_dev, kf2_dev, ep_dev, f12_dev,
                                                                           only_stereo, coarse, check_ori, matches12_dev, nmatches_dev,
                                                                           ctx->d_counters);
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    return ORBGPU_OK;
}

extern "C" int orbgpu_search_for_triangulation_batch(orbgpu_ctx *ctx, const orbgpu_kfset *s, int32_t n_pairs, const int32_t *kf1,
                                                     const int32_t *kf2, const float *ep, const float *f12, int32_t only_stereo,
                                                     int32_t coarse, int32_t check_ori, int32_t *matches12, int32_t *nmatches)
{
    ARG_TRY(ctx && s && n_pairs >= 0);
    ARG_TRY(n_pairs == 0 || (kf1 && kf2 && ep && f12 && matches12 && nmatches));
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    if (n_pairs == 0) return ORBGPU_OK;
    for (int p = 0; p < n_pairs; p++) ARG_TRY(kf1[p] >= 0 && kf1[p] < s->n_kf && kf2[p] >= 0 && kf2[p] < s->n_kf);
    const size_t P = (size_t)n_pairs;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o_k1 = take(P * 4), o_k2 = take(P * 4), o_ep = take(P * 8), o_f = take(P * 36);
    const size_t up = off;
    rc = stage_reserve(ctx, up);
    if (rc) return rc;
    rc = arena_reserve(ctx, up + align256(P * s->n_feat * 4) + align256(P * 4));
    if (rc) return rc;
    char *H = ctx->h_stage;
    memcpy(H + o_k1, kf1, P * 4);
    memcpy(H + o_k2, kf2, P * 4);
    memcpy(H + o_ep, ep, P * 8);
    memcpy(H + o_f, f12, P * 36);
    char *D = (char *)arena_take(ctx, up);
    int32_t *d_m = (int32_t *)arena_take(ctx, P * s->n_feat * 4), *d_nm = (int32_t *)arena_take(ctx, P * 4);
    if (!D || !d_m || !d_nm) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    CU_TRY(cudaMemcpyAsync(D, H, up, cudaMemcpyHostToDevice, ctx->stream));
    rc = orbgpu_search_for_triangulation_batch_dev(ctx, s, n_pairs, (const int32_t *)(D + o_k1), (const int32_t *)(D + o_k2),
                                                   (const float *)(D + o_ep), (const float *)(D + o_f), only_stereo, coarse, check_ori,
                                                   d_m, d_nm);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(matches12, d_m, P * s->n_feat * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaMemcpyAsync(nmatches, d_nm, P * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return ctx_fetch_comparisons(ctx);
}
