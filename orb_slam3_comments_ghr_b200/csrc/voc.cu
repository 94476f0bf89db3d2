// voc.cu -- DBoW2 vocabulary descent on the device.
//
// Replaces TemplatedVocabulary::transform (TemplatedVocabulary.h:1127-1194 batch, :1216-1258 per
// feature descent with FORB::distance, FORB.cpp:92-112), BowVector::addWeight/normalize
// (BowVector.cpp:35-85) and FeatureVector::addFeature (FeatureVector.cpp:32-46).
//
//   voc_transform_kernel  one warp per feature; lanes over the children of the current node; the
//                         child with the lexicographically smallest (distance, child position) wins,
//                         i.e. strict '<' / first child on ties (:1237-1248).
//   featvec_bow_build_kernel  two blocks per frame, side by side, each a bitonic sort in shared memory:
//     job 0 (featvec_build)  (node id, feature id) keys -> CSR with ascending node ids and ascending feature
//                            ids inside a node (the iteration order of the std::map<NodeId, vector<uint>>
//                            the reference walks);
//     job 1 (bow_build)      same sort on (word id, feature id); weight of a word = idf added `count` times in
//                            feature order (repeated double +=, not count*idf), then divided by the L1 norm
//                            accumulated in ascending word-id order.
#include <algorithm>
#include <cstring>

#include "internal.cuh"

namespace {

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr int SORT_THREADS = 1024;

__global__ void voc_transform_kernel(int n, const uint4 *__restrict__ desc, const uint4 *__restrict__ node_desc,
                                     const int32_t *__restrict__ child_offsets, const uint32_t *__restrict__ child_ids,
                                     const double *__restrict__ node_weight, const uint32_t *__restrict__ node_word, int L,
                                     int levelsup, uint32_t *__restrict__ word_id, uint32_t *__restrict__ node_id,
                                     double *__restrict__ weight, unsigned long long *__restrict__ counters)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const int lane = lane_id();
    const uint4 fa = desc[2 * i], fb = desc[2 * i + 1];
    const int nid_level = L - levelsup; // :1223
    uint32_t nid = 0;                   // root when nid_level <= 0 (:1224-1225)
    uint32_t final_id = 0;
    int level = 0, ncmp = 0;
    int c0 = child_offsets[0], c1 = child_offsets[1];
    while (c1 > c0) { // do { ... } while (!isLeaf) -- the root of a non-empty vocabulary has children
        ++level;
        uint32_t best = KEY_NONE;
        for (int c = c0 + lane; c < c1; c += 32) {
            const uint32_t id = child_ids[c];
            const uint32_t d = (uint32_t)ham256(fa, fb, node_desc[2 * id], node_desc[2 * id + 1]);
            best = min(best, (d << 20) | (uint32_t)(c - c0));
        }
        best = __reduce_min_sync(FULL_MASK, best);
        final_id = child_ids[c0 + (best & 0xFFFFF)];
        ncmp += c1 - c0;
        if (level == nid_level) nid = final_id; // :1250-1251
        c0 = child_offsets[final_id];
        c1 = child_offsets[final_id + 1];
    }
    if (lane == 0) {
        word_id[i] = node_word[final_id];
        weight[i] = node_weight[final_id];
        node_id[i] = nid;
        atomicAdd(&counters[0], (unsigned long long)ncmp);
    }
}

// batched form for a key-frame set: only the FeatureVector node survives (NODE id at level L - levelsup, 0xFFFFFFFF for a
// stopped word, w == 0, which the reference drops from the FeatureVector: TemplatedVocabulary.h:1157-1161)
template <int G> // lanes per feature: the smallest power of two >= the branching factor, so that a warp descends 32/G features at once
__global__ void voc_transform_nodes_kernel(long long n, const uint4 *__restrict__ desc, const uint4 *__restrict__ node_desc,
                                           const int32_t *__restrict__ child_offsets, const uint32_t *__restrict__ child_ids,
                                           const double *__restrict__ node_weight, int L, int levelsup, uint32_t *__restrict__ node_id,
                                           unsigned long long *__restrict__ counters)
{
    const int lane = lane_id(), sub = lane & (G - 1);
    const unsigned gmask = (G == 32) ? FULL_MASK : (((1u << G) - 1u) << (lane & ~(G - 1)));
    const long long group0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 / G) + lane / G;
    const long long n_groups = (((long long)gridDim.x * blockDim.x) >> 5) * (32 / G);
    const int nid_level = L - levelsup;
    unsigned long long ncmp = 0; // one atomic per group at the end: 16 M single-address atomics would serialise the kernel
    for (long long i = group0; i < n; i += n_groups) { // persistent grid: group-uniform trip count
        const uint4 fa = desc[2 * i], fb = desc[2 * i + 1];
        uint32_t nid = 0, final_id = 0;
        int level = 0;
        int c0 = child_offsets[0], c1 = child_offsets[1];
        while (c1 > c0) { // group-uniform
            ++level;
            uint32_t best = KEY_NONE;
            for (int c = c0 + sub; c < c1; c += G) {
                const uint32_t id = child_ids[c];
                const uint32_t d = (uint32_t)ham256(fa, fb, node_desc[2 * id], node_desc[2 * id + 1]);
                best = min(best, (d << 20) | (uint32_t)(c - c0));
            }
            best = __reduce_min_sync(gmask, best);
            final_id = child_ids[c0 + (best & 0xFFFFF)];
            ncmp += (unsigned)(c1 - c0);
            if (level == nid_level) nid = final_id;
            c0 = child_offsets[final_id];
            c1 = child_offsets[final_id + 1];
        }
        if (sub == 0) node_id[i] = node_weight[final_id] > 0.0 ? nid : 0xFFFFFFFFu;
    }
    if (sub == 0 && ncmp) atomicAdd(&counters[0], ncmp);
}

// in-place ascending bitonic sort of `cap` (power of two) 64-bit keys by one block
__device__ void block_bitonic_sort(unsigned long long *keys, int cap)
{
    for (int k = 2; k <= cap; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < cap; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        keys[i] = b;
                        keys[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// block-wide exclusive scan of one int per thread (blockDim.x == SORT_THREADS)
__device__ int block_exclusive_scan(int v, int *warp_sums, int &total)
{
    const int t = threadIdx.x;
    int incl = v;
    for (int off = 1; off < 32; off <<= 1) {
        const int u = __shfl_up_sync(FULL_MASK, incl, off);
        if ((t & 31) >= off) incl += u;
    }
    if ((t & 31) == 31) warp_sums[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        int w = warp_sums[t];
        for (int off = 1; off < 32; off <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, w, off);
            if (t >= off) w += u;
        }
        warp_sums[t] = w;
    }
    __syncthreads();
    const int prefix = ((t >> 5) > 0 ? warp_sums[(t >> 5) - 1] : 0) + incl - v;
    total = warp_sums[31];
    __syncthreads();
    return prefix;
}

// groups the sorted keys (high 32 bits = group id) into CSR: ids[], offsets[], and returns counts.
// Each thread handles a contiguous run of `per` sorted positions.
__device__ void csr_from_sorted(const unsigned long long *keys, int m, uint32_t *ids, int32_t *offsets, int *warp_sums,
                                int &n_groups)
{
    const int t = threadIdx.x;
    const int per = (m + SORT_THREADS - 1) / SORT_THREADS;
    const int s = min(m, t * per), e = min(m, s + per);
    int heads = 0;
    for (int i = s; i < e; i++)
        if (i == 0 || (uint32_t)(keys[i] >> 32) != (uint32_t)(keys[i - 1] >> 32)) heads++;
    int total;
    int g = block_exclusive_scan(heads, warp_sums, total);
    for (int i = s; i < e; i++)
        if (i == 0 || (uint32_t)(keys[i] >> 32) != (uint32_t)(keys[i - 1] >> 32)) {
            ids[g] = (uint32_t)(keys[i] >> 32);
            offsets[g] = i;
            g++;
        }
    if (t == 0) offsets[total] = m;
    n_groups = total;
    __syncthreads();
}

__device__ void featvec_build(int n, int cap, const uint32_t *__restrict__ node_id, const double *__restrict__ weight,
                              unsigned long long *keys, uint32_t *__restrict__ fv_node_ids, int32_t *__restrict__ fv_offsets,
                              uint32_t *__restrict__ fv_features, int32_t *__restrict__ meta)
{
    __shared__ int warp_sums[32];
    __shared__ int s_max;
    const int t = threadIdx.x;
    if (t == 0) s_max = 0;
    int valid = 0;
    for (int i = t; i < cap; i += SORT_THREADS) {
        unsigned long long k = ~0ull;
        if (i < n && weight[i] > 0.0) { // stopped words are dropped (:1157)
            k = ((unsigned long long)node_id[i] << 32) | (unsigned)i;
            valid++;
        }
        keys[i] = k;
    }
    __syncthreads();
    int m;
    block_exclusive_scan(valid, warp_sums, m);
    block_bitonic_sort(keys, cap);
    for (int i = t; i < m; i += SORT_THREADS) fv_features[i] = (uint32_t)(keys[i] & 0xFFFFFFFFull);
    int n_nodes;
    csr_from_sorted(keys, m, fv_node_ids, fv_offsets, warp_sums, n_nodes);
    int mx = 0;
    for (int g = t; g < n_nodes; g += SORT_THREADS) mx = max(mx, fv_offsets[g + 1] - fv_offsets[g]);
    if (mx) atomicMax(&s_max, mx);
    __syncthreads();
    if (t == 0) {
        meta[0] = n_nodes;
        meta[1] = m;
        meta[2] = s_max;
    }
}

__device__ void bow_build(int n, int cap, const uint32_t *__restrict__ word_id, const double *__restrict__ weight,
                          unsigned long long *keys, uint32_t *__restrict__ bow_words, int32_t *__restrict__ tmp_offsets,
                          double *__restrict__ bow_values, int32_t *__restrict__ meta)
{
    __shared__ int warp_sums[32];
    __shared__ double s_norm;
    const int t = threadIdx.x;
    int valid = 0;
    for (int i = t; i < cap; i += SORT_THREADS) {
        unsigned long long k = ~0ull;
        if (i < n && weight[i] > 0.0) {
            k = ((unsigned long long)word_id[i] << 32) | (unsigned)i;
            valid++;
        }
        keys[i] = k;
    }
    __syncthreads();
    int m;
    block_exclusive_scan(valid, warp_sums, m);
    block_bitonic_sort(keys, cap);
    int n_words;
    csr_from_sorted(keys, m, bow_words, tmp_offsets, warp_sums, n_words);
    // BowVector::addWeight (BowVector.cpp:35-50): first insert, then += per further feature
    for (int g = t; g < n_words; g += SORT_THREADS) {
        const int s = tmp_offsets[g], e = tmp_offsets[g + 1];
        double v = weight[(uint32_t)(keys[s] & 0xFFFFFFFFull)];
        for (int i = s + 1; i < e; i++) v = __dadd_rn(v, weight[(uint32_t)(keys[i] & 0xFFFFFFFFull)]);
        bow_values[g] = v;
    }
    __syncthreads();
    // BowVector::normalize(L1) (BowVector.cpp:63-85): norm accumulated in ascending word id
    if (t == 0) {
        double norm = 0.0;
        int g = 0;
        for (; g + 8 <= n_words; g += 8) { // the loads are batched, the additions stay in word order
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = fabs(bow_values[g + u]);
#pragma unroll
            for (int u = 0; u < 8; u++) norm = __dadd_rn(norm, v[u]);
        }
        for (; g < n_words; g++) norm = __dadd_rn(norm, fabs(bow_values[g]));
        s_norm = norm;
    }
    __syncthreads();
    if (s_norm > 0.0)
        for (int g = t; g < n_words; g += SORT_THREADS) bow_values[g] = __ddiv_rn(bow_values[g], s_norm);
    if (t == 0) meta[3] = n_words;
}

// FeatureVector (block with job 0) and BowVector (job 1) of one frame, built side by side: two independent sorts of the same
// features by node id / word id.  The keys are sorted in shared memory when they fit (`keys_global` null: 8 B per key, frames
// up to 16 k features); otherwise in the frame's global scratch, one job per launch.
__global__ void __launch_bounds__(SORT_THREADS)
featvec_bow_build_kernel(int job_base, int n, int cap, const uint32_t *__restrict__ node_id, const uint32_t *__restrict__ word_id,
                         const double *__restrict__ weight, unsigned long long *keys_global, uint32_t *__restrict__ fv_node_ids,
                         int32_t *__restrict__ fv_offsets, uint32_t *__restrict__ fv_features, uint32_t *__restrict__ bow_words,
                         int32_t *__restrict__ tmp_offsets, double *__restrict__ bow_values, int32_t *__restrict__ meta)
{
    extern __shared__ unsigned long long sort_smem[];
    unsigned long long *keys = keys_global ? keys_global : sort_smem;
    if (job_base + (int)blockIdx.x == 0) featvec_build(n, cap, node_id, weight, keys, fv_node_ids, fv_offsets, fv_features, meta);
    else bow_build(n, cap, word_id, weight, keys, bow_words, tmp_offsets, bow_values, meta);
}

} // namespace

int launch_voc_transform_nodes(orbgpu_ctx *ctx, const orbgpu_voc *voc, long long n, const uint4 *desc, int levelsup, uint32_t *node_id)
{
    if (n <= 0) return ORBGPU_OK;
    const int G = voc->k <= 8 ? 8 : (voc->k <= 16 ? 16 : 32);
    const long long warps = (n + 32 / G - 1) / (32 / G);
    long long blocks = (warps * 32 + 255) / 256;
    const long long resident = (long long)ctx->sm_count * 8; // 8 x 256 threads per SM; the kernel is a grid-stride loop
    if (blocks > resident) blocks = resident;
#define LAUNCH_NODES(GG)                                                                                                              \
    voc_transform_nodes_kernel<GG><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, desc, voc->node_desc, voc->child_offsets, voc->child_ids, \
                                                                              voc->weight, voc->L, levelsup, node_id, ctx->d_counters)
    if (G == 8) LAUNCH_NODES(8);
    else if (G == 16) LAUNCH_NODES(16);
    else LAUNCH_NODES(32);
#undef LAUNCH_NODES
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    return ORBGPU_OK;
}


extern "C" int orbgpu_voc_upload(orbgpu_ctx *ctx, const orbgpu_voc_host *h, orbgpu_voc **out)
{
    ARG_TRY(ctx && h && out);
    ARG_TRY(h->n_nodes >= 1 && h->node_desc && h->child_offsets && h->child_ids && h->weight && h->word_id);
    CU_TRY(cudaSetDevice(ctx->device));
    const int nn = h->n_nodes;
    const int nc = h->child_offsets[nn];
    ARG_TRY(nc >= 0);
    for (int i = 0; i < nn; i++) ARG_TRY(h->child_offsets[i + 1] >= h->child_offsets[i] && h->child_offsets[i + 1] - h->child_offsets[i] < (1 << 20));
    for (int c = 0; c < nc; c++) ARG_TRY(h->child_ids[c] > 0 && h->child_ids[c] < (uint32_t)nn);
    orbgpu_voc *v = new orbgpu_voc();
    OwnedHandle<orbgpu_voc, orbgpu_voc_destroy> owner(v);
    v->device = ctx->device;
    v->k = h->k; v->L = h->L; v->n_nodes = nn;
    CU_TRY(cudaMalloc(&v->node_desc, (size_t)nn * 32));
    CU_TRY(cudaMalloc(&v->child_offsets, (size_t)(nn + 1) * 4));
    CU_TRY(cudaMalloc(&v->child_ids, (size_t)std::max(nc, 1) * 4));
    CU_TRY(cudaMalloc(&v->weight, (size_t)nn * 8));
    CU_TRY(cudaMalloc(&v->word_id, (size_t)nn * 4));
    CU_TRY(cudaMemcpyAsync(v->node_desc, h->node_desc, (size_t)nn * 32, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(v->child_offsets, h->child_offsets, (size_t)(nn + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (nc) CU_TRY(cudaMemcpyAsync(v->child_ids, h->child_ids, (size_t)nc * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(v->weight, h->weight, (size_t)nn * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(v->word_id, h->word_id, (size_t)nn * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    *out = owner.release();
    return ORBGPU_OK;
}

extern "C" void orbgpu_voc_destroy(orbgpu_voc *v)
{
    if (!v) return;
    cudaSetDevice(v->device);
    cudaFree(v->node_desc); cudaFree(v->child_offsets); cudaFree(v->child_ids); cudaFree(v->weight); cudaFree(v->word_id);
    delete v;
}

extern "C" int orbgpu_transform(orbgpu_ctx *ctx, const orbgpu_voc *voc, orbgpu_frame *f, int32_t levelsup, int32_t store_featvec,
                                uint32_t *word_id, uint32_t *node_id, double *weight)
{
    ARG_TRY(ctx && voc && f);
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    const int n = f->n;
    if (n == 0) return ORBGPU_OK;
    rc = arena_reserve(ctx, align256((size_t)(n + 1) * 4) + 256);
    if (rc) return rc;
    int32_t *tmp_off = (int32_t *)arena_take(ctx, (size_t)(n + 1) * 4);
    voc_transform_kernel<<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>(n, f->desc, voc->node_desc, voc->child_offsets, voc->child_ids,
                                                                       voc->weight, voc->word_id, voc->L, levelsup, f->word_id,
                                                                       f->node_id, f->weight, ctx->d_counters);
    LAUNCH_COUNT(ctx);
    f->has_transform = true;
    if (store_featvec) {
        const size_t smem = (size_t)f->sort_cap * 8;
        if (smem <= 160 * 1024) {
            featvec_bow_build_kernel<<<2, SORT_THREADS, smem, ctx->stream>>>(0, n, f->sort_cap, f->node_id, f->word_id, f->weight, nullptr,
                                                                           f->fv_node_ids, f->fv_offsets, f->fv_features, f->bow_words, tmp_off,
                                                                           f->bow_values, f->fv_meta);
            LAUNCH_COUNT(ctx);
        } else {
            for (int job = 0; job < 2; job++) {
                featvec_bow_build_kernel<<<1, SORT_THREADS, 0, ctx->stream>>>(job, n, f->sort_cap, f->node_id, f->word_id, f->weight, f->sort_keys,
                                                                            f->fv_node_ids, f->fv_offsets, f->fv_features, f->bow_words, tmp_off,
                                                                            f->bow_values, f->fv_meta);
                LAUNCH_COUNT(ctx);
            }
        }
    }
    CU_TRY(cudaGetLastError());
    int32_t meta[4] = {0, 0, 0, 0};
    const OutPiece out[4] = {{word_id, f->word_id, (size_t)n * 4}, {node_id, f->node_id, (size_t)n * 4}, {weight, f->weight, (size_t)n * 8},
                             {store_featvec ? meta : nullptr, f->fv_meta, sizeof(meta)}};
    rc = ctx_download(ctx, out, 4);
    if (rc) return rc;
    if (store_featvec) {
        f->fv_n_nodes = meta[0];
        f->fv_total = meta[1];
        f->fv_max_node = meta[2];
        f->bow_n = meta[3];
    }
    return ORBGPU_OK;
}

// TemplatedVocabulary::transform(const vector<TDescriptor>&, BowVector&, FeatureVector&, int levelsup) (TemplatedVocabulary.h:
// 1127-1194) for a bare list of descriptors, host pointers in / host pointers out, ONE synchronisation: what the drop-in
// vocabulary wrapper (include/orbmatch_b200/ORBVocabulary.hpp) calls from Frame::ComputeBoW / KeyFrame::ComputeBoW.
extern "C" int orbgpu_transform_descriptors(orbgpu_ctx *ctx, const orbgpu_voc *voc, int32_t n, const uint8_t *desc, int32_t levelsup,
                                            int32_t *n_words, uint32_t *words, double *values, int32_t *n_nodes, uint32_t *node_ids,
                                            int32_t *offsets, uint32_t *features)
{
    ARG_TRY(ctx && voc && n >= 0 && n_words && n_nodes && (n == 0 || (desc && words && values && node_ids && offsets && features)));
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    *n_words = 0;
    *n_nodes = 0;
    if (offsets) offsets[0] = 0;
    if (n == 0) return ORBGPU_OK;
    const size_t N = (size_t)n;
    int cap = 1;
    while (cap < n) cap <<= 1;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o_desc = take(N * 32), o_w = take(N * 4), o_nid = take(N * 4), o_wt = take(N * 8), o_fvn = take(N * 4), o_fvo = take((N + 1) * 4),
                 o_fvf = take(N * 4), o_bw = take(N * 4), o_bv = take(N * 8), o_tmp = take((N + 1) * 4), o_sk = take((size_t)cap * 8),
                 o_meta = take(64);
    rc = arena_reserve(ctx, off + 256);
    if (rc) return rc;
    char *D = (char *)arena_take(ctx, off);
    if (!D) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    CU_TRY(cudaMemcpyAsync(D + o_desc, desc, N * 32, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t *d_w = (uint32_t *)(D + o_w), *d_nid = (uint32_t *)(D + o_nid);
    double *d_wt = (double *)(D + o_wt);
    voc_transform_kernel<<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>(n, (const uint4 *)(D + o_desc), voc->node_desc, voc->child_offsets,
                                                                       voc->child_ids, voc->weight, voc->word_id, voc->L, levelsup, d_w, d_nid,
                                                                       d_wt, ctx->d_counters);
    LAUNCH_COUNT(ctx);
    const size_t smem = (size_t)cap * 8;
    if (smem <= 160 * 1024) {
        featvec_bow_build_kernel<<<2, SORT_THREADS, smem, ctx->stream>>>(0, n, cap, d_nid, d_w, d_wt, nullptr, (uint32_t *)(D + o_fvn),
                                                                       (int32_t *)(D + o_fvo), (uint32_t *)(D + o_fvf), (uint32_t *)(D + o_bw),
                                                                       (int32_t *)(D + o_tmp), (double *)(D + o_bv), (int32_t *)(D + o_meta));
        LAUNCH_COUNT(ctx);
    } else {
        for (int job = 0; job < 2; job++) {
            featvec_bow_build_kernel<<<1, SORT_THREADS, 0, ctx->stream>>>(job, n, cap, d_nid, d_w, d_wt, (unsigned long long *)(D + o_sk),
                                                                        (uint32_t *)(D + o_fvn), (int32_t *)(D + o_fvo), (uint32_t *)(D + o_fvf),
                                                                        (uint32_t *)(D + o_bw), (int32_t *)(D + o_tmp), (double *)(D + o_bv),
                                                                        (int32_t *)(D + o_meta));
            LAUNCH_COUNT(ctx);
        }
    }
    CU_TRY(cudaGetLastError());
    int32_t meta[4] = {0, 0, 0, 0}; // n_nodes, total features, max node size, n_words
    const OutPiece out[6] = {{meta, D + o_meta, sizeof(meta)}, {words, D + o_bw, N * 4}, {values, D + o_bv, N * 8},
                             {node_ids, D + o_fvn, N * 4}, {offsets, D + o_fvo, (N + 1) * 4}, {features, D + o_fvf, N * 4}};
    rc = ctx_download(ctx, out, 6);
    if (rc) return rc;
    *n_nodes = meta[0];
    *n_words = meta[3];
    return ORBGPU_OK;
}

extern "C" int orbgpu_bowvector_download(orbgpu_ctx *ctx, const orbgpu_frame *f, int32_t *n_words, uint32_t *words, double *values)
{
    ARG_TRY(ctx && f && n_words);
    CU_TRY(cudaSetDevice(ctx->device));
    *n_words = f->bow_n;
    const OutPiece out[2] = {{words, f->bow_words, (size_t)f->bow_n * 4}, {values, f->bow_values, (size_t)f->bow_n * 8}};
    return ctx_download(ctx, out, 2);
}

extern "C" int orbgpu_featvec_download(orbgpu_ctx *ctx, const orbgpu_frame *f, int32_t *n_nodes, uint32_t *node_ids, int32_t *offsets,
                                       uint32_t *features)
{
    ARG_TRY(ctx && f && n_nodes);
    CU_TRY(cudaSetDevice(ctx->device));
    *n_nodes = f->fv_n_nodes;
    const OutPiece out[3] = {{node_ids, f->fv_node_ids, (size_t)f->fv_n_nodes * 4}, {offsets, f->fv_offsets, (size_t)(f->fv_n_nodes + 1) * 4},
                             {features, f->fv_features, (size_t)f->fv_total * 4}};
    return ctx_download(ctx, out, 3);
}

int voc_device_init() { return set_max_dyn_smem(featvec_bow_build_kernel); }
