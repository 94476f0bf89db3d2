// bowdb.cu -- KeyFrameDatabase candidate scoring: the query BowVector against the BowVectors of
// all key frames at once.
//   common words  KeyFrameDatabase.cc:928-943 (the inverted-file walk counts, per key frame, the words it shares with the
//                 query: mnRelocWords / mnLoopWords / mnMergeWords)
//   score         L1Scoring::score (Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:23-68): merge-join of the two word-sorted
//                 vectors, score += |vi - wi| - |vi| - |wi| over the common words in ascending word order (double), then
//                 -score / 2
// The key-frame BowVectors live in HBM as one CSR (the device form of the database the inverted file indexes).  One thread
// per key frame walks its list against the query held in shared memory: the double sum must run in word order to be
// bit-exact, so it is kept sequential per key frame and the parallelism is across key frames.  The candidate policy on
// top (minCommonWords = 0.8 * max, covisibility accumulation, :949-1031) is control plane and stays with the caller.
#include <cstring>

#include "internal.cuh"

struct orbgpu_bowdb {
    int device = 0;
    int n_kf = 0;
    int64_t total = 0;
    int32_t *offsets = nullptr; // [n_kf+1]
    uint32_t *words = nullptr;  // [total] ascending inside a key frame
    double *values = nullptr;   // [total]
};

namespace {

constexpr int BOW_THREADS = 128;

__global__ void __launch_bounds__(BOW_THREADS)
bow_score_l1_kernel(int n_kf, const int32_t *__restrict__ offsets, const uint32_t *__restrict__ words, const double *__restrict__ values,
                    int nq, const uint32_t *__restrict__ q_words, const double *__restrict__ q_values, int32_t *__restrict__ common,
                    double *__restrict__ scores)
{
    extern __shared__ unsigned char bow_smem[];
    double *sv = (double *)bow_smem;        // [nq]
    uint32_t *sw = (uint32_t *)(sv + nq);   // [nq]
    for (int i = threadIdx.x; i < nq; i += BOW_THREADS) {
        sv[i] = q_values[i];
        sw[i] = q_words[i];
    }
    __syncthreads();
    const int kf = blockIdx.x * BOW_THREADS + threadIdx.x;
    if (kf >= n_kf) return;
    int a = 0, b = offsets[kf];
    const int b_end = offsets[kf + 1];
    double score = 0.0;
    int n_common = 0;
    while (a < nq && b < b_end) { // ScoringObject.cpp:35-61 (lower_bound == advancing a sorted cursor)
        const uint32_t wa = sw[a], wb = words[b];
        if (wa == wb) {
            const double vi = sv[a], wi = values[b];
            score = __dadd_rn(score, __dsub_rn(__dsub_rn(fabs(__dsub_rn(vi, wi)), fabs(vi)), fabs(wi)));
            n_common++;
            a++; b++;
        } else if (wa < wb) a++;
        else b++;
    }
    common[kf] = n_common;
    scores[kf] = -score / 2.0;
}

} // namespace

extern "C" void orbgpu_bowdb_destroy(orbgpu_bowdb *d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    cudaFree(d->offsets); cudaFree(d->words); cudaFree(d->values);
    delete d;
}

extern "C" int orbgpu_bowdb_upload(orbgpu_ctx *ctx, const orbgpu_bowdb_host *h, orbgpu_bowdb **out)
{
    ARG_TRY(ctx && h && out && h->n_kf >= 0 && (h->n_kf == 0 || h->offsets));
    const int64_t total = h->n_kf ? h->offsets[h->n_kf] : 0;
    ARG_TRY(total >= 0 && (total == 0 || (h->words && h->values)));
    for (int i = 0; i < h->n_kf; i++) ARG_TRY(h->offsets[i] <= h->offsets[i + 1]);
    CU_TRY(cudaSetDevice(ctx->device));
    orbgpu_bowdb *d = new orbgpu_bowdb();
    OwnedHandle<orbgpu_bowdb, orbgpu_bowdb_destroy> owner(d);
    d->device = ctx->device; d->n_kf = h->n_kf; d->total = total;
    CU_TRY(cudaMalloc((void **)&d->offsets, (size_t)(h->n_kf + 1) * 4));
    CU_TRY(cudaMalloc((void **)&d->words, (size_t)(total > 0 ? total : 1) * 4));
    CU_TRY(cudaMalloc((void **)&d->values, (size_t)(total > 0 ? total : 1) * 8));
    if (h->n_kf) CU_TRY(cudaMemcpyAsync(d->offsets, h->offsets, (size_t)(h->n_kf + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (total) {
        CU_TRY(cudaMemcpyAsync(d->words, h->words, (size_t)total * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU_TRY(cudaMemcpyAsync(d->values, h->values, (size_t)total * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    *out = owner.release();
    return ORBGPU_OK;
}

extern "C" int orbgpu_bow_score_l1(orbgpu_ctx *ctx, const orbgpu_bowdb *db, int32_t nq_words, const uint32_t *q_words,
                                   const double *q_values, int32_t *common_words, double *scores)
{
    ARG_TRY(ctx && db && nq_words >= 0 && (nq_words == 0 || (q_words && q_values)));
    ARG_TRY(db->n_kf == 0 || (common_words && scores));
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    if (db->n_kf == 0) return ORBGPU_OK;
    const size_t smem = (size_t)nq_words * 12 + 16;
    if (smem > 200 * 1024) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "query BowVector too large for shared memory");
    const size_t qb = align256((size_t)nq_words * 8), qw = align256((size_t)nq_words * 4);
    const size_t ob = align256((size_t)db->n_kf * 8), oc = align256((size_t)db->n_kf * 4);
    rc = arena_reserve(ctx, qb + qw + ob + oc + 256);
    if (rc) return rc;
    double *d_qv = (double *)arena_take(ctx, qb + 8);
    uint32_t *d_qw = (uint32_t *)arena_take(ctx, qw + 4);
    double *d_sc = (double *)arena_take(ctx, ob);
    int32_t *d_cm = (int32_t *)arena_take(ctx, oc);
    if (!d_qv || !d_qw || !d_sc || !d_cm) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    if (nq_words) {
        CU_TRY(cudaMemcpyAsync(d_qv, q_values, (size_t)nq_words * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU_TRY(cudaMemcpyAsync(d_qw, q_words, (size_t)nq_words * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (smem > ORBGPU_SMEM_OPTIN - 1024) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "query BowVector too large for the shared-memory staging");
    bow_score_l1_kernel<<<(db->n_kf + BOW_THREADS - 1) / BOW_THREADS, BOW_THREADS, smem, ctx->stream>>>(
        db->n_kf, db->offsets, db->words, db->values, nq_words, d_qw, d_qv, d_cm, d_sc);
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    const OutPiece out[2] = {{common_words, d_cm, (size_t)db->n_kf * 4}, {scores, d_sc, (size_t)db->n_kf * 8}};
    return ctx_download(ctx, out, 2);
}

int bowdb_device_init() { return set_max_dyn_smem(bow_score_l1_kernel); }
