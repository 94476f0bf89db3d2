// knn2.cu -- brute-force 2-NN Hamming search with ratio test (config C5) and the batched
// DescriptorDistance primitive (ORBmatcher.cc:2388-2408).
//
// Semantics (the inner loop of SearchByBoW, ORBmatcher.cc:319-355 + accept rule :392-395):
//   best   = FIRST database index attaining the minimum distance (strict '<' update),
//   second = second smallest distance of the multiset,
//   match  = best_idx if best <= th_low && (float)best < nnratio*(float)second else -1.
// A lexicographic (dist, index) top-2 reduction reproduces this exactly, so every engine
// packs key = dist<<22 | chunk-relative index and keeps the two smallest keys.
//
// Engines (orbgpu_knn2_set_engine):
//   1  LOP3+POPC CUDA-core kernel: queries in registers, database tiles staged in shared memory
//      with cp.async double buffering, broadcast LDS.128 reads.
//   2  mma.sync m16n8k256 b1 and.popc (ptxas lowers it to IMMA on sm_100a; kept for the bake-off
//      BASELINE.json asks for).
//   3  tcgen05 +-1 fp8 contraction with TMEM accumulators, one CTA per SM (knn2_tc.cu).
//   4  the same with cta_group::2 (SM pairs, M = 256, N = 256 per MMA) -- the default.
#include <algorithm>

#include "internal.cuh"

int knn2_tc_run(orbgpu_ctx *ctx, const orbgpu_db *db, int64_t nq, const uint4 *q, int32_t th_low, float nnratio, int32_t *best_idx,
                int32_t *best_dist, int32_t *second_dist, int32_t *match, bool two_cta); // knn2_tc.cu
bool knn2_tc_supported();

namespace {

constexpr int KNN_THREADS = 256;
constexpr int KNN_QPT = 4;       // queries per thread
constexpr int KNN_TILE = 256;    // database descriptors per shared-memory tile
constexpr int KNN_IDX_BITS = 22; // chunk-relative index bits in the packed key
constexpr uint32_t KNN_KEY_NONE = 0xFFFFFFFFu;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// grid = (query tiles, database splits).  Each thread owns KNN_QPT queries (8 registers each) and
// scans its split of the database from shared memory.
__global__ void __launch_bounds__(KNN_THREADS, 2)
knn2_popc_kernel(const uint4 *__restrict__ q, int64_t nq, const uint4 *__restrict__ db, int64_t nd, int64_t chunk,
                 uint64_t *__restrict__ part_best, uint32_t *__restrict__ part_second, int64_t part_stride)
{
    __shared__ uint4 tile[2][KNN_TILE * 2];
    const int t = threadIdx.x;
    const int64_t qbase = (int64_t)blockIdx.x * (KNN_THREADS * KNN_QPT);
    const int64_t d0 = (int64_t)blockIdx.y * chunk;
    const int64_t d1 = min(nd, d0 + chunk);

    uint4 qa[KNN_QPT], qb[KNN_QPT];
    uint32_t best[KNN_QPT], second[KNN_QPT];
#pragma unroll
    for (int i = 0; i < KNN_QPT; i++) {
        const int64_t qi = qbase + t + (int64_t)i * KNN_THREADS;
        if (qi < nq) {
            qa[i] = q[2 * qi];
            qb[i] = q[2 * qi + 1];
        } else {
            qa[i] = make_uint4(0, 0, 0, 0);
            qb[i] = make_uint4(0, 0, 0, 0);
        }
        best[i] = KNN_KEY_NONE;
        second[i] = KNN_KEY_NONE;
    }

    const int ntiles = (int)((d1 - d0 + KNN_TILE - 1) / KNN_TILE);
    auto prefetch = [&](int it, int buf) {
        const int64_t di = d0 + (int64_t)it * KNN_TILE + t;
        if (di < d1) {
            cp_async16(&tile[buf][2 * t], &db[2 * di]);
            cp_async16(&tile[buf][2 * t + 1], &db[2 * di + 1]);
        }
    };
    if (ntiles > 0) prefetch(0, 0);
    cp_async_commit();
    for (int it = 0; it < ntiles; it++) {
        const int buf = it & 1;
        if (it + 1 < ntiles) prefetch(it + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const int cnt = (int)min((int64_t)KNN_TILE, d1 - d0 - (int64_t)it * KNN_TILE);
        const uint32_t rel0 = (uint32_t)it * KNN_TILE;
#pragma unroll 4
        for (int j = 0; j < cnt; j++) {
            const uint4 da = tile[buf][2 * j], dbb = tile[buf][2 * j + 1];
#pragma unroll
            for (int i = 0; i < KNN_QPT; i++) {
                const uint32_t dist = (uint32_t)ham256(qa[i], qb[i], da, dbb);
                const uint32_t key = (dist << KNN_IDX_BITS) | (rel0 + j);
                const uint32_t m = max(best[i], key);
                best[i] = min(best[i], key);
                second[i] = min(second[i], m);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < KNN_QPT; i++) {
        const int64_t qi = qbase + t + (int64_t)i * KNN_THREADS;
        if (qi < nq) {
            uint64_t kb = ~0ull;
            if (best[i] != KNN_KEY_NONE)
                kb = ((uint64_t)(best[i] >> KNN_IDX_BITS) << 32) | (uint64_t)(d0 + (best[i] & ((1u << KNN_IDX_BITS) - 1)));
            part_best[(int64_t)blockIdx.y * part_stride + qi] = kb;
            part_second[(int64_t)blockIdx.y * part_stride + qi] = (second[i] == KNN_KEY_NONE) ? 0xFFFFu : (second[i] >> KNN_IDX_BITS);
        }
    }
}

// ---- engine 2: mma.sync m16n8k256 b1 and.popc -------------------------------------------
// Each warp owns KNN_MT m16 query tiles (A fragments in registers) and sweeps its database split
// 8 descriptors at a time.  hamming = popc(a) + popc(b) - 2*popc(a&b).
constexpr int MMA_MT = 2;          // m16 tiles per warp -> 32 queries per warp
constexpr int MMA_WARPS = 8;
constexpr int MMA_TILE = 256;      // database descriptors per shared tile

__device__ __forceinline__ void mma_b1_and_popc(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(MMA_WARPS * 32, 2)
knn2_mma_b1_kernel(const uint32_t *__restrict__ q, int64_t nq, const uint32_t *__restrict__ db, int64_t nd, int64_t chunk,
                   uint64_t *__restrict__ part_best, uint32_t *__restrict__ part_second, int64_t part_stride)
{
    __shared__ uint32_t tile[MMA_TILE * 8];
    __shared__ uint16_t tile_pop[MMA_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int64_t qbase = ((int64_t)blockIdx.x * MMA_WARPS + warp) * (16 * MMA_MT);
    const int64_t d0 = (int64_t)blockIdx.y * chunk;
    const int64_t d1 = min(nd, d0 + chunk);

    uint32_t a[MMA_MT][4];
    int pa[MMA_MT][2];
    uint32_t best[MMA_MT][2], second[MMA_MT][2]; // rows g and g+8: this thread sees cols 2tq,2tq+1 of each n8 tile
#pragma unroll
    for (int m = 0; m < MMA_MT; m++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int64_t row = qbase + m * 16 + g + 8 * h;
            uint32_t w0 = 0, w1 = 0;
            if (row < nq) {
                w0 = q[row * 8 + tq];
                w1 = q[row * 8 + 4 + tq];
            }
            a[m][h] = w0;     // a0/a1: k in [32*tq, 32*tq+32)
            a[m][h + 2] = w1; // a2/a3: k in [128+32*tq, ...)
            int p = __popc(w0) + __popc(w1);
            p += __shfl_xor_sync(FULL_MASK, p, 1);
            p += __shfl_xor_sync(FULL_MASK, p, 2);
            pa[m][h] = p;
            best[m][h] = KNN_KEY_NONE;
            second[m][h] = KNN_KEY_NONE;
        }
    }
    const int ntiles = (int)((d1 - d0 + MMA_TILE - 1) / MMA_TILE);
    for (int it = 0; it < ntiles; it++) {
        __syncthreads();
        const int64_t tb = d0 + (int64_t)it * MMA_TILE;
        const int cnt = (int)min((int64_t)MMA_TILE, d1 - tb);
        for (int i = threadIdx.x; i < MMA_TILE * 8; i += blockDim.x) tile[i] = (i < cnt * 8) ? db[tb * 8 + i] : 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < MMA_TILE; i += blockDim.x) {
            int p = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) p += __popc(tile[i * 8 + w]);
            tile_pop[i] = (uint16_t)p;
        }
        __syncthreads();
        for (int n0 = 0; n0 < cnt; n0 += 8) {
            const uint32_t b0 = tile[(n0 + g) * 8 + tq], b1 = tile[(n0 + g) * 8 + 4 + tq];
            const int c0 = n0 + 2 * tq, c1 = c0 + 1;
            const int pb0 = tile_pop[c0], pb1 = tile_pop[c1];
            const uint32_t rel0 = (uint32_t)it * MMA_TILE + c0, rel1 = rel0 + 1;
#pragma unroll
            for (int m = 0; m < MMA_MT; m++) {
                int d[4] = {0, 0, 0, 0};
                mma_b1_and_popc(d, a[m], b0, b1);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (c0 < cnt) {
                        const uint32_t dist = (uint32_t)(pa[m][h] + pb0 - 2 * d[2 * h]);
                        const uint32_t key = (dist << KNN_IDX_BITS) | rel0;
                        const uint32_t mx = max(best[m][h], key);
                        best[m][h] = min(best[m][h], key);
                        second[m][h] = min(second[m][h], mx);
                    }
                    if (c1 < cnt) {
                        const uint32_t dist = (uint32_t)(pa[m][h] + pb1 - 2 * d[2 * h + 1]);
                        const uint32_t key = (dist << KNN_IDX_BITS) | rel1;
                        const uint32_t mx = max(best[m][h], key);
                        best[m][h] = min(best[m][h], key);
                        second[m][h] = min(second[m][h], mx);
                    }
                }
            }
        }
    }
    // merge the 4 threads of a quad (they hold disjoint column subsets of the same rows)
#pragma unroll
    for (int m = 0; m < MMA_MT; m++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t b = best[m][h], s = second[m][h];
#pragma unroll
            for (int off = 1; off <= 2; off <<= 1) {
                const uint32_t ob = __shfl_xor_sync(FULL_MASK, b, off), os = __shfl_xor_sync(FULL_MASK, s, off);
                const uint32_t nb = min(b, ob);
                s = min(min(s, os), max(b, ob));
                b = nb;
            }
            const int64_t row = qbase + m * 16 + g + 8 * h;
            if (tq == 0 && row < nq) {
                uint64_t kb = ~0ull;
                if (b != KNN_KEY_NONE) kb = ((uint64_t)(b >> KNN_IDX_BITS) << 32) | (uint64_t)(d0 + (b & ((1u << KNN_IDX_BITS) - 1)));
                part_best[(int64_t)blockIdx.y * part_stride + row] = kb;
                part_second[(int64_t)blockIdx.y * part_stride + row] = (s == KNN_KEY_NONE) ? 0xFFFFu : (s >> KNN_IDX_BITS);
            }
        }
}

// merges the per-split partial (best key, second dist) and applies the accept rule
__global__ void knn2_merge_kernel(const uint64_t *__restrict__ part_best, const uint32_t *__restrict__ part_second, int n_splits,
                                  int64_t part_stride, int64_t nq, int th_low, float nnratio, int32_t *__restrict__ best_idx,
                                  int32_t *__restrict__ best_dist, int32_t *__restrict__ second_dist, int32_t *__restrict__ match)
{
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    uint64_t b = ~0ull;
    int win = -1;
    for (int s = 0; s < n_splits; s++) {
        const uint64_t k = part_best[(int64_t)s * part_stride + qi];
        if (k < b) { b = k; win = s; }
    }
    uint32_t sec = 0xFFFFu;
    for (int s = 0; s < n_splits; s++) {
        const uint64_t k = part_best[(int64_t)s * part_stride + qi];
        const uint32_t cand = (s == win) ? part_second[(int64_t)s * part_stride + qi]
                                         : (k == ~0ull ? 0xFFFFu : (uint32_t)(k >> 32));
        sec = min(sec, cand);
    }
    int bd = 256, bi = -1, sd = 256; // initial values of the reference loop (ORBmatcher.cc:319-321)
    if (b != ~0ull) {
        bd = (int)(b >> 32);
        bi = (int)(b & 0xFFFFFFFFull);
    }
    if (sec < 256u) sd = (int)sec;
    int m = -1;
    if (bd <= th_low)
        if ((float)bd < __fmul_rn(nnratio, (float)sd)) m = bi;
    if (best_idx) best_idx[qi] = bi;
    if (best_dist) best_dist[qi] = bd;
    if (second_dist) second_dist[qi] = sd;
    if (match) match[qi] = m;
}

__global__ void descriptor_distance_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b, int64_t n, int32_t *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ham256(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}

} // namespace

extern "C" int orbgpu_knn2_set_engine(orbgpu_ctx *ctx, int32_t engine)
{
    ARG_TRY(ctx != nullptr && engine >= 0 && engine <= 4);
    if (engine >= 3 && !knn2_tc_supported())
        return orbgpu_fail(ORBGPU_ERR_INVALID, "tcgen05 engine not built into this library");
    ctx->knn_engine = engine;
    return ORBGPU_OK;
}

extern "C" int orbgpu_db_upload(orbgpu_ctx *ctx, int64_t nd, const uint8_t *db_desc, orbgpu_db **out)
{
    ARG_TRY(ctx && out && nd >= 0 && (nd == 0 || db_desc));
    CU_TRY(cudaSetDevice(ctx->device));
    orbgpu_db *d = new orbgpu_db();
    d->device = ctx->device;
    d->nd = nd;
    d->capacity = nd;
    d->owned = true;
    OwnedHandle<orbgpu_db, orbgpu_db_destroy> owner(d);
    void *p = nullptr;
    CU_TRY(cudaMalloc(&p, std::max<int64_t>(nd, 1) * 32));
    d->desc = (const uint4 *)p;
    if (nd > 0) CU_TRY(cudaMemcpyAsync(p, db_desc, nd * 32, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    *out = owner.release();
    return ORBGPU_OK;
}

// re-uploads descriptors into an owned database without reallocating (nd must not exceed the size it was created with)
extern "C" int orbgpu_db_update(orbgpu_ctx *ctx, orbgpu_db *db, int64_t nd, const uint8_t *db_desc)
{
    ARG_TRY(ctx && db && db->owned && nd >= 0 && nd <= db->capacity && (nd == 0 || db_desc));
    CU_TRY(cudaSetDevice(ctx->device));
    if (db->up_stream) CU_TRY(cudaStreamSynchronize(db->up_stream)); // an earlier chunked upload nobody consumed
    db->up_pending = 0;
    if (nd > 0) CU_TRY(cudaMemcpyAsync((void *)db->desc, db_desc, nd * 32, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    db->nd = nd;
    {
        std::lock_guard<std::mutex> lock(db->x_mu);
        db->x_valid = false; // the expanded copy is rebuilt by the next tensor-engine search
    }
    return ORBGPU_OK;
}

int db_wait_upload(orbgpu_ctx *ctx, const orbgpu_db *db)
{
    if (db->up_pending > 0) {
        CU_TRY(cudaStreamWaitEvent(ctx->stream, db->up_ev[db->up_pending - 1], 0)); // the chunks are copied in order on one stream
        db->up_pending = 0;
    }
    return ORBGPU_OK;
}

// starts the H2D copy of new descriptors into an owned database in chunks of whole tensor-engine splits (131 072 rows = 4 MiB) on the
// database's own stream; does not wait.  The caller's buffer must stay valid until a search that consumed the upload has synchronised.
static int db_begin_chunked_upload(orbgpu_ctx *ctx, orbgpu_db *db, int64_t nd, const uint8_t *db_desc)
{
    std::lock_guard<std::mutex> lock(db->x_mu);
    if (!db->up_stream) {
        CU_TRY(cudaStreamCreateWithFlags(&db->up_stream, cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&db->up_gate, cudaEventDisableTiming));
        for (int c = 0; c < 16; c++) CU_TRY(cudaEventCreateWithFlags(&db->up_ev[c], cudaEventDisableTiming));
    }
    // nothing enqueued so far on the caller's stream may still be reading the old descriptors when the copy lands
    CU_TRY(cudaEventRecord(db->up_gate, ctx->stream));
    CU_TRY(cudaStreamWaitEvent(db->up_stream, db->up_gate, 0));
    // three chunks of 1/8, 1/4 and 5/8 of the splits: the search of a chunk covers the copy of the next whether the call is
    // compute-bound (one GPU: 128 MiB cross PCIe in ~1 ms, the search takes 165) or closer to copy-bound (a query shard of 1/8 on each
    // of 8 GPUs sharing the host's memory system: ~5 ms against 20), and every extra launch costs a partial last wave
    const int64_t split_rows = 131072;
    const int64_t n_split = (nd + split_rows - 1) / split_rows;
    int chunks = 0;
    for (int64_t done = 0; done < n_split; chunks++) {
        int64_t take = chunks == 0 ? std::max<int64_t>(1, n_split / 8) : (chunks == 1 ? std::max<int64_t>(1, n_split / 4) : n_split - done);
        take = std::min(take, n_split - done);
        const int64_t r0 = done * split_rows, r1 = std::min(nd, (done + take) * split_rows);
        CU_TRY(cudaMemcpyAsync((char *)db->desc + r0 * 32, db_desc + r0 * 32, (r1 - r0) * 32, cudaMemcpyHostToDevice, db->up_stream));
        CU_TRY(cudaEventRecord(db->up_ev[chunks], db->up_stream));
        db->up_row_end[chunks] = r1;
        done += take;
    }
    db->nd = nd;
    db->x_valid = false;
    db->up_pending = chunks;
    return ORBGPU_OK;
}

// a database that borrows the caller's device memory (orbgpu_db_from_dev) cannot see its contents change: the caller says so
extern "C" int orbgpu_db_invalidate(orbgpu_db *db)
{
    ARG_TRY(db);
    std::lock_guard<std::mutex> lock(db->x_mu);
    db->x_valid = false;
    return ORBGPU_OK;
}

extern "C" int orbgpu_db_from_dev(orbgpu_ctx *ctx, int64_t nd, const void *db_desc_dev, orbgpu_db **out)
{
    ARG_TRY(ctx && out && nd >= 0 && (nd == 0 || db_desc_dev));
    ARG_TRY(((uintptr_t)db_desc_dev & 15) == 0);
    orbgpu_db *d = new orbgpu_db();
    d->device = ctx->device;
    d->nd = nd;
    d->owned = false;
    d->desc = (const uint4 *)db_desc_dev;
    *out = d;
    return ORBGPU_OK;
}

extern "C" void orbgpu_db_destroy(orbgpu_db *db)
{
    if (!db) return;
    cudaSetDevice(db->device);
    if (db->owned && db->desc) cudaFree((void *)db->desc);
    if (db->x_desc) cudaFree(db->x_desc);
    if (db->x_ready) cudaEventDestroy(db->x_ready);
    if (db->up_stream) {
        cudaStreamSynchronize(db->up_stream);
        cudaStreamDestroy(db->up_stream);
        cudaEventDestroy(db->up_gate);
        for (int c = 0; c < 16; c++) cudaEventDestroy(db->up_ev[c]);
    }
    delete db;
}

struct KnnPlan {
    int engine;
    int64_t qtiles, splits, chunk, stride, cap_splits;
    size_t part_bytes;
};

static KnnPlan knn2_plan(const orbgpu_ctx *ctx, int64_t nq, int64_t nd)
{
    KnnPlan p;
    p.engine = ctx->knn_engine;
    if (p.engine == 0) p.engine = knn2_tc_supported() ? 4 : 1; // auto: 2-CTA tcgen05 engine
    // the tensor engine works on 128-query x 256-database tiles: tiny problems go to the POPC kernel
    if (p.engine >= 3 && (nd < 1024 || nq < 128)) p.engine = 1;
    const int64_t per_tile = (p.engine == 2) ? MMA_WARPS * 16 * MMA_MT : KNN_THREADS * KNN_QPT;
    p.qtiles = std::max<int64_t>(1, (nq + per_tile - 1) / per_tile);
    // enough CTAs for ~4 waves over the SMs; chunk-relative indices must fit the packed key
    int64_t splits = std::max<int64_t>(1, (4LL * ctx->sm_count + p.qtiles - 1) / p.qtiles);
    splits = std::min<int64_t>(splits, std::max<int64_t>(1, nd / 2048));
    splits = std::max<int64_t>(splits, (nd + (1LL << KNN_IDX_BITS) - 1) >> KNN_IDX_BITS);
    splits = std::max<int64_t>(splits, 1);
    int64_t chunk = (nd + splits - 1) / splits;
    chunk = std::max<int64_t>(((chunk + 255) / 256) * 256, 256);
    p.chunk = chunk;
    p.splits = std::max<int64_t>(1, (nd + chunk - 1) / chunk);
    p.stride = std::max<int64_t>(nq, 1);
    p.cap_splits = p.splits;
    p.part_bytes = align256(p.cap_splits * p.stride * 8) + align256(p.cap_splits * p.stride * 4);
    return p;
}

// q, outputs: device pointers.  The arena must already hold plan.part_bytes free bytes.
static int knn2_launch(orbgpu_ctx *ctx, const orbgpu_db *db, int64_t nq, const uint4 *q, const KnnPlan &p, int32_t th_low,
                       float nnratio, int32_t *best_idx, int32_t *best_dist, int32_t *second_dist, int32_t *match)
{
    const int64_t nd = db->nd;
    uint64_t *part_best = (uint64_t *)arena_take(ctx, p.cap_splits * p.stride * 8);
    uint32_t *part_second = (uint32_t *)arena_take(ctx, p.cap_splits * p.stride * 4);
    if (!part_best || !part_second) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    ctx->last_comparisons = nq * nd;
    if (nq == 0) return ORBGPU_OK;
    if (p.engine >= 3 && nd > 0) // tensor engines: expansion + tcgen05 search + their own merge (4 = cta_group::2)
        return knn2_tc_run(ctx, db, nq, q, th_low, nnratio, best_idx, best_dist, second_dist, match, p.engine == 4);
    int n_splits = (int)p.splits;
    {
        int rc = db_wait_upload(ctx, db);
        if (rc) return rc;
    }
    if (nd > 0) {
        if (p.engine == 1) {
            dim3 grid((unsigned)p.qtiles, (unsigned)p.splits);
            knn2_popc_kernel<<<grid, KNN_THREADS, 0, ctx->stream>>>(q, nq, db->desc, nd, p.chunk, part_best, part_second, p.stride);
            LAUNCH_COUNT(ctx);
        } else if (p.engine == 2) {
            dim3 grid((unsigned)p.qtiles, (unsigned)p.splits);
            knn2_mma_b1_kernel<<<grid, MMA_WARPS * 32, 0, ctx->stream>>>((const uint32_t *)q, nq, (const uint32_t *)db->desc, nd,
                                                                         p.chunk, part_best, part_second, p.stride);
            LAUNCH_COUNT(ctx);
        }
        CU_TRY(cudaGetLastError());
    } else {
        n_splits = 0;
    }
    knn2_merge_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, ctx->stream>>>(part_best, part_second, n_splits, p.stride, nq, th_low,
                                                                            nnratio, best_idx, best_dist, second_dist, match);
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    return ORBGPU_OK;
}

extern "C" int orbgpu_knn2_ratio_dev(orbgpu_ctx *ctx, const orbgpu_db *db, int64_t nq, const void *q_desc_dev, int32_t th_low,
                                     float nnratio, int32_t *best_idx_dev, int32_t *best_dist_dev, int32_t *second_dist_dev,
                                     int32_t *match_dev)
{
    ARG_TRY(ctx && db && nq >= 0 && (nq == 0 || q_desc_dev));
    ARG_TRY(((uintptr_t)q_desc_dev & 15) == 0);
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    const KnnPlan p = knn2_plan(ctx, nq, db->nd);
    rc = arena_reserve(ctx, p.part_bytes);
    if (rc) return rc;
    return knn2_launch(ctx, db, nq, (const uint4 *)q_desc_dev, p, th_low, nnratio, best_idx_dev, best_dist_dev, second_dist_dev,
                       match_dev);
}

// host-pointer search; up_desc != nullptr: the database is re-uploaded first (chunked, on its own stream) -- AFTER the query copy has
// been enqueued: the H2D engine serves the streams in submission order, and the first search launch needs the queries, not the last chunk
static int knn2_ratio_host(orbgpu_ctx *ctx, orbgpu_db *db, int64_t up_nd, const uint8_t *up_desc, int64_t nq, const uint8_t *q_desc,
                           int32_t th_low, float nnratio, int32_t *best_idx, int32_t *best_dist, int32_t *second_dist, int32_t *match)
{
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    const KnnPlan p = knn2_plan(ctx, nq, up_desc ? up_nd : db->nd);
    const size_t qbytes = align256(nq * 32), rbytes = align256(nq * 4);
    rc = arena_reserve(ctx, p.part_bytes + qbytes + 4 * rbytes);
    if (rc) return rc;
    char *base = (char *)arena_take(ctx, qbytes + 4 * rbytes);
    if (!base) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    int32_t *dbi = (int32_t *)(base + qbytes), *dbd = (int32_t *)(base + qbytes + rbytes),
            *dsd = (int32_t *)(base + qbytes + 2 * rbytes), *dmt = (int32_t *)(base + qbytes + 3 * rbytes);
    CU_TRY(cudaMemcpyAsync(base, q_desc, nq * 32, cudaMemcpyHostToDevice, ctx->stream));
    if (up_desc) {
        rc = db_begin_chunked_upload(ctx, db, up_nd, up_desc);
        if (rc) return rc;
    }
    rc = knn2_launch(ctx, db, nq, (const uint4 *)base, p, th_low, nnratio, dbi, dbd, dsd, dmt);
    if (rc) return rc;
    if (best_idx) CU_TRY(cudaMemcpyAsync(best_idx, dbi, nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (best_dist) CU_TRY(cudaMemcpyAsync(best_dist, dbd, nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (second_dist) CU_TRY(cudaMemcpyAsync(second_dist, dsd, nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (match) CU_TRY(cudaMemcpyAsync(match, dmt, nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return ORBGPU_OK;
}

extern "C" int orbgpu_knn2_ratio(orbgpu_ctx *ctx, const orbgpu_db *db, int64_t nq, const uint8_t *q_desc, int32_t th_low,
                                 float nnratio, int32_t *best_idx, int32_t *best_dist, int32_t *second_dist, int32_t *match)
{
    ARG_TRY(ctx && db && nq >= 0 && (nq == 0 || q_desc));
    if (nq == 0) return ctx_begin(ctx);
    return knn2_ratio_host(ctx, const_cast<orbgpu_db *>(db), 0, nullptr, nq, q_desc, th_low, nnratio, best_idx, best_dist, second_dist, match);
}

// orbgpu_db_update + orbgpu_knn2_ratio in one call, overlapped: the new descriptors are copied in chunks on the database's own stream
// and the tensor engine searches each chunk as soon as it has arrived (what a SearchByNN(queries, database) on host matrices does).
// Returns after the results are on the host; the caller's buffers are free again.
extern "C" int orbgpu_knn2_ratio_update(orbgpu_ctx *ctx, orbgpu_db *db, int64_t nd, const uint8_t *db_desc, int64_t nq, const uint8_t *q_desc,
                                        int32_t th_low, float nnratio, int32_t *best_idx, int32_t *best_dist, int32_t *second_dist,
                                        int32_t *match)
{
    ARG_TRY(ctx && db && db->owned && nd >= 0 && nd <= db->capacity && (nd == 0 || db_desc) && nq >= 0 && (nq == 0 || q_desc));
    CU_TRY(cudaSetDevice(ctx->device));
    if (nq == 0) { // nothing to search: plain update
        return orbgpu_db_update(ctx, db, nd, db_desc);
    }
    const int rc = knn2_ratio_host(ctx, db, nd, db_desc, nq, q_desc, th_low, nnratio, best_idx, best_dist, second_dist, match);
    if (rc && db->up_stream) cudaStreamSynchronize(db->up_stream); // the caller's buffer is free again on every path
    return rc;
}

extern "C" int orbgpu_descriptor_distance(orbgpu_ctx *ctx, int64_t n, const uint8_t *a, const uint8_t *b, int32_t *out)
{
    ARG_TRY(ctx && n >= 0 && (n == 0 || (a && b && out)));
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    if (n == 0) return ORBGPU_OK;
    rc = arena_reserve(ctx, 2 * align256(n * 32) + align256(n * 4));
    if (rc) return rc;
    uint4 *da = (uint4 *)arena_take(ctx, n * 32), *dbb = (uint4 *)arena_take(ctx, n * 32);
    int32_t *dout = (int32_t *)arena_take(ctx, n * 4);
    CU_TRY(cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(dbb, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    descriptor_distance_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(da, dbb, n, dout);
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->last_comparisons = n;
    return ORBGPU_OK;
}
