// frame.cu -- device copy of a Frame / KeyFrame feature set, the CSR cell grid and the
// batched area query.
//
// Replaces Frame::AssignFeaturesToGrid (Frame.cc:469-507) + PosInGrid (:973-989) and
// Frame/KeyFrame::GetFeaturesInArea (Frame.cc:868-962, KeyFrame.cc:859-907).
//
// HBM layout (one slab per frame): features in feature order (descriptors as packed uint4
// pairs, float2 keypoints, octave, angle) plus a CELL-ORDERED copy -- `items` (x, y,
// octave|iy<<16, feature id) and `desc_sorted` -- so that a window query reads ONE contiguous
// span of memory: cells are numbered ix*rows+iy, hence the columns ix0..ix1 of a window are
// adjacent and already in the reference's iteration order (ix outer, iy inner, in-cell
// ascending feature id).  The y-range of the window is applied per item from the stored iy.
#include <algorithm>
#include <cstring>

#include <mutex>
#include <vector>

#include "internal.cuh"

namespace {

constexpr int GRID_THREADS = 1024;

// one block per frame
__global__ void __launch_bounds__(GRID_THREADS)
grid_build_kernel(int n, const float2 *__restrict__ xy, const int32_t *__restrict__ octave, const uint4 *__restrict__ desc,
                  float min_x, float min_y, float inv_w, float inv_h, int cols, int rows, int32_t *__restrict__ cell_of,
                  int32_t *__restrict__ cell_start, int32_t *__restrict__ cell_items, int4 *__restrict__ items,
                  uint4 *__restrict__ desc_sorted, int cof_in_smem)
{
    extern __shared__ int32_t sm[]; // counts[ncell] then cursor reuse; then (cof_in_smem) the features' cells [n]
    __shared__ int32_t warp_sums[32];
    const int ncell = cols * rows;
    const int t = threadIdx.x;
    // the ordered scatter below is one warp walking the features: from shared memory a step is ~100 cycles, from global
    // memory an L2 round trip
    int32_t *cof = cof_in_smem ? sm + ncell : cell_of;
    for (int c = t; c < ncell; c += GRID_THREADS) sm[c] = 0;
    __syncthreads();
    // PosInGrid (Frame.cc:973-989): round-half-away of (pt - min) * inv, reject outside the grid
    for (int i = t; i < n; i += GRID_THREADS) {
        const float2 p = xy[i];
        const int px = (int)roundf(__fmul_rn(__fsub_rn(p.x, min_x), inv_w));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(p.y, min_y), inv_h));
        int c = -1;
        if (!(px < 0 || px >= cols || py < 0 || py >= rows)) {
            c = px * rows + py;
            atomicAdd(&sm[c], 1);
        }
        cell_of[i] = c;
        if (cof_in_smem) cof[i] = c;
    }
    __syncthreads();
    // exclusive scan of the per-cell counts
    const int per = (ncell + GRID_THREADS - 1) / GRID_THREADS;
    int local = 0;
    for (int k = 0; k < per; k++) {
        const int c = t * per + k;
        if (c < ncell) local += sm[c];
    }
    int incl = local;
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(FULL_MASK, incl, off);
        if ((t & 31) >= off) incl += v;
    }
    if ((t & 31) == 31) warp_sums[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        int w = warp_sums[t];
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(FULL_MASK, w, off);
            if (t >= off) w += v;
        }
        warp_sums[t] = w;
    }
    __syncthreads();
    int run = incl - local + ((t >> 5) > 0 ? warp_sums[(t >> 5) - 1] : 0);
    for (int k = 0; k < per; k++) {
        const int c = t * per + k;
        if (c < ncell) {
            const int cnt = sm[c];
            cell_start[c] = run;
            sm[c] = run; // becomes the scatter cursor
            run += cnt;
        }
    }
    if (t == GRID_THREADS - 1) cell_start[ncell] = warp_sums[31];
    __syncthreads();
    // ordered scatter by ONE warp: features visited in ascending id so that in-cell order is
    // ascending feature id (push_back order of AssignFeaturesToGrid)
    if (t < 32) {
        for (int base = 0; base < n; base += 32) {
            const int i = base + t;
            const int c = (i < n) ? cof[i] : -1;
            const unsigned same = __match_any_sync(FULL_MASK, c);
            if (c >= 0) {
                const int rank = __popc(same & lanemask_lt());
                const int pos = sm[c] + rank;
                cell_items[pos] = i;
            }
            __syncwarp();
            if (c >= 0 && (t == 31 || (same >> (t + 1)) == 0)) sm[c] += __popc(same); // highest lane of the group advances the cursor
            __syncwarp();
        }
    }
    __syncthreads();
    const int n_in = cell_start[ncell];
    for (int s = t; s < n_in; s += GRID_THREADS) {
        const int i = cell_items[s];
        const float2 p = xy[i];
        const int iy = cof[i] % rows;
        items[s] = make_int4(__float_as_int(p.x), __float_as_int(p.y), (octave[i] & 0xffff) | (iy << 16), i);
        desc_sorted[2 * s] = desc[2 * i];
        desc_sorted[2 * s + 1] = desc[2 * i + 1];
    }
}

__global__ void area_count_kernel(FrameView f, int nq, const float *x, const float *y, const float *r, const int32_t *minl,
                                  const int32_t *maxl, int32_t *counts)
{
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    const int cnt = window_scan(f, x[q], y[q], r[q], minl[q], maxl[q], [](bool, int, int, int4) {});
    if (lane_id() == 0) counts[q] = cnt;
}

__global__ void exclusive_scan_kernel(const int32_t *in, int32_t *out, int n)
{
    // single block; n is small (number of queries)
    __shared__ int32_t carry;
    __shared__ int32_t wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = (i < n) ? in[i] : 0;
        int incl = v;
        for (int off = 1; off < 32; off <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, incl, off);
            if ((threadIdx.x & 31) >= off) incl += u;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = (threadIdx.x < (blockDim.x >> 5)) ? wsum[threadIdx.x] : 0;
            for (int off = 1; off < 32; off <<= 1) {
                const int u = __shfl_up_sync(FULL_MASK, w, off);
                if (threadIdx.x >= off) w += u;
            }
            wsum[threadIdx.x] = w;
        }
        __syncthreads();
        const int prefix = carry + ((threadIdx.x >> 5) > 0 ? wsum[(threadIdx.x >> 5) - 1] : 0);
        if (i < n) out[i] = prefix + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = prefix + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

__global__ void area_fill_kernel(FrameView f, int nq, const float *x, const float *y, const float *r, const int32_t *minl,
                                 const int32_t *maxl, const int32_t *offsets, int32_t *out, int64_t cap)
{
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    const int64_t off = offsets[q];
    window_scan(f, x[q], y[q], r[q], minl[q], maxl[q], [&](bool ok, int pos, int, int4 it) {
        if (ok && off + pos < cap) out[off + pos] = it.w;
    });
}

} // namespace

// exported to the search kernels (same translation-unit-local template is re-declared there)
extern "C" int orbgpu_frame_n(const orbgpu_frame *f) { return f ? f->n : 0; }

// Frame slabs are recycled: the drop-in adapter uploads its frames on every call, and a cudaMalloc / cudaFree pair per frame
// costs more than the upload itself (cudaFree also synchronises the device).  Process-wide, per device, bounded.  A slab that is
// given back carries an event recorded on its owner's stream; the next upload that takes it makes ITS stream wait on that event,
// so neither side ever stops the device -- the other threads' streams (Tracking / LocalMapping / LoopClosing each own a context)
// keep running.
namespace {
struct SlabPool {
    struct Entry { int device; char *p; size_t bytes; cudaEvent_t ev; };
    std::mutex mu;
    std::vector<Entry> free_list;
    size_t held = 0;
};
SlabPool g_slabs;
constexpr size_t SLAB_POOL_BYTES = size_t(256) << 20;
constexpr size_t SLAB_POOL_ENTRIES = 32;

char *slab_take(int device, size_t need, size_t *got, cudaEvent_t *ev)
{
    std::lock_guard<std::mutex> lock(g_slabs.mu);
    int best = -1;
    for (int i = 0; i < (int)g_slabs.free_list.size(); i++) {
        const SlabPool::Entry &e = g_slabs.free_list[i];
        if (e.device != device || e.bytes < need || e.bytes > 2 * need + (size_t(1) << 20)) continue;
        if (best < 0 || e.bytes < g_slabs.free_list[best].bytes) best = i;
    }
    if (best < 0) return nullptr;
    const SlabPool::Entry e = g_slabs.free_list[best];
    g_slabs.free_list.erase(g_slabs.free_list.begin() + best);
    g_slabs.held -= e.bytes;
    *got = e.bytes;
    *ev = e.ev;
    return e.p;
}
bool slab_give(int device, char *p, size_t bytes, cudaEvent_t ev)
{
    std::lock_guard<std::mutex> lock(g_slabs.mu);
    if (g_slabs.free_list.size() >= SLAB_POOL_ENTRIES || g_slabs.held + bytes > SLAB_POOL_BYTES) return false;
    g_slabs.free_list.push_back({device, p, bytes, ev});
    g_slabs.held += bytes;
    return true;
}
} // namespace

extern "C" void orbgpu_frame_destroy(orbgpu_frame *f)
{
    if (!f) return;
    cudaSetDevice(f->device);
    if (f->slab) {
        // work that still reads the frame (device-pointer entry points do not synchronise) is ordered before the slab's next use
        // by an event on the owner's stream; a frame whose owner is gone has nothing in flight (orbgpu_destroy synchronised)
        cudaEvent_t ev = nullptr;
        bool pooled = false;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
            if (!ctx_record_event_if_alive(f->owner, ev)) { cudaEventDestroy(ev); ev = nullptr; }
            pooled = slab_give(f->device, f->slab, f->slab_bytes, ev);
        }
        if (!pooled) {
            if (ev) cudaEventDestroy(ev);
            ctx_sync_if_alive(f->owner);
            cudaFree(f->slab);
        }
    }
    delete f;
}

static int next_pow2(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

extern "C" int orbgpu_frame_upload(orbgpu_ctx *ctx, const orbgpu_frame_host *h, orbgpu_frame **out)
{
    ARG_TRY(ctx && h && out);
    ARG_TRY(h->n >= 0 && h->grid_cols > 0 && h->grid_rows > 0 && h->grid_cols * h->grid_rows <= 8192);
    ARG_TRY(h->n == 0 || (h->desc && h->kp_xy && h->octave && h->angle));
    ARG_TRY(h->n_levels > 0 && h->n_levels <= 64 && h->scale_factors && h->level_sigma2);
    ARG_TRY(h->fv_n_nodes >= 0 && h->fv_n_nodes <= h->n);
    for (int i = 0; i < h->n; i++) ARG_TRY(h->octave[i] >= 0 && h->octave[i] < h->n_levels); // indexes the scale tables
    CU_TRY(cudaSetDevice(ctx->device));
    const int n = h->n, nl = h->n_levels, ncell = h->grid_cols * h->grid_rows;
    const size_t N = (size_t)(n > 0 ? n : 1);
    const int fv_total = h->fv_n_nodes > 0 ? h->fv_offsets[h->fv_n_nodes] : 0;
    ARG_TRY(fv_total >= 0 && fv_total <= n);

    orbgpu_frame *f = new orbgpu_frame();
    OwnedHandle<orbgpu_frame, orbgpu_frame_destroy> owner(f);
    f->device = ctx->device;
    f->n = n;
    f->n_levels = nl;
    f->min_x = h->min_x; f->min_y = h->min_y; f->max_x = h->max_x; f->max_y = h->max_y;
    f->inv_w = h->grid_inv_w; f->inv_h = h->grid_inv_h;
    f->cols = h->grid_cols; f->rows = h->grid_rows;
    f->sort_cap = next_pow2((int)N);

    // slab layout: [uploaded part | device-only part]
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o_desc = take(N * 32), o_xy = take(N * 8), o_oct = take(N * 4), o_ang = take(N * 4), o_ur = take(N * 4), o_r0 = take(N * 4),
                 o_sf = take(nl * 4), o_s2 = take(nl * 4), o_fvn = take(N * 4), o_fvo = take((N + 1) * 4), o_fvf = take(N * 4);
    const size_t upload_bytes = off;
    const size_t o_cs = take((size_t)(ncell + 1) * 4), o_ci = take(N * 4), o_it = take(N * 16), o_ds = take(N * 32),
                 o_w = take(N * 4), o_nid = take(N * 4), o_wt = take(N * 8), o_bw = take(N * 4), o_bv = take(N * 8),
                 o_cof = take(N * 4), o_sk = take((size_t)f->sort_cap * 8), o_meta = take(64);
    f->slab_bytes = off;
    f->owner = ctx;
    cudaEvent_t reuse_ev = nullptr;
    f->slab = slab_take(ctx->device, off, &f->slab_bytes, &reuse_ev);
    if (!f->slab) {
        f->slab_bytes = off;
        CU_TRY(cudaMalloc(&f->slab, f->slab_bytes));
    } else if (reuse_ev) { // the previous user's work on this slab comes first -- on the device, without stopping the host
        const cudaError_t we = cudaStreamWaitEvent(ctx->stream, reuse_ev, 0);
        cudaEventDestroy(reuse_ev);
        CU_TRY(we);
    }
    char *S = f->slab;
    f->desc = (uint4 *)(S + o_desc); f->xy = (float2 *)(S + o_xy); f->octave = (int32_t *)(S + o_oct);
    f->angle = (float *)(S + o_ang); f->u_right = h->u_right ? (float *)(S + o_ur) : nullptr;
    f->scale_factors = (float *)(S + o_sf); f->level_sigma2 = (float *)(S + o_s2); f->rank0 = (int32_t *)(S + o_r0);
    f->fv_node_ids = (uint32_t *)(S + o_fvn); f->fv_offsets = (int32_t *)(S + o_fvo); f->fv_features = (uint32_t *)(S + o_fvf);
    f->cell_start = (int32_t *)(S + o_cs); f->cell_items = (int32_t *)(S + o_ci); f->items = (int4 *)(S + o_it);
    f->desc_sorted = (uint4 *)(S + o_ds); f->word_id = (uint32_t *)(S + o_w); f->node_id = (uint32_t *)(S + o_nid);
    f->weight = (double *)(S + o_wt); f->bow_words = (uint32_t *)(S + o_bw); f->bow_values = (double *)(S + o_bv);
    f->cell_of = (int32_t *)(S + o_cof); f->sort_keys = (unsigned long long *)(S + o_sk); f->fv_meta = (int32_t *)(S + o_meta);

    // pack the host arrays into pinned staging in slab layout -> ONE H2D copy
    int rc = stage_reserve(ctx, upload_bytes);
    if (rc) return rc; // the owner guard releases the frame
    char *H = ctx->h_stage;
    if (n > 0) {
        memcpy(H + o_desc, h->desc, (size_t)n * 32);
        memcpy(H + o_xy, h->kp_xy, (size_t)n * 8);
        memcpy(H + o_oct, h->octave, (size_t)n * 4);
        memcpy(H + o_ang, h->angle, (size_t)n * 4);
        if (h->u_right) memcpy(H + o_ur, h->u_right, (size_t)n * 4);
        int32_t *r0 = (int32_t *)(H + o_r0);
        int n0 = 0;
        for (int i = 0; i < n; i++) r0[i] = h->octave[i] <= 0 ? n0++ : -1;
        f->n_level0 = n0;
    }
    memcpy(H + o_sf, h->scale_factors, (size_t)nl * 4);
    memcpy(H + o_s2, h->level_sigma2, (size_t)nl * 4);
    f->fv_n_nodes = h->fv_n_nodes;
    f->fv_total = fv_total;
    f->fv_max_node = 0;
    if (h->fv_n_nodes > 0) {
        memcpy(H + o_fvn, h->fv_node_ids, (size_t)h->fv_n_nodes * 4);
        memcpy(H + o_fvo, h->fv_offsets, (size_t)(h->fv_n_nodes + 1) * 4);
        memcpy(H + o_fvf, h->fv_features, (size_t)fv_total * 4);
        for (int a = 0; a < h->fv_n_nodes; a++) f->fv_max_node = std::max(f->fv_max_node, h->fv_offsets[a + 1] - h->fv_offsets[a]);
    }
    CU_TRY(cudaMemcpyAsync(S, H, upload_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const int cof_in_smem = ((size_t)ncell + N) * 4 <= 160 * 1024;
    const size_t grid_smem = cof_in_smem ? ((size_t)ncell + N) * 4 : (size_t)ncell * 4;
    grid_build_kernel<<<1, GRID_THREADS, grid_smem, ctx->stream>>>(n, f->xy, f->octave, f->desc, f->min_x, f->min_y, f->inv_w, f->inv_h,
                                                                  f->cols, f->rows, f->cell_of, f->cell_start, f->cell_items, f->items,
                                                                  f->desc_sorted, cof_in_smem);
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(ctx->stream)); // staging buffer is reusable after this
    *out = owner.release();
    return ORBGPU_OK;
}

extern "C" int orbgpu_frame_grid_download(orbgpu_ctx *ctx, const orbgpu_frame *f, int32_t *cell_start, int32_t *cell_items)
{
    ARG_TRY(ctx && f && cell_start && cell_items);
    CU_TRY(cudaSetDevice(ctx->device));
    const int ncell = f->cols * f->rows;
    CU_TRY(cudaMemcpyAsync(cell_start, f->cell_start, (size_t)(ncell + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (f->n > 0) CU_TRY(cudaMemcpyAsync(cell_items, f->cell_items, (size_t)f->n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return ORBGPU_OK;
}

extern "C" int orbgpu_features_in_area(orbgpu_ctx *ctx, const orbgpu_frame *f, int32_t nq, const float *x, const float *y,
                                       const float *r, const int32_t *min_level, const int32_t *max_level, int32_t *out_offsets,
                                       int32_t *out_idx, int64_t cap, int64_t *total)
{
    ARG_TRY(ctx && f && nq >= 0 && out_offsets && total && cap >= 0);
    ARG_TRY(nq == 0 || (x && y && r && min_level && max_level));
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    *total = 0;
    if (nq == 0) { out_offsets[0] = 0; return ORBGPU_OK; }
    const size_t qb = align256((size_t)nq * 4);
    rc = arena_reserve(ctx, 5 * qb + 2 * align256((size_t)(nq + 1) * 4) + align256((size_t)cap * 4));
    if (rc) return rc;
    float *dx = (float *)arena_take(ctx, nq * 4), *dy = (float *)arena_take(ctx, nq * 4), *dr = (float *)arena_take(ctx, nq * 4);
    int32_t *dmin = (int32_t *)arena_take(ctx, nq * 4), *dmax = (int32_t *)arena_take(ctx, nq * 4);
    int32_t *dcnt = (int32_t *)arena_take(ctx, (nq + 1) * 4), *doff = (int32_t *)arena_take(ctx, (nq + 1) * 4);
    int32_t *dout = (int32_t *)arena_take(ctx, std::max<int64_t>(cap, 1) * 4);
    CU_TRY(cudaMemcpyAsync(dx, x, nq * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(dy, y, nq * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(dr, r, nq * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(dmin, min_level, nq * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(dmax, max_level, nq * 4, cudaMemcpyHostToDevice, ctx->stream));
    const FrameView v = frame_view(f);
    const int blocks = (nq * 32 + 255) / 256;
    area_count_kernel<<<blocks, 256, 0, ctx->stream>>>(v, nq, dx, dy, dr, dmin, dmax, dcnt);
    exclusive_scan_kernel<<<1, 1024, 0, ctx->stream>>>(dcnt, doff, nq);
    area_fill_kernel<<<blocks, 256, 0, ctx->stream>>>(v, nq, dx, dy, dr, dmin, dmax, doff, dout, cap);
    ctx->launches += 3;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(out_offsets, doff, (size_t)(nq + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    *total = out_offsets[nq];
    const int64_t ncopy = std::min<int64_t>(*total, cap);
    if (ncopy > 0 && out_idx) {
        CU_TRY(cudaMemcpyAsync(out_idx, dout, (size_t)ncopy * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return ORBGPU_OK;
}

int frame_device_init() { return set_max_dyn_smem(grid_build_kernel); }
