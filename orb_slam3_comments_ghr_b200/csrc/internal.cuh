// internal.cuh -- shared device helpers and host-side context of liborbmatch_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/orbmatch_b200.h"

// ---------------------------------------------------------------------------------------
// error handling
void orbgpu_set_error(const std::string &msg);
int orbgpu_fail(int code, const std::string &msg);

#define CU_TRY(expr)                                                                                        \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return orbgpu_fail(ORBGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
    } while (0)

// owns a freshly created handle until the creating entry point succeeds: an early `return` (CU_TRY / ARG_TRY) releases
// whatever was allocated so far through the handle's own destroy function
template <class T, void (*Destroy)(T *)>
struct OwnedHandle {
    T *p;
    explicit OwnedHandle(T *q) : p(q) {}
    ~OwnedHandle() { if (p) Destroy(p); }
    OwnedHandle(const OwnedHandle &) = delete;
    OwnedHandle &operator=(const OwnedHandle &) = delete;
    T *release() { T *q = p; p = nullptr; return q; }
};

#define ARG_TRY(cond)                                                                                       \
    do {                                                                                                    \
        if (!(cond)) return orbgpu_fail(ORBGPU_ERR_INVALID, std::string("invalid argument: ") + #cond);     \
    } while (0)

// ---------------------------------------------------------------------------------------
// grow-only device scratch arena (per context, reset at the start of every API call)
struct Arena {
    char *base = nullptr;
    size_t cap = 0, used = 0;
};

struct orbgpu_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    Arena arena;
    int64_t launches = 0;
    int64_t last_comparisons = 0;
    bool gather_counters_clean = false; // the last call on this context was a fused-gather search (it leaves d_counters[0] == 0)
    size_t list_pool_hint = 0;          // candidate-list pool entries a projection search on this context has needed so far (grow-only)
    int cmp_slot = 0;                   // d_counters slot that holds the comparisons of the last search (2 after a fused-gather search)
    unsigned long long *d_counters = nullptr; // [8] device counters: [0] comparisons, [1] overflow flag, ...
    unsigned long long *h_counters = nullptr; // pinned mirror
    int knn_engine = 0;
    void *tri_timeline = nullptr; // debug time-stamp buffer (null = off)
    int tri_engine = 0; // 0 auto, 1 CTA-per-pair kernel, 2 persistent bulk-copy pipelined kernel
    // tcgen05 engine scratch (expanded queries + per-split partial results), owned by the context
    void *knn_expanded = nullptr;
    size_t knn_expanded_bytes = 0;
    // pinned host staging (grow-only) used to pack uploads into one H2D copy
    char *h_stage = nullptr;
    size_t h_stage_bytes = 0;
    char *h_out = nullptr; // pinned landing zone of ctx_download
    size_t h_out_bytes = 0;
};
int stage_reserve(orbgpu_ctx *ctx, size_t bytes);
// live-context registry (context.cu): objects that outlive a call (frame slabs) order their release after the owning context's
// stream with an event instead of a device-wide synchronisation.  Returns false when the context is gone (its stream was
// synchronised at orbgpu_destroy, so nothing can still be reading the object).
bool ctx_record_event_if_alive(orbgpu_ctx *ctx, cudaEvent_t ev);
bool ctx_sync_if_alive(orbgpu_ctx *ctx);

// returns a 256-byte aligned device pointer valid until the next arena_reset; grows (with a
// stream sync + realloc) when needed -- only ever called BEFORE any kernel of the API call
// that uses earlier arena pointers is launched, via arena_reserve.
int arena_reserve(orbgpu_ctx *ctx, size_t total_bytes);
void *arena_take(orbgpu_ctx *ctx, size_t bytes);
inline void arena_reset(orbgpu_ctx *ctx) { ctx->arena.used = 0; }
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

int ctx_begin(orbgpu_ctx *ctx); // set device, reset arena, zero counters
// candidate-list pool of the projection searches (search_proj.cu, search_projected.cu): entries to reserve for M points against an
// n-key-point frame -- 64 per point to begin with (a window holds a few dozen), never more than the dense M x n, at least what an
// earlier call on this context needed.  The kernels count every candidate in d_counters[1]; list_pool_overflowed() reads the
// downloaded count, and when it exceeded the capacity remembers it so that the repeated call fits.
inline size_t list_pool_entries(orbgpu_ctx *ctx, size_t M, size_t n)
{
    size_t want = M * 64 > (size_t(1) << 16) ? M * 64 : (size_t(1) << 16);
    if (ctx->list_pool_hint > want) want = ctx->list_pool_hint;
    if (want > M * n) want = M * n;
    return want > 0 ? want : 1;
}
inline bool list_pool_overflowed(orbgpu_ctx *ctx, size_t cap)
{
    const size_t needed = (size_t)ctx->h_counters[1];
    if (needed <= cap) return false;
    ctx->list_pool_hint = needed + needed / 8;
    return true;
}

// Per-device kernel attributes.  cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device, per-function limit: it is raised ONCE
// per device, to the opt-in maximum, when the first context on that device is created (context.cu: device_attrs_once) -- never per
// call, so two host threads cannot lower each other's limit between a set and a launch, and a second GPU in the same process gets
// its own opt-in.  Every translation unit lists its kernels in <unit>_device_init().
template <class K>
inline int set_max_dyn_smem(K kern)
{
    int dev = 0, optin = 0;
    CU_TRY(cudaGetDevice(&dev));
    CU_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cudaFuncAttributes a;
    CU_TRY(cudaFuncGetAttributes(&a, kern));
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)a.sharedSizeBytes));
    return ORBGPU_OK;
}
constexpr size_t ORBGPU_SMEM_OPTIN = 227 * 1024; // sm_100: 232448 bytes per block (static + dynamic)
int frame_device_init();
int knn2_tc_device_init();
int search_init_device_init();
int search_proj_device_init();
int search_projected_device_init();
int search_bow_device_init();
int triangulation_device_init();
int voc_device_init();
int bowdb_device_init();
struct orbgpu_voc;
// voc.cu: FeatureVector node of n descriptors (0xFFFFFFFF for stopped words), comparisons added to counters[0]
int launch_voc_transform_nodes(orbgpu_ctx *ctx, const orbgpu_voc *voc, long long n, const uint4 *desc, int levelsup, uint32_t *node_id);
int ctx_fetch_comparisons(orbgpu_ctx *ctx); // sync + read counter[0] into last_comparisons
// Results of a host-pointer call: every piece is copied device -> pinned memory asynchronously, ONE synchronisation, then the
// pieces go to the caller's (pageable) buffers.  A cudaMemcpyAsync straight into pageable memory blocks until it is done, so
// each result array used to cost its own round trip.  Also reads the comparison counter like ctx_fetch_comparisons.
struct OutPiece {
    void *host;
    const void *dev;
    size_t bytes;
};
int ctx_download(orbgpu_ctx *ctx, const OutPiece *pieces, int n);

#define LAUNCH_COUNT(ctx) ((ctx)->launches++)

// ---------------------------------------------------------------------------------------
// device-resident objects

struct orbgpu_frame {
    int device = 0;
    orbgpu_ctx *owner = nullptr; // uploading context (only dereferenced through the live-context registry)
    char *slab = nullptr;      // one device allocation holding every array below
    size_t slab_bytes = 0;
    int32_t *cell_of = nullptr; // [n] temp: cell id per feature (-1 outside the grid)
    unsigned long long *sort_keys = nullptr; // [pow2 >= n] temp for the FeatureVector / BowVector sorts
    int sort_cap = 0;
    int32_t *fv_meta = nullptr; // device [4]: n_nodes, total, max node size, bow_n
    int n = 0;
    int n_levels = 0;
    float min_x, min_y, max_x, max_y, inv_w, inv_h;
    int cols, rows;
    // feature arrays in feature order
    uint4 *desc = nullptr;      // [n][2]
    float2 *xy = nullptr;       // [n]
    int32_t *octave = nullptr;  // [n]
    float *angle = nullptr;     // [n]
    float *u_right = nullptr;   // [n] or null
    int32_t *rank0 = nullptr;   // [n] rank of the feature among the level-0 features (octave <= 0), -1 for the others
    int n_level0 = 0;           // number of level-0 features (SearchForInitialization only looks at those, ORBmatcher.cc:762,768)
    float *scale_factors = nullptr; // [n_levels]
    float *level_sigma2 = nullptr;
    // CSR cell index (cell = ix*rows+iy), in-cell ascending feature id
    int32_t *cell_start = nullptr; // [cols*rows+1]
    int32_t *cell_items = nullptr; // [n] feature ids in cell order
    // cell-ordered copies for coalesced window scans
    int4 *items = nullptr;         // [n] {x bits, y bits, octave | iy<<16, feature id}
    uint4 *desc_sorted = nullptr;  // [n][2]
    // FeatureVector CSR (by node id) -- uploaded from the host map or built by orbgpu_transform
    int fv_n_nodes = 0;
    int fv_total = 0;
    int fv_max_node = 0;           // largest node list (launch-shape hint)
    uint32_t *fv_node_ids = nullptr; // [<=n]
    int32_t *fv_offsets = nullptr;   // [<=n+1]
    uint32_t *fv_features = nullptr; // [<=n]
    // per-feature transform outputs (orbgpu_transform)
    uint32_t *word_id = nullptr, *node_id = nullptr;
    double *weight = nullptr;
    bool has_transform = false;
    // device BowVector
    int bow_n = 0;
    uint32_t *bow_words = nullptr;
    double *bow_values = nullptr;
};

struct orbgpu_voc {
    int device = 0;
    int k = 0, L = 0, n_nodes = 0;
    uint4 *node_desc = nullptr;     // [n_nodes][2]
    int32_t *child_offsets = nullptr;
    uint32_t *child_ids = nullptr;
    double *weight = nullptr;
    uint32_t *word_id = nullptr;
};

struct orbgpu_kfset {
    int device = 0;
    int n_kf = 0, n_feat = 0, n_levels = 0;
    uint4 *desc = nullptr;      // [n_kf][n_feat][2]
    float2 *xy = nullptr;       // [n_kf][n_feat]
    int32_t *octave = nullptr;
    float *angle = nullptr;
    uint8_t *has_mp = nullptr;
    float *u_right = nullptr;
    uint32_t *node_id = nullptr;
    float *scale_factors = nullptr, *level_sigma2 = nullptr;
    // per-keyframe FeatureVector CSR restricted to features WITHOUT a map point
    // (SearchForTriangulation skips the others: ORBmatcher.cc:1129,1165)
    int32_t *kf_n_nodes = nullptr;   // [n_kf]
    uint32_t *kf_node_ids = nullptr; // [n_kf][n_feat]
    int32_t *kf_node_off = nullptr;  // [n_kf][n_feat+1]
    int32_t *kf_feat = nullptr;      // [n_kf][n_feat] feature ids grouped by node, ascending inside
    // node-ordered copies of the map-point-free features: one contiguous, coalesced read per keyframe
    uint4 *desc_csr = nullptr;       // [n_kf][n_feat][2] descriptors in CSR slot order
    int4 *kp_csr = nullptr;          // [n_kf][n_feat] {x bits, y bits, octave, feature id} in CSR slot order
    int32_t *kf_n_free = nullptr;    // [n_kf] number of CSR slots
    int max_free = 0;                // max over keyframes (sizes the shared-memory staging)
    int max_nodes = 0;
    // per-keyframe stream blob: [lo halves][node offsets][node ids], fetched with ONE bulk copy; aux = per-slot
    // 32-byte record {hi half, keypoint} read only for prefilter survivors
    uint4 *aux = nullptr;            // [n_kf][n_feat][2]
    unsigned char *blob = nullptr;   // [n_kf][blob_stride]
    size_t blob_stride = 0;
    int32_t *blob_bytes = nullptr;   // [n_kf] bytes to copy (multiple of 16)
    int max_blob = 0;                // max over keyframes
};

struct orbgpu_db {
    int device = 0;
    int64_t nd = 0, capacity = 0;
    const uint4 *desc = nullptr; // [nd][2]
    bool owned = false;
    // tcgen05 engine: the database expanded to +-1 fp8 (256 B per row, rows padded to a multiple of 256 with zeros), built by the
    // first search that needs it and kept until the descriptors change (orbgpu_db_update / orbgpu_db_invalidate).  `x_ready` is
    // recorded on the expanding context's stream; other contexts make their stream wait on it.
    mutable std::mutex x_mu;
    mutable void *x_desc = nullptr;
    mutable size_t x_bytes = 0;
    mutable bool x_valid = false;
    mutable cudaEvent_t x_ready = nullptr;
    // chunked re-upload (orbgpu_knn2_ratio_update): the descriptors cross PCIe in three chunks on the database's own stream, one event
    // per chunk; up_pending > 0 until a search has consumed (or waited for) them
    mutable cudaStream_t up_stream = nullptr;
    mutable cudaEvent_t up_ev[16] = {};
    mutable cudaEvent_t up_gate = nullptr;
    mutable int up_pending = 0;
    mutable int64_t up_row_end[16] = {}; // chunk c = rows [up_row_end[c - 1], up_row_end[c]): whole 131 072-row splits, 1/8 + 1/4 + 5/8 of them
};
// makes ctx's stream wait for every chunk of a pending upload (readers that need the whole database)
int db_wait_upload(orbgpu_ctx *ctx, const orbgpu_db *db);

// POD view of a frame passed by value to kernels
struct FrameView {
    int n, n_levels, cols, rows;
    float min_x, min_y, inv_w, inv_h;
    const uint4 *desc;
    const float2 *xy;
    const int32_t *octave;
    const float *angle;
    const float *u_right;
    const float *scale_factors;
    const float *level_sigma2;
    const int32_t *cell_start;
    const int4 *items;
    const uint4 *desc_sorted;
    int fv_n_nodes;
    const uint32_t *fv_node_ids;
    const int32_t *fv_offsets;
    const uint32_t *fv_features;
};
inline FrameView frame_view(const orbgpu_frame *f)
{
    FrameView v;
    v.n = f->n; v.n_levels = f->n_levels; v.cols = f->cols; v.rows = f->rows;
    v.min_x = f->min_x; v.min_y = f->min_y; v.inv_w = f->inv_w; v.inv_h = f->inv_h;
    v.desc = f->desc; v.xy = f->xy; v.octave = f->octave; v.angle = f->angle; v.u_right = f->u_right;
    v.scale_factors = f->scale_factors; v.level_sigma2 = f->level_sigma2;
    v.cell_start = f->cell_start; v.items = f->items; v.desc_sorted = f->desc_sorted;
    v.fv_n_nodes = f->fv_n_nodes; v.fv_node_ids = f->fv_node_ids; v.fv_offsets = f->fv_offsets; v.fv_features = f->fv_features;
    return v;
}

// ---------------------------------------------------------------------------------------
// device helpers
#ifdef __CUDACC__

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ int ham128(const uint4 a, const uint4 b)
{
    return __popc(a.x ^ b.x) + __popc(a.y ^ b.y) + __popc(a.z ^ b.z) + __popc(a.w ^ b.w);
}
// ORBmatcher::DescriptorDistance (ORBmatcher.cc:2388-2408): 256-bit XOR + popcount
__device__ __forceinline__ int ham256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1)
{
    return ham128(a0, b0) + ham128(a1, b1);
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// rotation-histogram bin (e.g. ORBmatcher.cc:829-837), factor = 1.0f/HISTO_LENGTH as the code executes
__device__ __forceinline__ int rot_bin(float a1, float a2)
{
    const float factor = 1.0f / ORBGPU_HISTO_LENGTH;
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, factor));
    if (bin == ORBGPU_HISTO_LENGTH) bin = 0;
    return bin;
}
// ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2341-2383), serial over L bins
__device__ __forceinline__ void three_maxima(const int *histo, int L, int &ind1, int &ind2, int &ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = histo[i];
        if (s > max1) {
            max3 = max2; max2 = max1; max1 = s;
            ind3 = ind2; ind2 = ind1; ind1 = i;
        } else if (s > max2) {
            max3 = max2; max2 = s;
            ind3 = ind2; ind2 = i;
        } else if (s > max3) {
            max3 = s; ind3 = i;
        }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) {
        ind2 = -1; ind3 = -1;
    } else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) {
        ind3 = -1;
    }
}

// window geometry of Frame::GetFeaturesInArea (Frame.cc:886-914); returns false on the early outs
struct CellRange { int x0, x1, y0, y1; };
__device__ __forceinline__ bool window_cells(float x, float y, float r, float min_x, float min_y, float inv_w, float inv_h,
                                             int cols, int rows, CellRange &c)
{
    c.x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, min_x), r), inv_w)));
    if (c.x0 >= cols) return false;
    c.x1 = min(cols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, min_x), r), inv_w)));
    if (c.x1 < 0) return false;
    c.y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, min_y), r), inv_h)));
    if (c.y0 >= rows) return false;
    c.y1 = min(rows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, min_y), r), inv_h)));
    if (c.y1 < 0) return false;
    return true;
}
// per-item test of Frame::GetFeaturesInArea (:939-956) given the item record {x,y,octave|iy<<16,id}
__device__ __forceinline__ bool window_item_ok(const int4 it, const CellRange &c, float x, float y, float r, int min_level,
                                               int max_level)
{
    const int iy = it.z >> 16, oct = it.z & 0xffff;
    if (iy < c.y0 || iy > c.y1) return false;
    const bool check = (min_level > 0) || (max_level >= 0);
    if (check) {
        if (oct < min_level) return false;
        if (max_level >= 0 && oct > max_level) return false;
    }
    const float dx = __fsub_rn(__int_as_float(it.x), x), dy = __fsub_rn(__int_as_float(it.y), y);
    return fabsf(dx) < r && fabsf(dy) < r;
}
// warp-cooperative window scan in the reference's iteration order.  fn(ok, pos, slot, item) is
// called by all lanes for every 32-slot chunk; pos is the ordered output position of a passing item.
template <class P, class F>
__device__ __forceinline__ int window_scan_if(const FrameView &f, float x, float y, float r, int min_level, int max_level, P &&pred, F &&fn)
{
    CellRange c;
    if (!window_cells(x, y, r, f.min_x, f.min_y, f.inv_w, f.inv_h, f.cols, f.rows, c)) return 0;
    const int s = f.cell_start[c.x0 * f.rows], e = f.cell_start[(c.x1 + 1) * f.rows];
    const int lane = lane_id();
    int count = 0;
    for (int base = s; base < e; base += 32) {
        const int slot = base + lane;
        bool ok = false;
        int4 it = make_int4(0, 0, 0, 0);
        if (slot < e) {
            it = f.items[slot];
            ok = window_item_ok(it, c, x, y, r, min_level, max_level) && pred(it);
        }
        const unsigned bal = __ballot_sync(FULL_MASK, ok);
        const int pos = count + __popc(bal & lanemask_lt());
        fn(ok, pos, slot, it);
        count += __popc(bal);
    }
    return count;
}

template <class F>
__device__ __forceinline__ int window_scan(const FrameView &f, float x, float y, float r, int min_level, int max_level, F &&fn)
{
    return window_scan_if(f, x, y, r, min_level, max_level, [](const int4 &) { return true; }, fn);
}

// warp-wide two smallest of the per-lane (b1 <= b2) key pairs; keys are unique or KEY_NONE
__device__ __forceinline__ void warp_top2(uint32_t b1, uint32_t b2, uint32_t &m1, uint32_t &m2)
{
    m1 = __reduce_min_sync(FULL_MASK, b1);
    const uint32_t cand = (b1 == m1) ? b2 : b1;
    m2 = __reduce_min_sync(FULL_MASK, cand);
}
__device__ __forceinline__ void top2_push(uint32_t &b1, uint32_t &b2, uint32_t key)
{
    const uint32_t m = max(b1, key);
    b1 = min(b1, key);
    b2 = min(b2, m);
}
#endif
