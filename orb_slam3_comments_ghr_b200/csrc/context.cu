// context.cu -- context, arena, error reporting of liborbmatch_b200.so
#include <cstdio>
#include <cstring>
#include <mutex>

#include "internal.cuh"

static thread_local std::string g_last_error;

void orbgpu_set_error(const std::string &msg) { g_last_error = msg; }
int orbgpu_fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

extern "C" const char *orbgpu_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *orbgpu_version(void) { return "orbmatch_b200 0.1.0 (sm_100a)"; }

extern "C" int orbgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

namespace {
std::mutex g_live_mu;
std::vector<orbgpu_ctx *> g_live;
} // namespace
bool ctx_record_event_if_alive(orbgpu_ctx *ctx, cudaEvent_t ev)
{
    std::lock_guard<std::mutex> lock(g_live_mu);
    for (orbgpu_ctx *c : g_live)
        if (c == ctx) return cudaEventRecord(ev, ctx->stream) == cudaSuccess;
    return false;
}
bool ctx_sync_if_alive(orbgpu_ctx *ctx)
{
    std::lock_guard<std::mutex> lock(g_live_mu);
    for (orbgpu_ctx *c : g_live)
        if (c == ctx) return cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    return false;
}

// raises every kernel's dynamic shared-memory limit to the opt-in maximum, once per device and process
static int device_attrs_once(int device)
{
    static std::mutex mu;
    static bool done[64] = {false};
    std::lock_guard<std::mutex> lock(mu);
    if (device < 64 && done[device]) return ORBGPU_OK;
    int rc;
    if ((rc = frame_device_init()) || (rc = knn2_tc_device_init()) || (rc = search_init_device_init()) ||
        (rc = search_proj_device_init()) || (rc = search_projected_device_init()) || (rc = search_bow_device_init()) ||
        (rc = triangulation_device_init()) || (rc = voc_device_init()) || (rc = bowdb_device_init()))
        return rc;
    if (device < 64) done[device] = true;
    return ORBGPU_OK;
}

static int create_common(int device, cudaStream_t s, bool own, orbgpu_ctx **out)
{
    ARG_TRY(out != nullptr);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return orbgpu_fail(ORBGPU_ERR_NO_DEVICE, "no CUDA device visible: liborbmatch_b200 has no CPU fallback");
    }
    ARG_TRY(device >= 0 && device < n);
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return orbgpu_fail(ORBGPU_ERR_NO_DEVICE, std::string("device ") + prop.name +
                                                     " is not sm_100-class: this library is built for sm_100a only");
    int rc = device_attrs_once(device);
    if (rc) return rc;
    orbgpu_ctx *c = new orbgpu_ctx();
    OwnedHandle<orbgpu_ctx, orbgpu_destroy> owner(c); // an early return below releases the stream and the partial allocations
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (own) {
        CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    } else {
        c->stream = s;
    }
    c->own_stream = own;
    CU_TRY(cudaMalloc(&c->d_counters, 8 * sizeof(unsigned long long)));
    CU_TRY(cudaMallocHost(&c->h_counters, 8 * sizeof(unsigned long long)));
    CU_TRY(cudaMemsetAsync(c->d_counters, 0, 8 * sizeof(unsigned long long), c->stream));
    {
        std::lock_guard<std::mutex> lock(g_live_mu);
        g_live.push_back(c);
    }
    *out = owner.release();
    return ORBGPU_OK;
}

extern "C" int orbgpu_create(int device, orbgpu_ctx **out) { return create_common(device, nullptr, true, out); }
extern "C" int orbgpu_create_on_stream(int device, void *cuda_stream, orbgpu_ctx **out)
{
    return create_common(device, (cudaStream_t)cuda_stream, false, out);
}

extern "C" void orbgpu_destroy(orbgpu_ctx *ctx)
{
    if (!ctx) return;
    {
        std::lock_guard<std::mutex> lock(g_live_mu);
        for (size_t i = 0; i < g_live.size(); i++)
            if (g_live[i] == ctx) { g_live.erase(g_live.begin() + i); break; }
    }
    cudaSetDevice(ctx->device);
    if (ctx->stream || !ctx->own_stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->arena.base) cudaFree(ctx->arena.base);
    if (ctx->knn_expanded) cudaFree(ctx->knn_expanded);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int orbgpu_synchronize(orbgpu_ctx *ctx)
{
    ARG_TRY(ctx != nullptr);
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return ORBGPU_OK;
}
extern "C" void *orbgpu_stream(orbgpu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int64_t orbgpu_launch_count(orbgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int64_t orbgpu_last_comparisons(orbgpu_ctx *ctx) { return ctx ? ctx->last_comparisons : 0; }

int arena_reserve(orbgpu_ctx *ctx, size_t total_bytes)
{
    total_bytes = align256(total_bytes) + 4096;
    if (total_bytes <= ctx->arena.cap) return ORBGPU_OK;
    // all earlier work that used the old buffer must be finished before it is freed
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->arena.base) CU_TRY(cudaFree(ctx->arena.base));
    ctx->arena.base = nullptr;
    ctx->arena.cap = 0;
    size_t want = total_bytes + total_bytes / 4;
    CU_TRY(cudaMalloc(&ctx->arena.base, want));
    ctx->arena.cap = want;
    ctx->arena.used = 0;
    return ORBGPU_OK;
}

void *arena_take(orbgpu_ctx *ctx, size_t bytes)
{
    bytes = align256(bytes);
    if (ctx->arena.used + bytes > ctx->arena.cap) return nullptr;
    void *p = ctx->arena.base + ctx->arena.used;
    ctx->arena.used += bytes;
    return p;
}

int ctx_begin(orbgpu_ctx *ctx)
{
    ARG_TRY(ctx != nullptr);
    CU_TRY(cudaSetDevice(ctx->device));
    arena_reset(ctx);
    ctx->gather_counters_clean = false;
    ctx->cmp_slot = 0;
    CU_TRY(cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    return ORBGPU_OK;
}

int ctx_fetch_comparisons(orbgpu_ctx *ctx)
{
    CU_TRY(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                           ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->last_comparisons = (int64_t)ctx->h_counters[ctx->cmp_slot];
    return ORBGPU_OK;
}

int ctx_download(orbgpu_ctx *ctx, const OutPiece *pieces, int n)
{
    size_t total = 0;
    for (int i = 0; i < n; i++) total += (pieces[i].bytes + 63) & ~size_t(63);
    if (total > ctx->h_out_bytes) {
        CU_TRY(cudaStreamSynchronize(ctx->stream));
        if (ctx->h_out) CU_TRY(cudaFreeHost(ctx->h_out));
        ctx->h_out = nullptr;
        ctx->h_out_bytes = 0;
        const size_t want = total + total / 4 + 4096;
        CU_TRY(cudaMallocHost(&ctx->h_out, want));
        ctx->h_out_bytes = want;
    }
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        if (pieces[i].bytes && pieces[i].host)
            CU_TRY(cudaMemcpyAsync(ctx->h_out + off, pieces[i].dev, pieces[i].bytes, cudaMemcpyDeviceToHost, ctx->stream));
        off += (pieces[i].bytes + 63) & ~size_t(63);
    }
    int rc = ctx_fetch_comparisons(ctx); // the one synchronisation
    if (rc) return rc;
    off = 0;
    for (int i = 0; i < n; i++) {
        if (pieces[i].bytes && pieces[i].host) memcpy(pieces[i].host, ctx->h_out + off, pieces[i].bytes);
        off += (pieces[i].bytes + 63) & ~size_t(63);
    }
    return ORBGPU_OK;
}

// public: synchronise the context's stream and return the comparison counter of the last search (device-pointer entry
// points do not synchronise on their own)
extern "C" int orbgpu_fetch_comparisons(orbgpu_ctx *ctx, int64_t *out)
{
    ARG_TRY(ctx && out);
    int rc = ctx_fetch_comparisons(ctx);
    if (rc) return rc;
    *out = ctx->last_comparisons;
    return ORBGPU_OK;
}

int stage_reserve(orbgpu_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->h_stage_bytes) return ORBGPU_OK;
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_stage) CU_TRY(cudaFreeHost(ctx->h_stage));
    ctx->h_stage = nullptr;
    ctx->h_stage_bytes = 0;
    size_t want = bytes + bytes / 4 + 4096;
    CU_TRY(cudaMallocHost(&ctx->h_stage, want));
    ctx->h_stage_bytes = want;
    return ORBGPU_OK;
}
