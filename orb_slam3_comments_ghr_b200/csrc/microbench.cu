// microbench.cu -- measured integer-pipe peak for the Hamming rooflines ("measure with a pure-POPC
// microbenchmark at the under-load clock").  Not on the product path: bench.py calls it once per run so that the POPC roofline
// of the batched triangulation (C4) and of the LOP3+POPC all-pairs engine is a number measured on THIS GPU in THIS run instead
// of the paper figure 16 POPC32/clk/SM.
//
//   variant 0  POPC + IADD3 only: 8 independent chains per thread, x = popc(x) + c   (the POPC issue rate itself)
//   variant 1  the Hamming triple LOP3(xor) + POPC + IADD3: a += popc(v ^ w), what a 32-bit slice of DescriptorDistance costs
//
// All SMs, one CTA of 1024 threads per SM (one wave: CTA 0 spans the kernel); time from CUDA events, SM cycles from clock64 of
// CTA 0, so the result is also reported per clock per SM at the clock the kernel actually ran at.
#include "internal.cuh"

namespace {

constexpr int MB_THREADS = 1024;
constexpr int MB_CHAINS = 8;

template <int VARIANT>
__global__ void __launch_bounds__(MB_THREADS) popc_peak_kernel(int iters, uint32_t seed, uint32_t *__restrict__ sink, long long *__restrict__ cycles)
{
    uint32_t x[MB_CHAINS];
#pragma unroll
    for (int i = 0; i < MB_CHAINS; i++) x[i] = seed * (threadIdx.x + 1) + 0x9E3779B9u * (uint32_t)(i + blockIdx.x);
    uint32_t w = seed ^ threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < MB_CHAINS; i++) {
                if (VARIANT == 0) x[i] = __popc(x[i]) + seed;          // POPC + IADD3
                else x[i] += __popc(x[(i + 1) % MB_CHAINS] ^ w);        // LOP3 + POPC + IADD3
            }
            if (VARIANT == 1) w = w * 1664525u + 1013904223u;
        }
    }
    const long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < MB_CHAINS; i++) r ^= x[i];
    if (r == 0x12345678u) sink[0] = r; // keeps the chains alive
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
}

} // namespace

// popc_per_s: POPC32 instructions (thread-level) per second over the whole GPU; per_clk_sm: the same per SM clock per SM;
// sm_mhz: the SM clock the kernel ran at (cycles of CTA 0 / event time).
extern "C" int orbgpu_measure_popc_peak(orbgpu_ctx *ctx, int32_t variant, double *popc_per_s, double *per_clk_sm, double *sm_mhz)
{
    ARG_TRY(ctx && popc_per_s && (variant == 0 || variant == 1));
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    rc = arena_reserve(ctx, 1024);
    if (rc) return rc;
    uint32_t *sink = (uint32_t *)arena_take(ctx, 256);
    long long *cyc = (long long *)arena_take(ctx, 256);
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    const int grid = ctx->sm_count, iters = 8192;
    double best_ms = 1e30;
    long long best_cyc = 0;
    for (int rep = 0; rep < 6; rep++) { // first two repetitions warm the clocks up
        CU_TRY(cudaEventRecord(e0, ctx->stream));
        if (variant == 0) popc_peak_kernel<0><<<grid, MB_THREADS, 0, ctx->stream>>>(iters, 0x2545F491u + rep, sink, cyc);
        else popc_peak_kernel<1><<<grid, MB_THREADS, 0, ctx->stream>>>(iters, 0x2545F491u + rep, sink, cyc);
        CU_TRY(cudaEventRecord(e1, ctx->stream));
        LAUNCH_COUNT(ctx);
        CU_TRY(cudaGetLastError());
        long long c = 0;
        CU_TRY(cudaMemcpyAsync(&c, cyc, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep >= 2 && ms < best_ms) { best_ms = ms; best_cyc = c; }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double n_popc = (double)grid * MB_THREADS * (double)iters * 4 * MB_CHAINS;
    *popc_per_s = n_popc / (best_ms * 1e-3);
    // CTA 0 runs for the whole kernel: one CTA per SM = one wave
    const double mhz = (double)best_cyc / (best_ms * 1e-3) / 1e6;
    if (sm_mhz) *sm_mhz = mhz;
    if (per_clk_sm) *per_clk_sm = n_popc / ((double)best_cyc * ctx->sm_count);
    return ORBGPU_OK;
}
