// knn2_tc.cu -- tcgen05 engine of the brute-force 2-NN Hamming search (config C5).
//
// The all-pairs search is the one dense contraction of the hot path: with every descriptor bit b
// mapped to the fp8 (e4m3) value (-1)^b, the 256-term dot product of two descriptors is
// 256 - 2*hamming, so  hamming = 128 - dot/2  and the nearest neighbour is the LARGEST dot.
// All partial sums are integers of magnitude <= 256, exactly representable in fp16, so the
// tensor cores accumulate in fp16 without any rounding and the result is bit-exact.
//
//   expand_kernel       bits -> fp8 bytes (0x38 = +1, 0xB8 = -1), rows padded to a multiple of 128
//   knn2_tc_kernel      persistent, warp specialised, one CTA per SM:
//                         warp 0      TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B tiles)
//                         warp 1      MMA issuer     (tcgen05.mma kind::f8f6f4, M=128 N=128 K=32 x8)
//                         warp 2      TMEM allocator (512 columns = 4 accumulator buffers of 128)
//                         warps 4-11  epilogue       (two warpgroups on alternate accumulator buffers;
//                                                     thread == query row == TMEM lane)
//                       The epilogue reads the fp16 accumulators as packed half2 (tcgen05.ld ...pack::16b)
//                       and keeps the two largest dots per column parity with 3 HMNMX2 per 2 elements,
//                       plus the FIRST 128-row database stage in which the running maximum was reached.
//   knn2_tc_merge_kernel one warp per query: merges the per-split partials and re-scans the 128 rows
//                       of the winning stage with XOR+POPC to recover the first index attaining the
//                       minimum (strict-'<' first-index tie-break of the reference loop).
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "internal.cuh"

namespace {
// bisection switches used while bringing the pipeline up (1 = skip the top-2 arithmetic, 2 = also the TMEM loads, 4/8 = narrow MMA
// shapes, 16 = one k-block per stage); compiled out
constexpr int TC_DBG = 0;


constexpr int BM = 128;            // queries per CTA tile (UMMA M, TMEM lanes)
constexpr int BN = 128;            // database rows per stage (UMMA N)
constexpr int ROW_BYTES = 256;     // expanded descriptor: 256 fp8 values
constexpr int KB_BYTES = 128;      // one SWIZZLE_128B K-block
constexpr int NUM_KB = ROW_BYTES / KB_BYTES;
constexpr int TILE_BYTES = BM * KB_BYTES;      // 16 KB: 128 rows x 128 B
constexpr int STAGE_BYTES = NUM_KB * TILE_BYTES; // 32 KB
constexpr int NS = 4;              // B pipeline stages
constexpr int NA = 4;              // TMEM accumulator buffers (4 x 128 columns = 512)
constexpr int NUM_THREADS = 384;   // 12 warps
constexpr int EPI_WARP0 = 4;
constexpr uint32_t TMEM_COLS = 512;
constexpr int SMEM_BYTES = 2 * STAGE_BYTES + NS * STAGE_BYTES + 4096 /*barriers + merge scratch*/ + 1024 /*align slack*/;

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp8 operands, fp16 accumulators
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 64 fp16 columns -> 32 registers of packed half2 (column 2i in the low half)
__device__ __forceinline__ void tc_ld_64cols_packed(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4, [16,30) LBO >> 4 (unused for swizzled K-major), [32,46) SBO >> 4 = 1024 B between
//   8-row groups, [46,48) version = 1 (Blackwell), [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: c_format [4,6) = 0 (F16), a_format [7,10) = 0 (E4M3), b_format [10,13) = 0 (E4M3),
// a/b major = K (0), n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4
constexpr uint32_t IDESC = ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct TcParams {
    int64_t nq, nd;
    int n_qtiles, n_splits; // n_splits: the splits THIS launch sweeps, starting at split0 (a launch per uploaded chunk, or all at once)
    int split0;
    int stages_per_split; // 128-row stages per database split
    int total_stages;     // ceil(nd / 128)
    uint16_t *out_best;   // [n_splits][nq_pad] best hamming distance (0xFFFF none)
    uint16_t *out_second; // [n_splits][nq_pad]
    int32_t *out_stage;   // [n_splits][nq_pad] first database row of the winning stage
    int64_t out_stride;
};

__device__ __forceinline__ __half2 u2h2(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }

__global__ void __launch_bounds__(NUM_THREADS, 1)
knn2_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_db, TcParams P)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                        // [2][STAGE_BYTES]
    uint8_t *sB = smem + 2 * STAGE_BYTES;      // [NS][STAGE_BYTES]
    uint64_t *bars = (uint64_t *)(smem + (2 + NS) * STAGE_BYTES);
    // barrier indices
    uint64_t *b_full = bars, *b_empty = bars + NS, *a_full = bars + 2 * NS, *a_empty = a_full + 2, *t_full = a_empty + 2,
             *t_empty = t_full + NA;
    uint32_t *tmem_slot = (uint32_t *)(t_empty + NA);
    // merge scratch for the two epilogue warpgroups: [128 rows][4] (m1, m2 as raw half2, gmax stage)
    uint32_t *sMerge = tmem_slot + 4;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = P.n_qtiles * P.n_splits;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < NS; i++) { mbar_init(smem_u32(&b_full[i]), 1); mbar_init(smem_u32(&b_empty[i]), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(smem_u32(&a_full[i]), 1); mbar_init(smem_u32(&a_empty[i]), 1); }
        for (int i = 0; i < NA; i++) { mbar_init(smem_u32(&t_full[i]), 1); mbar_init(smem_u32(&t_empty[i]), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_db) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t sb = 0, pb = 0; // B ring stage / phase
            int it_local = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, it_local++) {
                const int qt = item % P.n_qtiles, split = P.split0 + item / P.n_qtiles;
                const int abuf = it_local & 1;
                mbar_wait(smem_u32(&a_empty[abuf]), ((it_local >> 1) & 1) ^ 1);
                mbar_expect_tx(smem_u32(&a_full[abuf]), STAGE_BYTES);
                for (int kb = 0; kb < NUM_KB; kb++)
                    tma_load_2d(smem_u32(sA + abuf * STAGE_BYTES + kb * TILE_BYTES), &map_q, smem_u32(&a_full[abuf]), kb * KB_BYTES, qt * BM);
                const int s0 = split * P.stages_per_split, s1 = min(P.total_stages, s0 + P.stages_per_split);
                for (int s = s0; s < s1; s++) {
                    mbar_wait(smem_u32(&b_empty[sb]), pb ^ 1);
                    mbar_expect_tx(smem_u32(&b_full[sb]), (TC_DBG & 16) ? TILE_BYTES : STAGE_BYTES);
                    for (int kb = 0; kb < ((TC_DBG & 16) ? 1 : NUM_KB); kb++)
                        tma_load_2d(smem_u32(sB + sb * STAGE_BYTES + kb * TILE_BYTES), &map_db, smem_u32(&b_full[sb]), kb * KB_BYTES, s * BN);
                    if (++sb == NS) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t sb = 0, pb = 0, ta = 0, pt = 0;
            int it_local = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, it_local++) {
                const int split = P.split0 + item / P.n_qtiles;
                const int abuf = it_local & 1;
                mbar_wait(smem_u32(&a_full[abuf]), (it_local >> 1) & 1);
                const int s0 = split * P.stages_per_split, s1 = min(P.total_stages, s0 + P.stages_per_split);
                for (int s = s0; s < s1; s++) {
                    mbar_wait(smem_u32(&t_empty[ta]), pt ^ 1); // epilogue drained this accumulator buffer
                    mbar_wait(smem_u32(&b_full[sb]), pb);       // TMA landed this stage
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + ta * BN;
#pragma unroll
                    for (int kb = 0; kb < NUM_KB; kb++) {
                        const uint64_t ad = make_desc(smem_u32(sA + abuf * STAGE_BYTES + kb * TILE_BYTES));
                        const uint64_t bd = make_desc(smem_u32(sB + sb * STAGE_BYTES + kb * TILE_BYTES));
#pragma unroll
                        for (int k = 0; k < KB_BYTES / 32; k++) // UMMA K = 32 bytes; advance start address by 32 B >> 4 = 2
                            tc_mma_f8(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), (TC_DBG & 4) ? (((uint32_t)(64 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24)) : ((TC_DBG & 8) ? (((uint32_t)(BN >> 3) << 17) | ((uint32_t)(64 >> 4) << 24)) : IDESC), (kb | k) ? 1u : 0u);
                    }
                    tc_commit(smem_u32(&b_empty[sb])); // smem stage reusable once these MMAs retire
                    tc_commit(smem_u32(&t_full[ta]));  // accumulator ready for the epilogue
                    if (++sb == NS) { sb = 0; pb ^= 1; }
                    if (++ta == NA) { ta = 0; pt ^= 1; }
                }
                tc_commit(smem_u32(&a_empty[abuf])); // A tile free for the item after next
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ================= epilogue: thread == query row == TMEM lane =================
        const int wg = (warp - EPI_WARP0) >> 2;  // 0 or 1: handles accumulator buffers wg, wg+2
        const int wq = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = wq * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
        const __half2 NEG_INF2 = __half2half2(__ushort_as_half((unsigned short)0xFC00));
        uint32_t pt[2] = {0, 0}; // phase of buffers wg and wg+2
        int it_local = 0;
        int stage_counter = 0;   // global (per CTA) stage counter to know which buffer a stage uses
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, it_local++) {
            const int qt = item % P.n_qtiles, split = P.split0 + item / P.n_qtiles;
            const int s0 = split * P.stages_per_split, s1 = min(P.total_stages, s0 + P.stages_per_split);
            __half2 m1 = NEG_INF2, m2 = NEG_INF2;
            float gmax = -1e30f;
            int best_stage = -1;
            for (int s = s0; s < s1; s++, stage_counter++) {
                const int buf = stage_counter & (NA - 1);
                if ((buf & 1) != wg) continue;
                const int pi = buf >> 1;
                mbar_wait(smem_u32(&t_full[buf]), pt[pi]);
                pt[pi] ^= 1;
                tc_fence_after();
                uint32_t r0[32], r1[32];
                if (!(TC_DBG & 2)) {
                    tc_ld_64cols_packed(lane_base + buf * BN, r0);
                    tc_ld_64cols_packed(lane_base + buf * BN + 64, r1);
                    tc_wait_ld();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i++) { r0[i] = 0; r1[i] = 0; }
                }
                // accumulator buffer can be overwritten as soon as the registers hold it
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&t_empty[buf]));
                const int valid = (int)min((int64_t)BN, P.nd - (int64_t)s * BN); // rows of this stage inside the database
                if (valid < BN) {
                    // last, partial stage: columns >= valid are padding -> -inf
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        const int c0 = 2 * i, c1 = 64 + 2 * i;
                        __half2 v0 = u2h2(r0[i]), v1 = u2h2(r1[i]);
                        if (c0 >= valid) v0 = NEG_INF2; else if (c0 + 1 >= valid) v0 = __halves2half2(__low2half(v0), __low2half(NEG_INF2));
                        if (c1 >= valid) v1 = NEG_INF2; else if (c1 + 1 >= valid) v1 = __halves2half2(__low2half(v1), __low2half(NEG_INF2));
                        r0[i] = *reinterpret_cast<uint32_t *>(&v0);
                        r1[i] = *reinterpret_cast<uint32_t *>(&v1);
                    }
                }
                if (!(TC_DBG & 1)) {
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const __half2 v = u2h2(r0[i]);
                    const __half2 lo = __hmin2(m1, v);
                    m1 = __hmax2(m1, v);
                    m2 = __hmax2(m2, lo);
                }
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const __half2 v = u2h2(r1[i]);
                    const __half2 lo = __hmin2(m1, v);
                    m1 = __hmax2(m1, v);
                    m2 = __hmax2(m2, lo);
                }
                } else { m1 = __hmax2(m1, u2h2(r0[0] ^ r1[31])); }
                const float cm = fmaxf(__low2float(m1), __high2float(m1));
                if (cm > gmax) { // strictly better: this is the FIRST stage (of this warpgroup) reaching the new maximum
                    gmax = cm;
                    best_stage = s;
                }
            }
            // ---- merge the two warpgroups (they saw alternate stages of the same rows) and write the partial
            // named barrier 1: the 256 epilogue threads only
            if (wg == 1) {
                sMerge[row * 4 + 0] = *reinterpret_cast<uint32_t *>(&m1);
                sMerge[row * 4 + 1] = *reinterpret_cast<uint32_t *>(&m2);
                sMerge[row * 4 + 2] = __float_as_uint(gmax);
                sMerge[row * 4 + 3] = (uint32_t)best_stage;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (wg == 0) {
                const __half2 o1 = u2h2(sMerge[row * 4 + 0]), o2 = u2h2(sMerge[row * 4 + 1]);
                const float og = __uint_as_float(sMerge[row * 4 + 2]);
                const int os = (int)sMerge[row * 4 + 3];
                // six candidates for the two largest dots: 4 lane-parity maxima and 4 runners-up
                float v[8] = {__low2float(m1), __high2float(m1), __low2float(m2), __high2float(m2),
                              __low2float(o1), __high2float(o1), __low2float(o2), __high2float(o2)};
                float t1 = -1e30f, t2 = -1e30f;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float x = v[i];
                    const float lo = fminf(t1, x);
                    t1 = fmaxf(t1, x);
                    t2 = fmaxf(t2, lo);
                }
                int bs = best_stage;
                if (og > gmax || (og == gmax && os >= 0 && (bs < 0 || os < bs))) bs = os;
                const int64_t q = (int64_t)qt * BM + row;
                if (q < P.nq) {
                    // hamming = 128 - dot/2 (dot is an even integer in [-256, 256]); -inf -> none
                    const uint16_t hb = (t1 < -1000.f) ? (uint16_t)0xFFFF : (uint16_t)(128 - (int)t1 / 2);
                    const uint16_t hs = (t2 < -1000.f) ? (uint16_t)0xFFFF : (uint16_t)(128 - (int)t2 / 2);
                    const int64_t o = (int64_t)split * P.out_stride + q;
                    P.out_best[o] = hb;
                    P.out_second[o] = hs;
                    P.out_stage[o] = (bs < 0) ? -1 : bs * BN;
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory"); // sMerge reusable
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// =============================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs (one SM pair) computes a 256-query x 128-row tile per
// MMA.  Each CTA holds its own 128 query rows (A) and HALF of every database stage (64 rows of B); the
// leader CTA issues tcgen05.mma.cta_group::2 (M = 256) and the accumulators land in both CTAs' TMEM
// (rows 0-127 in the leader, 128-255 in the peer).  Versus the 1-CTA kernel this halves the L2->SMEM
// traffic per SM (16 KB instead of 32 KB per stage) and the shared-memory operand reads per MMA
// (A 4 KB + B 2 KB instead of 4 + 4), which is what limited the 1-CTA kernel (ncu: 59 % tensor pipe,
// l1tex tc wavefronts at 59 %, ~5.6 KB/clk of L2 reads chip-wide).
constexpr int BN2 = 256;                        // database rows per stage of the 2-CTA kernel (UMMA N = 256)
constexpr int NS2 = 4;                          // B pipeline stages (this CTA's 128 rows = 32 KB each)
constexpr int NA2 = 2;                          // accumulator buffers (2 x 256 columns = 512)
constexpr int SMEM2_BYTES = 2 * STAGE_BYTES + NS2 * STAGE_BYTES + 4096 + 1024;
constexpr uint32_t IDESC2 = ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24); // M = 256, N = 256

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    // relaxed: the accumulator values are already in registers (tcgen05.wait::ld) and the tcgen05 fences order the
    // TMEM accesses; a release at cluster scope would cost a MEMBAR.ALL.GPU + ERRBAR per stage (ncu: 40 % of samples)
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f8_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// top-2 update of one 64-column chunk (32 packed half2 registers); columns >= valid are padding -> -inf
__device__ __forceinline__ void epi_consume(uint32_t (&r)[32], int base_col, int valid, __half2 &m1, __half2 &m2)
{
    const __half2 NEG_INF2 = __half2half2(__ushort_as_half((unsigned short)0xFC00));
    if (base_col + 64 > valid) {
#pragma unroll
        for (int i = 0; i < 32; i++) {
            const int c = base_col + 2 * i;
            __half2 v = u2h2(r[i]);
            if (c >= valid) v = NEG_INF2; else if (c + 1 >= valid) v = __halves2half2(__low2half(v), __low2half(NEG_INF2));
            r[i] = *reinterpret_cast<uint32_t *>(&v);
        }
    }
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const __half2 v = u2h2(r[i]);
        const __half2 lo = __hmin2(m1, v);
        m1 = __hmax2(m1, v);
        m2 = __hmax2(m2, lo);
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
knn2_tc2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_db, TcParams P)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                        // [2][STAGE_BYTES]    this CTA's 128 query rows
    uint8_t *sB = smem + 2 * STAGE_BYTES;      // [NS2][STAGE_BYTES]  this CTA's 128 rows of every 256-row stage
    uint64_t *bars = (uint64_t *)(smem + 2 * STAGE_BYTES + NS2 * STAGE_BYTES);
    uint64_t *b_full = bars, *b_empty = bars + NS2, *a_full = bars + 2 * NS2, *a_empty = a_full + 2, *t_full = a_empty + 2,
             *t_empty = t_full + NA2;
    uint32_t *tmem_slot = (uint32_t *)(t_empty + NA2);
    uint32_t *sMerge = tmem_slot + 4;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_qt2 = (P.n_qtiles + 1) / 2;           // query-tile PAIRS
    const int n_items = n_qt2 * P.n_splits;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < NS2; i++) { mbar_init(smem_u32(&b_full[i]), 1); mbar_init(smem_u32(&b_empty[i]), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(smem_u32(&a_full[i]), 1); mbar_init(smem_u32(&a_empty[i]), 1); }
        for (int i = 0; i < NA2; i++) { mbar_init(smem_u32(&t_full[i]), 1); mbar_init(smem_u32(&t_empty[i]), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_db) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all(); // barriers of both CTAs are initialised before any remote arrive / TMA signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer (both CTAs): own A tile + own 128 rows of every B stage =================
        if (lane == 0) {
            uint32_t sb = 0, pb = 0;
            int it_local = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters, it_local++) {
                const int qt = 2 * (item % n_qt2) + (int)rank, split = P.split0 + item / n_qt2;
                const int abuf = it_local & 1;
                mbar_wait(smem_u32(&a_empty[abuf]), ((it_local >> 1) & 1) ^ 1);
                const uint32_t afull_leader = mapa(smem_u32(&a_full[abuf]), 0);
                if (leader) mbar_expect_tx(smem_u32(&a_full[abuf]), 2 * STAGE_BYTES); // both CTAs' A tiles
                for (int kb = 0; kb < NUM_KB; kb++)
                    tma_load_2d_2sm(smem_u32(sA + abuf * STAGE_BYTES + kb * TILE_BYTES), &map_q, afull_leader, kb * KB_BYTES, qt * BM);
                const int s0 = split * P.stages_per_split, s1 = min(P.total_stages, s0 + P.stages_per_split);
                for (int s = s0; s < s1; s++) {
                    mbar_wait(smem_u32(&b_empty[sb]), pb ^ 1);
                    const uint32_t bfull_leader = mapa(smem_u32(&b_full[sb]), 0);
                    if (leader) mbar_expect_tx(smem_u32(&b_full[sb]), 2 * STAGE_BYTES); // both halves
                    for (int kb = 0; kb < NUM_KB; kb++)
                        tma_load_2d_2sm(smem_u32(sB + sb * STAGE_BYTES + kb * TILE_BYTES), &map_db, bfull_leader, kb * KB_BYTES,
                                        s * BN2 + (int)rank * (BN2 / 2));
                    if (++sb == NS2) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (leader && lane == 0) {
            uint32_t sb = 0, pb = 0, ta = 0, pt = 0;
            int it_local = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters, it_local++) {
                const int split = P.split0 + item / n_qt2;
                const int abuf = it_local & 1;
                mbar_wait(smem_u32(&a_full[abuf]), (it_local >> 1) & 1);
                const int s0 = split * P.stages_per_split, s1 = min(P.total_stages, s0 + P.stages_per_split);
                for (int s = s0; s < s1; s++) {
                    mbar_wait(smem_u32(&t_empty[ta]), pt ^ 1); // both CTAs' epilogues drained this accumulator buffer
                    mbar_wait(smem_u32(&b_full[sb]), pb);       // both halves of the stage landed
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + ta * BN2;
#pragma unroll
                    for (int kb = 0; kb < NUM_KB; kb++) {
                        const uint64_t ad = make_desc(smem_u32(sA + abuf * STAGE_BYTES + kb * TILE_BYTES));
                        const uint64_t bd = make_desc(smem_u32(sB + sb * STAGE_BYTES + kb * TILE_BYTES));
#pragma unroll
                        for (int k = 0; k < KB_BYTES / 32; k++)
                            tc_mma_f8_2sm(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), IDESC2, (kb | k) ? 1u : 0u);
                    }
                    tc_commit_2sm(smem_u32(&b_empty[sb])); // frees the stage in BOTH CTAs
                    tc_commit_2sm(smem_u32(&t_full[ta]));  // accumulator ready in BOTH CTAs
                    if (++sb == NS2) { sb = 0; pb ^= 1; }
                    if (++ta == NA2) { ta = 0; pt ^= 1; }
                }
                tc_commit_2sm(smem_u32(&a_empty[abuf]));
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ================= epilogue (both CTAs): thread == query row == TMEM lane; warpgroup g owns buffer g =================
        const int wg = (warp - EPI_WARP0) >> 2;
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16) + wg * BN2;
        const __half2 NEG_INF2 = __half2half2(__ushort_as_half((unsigned short)0xFC00));
        const uint32_t tempty_leader = mapa(smem_u32(&t_empty[wg]), 0);
        uint32_t pt = 0;
        int stage_counter = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters) {
            const int qt = 2 * (item % n_qt2) + (int)rank, split = P.split0 + item / n_qt2;
            const int s0 = split * P.stages_per_split, s1 = min(P.total_stages, s0 + P.stages_per_split);
            __half2 m1 = NEG_INF2, m2 = NEG_INF2;
            float gmax = -1e30f;
            int best_stage = -1;
            for (int s = s0; s < s1; s++, stage_counter++) {
                if ((stage_counter & 1) != wg) continue;
                mbar_wait(smem_u32(&t_full[wg]), pt);
                pt ^= 1;
                tc_fence_after();
                const int valid = (int)min((int64_t)BN2, P.nd - (int64_t)s * BN2);
                uint32_t r0[32], r1[32];
                tc_ld_64cols_packed(lane_base, r0);
                tc_ld_64cols_packed(lane_base + 64, r1);
                tc_wait_ld();
                epi_consume(r0, 0, valid, m1, m2);
                tc_ld_64cols_packed(lane_base + 128, r0);
                epi_consume(r1, 64, valid, m1, m2);
                tc_ld_64cols_packed(lane_base + 192, r1);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader); // the leader's barrier counts both CTAs (8 warps)
                epi_consume(r0, 128, valid, m1, m2);
                epi_consume(r1, 192, valid, m1, m2);
                const float cm = fmaxf(__low2float(m1), __high2float(m1));
                if (cm > gmax) { // strictly better: FIRST stage (of this warpgroup) reaching the new maximum
                    gmax = cm;
                    best_stage = s;
                }
            }
            if (wg == 1) {
                sMerge[row * 4 + 0] = *reinterpret_cast<uint32_t *>(&m1);
                sMerge[row * 4 + 1] = *reinterpret_cast<uint32_t *>(&m2);
                sMerge[row * 4 + 2] = __float_as_uint(gmax);
                sMerge[row * 4 + 3] = (uint32_t)best_stage;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (wg == 0) {
                const __half2 o1 = u2h2(sMerge[row * 4 + 0]), o2 = u2h2(sMerge[row * 4 + 1]);
                const float og = __uint_as_float(sMerge[row * 4 + 2]);
                const int os = (int)sMerge[row * 4 + 3];
                float v[8] = {__low2float(m1), __high2float(m1), __low2float(m2), __high2float(m2),
                              __low2float(o1), __high2float(o1), __low2float(o2), __high2float(o2)};
                float t1 = -1e30f, t2 = -1e30f;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float x = v[i];
                    const float lo = fminf(t1, x);
                    t1 = fmaxf(t1, x);
                    t2 = fmaxf(t2, lo);
                }
                int bs = best_stage;
                if (og > gmax || (og == gmax && os >= 0 && (bs < 0 || os < bs))) bs = os;
                const int64_t q = (int64_t)qt * BM + row;
                if (q < P.nq) {
                    const uint16_t hb = (t1 < -1000.f) ? (uint16_t)0xFFFF : (uint16_t)(128 - (int)t1 / 2);
                    const uint16_t hs = (t2 < -1000.f) ? (uint16_t)0xFFFF : (uint16_t)(128 - (int)t2 / 2);
                    const int64_t o = (int64_t)split * P.out_stride + q;
                    P.out_best[o] = hb;
                    P.out_second[o] = hs;
                    P.out_stage[o] = (bs < 0) ? -1 : bs * BN2;
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }
    // ---- teardown: nobody leaves while the pair may still signal its barriers or write its TMEM
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// bits -> fp8 (+1 = 0x38, -1 = 0xB8); one thread per 32-bit word -> 32 output bytes.  Rows >= n are zero (padding).
__global__ void expand_kernel(const uint32_t *__restrict__ in, int64_t n, int64_t n_pad, uint4 *__restrict__ out)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // word index: row = w / 8
    if (w >= n_pad * 8) return;
    const int64_t row = w >> 3;
    uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
    if (row < n) {
        const uint32_t x = in[w];
        uint32_t e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t nib = (x >> (4 * i)) & 0xF;
            e[i] = (((nib * 0x00204081u) & 0x01010101u) << 7) | 0x38383838u;
        }
        o0 = make_uint4(e[0], e[1], e[2], e[3]);
        o1 = make_uint4(e[4], e[5], e[6], e[7]);
    }
    out[2 * w] = o0;
    out[2 * w + 1] = o1;
}

// one warp per query: merge the per-split partials, then recover the first index attaining the minimum
__global__ void knn2_tc_merge_kernel(const uint4 *__restrict__ q, const uint4 *__restrict__ db, int64_t nq, int64_t nd, int n_splits, int bn,
                                     int64_t stride, const uint16_t *__restrict__ pb, const uint16_t *__restrict__ ps,
                                     const int32_t *__restrict__ pstage, int th_low, float nnratio, int32_t *__restrict__ best_idx,
                                     int32_t *__restrict__ best_dist, int32_t *__restrict__ second_dist, int32_t *__restrict__ match)
{
    const int64_t qi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (qi >= nq) return;
    const int lane = threadIdx.x & 31;
    // lexicographic (dist, split) minimum == first split attaining the minimum (splits are ascending database ranges)
    uint32_t b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;
    for (int s = lane; s < n_splits; s += 32) {
        const uint32_t d = pb[(int64_t)s * stride + qi];
        if (d != 0xFFFFu) top2_push(b1, b2, (d << 16) | (uint32_t)s);
    }
    uint32_t m1, m2;
    warp_top2(b1, b2, m1, m2);
    int bd = 256, bi = -1, sd = 256;
    if (m1 != 0xFFFFFFFFu) {
        bd = (int)(m1 >> 16);
        const int win = (int)(m1 & 0xFFFF);
        uint32_t sec = ps[(int64_t)win * stride + qi];
        if (m2 != 0xFFFFFFFFu) sec = min(sec, m2 >> 16);
        if (sec < 256u) sd = (int)sec;
        const int64_t st = pstage[(int64_t)win * stride + qi];
        const uint4 qa = q[2 * qi], qb = q[2 * qi + 1];
        int found = 0x7FFFFFFF;
        for (int j = lane; j < bn; j += 32) {
            const int64_t r = st + j;
            if (r >= 0 && r < nd && ham256(qa, qb, db[2 * r], db[2 * r + 1]) == bd) found = min(found, (int)r);
        }
        found = __reduce_min_sync(FULL_MASK, found);
        bi = (found == 0x7FFFFFFF) ? -2 : found; // -2 would flag an internal inconsistency (never expected)
    }
    if (lane == 0) {
        int m = -1;
        if (bd <= th_low)
            if ((float)bd < __fmul_rn(nnratio, (float)sd)) m = bi;
        if (best_idx) best_idx[qi] = bi;
        if (best_dist) best_dist[qi] = bd;
        if (second_dist) second_dist[qi] = sd;
        if (match) match[qi] = m;
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// rows x 256 bytes, box = 128 bytes x 128 rows, SWIZZLE_128B
int make_map(CUtensorMap *m, void *base, uint64_t rows, uint32_t box_rows)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) return orbgpu_fail(ORBGPU_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)ROW_BYTES, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ROW_BYTES};
    cuuint32_t box[2] = {(cuuint32_t)KB_BYTES, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return orbgpu_fail(ORBGPU_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    return ORBGPU_OK;
}

} // namespace

bool knn2_tc_supported() { return true; }

// q, outputs: device pointers.  Runs expansion + tcgen05 search + merge on ctx->stream.
int knn2_tc_run(orbgpu_ctx *ctx, const orbgpu_db *db, int64_t nq, const uint4 *q, int32_t th_low, float nnratio, int32_t *best_idx,
                int32_t *best_dist, int32_t *second_dist, int32_t *match, bool two_cta)
{
    const int64_t nd = db->nd;
    const int bn = two_cta ? BN2 : BN; // database rows per stage
    // rows padded to 256 for both engines, so that one expanded copy of the database serves either
    const int64_t nq_pad = (nq + 2 * BM - 1) / (2 * BM) * (2 * BM), nd_pad = (nd + BN2 - 1) / BN2 * BN2;
    const int n_qtiles = (int)(nq_pad / BM);
    const int total_stages = (int)((nd + bn - 1) / bn);
    const int n_units = two_cta ? (n_qtiles + 1) / 2 : n_qtiles; // work items per split
    const int n_workers = two_cta ? ctx->sm_count / 2 : ctx->sm_count;
    // database splits: each split (expanded: rows x 256 B) should stay L2 resident while every query tile sweeps it
    int stages_per_split = std::min(total_stages, (128 * 1024) / bn); // 128k rows = 32 MiB expanded per split
    int n_splits = (total_stages + stages_per_split - 1) / stages_per_split;
    if (n_units * n_splits < 2 * n_workers) {
        // few query tiles (one frame against the map: 2 k queries): the items must fill the persistent grid in whole waves --
        // the largest item count not above a multiple of the worker count, one wave when that still leaves >= 64 stages per item
        int want = std::max(1, n_workers / n_units); // one wave
        while (want > 1 && (total_stages + want - 1) / want < 8) want--;
        stages_per_split = (total_stages + want - 1) / want;
        n_splits = (total_stages + stages_per_split - 1) / stages_per_split;
    }
    // context scratch: expanded queries + partials.  The expanded database belongs to the database object and is kept.
    const size_t e_db = (size_t)nd_pad * ROW_BYTES, e_q = (size_t)nq_pad * ROW_BYTES;
    const size_t part = (size_t)n_splits * nq_pad;
    const size_t need = align256(e_q) + align256(part * 2) * 2 + align256(part * 4) + 4096;
    if (need > ctx->knn_expanded_bytes) {
        CU_TRY(cudaStreamSynchronize(ctx->stream));
        if (ctx->knn_expanded) CU_TRY(cudaFree(ctx->knn_expanded));
        ctx->knn_expanded = nullptr;
        ctx->knn_expanded_bytes = 0;
        CU_TRY(cudaMalloc(&ctx->knn_expanded, need));
        ctx->knn_expanded_bytes = need;
    }
    char *base = (char *)(((uintptr_t)ctx->knn_expanded + 1023) & ~(uintptr_t)1023);
    uint8_t *x_q = (uint8_t *)base;
    uint16_t *pb = (uint16_t *)(x_q + align256(e_q));
    uint16_t *ps = (uint16_t *)((char *)pb + align256(part * 2));
    int32_t *pst = (int32_t *)((char *)ps + align256(part * 2));
    uint8_t *x_db = nullptr;
    // a chunked upload in flight (orbgpu_knn2_ratio_update): the search of a chunk's splits starts as soon as the chunk has arrived
    // and been expanded, while the following chunks are still crossing PCIe on the database's own stream
    int up_chunks = 0;
    {
        std::lock_guard<std::mutex> lock(db->x_mu);
        if (db->x_bytes < e_db + 1024) {
            if (db->x_desc) { // grown database (or first use): nothing may still be reading the old copy
                CU_TRY(cudaDeviceSynchronize());
                CU_TRY(cudaFree(db->x_desc));
                db->x_desc = nullptr;
                db->x_bytes = 0;
            }
            CU_TRY(cudaMalloc(&db->x_desc, e_db + 1024));
            db->x_bytes = e_db + 1024;
            db->x_valid = false;
        }
        if (!db->x_ready) CU_TRY(cudaEventCreateWithFlags(&db->x_ready, cudaEventDisableTiming));
        x_db = (uint8_t *)(((uintptr_t)db->x_desc + 1023) & ~(uintptr_t)1023);
        const int64_t split_rows = (int64_t)stages_per_split * bn;
        if (db->up_pending > 0 && !db->x_valid && split_rows == 131072) { // the chunks are whole splits of this plan
            up_chunks = db->up_pending;
        } else {
            int rc = db_wait_upload(ctx, db); // whole database first (no upload pending: nothing to wait for)
            if (rc) return rc;
            if (!db->x_valid) {
                expand_kernel<<<(unsigned)((nd_pad * 8 + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)db->desc, nd, nd_pad, (uint4 *)x_db);
                LAUNCH_COUNT(ctx);
                CU_TRY(cudaGetLastError());
                CU_TRY(cudaEventRecord(db->x_ready, ctx->stream));
                db->x_valid = true;
            } else {
                CU_TRY(cudaStreamWaitEvent(ctx->stream, db->x_ready, 0)); // expanded on another context's stream, possibly still in flight
            }
        }
    }
    expand_kernel<<<(unsigned)((nq_pad * 8 + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)q, nq, nq_pad, (uint4 *)x_q);
    LAUNCH_COUNT(ctx);
    CUtensorMap mq, mdb;
    int rc = make_map(&mq, x_q, (uint64_t)nq_pad, BM);
    if (rc) return rc;
    rc = make_map(&mdb, x_db, (uint64_t)nd_pad, BN); // 128-row boxes: a whole 1-CTA stage / this CTA's half of a 2-CTA stage
    if (rc) return rc;
    TcParams P;
    P.nq = nq; P.nd = nd; P.n_qtiles = n_qtiles; P.n_splits = n_splits; P.split0 = 0; P.stages_per_split = stages_per_split;
    P.total_stages = total_stages; P.out_best = pb; P.out_second = ps; P.out_stage = pst; P.out_stride = nq_pad;
    auto launch = [&](int split0, int splits) {
        P.split0 = split0;
        P.n_splits = splits;
        if (two_cta) {
            const int grid = 2 * std::min(n_units * splits, n_workers); // whole clusters
            knn2_tc2_kernel<<<grid, NUM_THREADS, SMEM2_BYTES, ctx->stream>>>(mq, mdb, P);
        } else {
            const int grid = std::min(n_qtiles * splits, ctx->sm_count);
            knn2_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, ctx->stream>>>(mq, mdb, P);
        }
        LAUNCH_COUNT(ctx);
    };
    if (up_chunks > 0) {
        const int64_t split_rows = (int64_t)stages_per_split * bn;
        for (int c = 0; c < up_chunks; c++) {
            const int64_t r0 = c ? db->up_row_end[c - 1] : 0, r1 = db->up_row_end[c];
            const bool last = c == up_chunks - 1;
            CU_TRY(cudaStreamWaitEvent(ctx->stream, db->up_ev[c], 0));
            const int64_t rows_pad = (last ? nd_pad : r1) - r0;
            expand_kernel<<<(unsigned)((rows_pad * 8 + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)db->desc + r0 * 8, r1 - r0, rows_pad,
                                                                                           (uint4 *)(x_db + r0 * ROW_BYTES));
            LAUNCH_COUNT(ctx);
            const int s0 = (int)(r0 / split_rows), s1 = last ? n_splits : (int)(r1 / split_rows);
            launch(s0, s1 - s0);
        }
        CU_TRY(cudaGetLastError());
        std::lock_guard<std::mutex> lock(db->x_mu);
        CU_TRY(cudaEventRecord(db->x_ready, ctx->stream));
        db->x_valid = true;
        db->up_pending = 0;
    } else {
        launch(0, n_splits);
        CU_TRY(cudaGetLastError());
    }
    knn2_tc_merge_kernel<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, ctx->stream>>>(q, db->desc, nq, nd, n_splits, bn, nq_pad, pb, ps, pst, th_low,
                                                                                nnratio, best_idx, best_dist, second_dist, match);
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    return ORBGPU_OK;
}

int knn2_tc_device_init()
{
    int rc = set_max_dyn_smem(knn2_tc_kernel);
    if (rc) return rc;
    return set_max_dyn_smem(knn2_tc2_kernel);
}
