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
// HBM layout: all keyframes of a batch live in one kfset.  At upload a per-keyframe bitonic sort builds
// the CSR (by node id) of the map-point-free features AND node-ordered copies of their descriptors
// (`desc_csr`) and keypoints (`kp_csr`), so that a pair reads two contiguous spans.
// One CTA per pair:
//   phase A  cp.async the two descriptor spans into shared memory (coalesced 16-byte chunks, all loads
//            in flight at once; 16-byte chunk index XOR-swizzled so that 32-byte rows read by
//            consecutive lanes are bank-conflict free);
//   phase B  one warp per shared node (binary-search join of the two sorted node lists); the n1 x n2
//            candidate pairs of the node are flattened over the lanes; XOR+POPC from shared memory;
//            survivors (rare: planted matches) run the fp32 gates and atomicMin a packed key into
//            shared memory;
//   phase C  optional rotation histogram + cull, coalesced write of the match row, match count.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "internal.cuh"

namespace {

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr uint32_t NODE_NONE = 0xFFFFFFFFu;
constexpr int TRI_THREADS = 256;

// one block per keyframe: CSR (by node id) of the features WITHOUT a map point + node-ordered copies
__global__ void __launch_bounds__(1024)
kfset_csr_kernel(int n_feat, int cap, const uint32_t *__restrict__ node_id, const uint8_t *__restrict__ has_mp,
                 const uint4 *__restrict__ desc, const float2 *__restrict__ xy, const int32_t *__restrict__ octave,
                 int32_t *__restrict__ kf_n_nodes, uint32_t *__restrict__ kf_node_ids, int32_t *__restrict__ kf_node_off,
                 int32_t *__restrict__ kf_feat, uint4 *__restrict__ desc_csr, int4 *__restrict__ kp_csr,
                 int32_t *__restrict__ kf_n_free, int32_t *__restrict__ max_free, unsigned char *__restrict__ blob,
                 size_t blob_stride, int32_t *__restrict__ blob_bytes, uint4 *__restrict__ aux)
{
    extern __shared__ unsigned long long keys[]; // [cap]
    __shared__ int s_m;
    __shared__ int warp_sums[32];
    const int kf = blockIdx.x, t = threadIdx.x;
    const uint32_t *nid = node_id + (size_t)kf * n_feat;
    const uint8_t *mp = has_mp + (size_t)kf * n_feat;
    if (t == 0) s_m = 0;
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
    // stream blob of this keyframe (one bulk copy per keyframe in triangulation_stream_kernel):
    // [lo halves m x 16 B][node offsets (nn+1) x 4 B][node ids nn x 4 B]; the hi halves travel with the keypoint
    // in the 32-byte aux record of the slot, which only prefilter survivors ever read
    uint4 *b_lo = (uint4 *)(blob + (size_t)kf * blob_stride);
    int32_t *b_off = (int32_t *)(b_lo + m);
    // node-ordered copies
    for (int i = t; i < m; i += blockDim.x) {
        const int f = (int)(keys[i] & 0xFFFFFFFFull);
        feat[i] = f;
        const size_t src = (size_t)kf * n_feat + f, dst = (size_t)kf * n_feat + i;
        const uint4 lo = desc[2 * src], hi = desc[2 * src + 1];
        desc_csr[2 * dst] = lo;
        desc_csr[2 * dst + 1] = hi;
        b_lo[i] = lo;
        const float2 p = xy[src];
        const int4 kp = make_int4(__float_as_int(p.x), __float_as_int(p.y), octave[src], f);
        kp_csr[dst] = kp;
        aux[2 * dst] = hi;
        aux[2 * dst + 1] = make_uint4((unsigned)kp.x, (unsigned)kp.y, (unsigned)kp.z, (unsigned)kp.w);
    }
    // group heads: thread t owns a contiguous chunk of the sorted keys
    const int per = (m + blockDim.x - 1) / blockDim.x;
    const int s = min(m, t * per), e = min(m, s + per);
    int heads = 0;
    for (int i = s; i < e; i++)
        if (i == 0 || (uint32_t)(keys[i] >> 32) != (uint32_t)(keys[i - 1] >> 32)) heads++;
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
    uint32_t *b_ids = (uint32_t *)(b_off + total + 1);
    for (int i = s; i < e; i++)
        if (i == 0 || (uint32_t)(keys[i] >> 32) != (uint32_t)(keys[i - 1] >> 32)) {
            ids[g] = (uint32_t)(keys[i] >> 32);
            off[g] = i;
            b_ids[g] = (uint32_t)(keys[i] >> 32);
            b_off[g] = i;
            g++;
        }
    if (t == 0) {
        off[total] = m;
        b_off[total] = m;
        kf_n_nodes[kf] = total;
        kf_n_free[kf] = m;
        const int bytes = (16 * m + 8 * total + 4 + 15) & ~15;
        blob_bytes[kf] = bytes;
        atomicMax(max_free, m);
        atomicMax(max_free + 1, total);
        atomicMax(max_free + 2, bytes);
    }
}

struct KfSetView {
    int n_kf, n_feat;
    const float *angle;
    const float *u_right;
    const float *scale_factors;
    const float *level_sigma2;
    const int32_t *kf_n_nodes;
    const uint32_t *kf_node_ids;
    const int32_t *kf_node_off;
    const uint4 *desc_csr;
    const int4 *kp_csr;
    const int32_t *kf_n_free;
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

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
// 16-byte chunk c (= 2*row + half) -> swizzled chunk: rows 4..7 of every 8 swap their halves, so that 8 consecutive
// rows read with LDS.128 touch all 32 banks exactly once
__device__ __forceinline__ int swz(int c) { return c ^ ((c >> 3) & 1); }

constexpr int SURV_CAP = 2048;

// epipole gate (:1191-1203) + epipolar test (:1246) of one surviving candidate, then the (dist, -idx2) reduction
__device__ __forceinline__ void gate_survivor(uint32_t ent, const int4 *__restrict__ kp1, const int4 *__restrict__ kp2,
                                              const float *ur1, const float *ur2, const float *sEp, const float *sF,
                                              const float *__restrict__ scale_factors, const float *__restrict__ level_sigma2,
                                              int coarse, uint32_t *sBest)
{
    const int c1 = (int)(ent & 0x1FFF), c2 = (int)((ent >> 13) & 0x1FFF), dist = (int)(ent >> 26);
    const int4 q2 = kp2[c2];
    const int4 q1 = kp1[c1];
    const float x2 = __int_as_float(q2.x), y2 = __int_as_float(q2.y);
    bool st1 = false, st2 = false;
    if (ur1) {
        st1 = ur1[q1.w] >= 0.f;
        st2 = ur2[q2.w] >= 0.f;
    }
    if (!st1 && !st2) {
        const float dx = __fsub_rn(sEp[0], x2), dy = __fsub_rn(sEp[1], y2);
        if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, scale_factors[q2.z])) return;
    }
    if (!coarse && !epipolar_ok(sF, __int_as_float(q1.x), __int_as_float(q1.y), x2, y2, level_sigma2[q2.z])) return;
    atomicMin(&sBest[c1], ((uint32_t)dist << 20) | (0xFFFFFu - (uint32_t)q2.w));
}

template <bool HAS_RIGHT>
__global__ void __launch_bounds__(TRI_THREADS)
triangulation_pairs_kernel(KfSetView s, int n_pairs, int max_free, int max_nodes, const int32_t *__restrict__ kf1, const int32_t *__restrict__ kf2,
                           const float *__restrict__ ep, const float *__restrict__ f12, int only_stereo, int coarse, int check_ori,
                           int32_t *__restrict__ matches12, int32_t *__restrict__ nmatches, unsigned long long *__restrict__ counters)
{
    extern __shared__ uint4 sm_raw[];
    uint4 *sD1 = sm_raw;                      // [max_free][2] swizzled
    uint4 *sD2 = sm_raw + 2 * (size_t)max_free;
    uint32_t *sBest = (uint32_t *)(sm_raw + 4 * (size_t)max_free); // [max_free] key per CSR slot of keyframe 1
    int4 *sNode = (int4 *)(sBest + max_free);                      // [max_nodes] {s1, n1f, s2, n2f} of the joined nodes
    uint8_t *sBin = (uint8_t *)(sNode + max_nodes);                // [max_free]
    __shared__ int hist[ORBGPU_HISTO_LENGTH];
    __shared__ int ind[3];
    __shared__ float sF[9];
    __shared__ float sEp[2];
    __shared__ int s_count, s_nsurv;
    __shared__ uint32_t sSurv[SURV_CAP];
    const int p = blockIdx.x;
    if (p >= n_pairs) return;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = TRI_THREADS / 32;
    const int n = s.n_feat;
    const int k1 = kf1[p], k2 = kf2[p];
    const int m1 = s.kf_n_free[k1], m2 = s.kf_n_free[k2];
    const uint4 *g1 = s.desc_csr + (size_t)k1 * n * 2, *g2 = s.desc_csr + (size_t)k2 * n * 2;
    // ---- phase A: stage both descriptor spans (all chunks in flight)
    for (int c = t; c < 2 * m1; c += TRI_THREADS) cp_async16(&sD1[swz(c)], &g1[c]);
    for (int c = t; c < 2 * m2; c += TRI_THREADS) cp_async16(&sD2[swz(c)], &g2[c]);
    asm volatile("cp.async.commit_group;\n" ::);
    int32_t *row = matches12 + (size_t)p * n;
    for (int i = t; i < n; i += TRI_THREADS) row[i] = -1; // :1092 vMatches12(N, -1)
    for (int i = t; i < m1; i += TRI_THREADS) sBest[i] = KEY_NONE;
    if (t < ORBGPU_HISTO_LENGTH) hist[t] = 0;
    if (t < 9) sF[t] = f12[9 * (size_t)p + t];
    if (t < 2) sEp[t] = ep[2 * (size_t)p + t];
    if (t == 0) { s_count = 0; s_nsurv = 0; }
    const int4 *kp1 = s.kp_csr + (size_t)k1 * n, *kp2 = s.kp_csr + (size_t)k2 * n;
    const float *ur1 = s.u_right ? s.u_right + (size_t)k1 * n : nullptr, *ur2 = s.u_right ? s.u_right + (size_t)k2 * n : nullptr;
    const int nn1 = s.kf_n_nodes[k1], nn2 = s.kf_n_nodes[k2];
    const uint32_t *ids1 = s.kf_node_ids + (size_t)k1 * n, *ids2 = s.kf_node_ids + (size_t)k2 * n;
    const int32_t *off1 = s.kf_node_off + (size_t)k1 * (n + 1), *off2 = s.kf_node_off + (size_t)k2 * (n + 1);

    // ---- join: thread a looks node a of keyframe 1 up in keyframe 2's sorted node list (merge-join :1113-1292)
    for (int a = t; a < nn1; a += TRI_THREADS) {
        const uint32_t nid = ids1[a];
        int lo = 0, hi = nn2;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (ids2[mid] < nid) lo = mid + 1; else hi = mid;
        }
        int4 e = make_int4(0, 0, 0, 0);
        if (lo < nn2 && ids2[lo] == nid) {
            const int s1 = off1[a], s2 = off2[lo];
            e = make_int4(s1, off1[a + 1] - s1, s2, off2[lo + 1] - s2);
        }
        sNode[a] = e;
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();

    // ---- phase B
    unsigned long long ncmp = 0;
    for (int a = warp; a < nn1; a += nwarps) {
        const int4 e = sNode[a];
        const int s1 = e.x, n1f = e.y, s2 = e.z, n2f = e.w;
        if (n2f == 0 || n1f == 0 || (only_stereo && !HAS_RIGHT)) continue;
        // the n1f x n2f candidate pairs of the node are flattened over the lanes: pair index tp = lane, lane+32, ...
        // (i1, i2) = (tp / n2f, tp % n2f) is advanced incrementally (one division per node); trip counts are
        // warp-uniform so that the warp stays converged from one node to the next
        const int total = n1f * n2f;
        const int q32 = 32 / n2f, r32 = 32 - q32 * n2f;
        int i1 = lane / n2f, i2 = lane - i1 * n2f;
        for (int tp0 = 0; tp0 < total; tp0 += 32) {
            bool ok = (tp0 + lane) < total;
            const int c1 = s1 + (ok ? i1 : 0), c2 = s2 + (ok ? i2 : 0); // CSR slots
            if (HAS_RIGHT) {
                const bool st1 = ur1[kp1[c1].w] >= 0.f, st2 = ur2[kp2[c2].w] >= 0.f; // :1134, :1168
                if (only_stereo && (!st1 || !st2)) ok = false;                        // :1136-1138, :1170-1172
            }
            // first 128 bits; a random pair is already above TH_LOW here 99 % of the time
            int dist = ham128(sD1[swz(2 * c1)], sD2[swz(2 * c2)]);
            ncmp += ok ? 1 : 0;
            if (__any_sync(FULL_MASK, ok && dist <= ORBGPU_TH_LOW)) {
                dist += ham128(sD1[swz(2 * c1 + 1)], sD2[swz(2 * c2 + 1)]);
                if (ok && dist <= ORBGPU_TH_LOW) { // :1180 (bestDist starts at TH_LOW); rare: planted matches only
                    // the fp32 gates need keypoint data from global memory: defer them so that all survivors of the
                    // pair are gated in parallel instead of one lane at a time
                    const int slot = atomicAdd(&s_nsurv, 1);
                    const uint32_t ent = (uint32_t)c1 | ((uint32_t)c2 << 13) | ((uint32_t)dist << 26);
                    if (slot < SURV_CAP) sSurv[slot] = ent;
                    else gate_survivor(ent, kp1, kp2, ur1, ur2, sEp, sF, s.scale_factors, s.level_sigma2, coarse, sBest);
                }
            }
            i1 += q32;
            i2 += r32;
            if (i2 >= n2f) { i2 -= n2f; i1++; }
        }
        __syncwarp();
    }
    __syncthreads();
    {
        const int ns = min(s_nsurv, SURV_CAP);
        for (int i = t; i < ns; i += TRI_THREADS)
            gate_survivor(sSurv[i], kp1, kp2, ur1, ur2, sEp, sF, s.scale_factors, s.level_sigma2, coarse, sBest);
    }
    __syncthreads();
    // ---- phase C
    const float *ang1 = s.angle + (size_t)k1 * n, *ang2 = s.angle + (size_t)k2 * n;
    int mine = 0;
    if (check_ori) { // :1266-1277
        for (int c = t; c < m1; c += TRI_THREADS) {
            const uint32_t key = sBest[c];
            int bin = 255;
            if (key != KEY_NONE) {
                const int idx2 = (int)(0xFFFFFu - (key & 0xFFFFFu));
                bin = rot_bin(ang1[kp1[c].w], ang2[idx2]);
                if (bin >= 0 && bin < ORBGPU_HISTO_LENGTH) atomicAdd(&hist[bin], 1); else bin = 254;
            }
            sBin[c] = (uint8_t)bin;
        }
        __syncthreads();
        if (t == 0) three_maxima(hist, ORBGPU_HISTO_LENGTH, ind[0], ind[1], ind[2]);
        __syncthreads();
    }
    for (int c = t; c < m1; c += TRI_THREADS) {
        const uint32_t key = sBest[c];
        if (key == KEY_NONE) continue;
        if (check_ori) { // :1295-1314
            const int b = sBin[c];
            if (b < ORBGPU_HISTO_LENGTH && b != ind[0] && b != ind[1] && b != ind[2]) continue;
        }
        row[kp1[c].w] = (int)(0xFFFFFu - (key & 0xFFFFFu));
        mine++;
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


// =========================================================================================
// Engine 2: persistent, warp-specialised pipeline for monocular keyframe sets.
//
// One CTA per SM walks pairs p = blockIdx.x, += gridDim.x through a ring of S stages (S = 2..4, as shared memory
// allows).  A stage holds the two keyframes' stream blobs (lo halves | node offsets | node ids, ~17 KB each), the
// join table, the per-slot best keys and the candidate list of ONE pair.  Four roles, chained by mbarriers:
//
//   producer (1 warp)  waits empty[st]; reads the pair's metadata; ONE cp.async.bulk per keyframe -> full[st] (tx bytes)
//   join     (NJ warps) waits full[st]; node a of keyframe 1 binary-searched in
//                      keyframe 2's node list (merge-join :1113-1292); writes sCand[c1] = (first candidate, count)
//                      for every slot -> joined[st]
//   compare  (NC warps) waits joined[st]; 32-slot chunks dealt round-robin; lane owns CSR slot c1 of keyframe 1 and
//                      walks its node's candidates with a branch-free loop: 1 LDS.128 + 4 XOR + 4 POPC on the lo
//                      halves, bit j of a per-slot mask = candidate j is already <= TH_LOW (2 % of them are).
//                      Slots with a non-zero mask are compacted into the stage's list (one ballot per chunk) and
//                      the aux records of their candidates are prefetched into L2 -> compared[st]
//   post     (NG warps) waits compared[st]; one (slot, flagged candidate) entry per thread, ONE pass: fetches the 32-byte
//                      aux records {hi half, keypoint} of both, finishes the distance, fp32 gates, keeps the
//                      (dist, -idx2) minimum per slot and writes the match -> empty[st]
//
// The compare warps never meet a CTA-wide barrier: while they stream pair i, the post warps finish pair i-1, the join
// warps prepare pair i+1 and the bulk copies of pair i+2 are in flight.
constexpr int TS_OVF = 256; // (slot, candidate) entries of nodes with more than 32 candidates
constexpr uint32_t ENT_NONE = 0xFFFFFFFFu;
#ifdef TS_NO_PREFETCH
#define TS_PREFETCH(...) ((void)0)
#else
#define TS_PREFETCH(...) asm volatile(__VA_ARGS__)
#endif
struct TsStageCtl {
    int k1, k2, m1, m2, nn1, nn2;
    int n_list;     // slots with a non-zero mask listed so far
    int n_ovf;      // entries in the overflow list: second and later flagged candidates of a slot, candidates past the 32nd of a node
    int n_long;     // != 0: some node had more than 32 candidates (a slot can then have survivors without being listed)
    float geo[12];  // f12[9], ep[2]
};

__device__ __forceinline__ uint32_t ts_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ts_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void ts_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ts_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Blocking wait: try_wait with a suspend-time hint parks the warp in hardware instead of re-polling -- the polls of the
// waiting roles (post, join, producer and the compare warps between pairs) were 40 % of all issued instructions and took
// issue slots from the compare loop.
__device__ __forceinline__ void ts_mbar_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TS_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra TS_WAIT_DONE;\n\t"
        "bra TS_WAIT_LOOP;\n\t"
        "TS_WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
}
__device__ __forceinline__ void ts_mbar_wait(uint32_t bar, uint32_t parity) { ts_mbar_wait_hint(bar, parity, 4000u); }
// the roles that normally wait (producer, join, post).  The window's length is not critical: 300 ns ... 20 us measured within the
// run-to-run spread (0.116-0.122 ms per 4096 pairs on one box); what matters is that the warp is parked rather than polling.
__device__ __forceinline__ void ts_mbar_wait_relaxed(uint32_t bar, uint32_t parity) { ts_mbar_wait_hint(bar, parity, 1000u); }
__device__ __forceinline__ void ts_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void ts_bar_post(int nthreads) { asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory"); }

struct TsParams {
    const unsigned char *blob;
    size_t blob_stride;
    const int32_t *blob_bytes;
    const int32_t *kf_n_free;
    const int32_t *kf_n_nodes;
    const uint4 *aux; // [n_kf][n_feat][2] {hi half, keypoint {x, y, octave, feature id}} in CSR slot order
    const float *angle;
    const float *scale_factors;
    const float *level_sigma2;
    int n_feat, n_levels, cap_bytes, max_free, n_pairs, n_stages, stage_bytes;
    const int32_t *kf1, *kf2;
    const float *ep, *f12;
    int coarse, check_ori;
    // output targets: the match rows / counts of pair p go to row (pair0 + p) of EVERY target.  One target (the caller's own
    // buffers) for the plain call; the buffers of all ranks (peer memory over NVLink) for the fused all-gather, where the rows
    // were preset to -1 by their owners and only matches and counts cross the links.
    int n_targets, rows_preset; // rows_preset 2: "copy rows" -- target 0 is this rank's own buffer, built as in the plain call; the
                                // finished 8 KB row is then copied to the other targets with coalesced 128-bit stores (no preset needed)
                                // rows_preset 3: "compact pairs" -- the reference's vMatchedPairs form (:1317-1325): the matches of pair
                                // p, ascending idx1, as (idx1 << 16 | idx2) entries at tgt_m[r] + (pair0 + p) * n_feat; only the
                                // counts[p] valid entries (rounded up to 16 bytes) are stored, to every target, straight from shared memory
    long long pair0;
    int32_t *tgt_m[8], *tgt_nm[8];
    // compact mode, completion protocol of the fused all-gather: every CTA adds its pairs to `done` when all of them have been
    // shipped; the one that completes the batch publishes this rank's epoch in slot src_rank of every target's flag array
    // (system-scope fences order the pair stores before it)
    uint32_t *tgt_flag[8]; // target 0 = this rank's own flag array
    uint32_t *epoch_done;  // [0] current epoch, [1] pairs shipped so far; both advanced / cleared by the CTA that completes the batch
    uint32_t *status;      // [0] = 1 when a peer did not arrive within the time-out (may be null)
    uint32_t wait_mask;    // source ranks whose epoch the completing CTA waits for before the kernel ends
    int src_rank;
    unsigned long long *counters;
    long long *timeline; // debug: [n_my of CTA 0][16] SM-clock stamps, or null
};

// One prefilter survivor (slot c1 of keyframe 1 with its aux record {h1, q1} already loaded, slot c2 of keyframe 2):
// finish the distance (:1180; lo halves from the stage, hi halves from the aux records), epipole gate (:1191-1203,
// monocular), epipolar test (:1246).  Returns the (dist, -idx2) key, KEY_NONE when rejected.
__device__ __forceinline__ uint32_t ts_gate_loaded(const uint4 a_lo, const uint4 h1, const uint4 q1, const uint4 b_lo, const uint4 h2,
                                                   const uint4 q2, const float *geo, const float *sScale, const float *sSigma,
                                                   int coarse)
{
    const int dist = ham128(a_lo, b_lo) + ham128(h1, h2);
    if (dist > ORBGPU_TH_LOW) return KEY_NONE;
    const float x2 = __uint_as_float(q2.x), y2 = __uint_as_float(q2.y);
    const float dx = __fsub_rn(geo[9], x2), dy = __fsub_rn(geo[10], y2);
    if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, sScale[q2.z])) return KEY_NONE;
    if (!coarse && !epipolar_ok(geo, __uint_as_float(q1.x), __uint_as_float(q1.y), x2, y2, sSigma[q2.z])) return KEY_NONE;
    return ((uint32_t)dist << 20) | (0xFFFFFu - q2.w);
}
__device__ __forceinline__ uint32_t ts_gate(const uint4 a_lo, const uint4 h1, const uint4 q1, int c2, const uint4 *lo2,
                                            const uint4 *__restrict__ aux2, const float *geo, const float *sScale,
                                            const float *sSigma, int coarse)
{
    return ts_gate_loaded(a_lo, h1, q1, lo2[c2], aux2[2 * c2], aux2[2 * c2 + 1], geo, sScale, sSigma, coarse);
}

template <int NC, int NG, int NJ, int NGRP>
__global__ void __launch_bounds__((NC + NG + NJ + 1) * 32, 1) triangulation_stream_kernel(const TsParams P)
{
    constexpr int JT = NJ * 32;
    extern __shared__ __align__(128) unsigned char ts_smem[];
    __shared__ __align__(8) unsigned long long bars[4][4]; // [stage][full, joined, compared, empty]
    __shared__ TsStageCtl ctl[4];
    __shared__ float sScale[64], sSigma[64];
    __shared__ int s_cnt[2];    // post groups: matches of the pair in flight
    __shared__ unsigned long long s_ncmp; // gather mode: this CTA's comparisons (added to the global counter by the post leader)
    __shared__ int sBucket[64]; // join warps: slots per candidate-count bucket, then the buckets' bases
    __shared__ int hist2[2][ORBGPU_HISTO_LENGTH + 2];
    __shared__ int ind2[2][4];

    const int S = P.n_stages, cap = P.cap_bytes, mf = P.max_free;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int n = P.n_feat;
    const bool compact = P.rows_preset == 3;
    const int n_my = (P.n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t bar0 = ts_smem_u32(&bars[0][0]);
    auto bar_of = [&](int st, int which) { return bar0 + (uint32_t)(st * 4 + which) * 8u; };
    auto stamp = [&](int i, int ev) { if (P.timeline && blockIdx.x == 0) P.timeline[i * 16 + ev] = clock64(); };
    enum { B_FULL = 0, B_JOINED = 1, B_COMPARED = 2, B_EMPTY = 3 };

    if (t == 0) {
        for (int st = 0; st < 4; st++) {
            ts_mbar_init(bar_of(st, B_FULL), 1);
            ts_mbar_init(bar_of(st, B_JOINED), NJ);
            ts_mbar_init(bar_of(st, B_COMPARED), NC);
            ts_mbar_init(bar_of(st, B_EMPTY), NG / NGRP);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t == 0) s_ncmp = 0;
    if (t < 64) {
        sScale[t] = t < P.n_levels ? P.scale_factors[t] : 0.f;
        sSigma[t] = t < P.n_levels ? P.level_sigma2[t] : 0.f;
    }
    __syncthreads();

    // per-stage carve-up of the dynamic shared memory
    auto stage_base = [&](int st) { return ts_smem + (size_t)st * P.stage_bytes; };
    // compact mode: behind the stages, per post group a bitmap of the matched features and its per-word prefix counts
    const int n_words = (n + 31) >> 5;
    uint32_t *bm_all = (uint32_t *)(ts_smem + (size_t)S * P.stage_bytes);
    // [A cap][B cap][sCand mf x4][sMask mf x4][sBest mf x4][sList mf x2][sPerm mf x2][sOvf TS_OVF x4][sNodeE max_nodes x4]

    if (warp == NC + NG + NJ) {
        // ---------------- producer warp: the metadata of the next pair is fetched while the ring is still full, so
        // that only the bulk copies themselves follow the empty[] signal
        struct Next { int k1, k2, b1, b2, m1, m2, nn1, nn2; float g; };
        auto fetch = [&](int i) {
            Next x;
            x.g = 0.f;
            x.k1 = x.k2 = x.b1 = x.b2 = x.m1 = x.m2 = x.nn1 = x.nn2 = 0;
            if (i < n_my) {
                const int p = blockIdx.x + i * gridDim.x;
                x.k1 = P.kf1[p]; x.k2 = P.kf2[p];
                if (lane < 9) x.g = P.f12[9 * (size_t)p + lane];
                else if (lane < 11) x.g = P.ep[2 * (size_t)p + lane - 9];
                x.b1 = P.blob_bytes[x.k1]; x.b2 = P.blob_bytes[x.k2];
                x.m1 = P.kf_n_free[x.k1]; x.m2 = P.kf_n_free[x.k2];
                x.nn1 = P.kf_n_nodes[x.k1]; x.nn2 = P.kf_n_nodes[x.k2];
            }
            return x;
        };
        Next nx = fetch(0);
        for (int i = 0, st = 0, round = 0; i < n_my; i++) {
            const Next cur = nx;
            nx = fetch(i + 1);
            if (round > 0) ts_mbar_wait_relaxed(bar_of(st, B_EMPTY), (round - 1) & 1);
            if (lane == 0) stamp(i, 0);
            TsStageCtl &C = ctl[st];
            if (lane < 11) C.geo[lane] = cur.g;
            __syncwarp();
            if (lane == 0) {
                C.k1 = cur.k1; C.k2 = cur.k2;
                C.m1 = cur.m1; C.m2 = cur.m2;
                C.nn1 = cur.nn1; C.nn2 = cur.nn2;
                C.n_list = 0; C.n_ovf = 0; C.n_long = 0;
                const uint32_t bar = bar_of(st, B_FULL);
                ts_mbar_expect_tx(bar, (uint32_t)(cur.b1 + cur.b2));
                unsigned char *dst = stage_base(st);
                ts_bulk_load(ts_smem_u32(dst), P.blob + (size_t)cur.k1 * P.blob_stride, (uint32_t)cur.b1, bar);
                ts_bulk_load(ts_smem_u32(dst + cap), P.blob + (size_t)cur.k2 * P.blob_stride, (uint32_t)cur.b2, bar);
            }
            __syncwarp();
            if (++st == S) { st = 0; round++; }
        }
        return;
    }

    if (warp >= NC + NG) {
        // ---------------- join warps
        const int jt = (warp - (NC + NG)) * 32 + lane;
        for (int i = 0, st = 0, round = 0; i < n_my; i++) {
            ts_mbar_wait_relaxed(bar_of(st, B_FULL), round & 1);
            if (jt == 0) stamp(i, 1);
            const TsStageCtl &C = ctl[st];
            const int m1 = C.m1, m2 = C.m2, nn1 = C.nn1, nn2 = C.nn2;
            unsigned char *base = stage_base(st);
            const int32_t *off1 = (const int32_t *)((const uint4 *)base + m1);
            const uint32_t *ids1 = (const uint32_t *)(off1 + nn1 + 1);
            const int32_t *off2 = (const int32_t *)((const uint4 *)(base + cap) + m2);
            const uint32_t *ids2 = (const uint32_t *)(off2 + nn2 + 1);
            uint32_t *sCand = (uint32_t *)(base + 2 * (size_t)cap), *sBest = sCand + 2 * mf;
            uint16_t *sPerm = (uint16_t *)(sBest + mf) + mf;          // after the (16-bit) slot list
            uint32_t *sNodeE = (uint32_t *)(sPerm + mf) + TS_OVF;     // after the overflow list: per node (s2 | n2f << 16)
            if (jt < 64) sBucket[jt] = 0;
            asm volatile("bar.sync 4, %0;" ::"r"(JT) : "memory");
            for (int a = jt; a < nn1; a += JT) {
                const uint32_t nid = ids1[a];
                int lo = 0, hi = nn2;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (ids2[mid] < nid) lo = mid + 1; else hi = mid;
                }
                uint32_t e = 0;
                if (lo < nn2 && ids2[lo] == nid) {
                    const int s2 = off2[lo];
                    e = (uint32_t)s2 | ((uint32_t)(off2[lo + 1] - s2) << 16);
                }
                sNodeE[a] = e;
                atomicAdd(&sBucket[min((int)(e >> 16), 63)], off1[a + 1] - off1[a]);
            }
            asm volatile("bar.sync 4, %0;" ::"r"(JT) : "memory");
            // compare order: nodes bucketed by their candidate count (descending), so that the 32 slots a compare warp
            // takes at a time walk candidate lists of (nearly) the same length.  Counting sort: slots per bucket above,
            // exclusive scan over the 64 buckets by one warp, then every node reserves its range in its bucket (the order
            // of equal-count nodes is irrelevant: results do not depend on the compare order).
            if (jt < 32) {
                const int v0 = sBucket[63 - 2 * jt], v1 = sBucket[62 - 2 * jt];
                int incl = v0 + v1;
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(FULL_MASK, incl, o);
                    if (jt >= o) incl += u;
                }
                const int excl = incl - v0 - v1;
                sBucket[63 - 2 * jt] = excl;
                sBucket[62 - 2 * jt] = excl + v0;
            }
            asm volatile("bar.sync 4, %0;" ::"r"(JT) : "memory");
            for (int a = jt; a < nn1; a += JT) {
                const uint32_t e = sNodeE[a];
                const int s1 = off1[a], e1 = off1[a + 1];
                const int pos = atomicAdd(&sBucket[min((int)(e >> 16), 63)], e1 - s1);
                for (int c = s1; c < e1; c++) {
                    sCand[c] = e;
                    sBest[c] = KEY_NONE;
                    sPerm[pos + (c - s1)] = (uint16_t)c;
                }
            }
            asm volatile("bar.sync 4, %0;" ::"r"(JT) : "memory"); // sBucket is reused by the next pair
            __syncwarp();
            if (jt == 0) stamp(i, 2);
            if (lane == 0) ts_mbar_arrive(bar_of(st, B_JOINED));
            if (++st == S) { st = 0; round++; }
        }
        return;
    }

    if (warp < NC) {
        // ---------------- compare warps
        unsigned long long ncmp = 0;
        for (int i = 0, st = 0, round = 0; i < n_my; i++) {
            ts_mbar_wait(bar_of(st, B_JOINED), round & 1);
            if (t == 0) stamp(i, 3);
            TsStageCtl &C = ctl[st];
            const int k1 = C.k1, k2 = C.k2, m1 = C.m1;
            unsigned char *base = stage_base(st);
            const uint4 *lo1 = (const uint4 *)base, *lo2 = (const uint4 *)(base + cap);
            const uint32_t *sCand = (const uint32_t *)(base + 2 * (size_t)cap);
            uint32_t *sMask = (uint32_t *)sCand + mf, *sBest = sMask + mf;
            uint16_t *sList = (uint16_t *)(sBest + mf);
            const uint16_t *sPerm = sList + mf;
            uint32_t *sOvf = (uint32_t *)(sPerm + mf);
            const uint4 *aux1 = P.aux + (size_t)k1 * n * 2, *aux2 = P.aux + (size_t)k2 * n * 2;
            // 32-slot chunks are dealt to the warps in snake order (rounds alternate direction): the slots are sorted by candidate
            // count, descending, so warp w's chunks w, 2 NC - 1 - w, 2 NC + w, ... add up to nearly the same number of trips for
            // every warp (plain round-robin gives the first warp the two longest chunks of their rounds, the last the two shortest).
            // Rotated from pair to pair so that the odd chunk does not always land on the same warps.
            const int wr = (warp + NC - (i % NC)) % NC;
            for (int r = 0, c0 = wr * 32; c0 < m1; r++, c0 = (r * NC + ((r & 1) ? NC - 1 - wr : wr)) * 32) {
                int c1 = c0 + lane;
                uint32_t cand = 0;
                uint4 a_lo = make_uint4(0, 0, 0, 0);
                if (c1 < m1) {
                    c1 = sPerm[c1]; // position in compare order -> CSR slot
                    cand = sCand[c1];
                    a_lo = lo1[c1];
                }
                const int n2f = (int)(cand >> 16), s2 = (int)(cand & 0xFFFF);
                ncmp += (unsigned)n2f;
                // branch-free inner loop: bit j of mask = candidate j passes the 128-bit prefilter
                const uint4 *pb = lo2 + s2, *pe = pb + min(n2f, 32);
                uint32_t mask = 0, bit = 1;
#pragma unroll 2
                for (; pb < pe; ++pb) {
                    if (ham128(a_lo, *pb) <= ORBGPU_TH_LOW) mask |= bit; // a random pair fails here 99.6 % of the time
                    bit <<= 1;
                }
                __syncwarp();
                // list the slots that have candidates left (one reservation per chunk), start pulling their aux records
                const unsigned bal = __ballot_sync(FULL_MASK, mask != 0);
                if (bal) {
                    int slot0 = 0;
                    if (lane == 0) slot0 = atomicAdd(&C.n_list, __popc(bal));
                    slot0 = __shfl_sync(FULL_MASK, slot0, 0);
                    if (mask) {
                        // the slot is listed with its FIRST flagged candidate; a second one is rare (a false positive of the 128-bit
                        // prefilter next to the true match: ~7 % of the listed slots) and goes to the overflow list as its own
                        // (slot, candidate) entry, so that the post warps gate every entry in ONE pass
                        sMask[c1] = mask;
                        sList[slot0 + __popc(bal & lanemask_lt())] = (uint16_t)c1;
                        TS_PREFETCH("prefetch.global.L2 [%0];" ::"l"(aux1 + 2 * c1));
                        TS_PREFETCH("prefetch.global.L2 [%0];" ::"l"(aux2 + 2 * (s2 + __ffs(mask) - 1)));
                        mask &= mask - 1;
                        while (mask) {
                            const int c2 = s2 + __ffs(mask) - 1;
                            mask &= mask - 1;
                            TS_PREFETCH("prefetch.global.L2 [%0];" ::"l"(aux2 + 2 * c2));
                            const int slot = atomicAdd(&C.n_ovf, 1);
                            if (slot < TS_OVF) sOvf[slot] = (uint32_t)c1 | ((uint32_t)c2 << 13);
                            else { // list full (adversarial inputs only): gate in place
                                const uint32_t key = ts_gate(a_lo, aux1[2 * c1], aux1[2 * c1 + 1], c2, lo2, aux2, C.geo, sScale, sSigma, P.coarse);
                                if (key != KEY_NONE) atomicMin(&sBest[c1], key);
                            }
                        }
                    }
                }
                // a node with more than 32 candidates (rare with a real vocabulary): the rest goes through the overflow list
                if (n2f > 32) C.n_long = 1;
                for (int j = 32; j < n2f; j++) {
                    if (ham128(a_lo, lo2[s2 + j]) > ORBGPU_TH_LOW) continue;
                    const int slot = atomicAdd(&C.n_ovf, 1);
                    if (slot < TS_OVF) sOvf[slot] = (uint32_t)c1 | ((uint32_t)(s2 + j) << 13);
                    else { // list full (adversarial inputs only): gate in place
                        const uint32_t key = ts_gate(a_lo, aux1[2 * c1], aux1[2 * c1 + 1], s2 + j, lo2, aux2, C.geo, sScale, sSigma, P.coarse);
                        if (key != KEY_NONE) atomicMin(&sBest[c1], key);
                    }
                }
                __syncwarp();
            }
            __syncwarp();
            if (t == 0) stamp(i, 4);
            if (P.epoch_done && i == n_my - 1) { // gather mode: the CTA's total is complete when the last pair has been compared
                for (int o = 16; o; o >>= 1) ncmp += __shfl_xor_sync(FULL_MASK, ncmp, o);
                if (lane == 0 && ncmp) atomicAdd(&s_ncmp, ncmp);
            }
            if (lane == 0) ts_mbar_arrive(bar_of(st, B_COMPARED));
            if (++st == S) { st = 0; round++; }
        }
        if (!P.epoch_done) {
            for (int o = 16; o; o >>= 1) ncmp += __shfl_xor_sync(FULL_MASK, ncmp, o);
            if (lane == 0 && ncmp) atomicAdd(&P.counters[0], ncmp);
        }
        return;
    }

    // ---------------- post warps: two groups of NG/2 warps take alternate pairs, so that each group has two compare
    // periods to cover the round trips of its pair's gating step
    constexpr int GT = (NG / NGRP) * 32;
    const int grp = (warp - NC) / (NG / NGRP), gt = (warp - NC - grp * (NG / NGRP)) * 32 + lane;
    const int bar_id = 2 + grp;
    auto bar_post = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(GT) : "memory"); };
    int *hist = hist2[grp], *ind = ind2[grp];
    for (int i = 0, st = 0, round = 0; i < n_my; i++, st = (st + 1 == S ? 0 : st + 1), round += (st == 0)) {
        if ((i % NGRP) != grp) continue;
        const int p = blockIdx.x + i * gridDim.x;
        const size_t row_off = (size_t)(P.pair0 + p) * n;
        // vMatches12(N, -1) (:1092), while the pair is still being compared
        const int n_scatter = P.rows_preset == 2 ? 1 : P.n_targets; // targets that receive individual match stores
        uint32_t *bm = bm_all + (size_t)grp * 2 * n_words, *wbase = bm + n_words;
        if (compact) {
            for (int x = gt; x < n_words; x += GT) bm[x] = 0;
        } else if (P.rows_preset != 1) {
            for (int r = 0; r < n_scatter; r++) {
                int32_t *row = P.tgt_m[r] + row_off;
                if ((n & 3) == 0) {
                    int4 *row4 = (int4 *)row;
                    for (int x = gt; x < (n >> 2); x += GT) row4[x] = make_int4(-1, -1, -1, -1);
                } else {
                    for (int x = gt; x < n; x += GT) row[x] = -1;
                }
            }
        }
        auto put_match = [&](int f1, int idx2) {
            for (int r = 0; r < n_scatter; r++) P.tgt_m[r][row_off + f1] = idx2;
        };
        if (gt == 0) s_cnt[grp] = 0;
        if (P.check_ori && gt < ORBGPU_HISTO_LENGTH) hist[gt] = 0;
        bar_post(); // the row is initialised before any thread of the group writes a match into it
        if (gt == 0) stamp(i, 5);
        ts_mbar_wait_relaxed(bar_of(st, B_COMPARED), round & 1);
        if (gt == 0) stamp(i, 6);
        TsStageCtl &C = ctl[st];
        const int k1 = C.k1, k2 = C.k2, m1 = C.m1, ns = C.n_list;
        unsigned char *base = stage_base(st);
        const uint4 *lo1 = (const uint4 *)base, *lo2 = (const uint4 *)(base + cap);
        const uint32_t *sCand = (const uint32_t *)(base + 2 * (size_t)cap), *sMask = sCand + mf;
        uint32_t *best = (uint32_t *)sMask + mf;
        const uint16_t *sList = (const uint16_t *)(best + mf);
        uint32_t *sOvf = (uint32_t *)(sList + 2 * mf);
        const uint4 *aux1 = P.aux + (size_t)k1 * n * 2, *aux2 = P.aux + (size_t)k2 * n * 2;
        // the feature id of a listed slot replaces its (consumed) mask; compact mode: its entry does, and the ordered list is built
        // over the (dead) join table
        uint32_t *sEnt = (uint32_t *)sMask, *sOut = (uint32_t *)sCand;
        auto mark = [&](int slot, int f1, int idx2) {
            sEnt[slot] = ((uint32_t)f1 << 16) | (uint32_t)idx2;
            atomicOr(&bm[f1 >> 5], 1u << (f1 & 31));
        };
        int mine = 0;
        // ---- gate: one entry per thread -- the listed slots with their first flagged candidate, then the overflow entries (further
        // candidates of a slot, candidates past the 32nd of a node).  The per-slot minimum (last-wins ties of :1180) is collected in sBest.
        const int n_ovf = min(C.n_ovf, TS_OVF), n_ent = ns + n_ovf;
        auto entry = [&](int e, int &c1, int &ca) {
            if (e < ns) {
                c1 = (int)sList[e];
                ca = (int)(sCand[c1] & 0xFFFF) + __ffs(sMask[c1]) - 1;
            } else {
                const uint32_t en = sOvf[e - ns];
                c1 = (int)(en & 0x1FFF);
                ca = (int)((en >> 13) & 0x1FFF);
            }
        };
        if (gt + GT < n_ent) { // ~300 entries for 192 threads: the records of this thread's second entry are pulled into L1 while
            int c1, ca;        // the first is gated (holding both in registers spills: 64 registers per thread at 1024 threads)
            entry(gt + GT, c1, ca);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(aux1 + 2 * c1));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(aux2 + 2 * ca));
        }
        for (int e = gt; e < n_ent; e += GT) {
            int c1, ca;
            const bool first = e < ns;
            entry(e, c1, ca);
            const uint4 h1 = aux1[2 * c1], q1 = aux1[2 * c1 + 1];
            const uint4 ha = aux2[2 * ca], qa = aux2[2 * ca + 1];
            const uint32_t key = ts_gate_loaded(lo1[c1], h1, q1, lo2[ca], ha, qa, C.geo, sScale, sSigma, P.coarse);
            if (key != KEY_NONE) atomicMin(&best[c1], key);
            if (first) sEnt[c1] = q1.w; // feature id of the slot, for the output pass
        }
        bar_post();
        if (gt == 0) { stamp(i, 10); if (P.timeline && blockIdx.x == 0) { P.timeline[i * 16 + 12] = ns; P.timeline[i * 16 + 13] = n_ovf; P.timeline[i * 16 + 14] = C.n_long; } }
        // ---- output by the slots' owners: the listed slots, or every slot when the compare warps found a node with more than 32
        // candidates (such a slot can have survivors without being listed)
        const bool all_slots = C.n_long != 0;
        const int n_own = all_slots ? m1 : ns;
        const float *ang1 = P.angle + (size_t)k1 * n, *ang2 = P.angle + (size_t)k2 * n;
        auto slot_of = [&](int e) { return all_slots ? e : (int)sList[e]; };
        auto feat_of = [&](int c) { return all_slots ? (int)aux1[2 * c + 1].w : (int)sEnt[c]; };
        if (P.check_ori) { // :1266-1277, :1295-1314
            for (int e = gt; e < n_own; e += GT) {
                const int c = slot_of(e);
                const uint32_t key = best[c];
                if (key == KEY_NONE) continue;
                const int bin = rot_bin(ang1[feat_of(c)], ang2[(int)(0xFFFFFu - (key & 0xFFFFFu))]);
                if (bin >= 0 && bin < ORBGPU_HISTO_LENGTH) atomicAdd(&hist[bin], 1);
            }
            bar_post();
            if (gt == 0) three_maxima(hist, ORBGPU_HISTO_LENGTH, ind[0], ind[1], ind[2]);
            bar_post();
        }
        for (int e = gt; e < n_own; e += GT) {
            const int c = slot_of(e);
            const uint32_t key = best[c];
            const int f1 = key != KEY_NONE || !compact ? feat_of(c) : 0;
            if (compact) sEnt[c] = ENT_NONE;
            if (key == KEY_NONE) continue;
            const int idx2 = (int)(0xFFFFFu - (key & 0xFFFFFu));
            if (P.check_ori) {
                const int bin = rot_bin(ang1[f1], ang2[idx2]);
                if (bin >= 0 && bin < ORBGPU_HISTO_LENGTH && bin != ind[0] && bin != ind[1] && bin != ind[2]) continue;
            }
            if (compact) mark(c, f1, idx2); else put_match(f1, idx2);
            mine++;
        }
        if (gt == 0) stamp(i, 8);
        for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(FULL_MASK, mine, o);
        if (lane == 0 && mine) atomicAdd(&s_cnt[grp], mine);
        bar_post(); // the group's count is complete; hist / ind may be reused by the next pair
        if (gt == 0) {
            const int cnt = s_cnt[grp];
            for (int r = 0; r < P.n_targets; r++) P.tgt_nm[r][P.pair0 + p] = cnt;
        }
        if (compact) {
            // ascending idx1 (:1319-1324): the rank of a match is the number of matched features before it -- prefix counts of the
            // bitmap words by one warp, then every entry finds its place with one popcount
            if (gt < 32) {
                const int per = (n_words + 31) >> 5, w0 = min(n_words, gt * per), w1 = min(n_words, w0 + per);
                int sum = 0;
                for (int w = w0; w < w1; w++) sum += __popc(bm[w]);
                int incl = sum;
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(FULL_MASK, incl, o);
                    if (gt >= o) incl += u;
                }
                int run = incl - sum;
                for (int w = w0; w < w1; w++) { wbase[w] = (uint32_t)run; run += __popc(bm[w]); }
            }
            bar_post();
            const int cnt = s_cnt[grp];
            auto place = [&](int slot) {
                const uint32_t en = sEnt[slot];
                if (en == ENT_NONE) return;
                const int f1 = (int)(en >> 16);
                sOut[wbase[f1 >> 5] + __popc(bm[f1 >> 5] & ((1u << (f1 & 31)) - 1u))] = en;
            };
            for (int e = gt; e < n_own; e += GT) place(slot_of(e));
            const int n16 = (cnt + 3) >> 2;
            if (gt < 4 && cnt + gt < 4 * n16) sOut[cnt + gt] = ENT_NONE; // pads the last 16-byte chunk (4 * n16 <= max_free: a multiple of 4)
            bar_post();
            if (gt == 0) stamp(i, 9);
            // ship: only the valid prefix crosses the links, 16 bytes per store, every target (this rank's own buffer included)
            for (int r = 0; r < P.n_targets; r++) {
                uint4 *dst = (uint4 *)(P.tgt_m[r] + row_off);
                for (int x = gt; x < n16; x += GT) dst[x] = ((const uint4 *)sOut)[x];
            }
        }
        if (P.rows_preset == 2) { // the row is complete in this rank's buffer (all stores precede the barrier above): ship it whole
            const int32_t *src = P.tgt_m[0] + row_off;
            for (int r = 1; r < P.n_targets; r++) {
                int32_t *dst = P.tgt_m[r] + row_off;
                if ((n & 3) == 0) {
                    for (int x = gt; x < (n >> 2); x += GT) ((int4 *)dst)[x] = __ldcg((const int4 *)src + x);
                } else {
                    for (int x = gt; x < n; x += GT) dst[x] = __ldcg(src + x);
                }
            }
        }
        __syncwarp();
        if (gt == 0) stamp(i, 7);
        if (lane == 0) ts_mbar_arrive(bar_of(st, B_EMPTY));
    }
    if (compact && P.epoch_done) {
        // completion protocol of the fused all-gather, once per CTA: both post groups have issued the stores of all their pairs
        // (named barrier), then ONE thread orders them at system scope and adds the CTA's pairs to the rank's counter.  The CTA that
        // completes the batch publishes the rank's epoch in every target's flag array, waits until every source in wait_mask has
        // published it too (or a later one: a fast rank may already have finished the next step into the other buffer), and
        // advances the epoch -- the kernel ends when the gathered result is complete on this rank.  No memset, no second kernel.
        asm volatile("bar.sync 5, %0;" ::"r"(NG * 32) : "memory");
        if (warp == NC && lane == 0) {
            if (s_ncmp) atomicAdd(&P.counters[0], s_ncmp);
            __threadfence_system();
            const unsigned old = atomicAdd(&P.epoch_done[1], (unsigned)n_my);
            if (old + (unsigned)n_my == (unsigned)P.n_pairs) {
                __threadfence_system();
                const unsigned epoch = *(volatile uint32_t *)&P.epoch_done[0];
                for (int r = 0; r < P.n_targets; r++) *(volatile uint32_t *)(P.tgt_flag[r] + P.src_rank) = epoch;
                P.counters[2] = atomicExch(&P.counters[0], 0ull); // comparisons of this step; the running counter restarts at zero
                bool ok = true;
                unsigned long long t0;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                for (int src = 0; src < P.n_targets && ok; src++) {
                    if (!((P.wait_mask >> src) & 1u)) continue;
                    while ((int32_t)(*(volatile const uint32_t *)(P.tgt_flag[0] + src) - epoch) < 0) {
                        unsigned long long t1;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                        if (t1 - t0 > 4000000000ull) { ok = false; break; } // 4 s: a peer that never arrives must not hang the GPU
                        __nanosleep(100);
                    }
                }
                __threadfence_system(); // acquire side: the peers' pair data precedes their flags
                if (!ok && P.status) P.status[0] = 1;
                *(volatile uint32_t *)&P.epoch_done[1] = 0;
                *(volatile uint32_t *)&P.epoch_done[0] = epoch + 1;
            }
        }
    }
}

// ---- compact form of the result: the reference's vMatchedPairs (:1317-1325), pairs (idx1, idx2) in ascending idx1
__global__ void tri_offsets_kernel(int n_pairs, const int32_t *__restrict__ nmatches, int32_t *__restrict__ offsets)
{
    // exclusive scan by one block of 1024 threads over contiguous chunks
    __shared__ int warp_sum[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int per = (n_pairs + 1023) / 1024, lo = min(n_pairs, t * per), hi = min(n_pairs, lo + per);
    int mine = 0;
    for (int i = lo; i < hi; i++) mine += nmatches[i];
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sum[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, w, o);
            if (lane >= o) w += u;
        }
        warp_sum[lane] = w;
    }
    __syncthreads();
    int run = (warp ? warp_sum[warp - 1] : 0) + incl - mine;
    for (int i = lo; i < hi; i++) {
        offsets[i] = run;
        run += nmatches[i];
    }
    if (t == 1023) offsets[n_pairs] = warp_sum[31];
}

__global__ void tri_compact_kernel(int n_pairs, int n_feat, const int32_t *__restrict__ matches12, const int32_t *__restrict__ offsets,
                                   int2 *__restrict__ pairs, long long cap)
{
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; // one warp per key-frame pair
    if (p >= n_pairs) return;
    const int lane = lane_id();
    const int32_t *row = matches12 + (size_t)p * n_feat;
    long long out = offsets[p];
    for (int b = 0; b < n_feat; b += 32) {
        const int i = b + lane;
        const int m = i < n_feat ? row[i] : -1;
        const unsigned bal = __ballot_sync(FULL_MASK, m >= 0);
        const long long pos = out + __popc(bal & lanemask_lt());
        if (m >= 0 && pos < cap) pairs[pos] = make_int2(i, m);
        out += __popc(bal);
    }
}

// compact form (counts + fixed-stride (idx1 << 16 | idx2) entries) -> contiguous (idx1, idx2) pairs at the scanned offsets
__global__ void tri_pack_entries_kernel(int n_pairs, int n_feat, const uint32_t *__restrict__ entries, const int32_t *__restrict__ offsets,
                                        int2 *__restrict__ pairs, long long cap)
{
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; // one warp per key-frame pair
    if (p >= n_pairs) return;
    const int lane = lane_id();
    const long long o = offsets[p];
    const int cnt = offsets[p + 1] - offsets[p];
    const uint32_t *src = entries + (size_t)p * n_feat;
    for (int j = lane; j < cnt; j += 32) {
        const uint32_t e = src[j];
        if (o + j < cap) pairs[o + j] = make_int2((int)(e >> 16), (int)(e & 0xFFFFu));
    }
}

KfSetView kfset_view(const orbgpu_kfset *s)
{
    KfSetView v;
    v.n_kf = s->n_kf; v.n_feat = s->n_feat;
    v.angle = s->angle; v.u_right = s->u_right;
    v.scale_factors = s->scale_factors; v.level_sigma2 = s->level_sigma2;
    v.kf_n_nodes = s->kf_n_nodes; v.kf_node_ids = s->kf_node_ids; v.kf_node_off = s->kf_node_off;
    v.desc_csr = s->desc_csr; v.kp_csr = s->kp_csr; v.kf_n_free = s->kf_n_free;
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
    cudaFree(s->desc_csr); cudaFree(s->kp_csr); cudaFree(s->kf_n_free);
    cudaFree(s->blob); cudaFree(s->blob_bytes); cudaFree(s->aux);
    delete s;
}

// (re)builds the per-key-frame CSR, the node-ordered copies, the stream blobs and the aux records from s->node_id / s->has_mp
static int kfset_build_csr(orbgpu_ctx *ctx, orbgpu_kfset *s)
{
    int32_t *d_max = s->kf_n_free + s->n_kf;
    CU_TRY(cudaMemsetAsync(d_max, 0, 12, ctx->stream));
    int cap = 1;
    while (cap < s->n_feat) cap <<= 1;
    const size_t smem = (size_t)cap * 8;
    kfset_csr_kernel<<<s->n_kf, 1024, smem, ctx->stream>>>(s->n_feat, cap, s->node_id, s->has_mp, s->desc, s->xy, s->octave, s->kf_n_nodes,
                                                          s->kf_node_ids, s->kf_node_off, s->kf_feat, s->desc_csr, s->kp_csr, s->kf_n_free,
                                                          d_max, s->blob, s->blob_stride, s->blob_bytes, s->aux);
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    int32_t mx[3] = {0, 0, 0};
    CU_TRY(cudaMemcpyAsync(mx, d_max, 12, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    s->max_free = ((mx[0] > 0 ? mx[0] : 1) + 3) & ~3; // multiple of 4: keeps the int4 node table 16-byte aligned
    s->max_nodes = mx[1] > 0 ? mx[1] : 1;
    s->max_blob = mx[2] > 0 ? mx[2] : 16;
    return ORBGPU_OK;
}

extern "C" int orbgpu_kfset_upload(orbgpu_ctx *ctx, const orbgpu_kfset_host *h, orbgpu_kfset **out)
{
    ARG_TRY(ctx && h && out);
    ARG_TRY(h->n_kf > 0 && h->n_feat > 0 && h->n_feat <= 8192 && h->n_feat < (1 << 20));
    ARG_TRY(h->desc && h->kp_xy && h->octave && h->angle && h->has_mp && h->scale_factors && h->level_sigma2);
    ARG_TRY(h->n_levels > 0 && h->n_levels <= 64);
    for (size_t i = 0, T0 = (size_t)h->n_kf * h->n_feat; i < T0; i++)
        ARG_TRY(h->octave[i] >= 0 && h->octave[i] < h->n_levels); // indexes the scale tables in the gates
    CU_TRY(cudaSetDevice(ctx->device));
    orbgpu_kfset *s = new orbgpu_kfset();
    OwnedHandle<orbgpu_kfset, orbgpu_kfset_destroy> owner(s);
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
    if (h->node_id) UP(s->node_id, h->node_id, T * 4);
    else { // no FeatureVector yet: orbgpu_kfset_transform fills it on the device
        CU_TRY(cudaMalloc((void **)&s->node_id, T * 4));
        CU_TRY(cudaMemsetAsync(s->node_id, 0xFF, T * 4, ctx->stream));
    }
    UP(s->scale_factors, h->scale_factors, (size_t)h->n_levels * 4);
    UP(s->level_sigma2, h->level_sigma2, (size_t)h->n_levels * 4);
#undef UP
    CU_TRY(cudaMalloc((void **)&s->kf_n_nodes, (size_t)h->n_kf * 4));
    CU_TRY(cudaMalloc((void **)&s->kf_node_ids, T * 4));
    CU_TRY(cudaMalloc((void **)&s->kf_node_off, (size_t)h->n_kf * (h->n_feat + 1) * 4));
    CU_TRY(cudaMalloc((void **)&s->kf_feat, T * 4));
    CU_TRY(cudaMalloc((void **)&s->desc_csr, T * 32));
    CU_TRY(cudaMalloc((void **)&s->kp_csr, T * 16));
    CU_TRY(cudaMalloc((void **)&s->kf_n_free, (size_t)h->n_kf * 4 + 256));
    s->blob_stride = ((size_t)h->n_feat * 24 + 4 + 127) & ~size_t(127);
    CU_TRY(cudaMalloc((void **)&s->blob, s->blob_stride * h->n_kf));
    CU_TRY(cudaMalloc((void **)&s->aux, T * 32));
    CU_TRY(cudaMalloc((void **)&s->blob_bytes, (size_t)h->n_kf * 4));
    int rc = kfset_build_csr(ctx, s);
    if (rc) return rc;
    *out = owner.release();
    return ORBGPU_OK;
}

// development aid (not in the public header): device buffer receiving CTA 0's pipeline time stamps
extern "C" int orbgpu_debug_triangulation_timeline(orbgpu_ctx *ctx, void *dev_buf)
{
    ARG_TRY(ctx);
    ctx->tri_timeline = dev_buf;
    return ORBGPU_OK;
}

// TemplatedVocabulary::transform (TemplatedVocabulary.h:1127-1194, 1216-1258) for every feature of every key frame of the set,
// on the device: node ids of the FeatureVectors (level L - levelsup; stopped words dropped), then the CSR / stream blobs are
// rebuilt.  This is KeyFrame::ComputeBoW (KeyFrame.cc:102-117) for the whole batch, without a host round trip.
extern "C" int orbgpu_kfset_transform(orbgpu_ctx *ctx, const orbgpu_voc *voc, orbgpu_kfset *s, int32_t levelsup)
{
    ARG_TRY(ctx && voc && s);
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    rc = launch_voc_transform_nodes(ctx, voc, (long long)s->n_kf * s->n_feat, s->desc, levelsup, s->node_id);
    if (rc) return rc;
    rc = kfset_build_csr(ctx, s);
    if (rc) return rc;
    return ctx_fetch_comparisons(ctx);
}

extern "C" int orbgpu_triangulation_set_engine(orbgpu_ctx *ctx, int32_t engine)
{
    ARG_TRY(ctx && engine >= 0 && engine <= 2);
    ctx->tri_engine = engine;
    return ORBGPU_OK;
}

// n_targets > 1 (or pair_offset != 0, rows_preset): output into the given targets (engine 2 only); otherwise the plain call
static int tri_launch(orbgpu_ctx *ctx, const orbgpu_kfset *s, int32_t n_pairs, const int32_t *kf1_dev, const int32_t *kf2_dev,
                      const float *ep_dev, const float *f12_dev, int32_t only_stereo, int32_t coarse, int32_t check_ori,
                      int32_t *matches12_dev, int32_t *nmatches_dev, int n_targets, void *const *tgt_m, void *const *tgt_nm,
                      int64_t pair_offset, int rows_preset, const orbgpu_tri_gather *gather = nullptr)
{
    ARG_TRY(ctx && s && n_pairs >= 0);
    ARG_TRY(n_pairs == 0 || (kf1_dev && kf2_dev && ep_dev && f12_dev && matches12_dev && nmatches_dev));
    CU_TRY(cudaSetDevice(ctx->device));
    // the gather form keeps the comparison counter itself (the completing CTA moves it to slot 2 and clears it): no memset node in
    // the step unless another call has used the counters in between
    if (!gather || !ctx->gather_counters_clean) CU_TRY(cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    ctx->gather_counters_clean = gather != nullptr;
    ctx->cmp_slot = gather ? 2 : 0;
    if (n_pairs == 0) return ORBGPU_OK;
    // engine 2: persistent warp-specialised pipeline (monocular sets; bOnlyStereo on a monocular set matches nothing
    // and is left to the per-pair kernel)
    {
        constexpr int NC = 16, NG = 12, NJ = 3; // compare / post (two groups) / join warps (+ 1 producer warp = 1024 threads)
        const int cap = (s->max_blob + 127) & ~127;
        const size_t stage_bytes = (2 * (size_t)cap + (size_t)s->max_free * 16 + (size_t)TS_OVF * 4 + (size_t)s->max_nodes * 4 + 127) & ~size_t(127);
        const size_t bm_bytes = rows_preset == 3 ? (size_t)4 * ((s->n_feat + 31) / 32) * 4 : 0; // 2 post groups x {bitmap, prefix counts}
        int n_stages = (int)((227 * 1024 - 2048 - bm_bytes) / stage_bytes);
        if (n_stages > 4) n_stages = 4;
        const bool can = !s->u_right && !only_stereo && n_stages >= 3 && s->max_free <= 8192; // each post group holds a stage
        if ((ctx->tri_engine == 2 || tgt_m) && !can)
            return orbgpu_fail(ORBGPU_ERR_INVALID, "triangulation engine 2 needs a monocular keyframe set that fits the shared-memory ring");
        if (can && (ctx->tri_engine != 1 || tgt_m)) {
            TsParams P;
            P.blob = s->blob; P.blob_stride = s->blob_stride; P.blob_bytes = s->blob_bytes;
            P.kf_n_free = s->kf_n_free; P.kf_n_nodes = s->kf_n_nodes; P.aux = s->aux; P.angle = s->angle;
            P.scale_factors = s->scale_factors; P.level_sigma2 = s->level_sigma2;
            P.n_feat = s->n_feat; P.n_levels = s->n_levels; P.cap_bytes = cap; P.max_free = s->max_free; P.n_pairs = n_pairs;
            P.n_stages = n_stages; P.stage_bytes = (int)stage_bytes;
            P.kf1 = kf1_dev; P.kf2 = kf2_dev; P.ep = ep_dev; P.f12 = f12_dev;
            P.coarse = coarse; P.check_ori = check_ori;
            P.counters = ctx->d_counters;
            if (tgt_m) {
                P.n_targets = n_targets; P.rows_preset = rows_preset; P.pair0 = pair_offset;
                for (int r = 0; r < 8; r++) {
                    P.tgt_m[r] = r < n_targets ? (int32_t *)tgt_m[r] : nullptr;
                    P.tgt_nm[r] = r < n_targets ? (int32_t *)tgt_nm[r] : nullptr;
                }
            } else {
                P.n_targets = 1; P.rows_preset = 0; P.pair0 = 0;
                for (int r = 0; r < 8; r++) { P.tgt_m[r] = nullptr; P.tgt_nm[r] = nullptr; }
                P.tgt_m[0] = matches12_dev; P.tgt_nm[0] = nmatches_dev;
            }
            P.timeline = (long long *)ctx->tri_timeline;
            for (int r = 0; r < 8; r++) P.tgt_flag[r] = nullptr;
            P.epoch_done = nullptr;
            P.status = nullptr;
            P.wait_mask = 0;
            P.src_rank = 0;
            if (gather) {
                for (int r = 0; r < n_targets; r++) P.tgt_flag[r] = (uint32_t *)gather->flags[r];
                P.epoch_done = (uint32_t *)gather->epoch_done;
                P.status = (uint32_t *)gather->status;
                P.src_rank = gather->rank;
                // wait_mask is given by rank; the kernel's targets are in ring order starting at this rank
                for (int i = 0; i < n_targets; i++)
                    if ((gather->wait_mask >> ((gather->rank + i) % n_targets)) & 1u) P.wait_mask |= 1u << i;
            }
            auto kern2 = triangulation_stream_kernel<NC, NG, NJ, 2>;
            const size_t smem2 = stage_bytes * n_stages + bm_bytes;
            if (getenv("ORBGPU_DEBUG"))
                fprintf(stderr, "[orbgpu] triangulation_stream_kernel: %d stages x %zu B (+%zu B), max_free %d, max_blob %d, mode %d\n", n_stages,
                        stage_bytes, bm_bytes, s->max_free, s->max_blob, rows_preset);
            const int grid = n_pairs < ctx->sm_count ? n_pairs : ctx->sm_count;
            kern2<<<grid, (NC + NG + NJ + 1) * 32, smem2, ctx->stream>>>(P);
            LAUNCH_COUNT(ctx);
            CU_TRY(cudaGetLastError());
            return ORBGPU_OK;
        }
    }
    // 2 x max_free descriptors (32 B) + key (4 B) + histogram bin (1 B) per CSR slot
    const size_t smem = (size_t)s->max_free * (64 + 4 + 1) + (size_t)s->max_nodes * 16 + 64;
    if (smem > 227 * 1024) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "keyframes too large for the shared-memory staging");
    auto kern = s->u_right ? triangulation_pairs_kernel<true> : triangulation_pairs_kernel<false>;
    kern<<<n_pairs, TRI_THREADS, smem, ctx->stream>>>(kfset_view(s), n_pairs, s->max_free, s->max_nodes, kf1_dev, kf2_dev, ep_dev, f12_dev,
                                                     only_stereo, coarse, check_ori, matches12_dev, nmatches_dev, ctx->d_counters);
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    return ORBGPU_OK;
}

extern "C" int orbgpu_search_for_triangulation_batch_dev(orbgpu_ctx *ctx, const orbgpu_kfset *s, int32_t n_pairs, const int32_t *kf1_dev,
                                                         const int32_t *kf2_dev, const float *ep_dev, const float *f12_dev,
                                                         int32_t only_stereo, int32_t coarse, int32_t check_ori, int32_t *matches12_dev,
                                                         int32_t *nmatches_dev)
{
    return tri_launch(ctx, s, n_pairs, kf1_dev, kf2_dev, ep_dev, f12_dev, only_stereo, coarse, check_ori, matches12_dev, nmatches_dev, 0,
                      nullptr, nullptr, 0, 0);
}

// Fused search + all-gather: the match rows and counts of this rank's pairs are stored straight into the result buffers of ALL
// ranks (peer memory mapped over NVLink / NVSwitch, e.g. torch symmetric memory), at rows pair_offset + p.  With rows_preset the
// owners have filled their buffers with -1 beforehand and only the matches (a few hundred 4-byte stores per pair) and the
// counts cross the links; the caller separates steps with a cross-rank barrier.
extern "C" int orbgpu_search_for_triangulation_batch_peers_dev(orbgpu_ctx *ctx, const orbgpu_kfset *s, int32_t n_pairs,
                                                               const int32_t *kf1_dev, const int32_t *kf2_dev, const float *ep_dev,
                                                               const float *f12_dev, int32_t coarse, int32_t check_ori, int32_t n_targets,
                                                               void *const *target_matches, void *const *target_nmatches,
                                                               int64_t pair_offset, int32_t rows_preset)
{
    ARG_TRY(ctx && s && n_pairs >= 0 && n_targets >= 1 && n_targets <= 8 && target_matches && target_nmatches && pair_offset >= 0);
    for (int r = 0; r < n_targets; r++) ARG_TRY(target_matches[r] && target_nmatches[r]);
    return tri_launch(ctx, s, n_pairs, kf1_dev, kf2_dev, ep_dev, f12_dev, 0, coarse, check_ori, (int32_t *)target_matches[0],
                      (int32_t *)target_nmatches[0], n_targets, target_matches, target_nmatches, pair_offset,
                      rows_preset == 2 ? 2 : (rows_preset ? 1 : 0));
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

// same search, result in the reference's vMatchedPairs form: pairs[pair_offsets[p] .. pair_offsets[p+1]) = (idx1, idx2) of pair p,
// ascending idx1.  The dense rows stay on the device; only the pairs travel back (about 1/20 of the bytes).
extern "C" int orbgpu_search_for_triangulation_batch_pairs(orbgpu_ctx *ctx, const orbgpu_kfset *s, int32_t n_pairs, const int32_t *kf1,
                                                           const int32_t *kf2, const float *ep, const float *f12, int32_t only_stereo,
                                                           int32_t coarse, int32_t check_ori, int32_t *pair_offsets, int32_t *pairs,
                                                           int64_t cap, int64_t *total)
{
    ARG_TRY(ctx && s && n_pairs >= 0 && pair_offsets && total && cap >= 0 && (cap == 0 || pairs));
    ARG_TRY(n_pairs == 0 || (kf1 && kf2 && ep && f12));
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    *total = 0;
    pair_offsets[0] = 0;
    if (n_pairs == 0) return ORBGPU_OK;
    for (int p = 0; p < n_pairs; p++) ARG_TRY(kf1[p] >= 0 && kf1[p] < s->n_kf && kf2[p] >= 0 && kf2[p] < s->n_kf);
    const size_t P = (size_t)n_pairs;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o_k1 = take(P * 4), o_k2 = take(P * 4), o_ep = take(P * 8), o_f = take(P * 36);
    const size_t up = off;
    rc = stage_reserve(ctx, up);
    if (rc) return rc;
    rc = arena_reserve(ctx, up + align256(P * s->n_feat * 4) + 2 * align256((P + 1) * 4) + align256((size_t)cap * 8 + 8));
    if (rc) return rc;
    char *H = ctx->h_stage;
    memcpy(H + o_k1, kf1, P * 4);
    memcpy(H + o_k2, kf2, P * 4);
    memcpy(H + o_ep, ep, P * 8);
    memcpy(H + o_f, f12, P * 36);
    char *D = (char *)arena_take(ctx, up);
    int32_t *d_m = (int32_t *)arena_take(ctx, P * s->n_feat * 4), *d_nm = (int32_t *)arena_take(ctx, (P + 1) * 4),
            *d_off = (int32_t *)arena_take(ctx, (P + 1) * 4);
    int2 *d_pairs = (int2 *)arena_take(ctx, (size_t)cap * 8 + 8);
    if (!D || !d_m || !d_nm || !d_off || !d_pairs) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    CU_TRY(cudaMemcpyAsync(D, H, up, cudaMemcpyHostToDevice, ctx->stream));
    rc = orbgpu_search_for_triangulation_batch_dev(ctx, s, n_pairs, (const int32_t *)(D + o_k1), (const int32_t *)(D + o_k2),
                                                   (const float *)(D + o_ep), (const float *)(D + o_f), only_stereo, coarse, check_ori,
                                                   d_m, d_nm);
    if (rc) return rc;
    tri_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(n_pairs, d_nm, d_off);
    tri_compact_kernel<<<(unsigned)((P * 32 + 255) / 256), 256, 0, ctx->stream>>>(n_pairs, s->n_feat, d_m, d_off, d_pairs, (long long)cap);
    ctx->launches += 2;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(pair_offsets, d_off, (P + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    rc = ctx_fetch_comparisons(ctx); // synchronises: the total is known
    if (rc) return rc;
    *total = pair_offsets[n_pairs];
    const int64_t n_copy = *total < cap ? *total : cap;
    if (n_copy > 0) {
        CU_TRY(cudaMemcpyAsync(pairs, d_pairs, (size_t)n_copy * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return *total <= cap ? ORBGPU_OK : orbgpu_fail(ORBGPU_ERR_OVERFLOW, "pairs capacity too small: see *total");
}

// Fused search + all-gather in the reference's vMatchedPairs form (see orbgpu_tri_gather in the header): compact (idx1, idx2)
// entries of this rank's pairs stored straight from shared memory into every rank's buffers, epoch flags instead of a barrier.
extern "C" int orbgpu_search_for_triangulation_batch_gather_dev(orbgpu_ctx *ctx, const orbgpu_kfset *s, int32_t n_pairs,
                                                                const int32_t *kf1_dev, const int32_t *kf2_dev, const float *ep_dev,
                                                                const float *f12_dev, int32_t coarse, int32_t check_ori,
                                                                const orbgpu_tri_gather *g, int64_t pair_offset)
{
    ARG_TRY(ctx && s && g && n_pairs > 0 && pair_offset >= 0); // every rank ships at least one pair (its peers wait for its epoch)
    ARG_TRY(g->n_ranks >= 1 && g->n_ranks <= 8 && g->rank >= 0 && g->rank < g->n_ranks && g->epoch_done);
    ARG_TRY((s->n_feat & 3) == 0 && s->n_feat <= 65535); // 16-byte stores of (idx1 << 16 | idx2) entries
    for (int r = 0; r < g->n_ranks; r++) ARG_TRY(g->pairs[r] && g->counts[r] && g->flags[r]);
    // target 0 = this rank's own buffers, then the peers in ring order (spreads the first stores of a step over the links)
    void *tm[8], *tn[8];
    orbgpu_tri_gather go = *g;
    for (int i = 0; i < g->n_ranks; i++) {
        const int r = (g->rank + i) % g->n_ranks;
        tm[i] = g->pairs[r]; tn[i] = g->counts[r]; go.flags[i] = g->flags[r];
    }
    return tri_launch(ctx, s, n_pairs, kf1_dev, kf2_dev, ep_dev, f12_dev, 0, coarse, check_ori, (int32_t *)tm[0], (int32_t *)tn[0], g->n_ranks,
                      tm, tn, pair_offset, 3, &go);
}

// The gathered result of orbgpu_search_for_triangulation_batch_gather_dev on the host, in the form of
// orbgpu_search_for_triangulation_batch_pairs: offsets scan + pack on the device, one download of the valid pairs only.
extern "C" int orbgpu_tri_gather_download(orbgpu_ctx *ctx, int32_t n_pairs, int32_t n_feat, const void *counts_dev, const void *entries_dev,
                                          int32_t *pair_offsets, int32_t *pairs, int64_t cap, int64_t *total)
{
    ARG_TRY(ctx && n_pairs >= 0 && n_feat > 0 && pair_offsets && total && cap >= 0 && (cap == 0 || pairs));
    ARG_TRY(n_pairs == 0 || (counts_dev && entries_dev));
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    *total = 0;
    pair_offsets[0] = 0;
    if (n_pairs == 0) return ORBGPU_OK;
    const size_t P = (size_t)n_pairs;
    rc = arena_reserve(ctx, align256((P + 1) * 4) + align256((size_t)cap * 8 + 8) + 512);
    if (rc) return rc;
    int32_t *d_off = (int32_t *)arena_take(ctx, (P + 1) * 4);
    int2 *d_pairs = (int2 *)arena_take(ctx, (size_t)cap * 8 + 8);
    if (!d_off || !d_pairs) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    tri_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(n_pairs, (const int32_t *)counts_dev, d_off);
    tri_pack_entries_kernel<<<(unsigned)((P * 32 + 255) / 256), 256, 0, ctx->stream>>>(n_pairs, n_feat, (const uint32_t *)entries_dev, d_off, d_pairs,
                                                                                        (long long)cap);
    ctx->launches += 2;
    CU_TRY(cudaGetLastError());
    // offsets and (speculatively) the whole capacity of pairs in one go would move unused bytes: the total comes first
    CU_TRY(cudaMemcpyAsync(pair_offsets, d_off, (P + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    *total = pair_offsets[n_pairs];
    const int64_t n_copy = *total < cap ? *total : cap;
    if (n_copy > 0) {
        CU_TRY(cudaMemcpyAsync(pairs, d_pairs, (size_t)n_copy * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return *total <= cap ? ORBGPU_OK : orbgpu_fail(ORBGPU_ERR_OVERFLOW, "pairs capacity too small: see *total");
}

int triangulation_device_init()
{
    int rc;
    if ((rc = set_max_dyn_smem(kfset_csr_kernel)) || (rc = set_max_dyn_smem(triangulation_stream_kernel<16, 12, 3, 2>)) ||
        (rc = set_max_dyn_smem(triangulation_pairs_kernel<true>)) || (rc = set_max_dyn_smem(triangulation_pairs_kernel<false>)))
        return rc;
    return ORBGPU_OK;
}
