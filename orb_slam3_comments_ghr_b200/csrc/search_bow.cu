// search_bow.cu -- ORBmatcher::SearchByBoW, KeyFrame<->Frame (ORBmatcher.cc:262-496, Nleft == -1
// path) and KeyFrame<->KeyFrame (ORBmatcher.cc:890-1043), config C3.
//
// The reference merge-joins the two FeatureVector maps on NodeId and, inside a shared node, walks
// the keyframe features in order; a frame feature matched by an earlier keyframe feature is
// skipped by later ones (:335 / vbMatched2 :962,:990).  A feature belongs to exactly one node, so
// that dependency is node-local: one CTA per keyframe node (binary search for the partner node in
// the other FeatureVector), keyframe features replayed in order inside the CTA, threads over the
// partner's features with a block-wide lexicographic (distance, position) top-2.  The rotation
// histogram is accumulated with atomics (bin sizes are order independent) and culled afterwards.
#include <algorithm>
#include <cstring>
#include <vector>

#include "internal.cuh"

namespace {

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr int BOW_MAX_WARPS = 32;

// Big node pairs (e.g. the single root bucket when levelsup >= L, which is what Frame.cc:1008's levelsup = 4 gives with an
// L = 4 vocabulary: 1200 x 2000 features in one node) are not replayed feature by feature.  The "partner already matched"
// rule (:335 / :962) is solved as a fixed point over LOCK TIMES, as in the projection searches: lock[p] = position of the
// first keyframe feature that takes partner p.  Given the locks, every keyframe feature's outcome is independent (first two
// entries of its sorted candidate list that are not locked by an earlier feature); the outcomes give new locks; iterate
// until nothing changes.  The outcome of feature i depends only on features before i, so by induction the k-th iteration
// has the first k features right and the fixed point is the sequential result -- in practice a handful of iterations.
constexpr int BOW_BIG_N1 = 32, BOW_BIG_N2 = 256; // node pair taken by the fixed-point path: n1 >= .. && n2 >= ..
constexpr int BOW_LIST_K = 8;                    // sorted candidate list per keyframe feature
__device__ __forceinline__ bool bow_big(int n1, int n2) { return n1 >= BOW_BIG_N1 && n2 >= BOW_BIG_N2; }
__device__ __forceinline__ int bow_partner(const FrameView &f, uint32_t nid)
{ // lower_bound merge-join (:292-467) == look the node id up in the other sorted list
    int lo = 0, hi = f.fv_n_nodes;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (f.fv_node_ids[mid] < nid) lo = mid + 1; else hi = mid;
    }
    return (lo < f.fv_n_nodes && f.fv_node_ids[lo] == nid) ? lo : -1;
}

// MODE 0: KF <-> F   (match indexed by F feature, value = KF feature)
// MODE 1: KF1 <-> KF2 (match indexed by KF1 feature, value = KF2 feature)
template <int MODE>
__device__ __forceinline__ void bow_match_body(const int a, const FrameView &kf, const FrameView &f, const uint8_t *__restrict__ kf_valid,
                                               const uint8_t *__restrict__ f_valid, float nnratio, int check_ori, int32_t *match,
                                               uint8_t *matched2, int32_t *__restrict__ bin_of, int *__restrict__ hist,
                                               int *__restrict__ nmatches, unsigned long long *__restrict__ counters, int stage_cap,
                                               int skip_big)
{
    // optional staging (dynamic shared memory, stage_cap descriptors): when the partner node fits, its descriptors and
    // "already matched" flags are copied once and every keyframe feature of the node is replayed against shared memory --
    // a large node (e.g. the single root bucket of levelsup >= L) is otherwise one global round trip per keyframe feature
    extern __shared__ uint4 bow_smem[];
    __shared__ uint32_t wm1[BOW_MAX_WARPS], wm2[BOW_MAX_WARPS];
    __shared__ int s_b;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
    if (t == 0) s_b = bow_partner(f, kf.fv_node_ids[a]);
    __syncthreads();
    const int b = s_b;
    if (b < 0) return;
    const int s1 = kf.fv_offsets[a], e1 = kf.fv_offsets[a + 1];
    const int s2 = f.fv_offsets[b], n2 = f.fv_offsets[b + 1] - s2;
    if (skip_big && bow_big(e1 - s1, n2)) return; // bow_big_lists_kernel + bow_big_resolve_kernel take this node pair
    const bool staged = n2 <= stage_cap && (e1 - s1) >= 4; // worth it only when several keyframe features share the copy
    int *sIdx2 = (int *)(bow_smem + 2 * (size_t)stage_cap);   // [stage_cap] feature id of the partner's entry
    float *sAng2 = (float *)(sIdx2 + stage_cap);               // [stage_cap] its angle
    uint8_t *sTaken = (uint8_t *)(sAng2 + stage_cap);          // [stage_cap] already matched / not eligible
    if (staged) {
        for (int p = t; p < n2; p += blockDim.x) {
            const int idx2 = (int)f.fv_features[s2 + p];
            bow_smem[2 * p] = f.desc[2 * idx2];
            bow_smem[2 * p + 1] = f.desc[2 * idx2 + 1];
            sIdx2[p] = idx2;
            sAng2[p] = f.angle[idx2];
            sTaken[p] = (MODE == 0) ? (uint8_t)(match[idx2] >= 0) : (uint8_t)(matched2[idx2] || !f_valid[idx2]);
        }
        __syncthreads();
    }
    int my_matches = 0;
    unsigned long long ncmp = 0;
    // the keyframe feature of the NEXT iteration is fetched while the current one is scanned (three dependent global loads)
    int n_idx1 = (int)kf.fv_features[s1];
    bool n_valid = kf_valid[n_idx1] != 0;
    uint4 n_qa = kf.desc[2 * n_idx1], n_qb = kf.desc[2 * n_idx1 + 1];
    float n_ang = check_ori ? kf.angle[n_idx1] : 0.f;
    for (int iKF = s1; iKF < e1; iKF++) {
        const int idx1 = n_idx1;
        const bool valid1 = n_valid;
        const uint4 qa = n_qa, qb = n_qb;
        const float ang1 = n_ang;
        if (iKF + 1 < e1) {
            n_idx1 = (int)kf.fv_features[iKF + 1];
            n_valid = kf_valid[n_idx1] != 0;
            n_qa = kf.desc[2 * n_idx1]; n_qb = kf.desc[2 * n_idx1 + 1];
            if (check_ori) n_ang = kf.angle[n_idx1];
        }
        if (!valid1) continue; // :311-315 / :937-941 (block-uniform)
        uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
        if (staged) {
            for (int p = t; p < n2; p += blockDim.x) {
                if (sTaken[p]) continue; // :335 / :962-966
                const uint32_t d = (uint32_t)ham256(qa, qb, bow_smem[2 * p], bow_smem[2 * p + 1]);
                top2_push(b1, b2, (d << 20) | (uint32_t)p);
                ncmp++;
            }
        } else {
            for (int p = t; p < n2; p += blockDim.x) {
                const int idx2 = (int)f.fv_features[s2 + p];
                if (MODE == 0) {
                    if (match[idx2] >= 0) continue; // :335
                } else {
                    if (matched2[idx2] || !f_valid[idx2]) continue; // :962-966
                }
                const uint32_t d = (uint32_t)ham256(qa, qb, f.desc[2 * idx2], f.desc[2 * idx2 + 1]);
                top2_push(b1, b2, (d << 20) | (uint32_t)p);
                ncmp++;
            }
        }
        uint32_t m1, m2;
        warp_top2(b1, b2, m1, m2);
        if (nwarps > 1) {
            if (lane == 0) { wm1[warp] = m1; wm2[warp] = m2; }
            __syncthreads();
            if (warp == 0) { // the per-warp pairs are merged by one warp-wide top-2 (keys are unique or KEY_NONE)
                const uint32_t c1 = lane < nwarps ? wm1[lane] : KEY_NONE, c2 = lane < nwarps ? wm2[lane] : KEY_NONE;
                warp_top2(c1, c2, m1, m2);
            }
        }
        if (t == 0 && m1 != KEY_NONE) {
            const int bestDist1 = (int)(m1 >> 20);
            const int bestDist2 = (m2 == KEY_NONE) ? 256 : (int)(m2 >> 20);
            const bool th_ok = (MODE == 0) ? (bestDist1 <= ORBGPU_TH_LOW) : (bestDist1 < ORBGPU_TH_LOW); // :392 / :985
            if (th_ok && (float)bestDist1 < __fmul_rn(nnratio, (float)bestDist2)) {                     // :395 / :987
                const int best2 = staged ? sIdx2[m1 & 0xFFFFF] : (int)f.fv_features[s2 + (m1 & 0xFFFFF)];
                const int slot = (MODE == 0) ? best2 : idx1;
                if (MODE == 0) {
                    match[best2] = idx1; // vpMapPointMatches[bestIdxF] = pMP (of KF feature idx1)
                } else {
                    match[idx1] = best2; // vpMatches12[idx1] = vpMapPoints2[bestIdx2]
                    matched2[best2] = 1;
                }
                if (staged) sTaken[m1 & 0xFFFFF] = 1;
                if (check_ori) { // :405-419 / :992-1002
                    const int bin = rot_bin(ang1, staged ? sAng2[m1 & 0xFFFFF] : f.angle[best2]);
                    if (bin >= 0 && bin < ORBGPU_HISTO_LENGTH) {
                        atomicAdd(&hist[bin], 1);
                        bin_of[slot] = bin;
                    }
                }
                my_matches++;
            }
        }
        __syncthreads(); // the new match must be visible before the next keyframe feature scans
    }
    for (int off = 16; off; off >>= 1) ncmp += __shfl_xor_sync(FULL_MASK, ncmp, off);
    if (lane == 0 && ncmp) atomicAdd(&counters[0], ncmp);
    if (t == 0 && my_matches) atomicAdd(nmatches, my_matches);
}

// ---- fixed-point path, step 1: one warp per keyframe feature of a big node pair lists its BOW_LIST_K best eligible partners,
// ascending in (distance, position).  lists[q][k]: keys (dist << 20 | position in the partner node), KEY_NONE padded;
// meta[q] = m | more << 8: the first m entries are exactly the m smallest keys, `more` = eligible partners exist beyond them.
template <int MODE>
__global__ void bow_big_lists_kernel(FrameView kf, FrameView f, const uint8_t *__restrict__ kf_valid, const uint8_t *__restrict__ f_valid,
                                     uint32_t *__restrict__ lists, uint32_t *__restrict__ meta)
{
    const int q = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5); // position in the keyframe's FeatureVector
    const int lane = threadIdx.x & 31;
    if (q >= kf.fv_offsets[kf.fv_n_nodes]) return;
    int lo = 0, hi = kf.fv_n_nodes - 1; // node of position q: last a with fv_offsets[a] <= q
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (kf.fv_offsets[mid] <= q) lo = mid; else hi = mid - 1;
    }
    const int a = lo;
    const int b = bow_partner(f, kf.fv_node_ids[a]);
    if (b < 0) return;
    const int s2 = f.fv_offsets[b], n2 = f.fv_offsets[b + 1] - s2;
    if (!bow_big(kf.fv_offsets[a + 1] - kf.fv_offsets[a], n2)) return;
    const int idx1 = (int)kf.fv_features[q];
    if (!kf_valid[idx1]) return; // :311-315 / :937-941
    const uint4 qa = kf.desc[2 * idx1], qb = kf.desc[2 * idx1 + 1];
    uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
    int cnt = 0;
    for (int p = lane; p < n2; p += 32) {
        const int idx2 = (int)f.fv_features[s2 + p];
        if (MODE == 1 && !f_valid[idx2]) continue; // :962-966 (nothing is matched yet: :335 only bites through the locks)
        const uint32_t d = (uint32_t)ham256(qa, qb, f.desc[2 * idx2], f.desc[2 * idx2 + 1]);
        top2_push(b1, b2, (d << 20) | (uint32_t)p);
        cnt++;
    }
    // pop the lanes' pairs in ascending order; the prefix stays exact until a lane that dropped keys runs empty
    uint32_t out = KEY_NONE;
    int m = 0, popped = 0;
    for (int k = 0; k < BOW_LIST_K; k++) {
        const uint32_t head = popped == 0 ? b1 : (popped == 1 ? b2 : KEY_NONE);
        const uint32_t mn = __reduce_min_sync(FULL_MASK, head);
        if (mn == KEY_NONE) break;
        if (lane == k) out = mn;
        m++;
        if (head == mn) popped++;
        if (__any_sync(FULL_MASK, head == mn && popped == 2 && cnt > 2)) break;
    }
    int total = cnt;
    for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(FULL_MASK, total, o);
    if (lane < BOW_LIST_K) lists[(size_t)q * BOW_LIST_K + lane] = out;
    if (lane == 0) meta[q] = (uint32_t)m | ((total > m ? 1u : 0u) << 8);
}

// ---- step 2: one CTA per big node pair iterates outcomes <-> lock times to the fixed point, then writes the matches
template <int MODE>
__global__ void bow_big_resolve_kernel(FrameView kf, FrameView f, const uint8_t *__restrict__ kf_valid, const uint8_t *__restrict__ f_valid,
                                       const uint32_t *__restrict__ lists, const uint32_t *__restrict__ meta, float nnratio, int check_ori,
                                       int32_t *match, int32_t *__restrict__ bin_of, int *__restrict__ hist, int *__restrict__ nmatches,
                                       unsigned long long *__restrict__ counters, int n1_cap, int n2_cap)
{
    extern __shared__ int big_smem[];
    int *lock = big_smem;          // [n2_cap] position of the first keyframe feature that takes the partner
    int *choice = lock + n2_cap;   // [n1_cap] partner taken by the feature, -1 none
    int *queue = choice + n1_cap;  // [n1_cap] features whose list ran out: full rescan
    __shared__ uint32_t wm1[BOW_MAX_WARPS], wm2[BOW_MAX_WARPS];
    __shared__ int s_b, s_changed, s_nq, s_elig, s_carry, s_warp[BOW_MAX_WARPS];
    __shared__ unsigned long long s_cmp;
    const int a = blockIdx.x;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
    if (t == 0) {
        s_b = bow_partner(f, kf.fv_node_ids[a]);
        s_elig = 0; s_carry = 0; s_cmp = 0;
    }
    __syncthreads();
    const int b = s_b;
    if (b < 0) return;
    const int s1 = kf.fv_offsets[a], n1 = kf.fv_offsets[a + 1] - s1;
    const int s2 = f.fv_offsets[b], n2 = f.fv_offsets[b + 1] - s2;
    if (!bow_big(n1, n2)) return;
    for (int p = t; p < n2; p += blockDim.x) lock[p] = INT_MAX;
    for (int i = t; i < n1; i += blockDim.x) choice[i] = -2;
    if (MODE == 1) { // partners that can be compared at all (:962-966)
        int c = 0;
        for (int p = t; p < n2; p += blockDim.x) c += f_valid[f.fv_features[s2 + p]] != 0;
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
        if (lane == 0 && c) atomicAdd(&s_elig, c);
    }
    __syncthreads();
    const int n_elig = MODE == 1 ? s_elig : n2;

    // accept test of :392-395 / :985-987 on the (best, second) keys
    auto decide = [&](uint32_t k1, uint32_t k2) -> int {
        if (k1 == KEY_NONE) return -1;
        const int d1 = (int)(k1 >> 20), d2 = (k2 == KEY_NONE) ? 256 : (int)(k2 >> 20);
        const bool th_ok = (MODE == 0) ? (d1 <= ORBGPU_TH_LOW) : (d1 < ORBGPU_TH_LOW);
        return (th_ok && (float)d1 < __fmul_rn(nnratio, (float)d2)) ? (int)(k1 & 0xFFFFFu) : -1;
    };

    for (int iter = 0; iter <= n1 + 1; iter++) {
        if (t == 0) { s_changed = 0; s_nq = 0; }
        __syncthreads();
        for (int i = t; i < n1; i += blockDim.x) {
            const int idx1 = (int)kf.fv_features[s1 + i];
            if (!kf_valid[idx1]) continue;
            const uint32_t mt = meta[s1 + i];
            const int m = (int)(mt & 0xFF);
            uint32_t k1 = KEY_NONE, k2 = KEY_NONE;
            for (int e = 0; e < m; e++) {
                const uint32_t key = lists[(size_t)(s1 + i) * BOW_LIST_K + e];
                if (lock[key & 0xFFFFFu] < i) continue; // taken by an earlier keyframe feature (:335 / :962)
                if (k1 == KEY_NONE) k1 = key;
                else { k2 = key; break; }
            }
            if (k2 == KEY_NONE && (mt >> 8)) { // the list does not reach the second free partner
                queue[atomicAdd(&s_nq, 1)] = i;
                continue;
            }
            const int c = decide(k1, k2);
            if (c != choice[i]) { choice[i] = c; s_changed = 1; }
        }
        __syncthreads();
        const int nq = s_nq;
        for (int qi = 0; qi < nq; qi++) { // rare: block-wide scan of the whole partner node for one feature
            const int i = queue[qi];
            const int idx1 = (int)kf.fv_features[s1 + i];
            const uint4 qa = kf.desc[2 * idx1], qb = kf.desc[2 * idx1 + 1];
            uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
            for (int p = t; p < n2; p += blockDim.x) {
                if (lock[p] < i) continue;
                const int idx2 = (int)f.fv_features[s2 + p];
                if (MODE == 1 && !f_valid[idx2]) continue;
                const uint32_t d = (uint32_t)ham256(qa, qb, f.desc[2 * idx2], f.desc[2 * idx2 + 1]);
                top2_push(b1, b2, (d << 20) | (uint32_t)p);
            }
            uint32_t m1, m2;
            warp_top2(b1, b2, m1, m2);
            if (lane == 0) { wm1[warp] = m1; wm2[warp] = m2; }
            __syncthreads();
            if (warp == 0) {
                const uint32_t c1 = lane < nwarps ? wm1[lane] : KEY_NONE, c2 = lane < nwarps ? wm2[lane] : KEY_NONE;
                warp_top2(c1, c2, m1, m2);
                if (lane == 0) {
                    const int c = decide(m1, m2);
                    if (c != choice[i]) { choice[i] = c; s_changed = 1; }
                }
            }
            __syncthreads();
        }
        if (!s_changed) break;
        for (int p = t; p < n2; p += blockDim.x) lock[p] = INT_MAX;
        __syncthreads();
        for (int i = t; i < n1; i += blockDim.x)
            if (choice[i] >= 0) atomicMin(&lock[choice[i]], i);
        __syncthreads();
    }

    // ---- the fixed point: every accepted feature is the only holder of its partner.  Outputs as in the replay kernel;
    // DescriptorDistance calls of the sequential scan: feature i compares the eligible partners not taken before it
    int my_matches = 0;
    unsigned long long ncmp = 0;
    for (int base = 0; base < n1; base += blockDim.x) {
        const int i = base + t;
        int acc = 0, idx1 = -1;
        bool valid = false;
        if (i < n1) {
            idx1 = (int)kf.fv_features[s1 + i];
            valid = kf_valid[idx1] != 0;
            acc = (valid && choice[i] >= 0) ? 1 : 0;
        }
        int incl = acc;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = lane < nwarps ? s_warp[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(FULL_MASK, w, o);
                if (lane >= o) w += u;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int before = s_carry + (warp ? s_warp[warp - 1] : 0) + incl - acc; // accepted features ahead of i
        if (valid) ncmp += (unsigned long long)(n_elig - before);
        if (acc) {
            const int best2 = (int)f.fv_features[s2 + choice[i]];
            const int slot = (MODE == 0) ? best2 : idx1;
            if (MODE == 0) match[best2] = idx1; else match[idx1] = best2;
            if (check_ori) { // :405-419 / :992-1002
                const int bin = rot_bin(kf.angle[idx1], f.angle[best2]);
                if (bin >= 0 && bin < ORBGPU_HISTO_LENGTH) {
                    atomicAdd(&hist[bin], 1);
                    bin_of[slot] = bin;
                }
            }
            my_matches++;
        }
        __syncthreads();
        if (t == 0) s_carry += s_warp[nwarps - 1];
        __syncthreads();
    }
    for (int o = 16; o; o >>= 1) {
        ncmp += __shfl_xor_sync(FULL_MASK, ncmp, o);
        my_matches += __shfl_xor_sync(FULL_MASK, my_matches, o);
    }
    if (lane == 0 && ncmp) atomicAdd(&counters[0], ncmp);
    if (lane == 0 && my_matches) atomicAdd(nmatches, my_matches);
}

template <int MODE>
__global__ void bow_match_kernel(FrameView kf, FrameView f, const uint8_t *__restrict__ kf_valid,
                                 const uint8_t *__restrict__ f_valid, float nnratio, int check_ori, int32_t *match,
                                 uint8_t *matched2, int32_t *__restrict__ bin_of, int *__restrict__ hist, int *__restrict__ nmatches,
                                 unsigned long long *__restrict__ counters, int stage_cap, int skip_big)
{
    bow_match_body<MODE>(blockIdx.x, kf, f, kf_valid, f_valid, nnratio, check_ori, match, matched2, bin_of, hist, nmatches, counters, stage_cap,
                         skip_big);
}

// One frame (or key frame) against K key frames -- the relocalisation loop (MODE 0: the K key frames are the `kf` side, the frame is
// fixed) and the loop / merge candidate window (MODE 1: the current key frame is the fixed `kf` side, the K window key frames are the
// `f` side).  blockIdx.y = pair, blockIdx.x = node of the pair's `kf` side.  ONE launch for all K searches; the per-pair arguments are
// read from a device array.
struct BowBatchArg {
    FrameView other;            // the side that varies over the batch
    const uint8_t *other_valid; // its map-point mask
    int32_t *match, *bin_of;
    uint8_t *matched2;
    int *hist; // [64]: [0..29] histogram, [32] nmatches
    int n_nodes; // blocks of this pair (0: it went through its own launches, or has nothing to join)
};
template <int MODE>
__global__ void bow_match_batch_kernel(const BowBatchArg *__restrict__ args, FrameView fixed, const uint8_t *__restrict__ fixed_valid,
                                       float nnratio, int check_ori, unsigned long long *__restrict__ counters, int stage_cap)
{
    const BowBatchArg &A = args[blockIdx.y];
    if ((int)blockIdx.x >= A.n_nodes) return;
    if (MODE == 0)
        bow_match_body<0>(blockIdx.x, A.other, fixed, A.other_valid, nullptr, nnratio, check_ori, A.match, A.matched2, A.bin_of, A.hist,
                          A.hist + 32, counters, stage_cap, 0);
    else
        bow_match_body<1>(blockIdx.x, fixed, A.other, fixed_valid, A.other_valid, nnratio, check_ori, A.match, A.matched2, A.bin_of, A.hist,
                          A.hist + 32, counters, stage_cap, 0);
}

__device__ __forceinline__ void bow_cull_body(int n, int check_ori, int32_t *__restrict__ match, const int32_t *__restrict__ bin_of,
                                              const int *__restrict__ hist, int *__restrict__ nmatches);
__global__ void bow_cull_batch_kernel(const BowBatchArg *__restrict__ args, int n, int check_ori)
{
    const BowBatchArg &A = args[blockIdx.x];
    if (!A.match) return; // a pair that went through its own launches (big node pairs)
    bow_cull_body(n, check_ori, A.match, A.bin_of, A.hist, A.hist + 32);
}

__global__ void bow_cull_kernel(int n, int check_ori, int32_t *__restrict__ match, const int32_t *__restrict__ bin_of,
                                const int *__restrict__ hist, int *__restrict__ nmatches)
{
    bow_cull_body(n, check_ori, match, bin_of, hist, nmatches);
}

__device__ __forceinline__ void bow_cull_body(int n, int check_ori, int32_t *__restrict__ match, const int32_t *__restrict__ bin_of,
                                              const int *__restrict__ hist, int *__restrict__ nmatches)
{
    __shared__ int ind[3];
    __shared__ int removed;
    if (threadIdx.x == 0) {
        removed = 0;
        three_maxima(hist, ORBGPU_HISTO_LENGTH, ind[0], ind[1], ind[2]);
    }
    __syncthreads();
    if (!check_ori) return;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { // :470-493 / :1022-1040
        const int b = bin_of[i];
        if (b >= 0 && b != ind[0] && b != ind[1] && b != ind[2]) {
            match[i] = -1;
            atomicAdd(&removed, 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *nmatches -= removed;
}

// arena bytes one pair needs (scratch + masks); the match vector and the histogram block are the caller's
size_t bow_scratch_bytes(const orbgpu_frame *kf, const orbgpu_frame *f, int mode, bool *big_out)
{
    const int n_out = (mode == 0) ? f->n : kf->n;
    // fixed-point path for big node pairs (both sizes are known on the host as launch hints)
    const bool big = kf->fv_max_node >= BOW_BIG_N1 && f->fv_max_node >= BOW_BIG_N2 && kf->fv_n_nodes > 0 && f->fv_n_nodes > 0 &&
                     ((size_t)f->fv_max_node + 2 * (size_t)kf->fv_max_node) * 4 <= 200 * 1024; // else: replay kernel for every node
    if (big_out) *big_out = big;
    (void)n_out;
    return 256 + (big ? align256((size_t)kf->n * BOW_LIST_K * 4) + align256((size_t)kf->n * 4) : 0);
}
// per-pair result / state block, initialised by the caller with two memsets (0xFF part, zero part) -- for a batch, two memsets in all
struct BowBlock {
    int32_t *d_match, *d_bin; // [n_out] each, 0xFF
    int *d_hist;              // [64] zero: [0..29] histogram, [32] nmatches
    uint8_t *d_m2;            // [f->n + 1] zero
};
inline size_t bow_ff_bytes(int n_out) { return 2 * align256((size_t)n_out * 4); }
inline size_t bow_zero_bytes(int fn) { return 256 + align256((size_t)fn + 1); }
inline BowBlock bow_block(char *ff, char *zero, int n_out)
{
    BowBlock b;
    b.d_match = (int32_t *)ff; b.d_bin = (int32_t *)(ff + align256((size_t)n_out * 4));
    b.d_hist = (int *)zero; b.d_m2 = (uint8_t *)(zero + 256);
    return b;
}

// enqueues one SearchByBoW on the context's stream: masks already on the device, results left in d_match [n_out] and d_hist[32]
// (= nmatches).  The arena must hold bow_scratch_bytes() more bytes.
int bow_enqueue(orbgpu_ctx *ctx, int mode, const orbgpu_frame *kf, const orbgpu_frame *f, const uint8_t *d_kfv, const uint8_t *d_fv,
                float nnratio, int check_ori, const BowBlock &blk)
{
    const int n_out = (mode == 0) ? f->n : kf->n;
    bool big = false;
    bow_scratch_bytes(kf, f, mode, &big);
    int32_t *d_match = blk.d_match, *d_bin = blk.d_bin;
    uint8_t *d_m2 = blk.d_m2;
    int *d_hist = blk.d_hist, *d_nm = d_hist + 32; // [0..29] histogram, [32] nmatches
    if (kf->fv_n_nodes > 0 && f->fv_n_nodes > 0) {
        const int threads = f->fv_max_node <= 32 ? 32 : (f->fv_max_node <= 512 ? 128 : (f->fv_max_node <= 1024 ? 256 : 1024));
        // staging capacity: the largest partner node, as far as shared memory goes (33 B per descriptor)
        int stage_cap = f->fv_max_node > 64 ? f->fv_max_node : 0;
        if (stage_cap > 4800) stage_cap = 4800;
        stage_cap = (stage_cap + 15) & ~15;
        const size_t smem = (size_t)stage_cap * 41; // 32 B descriptor + feature id + angle + flag
        const FrameView vk = frame_view(kf), vf = frame_view(f);
        if (mode == 0) {
            bow_match_kernel<0><<<kf->fv_n_nodes, threads, smem, ctx->stream>>>(vk, vf, d_kfv, d_fv, nnratio, check_ori, d_match, d_m2, d_bin,
                                                                               d_hist, d_nm, ctx->d_counters, stage_cap, big ? 1 : 0);
        } else {
            bow_match_kernel<1><<<kf->fv_n_nodes, threads, smem, ctx->stream>>>(vk, vf, d_kfv, d_fv, nnratio, check_ori, d_match, d_m2, d_bin,
                                                                               d_hist, d_nm, ctx->d_counters, stage_cap, big ? 1 : 0);
        }
        LAUNCH_COUNT(ctx);
        if (big) {
            uint32_t *d_lists = (uint32_t *)arena_take(ctx, (size_t)kf->n * BOW_LIST_K * 4), *d_meta = (uint32_t *)arena_take(ctx, (size_t)kf->n * 4);
            if (!d_lists || !d_meta) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
            const int n1_cap = kf->fv_max_node, n2_cap = f->fv_max_node;
            const size_t smem_big = ((size_t)n2_cap + 2 * (size_t)n1_cap) * 4;
            const int blocks = (int)(((size_t)kf->n * 32 + 255) / 256);
            if (smem_big > ORBGPU_SMEM_OPTIN - 4096) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "node too large for the shared-memory lock table");
            if (mode == 0) {
                bow_big_lists_kernel<0><<<blocks, 256, 0, ctx->stream>>>(vk, vf, d_kfv, d_fv, d_lists, d_meta);
                bow_big_resolve_kernel<0><<<kf->fv_n_nodes, 1024, smem_big, ctx->stream>>>(vk, vf, d_kfv, d_fv, d_lists, d_meta, nnratio, check_ori, d_match,
                                                                                         d_bin, d_hist, d_nm, ctx->d_counters, n1_cap, n2_cap);
            } else {
                bow_big_lists_kernel<1><<<blocks, 256, 0, ctx->stream>>>(vk, vf, d_kfv, d_fv, d_lists, d_meta);
                bow_big_resolve_kernel<1><<<kf->fv_n_nodes, 1024, smem_big, ctx->stream>>>(vk, vf, d_kfv, d_fv, d_lists, d_meta, nnratio, check_ori, d_match,
                                                                                         d_bin, d_hist, d_nm, ctx->d_counters, n1_cap, n2_cap);
            }
            LAUNCH_COUNT(ctx);
            LAUNCH_COUNT(ctx);
        }
        bow_cull_kernel<<<1, 256, 0, ctx->stream>>>(n_out, check_ori, d_match, d_bin, d_hist, d_nm);
        LAUNCH_COUNT(ctx);
        CU_TRY(cudaGetLastError());
    }
    return ORBGPU_OK;
}

int run_bow(orbgpu_ctx *ctx, int mode, const orbgpu_frame *kf, const orbgpu_frame *f, const uint8_t *kf_valid, const uint8_t *f_valid,
            float nnratio, int check_ori, int32_t *match_out, int32_t *nmatches)
{
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    *nmatches = 0;
    const int n_out = (mode == 0) ? f->n : kf->n; // size of the match vector
    if (n_out == 0) return ORBGPU_OK;
    const size_t ffb = bow_ff_bytes(n_out), zb = bow_zero_bytes(f->n);
    rc = arena_reserve(ctx, bow_scratch_bytes(kf, f, mode, nullptr) + ffb + zb + align256(kf->n + 1) + align256(f->n + 1) + 1024);
    if (rc) return rc;
    char *ff = (char *)arena_take(ctx, ffb), *zero = (char *)arena_take(ctx, zb);
    uint8_t *d_kfv = (uint8_t *)arena_take(ctx, kf->n + 1), *d_fv = (uint8_t *)arena_take(ctx, f->n + 1);
    if (!ff || !zero || !d_kfv || !d_fv) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    CU_TRY(cudaMemsetAsync(ff, 0xFF, ffb, ctx->stream));
    CU_TRY(cudaMemsetAsync(zero, 0, zb, ctx->stream));
    if (kf->n) CU_TRY(cudaMemcpyAsync(d_kfv, kf_valid, kf->n, cudaMemcpyHostToDevice, ctx->stream));
    if (mode == 1 && f->n) CU_TRY(cudaMemcpyAsync(d_fv, f_valid, f->n, cudaMemcpyHostToDevice, ctx->stream));
    const BowBlock blk = bow_block(ff, zero, n_out);
    rc = bow_enqueue(ctx, mode, kf, f, d_kfv, d_fv, nnratio, check_ori, blk);
    if (rc) return rc;
    const OutPiece out[2] = {{match_out, blk.d_match, (size_t)n_out * 4}, {nmatches, blk.d_hist + 32, 4}};
    return ctx_download(ctx, out, 2);
}

// One frame against K candidate key frames (mode 0: Tracking::Relocalization, Tracking.cc:4469-4495) or the current key frame against
// the K key frames of a candidate's covisibility window (mode 1: LoopClosing.cc:909-925): ONE match launch over (node, pair) and ONE
// cull launch for all K searches, one upload of the arguments and validity masks, two memsets, one download and one synchronisation.
// Pairs that need the big-node fixed point keep their own launches.
int run_bow_batch(orbgpu_ctx *ctx, int mode, int K, const orbgpu_frame *const *others, const orbgpu_frame *fixed,
                  const uint8_t *const *other_valid, const uint8_t *fixed_valid, float nnratio, int check_ori, int32_t *match_out,
                  int32_t *nmatches)
{
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    for (int k = 0; k < K; k++) nmatches[k] = 0;
    const int n_out = fixed->n; // mode 0: indexed by the frame's features; mode 1: by the current key frame's
    if (n_out == 0 || K == 0) return ORBGPU_OK;
    auto kf_of = [&](int k) { return mode == 0 ? others[k] : fixed; };
    auto f_of = [&](int k) { return mode == 0 ? fixed : others[k]; };
    int f_n_max = 0, f_node_max = 0;
    for (int k = 0; k < K; k++) {
        f_n_max = std::max(f_n_max, f_of(k)->n);
        f_node_max = std::max(f_node_max, f_of(k)->fv_max_node);
    }
    const size_t ffb = bow_ff_bytes(n_out), zb = bow_zero_bytes(f_n_max);
    size_t total = K * (ffb + zb) + align256((size_t)K * n_out * 4) + 2048, mask_bytes = align256(fixed->n + 1);
    std::vector<size_t> mask_off(K);
    for (int k = 0; k < K; k++) {
        total += bow_scratch_bytes(kf_of(k), f_of(k), mode, nullptr);
        mask_off[k] = mask_bytes;
        mask_bytes += align256(others[k]->n + 1);
    }
    const size_t arg_bytes = align256((size_t)K * sizeof(BowBatchArg));
    rc = stage_reserve(ctx, mask_bytes + arg_bytes + 256);
    if (rc) return rc;
    rc = arena_reserve(ctx, total + align256(mask_bytes) + arg_bytes + 1024);
    if (rc) return rc;
    char *ff = (char *)arena_take(ctx, K * ffb), *zero = (char *)arena_take(ctx, K * zb);
    int32_t *d_out = (int32_t *)arena_take(ctx, (size_t)K * n_out * 4 + 4 * K);
    uint8_t *d_up = (uint8_t *)arena_take(ctx, mask_bytes + arg_bytes);
    if (!ff || !zero || !d_out || !d_up) return orbgpu_fail(ORBGPU_ERR_OVERFLOW, "arena exhausted");
    uint8_t *d_masks = d_up + arg_bytes; // [0] the fixed side's mask (mode 1), then one per pair
    // per-pair arguments of the batch kernels + the validity masks: one pinned block, one H2D copy
    BowBatchArg *h_args = (BowBatchArg *)ctx->h_stage;
    int max_nodes = 0, n_big = 0;
    std::vector<char> is_big(K, 0);
    if (mode == 1 && fixed->n) memcpy(ctx->h_stage + arg_bytes, fixed_valid, fixed->n);
    for (int k = 0; k < K; k++) {
        bool big = false;
        bow_scratch_bytes(kf_of(k), f_of(k), mode, &big);
        is_big[k] = big;
        n_big += big;
        const BowBlock blk = bow_block(ff + k * ffb, zero + k * zb, n_out);
        BowBatchArg &A = h_args[k];
        A.other = frame_view(others[k]);
        A.other_valid = d_masks + mask_off[k];
        A.match = big ? nullptr : blk.d_match;
        A.bin_of = blk.d_bin; A.matched2 = blk.d_m2; A.hist = blk.d_hist;
        A.n_nodes = (big || kf_of(k)->fv_n_nodes == 0 || f_of(k)->fv_n_nodes == 0) ? 0 : kf_of(k)->fv_n_nodes; // else its blocks exit at once
        max_nodes = std::max(max_nodes, A.n_nodes);
        if (others[k]->n) memcpy(ctx->h_stage + arg_bytes + mask_off[k], other_valid[k], others[k]->n);
    }
    CU_TRY(cudaMemcpyAsync(d_up, ctx->h_stage, arg_bytes + mask_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemsetAsync(ff, 0xFF, K * ffb, ctx->stream)); // the match / bin vectors of ALL pairs
    CU_TRY(cudaMemsetAsync(zero, 0, K * zb, ctx->stream));   // histograms, counts and taken flags of ALL pairs
    if (max_nodes > 0) { // ONE launch for all the ordinary pairs: blockIdx.y = pair, blockIdx.x = node
        const int threads = f_node_max <= 32 ? 32 : (f_node_max <= 512 ? 128 : (f_node_max <= 1024 ? 256 : 1024));
        int stage_cap = f_node_max > 64 ? f_node_max : 0;
        if (stage_cap > 4800) stage_cap = 4800;
        stage_cap = (stage_cap + 15) & ~15;
        const size_t smem = (size_t)stage_cap * 41;
        if (mode == 0)
            bow_match_batch_kernel<0><<<dim3(max_nodes, K), threads, smem, ctx->stream>>>((const BowBatchArg *)d_up, frame_view(fixed), d_masks, nnratio,
                                                                                          check_ori, ctx->d_counters, stage_cap);
        else
            bow_match_batch_kernel<1><<<dim3(max_nodes, K), threads, smem, ctx->stream>>>((const BowBatchArg *)d_up, frame_view(fixed), d_masks, nnratio,
                                                                                          check_ori, ctx->d_counters, stage_cap);
        bow_cull_batch_kernel<<<K, 256, 0, ctx->stream>>>((const BowBatchArg *)d_up, n_out, check_ori);
        ctx->launches += 2;
        CU_TRY(cudaGetLastError());
    }
    for (int k = 0; k < K && n_big; k++) { // pairs with a big node pair (one root bucket): the fixed-point path, their own launches
        if (!is_big[k]) continue;
        const uint8_t *d_kfv = mode == 0 ? d_masks + mask_off[k] : d_masks, *d_fv = mode == 0 ? d_masks : d_masks + mask_off[k];
        rc = bow_enqueue(ctx, mode, kf_of(k), f_of(k), d_kfv, d_fv, nnratio, check_ori, bow_block(ff + k * ffb, zero + k * zb, n_out));
        if (rc) return rc;
    }
    // results to one contiguous block: [K][n_out] matches (pitch ffb -> n_out * 4), then the K counts (zero + k * zb + 128)
    int32_t *d_nm = d_out + (size_t)K * n_out;
    CU_TRY(cudaMemcpy2DAsync(d_out, (size_t)n_out * 4, ff, ffb, (size_t)n_out * 4, K, cudaMemcpyDeviceToDevice, ctx->stream));
    CU_TRY(cudaMemcpy2DAsync(d_nm, 4, zero + 128, zb, 4, K, cudaMemcpyDeviceToDevice, ctx->stream));
    const OutPiece out[2] = {{match_out, d_out, (size_t)K * n_out * 4}, {nmatches, d_nm, (size_t)K * 4}};
    return ctx_download(ctx, out, 2);
}

} // namespace

__global__ void three_maxima_kernel(const int *histo, int L, int *ind)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) three_maxima(histo, L, ind[0], ind[1], ind[2]);
}

// ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2341-2383) on the device, exposed for parity tests
extern "C" int orbgpu_compute_three_maxima(orbgpu_ctx *ctx, const int32_t *histo, int32_t L, int32_t *ind)
{
    ARG_TRY(ctx && histo && ind && L > 0 && L <= 4096);
    int rc = ctx_begin(ctx);
    if (rc) return rc;
    rc = arena_reserve(ctx, align256((size_t)L * 4) + 256);
    if (rc) return rc;
    int *d_h = (int *)arena_take(ctx, (size_t)L * 4), *d_i = (int *)arena_take(ctx, 256);
    CU_TRY(cudaMemcpyAsync(d_h, histo, (size_t)L * 4, cudaMemcpyHostToDevice, ctx->stream));
    three_maxima_kernel<<<1, 32, 0, ctx->stream>>>(d_h, L, d_i);
    LAUNCH_COUNT(ctx);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(ind, d_i, 12, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return ORBGPU_OK;
}

extern "C" int orbgpu_search_by_bow_kf_f(orbgpu_ctx *ctx, const orbgpu_frame *kf, const orbgpu_frame *f, const uint8_t *kf_mp_valid,
                                         float nnratio, int32_t check_ori, int32_t *match_f2kf, int32_t *nmatches)
{
    ARG_TRY(ctx && kf && f && nmatches && (f->n == 0 || match_f2kf) && (kf->n == 0 || kf_mp_valid));
    ARG_TRY(f->n < (1 << 20));
    return run_bow(ctx, 0, kf, f, kf_mp_valid, nullptr, nnratio, check_ori, match_f2kf, nmatches);
}

extern "C" int orbgpu_search_by_bow_kf_kf(orbgpu_ctx *ctx, const orbgpu_frame *kf1, const orbgpu_frame *kf2,
                                          const uint8_t *kf1_mp_valid, const uint8_t *kf2_mp_valid, float nnratio, int32_t check_ori,
                                          int32_t *match_12, int32_t *nmatches)
{
    ARG_TRY(ctx && kf1 && kf2 && nmatches && (kf1->n == 0 || (match_12 && kf1_mp_valid)) && (kf2->n == 0 || kf2_mp_valid));
    ARG_TRY(kf2->n < (1 << 20));
    return run_bow(ctx, 1, kf1, kf2, kf1_mp_valid, kf2_mp_valid, nnratio, check_ori, match_12, nmatches);
}

extern "C" int orbgpu_search_by_bow_kf_f_batch(orbgpu_ctx *ctx, int32_t n_kf, const orbgpu_frame *const *kfs, const orbgpu_frame *f,
                                               const uint8_t *const *kf_mp_valid, float nnratio, int32_t check_ori, int32_t *match_f2kf,
                                               int32_t *nmatches)
{
    ARG_TRY(ctx && f && n_kf >= 0 && (n_kf == 0 || (kfs && kf_mp_valid && nmatches)) && (f->n == 0 || n_kf == 0 || match_f2kf));
    ARG_TRY(f->n < (1 << 20));
    for (int k = 0; k < n_kf; k++) ARG_TRY(kfs[k] && (kfs[k]->n == 0 || kf_mp_valid[k]));
    return run_bow_batch(ctx, 0, n_kf, kfs, f, kf_mp_valid, nullptr, nnratio, check_ori, match_f2kf, nmatches);
}

extern "C" int orbgpu_search_by_bow_kf_kf_batch(orbgpu_ctx *ctx, const orbgpu_frame *kf1, const uint8_t *kf1_mp_valid, int32_t n_kf,
                                                const orbgpu_frame *const *kf2s, const uint8_t *const *kf2_mp_valid, float nnratio,
                                                int32_t check_ori, int32_t *match_12, int32_t *nmatches)
{
    ARG_TRY(ctx && kf1 && n_kf >= 0 && (n_kf == 0 || (kf2s && kf2_mp_valid && nmatches)) && (kf1->n == 0 || n_kf == 0 || (match_12 && kf1_mp_valid)));
    for (int k = 0; k < n_kf; k++) ARG_TRY(kf2s[k] && kf2s[k]->n < (1 << 20) && (kf2s[k]->n == 0 || kf2_mp_valid[k]));
    return run_bow_batch(ctx, 1, n_kf, kf2s, kf1, kf2_mp_valid, kf1_mp_valid, nnratio, check_ori, match_12, nmatches);
}

int search_bow_device_init()
{
    int rc;
    if ((rc = set_max_dyn_smem(bow_match_batch_kernel<0>)) || (rc = set_max_dyn_smem(bow_match_batch_kernel<1>))) return rc;
    if ((rc = set_max_dyn_smem(bow_match_kernel<0>)) || (rc = set_max_dyn_smem(bow_match_kernel<1>)) ||
        (rc = set_max_dyn_smem(bow_big_resolve_kernel<0>)) || (rc = set_max_dyn_smem(bow_big_resolve_kernel<1>)))
        return rc;
    return ORBGPU_OK;
}
