/*
 * orb_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See orb_oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off pins the reference's written fp32 operation order (no FMA fusion);
 * the GPU kernels use __fmul_rn/__fadd_rn for the same reason.
 */
#include "orb_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static __thread int64_t g_comparisons = 0;

int64_t oracle_comparisons(void) { return g_comparisons; }
void oracle_comparisons_reset(void) { g_comparisons = 0; }

/* ORBmatcher.cc:2388-2408: 8 x (xor, SWAR popcount) over int32 words */
int oracle_descriptor_distance(const uint8_t *a, const uint8_t *b)
{
    int dist = 0;
    g_comparisons++;
    for (int i = 0; i < 8; i++) {
        uint32_t wa, wb;
        memcpy(&wa, a + 4 * i, 4);
        memcpy(&wb, b + 4 * i, 4);
        uint32_t v = wa ^ wb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

/* Frame.cc:973-989 PosInGrid: round-half-away, reject outside [0,cols) x [0,rows) */
static int pos_in_grid(const orbgpu_frame_host *f, float x, float y, int *px, int *py)
{
    int posX = (int)roundf((x - f->min_x) * f->grid_inv_w);
    int posY = (int)roundf((y - f->min_y) * f->grid_inv_h);
    if (posX < 0 || posX >= f->grid_cols || posY < 0 || posY >= f->grid_rows)
        return 0;
    *px = posX;
    *py = posY;
    return 1;
}

/* Frame.cc:469-507 AssignFeaturesToGrid: push_back in ascending i => in-cell ascending id */
void oracle_grid_build(const orbgpu_frame_host *f, int32_t *cell_start, int32_t *cell_items)
{
    const int ncell = f->grid_cols * f->grid_rows;
    int32_t *cnt = (int32_t *)calloc((size_t)ncell + 1, sizeof(int32_t));
    int32_t *cell_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)(f->n > 0 ? f->n : 1));
    for (int i = 0; i < f->n; i++) {
        int px, py;
        if (pos_in_grid(f, f->kp_xy[2 * i], f->kp_xy[2 * i + 1], &px, &py)) {
            cell_of[i] = px * f->grid_rows + py;
            cnt[cell_of[i]]++;
        } else
            cell_of[i] = -1;
    }
    int32_t acc = 0;
    for (int c = 0; c < ncell; c++) {
        cell_start[c] = acc;
        acc += cnt[c];
        cnt[c] = cell_start[c];
    }
    cell_start[ncell] = acc;
    for (int i = 0; i < f->n; i++)
        if (cell_of[i] >= 0)
            cell_items[cnt[cell_of[i]]++] = i;
    free(cnt);
    free(cell_of);
}

/* Frame.cc:868-962 */
int oracle_features_in_area(const orbgpu_frame_host *f, const int32_t *cell_start, const int32_t *cell_items, float x,
                            float y, float r, int min_level, int max_level, int32_t *out_idx)
{
    int n = 0;
    const float factorX = r, factorY = r;
    int nMinCellX = (int)floorf((x - f->min_x - factorX) * f->grid_inv_w); /* :886 */
    if (nMinCellX < 0) nMinCellX = 0;
    if (nMinCellX >= f->grid_cols) return 0; /* :889 */
    int nMaxCellX = (int)ceilf((x - f->min_x + factorX) * f->grid_inv_w); /* :895 */
    if (nMaxCellX > f->grid_cols - 1) nMaxCellX = f->grid_cols - 1;
    if (nMaxCellX < 0) return 0; /* :898 */
    int nMinCellY = (int)floorf((y - f->min_y - factorY) * f->grid_inv_h); /* :904 */
    if (nMinCellY < 0) nMinCellY = 0;
    if (nMinCellY >= f->grid_rows) return 0;
    int nMaxCellY = (int)ceilf((y - f->min_y + factorY) * f->grid_inv_h); /* :910 */
    if (nMaxCellY > f->grid_rows - 1) nMaxCellY = f->grid_rows - 1;
    if (nMaxCellY < 0) return 0;

    const int bCheckLevels = (min_level > 0) || (max_level >= 0); /* :919 */

    for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const int c = ix * f->grid_rows + iy;
            for (int j = cell_start[c]; j < cell_start[c + 1]; j++) {
                const int idx = cell_items[j];
                if (bCheckLevels) { /* :939-948 */
                    if (f->octave[idx] < min_level) continue;
                    if (max_level >= 0)
                        if (f->octave[idx] > max_level) continue;
                }
                const float distx = f->kp_xy[2 * idx] - x;
                const float disty = f->kp_xy[2 * idx + 1] - y;
                if (fabsf(distx) < factorX && fabsf(disty) < factorY) /* :955 strict */
                    out_idx[n++] = idx;
            }
        }
    }
    return n;
}

/* ORBmatcher.cc:2341-2383 */
void oracle_compute_three_maxima(const int32_t *histo_sizes, int L, int32_t *ind)
{
    int max1 = 0, max2 = 0, max3 = 0;
    int ind1 = -1, ind2 = -1, ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = histo_sizes[i];
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
    if ((float)max2 < 0.1f * (float)max1) {
        ind2 = -1; ind3 = -1;
    } else if ((float)max3 < 0.1f * (float)max1) {
        ind3 = -1;
    }
    ind[0] = ind1; ind[1] = ind2; ind[2] = ind3;
}

/* rotation-histogram bin, e.g. ORBmatcher.cc:829-837: factor = 1.0f/HISTO_LENGTH */
static int rot_bin(float a1, float a2)
{
    const float factor = 1.0f / ORBGPU_HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == ORBGPU_HISTO_LENGTH) bin = 0;
    return bin;
}

/* growable per-bin lists (vector<int> rotHist[HISTO_LENGTH]) */
typedef struct {
    int32_t *v[ORBGPU_HISTO_LENGTH];
    int32_t n[ORBGPU_HISTO_LENGTH];
    int32_t cap[ORBGPU_HISTO_LENGTH];
} rot_hist;

static void rh_init(rot_hist *h)
{
    for (int i = 0; i < ORBGPU_HISTO_LENGTH; i++) {
        h->cap[i] = 64;
        h->n[i] = 0;
        h->v[i] = (int32_t *)malloc(sizeof(int32_t) * 64);
    }
}
static void rh_push(rot_hist *h, int bin, int32_t val)
{
    if (bin < 0 || bin >= ORBGPU_HISTO_LENGTH) return; /* assert in the reference (:838) */
    if (h->n[bin] == h->cap[bin]) {
        h->cap[bin] *= 2;
        h->v[bin] = (int32_t *)realloc(h->v[bin], sizeof(int32_t) * (size_t)h->cap[bin]);
    }
    h->v[bin][h->n[bin]++] = val;
}
static void rh_free(rot_hist *h)
{
    for (int i = 0; i < ORBGPU_HISTO_LENGTH; i++) free(h->v[i]);
}

/* ORBmatcher.cc:735-878 */
int oracle_search_for_initialization(const orbgpu_frame_host *f1, const orbgpu_frame_host *f2, float *prev_matched_xy,
                                     int window_size, float nnratio, int check_ori, int32_t *matches12)
{
    int nmatches = 0;
    const int n1 = f1->n, n2 = f2->n;
    for (int i = 0; i < n1; i++) matches12[i] = -1; /* :739 */
    rot_hist rh;
    rh_init(&rh);
    int32_t *cs = (int32_t *)malloc(sizeof(int32_t) * (size_t)(f2->grid_cols * f2->grid_rows + 1));
    int32_t *ci = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n2 + 1));
    oracle_grid_build(f2, cs, ci);
    int32_t *vMatchedDistance = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n2 + 1));
    int32_t *vnMatches21 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n2 + 1));
    int32_t *vIndices2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n2 + 1));
    for (int i = 0; i < n2; i++) {
        vMatchedDistance[i] = INT_MAX; /* :752 */
        vnMatches21[i] = -1;           /* :754 */
    }
    for (int i1 = 0; i1 < n1; i1++) {
        const int level1 = f1->octave[i1];
        if (level1 > 0) continue; /* :762 */
        const int nc = oracle_features_in_area(f2, cs, ci, prev_matched_xy[2 * i1], prev_matched_xy[2 * i1 + 1],
                                               (float)window_size, level1, level1, vIndices2); /* :768 */
        if (nc == 0) continue;
        const uint8_t *d1 = f1->desc + 32 * (size_t)i1;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int c = 0; c < nc; c++) {
            const int i2 = vIndices2[c];
            const int dist = oracle_descriptor_distance(d1, f2->desc + 32 * (size_t)i2);
            if (vMatchedDistance[i2] <= dist) continue; /* :790 */
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2;
            } else if (dist < bestDist2) {
                bestDist2 = dist;
            }
        }
        if (bestDist <= ORBGPU_TH_LOW) { /* :807 */
            if ((float)bestDist < (float)bestDist2 * nnratio) { /* :810 */
                if (vnMatches21[bestIdx2] >= 0) { /* :813-817 */
                    matches12[vnMatches21[bestIdx2]] = -1;
                    nmatches--;
                }
                matches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (check_ori) /* :826-840 */
                    rh_push(&rh, rot_bin(f1->angle[i1], f2->angle[bestIdx2]), i1);
            }
        }
    }
    if (check_ori) { /* :846-869 */
        int32_t ind[3];
        oracle_compute_three_maxima(rh.n, ORBGPU_HISTO_LENGTH, ind);
        for (int i = 0; i < ORBGPU_HISTO_LENGTH; i++) {
            if (i == ind[0] || i == ind[1] || i == ind[2]) continue;
            for (int j = 0; j < rh.n[i]; j++) {
                const int idx1 = rh.v[i][j];
                if (matches12[idx1] >= 0) {
                    matches12[idx1] = -1;
                    nmatches--;
                }
            }
        }
    }
    for (int i1 = 0; i1 < n1; i1++) /* :873-875 */
        if (matches12[i1] >= 0) {
            prev_matched_xy[2 * i1] = f2->kp_xy[2 * matches12[i1]];
            prev_matched_xy[2 * i1 + 1] = f2->kp_xy[2 * matches12[i1] + 1];
        }
    rh_free(&rh);
    free(cs); free(ci); free(vMatchedDistance); free(vnMatches21); free(vIndices2);
    return nmatches;
}

/* ORBmatcher.cc:245-252 */
static float radius_by_viewing_cos(float viewCos)
{
    if (viewCos > 0.998) /* float -> double compare */
        return 2.5f;
    else
        return 4.0f;
}

/* ORBmatcher.cc:44-242, Nleft == -1 */
int oracle_search_by_projection_local(const orbgpu_frame_host *f, const orbgpu_mappoints_host *mps, float th,
                                      int far_points, float th_far_points, float nnratio, const int32_t *kp_prior_obs,
                                      int32_t *kp_mp)
{
    int nmatches = 0;
    const int bFactor = th != 1.0; /* :49 (float vs double literal) */
    int32_t *cs = (int32_t *)malloc(sizeof(int32_t) * (size_t)(f->grid_cols * f->grid_rows + 1));
    int32_t *ci = (int32_t *)malloc(sizeof(int32_t) * (size_t)(f->n + 1));
    int32_t *vIndices = (int32_t *)malloc(sizeof(int32_t) * (size_t)(f->n + 1));
    /* Observations() of the map point currently held by each keypoint (0 == none/unobserved) */
    int32_t *cur_obs = (int32_t *)malloc(sizeof(int32_t) * (size_t)(f->n + 1));
    oracle_grid_build(f, cs, ci);
    for (int i = 0; i < f->n; i++) cur_obs[i] = kp_prior_obs[i];

    for (int iMP = 0; iMP < mps->n; iMP++) {
        if (!mps->in_view[iMP]) continue;                                   /* :55 (no right view) */
        if (far_points && mps->depth[iMP] > th_far_points) continue;        /* :58 */
        if (mps->bad[iMP]) continue;                                        /* :61 */
        const int nPredictedLevel = mps->scale_level[iMP];
        float r = radius_by_viewing_cos(mps->view_cos[iMP]);                /* :71 */
        if (bFactor) r *= th;                                               /* :74 */
        const int nc = oracle_features_in_area(f, cs, ci, mps->proj_xy[2 * iMP], mps->proj_xy[2 * iMP + 1],
                                               r * f->scale_factors[nPredictedLevel], nPredictedLevel - 1,
                                               nPredictedLevel, vIndices);  /* :78-81 */
        if (nc == 0) continue;
        const uint8_t *dMP = mps->desc + 32 * (size_t)iMP;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int c = 0; c < nc; c++) {
            const int idx = vIndices[c];
            if (cur_obs[idx] > 0) continue; /* :102-104 */
            if (f->u_right && f->u_right[idx] > 0) { /* :107-117 */
                const float er = fabsf(mps->proj_xr[iMP] - f->u_right[idx]);
                if (er > r * f->scale_factors[nPredictedLevel]) continue;
            }
            const int dist = oracle_descriptor_distance(dMP, f->desc + 32 * (size_t)idx);
            if (dist < bestDist) { /* :125-141 */
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = f->octave[idx];
                bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = f->octave[idx];
                bestDist2 = dist;
            }
        }
        if (bestDist <= ORBGPU_TH_HIGH) { /* :147 */
            if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue; /* :151 */
            if (bestLevel != bestLevel2 || (float)bestDist <= nnratio * (float)bestDist2) {        /* :154 */
                kp_mp[bestIdx] = iMP; /* :156 */
                cur_obs[bestIdx] = mps->n_obs[iMP];
                nmatches++;
            }
        }
    }
    free(cs); free(ci); free(vIndices); free(cur_obs);
    return nmatches;
}

/* The common search core of the self-projecting overloads, sequential restatement:
 *   SearchByProjection(Cur, Last)            ORBmatcher.cc:2026-2101  (window :2031-2041, skip :2046-2049, stereo :2052-2059,
 *                                            best-only :2062-2068, accept :2071-2074, histogram :2076-2088, cull :2163-2186)
 *   SearchByProjection(Cur, KF, found)       :2262-2320  (skip `if (CurrentFrame.mvpMapPoints[*vit]) continue`)
 *   SearchByProjection(KF, Sim3, ...)        :566-614    (skip `if (vpMatched[idx]) continue`, level filter :583-584)
 *   Fuse                                     :1452-1510  (level filter :1461-1462, chi2 gate :1463-1492)
 *   Fuse (Sim3), SearchBySim3                :1636-1650, :1780-1800 / :1870-1890
 * The per-point prologue (pose transform, projection, frustum gates, PredictScale) is an input. */
int oracle_search_projected(const orbgpu_frame_host *f, const orbgpu_projpoints_host *pts, const orbgpu_projsearch_params *prm,
                            const uint8_t *kp_locked, int32_t *best_idx, int32_t *best_dist, int32_t *kp_owner)
{
    int nmatches = 0;
    int32_t *cs = (int32_t *)malloc(sizeof(int32_t) * (size_t)(f->grid_cols * f->grid_rows + 1));
    int32_t *ci = (int32_t *)malloc(sizeof(int32_t) * (size_t)(f->n + 1));
    int32_t *vIndices = (int32_t *)malloc(sizeof(int32_t) * (size_t)(f->n + 1));
    uint8_t *locked = (uint8_t *)calloc((size_t)f->n + 1, 1);
    int32_t *owner = (int32_t *)malloc(sizeof(int32_t) * (size_t)(f->n + 1));
    /* rotHist[bin] holds keypoint indices in acceptance order; one flat array of (bin, keypoint) is enough */
    int32_t *hist_bin = (int32_t *)malloc(sizeof(int32_t) * (size_t)(pts->n + 1));
    int32_t *hist_kp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(pts->n + 1));
    int32_t histo[ORBGPU_HISTO_LENGTH];
    int n_hist = 0;
    const float factor = 1.0f / ORBGPU_HISTO_LENGTH;
    memset(histo, 0, sizeof(histo));
    oracle_grid_build(f, cs, ci);
    for (int i = 0; i < f->n; i++) {
        locked[i] = kp_locked ? kp_locked[i] : 0;
        owner[i] = -1;
    }
    for (int m = 0; m < pts->n; m++) {
        best_idx[m] = -1;
        best_dist[m] = 256;
        if (!pts->active[m]) continue;
        const float u = pts->uv[2 * m], v = pts->uv[2 * m + 1], radius = pts->radius[m];
        const int nc = oracle_features_in_area(f, cs, ci, u, v, radius, pts->min_level[m], pts->max_level[m], vIndices);
        if (nc == 0) continue;
        const uint8_t *dMP = pts->desc + 32 * (size_t)m;
        int bestDist = 256, bestIdx = -1;
        for (int c = 0; c < nc; c++) {
            const int idx = vIndices[c];
            if (prm->ordered && locked[idx]) continue;
            if (prm->stereo_gate && f->u_right && f->u_right[idx] > 0) {
                const float er = fabsf(pts->ur[m] - f->u_right[idx]);
                if (er > radius) continue;
            }
            if (prm->chi2_gate) {
                const float ex = u - f->kp_xy[2 * idx], ey = v - f->kp_xy[2 * idx + 1];
                const float inv = prm->inv_level_sigma2[f->octave[idx]];
                if (f->u_right && f->u_right[idx] >= 0) {
                    const float er = pts->ur[m] - f->u_right[idx];
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * inv > 7.8) continue;
                } else {
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * inv > 5.99) continue;
                }
            }
            const int dist = oracle_descriptor_distance(dMP, f->desc + 32 * (size_t)idx);
            if (dist < bestDist) {
                bestDist = dist;
                bestIdx = idx;
            }
        }
        if (bestIdx >= 0 && (float)bestDist <= prm->max_dist) {
            best_idx[m] = bestIdx;
            best_dist[m] = bestDist;
            owner[bestIdx] = m;
            if (!pts->locks || pts->locks[m]) locked[bestIdx] = 1;
            nmatches++;
            if (prm->check_ori) {
                float rot = pts->angle[m] - f->angle[bestIdx];
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == ORBGPU_HISTO_LENGTH) bin = 0;
                if (bin >= 0 && bin < ORBGPU_HISTO_LENGTH) {
                    histo[bin]++;
                    hist_bin[n_hist] = bin;
                    hist_kp[n_hist] = bestIdx;
                    n_hist++;
                }
            }
        }
    }
    if (prm->check_ori) {
        int32_t ind[3];
        oracle_compute_three_maxima(histo, ORBGPU_HISTO_LENGTH, ind);
        for (int e = 0; e < n_hist; e++)
            if (hist_bin[e] != ind[0] && hist_bin[e] != ind[1] && hist_bin[e] != ind[2]) {
                owner[hist_kp[e]] = -1;
                nmatches--;
            }
    }
    if (kp_owner) memcpy(kp_owner, owner, sizeof(int32_t) * (size_t)f->n);
    free(cs); free(ci); free(vIndices); free(locked); free(owner); free(hist_bin); free(hist_kp);
    return nmatches;
}

/* KeyFrameDatabase candidate scoring: common words per key frame (the inverted-file walk of
 * KeyFrameDatabase.cc:928-943 counts one per (query word, key frame holding it)) and L1Scoring::score
 * (Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:23-68) of the query against every key frame. */
void oracle_bow_score_l1(const orbgpu_bowdb_host *db, int32_t nq, const uint32_t *q_words, const double *q_values,
                         int32_t *common_words, double *scores)
{
    for (int kf = 0; kf < db->n_kf; kf++) {
        int a = 0, b = db->offsets[kf];
        const int b_end = db->offsets[kf + 1];
        double score = 0;
        int n_common = 0;
        while (a < nq && b < b_end) {
            const double vi = q_values[a], wi = db->values[b];
            if (q_words[a] == db->words[b]) {
                score += fabs(vi - wi) - fabs(vi) - fabs(wi);
                n_common++;
                ++a; ++b;
            } else if (q_words[a] < db->words[b]) {
                while (a < nq && q_words[a] < db->words[b]) ++a; /* v1.lower_bound(v2_it->first) */
            } else {
                while (b < b_end && db->words[b] < q_words[a]) ++b; /* v2.lower_bound(v1_it->first) */
            }
        }
        common_words[kf] = n_common;
        scores[kf] = -score / 2.0;
    }
}

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:444-535) for a batch of map points whose observation
 * descriptors are given as a CSR.  The real MapPoint.cc cannot be compiled here (OpenCV / Eigen / Boost / g2o), so this is a
 * restatement only; tests/test_oracle_golden.py cross-checks it with an independent numpy evaluation. */
static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }
void oracle_compute_distinctive_descriptors(int32_t n_mp, const int32_t *offsets, const uint8_t *desc, int32_t *best_idx,
                                            int32_t *best_median)
{
    for (int p = 0; p < n_mp; p++) {
        const int s = offsets[p], N = offsets[p + 1] - s;
        best_idx[p] = -1;
        if (best_median) best_median[p] = -1;
        if (N <= 0) continue; /* :455-456, :481-482 */
        float *Distances = (float *)malloc(sizeof(float) * (size_t)N * (size_t)N); /* float Distances[N][N] (:487) */
        int *vDists = (int *)malloc(sizeof(int) * (size_t)N);
        for (int i = 0; i < N; i++) {
            Distances[(size_t)i * N + i] = 0;
            for (int j = i + 1; j < N; j++) {
                const int distij = oracle_descriptor_distance(desc + 32 * (size_t)(s + i), desc + 32 * (size_t)(s + j));
                Distances[(size_t)i * N + j] = (float)distij;
                Distances[(size_t)j * N + i] = (float)distij;
            }
        }
        int BestMedian = 0x7FFFFFFF, BestIdx = 0;
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < N; j++) vDists[j] = (int)Distances[(size_t)i * N + j];
            qsort(vDists, (size_t)N, sizeof(int), cmp_int);
            const int median = vDists[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) {
                BestMedian = median;
                BestIdx = i;
            }
        }
        best_idx[p] = BestIdx;
        if (best_median) best_median[p] = BestMedian;
        free(Distances); free(vDists);
    }
}

/* Coarse stage of Frame::ComputeStereoMatches (src/Frame.cc:1117-1247).  Pinned on the reference's own text of that stage
 * (oracle/extract_ref.py cuts Frame.cc:1117-1247 into oracle/_ref, ref_stereo_coarse_match; tests/test_oracle_vs_ref.py).  Rows
 * outside [0, n_rows) are undefined behaviour in the reference (vRowIndices[yi]) and are ignored here. */
void oracle_stereo_coarse_match(int32_t n_left, const uint8_t *desc_l, const float *kp_xy_l, const int32_t *octave_l, int32_t n_right,
                                const uint8_t *desc_r, const float *kp_xy_r, const int32_t *octave_r, const float *scale_factors,
                                int32_t n_rows, float mb, float mbf, int32_t *best_idx_r, int32_t *best_dist)
{
    const int thOrbDist = (ORBGPU_TH_HIGH + ORBGPU_TH_LOW) / 2; /* :1139 */
    int32_t *cnt = (int32_t *)calloc((size_t)n_rows + 1, sizeof(int32_t));
    int32_t *start = (int32_t *)calloc((size_t)n_rows + 2, sizeof(int32_t));
    for (int pass = 0; pass < 2; pass++) { /* vRowIndices as a CSR: count, then fill in ascending iR (:1147-1156) */
        int32_t *items = NULL;
        if (pass == 1) {
            for (int r = 0; r < n_rows; r++) start[r + 1] = start[r] + cnt[r];
            items = (int32_t *)malloc(sizeof(int32_t) * (size_t)(start[n_rows] + 1));
            memset(cnt, 0, sizeof(int32_t) * (size_t)n_rows);
        }
        for (int iR = 0; iR < n_right; iR++) {
            const float kpY = kp_xy_r[2 * iR + 1];
            const float r = 2.0f * scale_factors[octave_r[iR]];
            const int maxr = (int)ceilf(kpY + r), minr = (int)floorf(kpY - r);
            for (int yi = minr; yi <= maxr; yi++) {
                if (yi < 0 || yi >= n_rows) continue;
                if (pass == 1) items[start[yi] + cnt[yi]] = iR;
                cnt[yi]++;
            }
        }
        if (pass == 1) {
            const float minZ = mb, minD = 0, maxD = mbf / minZ; /* :1160-1163 */
            for (int iL = 0; iL < n_left; iL++) {
                best_idx_r[iL] = -1;
                best_dist[iL] = ORBGPU_TH_HIGH;
                const int levelL = octave_l[iL];
                const float vL = kp_xy_l[2 * iL + 1], uL = kp_xy_l[2 * iL];
                if (vL < 0 || (int)vL >= n_rows) continue;
                const int row = (int)vL;
                if (cnt[row] == 0) continue; /* :1180 */
                const float minU = uL - maxD, maxU = uL - minD;
                if (maxU < 0) continue; /* :1186 */
                int bestDist = ORBGPU_TH_HIGH, bestIdxR = 0;
                for (int iC = 0; iC < cnt[row]; iC++) {
                    const int iR = items[start[row] + iC];
                    if (octave_r[iR] < levelL - 1 || octave_r[iR] > levelL + 1) continue; /* :1197 */
                    const float uR = kp_xy_r[2 * iR];
                    if (uR >= minU && uR <= maxU) {
                        const int dist = oracle_descriptor_distance(desc_l + 32 * (size_t)iL, desc_r + 32 * (size_t)iR);
                        if (dist < bestDist) {
                            bestDist = dist;
                            bestIdxR = iR;
                        }
                    }
                }
                best_dist[iL] = bestDist;
                if (bestDist < thOrbDist) best_idx_r[iL] = bestIdxR; /* :1214 */
            }
            free(items);
        }
    }
    free(cnt); free(start);
}

/* TemplatedVocabulary.h:1216-1258 */
void oracle_voc_transform(const orbgpu_voc_host *v, int32_t n, const uint8_t *desc, int levelsup, uint32_t *word_id,
                          uint32_t *node_id, double *weight)
{
    const int nid_level = v->L - levelsup;
    for (int i = 0; i < n; i++) {
        const uint8_t *feature = desc + 32 * (size_t)i;
        uint32_t nid = 0; /* the reference leaves it uninitialised when never written (:1151); see DESIGN.md */
        uint32_t final_id = 0;
        int current_level = 0;
        do {
            ++current_level;
            const int c0 = v->child_offsets[final_id], c1 = v->child_offsets[final_id + 1];
            final_id = v->child_ids[c0];
            double best_d = (double)oracle_descriptor_distance(feature, v->node_desc + 32 * (size_t)final_id);
            for (int c = c0 + 1; c < c1; c++) {
                const uint32_t id = v->child_ids[c];
                const double d = (double)oracle_descriptor_distance(feature, v->node_desc + 32 * (size_t)id);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (v->child_offsets[final_id + 1] > v->child_offsets[final_id]);
        if (word_id) word_id[i] = v->word_id[final_id];
        if (weight) weight[i] = v->weight[final_id];
        if (node_id) node_id[i] = nid;
    }
}

typedef struct { uint32_t key; int32_t idx; } kv_t;
static int kv_cmp(const void *a, const void *b)
{
    const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

/* BowVector.cpp:35-50 addWeight in feature order, :63-85 normalize(L1) */
int oracle_bowvector(int32_t n, const uint32_t *word_id, const double *weight, uint32_t *words, double *values)
{
    kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(n + 1));
    int m = 0;
    for (int i = 0; i < n; i++)
        if (weight[i] > 0) { kv[m].key = word_id[i]; kv[m].idx = i; m++; }
    qsort(kv, (size_t)m, sizeof(kv_t), kv_cmp);
    int nw = 0;
    for (int i = 0; i < m;) {
        int j = i;
        double acc = weight[kv[i].idx]; /* first insert */
        for (j = i + 1; j < m && kv[j].key == kv[i].key; j++) acc += weight[kv[j].idx]; /* += in feature order */
        words[nw] = kv[i].key;
        values[nw] = acc;
        nw++;
        i = j;
    }
    double norm = 0.0;
    for (int i = 0; i < nw; i++) norm += fabs(values[i]); /* ascending word id */
    if (norm > 0.0)
        for (int i = 0; i < nw; i++) values[i] /= norm;
    free(kv);
    return nw;
}

/* FeatureVector.cpp:32-46 */
int oracle_featvec(int32_t n, const uint32_t *node_id, const double *weight, uint32_t *node_ids, int32_t *offsets,
                   uint32_t *features)
{
    kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(n + 1));
    int m = 0;
    for (int i = 0; i < n; i++)
        if (weight[i] > 0) { kv[m].key = node_id[i]; kv[m].idx = i; m++; }
    qsort(kv, (size_t)m, sizeof(kv_t), kv_cmp);
    int nn = 0;
    for (int i = 0; i < m; i++) {
        if (i == 0 || kv[i].key != kv[i - 1].key) {
            node_ids[nn] = kv[i].key;
            offsets[nn] = i;
            nn++;
        }
        features[i] = (uint32_t)kv[i].idx;
    }
    offsets[nn] = m;
    free(kv);
    return nn;
}

/* ORBmatcher.cc:262-496, Nleft == -1 */
int oracle_search_by_bow_kf_f(const orbgpu_frame_host *kf, const orbgpu_frame_host *f, const uint8_t *kf_mp_valid,
                              float nnratio, int check_ori, int32_t *match_f2kf)
{
    int nmatches = 0;
    for (int i = 0; i < f->n; i++) match_f2kf[i] = -1; /* :268 */
    rot_hist rh;
    rh_init(&rh);
    int a = 0, b = 0;
    while (a < kf->fv_n_nodes && b < f->fv_n_nodes) { /* :292 */
        if (kf->fv_node_ids[a] == f->fv_node_ids[b]) {
            for (int iKF = kf->fv_offsets[a]; iKF < kf->fv_offsets[a + 1]; iKF++) {
                const int realIdxKF = (int)kf->fv_features[iKF];
                if (!kf_mp_valid[realIdxKF]) continue; /* :311-315 */
                const uint8_t *dKF = kf->desc + 32 * (size_t)realIdxKF;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int iF = f->fv_offsets[b]; iF < f->fv_offsets[b + 1]; iF++) {
                    const int realIdxF = (int)f->fv_features[iF];
                    if (match_f2kf[realIdxF] >= 0) continue; /* :335 */
                    const int dist = oracle_descriptor_distance(dKF, f->desc + 32 * (size_t)realIdxF);
                    if (dist < bestDist1) {
                        bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF;
                    } else if (dist < bestDist2) {
                        bestDist2 = dist;
                    }
                }
                if (bestDist1 <= ORBGPU_TH_LOW) { /* :392 */
                    if ((float)bestDist1 < nnratio * (float)bestDist2) { /* :395 */
                        match_f2kf[bestIdxF] = realIdxKF;
                        if (check_ori)
                            rh_push(&rh, rot_bin(kf->angle[realIdxKF], f->angle[bestIdxF]), bestIdxF);
                        nmatches++;
                    }
                }
            }
            a++; b++;
        } else if (kf->fv_node_ids[a] < f->fv_node_ids[b]) {
            a++; /* lower_bound(:460) lands on the first id >= the other: same as stepping */
        } else {
            b++;
        }
    }
    if (check_ori) { /* :470-493 */
        int32_t ind[3];
        oracle_compute_three_maxima(rh.n, ORBGPU_HISTO_LENGTH, ind);
        for (int i = 0; i < ORBGPU_HISTO_LENGTH; i++) {
            if (i == ind[0] || i == ind[1] || i == ind[2]) continue;
            for (int j = 0; j < rh.n[i]; j++) {
                match_f2kf[rh.v[i][j]] = -1;
                nmatches--;
            }
        }
    }
    rh_free(&rh);
    return nmatches;
}

/* ORBmatcher.cc:890-1043 */
int oracle_search_by_bow_kf_kf(const orbgpu_frame_host *kf1, const orbgpu_frame_host *kf2, const uint8_t *kf1_mp_valid,
                               const uint8_t *kf2_mp_valid, float nnratio, int check_ori, int32_t *match_12)
{
    int nmatches = 0;
    for (int i = 0; i < kf1->n; i++) match_12[i] = -1; /* :904 */
    uint8_t *vbMatched2 = (uint8_t *)calloc((size_t)kf2->n + 1, 1);
    rot_hist rh;
    rh_init(&rh);
    int a = 0, b = 0;
    while (a < kf1->fv_n_nodes && b < kf2->fv_n_nodes) {
        if (kf1->fv_node_ids[a] == kf2->fv_node_ids[b]) {
            for (int i1 = kf1->fv_offsets[a]; i1 < kf1->fv_offsets[a + 1]; i1++) {
                const int idx1 = (int)kf1->fv_features[i1];
                if (!kf1_mp_valid[idx1]) continue; /* :937-941 */
                const uint8_t *d1 = kf1->desc + 32 * (size_t)idx1;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int i2 = kf2->fv_offsets[b]; i2 < kf2->fv_offsets[b + 1]; i2++) {
                    const int idx2 = (int)kf2->fv_features[i2];
                    if (vbMatched2[idx2] || !kf2_mp_valid[idx2]) continue; /* :962-966 */
                    const int dist = oracle_descriptor_distance(d1, kf2->desc + 32 * (size_t)idx2);
                    if (dist < bestDist1) {
                        bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2;
                    } else if (dist < bestDist2) {
                        bestDist2 = dist;
                    }
                }
                if (bestDist1 < ORBGPU_TH_LOW) { /* :985 strict */
                    if ((float)bestDist1 < nnratio * (float)bestDist2) { /* :987 */
                        match_12[idx1] = bestIdx2;
                        vbMatched2[bestIdx2] = 1;
                        if (check_ori)
                            rh_push(&rh, rot_bin(kf1->angle[idx1], kf2->angle[bestIdx2]), idx1);
                        nmatches++;
                    }
                }
            }
            a++; b++;
        } else if (kf1->fv_node_ids[a] < kf2->fv_node_ids[b]) {
            a++;
        } else {
            b++;
        }
    }
    if (check_ori) { /* :1022-1040 */
        int32_t ind[3];
        oracle_compute_three_maxima(rh.n, ORBGPU_HISTO_LENGTH, ind);
        for (int i = 0; i < ORBGPU_HISTO_LENGTH; i++) {
            if (i == ind[0] || i == ind[1] || i == ind[2]) continue;
            for (int j = 0; j < rh.n[i]; j++) {
                match_12[rh.v[i][j]] = -1;
                nmatches--;
            }
        }
    }
    rh_free(&rh);
    free(vbMatched2);
    return nmatches;
}

/* Pinhole.cpp:203-218 with F12 given (row-major); unc = mvLevelSigma2[kp2.octave] */
static int epipolar_constrain(const float *F12, float x1, float y1, float x2, float y2, float unc)
{
    const float a = x1 * F12[0] + y1 * F12[3] + F12[6];
    const float b = x1 * F12[1] + y1 * F12[4] + F12[7];
    const float c = x1 * F12[2] + y1 * F12[5] + F12[8];
    const float num = a * x2 + b * y2 + c;
    const float den = a * a + b * b;
    if (den == 0) return 0;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * unc; /* double compare */
}

int oracle_epipolar_constrain(const float *F12, float x1, float y1, float x2, float y2, float unc)
{
    return epipolar_constrain(F12, x1, y1, x2, y2, unc);
}

/* Pinhole.cpp:64-71 */
void oracle_pinhole_project(const float *K, const float *xyz, float *uv)
{
    uv[0] = K[0] * xyz[0] / xyz[2] + K[2];
    uv[1] = K[1] * xyz[1] / xyz[2] + K[3];
}

/* MapPoint.cc:695-738: ratio = mfMaxDistance / currentDist; ceil(log(ratio) / mfLogScaleFactor) with the float overload of log,
 * clamped to [0, nLevels - 1] */
int oracle_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels)
{
    const float ratio = max_distance / current_dist;
    int nScale = (int)ceilf(logf(ratio) / log_scale_factor);
    if (nScale < 0)
        nScale = 0;
    else if (nScale >= n_levels)
        nScale = n_levels - 1;
    return nScale;
}

/* Frame::isInFrustum, Nleft == -1 (Frame.cc:676-782).  The frame's pose members are inputs exactly as the reference holds them:
 * mRcw (row-major 3x3), mtcw, mOw (Frame.h; filled by UpdatePoseMatrices with the host's Sophus).  Outputs = the MapPoint members
 * the function writes; a point rejected after the image-bounds test keeps its projection in proj_xy (:712-713) with in_view 0. */
void oracle_is_in_frustum(const orbgpu_frustum_host *fr, int32_t n, const float *world_pos, const float *normal,
                          const float *min_distance, const float *max_distance, uint8_t *in_view, float *proj_xy, float *proj_xr,
                          float *depth, int32_t *scale_level, float *view_cos)
{
    const float *R = fr->Rcw, *t = fr->tcw, *Ow = fr->Ow;
    for (int i = 0; i < n; i++) {
        in_view[i] = 0;            /* :682 */
        proj_xy[2 * i] = -1.f;     /* :683-684 */
        proj_xy[2 * i + 1] = -1.f;
        proj_xr[i] = 0.f; depth[i] = 0.f; scale_level[i] = 0; view_cos[i] = 0.f; /* untouched members: reported as 0 */
        const float *P = world_pos + 3 * i;
        float Pc[3];
        for (int r = 0; r < 3; r++) Pc[r] = (R[3 * r] * P[0] + R[3 * r + 1] * P[1] + R[3 * r + 2] * P[2]) + t[r]; /* :695 */
        const float Pc_dist = sqrtf(Pc[0] * Pc[0] + Pc[1] * Pc[1] + Pc[2] * Pc[2]);
        const float PcZ = Pc[2];
        const float invz = 1.0f / PcZ;
        if (PcZ < 0.0f) continue; /* :701 */
        float uv[2];
        oracle_pinhole_project(fr->K, Pc, uv); /* :704 */
        if (uv[0] < fr->min_x || uv[0] > fr->max_x) continue; /* :707 */
        if (uv[1] < fr->min_y || uv[1] > fr->max_y) continue;
        proj_xy[2 * i] = uv[0]; /* :712-713 */
        proj_xy[2 * i + 1] = uv[1];
        const float maxDistance = 1.2f * max_distance[i], minDistance = 0.8f * min_distance[i]; /* MapPoint.cc:665-678 */
        const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
        const float dist = sqrtf(PO[0] * PO[0] + PO[1] * PO[1] + PO[2] * PO[2]);
        if (dist < minDistance || dist > maxDistance) continue; /* :723 */
        const float *Pn = normal + 3 * i;
        const float viewCos = (PO[0] * Pn[0] + PO[1] * Pn[1] + PO[2] * Pn[2]) / dist; /* :730 */
        if (viewCos < fr->viewing_cos_limit) continue;
        const int nPredictedLevel = oracle_predict_scale(max_distance[i], dist, fr->log_scale_factor, fr->n_levels); /* :737 */
        in_view[i] = 1;
        proj_xr[i] = uv[0] - fr->mbf * invz; /* :743 */
        depth[i] = Pc_dist;
        scale_level[i] = nPredictedLevel;
        view_cos[i] = viewCos;
    }
}

/* per-keyframe FeatureVector CSR from the per-feature node ids of a kfset */
static int kf_featvec(const orbgpu_kfset_host *s, int kf, uint32_t *node_ids, int32_t *offsets, uint32_t *features)
{
    const int n = s->n_feat;
    const uint32_t *nid = s->node_id + (size_t)kf * n;
    kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(n + 1));
    int m = 0;
    for (int i = 0; i < n; i++)
        if (nid[i] != 0xFFFFFFFFu) { kv[m].key = nid[i]; kv[m].idx = i; m++; }
    qsort(kv, (size_t)m, sizeof(kv_t), kv_cmp);
    int nn = 0;
    for (int i = 0; i < m; i++) {
        if (i == 0 || kv[i].key != kv[i - 1].key) { node_ids[nn] = kv[i].key; offsets[nn] = i; nn++; }
        features[i] = (uint32_t)kv[i].idx;
    }
    offsets[nn] = m;
    free(kv);
    return nn;
}

/* ORBmatcher.cc:1045-1328, monocular pinhole path (mpCamera2 == NULL, NLeft == -1) */
int oracle_search_for_triangulation(const orbgpu_kfset_host *s, int kf1, int kf2, const float *ep, const float *f12,
                                    int only_stereo, int coarse, int check_ori, int32_t *matches12)
{
    const int n = s->n_feat;
    int nmatches = 0;
    uint32_t *nid1 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + 1)), *nid2 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + 1));
    int32_t *off1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 2)), *off2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 2));
    uint32_t *ft1 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + 1)), *ft2 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + 1));
    const int nn1 = kf_featvec(s, kf1, nid1, off1, ft1), nn2 = kf_featvec(s, kf2, nid2, off2, ft2);
    const uint8_t *desc1 = s->desc + (size_t)kf1 * n * 32, *desc2 = s->desc + (size_t)kf2 * n * 32;
    const float *xy1 = s->kp_xy + (size_t)kf1 * n * 2, *xy2 = s->kp_xy + (size_t)kf2 * n * 2;
    const int32_t *oct2 = s->octave + (size_t)kf2 * n;
    const float *ang1 = s->angle + (size_t)kf1 * n, *ang2 = s->angle + (size_t)kf2 * n;
    const uint8_t *mp1 = s->has_mp + (size_t)kf1 * n, *mp2 = s->has_mp + (size_t)kf2 * n;
    const float *ur1 = s->u_right ? s->u_right + (size_t)kf1 * n : NULL;
    const float *ur2 = s->u_right ? s->u_right + (size_t)kf2 * n : NULL;
    for (int i = 0; i < n; i++) matches12[i] = -1; /* :1092 */
    rot_hist rh;
    rh_init(&rh);
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) { /* :1113 */
        if (nid1[a] == nid2[b]) {
            for (int i1 = off1[a]; i1 < off1[a + 1]; i1++) {
                const int idx1 = (int)ft1[i1];
                if (mp1[idx1]) continue; /* :1125-1132 */
                const int bStereo1 = ur1 ? (ur1[idx1] >= 0) : 0; /* :1134 */
                if (only_stereo && !bStereo1) continue;
                const uint8_t *d1 = desc1 + 32 * (size_t)idx1;
                int bestDist = ORBGPU_TH_LOW, bestIdx2 = -1; /* :1151 */
                for (int i2 = off2[b]; i2 < off2[b + 1]; i2++) {
                    const int idx2 = (int)ft2[i2];
                    if (mp2[idx2]) continue; /* :1165 (vbMatched2 is never set: :1261-1262) */
                    const int bStereo2 = ur2 ? (ur2[idx2] >= 0) : 0;
                    if (only_stereo && !bStereo2) continue;
                    const int dist = oracle_descriptor_distance(d1, desc2 + 32 * (size_t)idx2);
                    if (dist > ORBGPU_TH_LOW || dist > bestDist) continue; /* :1180 */
                    if (!bStereo1 && !bStereo2) { /* :1191-1203 */
                        const float distex = ep[0] - xy2[2 * idx2];
                        const float distey = ep[1] - xy2[2 * idx2 + 1];
                        if (distex * distex + distey * distey < 100 * s->scale_factors[oct2[idx2]]) continue;
                    }
                    if (coarse || epipolar_constrain(f12, xy1[2 * idx1], xy1[2 * idx1 + 1], xy2[2 * idx2],
                                                     xy2[2 * idx2 + 1], s->level_sigma2[oct2[idx2]])) { /* :1246 */
                        bestIdx2 = idx2;
                        bestDist = dist;
                    }
                }
                if (bestIdx2 >= 0) { /* :1254-1278 */
                    matches12[idx1] = bestIdx2;
                    nmatches++;
                    if (check_ori) rh_push(&rh, rot_bin(ang1[idx1], ang2[bestIdx2]), idx1);
                }
            }
            a++; b++;
        } else if (nid1[a] < nid2[b]) {
            a++;
        } else {
            b++;
        }
    }
    if (check_ori) { /* :1295-1314 */
        int32_t ind[3];
        oracle_compute_three_maxima(rh.n, ORBGPU_HISTO_LENGTH, ind);
        for (int i = 0; i < ORBGPU_HISTO_LENGTH; i++) {
            if (i == ind[0] || i == ind[1] || i == ind[2]) continue;
            for (int j = 0; j < rh.n[i]; j++) {
                matches12[rh.v[i][j]] = -1;
                nmatches--;
            }
        }
    }
    rh_free(&rh);
    free(nid1); free(nid2); free(off1); free(off2); free(ft1); free(ft2);
    return nmatches;
}

typedef struct {
    const orbgpu_kfset_host *s;
    int p0, p1;
    const int32_t *kf1, *kf2;
    const float *ep, *f12;
    int only_stereo, coarse, check_ori;
    int32_t *matches12, *nmatches;
    int64_t comparisons;
} tri_job;

static void *tri_worker(void *arg)
{
    tri_job *j = (tri_job *)arg;
    oracle_comparisons_reset();
    for (int p = j->p0; p < j->p1; p++)
        j->nmatches[p] = oracle_search_for_triangulation(j->s, j->kf1[p], j->kf2[p], j->ep + 2 * (size_t)p,
                                                         j->f12 + 9 * (size_t)p, j->only_stereo, j->coarse,
                                                         j->check_ori, j->matches12 + (size_t)p * j->s->n_feat);
    j->comparisons = oracle_comparisons();
    return NULL;
}

void oracle_search_for_triangulation_batch(const orbgpu_kfset_host *s, int n_pairs, const int32_t *kf1,
                                           const int32_t *kf2, const float *ep, const float *f12, int only_stereo,
                                           int coarse, int check_ori, int32_t *matches12, int32_t *nmatches,
                                           int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    tri_job jobs[256];
    for (int t = 0; t < n_threads; t++) {
        tri_job *j = &jobs[t];
        j->s = s; j->kf1 = kf1; j->kf2 = kf2; j->ep = ep; j->f12 = f12;
        j->only_stereo = only_stereo; j->coarse = coarse; j->check_ori = check_ori;
        j->matches12 = matches12; j->nmatches = nmatches; j->comparisons = 0;
        j->p0 = (int)((int64_t)n_pairs * t / n_threads);
        j->p1 = (int)((int64_t)n_pairs * (t + 1) / n_threads);
        pthread_create(&th[t], NULL, tri_worker, j);
    }
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        g_comparisons += jobs[t].comparisons;
    }
}

typedef struct {
    int64_t q0, q1, nd;
    const uint8_t *q, *db;
    int th_low;
    float nnratio;
    int32_t *best_idx, *best_dist, *second_dist, *match;
    int64_t comparisons;
} knn_job;

/* semantics of the SearchByBoW inner loop, ORBmatcher.cc:319-355 + :392-395 */
static void *knn_worker(void *arg)
{
    knn_job *j = (knn_job *)arg;
    oracle_comparisons_reset();
    for (int64_t i = j->q0; i < j->q1; i++) {
        const uint8_t *dq = j->q + 32 * (size_t)i;
        int bestDist1 = 256, bestIdx = -1, bestDist2 = 256;
        for (int64_t d = 0; d < j->nd; d++) {
            const int dist = oracle_descriptor_distance(dq, j->db + 32 * (size_t)d);
            if (dist < bestDist1) {
                bestDist2 = bestDist1; bestDist1 = dist; bestIdx = (int)d;
            } else if (dist < bestDist2) {
                bestDist2 = dist;
            }
        }
        int m = -1;
        if (bestDist1 <= j->th_low)
            if ((float)bestDist1 < j->nnratio * (float)bestDist2) m = bestIdx;
        if (j->best_idx) j->best_idx[i] = bestIdx;
        if (j->best_dist) j->best_dist[i] = bestDist1;
        if (j->second_dist) j->second_dist[i] = bestDist2;
        if (j->match) j->match[i] = m;
    }
    j->comparisons = oracle_comparisons();
    return NULL;
}

void oracle_knn2_ratio(int64_t nq, const uint8_t *q, int64_t nd, const uint8_t *db, int th_low, float nnratio,
                       int32_t *best_idx, int32_t *best_dist, int32_t *second_dist, int32_t *match, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    knn_job jobs[256];
    for (int t = 0; t < n_threads; t++) {
        knn_job *j = &jobs[t];
        j->q0 = nq * t / n_threads; j->q1 = nq * (t + 1) / n_threads; j->nd = nd;
        j->q = q; j->db = db; j->th_low = th_low; j->nnratio = nnratio;
        j->best_idx = best_idx; j->best_dist = best_dist; j->second_dist = second_dist; j->match = match;
        j->comparisons = 0;
        pthread_create(&th[t], NULL, knn_worker, j);
    }
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        g_comparisons += jobs[t].comparisons;
    }
}
