/*
 * orb_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's descriptor-matching hot path
 * (Herong1212/ORB_SLAM3_comments_ghr: src/ORBmatcher.cc, src/Frame.cc, src/KeyFrame.cc,
 * src/CameraModels/Pinhole.cpp, Thirdparty/DBoW2).  Every function cites the reference
 * file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (liborbmatch_b200.so) never does.
 *
 * Parity pin: this restatement is checked against the reference's own ORBmatcher.cc / DBoW2
 * sources compiled unmodified into oracle/_ref/libref_orbmatcher.so (see oracle/Makefile,
 * tests/test_oracle_vs_ref.py) and against the golden vectors generated from that build
 * (tests/golden/, scripts/make_golden.py).
 *
 * Every function here is pinned on the reference's own compiled code: the matcher and DBoW2 sources compiled whole, the helper
 * functions of Frame.cc / KeyFrame.cc / MapPoint.cc / Pinhole.cpp (grid, GetFeaturesInArea, IsInImage, isInFrustum, PredictScale,
 * project, epipolarConstrain, the coarse stage of ComputeStereoMatches) from their own text cut out by line range at build time
 * (oracle/extract_ref.py), oracle_search_projected via the harnesses of ref_adapter.cc, oracle_bow_score_l1 (ScoringObject.cpp),
 * oracle_compute_distinctive_descriptors (src/MapPoint.cc compiled unmodified with its real include/MapPoint.h against shim_mp/).
 * What stays unpinned is only the fp32 rounding INSIDE the un-vendored Sophus / Eigen (quaternion-form SE3 * point, 3x3 inverse):
 * poses, epipoles and F12 are therefore inputs.
 *
 * The flat input structs are the ones of the product ABI (include/orbmatch_b200.h) so that
 * oracle and GPU consume byte-identical inputs.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H

#include "../include/orbmatch_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* number of DescriptorDistance calls made by this thread since the last reset */
int64_t oracle_comparisons(void);
void oracle_comparisons_reset(void);

/* ORBmatcher.cc:2388-2408 / FORB.cpp:92-112 */
int oracle_descriptor_distance(const uint8_t *a, const uint8_t *b);

/* Frame.cc:469-507 + :973-989.  cell = ix*rows+iy.  cell_start[cols*rows+1], cell_items[n]. */
void oracle_grid_build(const orbgpu_frame_host *f, int32_t *cell_start, int32_t *cell_items);

/* Frame.cc:868-962 (KeyFrame.cc:859-907 when min_level=-1,max_level=-1). returns count */
int oracle_features_in_area(const orbgpu_frame_host *f, const int32_t *cell_start, const int32_t *cell_items, float x,
                            float y, float r, int min_level, int max_level, int32_t *out_idx);

/* ORBmatcher.cc:2341-2383 */
void oracle_compute_three_maxima(const int32_t *histo_sizes, int L, int32_t *ind);

/* ORBmatcher.cc:735-878 */
int oracle_search_for_initialization(const orbgpu_frame_host *f1, const orbgpu_frame_host *f2, float *prev_matched_xy,
                                     int window_size, float nnratio, int check_ori, int32_t *matches12);

/* ORBmatcher.cc:44-242 (Nleft == -1 path) */
int oracle_search_by_projection_local(const orbgpu_frame_host *f, const orbgpu_mappoints_host *mps, float th,
                                      int far_points, float th_far_points, float nnratio, const int32_t *kp_prior_obs,
                                      int32_t *kp_mp);

/* search core of the self-projecting overloads (row a6: ORBmatcher.cc:498-733, 1330-1955, 1957-2330); returns nmatches */
int oracle_search_projected(const orbgpu_frame_host *f, const orbgpu_projpoints_host *pts, const orbgpu_projsearch_params *prm,
                            const uint8_t *kp_locked, int32_t *best_idx, int32_t *best_dist, int32_t *kp_owner);

/* KeyFrameDatabase.cc:928-943 (common words) + ScoringObject.cpp:23-68 (L1Scoring::score), query vs every key frame */
void oracle_bow_score_l1(const orbgpu_bowdb_host *db, int32_t nq, const uint32_t *q_words, const double *q_values,
                         int32_t *common_words, double *scores);

/* MapPoint.cc:444-535 batched over a CSR of observation descriptors (restatement only: MapPoint.cc does not compile here) */
void oracle_compute_distinctive_descriptors(int32_t n_mp, const int32_t *offsets, const uint8_t *desc, int32_t *best_idx,
                                            int32_t *best_median);

/* Frame.cc:1139-1216, coarse stage of ComputeStereoMatches (restatement only: Frame.cc does not compile here) */
void oracle_stereo_coarse_match(int32_t n_left, const uint8_t *desc_l, const float *kp_xy_l, const int32_t *octave_l, int32_t n_right,
                                const uint8_t *desc_r, const float *kp_xy_r, const int32_t *octave_r, const float *scale_factors,
                                int32_t n_rows, float mb, float mbf, int32_t *best_idx_r, int32_t *best_dist);

/* Pinhole.cpp:203-218 with F12 given (row-major); Pinhole.cpp:64-71; MapPoint.cc:695-738 */
int oracle_epipolar_constrain(const float *F12, float x1, float y1, float x2, float y2, float unc);
void oracle_pinhole_project(const float *K, const float *xyz, float *uv);
int oracle_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels);
/* Frame.cc:676-782 (Nleft == -1) */
void oracle_is_in_frustum(const orbgpu_frustum_host *fr, int32_t n, const float *world_pos, const float *normal,
                          const float *min_distance, const float *max_distance, uint8_t *in_view, float *proj_xy, float *proj_xr,
                          float *depth, int32_t *scale_level, float *view_cos);

/* TemplatedVocabulary.h:1216-1258 per feature */
void oracle_voc_transform(const orbgpu_voc_host *v, int32_t n, const uint8_t *desc, int levelsup, uint32_t *word_id,
                          uint32_t *node_id, double *weight);
/* TemplatedVocabulary.h:1157-1161,1192-1193 + BowVector.cpp:35-85: returns number of words */
int oracle_bowvector(int32_t n, const uint32_t *word_id, const double *weight, uint32_t *words, double *values);
/* TemplatedVocabulary.h:1157-1161 + FeatureVector.cpp:32-46: returns number of nodes */
int oracle_featvec(int32_t n, const uint32_t *node_id, const double *weight, uint32_t *node_ids, int32_t *offsets,
                   uint32_t *features);

/* ORBmatcher.cc:262-496 (Nleft == -1) */
int oracle_search_by_bow_kf_f(const orbgpu_frame_host *kf, const orbgpu_frame_host *f, const uint8_t *kf_mp_valid,
                              float nnratio, int check_ori, int32_t *match_f2kf);
/* ORBmatcher.cc:890-1043 */
int oracle_search_by_bow_kf_kf(const orbgpu_frame_host *kf1, const orbgpu_frame_host *kf2, const uint8_t *kf1_mp_valid,
                               const uint8_t *kf2_mp_valid, float nnratio, int check_ori, int32_t *match_12);

/* ORBmatcher.cc:1045-1328 for one pair of keyframes of a kfset (mono pinhole path);
 * Pinhole.cpp:203-218 for the epipolar test with the given F12. */
int oracle_search_for_triangulation(const orbgpu_kfset_host *s, int kf1, int kf2, const float *ep, const float *f12,
                                    int only_stereo, int coarse, int check_ori, int32_t *matches12);
/* all pairs, n_threads host threads (pairs are independent) */
void oracle_search_for_triangulation_batch(const orbgpu_kfset_host *s, int n_pairs, const int32_t *kf1,
                                           const int32_t *kf2, const float *ep, const float *f12, int only_stereo,
                                           int coarse, int check_ori, int32_t *matches12, int32_t *nmatches,
                                           int n_threads);

/* brute-force 2-NN + ratio (semantics: ORBmatcher.cc:319-355, 392-395); n_threads host threads */
void oracle_knn2_ratio(int64_t nq, const uint8_t *q, int64_t nd, const uint8_t *db, int th_low, float nnratio,
                       int32_t *best_idx, int32_t *best_dist, int32_t *second_dist, int32_t *match, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
