/*
 * orbmatch_b200.h -- C ABI of the B200-native ORB descriptor-matching hot path.
 *
 * This is the drop-in boundary for ORB-SLAM3's ORBmatcher family and the DBoW2
 * vocabulary descent (reference: Herong1212/ORB_SLAM3_comments_ghr).  The
 * reference has no FFI layer; its boundary is the C++ class ORB_SLAM3::ORBmatcher
 * (include/ORBmatcher.h:34-99) and ORBVocabulary::transform
 * (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:146-147).  Every entry point below
 * names the reference function it replaces.  The C++ adapter with the reference's
 * exact signatures lives in include/orbmatch_b200/ORBmatcher.hpp and only packs /
 * unpacks STL containers around these calls.
 *
 * Conventions
 *   - extern "C", POD arguments only, no CUDA or torch types in the signatures.
 *   - every function returns an int status: ORBGPU_OK (0) or a negative error code;
 *     orbgpu_last_error() returns a human readable message for the calling thread.
 *   - "host" pointers are ordinary CPU memory (pinned memory makes copies async).
 *     "_dev" entry points take CUDA device pointers (for callers that keep data
 *     resident in HBM, e.g. torch tensors via data_ptr()).
 *   - descriptors are N x 32 bytes row-major (cv::Mat CV_8U rows, Frame.h:247); on the
 *     device they live as packed uint4 pairs (two 128-bit words per descriptor).
 *   - all index outputs are int32; "no match" is -1 (reference: NULL MapPoint* / -1).
 *   - there is NO CPU fallback: every search runs on the GPU or fails with an error.
 *   - a context owns one CUDA stream and its workspaces; contexts are independent so
 *     the three reference threads (Tracking / LocalMapping / LoopClosing,
 *     System.cc:234,254) each use their own.  A context is not thread-safe itself.
 */
#ifndef ORBMATCH_B200_H
#define ORBMATCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBGPU_OK 0
#define ORBGPU_ERR_INVALID (-1)   /* bad argument */
#define ORBGPU_ERR_CUDA (-2)      /* CUDA runtime error (message in orbgpu_last_error) */
#define ORBGPU_ERR_NO_DEVICE (-3) /* no usable sm_100 GPU: there is no CPU fallback */
#define ORBGPU_ERR_OVERFLOW (-4)  /* internal capacity exceeded even after regrow */

/* ORBmatcher.cc:34-36 */
#define ORBGPU_TH_HIGH 100
#define ORBGPU_TH_LOW 50
#define ORBGPU_HISTO_LENGTH 30
/* Frame.h:44-45 */
#define ORBGPU_FRAME_GRID_ROWS 48
#define ORBGPU_FRAME_GRID_COLS 64

typedef struct orbgpu_ctx orbgpu_ctx;
typedef struct orbgpu_frame orbgpu_frame;   /* device copy of one Frame/KeyFrame feature set */
typedef struct orbgpu_voc orbgpu_voc;       /* device copy of a DBoW2 vocabulary tree */
typedef struct orbgpu_kfset orbgpu_kfset;   /* device-resident batch of keyframes (config C4) */
typedef struct orbgpu_db orbgpu_db;         /* device-resident descriptor database (config C5) */

/* Host-side view of the Frame / KeyFrame members the matcher reads.
 * Frame.h:219-296, KeyFrame.h:378-406. */
typedef struct orbgpu_frame_host {
    int32_t n;              /* Frame::N */
    const uint8_t *desc;    /* [n][32]  mDescriptors */
    const float *kp_xy;     /* [n][2]   mvKeysUn[i].pt */
    const int32_t *octave;  /* [n]      mvKeysUn[i].octave */
    const float *angle;     /* [n]      mvKeysUn[i].angle */
    const float *u_right;   /* [n] mvuRight, or NULL == monocular (all -1) */
    float min_x, min_y, max_x, max_y; /* mnMinX.. (KeyFrame stores ints: pass them converted) */
    float grid_inv_w, grid_inv_h;     /* mfGridElementWidthInv / HeightInv */
    int32_t grid_cols, grid_rows;     /* 64 x 48 */
    int32_t n_levels;                 /* mnScaleLevels (<= 16) */
    const float *scale_factors;       /* [n_levels] mvScaleFactors */
    const float *level_sigma2;        /* [n_levels] mvLevelSigma2 */
    /* optional FeatureVector in CSR form (flattened std::map<NodeId, vector<uint>>,
     * FeatureVector.h): node ids ascending, feature ids ascending inside a node.
     * fv_n_nodes == 0 means "no FeatureVector yet" (orbgpu_transform can fill it). */
    int32_t fv_n_nodes;
    const uint32_t *fv_node_ids;   /* [fv_n_nodes] */
    const int32_t *fv_offsets;     /* [fv_n_nodes+1] */
    const uint32_t *fv_features;   /* [fv_offsets[fv_n_nodes]] */
} orbgpu_frame_host;

/* Host-side view of a DBoW2 vocabulary (TemplatedVocabulary.h:361-435: m_k, m_L, m_nodes).
 * Node 0 is the root.  children of node i are child_ids[child_offsets[i] .. child_offsets[i+1]),
 * in the order of Node::children.  A node without children is a leaf (Node::isLeaf). */
typedef struct orbgpu_voc_host {
    int32_t k, L;
    int32_t n_nodes;
    const uint8_t *node_desc;       /* [n_nodes][32] Node::descriptor (root row unused) */
    const int32_t *child_offsets;   /* [n_nodes+1] */
    const uint32_t *child_ids;      /* [child_offsets[n_nodes]] */
    const double *weight;           /* [n_nodes] Node::weight */
    const uint32_t *word_id;        /* [n_nodes] Node::word_id (valid on leaves) */
} orbgpu_voc_host;

/* ---- context ------------------------------------------------------------------------- */
int orbgpu_device_count(void);
int orbgpu_create(int device, orbgpu_ctx **out);
int orbgpu_create_on_stream(int device, void *cuda_stream, orbgpu_ctx **out);
void orbgpu_destroy(orbgpu_ctx *ctx);
int orbgpu_synchronize(orbgpu_ctx *ctx);
void *orbgpu_stream(orbgpu_ctx *ctx); /* cudaStream_t the context launches on */
const char *orbgpu_last_error(void);
const char *orbgpu_version(void);
/* number of this library's kernels launched through ctx since creation (bench gpu_launches) */
int64_t orbgpu_launch_count(orbgpu_ctx *ctx);
/* number of Hamming comparisons (DescriptorDistance-equivalents) the last search call
 * executed, counted on the device exactly where the reference calls DescriptorDistance. */
int64_t orbgpu_last_comparisons(orbgpu_ctx *ctx);
/* same, after synchronising the stream (the *_dev entry points return without synchronising) */
int orbgpu_fetch_comparisons(orbgpu_ctx *ctx, int64_t *out);

/* ---- a1: ORBmatcher::DescriptorDistance (ORBmatcher.cc:2388-2408), FORB::distance (FORB.cpp:92-112)
 * batched: out[i] = hamming(a[i], b[i]).  Host pointers. */
int orbgpu_descriptor_distance(orbgpu_ctx *ctx, int64_t n, const uint8_t *a, const uint8_t *b, int32_t *out);

/* ---- a2/a3: Frame::AssignFeaturesToGrid (Frame.cc:469-507), PosInGrid (:973-989),
 * Frame::GetFeaturesInArea (:868-962), KeyFrame::GetFeaturesInArea (KeyFrame.cc:859-907).
 * Upload builds the device CSR cell index (cell = ix*rows+iy, in-cell ascending feature id). */
int orbgpu_frame_upload(orbgpu_ctx *ctx, const orbgpu_frame_host *f, orbgpu_frame **out);
void orbgpu_frame_destroy(orbgpu_frame *f);
int orbgpu_frame_n(const orbgpu_frame *f);
/* copy the device CSR grid back: cell_start[cols*rows+1], cell_items[n] (n_in_grid <= n valid) */
int orbgpu_frame_grid_download(orbgpu_ctx *ctx, const orbgpu_frame *f, int32_t *cell_start, int32_t *cell_items);
/* batched GetFeaturesInArea: for query q the indices in reference order are
 * out_idx[out_offsets[q] .. out_offsets[q+1]).  min_level/max_level as in Frame.cc:868 (pass -1,-1
 * for the KeyFrame variant).  out_idx capacity is cap entries; total is returned in *total. */
int orbgpu_features_in_area(orbgpu_ctx *ctx, const orbgpu_frame *f, int32_t nq, const float *x, const float *y,
                            const float *r, const int32_t *min_level, const int32_t *max_level,
                            int32_t *out_offsets, int32_t *out_idx, int64_t cap, int64_t *total);

/* ---- a4: ORBmatcher::SearchForInitialization (ORBmatcher.cc:735-878)
 * prev_matched_xy [n1][2] in/out (vbPrevMatched); matches12 [n1] out (vnMatches12). */
int orbgpu_search_for_initialization(orbgpu_ctx *ctx, const orbgpu_frame *f1, const orbgpu_frame *f2,
                                     float *prev_matched_xy, int32_t window_size, float nnratio, int32_t check_ori,
                                     int32_t *matches12, int32_t *nmatches);

/* ---- a5: ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints)
 * (ORBmatcher.cc:44-242, monocular/RGB-D path Nleft==-1) + RadiusByViewingCos (:245-252).
 * Map-point inputs are the members the function reads (MapPoint.h:166-177, 209-239). */
typedef struct orbgpu_mappoints_host {
    int32_t n;
    const uint8_t *desc;          /* [n][32] MapPoint::GetDescriptor() */
    const float *proj_xy;         /* [n][2]  mTrackProjX, mTrackProjY */
    const float *proj_xr;         /* [n]     mTrackProjXR (stereo gate), or NULL */
    const int32_t *scale_level;   /* [n]     mnTrackScaleLevel */
    const float *view_cos;        /* [n]     mTrackViewCos */
    const float *depth;           /* [n]     mTrackDepth */
    const uint8_t *in_view;       /* [n]     mbTrackInView */
    const uint8_t *bad;           /* [n]     isBad() */
    const int32_t *n_obs;         /* [n]     Observations() */
} orbgpu_mappoints_host;
/* kp_mp [F.N] in/out: index of the map point (into mps) assigned to each keypoint, -1 none
 * (F.mvpMapPoints).  kp_prior_obs [F.N] in: Observations() of the map point a keypoint
 * already holds on entry (0 when none) -- used by the skip rule at :102-104. */
int orbgpu_search_by_projection_local(orbgpu_ctx *ctx, const orbgpu_frame *f, const orbgpu_mappoints_host *mps,
                                      float th, int32_t far_points, float th_far_points, float nnratio,
                                      const int32_t *kp_prior_obs, int32_t *kp_mp, int32_t *nmatches);

/* ---- 8(f) rank 1: Frame::isInFrustum (Frame.cc:676-782, Nleft == -1) for a whole list of map points, on the device, and the fused
 * Tracking::SearchLocalPoints step: isInFrustum + SearchByProjection(Frame&, vector<MapPoint*>&) without the host loop between them.
 * The frame's pose members are inputs exactly as the reference holds them (mRcw, mtcw, mOw: Frame.h, filled by
 * Frame::UpdatePoseMatrices with the host's own Sophus); everything from `mRcw * P + mtcw` on is evaluated on the device in the
 * reference's fp32 operation order.  PredictScale (MapPoint.cc:695-738) takes the logarithm in double precision and rounds it to
 * float: it can differ from the host libm's logf by one ulp, which moves the predicted level only when log(ratio) / logScaleFactor
 * lies within one ulp of an integer -- the C++ adapter's checker mode counts such disagreements (north_star target: zero). */
typedef struct orbgpu_frustum_host {
    float Rcw[9];            /* Frame::mRcw, row-major */
    float tcw[3];            /* Frame::mtcw */
    float Ow[3];             /* Frame::mOw */
    float K[4];              /* pinhole fx fy cx cy (mpCamera->project, Pinhole.cpp:64-71) */
    float mbf;               /* Frame::mbf (mTrackProjXR = u - mbf / z) */
    float min_x, min_y, max_x, max_y; /* Frame::mnMinX .. mnMaxY */
    float viewing_cos_limit; /* 0.5 in Tracking::SearchLocalPoints */
    float log_scale_factor;  /* Frame::mfLogScaleFactor */
    int32_t n_levels;        /* Frame::mnScaleLevels */
} orbgpu_frustum_host;
/* world_pos [n][3], normal [n][3], min_distance / max_distance [n] = MapPoint::mfMinDistance / mfMaxDistance (the 0.8 / 1.2
 * invariance factors of MapPoint.cc:665-678 are applied inside).  Outputs [n] (each may be NULL): the members isInFrustum writes --
 * in_view = mbTrackInView, proj_xy = mTrackProjX/Y (-1 when rejected before the image-bounds test passed), proj_xr, depth,
 * scale_level, view_cos (0 when the point is rejected). */
int orbgpu_is_in_frustum(orbgpu_ctx *ctx, const orbgpu_frustum_host *fr, int32_t n, const float *world_pos, const float *normal,
                         const float *min_distance, const float *max_distance, uint8_t *in_view, float *proj_xy, float *proj_xr,
                         float *depth, int32_t *scale_level, float *view_cos);
/* Tracking::SearchLocalPoints (Tracking.cc: isInFrustum over mvpLocalMapPoints, then SearchByProjection): map points as the
 * reference's loop sees them.  skip [n] = 1 for points the loop does not test (already matched in the current frame, isBad): they
 * keep mbTrackInView = false.  in_view [n] out (may be NULL) tells the caller which points to IncreaseVisible(); the rest is
 * orbgpu_search_by_projection_local on the device-resident projections. */
typedef struct orbgpu_localpoints_host {
    int32_t n;
    const uint8_t *desc;         /* [n][32] MapPoint::GetDescriptor() */
    const float *world_pos;      /* [n][3] */
    const float *normal;         /* [n][3] */
    const float *min_distance;   /* [n] mfMinDistance */
    const float *max_distance;   /* [n] mfMaxDistance */
    const uint8_t *skip;         /* [n] or NULL */
    const uint8_t *bad;          /* [n] isBad() at matching time (ORBmatcher.cc:62) */
    const int32_t *n_obs;        /* [n] Observations() */
} orbgpu_localpoints_host;
int orbgpu_search_local_points(orbgpu_ctx *ctx, const orbgpu_frame *f, const orbgpu_frustum_host *fr, const orbgpu_localpoints_host *pts,
                               float th, int32_t far_points, float th_far_points, float nnratio, const int32_t *kp_prior_obs,
                               int32_t *kp_mp, uint8_t *in_view, int32_t *nmatches);

/* ---- a6: the projection-gated searches that do their own projection: SearchByProjection(Cur, Last) (ORBmatcher.cc:1957-2191),
 * SearchByProjection(Cur, KF, sAlreadyFound) (:2203-2330), SearchByProjection(KF, Sim3, ...) x2 (:498-733),
 * Fuse x2 (:1330-1682), SearchBySim3 (:1684-1955).  They share one skeleton after the per-point prologue
 * (SE3/Sim3 transform, camera projection, frustum / distance / viewing-angle gates, PredictScale): a window search
 * around (u, v) with an octave filter, optional per-candidate gates, best-only selection (strict <, first
 * candidate in GetFeaturesInArea order wins) and an acceptance threshold.  The prologue is evaluated by the caller
 * with its own Sophus / Eigen (include/orbmatch_b200/ORBmatcher.hpp), because that arithmetic is not vendored in the
 * reference; everything from the window on runs here. */
typedef struct orbgpu_projpoints_host {
    int32_t n;
    const uint8_t *desc;       /* [n][32] MapPoint::GetDescriptor() */
    const float *uv;           /* [n][2]  projection into the target frame */
    const float *radius;       /* [n]     th * mvScaleFactors[level] */
    const int32_t *min_level;  /* [n]     minLevel of Frame::GetFeaturesInArea (Frame.cc:868, :919); the KeyFrame callers' */
    const int32_t *max_level;  /* [n]     `kpLevel < nPredictedLevel-1 || kpLevel > nPredictedLevel` is (level-1, level) */
    const float *ur;           /* [n] or NULL: predicted right-image coordinate uv(0) - mbf*invz (stereo / chi2 gates) */
    const uint8_t *active;     /* [n] 0: dropped by the caller's prologue gates */
    const uint8_t *locks;      /* [n] or NULL (all 1): once this point takes a keypoint, later points skip that keypoint */
    const float *angle;        /* [n] or NULL: angle of the source keypoint (rotation histogram) */
} orbgpu_projpoints_host;
typedef struct orbgpu_projsearch_params {
    float max_dist;       /* accept iff (float)bestDist <= max_dist (TH_HIGH, ORBdist, TH_LOW*ratioHamming, TH_LOW) */
    int32_t ordered;      /* 1: points in order, a keypoint that is locked (kp_locked on entry, or taken by an earlier locking
                             point) is skipped (:2046-2049, :580-581); 0: every point independently (Fuse, SearchBySim3) */
    int32_t stereo_gate;  /* 1: skip a candidate with u_right > 0 and |ur - u_right| > radius (:2052-2059) */
    int32_t chi2_gate;    /* 1: Fuse reprojection gate (:1463-1492): e2*invSigma2[level] > 7.8 (u_right >= 0) / 5.99 */
    int32_t check_ori;    /* 1: rotation-histogram cull over the accepted matches (:2163-2186) */
    const float *inv_level_sigma2; /* [frame n_levels] mvInvLevelSigma2, needed when chi2_gate */
} orbgpu_projsearch_params;
/* kp_locked [frame n] in (may be NULL): keypoints unavailable on entry.
 * best_idx / best_dist [n] out: accepted keypoint of every point (-1 none) BEFORE the histogram cull, and its distance.
 * kp_owner [frame n] out (may be NULL): index of the LAST point that took each keypoint, -1 when none or culled
 * (what the reference leaves in CurrentFrame.mvpMapPoints / vpMatched).  nmatches: accepted minus culled entries. */
int orbgpu_search_projected(orbgpu_ctx *ctx, const orbgpu_frame *f, const orbgpu_projpoints_host *pts,
                            const orbgpu_projsearch_params *prm, const uint8_t *kp_locked, int32_t *best_idx,
                            int32_t *best_dist, int32_t *kp_owner, int32_t *nmatches);

/* ---- a11/a12: TemplatedVocabulary::transform (TemplatedVocabulary.h:1127-1194, 1216-1258) */
int orbgpu_voc_upload(orbgpu_ctx *ctx, const orbgpu_voc_host *v, orbgpu_voc **out);
void orbgpu_voc_destroy(orbgpu_voc *v);
/* Descends every feature of frame f.  Outputs (host, each [n], any may be NULL):
 *   word_id, node_id (node at level L-levelsup, 0 == root when L-levelsup<=0), weight.
 * When store_featvec != 0 the frame's device FeatureVector (CSR by node id, stopped
 * words w==0 dropped, TemplatedVocabulary.h:1157) is (re)built on the device so that
 * SearchByBoW / SearchForTriangulation can run without a host round trip. */
int orbgpu_transform(orbgpu_ctx *ctx, const orbgpu_voc *voc, orbgpu_frame *f, int32_t levelsup, int32_t store_featvec,
                     uint32_t *word_id, uint32_t *node_id, double *weight);
/* The same transform for a bare list of descriptors with the reference's outputs in flat form, one call and one synchronisation
 * (what Frame::ComputeBoW, Frame.cc:1005-1008, and KeyFrame::ComputeBoW, KeyFrame.cc:113, need): BowVector as (words ascending,
 * values) with n_words entries, FeatureVector as CSR (node ids ascending, offsets, feature ids ascending inside a node) with
 * n_nodes nodes.  Capacities: n entries each (n + 1 offsets). */
int orbgpu_transform_descriptors(orbgpu_ctx *ctx, const orbgpu_voc *voc, int32_t n, const uint8_t *desc, int32_t levelsup,
                                 int32_t *n_words, uint32_t *words, double *values, int32_t *n_nodes, uint32_t *node_ids,
                                 int32_t *offsets, uint32_t *features);
/* device-built BowVector (BowVector.cpp:35-85): words ascending, value = idf added count
 * times in feature order, then L1-normalised in ascending word order.  Capacity n each. */
int orbgpu_bowvector_download(orbgpu_ctx *ctx, const orbgpu_frame *f, int32_t *n_words, uint32_t *words, double *values);
/* device FeatureVector back to host CSR (capacities: n nodes, n+1 offsets, n features) */
int orbgpu_featvec_download(orbgpu_ctx *ctx, const orbgpu_frame *f, int32_t *n_nodes, uint32_t *node_ids,
                            int32_t *offsets, uint32_t *features);

/* ---- a7: ORBmatcher::SearchByBoW (KeyFrame*, Frame&, vector<MapPoint*>&) (ORBmatcher.cc:262-496)
 * kf_mp_valid [kf.N]: 1 when vpMapPointsKF[i] != NULL && !isBad().
 * match_f2kf [f.N] out: index of the KF feature whose MapPoint was assigned to F feature i
 * (vpMapPointMatches[i] = vpMapPointsKF[match]), -1 none. */
int orbgpu_search_by_bow_kf_f(orbgpu_ctx *ctx, const orbgpu_frame *kf, const orbgpu_frame *f, const uint8_t *kf_mp_valid,
                              float nnratio, int32_t check_ori, int32_t *match_f2kf, int32_t *nmatches);
/* One frame against n_kf candidate key frames -- the relocalisation loop (Tracking.cc:4469-4495: SearchByBoW(vpCandidateKFs[i],
 * mCurrentFrame, vvpMapPointMatches[i]) for every candidate) and the loop / merge candidate loops of LoopClosing -- in ONE call: one
 * upload of the validity masks, the searches enqueued back to back, one download, one synchronisation.  match_f2kf [n_kf][f.N],
 * nmatches [n_kf]; results identical to n_kf separate calls. */
int orbgpu_search_by_bow_kf_f_batch(orbgpu_ctx *ctx, int32_t n_kf, const orbgpu_frame *const *kfs, const orbgpu_frame *f,
                                    const uint8_t *const *kf_mp_valid, float nnratio, int32_t check_ori, int32_t *match_f2kf,
                                    int32_t *nmatches);
/* KeyFrame <-> KeyFrame variant (ORBmatcher.cc:890-1043): strict best < TH_LOW, vbMatched2.
 * match_12 [kf1.N] out: index of the KF2 feature matched to KF1 feature i, -1 none. */
int orbgpu_search_by_bow_kf_kf(orbgpu_ctx *ctx, const orbgpu_frame *kf1, const orbgpu_frame *kf2,
                               const uint8_t *kf1_mp_valid, const uint8_t *kf2_mp_valid, float nnratio, int32_t check_ori,
                               int32_t *match_12, int32_t *nmatches);
/* The current key frame against the n_kf key frames of a candidate's covisibility window -- the loop / merge detection loop
 * (LoopClosing.cc:909-925: matcherBoW.SearchByBoW(mpCurrentKF, vpCovKFi[j], vvpMatchedMPs[j]) for every j) -- in ONE call, one launch
 * over (node, window key frame).  match_12 [n_kf][kf1.N], nmatches [n_kf]; results identical to n_kf separate calls. */
int orbgpu_search_by_bow_kf_kf_batch(orbgpu_ctx *ctx, const orbgpu_frame *kf1, const uint8_t *kf1_mp_valid, int32_t n_kf,
                                     const orbgpu_frame *const *kf2s, const uint8_t *const *kf2_mp_valid, float nnratio,
                                     int32_t check_ori, int32_t *match_12, int32_t *nmatches);

/* ---- a8/a9: batched ORBmatcher::SearchForTriangulation (ORBmatcher.cc:1045-1328) with
 * Pinhole::epipolarConstrain (Pinhole.cpp:189-219), monocular pinhole path.
 * A kfset holds n_kf keyframes of n_feat features each, resident in HBM. */
typedef struct orbgpu_kfset_host {
    int32_t n_kf, n_feat;
    const uint8_t *desc;         /* [n_kf][n_feat][32] */
    const float *kp_xy;          /* [n_kf][n_feat][2] */
    const int32_t *octave;       /* [n_kf][n_feat] */
    const float *angle;          /* [n_kf][n_feat] */
    const uint8_t *has_mp;       /* [n_kf][n_feat] GetMapPoint(i) != NULL */
    const float *u_right;        /* [n_kf][n_feat] mvuRight or NULL (mono) */
    const uint32_t *node_id;     /* [n_kf][n_feat] FeatureVector node of each feature, 0xFFFFFFFF == none; NULL: see orbgpu_kfset_transform */
    int32_t n_levels;
    const float *scale_factors;  /* [n_levels] */
    const float *level_sigma2;   /* [n_levels] */
} orbgpu_kfset_host;
int orbgpu_kfset_upload(orbgpu_ctx *ctx, const orbgpu_kfset_host *s, orbgpu_kfset **out);
void orbgpu_kfset_destroy(orbgpu_kfset *s);
/* KeyFrame::ComputeBoW (KeyFrame.cc:102-117) for the whole set on the device: TemplatedVocabulary::transform of every feature
 * (TemplatedVocabulary.h:1127-1194, 1216-1258) -> FeatureVector node ids (level L - levelsup, stopped words dropped), then the
 * per-key-frame CSR is rebuilt.  A set uploaded with node_id == NULL has no FeatureVector until this is called. */
int orbgpu_kfset_transform(orbgpu_ctx *ctx, const orbgpu_voc *voc, orbgpu_kfset *s, int32_t levelsup);
/* Per pair p: keyframes (kf1[p], kf2[p]); geometry precomputed by the caller with the
 * reference's own host algebra (ORBmatcher.cc:1053-1071, Pinhole.cpp:194-197):
 *   ep[p][2]  epipole of camera 1 in image 2;  f12[p][9] row-major F12 = K1^-T [t12]x R12 K2^-1.
 * matches12 [n_pairs][n_feat] out (vMatches12, -1 none), nmatches [n_pairs] out.
 * Pointers are HOST pointers; the _dev variant takes device pointers and does no copies. */
int orbgpu_search_for_triangulation_batch(orbgpu_ctx *ctx, const orbgpu_kfset *s, int32_t n_pairs, const int32_t *kf1,
                                          const int32_t *kf2, const float *ep, const float *f12, int32_t only_stereo,
                                          int32_t coarse, int32_t check_ori, int32_t *matches12, int32_t *nmatches);
int orbgpu_search_for_triangulation_batch_dev(orbgpu_ctx *ctx, const orbgpu_kfset *s, int32_t n_pairs, const int32_t *kf1_dev,
                                              const int32_t *kf2_dev, const float *ep_dev, const float *f12_dev,
                                              int32_t only_stereo, int32_t coarse, int32_t check_ori, int32_t *matches12_dev,
                                              int32_t *nmatches_dev);

/* Fused search + all-gather for the sharded batch (monocular sets): the rows / counts of this rank's pairs are stored into the
 * [P_total][n_feat] / [P_total] result buffers of ALL ranks -- target_matches[r], target_nmatches[r] are device addresses valid
 * on THIS GPU (peer memory over NVLink / NVSwitch, e.g. torch symmetric memory) -- at rows pair_offset + p.
 *   rows_preset 0: every target row is filled with -1 and receives the individual match stores;
 *   rows_preset 1: every owner filled its buffer with -1 beforehand, only matches and counts cross the links;
 *   rows_preset 2: target 0 must be THIS rank's buffer; the row is built there and then copied whole to the other targets
 *                  with coalesced 128-bit stores (no preset; small scattered stores do not coalesce on the links).
 * The caller separates steps with a cross-rank barrier.  n_targets <= 8 (one box). */
int orbgpu_search_for_triangulation_batch_peers_dev(orbgpu_ctx *ctx, const orbgpu_kfset *s, int32_t n_pairs, const int32_t *kf1_dev,
                                                    const int32_t *kf2_dev, const float *ep_dev, const float *f12_dev,
                                                    int32_t coarse, int32_t check_ori, int32_t n_targets, void *const *target_matches,
                                                    void *const *target_nmatches, int64_t pair_offset, int32_t rows_preset);
/* Fused search + all-gather in the reference's own result form, vMatchedPairs (ORBmatcher.cc:1317-1325): the matches of pair p,
 * ascending idx1, as 32-bit entries (idx1 << 16 | idx2) at pairs[(pair_offset + p) * n_feat ...], counts[pair_offset + p] of
 * them valid (the store granularity is 16 bytes: up to 3 entries 0xFFFFFFFF may follow).  Only the valid prefix of every pair
 * crosses NVLink -- about 1/10 of the dense rows of the _peers_dev variant -- and it is stored from inside the search kernel,
 * straight from shared memory, into the buffers of ALL ranks while later pairs are still being compared.  There is no
 * cross-rank barrier: when a rank has shipped its last pair it publishes its epoch in slot `rank` of every rank's flag array
 * (system-scope fences order the data before it); the CTA that completes the rank's batch then waits until every source in
 * wait_mask has published the current epoch and advances epoch_done[0] -- the kernel ends when the gathered result is complete
 * on this rank, one launch per step.  Buffers are peer-mapped device addresses valid on
 * THIS GPU (e.g. torch symmetric memory); the caller double-buffers pairs / counts between consecutive steps (a rank may be one
 * step ahead of its peers), flags / epoch_done / status are shared by both buffers.  n_feat must be a multiple of 4.
 * Initial state: flags = 0, epoch_done = {1, 0}, status = {0}.  status[0] becomes 1 if a source did not arrive within 4 s. */
typedef struct orbgpu_tri_gather {
    int32_t n_ranks, rank;
    void *pairs[8];    /* [n_ranks] -> uint32 [P_total][n_feat] */
    void *counts[8];   /* [n_ranks] -> int32  [P_total] */
    void *flags[8];    /* [n_ranks] -> uint32 [n_ranks]: flags[t][s] = last epoch rank s has completed into rank t's buffers */
    void *epoch_done;  /* local uint32 [2] */
    void *status;      /* local uint32 [1], may be NULL */
    uint32_t wait_mask; /* sources to wait for (bit s); (1 << n_ranks) - 1 in normal use */
} orbgpu_tri_gather;
int orbgpu_search_for_triangulation_batch_gather_dev(orbgpu_ctx *ctx, const orbgpu_kfset *s, int32_t n_pairs, const int32_t *kf1_dev,
                                                     const int32_t *kf2_dev, const float *ep_dev, const float *f12_dev, int32_t coarse,
                                                     int32_t check_ori, const orbgpu_tri_gather *g, int64_t pair_offset);
/* The gathered result on the host, in the form of orbgpu_search_for_triangulation_batch_pairs below (pair_offsets [n_pairs + 1],
 * pairs [total][2] = (idx1, idx2), ascending idx1 inside a pair): offsets scan and packing on the device, then only the valid
 * pairs cross PCIe.  counts_dev / entries_dev are this rank's gather buffers (device pointers); the call runs on ctx's stream, so
 * it is ordered after a gather step launched on the same context. */
int orbgpu_tri_gather_download(orbgpu_ctx *ctx, int32_t n_pairs, int32_t n_feat, const void *counts_dev, const void *entries_dev,
                               int32_t *pair_offsets, int32_t *pairs, int64_t cap, int64_t *total);
/* Same search with the result in the reference's vMatchedPairs form (ORBmatcher.cc:1317-1325): for pair p the matches are
 * pairs[2*j], pairs[2*j+1] = (idx1, idx2), j in [pair_offsets[p], pair_offsets[p+1]), ascending idx1.  pair_offsets has
 * n_pairs+1 entries; cap = capacity of pairs in (idx1, idx2) entries; *total = entries produced (ORBGPU_ERR_OVERFLOW and the
 * first cap entries when it exceeds cap).  Only the pairs cross the bus: about 1/20 of the dense rows. */
int orbgpu_search_for_triangulation_batch_pairs(orbgpu_ctx *ctx, const orbgpu_kfset *s, int32_t n_pairs, const int32_t *kf1,
                                                const int32_t *kf2, const float *ep, const float *f12, int32_t only_stereo,
                                                int32_t coarse, int32_t check_ori, int32_t *pair_offsets, int32_t *pairs, int64_t cap,
                                                int64_t *total);
/* selects the batched-triangulation kernel: 0 = auto, 1 = one CTA per pair (cp.async staging; the only one
 * that handles mvuRight / bOnlyStereo), 2 = persistent CTAs, producer warp + double-buffered bulk copies
 * (cp.async.bulk + mbarrier) of per-keyframe stream blobs (what auto resolves to for monocular sets). */
int orbgpu_triangulation_set_engine(orbgpu_ctx *ctx, int32_t engine);

/* ---- brute-force 2-NN + ratio test (north_star "SearchByNN"; not in this fork: semantics are the
 * inner loop of SearchByBoW, ORBmatcher.cc:327-355 + accept rule :392-395).
 * For each query: best = first db index attaining the minimum distance, second = second
 * smallest distance of the multiset (256 when nd < 2).  match = best_idx when
 * best <= th_low && (float)best < nnratio*(float)second, else -1. */
int orbgpu_db_upload(orbgpu_ctx *ctx, int64_t nd, const uint8_t *db_desc, orbgpu_db **out);
/* refresh an uploaded database in place (nd <= the size it was created with): one H2D copy, no allocation */
int orbgpu_db_update(orbgpu_ctx *ctx, orbgpu_db *db, int64_t nd, const uint8_t *db_desc);
int orbgpu_db_from_dev(orbgpu_ctx *ctx, int64_t nd, const void *db_desc_dev, orbgpu_db **out); /* borrows the pointer */
/* The tensor engine keeps an expanded (+-1 fp8) copy of the database inside the orbgpu_db, built by the first search and reused by
 * the following ones (the map changes at key-frame rate, relocalisation queries it per frame).  orbgpu_db_update drops it; for a
 * borrowed database (orbgpu_db_from_dev) the caller reports changed contents with orbgpu_db_invalidate. */
int orbgpu_db_invalidate(orbgpu_db *db);
void orbgpu_db_destroy(orbgpu_db *db);
int orbgpu_knn2_ratio(orbgpu_ctx *ctx, const orbgpu_db *db, int64_t nq, const uint8_t *q_desc, int32_t th_low, float nnratio,
                      int32_t *best_idx, int32_t *best_dist, int32_t *second_dist, int32_t *match);
/* orbgpu_db_update + orbgpu_knn2_ratio in one call with the upload overlapped: the descriptors are copied in chunks on the database's
 * own stream and every chunk is searched as soon as it has arrived (SearchByNN(queries, database) on two host matrices).  db must be
 * an owned database (orbgpu_db_upload) with capacity >= nd.  Results identical to the two separate calls. */
int orbgpu_knn2_ratio_update(orbgpu_ctx *ctx, orbgpu_db *db, int64_t nd, const uint8_t *db_desc, int64_t nq, const uint8_t *q_desc,
                             int32_t th_low, float nnratio, int32_t *best_idx, int32_t *best_dist, int32_t *second_dist, int32_t *match);
int orbgpu_knn2_ratio_dev(orbgpu_ctx *ctx, const orbgpu_db *db, int64_t nq, const void *q_desc_dev, int32_t th_low,
                          float nnratio, int32_t *best_idx_dev, int32_t *best_dist_dev, int32_t *second_dist_dev,
                          int32_t *match_dev);
/* selects the all-pairs engine: 0 = auto, 1 = LOP3+POPC CUDA-core kernel,
 * 2 = mma.sync b1 and.popc, 3 = tcgen05 (+-1 fp8 contraction, TMEM accumulators) one CTA per SM,
 * 4 = tcgen05 with cta_group::2 SM pairs (what auto resolves to for nq >= 128 and nd >= 1024). */
int orbgpu_knn2_set_engine(orbgpu_ctx *ctx, int32_t engine);

/* ---- 8(f) rank 2: KeyFrameDatabase candidate scoring -- the query BowVector against every key frame's BowVector.
 * common_words[kf] = number of words shared with the query (the inverted-file walk of KeyFrameDatabase.cc:928-943:
 * mnRelocWords / mnLoopWords / mnMergeWords); scores[kf] = L1Scoring::score(query, kf) (ScoringObject.cpp:23-68), double,
 * summed in ascending word order.  The caller casts to float (`float si = mpVoc->score(...)`, :970) and applies the
 * candidate policy (:949-1031).  The database is the CSR of the key frames' BowVectors (words ascending per key frame). */
typedef struct orbgpu_bowdb orbgpu_bowdb;
typedef struct orbgpu_bowdb_host {
    int32_t n_kf;
    const int32_t *offsets;  /* [n_kf+1] */
    const uint32_t *words;   /* [offsets[n_kf]] */
    const double *values;    /* [offsets[n_kf]] */
} orbgpu_bowdb_host;
int orbgpu_bowdb_upload(orbgpu_ctx *ctx, const orbgpu_bowdb_host *h, orbgpu_bowdb **out);
void orbgpu_bowdb_destroy(orbgpu_bowdb *d);
int orbgpu_bow_score_l1(orbgpu_ctx *ctx, const orbgpu_bowdb *db, int32_t nq_words, const uint32_t *q_words,
                        const double *q_values, int32_t *common_words, double *scores);

/* ---- 8(f) rank 4: batched MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:444-535).  Map point p observes the
 * descriptors desc[offsets[p] .. offsets[p+1]) (observation order, left then right index per key frame, :465-477).
 * best_idx[p] = the row with the smallest median distance to the others (first wins ties, :510-514), -1 for an empty
 * list; best_median[p] (may be NULL) = that median.  The caller clones vDescriptors[best_idx] into mDescriptor (:518). */
int orbgpu_compute_distinctive_descriptors(orbgpu_ctx *ctx, int32_t n_mp, const int32_t *offsets, const uint8_t *desc,
                                           int32_t *best_idx, int32_t *best_median);

/* ---- 8(f) rank 3: coarse stage of Frame::ComputeStereoMatches (Frame.cc:1139-1216): per left keypoint the right keypoint of
 * its row band (right keypoints are listed in rows [floor(y-r), ceil(y+r)], r = 2*scaleFactor[octave], :1143-1156) with the
 * smallest descriptor distance among those within one octave and with uR in [uL - mbf/mb, uL] (:1194-1200); bestDist starts
 * at TH_HIGH, strict <, first candidate (ascending right index) wins; accepted iff bestDist < (TH_HIGH+TH_LOW)/2 (:1214).
 * best_idx_r[i] = accepted right keypoint or -1; best_dist[i] = bestDist (TH_HIGH when no candidate).  Rows outside
 * [0, n_rows) -- undefined behaviour in the reference -- are ignored.  The SAD sub-pixel refinement (:1216-1290) reads the
 * image pyramids and stays with the caller. */
int orbgpu_stereo_coarse_match(orbgpu_ctx *ctx, int32_t n_left, const uint8_t *desc_l, const float *kp_xy_l, const int32_t *octave_l,
                               int32_t n_right, const uint8_t *desc_r, const float *kp_xy_r, const int32_t *octave_r,
                               const float *scale_factors, int32_t n_levels, int32_t n_rows, float mb, float mbf,
                               int32_t *best_idx_r, int32_t *best_dist);

/* ---- measurement aid: integer-pipe peak of this GPU, measured with a register-only microbenchmark on all SMs.
 * variant 0 = POPC + IADD3 only, variant 1 = the Hamming triple LOP3(xor) + POPC + IADD3 (one 32-bit slice of
 * DescriptorDistance, ORBmatcher.cc:2397-2405).  popc_per_s = POPC32 per second over the whole GPU; per_clk_sm = per SM clock
 * per SM; sm_mhz = the SM clock the kernel ran at.  Used by bench.py as the denominator of the POPC rooflines. */
int orbgpu_measure_popc_peak(orbgpu_ctx *ctx, int32_t variant, double *popc_per_s, double *per_clk_sm, double *sm_mhz);

/* ---- a10: ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2341-2383) exposed for testing:
 * histo[30] bin sizes -> ind[3]. Runs on the device. */
int orbgpu_compute_three_maxima(orbgpu_ctx *ctx, const int32_t *histo, int32_t L, int32_t *ind);

#ifdef __cplusplus
}
#endif
#endif /* ORBMATCH_B200_H */
