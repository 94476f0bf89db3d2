// ORACLE (test infrastructure, NOT product code) -- flat-array entry points around the
// reference's OWN matcher and vocabulary code.
//
// This translation unit is linked with the UNMODIFIED reference sources
//   /root/reference/src/ORBmatcher.cc
//   /root/reference/Thirdparty/DBoW2/DBoW2/{FORB,BowVector,FeatureVector,ScoringObject}.cpp
//   /root/reference/Thirdparty/DBoW2/DUtils/{Random,Timestamp}.cpp
// (compiled where they lie, see oracle/Makefile) into oracle/_ref/libref_orbmatcher.so.
// It builds stub Frame / KeyFrame / MapPoint objects (oracle/shim/orbslam_stubs.h) from the
// same flat structs the product ABI takes, calls ORB_SLAM3::ORBmatcher / ORBVocabulary, and
// flattens the STL results.  It is what pins the C restatement (orb_oracle.c) and the golden
// vectors (tests/golden/) to the reference's real behaviour.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "orbmatch_b200.h"

#include "ORBmatcher.h" // the reference's include/ORBmatcher.h (stubs resolve MapPoint.h/KeyFrame.h/Frame.h)

#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/ScoringObject.h"
// The fork added `std::vector<std::pair<TDescriptor, F>> vocabulary;` (TemplatedVocabulary.h:435)
// with F = FORB abstract, which does not compile as shipped.  Giving that one pair type an
// (empty) explicit specialisation lets the header compile unmodified; the member is unused.
namespace std
{
    template <>
    struct pair<cv::Mat, DBoW2::FORB>
    {
    };
}
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"
#include "Thirdparty/DBoW2/DUtils/Random.h"

using namespace ORB_SLAM3;

namespace
{
    struct TaggedMapPoint : public MapPoint
    {
        int tag = -1;
    };

    void fill_feature_set(FeatureSet &s, const orbgpu_frame_host *h)
    {
        s.N = h->n;
        s.mvKeysUn.resize(h->n);
        for (int i = 0; i < h->n; i++)
        {
            cv::KeyPoint kp;
            kp.pt.x = h->kp_xy[2 * i];
            kp.pt.y = h->kp_xy[2 * i + 1];
            kp.octave = h->octave[i];
            kp.angle = h->angle[i];
            s.mvKeysUn[i] = kp;
        }
        s.mvKeys = s.mvKeysUn;
        s.mvuRight.assign(h->n, -1.f);
        if (h->u_right)
            for (int i = 0; i < h->n; i++) s.mvuRight[i] = h->u_right[i];
        s.mDescriptors.create(h->n > 0 ? h->n : 1, 32, CV_8U);
        if (h->n > 0) std::memcpy(s.mDescriptors.data, h->desc, (size_t)h->n * 32);
        s.mnScaleLevels = h->n_levels;
        s.mvScaleFactors.assign(h->scale_factors, h->scale_factors + h->n_levels);
        s.mvLevelSigma2.assign(h->level_sigma2, h->level_sigma2 + h->n_levels);
        s.mvInvLevelSigma2.resize(h->n_levels);
        for (int i = 0; i < h->n_levels; i++) s.mvInvLevelSigma2[i] = 1.0f / s.mvLevelSigma2[i];
        s.mfLogScaleFactor = h->n_levels > 1 ? std::log(h->scale_factors[1]) : 1.f;
        s.mfGridElementWidthInv = h->grid_inv_w;
        s.mfGridElementHeightInv = h->grid_inv_h;
        s.mvpMapPoints.assign(h->n, nullptr);
        s.mFeatVec.clear();
        for (int a = 0; a < h->fv_n_nodes; a++)
            for (int j = h->fv_offsets[a]; j < h->fv_offsets[a + 1]; j++)
                s.mFeatVec.addFeature(h->fv_node_ids[a], h->fv_features[j]);
    }
    void fill_frame(Frame &F, const orbgpu_frame_host *h)
    {
        fill_feature_set(F, h);
        F.mnMinX = h->min_x; F.mnMinY = h->min_y; F.mnMaxX = h->max_x; F.mnMaxY = h->max_y;
        F.mvbOutlier.assign(h->n, false);
        F.AssignFeaturesToGrid(); // the reference's own Frame.cc:469-507 / :973-989 (oracle/_ref/gen/ref_extracted.cc)
    }
    void fill_keyframe(KeyFrame &K, const orbgpu_frame_host *h)
    {
        fill_feature_set(K, h);
        Frame F; // a key frame takes the grid of the frame it is made from (KeyFrame.cc:66-82)
        F.N = K.N; F.mvKeysUn = K.mvKeysUn;
        F.mfGridElementWidthInv = K.mfGridElementWidthInv; F.mfGridElementHeightInv = K.mfGridElementHeightInv;
        F.mnMinX = h->min_x; F.mnMinY = h->min_y; F.mnMaxX = h->max_x; F.mnMaxY = h->max_y;
        F.AssignFeaturesToGrid();
        K.CopyGridFrom(F);
    }

    struct ExposedMatcher : public ORBmatcher
    {
        ExposedMatcher() : ORBmatcher(0.6f, true) {}
        void three(std::vector<int> *h, int L, int &a, int &b, int &c) { ComputeThreeMaxima(h, L, a, b, c); }
    };

    typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> RefVocBase;
    struct RefVoc : public RefVocBase
    {
        RefVoc(int k, int L) : RefVocBase(k, L, DBoW2::TF_IDF, DBoW2::L1_NORM) {}
        using RefVocBase::m_nodes;
        using RefVocBase::m_words;
        using RefVocBase::m_k;
        using RefVocBase::m_L;
        void descend(const cv::Mat &f, DBoW2::WordId &id, DBoW2::WordValue &w, DBoW2::NodeId *nid, int levelsup) const
        {
            RefVocBase::transform(f, id, w, nid, levelsup);
        }
    };
} // namespace

extern "C"
{
    int ref_descriptor_distance(const uint8_t *a, const uint8_t *b)
    {
        cv::Mat ma(1, 32, CV_8U, (void *)a), mb(1, 32, CV_8U, (void *)b);
        return ORBmatcher::DescriptorDistance(ma, mb);
    }

    void ref_compute_three_maxima(const int32_t *histo_sizes, int L, int32_t *ind)
    {
        std::vector<std::vector<int>> h(L);
        for (int i = 0; i < L; i++) h[i].assign(histo_sizes[i], 0);
        ExposedMatcher m;
        int a = -1, b = -1, c = -1;
        m.three(h.data(), L, a, b, c);
        ind[0] = a; ind[1] = b; ind[2] = c;
    }

    // Frame::GetFeaturesInArea / AssignFeaturesToGrid / PosInGrid: the reference's own text (Frame.cc:868-962, 469-507, 973-989)
    int ref_features_in_area(const orbgpu_frame_host *f, float x, float y, float r, int min_level, int max_level,
                             int32_t *out_idx)
    {
        Frame F;
        fill_frame(F, f);
        std::vector<size_t> v = F.GetFeaturesInArea(x, y, r, min_level, max_level);
        for (size_t i = 0; i < v.size(); i++) out_idx[i] = (int32_t)v[i];
        return (int)v.size();
    }

    void ref_grid(const orbgpu_frame_host *f, int32_t *cell_start, int32_t *cell_items)
    {
        Frame F;
        fill_frame(F, f);
        int acc = 0;
        for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
            for (int iy = 0; iy < FRAME_GRID_ROWS; iy++)
            {
                cell_start[ix * FRAME_GRID_ROWS + iy] = acc;
                for (size_t j = 0; j < F.mGrid[ix][iy].size(); j++) cell_items[acc++] = (int32_t)F.mGrid[ix][iy][j];
            }
        cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = acc;
    }

    // KeyFrame::GetFeaturesInArea (KeyFrame.cc:859-907) + IsInImage (:910-913), the reference's own text; bounds as the KeyFrame
    // holds them (ints).  in_image [1] out.
    int ref_keyframe_features_in_area(const orbgpu_frame_host *f, float x, float y, float r, int32_t *out_idx, int32_t *in_image)
    {
        KeyFrame K;
        fill_keyframe(K, f);
        std::vector<size_t> v = K.GetFeaturesInArea(x, y, r);
        for (size_t i = 0; i < v.size(); i++) out_idx[i] = (int32_t)v[i];
        if (in_image) *in_image = K.IsInImage(x, y) ? 1 : 0;
        return (int)v.size();
    }

    // MapPoint::PredictScale (MapPoint.cc:695-738, both overloads) and Get{Min,Max}DistanceInvariance (:665-678)
    void ref_predict_scale(int n, const float *max_distance, const float *min_distance, const float *current_dist, float log_scale_factor,
                           int n_levels, int32_t *level_kf, int32_t *level_f, float *min_inv, float *max_inv)
    {
        KeyFrame K;
        Frame F;
        K.mfLogScaleFactor = F.mfLogScaleFactor = log_scale_factor;
        K.mnScaleLevels = F.mnScaleLevels = n_levels;
        for (int i = 0; i < n; i++)
        {
            MapPoint mp;
            mp.mfMaxDistance = max_distance[i];
            mp.mfMinDistance = min_distance[i];
            level_kf[i] = mp.PredictScale(current_dist[i], &K);
            level_f[i] = mp.PredictScale(current_dist[i], &F);
            min_inv[i] = mp.GetMinDistanceInvariance();
            max_inv[i] = mp.GetMaxDistanceInvariance();
        }
    }

    // Pinhole::project (Pinhole.cpp:64-71) and Pinhole::epipolarConstrain (:189-219) with the camera pair (K1, K2 = fx fy cx cy)
    void ref_pinhole_project(const float *K, int n, const float *xyz, float *uv)
    {
        Pinhole cam(K[0], K[1], K[2], K[3]);
        for (int i = 0; i < n; i++)
        {
            const Eigen::Vector2f p = cam.project(Eigen::Vector3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
            uv[2 * i] = p(0);
            uv[2 * i + 1] = p(1);
        }
    }
    void ref_epipolar_constrain(const float *K1, const float *K2, const float *R12, const float *t12, int n, const float *kp1_xy,
                                const float *kp2_xy, const float *unc, uint8_t *ok, float *f12_out)
    {
        Pinhole c1(K1[0], K1[1], K1[2], K1[3]), c2(K2[0], K2[1], K2[2], K2[3]);
        Eigen::Matrix3f R;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) R(r, c) = R12[3 * r + c];
        const Eigen::Vector3f t(t12[0], t12[1], t12[2]);
        for (int i = 0; i < n; i++)
        {
            cv::KeyPoint a, b;
            a.pt.x = kp1_xy[2 * i]; a.pt.y = kp1_xy[2 * i + 1];
            b.pt.x = kp2_xy[2 * i]; b.pt.y = kp2_xy[2 * i + 1];
            ok[i] = c1.epipolarConstrain(&c2, a, b, R, t, 1.f, unc[i]) ? 1 : 0;
        }
        if (f12_out) // the F12 the function forms (Pinhole.cpp:194-197), for callers that feed it to the flat-input oracle / the GPU
        {
            const Eigen::Matrix3f F = c1.fundamental(&c2, R, t);
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) f12_out[3 * r + c] = F(r, c);
        }
    }

    // Frame::isInFrustum (Frame.cc:676-782, Nleft == -1 path): the reference's own text on a frame with pose Tcw (row-major 3x4),
    // pinhole K, image bounds of the frame.  Outputs = the members the function writes into each MapPoint.
    void ref_is_in_frustum(const orbgpu_frame_host *f, const float *Tcw34, const float *K, float mbf, float viewing_cos_limit, int n,
                           const float *world_pos, const float *normal, const float *min_distance, const float *max_distance,
                           uint8_t *in_view, float *proj_xy, float *proj_xr, float *depth, int32_t *scale_level, float *view_cos, uint8_t *ret)
    {
        Frame F;
        fill_frame(F, f);
        Eigen::Matrix3f R;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) R(r, c) = Tcw34[4 * r + c];
        F.SetPose(Sophus::SE3f(R, Eigen::Vector3f(Tcw34[3], Tcw34[7], Tcw34[11])));
        Pinhole cam(K[0], K[1], K[2], K[3]);
        F.mpCamera = &cam;
        F.mbf = mbf;
        for (int i = 0; i < n; i++)
        {
            MapPoint mp;
            mp.worldPos_ = Eigen::Vector3f(world_pos[3 * i], world_pos[3 * i + 1], world_pos[3 * i + 2]);
            mp.normal_ = Eigen::Vector3f(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
            mp.mfMinDistance = min_distance[i];
            mp.mfMaxDistance = max_distance[i];
            ret[i] = F.isInFrustum(&mp, viewing_cos_limit) ? 1 : 0;
            in_view[i] = mp.mbTrackInView ? 1 : 0;
            proj_xy[2 * i] = mp.mTrackProjX;
            proj_xy[2 * i + 1] = mp.mTrackProjY;
            proj_xr[i] = mp.mTrackProjXR;
            depth[i] = mp.mTrackDepth;
            scale_level[i] = mp.mnTrackScaleLevel;
            view_cos[i] = mp.mTrackViewCos;
        }
    }

    // coarse stage of Frame::ComputeStereoMatches (Frame.cc:1117-1247): the reference's own text up to the SAD refinement
    void ref_stereo_coarse_match(int32_t n_left, const uint8_t *desc_l, const float *kp_xy_l, const int32_t *octave_l, int32_t n_right,
                                 const uint8_t *desc_r, const float *kp_xy_r, const int32_t *octave_r, const float *scale_factors,
                                 int32_t n_levels, int32_t n_rows, float mb, float mbf, int32_t *best_idx_r, int32_t *best_dist)
    {
        Frame F;
        F.N = n_left;
        F.mvKeys.resize(n_left);
        for (int i = 0; i < n_left; i++)
        {
            F.mvKeys[i].pt.x = kp_xy_l[2 * i]; F.mvKeys[i].pt.y = kp_xy_l[2 * i + 1]; F.mvKeys[i].octave = octave_l[i];
        }
        F.mvKeysRight.resize(n_right);
        for (int i = 0; i < n_right; i++)
        {
            F.mvKeysRight[i].pt.x = kp_xy_r[2 * i]; F.mvKeysRight[i].pt.y = kp_xy_r[2 * i + 1]; F.mvKeysRight[i].octave = octave_r[i];
        }
        F.mDescriptors.create(n_left > 0 ? n_left : 1, 32, CV_8U);
        if (n_left > 0) std::memcpy(F.mDescriptors.data, desc_l, (size_t)n_left * 32);
        F.mDescriptorsRight.create(n_right > 0 ? n_right : 1, 32, CV_8U);
        if (n_right > 0) std::memcpy(F.mDescriptorsRight.data, desc_r, (size_t)n_right * 32);
        F.mvScaleFactors.assign(scale_factors, scale_factors + n_levels);
        F.mvInvScaleFactors.resize(n_levels);
        for (int i = 0; i < n_levels; i++) F.mvInvScaleFactors[i] = 1.0f / scale_factors[i];
        F.mb = mb;
        F.mbf = mbf;
        ORBextractorStub ex;
        ex.mvImagePyramid.resize(1);
        ex.mvImagePyramid[0].rows = n_rows;
        F.mpORBextractorLeft = F.mpORBextractorRight = &ex;
        F.stereo_best_idx_.assign(n_left > 0 ? n_left : 1, -1);
        F.stereo_best_dist_.assign(n_left > 0 ? n_left : 1, ORBmatcher::TH_HIGH);
        F.ComputeStereoMatches();
        for (int i = 0; i < n_left; i++)
        {
            best_idx_r[i] = F.stereo_best_idx_[i];
            best_dist[i] = F.stereo_best_dist_[i];
        }
    }

    int ref_search_for_initialization(const orbgpu_frame_host *f1, const orbgpu_frame_host *f2, float *prev_matched_xy,
                                      int window_size, float nnratio, int check_ori, int32_t *matches12)
    {
        Frame F1, F2;
        fill_frame(F1, f1);
        fill_frame(F2, f2);
        std::vector<cv::Point2f> prev(f1->n);
        for (int i = 0; i < f1->n; i++) prev[i] = cv::Point2f(prev_matched_xy[2 * i], prev_matched_xy[2 * i + 1]);
        std::vector<int> m12;
        ORBmatcher matcher(nnratio, check_ori != 0);
        const int n = matcher.SearchForInitialization(F1, F2, prev, m12, window_size);
        for (int i = 0; i < f1->n; i++)
        {
            matches12[i] = m12[i];
            prev_matched_xy[2 * i] = prev[i].x;
            prev_matched_xy[2 * i + 1] = prev[i].y;
        }
        return n;
    }

    int ref_search_by_projection_local(const orbgpu_frame_host *f, const orbgpu_mappoints_host *mps, float th,
                                       int far_points, float th_far_points, float nnratio, const int32_t *kp_prior_obs,
                                       int32_t *kp_mp)
    {
        Frame F;
        fill_frame(F, f);
        std::vector<TaggedMapPoint> pts(mps->n);
        std::vector<MapPoint *> vp(mps->n);
        for (int i = 0; i < mps->n; i++)
        {
            TaggedMapPoint &p = pts[i];
            p.tag = i;
            p.descriptor_.create(1, 32, CV_8U);
            std::memcpy(p.descriptor_.data, mps->desc + 32 * (size_t)i, 32);
            p.mTrackProjX = mps->proj_xy[2 * i];
            p.mTrackProjY = mps->proj_xy[2 * i + 1];
            p.mTrackProjXR = mps->proj_xr ? mps->proj_xr[i] : 0.f;
            p.mnTrackScaleLevel = mps->scale_level[i];
            p.mTrackViewCos = mps->view_cos[i];
            p.mTrackDepth = mps->depth[i];
            p.mbTrackInView = mps->in_view[i] != 0;
            p.mbTrackInViewR = false;
            p.bad_ = mps->bad[i] != 0;
            p.nObs_ = mps->n_obs[i];
            vp[i] = &p;
        }
        // map points the frame already holds on entry (not part of vpMapPoints)
        std::vector<TaggedMapPoint> prior(f->n);
        for (int i = 0; i < f->n; i++)
        {
            prior[i].tag = -2 - i;
            prior[i].nObs_ = kp_prior_obs[i];
            if (kp_mp[i] >= 0 || kp_prior_obs[i] > 0) F.mvpMapPoints[i] = &prior[i];
        }
        ORBmatcher matcher(nnratio, true);
        const int n = matcher.SearchByProjection(F, vp, th, far_points != 0, th_far_points);
        for (int i = 0; i < f->n; i++)
        {
            MapPoint *p = F.mvpMapPoints[i];
            if (!p) kp_mp[i] = -1;
            else
            {
                const int tag = static_cast<TaggedMapPoint *>(p)->tag;
                if (tag >= 0) kp_mp[i] = tag; /* else: keeps the caller's prior value */
            }
        }
        return n;
    }

    // ------------------------------------------------------------------------------------------------------------
    // Self-projecting overloads driven from flat "already projected" points.  The geometry is
    // degenerate on purpose: identity pose, pinhole fx = fy = 1, cx = cy = 0 and world points (u, v, 1), so that the
    // reference's OWN prologue (Tcw * x3Dw, mpCamera->project, PredictScale, ...) reproduces the given (u, v), radius
    // and level exactly (1*u + 0*v + 0*1 + 0 and 1*u/1 + 0 are exact in fp32).  What is pinned here is everything
    // from the window on; the general-pose prologue runs in tests/cpp/adapter_parity.cc.
    namespace
    {
        struct ProjWorld
        {
            std::vector<TaggedMapPoint> pts, locked;
            Pinhole cam{1.f, 1.f, 0.f, 0.f};
        };
        // level of point m as the generator defined it from (min_level, max_level) and the window mode
        int level_of(const orbgpu_projpoints_host *p, int m, int mode)
        {
            if (mode == 1) return p->min_level[m];      // fwd: (l, -1)
            if (mode == 2) return p->max_level[m];      // bwd: (0, l)
            if (mode == 3) return p->max_level[m];      // pred: (l-1, l)
            return p->min_level[m] + 1;                 // pm1: (l-1, l+1)
        }
        void make_points(ProjWorld &w, const orbgpu_projpoints_host *p, int mode, const float *scale_log, int n_levels)
        {
            (void)scale_log; (void)n_levels;
            w.pts.resize(p->n);
            for (int i = 0; i < p->n; i++)
            {
                TaggedMapPoint &mp = w.pts[i];
                mp.tag = i;
                mp.descriptor_.create(1, 32, CV_8U);
                std::memcpy(mp.descriptor_.data, p->desc + 32 * (size_t)i, 32);
                mp.worldPos_ = Eigen::Vector3f(p->uv[2 * i], p->uv[2 * i + 1], 1.f);
                mp.nObs_ = (!p->locks || p->locks[i]) ? 1 : 0;
                mp.bad_ = false;
                // distance gates and PredictScale (MapPoint.cc:695-738): dist = |PO| with Ow = 0
                const float dist = mp.worldPos_.norm();
                mp.normal_ = Eigen::Vector3f(mp.worldPos_(0) / dist, mp.worldPos_(1) / dist, mp.worldPos_(2) / dist);
                mp.mfMinDistance = 0.f;
                mp.mfMaxDistance = (float)((double)dist * std::pow(1.2, (double)level_of(p, i, mode) - 0.5));
            }
        }
        void lock_keypoints(ProjWorld &w, std::vector<MapPoint *> &slots, const uint8_t *kp_locked, int n)
        {
            w.locked.resize(n);
            for (int k = 0; k < n; k++)
            {
                w.locked[k].tag = -2;
                w.locked[k].nObs_ = 1;
                if (kp_locked && kp_locked[k]) slots[k] = &w.locked[k];
            }
        }
        void read_owners(const std::vector<MapPoint *> &slots, int32_t *kp_owner)
        {
            for (size_t k = 0; k < slots.size(); k++)
            {
                const int tag = slots[k] ? static_cast<TaggedMapPoint *>(slots[k])->tag : -1;
                kp_owner[k] = tag >= 0 ? tag : -1;
            }
        }
    } // namespace

    // ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (ORBmatcher.cc:1957-2191); mode 0 pm1 / 1 fwd / 2 bwd
    int ref_projected_cur_last(const orbgpu_frame_host *cur, const orbgpu_projpoints_host *p, float th, int mode, float mbf,
                               const uint8_t *kp_locked, int check_ori, int32_t *kp_owner)
    {
        ProjWorld w;
        Frame C, L;
        fill_frame(C, cur);
        C.mpCamera = &w.cam;
        C.mb = 0.5f;
        C.mbf = mbf;
        make_points(w, p, mode, nullptr, 0);
        lock_keypoints(w, C.mvpMapPoints, kp_locked, cur->n);
        L.N = p->n;
        L.mvKeys.resize(p->n);
        L.mvKeysUn.resize(p->n);
        L.mvpMapPoints.assign(p->n, nullptr);
        L.mvbOutlier.assign(p->n, false);
        for (int i = 0; i < p->n; i++)
        {
            cv::KeyPoint kp;
            kp.octave = level_of(p, i, mode);
            kp.angle = p->angle ? p->angle[i] : 0.f;
            L.mvKeys[i] = kp;
            L.mvKeysUn[i] = kp;
            if (p->active[i]) L.mvpMapPoints[i] = &w.pts[i];
        }
        // tlc = Tlw * twc with twc = 0: the sign of its z against mb selects forward / backward (:1969-1972)
        const float tz = mode == 1 ? 1.f : (mode == 2 ? -1.f : 0.f);
        L.mTcw = Sophus::SE3f(Eigen::Matrix3f::Identity(), Eigen::Vector3f(0.f, 0.f, tz));
        ORBmatcher matcher(0.9f, check_ori != 0);
        const int n = matcher.SearchByProjection(C, L, th, false);
        read_owners(C.mvpMapPoints, kp_owner);
        return n;
    }

    // ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist) (:2203-2330)
    int ref_projected_reloc(const orbgpu_frame_host *cur, const orbgpu_projpoints_host *p, float th, int orb_dist,
                            const uint8_t *kp_locked, int check_ori, int32_t *kp_owner)
    {
        ProjWorld w;
        Frame C;
        KeyFrame K;
        fill_frame(C, cur);
        C.mpCamera = &w.cam;
        make_points(w, p, 0, nullptr, 0);
        lock_keypoints(w, C.mvpMapPoints, kp_locked, cur->n);
        K.N = p->n;
        K.mvKeysUn.resize(p->n);
        K.mvpMapPoints.assign(p->n, nullptr);
        std::set<MapPoint *> found;
        for (int i = 0; i < p->n; i++)
        {
            K.mvKeysUn[i].angle = p->angle ? p->angle[i] : 0.f;
            K.mvpMapPoints[i] = &w.pts[i];
            if (!p->active[i]) found.insert(&w.pts[i]); // sAlreadyFound skips it (:2222)
        }
        ORBmatcher matcher(0.9f, check_ori != 0);
        const int n = matcher.SearchByProjection(C, &K, found, th, orb_dist);
        read_owners(C.mvpMapPoints, kp_owner);
        return n;
    }

    // ORBmatcher::SearchByProjection(KeyFrame*, Sim3f&, vpPoints, vpMatched, th, ratioHamming) (:498-620)
    int ref_projected_sim3(const orbgpu_frame_host *kf, const orbgpu_projpoints_host *p, int th, float ratio_hamming,
                           const uint8_t *kp_locked, int32_t *kp_owner)
    {
        ProjWorld w;
        KeyFrame K;
        fill_keyframe(K, kf);
        K.mpCamera = &w.cam;
        make_points(w, p, 3, nullptr, 0);
        std::vector<MapPoint *> vp(p->n), vpMatched(kf->n, nullptr);
        for (int i = 0; i < p->n; i++)
        {
            vp[i] = &w.pts[i];
            w.pts[i].bad_ = !p->active[i];
        }
        lock_keypoints(w, vpMatched, kp_locked, kf->n);
        Sophus::Sim3f Scw;
        ORBmatcher matcher(0.9f, true);
        const int n = matcher.SearchByProjection(&K, Scw, vp, vpMatched, th, ratio_hamming);
        read_owners(vpMatched, kp_owner);
        return n;
    }

    // ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th, bRight=false) (:1330-1545): best_idx per point from the
    // side effects (AddObservation :1519-1520, or Replace on the point the keypoint already holds :1505-1515)
    int ref_projected_fuse(const orbgpu_frame_host *kf, const orbgpu_projpoints_host *p, float th, float bf, int32_t *best_idx)
    {
        ProjWorld w;
        KeyFrame K;
        fill_keyframe(K, kf);
        K.mpCamera = &w.cam;
        K.mbf = bf;
        make_points(w, p, 3, nullptr, 0);
        std::vector<MapPoint *> vp(p->n);
        for (int i = 0; i < p->n; i++)
        {
            w.pts[i].nObs_ = 0;
            vp[i] = p->active[i] ? &w.pts[i] : nullptr;
            best_idx[i] = -1;
        }
        g_replace_log().clear();
        ORBmatcher matcher(0.9f, true);
        const int n = matcher.Fuse(&K, vp, th, false);
        for (int i = 0; i < p->n; i++)
            if (w.pts[i].observations_.count(&K)) best_idx[i] = std::get<0>(w.pts[i].observations_[&K]);
        for (const auto &e : g_replace_log()) // pMP->Replace(pMPinKF) or pMPinKF->Replace(pMP) (:1508-1511): the one that
        {                                     // observes the keyframe is the point held by the keypoint
            TaggedMapPoint *a = static_cast<TaggedMapPoint *>(e.first), *b = static_cast<TaggedMapPoint *>(e.second);
            TaggedMapPoint *held = a->observations_.count(&K) ? a : b, *fused = held == a ? b : a;
            best_idx[fused->tag] = std::get<0>(held->observations_[&K]);
        }
        return n;
    }

    // ORBmatcher::Fuse(KeyFrame*, Sim3f&, vpPoints, th, vpReplacePoint) (:1547-1682)
    int ref_projected_fuse_sim3(const orbgpu_frame_host *kf, const orbgpu_projpoints_host *p, float th, int32_t *best_idx)
    {
        ProjWorld w;
        KeyFrame K;
        fill_keyframe(K, kf);
        K.mpCamera = &w.cam;
        make_points(w, p, 3, nullptr, 0);
        std::vector<MapPoint *> vp(p->n), repl(p->n, nullptr);
        for (int i = 0; i < p->n; i++)
        {
            vp[i] = &w.pts[i];
            w.pts[i].bad_ = !p->active[i];
            best_idx[i] = -1;
        }
        Sophus::Sim3f Scw;
        ORBmatcher matcher(0.9f, true);
        const int n = matcher.Fuse(&K, Scw, vp, th, repl);
        for (int i = 0; i < p->n; i++)
        {
            if (w.pts[i].observations_.count(&K)) best_idx[i] = std::get<0>(w.pts[i].observations_[&K]);
            if (repl[i]) best_idx[i] = std::get<0>(static_cast<TaggedMapPoint *>(repl[i])->observations_[&K]);
        }
        return n;
    }

    // DBoW2::L1Scoring::score (the reference's ScoringObject.cpp, compiled unmodified) of the query against every key
    // frame of a CSR database; common words counted from the same std::map objects (KeyFrameDatabase.cc:928-943 walks an
    // inverted file over exactly these (word, key frame) incidences)
    void ref_bow_score_l1(const orbgpu_bowdb_host *db, int32_t nq, const uint32_t *q_words, const double *q_values,
                          int32_t *common_words, double *scores)
    {
        DBoW2::BowVector q;
        for (int i = 0; i < nq; i++) q.insert(std::make_pair(q_words[i], q_values[i]));
        DBoW2::L1Scoring l1;
        for (int kf = 0; kf < db->n_kf; kf++)
        {
            DBoW2::BowVector v;
            int n_common = 0;
            for (int j = db->offsets[kf]; j < db->offsets[kf + 1]; j++)
            {
                v.insert(std::make_pair(db->words[j], db->values[j]));
                n_common += (int)q.count(db->words[j]);
            }
            common_words[kf] = n_common;
            scores[kf] = l1.score(q, v);
        }
    }

    int ref_search_by_bow_kf_f(const orbgpu_frame_host *kf, const orbgpu_frame_host *f, const uint8_t *kf_mp_valid,
                               float nnratio, int check_ori, int32_t *match_f2kf)
    {
        KeyFrame K;
        Frame F;
        fill_keyframe(K, kf);
        fill_frame(F, f);
        std::vector<TaggedMapPoint> pts(kf->n);
        for (int i = 0; i < kf->n; i++)
        {
            pts[i].tag = i;
            if (kf_mp_valid[i]) K.mvpMapPoints[i] = &pts[i];
        }
        std::vector<MapPoint *> out;
        ORBmatcher matcher(nnratio, check_ori != 0);
        const int n = matcher.SearchByBoW(&K, F, out);
        for (int i = 0; i < f->n; i++) match_f2kf[i] = out[i] ? static_cast<TaggedMapPoint *>(out[i])->tag : -1;
        return n;
    }

    int ref_search_by_bow_kf_kf(const orbgpu_frame_host *kf1, const orbgpu_frame_host *kf2, const uint8_t *kf1_mp_valid,
                                const uint8_t *kf2_mp_valid, float nnratio, int check_ori, int32_t *match_12)
    {
        KeyFrame K1, K2;
        fill_keyframe(K1, kf1);
        fill_keyframe(K2, kf2);
        std::vector<TaggedMapPoint> p1(kf1->n), p2(kf2->n);
        for (int i = 0; i < kf1->n; i++)
        {
            p1[i].tag = i;
            if (kf1_mp_valid[i]) K1.mvpMapPoints[i] = &p1[i];
        }
        for (int i = 0; i < kf2->n; i++)
        {
            p2[i].tag = i;
            if (kf2_mp_valid[i]) K2.mvpMapPoints[i] = &p2[i];
        }
        std::vector<MapPoint *> out;
        ORBmatcher matcher(nnratio, check_ori != 0);
        const int n = matcher.SearchByBoW(&K1, &K2, out);
        for (int i = 0; i < kf1->n; i++) match_12[i] = out[i] ? static_cast<TaggedMapPoint *>(out[i])->tag : -1;
        return n;
    }

    // ---- triangulation ------------------------------------------------------------------
    static void pose_from_flat(const float *T, Sophus::SE3f &out)
    { // T = [R row-major (9) | t (3)]
        Eigen::Matrix3f R;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) R(r, c) = T[3 * r + c];
        out = Sophus::SE3f(R, Eigen::Vector3f(T[9], T[10], T[11]));
    }

    // The host-side pose algebra exactly as ORBmatcher.cc:1053-1071 + Pinhole.cpp:194-197 run it
    // in this build: ep (epipole of camera 1 in image 2) and F12 (row-major).  Its outputs are the
    // inputs of both orbgpu_search_for_triangulation_batch and oracle_search_for_triangulation.
    void ref_triangulation_geometry(const float *T1w, const float *T2w, const float *K1, const float *K2, float *ep,
                                    float *f12)
    {
        KeyFrame A, B;
        pose_from_flat(T1w, A.mTcw);
        pose_from_flat(T2w, B.mTcw);
        Pinhole c1(K1[0], K1[1], K1[2], K1[3]), c2(K2[0], K2[1], K2[2], K2[3]);
        Sophus::SE3f T1 = A.GetPose();
        Sophus::SE3f T2 = B.GetPose();
        Sophus::SE3f Tw2 = B.GetPoseInverse();
        Eigen::Vector3f Cw = A.GetCameraCenter();
        Eigen::Vector3f C2 = T2 * Cw;
        Eigen::Vector2f e = c2.project(C2);
        Sophus::SE3f T12 = T1 * Tw2;
        Eigen::Matrix3f R12 = T12.rotationMatrix();
        Eigen::Vector3f t12 = T12.translation();
        Eigen::Matrix3f F = c1.fundamental(&c2, R12, t12);
        ep[0] = e(0); ep[1] = e(1);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) f12[3 * r + c] = F(r, c);
    }

    static void fill_kf_from_set(KeyFrame &K, const orbgpu_kfset_host *s, int kf, std::vector<TaggedMapPoint> &pts)
    {
        const int n = s->n_feat;
        K.N = n;
        K.mvKeysUn.resize(n);
        const float *xy = s->kp_xy + (size_t)kf * n * 2;
        for (int i = 0; i < n; i++)
        {
            cv::KeyPoint kp;
            kp.pt.x = xy[2 * i]; kp.pt.y = xy[2 * i + 1];
            kp.octave = s->octave[(size_t)kf * n + i];
            kp.angle = s->angle[(size_t)kf * n + i];
            K.mvKeysUn[i] = kp;
        }
        K.mvKeys = K.mvKeysUn;
        K.mvuRight.assign(n, -1.f);
        if (s->u_right)
            for (int i = 0; i < n; i++) K.mvuRight[i] = s->u_right[(size_t)kf * n + i];
        K.mDescriptors = cv::Mat(n, 32, CV_8U, (void *)(s->desc + (size_t)kf * n * 32));
        K.mnScaleLevels = s->n_levels;
        K.mvScaleFactors.assign(s->scale_factors, s->scale_factors + s->n_levels);
        K.mvLevelSigma2.assign(s->level_sigma2, s->level_sigma2 + s->n_levels);
        pts.resize(n);
        K.mvpMapPoints.assign(n, nullptr);
        for (int i = 0; i < n; i++)
            if (s->has_mp[(size_t)kf * n + i]) K.mvpMapPoints[i] = &pts[i];
        K.mFeatVec.clear();
        for (int i = 0; i < n; i++)
        {
            const uint32_t nid = s->node_id[(size_t)kf * n + i];
            if (nid != 0xFFFFFFFFu) K.mFeatVec.addFeature(nid, (unsigned)i);
        }
    }

    int ref_search_for_triangulation(const orbgpu_kfset_host *s, int kf1, int kf2, const float *T1w, const float *T2w,
                                     const float *K1, const float *K2, int only_stereo, int coarse, int check_ori,
                                     float nnratio, int32_t *matches12)
    {
        KeyFrame A, B;
        std::vector<TaggedMapPoint> pa, pb;
        fill_kf_from_set(A, s, kf1, pa);
        fill_kf_from_set(B, s, kf2, pb);
        pose_from_flat(T1w, A.mTcw);
        pose_from_flat(T2w, B.mTcw);
        Pinhole c1(K1[0], K1[1], K1[2], K1[3]), c2(K2[0], K2[1], K2[2], K2[3]);
        A.mpCamera = &c1;
        B.mpCamera = &c2;
        std::vector<std::pair<size_t, size_t>> pairs;
        ORBmatcher matcher(nnratio, check_ori != 0);
        const int n = matcher.SearchForTriangulation(&A, &B, pairs, only_stereo != 0, coarse != 0);
        for (int i = 0; i < s->n_feat; i++) matches12[i] = -1;
        for (size_t i = 0; i < pairs.size(); i++) matches12[pairs[i].first] = (int32_t)pairs[i].second;
        return n;
    }

    // all pairs across n_threads host threads (the reference itself is single-threaded per call;
    // LocalMapping.cc:556-630 runs one call per neighbour keyframe)
    void ref_search_for_triangulation_batch(const orbgpu_kfset_host *s, int n_pairs, const int32_t *kf1, const int32_t *kf2,
                                            const float *T1w, const float *T2w, const float *K, int only_stereo, int coarse,
                                            int check_ori, float nnratio, int32_t *matches12, int32_t *nmatches,
                                            int n_threads)
    {
        if (n_threads < 1) n_threads = 1;
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++)
            th.emplace_back([=]() {
                const int p0 = (int)((int64_t)n_pairs * t / n_threads), p1 = (int)((int64_t)n_pairs * (t + 1) / n_threads);
                for (int p = p0; p < p1; p++)
                    nmatches[p] = ref_search_for_triangulation(s, kf1[p], kf2[p], T1w + 12 * (size_t)p, T2w + 12 * (size_t)p, K,
                                                               K, only_stereo, coarse, check_ori, nnratio,
                                                               matches12 + (size_t)p * s->n_feat);
            });
        for (auto &x : th) x.join();
    }

    // brute-force 2-NN + ratio over the reference's DescriptorDistance (the function itself is
    // "defined by this repo": SearchByNN does not exist in this fork)
    void ref_knn2_ratio(int64_t nq, const uint8_t *q, int64_t nd, const uint8_t *db, int th_low, float nnratio,
                        int32_t *best_idx, int32_t *best_dist, int32_t *second_dist, int32_t *match, int n_threads)
    {
        if (n_threads < 1) n_threads = 1;
        cv::Mat Q((int)nq, 32, CV_8U, (void *)q), D((int)nd, 32, CV_8U, (void *)db);
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++)
            th.emplace_back([&, t]() {
                const int64_t q0 = nq * t / n_threads, q1 = nq * (t + 1) / n_threads;
                for (int64_t i = q0; i < q1; i++)
                {
                    const cv::Mat &dq = Q.row((int)i);
                    int bestDist1 = 256, bestIdx = -1, bestDist2 = 256;
                    for (int64_t d = 0; d < nd; d++)
                    {
                        const cv::Mat &dd = D.row((int)d);
                        const int dist = ORBmatcher::DescriptorDistance(dq, dd);
                        if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx = (int)d; }
                        else if (dist < bestDist2) { bestDist2 = dist; }
                    }
                    int m = -1;
                    if (bestDist1 <= th_low)
                        if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) m = bestIdx;
                    if (best_idx) best_idx[i] = bestIdx;
                    if (best_dist) best_dist[i] = bestDist1;
                    if (second_dist) second_dist[i] = bestDist2;
                    if (match) match[i] = m;
                }
            });
        for (auto &x : th) x.join();
    }

    // ---- vocabulary -----------------------------------------------------------------------
    void *ref_voc_create(int k, int L, int n_images, int n_per_image, const uint8_t *desc, int seed)
    {
        DUtils::Random::SeedRandOnce(seed);
        DUtils::Random::SeedRand(seed);
        std::vector<std::vector<cv::Mat>> training(n_images);
        for (int im = 0; im < n_images; im++)
        {
            training[im].resize(n_per_image);
            for (int j = 0; j < n_per_image; j++)
            {
                cv::Mat m(1, 32, CV_8U);
                std::memcpy(m.data, desc + ((size_t)im * n_per_image + j) * 32, 32);
                training[im][j] = m;
            }
        }
        RefVoc *v = new RefVoc(k, L);
        v->create(training, k, L);
        return v;
    }

    void *ref_voc_from_flat(const orbgpu_voc_host *h)
    {
        RefVoc *v = new RefVoc(h->k, h->L);
        v->m_nodes.clear();
        v->m_nodes.resize(h->n_nodes);
        size_t n_words = 0;
        for (int i = 0; i < h->n_nodes; i++)
        {
            auto &nd = v->m_nodes[i];
            nd.id = i;
            nd.weight = h->weight[i];
            nd.word_id = h->word_id[i];
            nd.descriptor.create(1, 32, CV_8U);
            std::memcpy(nd.descriptor.data, h->node_desc + 32 * (size_t)i, 32);
            for (int c = h->child_offsets[i]; c < h->child_offsets[i + 1]; c++)
            {
                nd.children.push_back(h->child_ids[c]);
                v->m_nodes[h->child_ids[c]].parent = i;
            }
            if (i > 0 && nd.children.empty()) n_words = std::max(n_words, (size_t)h->word_id[i] + 1);
        }
        v->m_words.assign(n_words, nullptr);
        for (int i = 1; i < h->n_nodes; i++)
            if (v->m_nodes[i].children.empty()) v->m_words[h->word_id[i]] = &v->m_nodes[i];
        return v;
    }

    void ref_voc_destroy(void *p) { delete static_cast<RefVoc *>(p); }
    int ref_voc_n_nodes(void *p) { return (int)static_cast<RefVoc *>(p)->m_nodes.size(); }
    int ref_voc_n_children(void *p)
    {
        size_t n = 0;
        for (auto &nd : static_cast<RefVoc *>(p)->m_nodes) n += nd.children.size();
        return (int)n;
    }
    int ref_voc_n_words(void *p) { return (int)static_cast<RefVoc *>(p)->m_words.size(); }
    // export into caller-allocated flat arrays (layout of orbgpu_voc_host)
    void ref_voc_export(void *p, uint8_t *node_desc, int32_t *child_offsets, uint32_t *child_ids, double *weight,
                        uint32_t *word_id)
    {
        RefVoc *v = static_cast<RefVoc *>(p);
        int acc = 0;
        for (size_t i = 0; i < v->m_nodes.size(); i++)
        {
            auto &nd = v->m_nodes[i];
            child_offsets[i] = acc;
            for (auto c : nd.children) child_ids[acc++] = c;
            weight[i] = nd.weight;
            word_id[i] = nd.word_id;
            if (!nd.descriptor.empty()) std::memcpy(node_desc + 32 * i, nd.descriptor.data, 32);
            else std::memset(node_desc + 32 * i, 0, 32);
        }
        child_offsets[v->m_nodes.size()] = acc;
    }

    // TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup) -- the batch
    // call Frame::ComputeBoW makes (Frame.cc:1005-1008) -- plus the per-feature descent outputs.
    // Returns n_words; bow arrays capacity n, featvec capacities n / n+1 / n.
    int ref_voc_transform(void *p, int n, const uint8_t *desc, int levelsup, uint32_t *word_id, uint32_t *node_id,
                          double *weight, uint32_t *bow_words, double *bow_values, int32_t *fv_n_nodes,
                          uint32_t *fv_node_ids, int32_t *fv_offsets, uint32_t *fv_features)
    {
        RefVoc *v = static_cast<RefVoc *>(p);
        std::vector<cv::Mat> feats(n);
        for (int i = 0; i < n; i++) feats[i] = cv::Mat(1, 32, CV_8U, (void *)(desc + 32 * (size_t)i));
        for (int i = 0; i < n; i++)
        {
            DBoW2::WordId id = 0;
            DBoW2::WordValue w = 0;
            DBoW2::NodeId nid = 0;
            v->descend(feats[i], id, w, &nid, levelsup);
            if (word_id) word_id[i] = id;
            if (node_id) node_id[i] = nid;
            if (weight) weight[i] = w;
        }
        DBoW2::BowVector bv;
        DBoW2::FeatureVector fv;
        v->transform(feats, bv, fv, levelsup);
        int nw = 0;
        for (auto &kv : bv)
        {
            if (bow_words) bow_words[nw] = kv.first;
            if (bow_values) bow_values[nw] = kv.second;
            nw++;
        }
        int nn = 0, acc = 0;
        for (auto &kv : fv)
        {
            if (fv_node_ids) fv_node_ids[nn] = kv.first;
            if (fv_offsets) fv_offsets[nn] = acc;
            for (auto f : kv.second)
            {
                if (fv_features) fv_features[acc] = f;
                acc++;
            }
            nn++;
        }
        if (fv_offsets) fv_offsets[nn] = acc;
        if (fv_n_nodes) *fv_n_nodes = nn;
        return nw;
    }
}
