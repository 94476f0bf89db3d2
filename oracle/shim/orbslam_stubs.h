// ORACLE BUILD SHIM (test infrastructure) -- stand-ins for ORB_SLAM3::MapPoint, KeyFrame,
// Frame and GeometricCamera exposing exactly the members the reference's ORBmatcher.cc
// touches, so that /root/reference/src/ORBmatcher.cc + include/ORBmatcher.h compile
// UNMODIFIED into oracle/_ref/libref_orbmatcher.so (oracle/Makefile).  The real headers
// pull in OpenCV/Eigen/Sophus/Boost/g2o/Pangolin, none of which exist in this image.
//
// The helper functions ORBmatcher.cc calls INTO these classes are only DECLARED here.  Their bodies are the reference's own
// text, cut out of /root/reference by line range at build time (oracle/extract_ref.py -> oracle/_ref/gen/ref_extracted.cc):
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea / isInFrustum   src/Frame.cc:469-507, 973-989, 868-962, 676-782
//   Frame::ComputeStereoMatches, coarse stage                                  src/Frame.cc:1117-1247
//   KeyFrame::GetFeaturesInArea / IsInImage                                    src/KeyFrame.cc:859-913
//   MapPoint::PredictScale x2 / Get{Min,Max}DistanceInvariance                 src/MapPoint.cc:665-738
//   Pinhole::project / toK_ / epipolarConstrain                                src/CameraModels/Pinhole.cpp:64-71, 171-176, 189-219
// What IS written here: data members, trivial accessors, and the grid copy of the KeyFrame constructor (KeyFrame.cc:66-82, a
// member-wise copy).  Nothing here is product code.
#ifndef ORB_ORACLE_SHIM_ORBSLAM_STUBS_H
#define ORB_ORACLE_SHIM_ORBSLAM_STUBS_H

#include <cassert>
#include <cmath>
#include <map>
#include <mutex>
#include <set>
#include <stdexcept>
#include <tuple>
#include <vector>

#include <opencv2/core/core.hpp>
#include <sophus/sim3.hpp>

#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

using namespace std; // the reference headers rely on it (ORBmatcher.h:73,81)

namespace ORB_SLAM3
{
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

    class KeyFrame;
    class Frame;

    class GeometricCamera
    {
    public:
        virtual ~GeometricCamera() {}
        virtual Eigen::Vector2f project(const Eigen::Vector3f &v3D) = 0;
        virtual Eigen::Matrix3f toK_() = 0;
        virtual bool epipolarConstrain(GeometricCamera *pCamera2, const cv::KeyPoint &kp1, const cv::KeyPoint &kp2,
                                       const Eigen::Matrix3f &R12, const Eigen::Vector3f &t12, const float sigmaLevel,
                                       const float unc) = 0;
    };

    class Pinhole : public GeometricCamera
    {
    public:
        float mvParameters[4]; // fx fy cx cy
        Pinhole(float fx, float fy, float cx, float cy) : mvParameters{fx, fy, cx, cy} {}
        Eigen::Vector2f project(const Eigen::Vector3f &v3D) override; // Pinhole.cpp:64-71 (extracted)
        Eigen::Matrix3f toK_() override;                              // Pinhole.cpp:171-176 (extracted)
        bool epipolarConstrain(GeometricCamera *pCamera2, const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const Eigen::Matrix3f &R12,
                               const Eigen::Vector3f &t12, const float sigmaLevel, const float unc) override; // :189-219 (extracted)
        // harness helper (NOT reference text): F12 as Pinhole.cpp:194-197 forms it, for the entry points that take F12 as an input
        Eigen::Matrix3f fundamental(GeometricCamera *pCamera2, const Eigen::Matrix3f &R12, const Eigen::Vector3f &t12)
        {
            Eigen::Matrix3f t12x = Sophus::SO3f::hat(t12);
            Eigen::Matrix3f K1 = this->toK_();
            Eigen::Matrix3f K2 = pCamera2->toK_();
            return K1.transpose().inverse() * t12x * R12 * K2.inverse();
        }
    };

    class MapPoint;
    // (held, incoming) pairs of MapPoint::Replace calls, so that the harness can recover Fuse's per-point result
    inline std::vector<std::pair<MapPoint *, MapPoint *>> &g_replace_log()
    {
        static thread_local std::vector<std::pair<MapPoint *, MapPoint *>> log;
        return log;
    }

    // the stand-in map points are copied around by the harnesses; a copy gets a fresh (unlocked) mutex
    struct CopyableMutex : public std::mutex
    {
        CopyableMutex() {}
        CopyableMutex(const CopyableMutex &) {}
        CopyableMutex &operator=(const CopyableMutex &) { return *this; }
    };

    class MapPoint
    {
    public:
        // tracking scratch, MapPoint.h:166-177
        float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0;
        float mTrackProjXR = 0, mTrackProjYR = 0;
        bool mbTrackInView = false, mbTrackInViewR = false;
        int mnTrackScaleLevel = 0, mnTrackScaleLevelR = -1;
        float mTrackViewCos = 0, mTrackViewCosR = 0;

        // state behind the accessors
        bool bad_ = false;
        int nObs_ = 0;
        cv::Mat descriptor_;
        Eigen::Vector3f worldPos_, normal_;
        float mfMinDistance = 0, mfMaxDistance = 0;
        long unsigned int mnLastFrameSeen = 0; // MapPoint.h:184 (Tracking::SearchLocalPoints)
        // the two accessors INTEGRATION.md asks a maintainer to add to MapPoint.h for the device-side isInFrustum (PredictScale reads
        // the raw mfMaxDistance, MapPoint.cc:699,:726; the public getters only return 0.8 x / 1.2 x of the limits)
        float GetMinDistanceRaw() { return mfMinDistance; }
        float GetMaxDistanceRaw() { return mfMaxDistance; }
        CopyableMutex mMutexPos; // taken by the extracted bodies (MapPoint.cc:667, :698)
        std::map<KeyFrame *, std::tuple<int, int>> observations_;

        bool isBad() { return bad_; }
        int Observations() { return nObs_; }
        cv::Mat GetDescriptor() { return descriptor_.clone(); } // MapPoint.cc:540-544
        Eigen::Vector3f GetWorldPos() { return worldPos_; }
        Eigen::Vector3f GetNormal() { return normal_; }
        float GetMinDistanceInvariance();                          // MapPoint.cc:665-669 (extracted)
        float GetMaxDistanceInvariance();                          // MapPoint.cc:674-678 (extracted)
        int PredictScale(const float &currentDist, KeyFrame *pKF); // MapPoint.cc:695-713 (extracted)
        int PredictScale(const float &currentDist, Frame *pF);     // MapPoint.cc:722-738 (extracted)
        void Replace(MapPoint *pMP) { g_replace_log().push_back(std::make_pair(this, pMP)); }
        void AddObservation(KeyFrame *pKF, int idx) { observations_[pKF] = std::make_tuple(idx, -1); nObs_++; }
        bool IsInKeyFrame(KeyFrame *pKF) { return observations_.count(pKF) > 0; }
        std::tuple<int, int> GetIndexInKeyFrame(KeyFrame *pKF)
        {
            auto it = observations_.find(pKF);
            return it == observations_.end() ? std::make_tuple(-1, -1) : it->second;
        }
    };

    // shared by Frame and KeyFrame stubs: the feature arrays + the cell grid
    class FeatureSet
    {
    public:
        long unsigned int mnId = next_id_(); // Frame.h:278, KeyFrame.h:243: mnId = nNextId++ in the constructors, kept by copies
        static long unsigned int next_id_()
        {
            static long unsigned int n = 1;
            return n++;
        }
        int N = 0;
        std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
        std::vector<float> mvuRight;
        cv::Mat mDescriptors;
        DBoW2::BowVector mBowVec;
        DBoW2::FeatureVector mFeatVec;
        std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
        int mnScaleLevels = 0;
        float mfLogScaleFactor = 0;
        float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
        GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
        float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mb = 0;
        std::vector<float> mvInvScaleFactors;
        std::vector<MapPoint *> mvpMapPoints;
        std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
        std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
        Sophus::SE3f mTcw;
    };

    // what Frame::ComputeStereoMatches reads of the ORB extractors: the size of the pyramid's base image (Frame.cc:1143)
    struct ImageSizeStub { int rows = 0, cols = 0; };
    struct ORBextractorStub { std::vector<ImageSizeStub> mvImagePyramid; };
    // closes the extracted coarse stage of ComputeStereoMatches (oracle/extract_ref.py): bestIdxR is accepted iff bestDist <
    // thOrbDist (Frame.cc:1248); a left key point that `continue`s earlier keeps the defaults (-1, TH_HIGH)
#define ORB_ORACLE_STEREO_COARSE_HOOK(iL, bestDist, bestIdxR, thOrbDist) \
    do {                                                                   \
        stereo_best_dist_[iL] = (bestDist);                                \
        stereo_best_idx_[iL] = (bestDist) < (thOrbDist) ? (int)(bestIdxR) : -1; \
    } while (0)

    class Frame : public FeatureSet
    {
    public:
        int Nleft = -1;
        float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0; // static float in the reference (Frame.h:293-296)
        std::vector<bool> mvbOutlier;
        std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
        Sophus::SE3f mTrl;
        // pose pieces Frame::isInFrustum reads (Frame.h: mRcw, mtcw, mOw), kept in step with mTcw by SetPose
        Eigen::Matrix3f mRcw;
        Eigen::Vector3f mtcw, mOw;
        // stereo members of ComputeStereoMatches
        std::vector<float> mvDepth;
        cv::Mat mDescriptorsRight;
        ORBextractorStub *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
        std::vector<int> stereo_best_idx_, stereo_best_dist_; // filled by ORB_ORACLE_STEREO_COARSE_HOOK

        Sophus::SE3f GetPose() const { return mTcw; }
        Eigen::Vector3f GetCameraCenter() { return mOw; } // Frame.h:135-138
        Sophus::SE3f GetRelativePoseTrl() { return mTrl; }
        void SetPose(const Sophus::SE3f &Tcw) // Frame::SetPose + UpdatePoseMatrices (Frame.cc:600-660): member-wise
        {
            mTcw = Tcw;
            Sophus::SE3f Twc = mTcw.inverse();
            mOw = Twc.translation();
            mRcw = mTcw.rotationMatrix();
            mtcw = mTcw.translation();
        }

        void AssignFeaturesToGrid();                                    // Frame.cc:469-507 (extracted)
        bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY);   // Frame.cc:973-989 (extracted)
        vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1, const int maxLevel = -1,
                                         const bool bRight = false) const; // Frame.cc:868-962 (extracted)
        bool isInFrustum(MapPoint *pMP, float viewingCosLimit);         // Frame.cc:676-782 (extracted)
        bool isInFrustumChecks(MapPoint *, float, bool = false) { throw std::logic_error("stereo-fisheye path (Nleft != -1) is out of scope"); }
        void ComputeStereoMatches();                                    // Frame.cc:1117-1247, coarse stage (extracted + hook)
    };

    class KeyFrame : public FeatureSet
    {
    public:
        int NLeft = -1;
        int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0; // const int in the reference (KeyFrame.h:403-406)
        int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
        Sophus::SE3f mTrl;

        vector<MapPoint *> GetMapPointMatches() { return mvpMapPoints; }
        MapPoint *GetMapPoint(const size_t &idx) { return mvpMapPoints[idx]; }
        std::set<MapPoint *> GetMapPoints()
        {
            std::set<MapPoint *> s;
            for (MapPoint *p : mvpMapPoints)
                if (p && !p->isBad())
                    s.insert(p);
            return s;
        }
        void AddMapPoint(MapPoint *pMP, const size_t &idx) { mvpMapPoints[idx] = pMP; }
        Sophus::SE3f GetPose() { return mTcw; }
        Sophus::SE3f GetPoseInverse() { return mTcw.inverse(); }
        Eigen::Vector3f GetCameraCenter() { return mTcw.inverse().translation(); }
        Sophus::SE3f GetRightPose() { return mTrl * mTcw; }
        Sophus::SE3f GetRightPoseInverse() { return (mTrl * mTcw).inverse(); }
        Eigen::Vector3f GetRightCameraCenter() { return (mTrl * mTcw).inverse().translation(); }

        bool IsInImage(const float &x, const float &y) const; // KeyFrame.cc:910-913 (extracted)
        vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const bool bRight = false) const; // :859-907 (extracted)
        // grid of a key frame = the grid of the Frame it was made from (KeyFrame.cc:66-82: mGrid[i][j] = F.mGrid[i][j]); the bounds
        // become ints (KeyFrame.h:403-406 `const int mnMinX` initialised from the Frame's floats)
        void CopyGridFrom(const Frame &F)
        {
            mnMinX = F.mnMinX; mnMinY = F.mnMinY; mnMaxX = F.mnMaxX; mnMaxY = F.mnMaxY;
            for (int i = 0; i < mnGridCols; i++)
                for (int j = 0; j < mnGridRows; j++) mGrid[i][j] = F.mGrid[i][j];
        }
    };
} // namespace ORB_SLAM3

#endif
