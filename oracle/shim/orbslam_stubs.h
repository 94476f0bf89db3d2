// ORACLE BUILD SHIM (test infrastructure) -- stand-ins for ORB_SLAM3::MapPoint, KeyFrame,
// Frame and GeometricCamera exposing exactly the members the reference's ORBmatcher.cc
// touches, so that /root/reference/src/ORBmatcher.cc + include/ORBmatcher.h compile
// UNMODIFIED into oracle/_ref/libref_orbmatcher.so (oracle/Makefile).  The real headers
// pull in OpenCV/Eigen/Sophus/Boost/g2o/Pangolin, none of which exist in this image.
//
// The few helper functions ORBmatcher.cc calls INTO these classes are restated here,
// citing the reference lines they follow:
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea   src/Frame.cc:469-507, 973-989, 868-962
//   KeyFrame::GetFeaturesInArea / IsInImage                       src/KeyFrame.cc:859-913
//   MapPoint::PredictScale / Get{Min,Max}DistanceInvariance       src/MapPoint.cc:665-738
//   Pinhole::project / epipolarConstrain / toK_                   src/CameraModels/Pinhole.cpp:64-71, 171-219
// Nothing here is product code.
#ifndef ORB_ORACLE_SHIM_ORBSLAM_STUBS_H
#define ORB_ORACLE_SHIM_ORBSLAM_STUBS_H

#include <cassert>
#include <cmath>
#include <map>
#include <set>
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
        // Pinhole.cpp:64-71
        Eigen::Vector2f project(const Eigen::Vector3f &v3D) override
        {
            Eigen::Vector2f res;
            res[0] = mvParameters[0] * v3D[0] / v3D[2] + mvParameters[2];
            res[1] = mvParameters[1] * v3D[1] / v3D[2] + mvParameters[3];
            return res;
        }
        // Pinhole.cpp:171-176
        Eigen::Matrix3f toK_() override
        {
            Eigen::Matrix3f K;
            K(0, 0) = mvParameters[0]; K(0, 2) = mvParameters[2];
            K(1, 1) = mvParameters[1]; K(1, 2) = mvParameters[3];
            K(2, 2) = 1.f;
            return K;
        }
        // Pinhole.cpp:194-197
        Eigen::Matrix3f fundamental(GeometricCamera *pCamera2, const Eigen::Matrix3f &R12, const Eigen::Vector3f &t12)
        {
            Eigen::Matrix3f t12x = Sophus::SO3f::hat(t12);
            Eigen::Matrix3f K1 = this->toK_();
            Eigen::Matrix3f K2 = pCamera2->toK_();
            return K1.transpose().inverse() * t12x * R12 * K2.inverse();
        }
        // Pinhole.cpp:189-219
        bool epipolarConstrain(GeometricCamera *pCamera2, const cv::KeyPoint &kp1, const cv::KeyPoint &kp2,
                               const Eigen::Matrix3f &R12, const Eigen::Vector3f &t12, const float sigmaLevel,
                               const float unc) override
        {
            (void)sigmaLevel;
            Eigen::Matrix3f F12 = fundamental(pCamera2, R12, t12);
            const float a = kp1.pt.x * F12(0, 0) + kp1.pt.y * F12(1, 0) + F12(2, 0);
            const float b = kp1.pt.x * F12(0, 1) + kp1.pt.y * F12(1, 1) + F12(2, 1);
            const float c = kp1.pt.x * F12(0, 2) + kp1.pt.y * F12(1, 2) + F12(2, 2);
            const float num = a * kp2.pt.x + b * kp2.pt.y + c;
            const float den = a * a + b * b;
            if (den == 0)
                return false;
            const float dsqr = num * num / den;
            return dsqr < 3.84 * unc;
        }
    };

    class MapPoint;
    // (held, incoming) pairs of MapPoint::Replace calls, so that the harness can recover Fuse's per-point result
    inline std::vector<std::pair<MapPoint *, MapPoint *>> &g_replace_log()
    {
        static thread_local std::vector<std::pair<MapPoint *, MapPoint *>> log;
        return log;
    }

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
        std::map<KeyFrame *, std::tuple<int, int>> observations_;

        bool isBad() { return bad_; }
        int Observations() { return nObs_; }
        cv::Mat GetDescriptor() { return descriptor_.clone(); } // MapPoint.cc:540-544
        Eigen::Vector3f GetWorldPos() { return worldPos_; }
        Eigen::Vector3f GetNormal() { return normal_; }
        float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; } // MapPoint.cc:665-669
        float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; } // MapPoint.cc:674-678
        int PredictScale(const float &currentDist, KeyFrame *pKF);        // MapPoint.cc:695-713
        int PredictScale(const float &currentDist, Frame *pF);            // MapPoint.cc:722-738
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
        std::vector<MapPoint *> mvpMapPoints;
        std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
        std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
        Sophus::SE3f mTcw;

        // Frame.cc:973-989
        bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY, float minX, float minY)
        {
            posX = round((kp.pt.x - minX) * mfGridElementWidthInv);
            posY = round((kp.pt.y - minY) * mfGridElementHeightInv);
            if (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS)
                return false;
            return true;
        }
        // Frame.cc:469-507
        void AssignFeaturesToGrid(float minX, float minY)
        {
            for (int i = 0; i < N; i++)
            {
                const cv::KeyPoint &kp = mvKeysUn[i];
                int nGridPosX, nGridPosY;
                if (PosInGrid(kp, nGridPosX, nGridPosY, minX, minY))
                    mGrid[nGridPosX][nGridPosY].push_back(i);
            }
        }
    };

    class Frame : public FeatureSet
    {
    public:
        int Nleft = -1;
        float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0; // static float in the reference (Frame.h:293-296)
        std::vector<bool> mvbOutlier;
        std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
        Sophus::SE3f mTrl;

        Sophus::SE3f GetPose() const { return mTcw; }
        Sophus::SE3f GetRelativePoseTrl() { return mTrl; }

        // Frame.cc:868-962
        vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1,
                                         const int maxLevel = -1, const bool bRight = false) const
        {
            vector<size_t> vIndices;
            vIndices.reserve(N);
            float factorX = r;
            float factorY = r;
            const int nMinCellX = max(0, (int)floor((x - mnMinX - factorX) * mfGridElementWidthInv));
            if (nMinCellX >= FRAME_GRID_COLS)
                return vIndices;
            const int nMaxCellX = min((int)FRAME_GRID_COLS - 1, (int)ceil((x - mnMinX + factorX) * mfGridElementWidthInv));
            if (nMaxCellX < 0)
                return vIndices;
            const int nMinCellY = max(0, (int)floor((y - mnMinY - factorY) * mfGridElementHeightInv));
            if (nMinCellY >= FRAME_GRID_ROWS)
                return vIndices;
            const int nMaxCellY = min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - mnMinY + factorY) * mfGridElementHeightInv));
            if (nMaxCellY < 0)
                return vIndices;
            const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
            for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            {
                for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
                {
                    const vector<size_t> &vCell = (!bRight) ? mGrid[ix][iy] : mGridRight[ix][iy];
                    if (vCell.empty())
                        continue;
                    for (size_t j = 0, jend = vCell.size(); j < jend; j++)
                    {
                        const cv::KeyPoint &kpUn = mvKeysUn[vCell[j]];
                        if (bCheckLevels)
                        {
                            if (kpUn.octave < minLevel)
                                continue;
                            if (maxLevel >= 0)
                                if (kpUn.octave > maxLevel)
                                    continue;
                        }
                        const float distx = kpUn.pt.x - x;
                        const float disty = kpUn.pt.y - y;
                        if (fabs(distx) < factorX && fabs(disty) < factorY)
                            vIndices.push_back(vCell[j]);
                    }
                }
            }
            return vIndices;
        }
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

        // KeyFrame.cc:910-913
        bool IsInImage(const float &x, const float &y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }

        // KeyFrame.cc:859-907
        vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const bool bRight = false) const
        {
            vector<size_t> vIndices;
            vIndices.reserve(N);
            float factorX = r;
            float factorY = r;
            const int nMinCellX = max(0, (int)floor((x - mnMinX - factorX) * mfGridElementWidthInv));
            if (nMinCellX >= mnGridCols)
                return vIndices;
            const int nMaxCellX = min((int)mnGridCols - 1, (int)ceil((x - mnMinX + factorX) * mfGridElementWidthInv));
            if (nMaxCellX < 0)
                return vIndices;
            const int nMinCellY = max(0, (int)floor((y - mnMinY - factorY) * mfGridElementHeightInv));
            if (nMinCellY >= mnGridRows)
                return vIndices;
            const int nMaxCellY = min((int)mnGridRows - 1, (int)ceil((y - mnMinY + factorY) * mfGridElementHeightInv));
            if (nMaxCellY < 0)
                return vIndices;
            for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            {
                for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
                {
                    const vector<size_t> &vCell = (!bRight) ? mGrid[ix][iy] : mGridRight[ix][iy];
                    for (size_t j = 0, jend = vCell.size(); j < jend; j++)
                    {
                        const cv::KeyPoint &kpUn = mvKeysUn[vCell[j]];
                        const float distx = kpUn.pt.x - x;
                        const float disty = kpUn.pt.y - y;
                        if (fabs(distx) < r && fabs(disty) < r)
                            vIndices.push_back(vCell[j]);
                    }
                }
            }
            return vIndices;
        }
    };

    // MapPoint.cc:695-713
    inline int MapPoint::PredictScale(const float &currentDist, KeyFrame *pKF)
    {
        float ratio = mfMaxDistance / currentDist;
        int nScale = ceil(log(ratio) / pKF->mfLogScaleFactor);
        if (nScale < 0)
            nScale = 0;
        else if (nScale >= pKF->mnScaleLevels)
            nScale = pKF->mnScaleLevels - 1;
        return nScale;
    }
    // MapPoint.cc:722-738
    inline int MapPoint::PredictScale(const float &currentDist, Frame *pF)
    {
        float ratio = mfMaxDistance / currentDist;
        int nScale = ceil(log(ratio) / pF->mfLogScaleFactor);
        if (nScale < 0)
            nScale = 0;
        else if (nScale >= pF->mnScaleLevels)
            nScale = pF->mnScaleLevels - 1;
        return nScale;
    }
} // namespace ORB_SLAM3

#endif
