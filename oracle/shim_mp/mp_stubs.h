// ORACLE BUILD SHIM (test infrastructure) -- stand-ins for the classes the reference's src/MapPoint.cc touches, so that
// MapPoint.cc compiles UNMODIFIED (with the reference's real include/MapPoint.h) into oracle/_ref/libref_mappoint.so.
// Used to pin oracle_compute_distinctive_descriptors against MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:444-535).
// Only the members MapPoint.cc reads exist; nothing here is product code.
#ifndef ORB_ORACLE_SHIM_MP_STUBS_H
#define ORB_ORACLE_SHIM_MP_STUBS_H
#include <Eigen/Core>
#include <sophus/se3.hpp>
#include <opencv2/core/core.hpp>
#include <algorithm>
#include <climits>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>

using namespace std; // the reference's headers leak it (include/KeyFrame.h) and MapPoint.cc relies on that

template <class Archive> void serializeMatrix(Archive &, cv::Mat &, const unsigned int) {}

namespace ORB_SLAM3
{
    class MapPoint;
    class GeometricCamera;

    class Map
    {
    public:
        long unsigned int GetId() { return 0; }
        void EraseMapPoint(MapPoint *) {}
        std::mutex mMutexPointCreation;
    };

    class KeyFrame
    {
    public:
        long unsigned int mnId = 0, mnFrameId = 0;
        int NLeft = -1;
        GeometricCamera *mpCamera2 = nullptr;
        std::vector<float> mvuRight;
        std::vector<cv::KeyPoint> mvKeysUn, mvKeys, mvKeysRight;
        std::vector<float> mvScaleFactors;
        int mnScaleLevels = 8;
        float mfLogScaleFactor = 0.18232156f;
        cv::Mat mDescriptors;
        bool bad = false;
        bool isBad() { return bad; }
        Eigen::Vector3f GetCameraCenter() { return Eigen::Vector3f(); }
        Eigen::Vector3f GetRightCameraCenter() { return Eigen::Vector3f(); }
        void EraseMapPointMatch(int) {}
        void ReplaceMapPointMatch(int, MapPoint *) {}
        Map *GetMap() { return nullptr; }
    };

    class Frame
    {
    public:
        long unsigned int mnId = 0;
        int Nleft = -1;
        std::vector<cv::KeyPoint> mvKeysUn, mvKeys, mvKeysRight;
        std::vector<float> mvScaleFactors;
        int mnScaleLevels = 8;
        float mfLogScaleFactor = 0.18232156f;
        cv::Mat mDescriptors;
        Eigen::Vector3f GetCameraCenter() { return Eigen::Vector3f(); }
        Eigen::Matrix3f GetRwc() { return Eigen::Matrix3f::Identity(); }
        Sophus::SE3f GetRelativePoseTlr() { return Sophus::SE3f(); }
        Eigen::Vector3f GetOw() { return Eigen::Vector3f(); }
    };

    // the one ORBmatcher entry MapPoint.cc calls (:500); defined by the reference's ORBmatcher.cc in libref_orbmatcher.so
    class ORBmatcher
    {
    public:
        static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
    };
} // namespace ORB_SLAM3
#endif
