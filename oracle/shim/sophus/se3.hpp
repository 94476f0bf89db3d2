// ORACLE BUILD SHIM (test infrastructure) -- matrix-form SE3f/Sim3f with the member functions
// the reference's ORBmatcher.cc calls.  Sophus is not installed in this image (and is an
// un-vendored, un-pinned system dependency of the reference: Examples/ROS/ORB_SLAM3/
// CMakeLists.txt:58), so its quaternion-form fp32 rounding is NOT reproduced here; every
// float that crosses into the hot path is fed to the reference build and the GPU alike.
#ifndef ORB_ORACLE_SHIM_SOPHUS_SE3
#define ORB_ORACLE_SHIM_SOPHUS_SE3
#include <Eigen/Core>

namespace Sophus
{
    template <class T>
    class SO3
    {
    public:
        static Eigen::Matrix3f hat(const Eigen::Vector3f &w)
        {
            Eigen::Matrix3f O;
            O(0, 1) = -w(2); O(0, 2) = w(1);
            O(1, 0) = w(2);  O(1, 2) = -w(0);
            O(2, 0) = -w(1); O(2, 1) = w(0);
            return O;
        }
    };
    typedef SO3<float> SO3f;

    template <class T>
    class SE3
    {
    public:
        SE3() : R_(Eigen::Matrix3f::Identity()) {}
        SE3(const Eigen::Matrix3f &R, const Eigen::Vector3f &t) : R_(R), t_(t) {}
        SE3 inverse() const
        {
            Eigen::Matrix3f Rt = R_.transpose();
            return SE3(Rt, -(Rt * t_));
        }
        const Eigen::Vector3f &translation() const { return t_; }
        Eigen::Matrix3f rotationMatrix() const { return R_; }
        SE3 operator*(const SE3 &o) const { return SE3(R_ * o.R_, R_ * o.t_ + t_); }
        Eigen::Vector3f operator*(const Eigen::Vector3f &p) const { return R_ * p + t_; }

    private:
        Eigen::Matrix3f R_;
        Eigen::Vector3f t_;
    };
    typedef SE3<float> SE3f;
} // namespace Sophus
#endif
