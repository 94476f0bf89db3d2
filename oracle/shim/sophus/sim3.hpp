// ORACLE BUILD SHIM (test infrastructure) -- see se3.hpp.
#ifndef ORB_ORACLE_SHIM_SOPHUS_SIM3
#define ORB_ORACLE_SHIM_SOPHUS_SIM3
#include <sophus/se3.hpp>

namespace Sophus
{
    template <class T>
    class Sim3
    {
    public:
        Sim3() : R_(Eigen::Matrix3f::Identity()), s_(1.f) {}
        Sim3(float s, const Eigen::Matrix3f &R, const Eigen::Vector3f &t) : R_(R), t_(t), s_(s) {}
        Eigen::Matrix3f rotationMatrix() const { return R_; }
        const Eigen::Vector3f &translation() const { return t_; }
        float scale() const { return s_; }
        Sim3 inverse() const
        {
            Eigen::Matrix3f Rt = R_.transpose();
            const float si = 1.0f / s_;
            return Sim3(si, Rt, -((Rt * t_) * si));
        }
        Eigen::Vector3f operator*(const Eigen::Vector3f &p) const { return (R_ * p) * s_ + t_; }

    private:
        Eigen::Matrix3f R_;
        Eigen::Vector3f t_;
        float s_;
    };
    typedef Sim3<float> Sim3f;
} // namespace Sophus
#endif
