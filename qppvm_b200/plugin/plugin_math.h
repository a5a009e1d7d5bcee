// Small host-side helpers shared by the two plugin drop-ins (pose errors as OpenSoT's cartesian_utils computes them).
#pragma once
#include <Eigen/Dense>
#include <cmath>

namespace qppvm_plugin {

struct Quat { double x, y, z, w; };

inline Quat quat_from_R(const Eigen::Matrix3d& R)
{
    Quat q;
    const double tr = R(0, 0) + R(1, 1) + R(2, 2);
    if (tr > 0) {
        const double s = std::sqrt(tr + 1.0) * 2;
        q.w = 0.25 * s; q.x = (R(2, 1) - R(1, 2)) / s; q.y = (R(0, 2) - R(2, 0)) / s; q.z = (R(1, 0) - R(0, 1)) / s;
    } else if (R(0, 0) > R(1, 1) && R(0, 0) > R(2, 2)) {
        const double s = std::sqrt(1.0 + R(0, 0) - R(1, 1) - R(2, 2)) * 2;
        q.w = (R(2, 1) - R(1, 2)) / s; q.x = 0.25 * s; q.y = (R(0, 1) + R(1, 0)) / s; q.z = (R(0, 2) + R(2, 0)) / s;
    } else if (R(1, 1) > R(2, 2)) {
        const double s = std::sqrt(1.0 + R(1, 1) - R(0, 0) - R(2, 2)) * 2;
        q.w = (R(0, 2) - R(2, 0)) / s; q.x = (R(0, 1) + R(1, 0)) / s; q.y = 0.25 * s; q.z = (R(1, 2) + R(2, 1)) / s;
    } else {
        const double s = std::sqrt(1.0 + R(2, 2) - R(0, 0) - R(1, 1)) * 2;
        q.w = (R(1, 0) - R(0, 1)) / s; q.x = (R(0, 2) + R(2, 0)) / s; q.y = (R(1, 2) + R(2, 1)) / s; q.z = 0.25 * s;
    }
    return q;
}

// Orientation error e_o = eta * eps_d - eta_d * eps - eps_d x eps  (desired d, actual without suffix); zero when equal.
inline void orientation_error(const Eigen::Matrix3d& Rd, const Eigen::Matrix3d& R, double* e)
{
    Quat qd = quat_from_R(Rd), q = quat_from_R(R);
    if (qd.w * q.w + qd.x * q.x + qd.y * q.y + qd.z * q.z < 0) { qd.w = -qd.w; qd.x = -qd.x; qd.y = -qd.y; qd.z = -qd.z; }
    e[0] = q.w * qd.x - qd.w * q.x - (qd.y * q.z - qd.z * q.y);
    e[1] = q.w * qd.y - qd.w * q.y - (qd.z * q.x - qd.x * q.z);
    e[2] = q.w * qd.z - qd.w * q.z - (qd.x * q.y - qd.y * q.x);
}

}  // namespace qppvm_plugin
