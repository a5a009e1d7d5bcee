/*
 * Drop-in for the reference's QPPVM RT plugin (ref:include/QPPVM_RT_plugin/QPPVMPlugin.h:37-106): same class name,
 * namespace, virtual surface and registration symbol; the OpenSoT torque stack + QPOases_sot members are replaced
 * by one qppvm_handle of kind QPPVM_KIND_TORQUE (include/qppvm_b200.h).
 */
#ifndef __QPPVM_PLUGIN_H__
#define __QPPVM_PLUGIN_H__

#include <XCM/XBotControlPlugin.h>
#include <string>
#include <vector>
#include "../../include/qppvm_b200.h"

namespace demo {

class QPPVMPlugin : public XBot::XBotControlPlugin {
public:
    QPPVMPlugin();
    virtual ~QPPVMPlugin();
    virtual bool init_control_plugin(XBot::Handle::Ptr handle);
    virtual void on_start(double time);
    virtual void control_loop(double time, double period);
    virtual bool close();

    const std::vector<double>& last_record() const { return _record; }
    const std::vector<double>& last_output() const { return _out; }
    int last_status() const { return _status; }

private:
    double _start_time = 0.0;
    bool _set_ref = false;
    void syncFromMotorSide(XBot::RobotInterface::Ptr robot, XBot::ModelInterface::Ptr model);
    void sense();
    void QPPVMControl(const double time);
    void impedance_wrench(const std::string& link, const Eigen::Affine3d& ref, const Eigen::MatrixXd& J, double K, double D, double* F6) const;

    XBot::JointIdMap _jidmap;
    XBot::RobotInterface::Ptr _robot;
    XBot::ModelInterface::Ptr _model;
    XBot::MatLogger::Ptr _matlogger;
    Eigen::VectorXd _q, _dq, _q_ref, _q_home, _k, _d, _tau_d, _h;
    Eigen::VectorXd _tau_max, _tau_min, _tau_max_const, _tau_min_const;
    Eigen::Affine3d _ref_left, _ref_right, _start_pose;
    Eigen::MatrixXd _Jtmp, _M;
    double _Kc = 700.0, _Dc = 70.0, _Kj = 5.0, _Dj = 2.0;     // ref:src/QPPVMPlugin.cpp:105-106, 136-137, 148-149
    // The tasks / constraints the reference constructs next to its stack (ref:src/QPPVMPlugin.cpp:154-171, members
    // ref:include/QPPVM_RT_plugin/QPPVMPlugin.h:69-70,91-92).  QPPVM_PLUGIN_STACK="elbows", "joint_limits" or both
    // (comma-separated) stacks them the way the commented lines :169-171 / :177-178 do; unset = the shipped stack.
    int _stack_flags = 0;
    Eigen::VectorXd _q_max, _q_min, _k_jl, _d_jl;             // :120-123, gains k0 * 10, d0 * 20 (:170)
    Eigen::Affine3d _ref_elbow_left, _ref_elbow_right;        // the elbow tasks keep their construction-time reference
    double _Ke = 100.0, _De = 1.0;                            // CartesianImpedanceCtrl defaults: never overridden for the elbows

    qppvm_handle* _solver = nullptr;
    qppvm_layout _L;
    std::vector<double> _record, _out;
    int _status = 0;
};

}  // namespace demo
#endif
