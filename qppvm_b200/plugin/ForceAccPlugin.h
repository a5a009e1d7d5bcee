/*
 * Drop-in for the reference's ForceAcc RT plugin (ref:include/ForceAccPlugin/ForceAcc.h:36-88): same class name,
 * namespace, virtual surface (init_control_plugin / on_start / on_stop / control_loop / close) and registration
 * symbol; the OpenSoT stack + QPOases_sot members are replaced by one qppvm_handle (include/qppvm_b200.h).
 */
#ifndef ForceAccExample_PLUGIN_H_
#define ForceAccExample_PLUGIN_H_

#include <XCM/XBotControlPlugin.h>
#include <string>
#include <vector>
#include "../../include/qppvm_b200.h"

namespace XBotPlugin {

class ForceAccExample : public XBot::XBotControlPlugin {
public:
    virtual bool init_control_plugin(XBot::Handle::Ptr handle);
    virtual bool close();
    virtual void on_start(double time);
    virtual void on_stop(double time) {}
    virtual ~ForceAccExample();

    // not part of the reference surface: lets the boundary test read what was handed to / returned by the solver
    const std::vector<double>& last_record() const { return _record; }
    const std::vector<double>& last_output() const { return _out; }
    int last_status() const { return _status; }

protected:
    virtual void control_loop(double time, double period);

private:
    struct CartesianRef {           // what OpenSoT::tasks::acceleration::Cartesian keeps between ticks
        Eigen::Affine3d pose;       // reference pose (resetReference(): pose at on_start, zero twist / acceleration)
    };
    void sync_model();
    void cartesian_rhs(const std::string& link, const CartesianRef& ref, const Eigen::MatrixXd& J, double* rhs6) const;
    void build_record();

    XBot::RobotInterface::Ptr _robot;
    XBot::ModelInterface::Ptr _model;
    XBot::ImuSensor::ConstPtr _imu;
    XBot::SharedObject<Eigen::Vector3d> _sh_fb_pos, _sh_fb_vel;
    XBot::MatLogger::Ptr _logger;
    double _start_time = 0.0;
    Eigen::VectorXd _k, _d, _q, _qdot, _qhome, _tau, _tau_c, _qddot_value, _x, _h;
    Eigen::Vector3d _initial_com;
    std::vector<std::string> _contact_links;
    std::vector<Eigen::VectorXd> _wrench_value;
    std::vector<CartesianRef> _feet_ref;
    CartesianRef _waist_ref;
    Eigen::MatrixXd _Jtmp, _M;
    double _lambda = 100.0, _lambda2 = 20.0;       // OpenSoT acceleration-task defaults
    // _com_task (OpenSoT tasks::force::CoM, ref:src/ForceAcc.cpp:103; member ref:include/ForceAccPlugin/ForceAcc.h): the
    // reference constructs it and never stacks it.  FORCEACC_PLUGIN_STACK=com adds it to level 1
    // (_postural_task + feet_cart_aggr + _com_task); unset = the shipped stack.
    bool _stack_com = false;
    Eigen::Vector3d _com_ref;                      // centre of mass at on_start (the task's construction-time reference)
    void com_task_rows(double* A_com, double* b_com) const;

    qppvm_handle* _solver = nullptr;
    qppvm_layout _L;
    std::vector<double> _record, _out;
    int _status = 0;
};

}  // namespace XBotPlugin
#endif
