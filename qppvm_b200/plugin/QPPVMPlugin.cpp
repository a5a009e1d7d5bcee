/*
 * demo::QPPVMPlugin drop-in: ref:src/QPPVMPlugin.cpp:42-353 with `_autostack->update(_q)` (:226) and
 * `_solver->solve(_tau_d)` (:246) replaced by one call through the C-ABI to the B200 kernel (Torque kind).
 */
#include "QPPVMPlugin.h"
#include "plugin_math.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

REGISTER_XBOT_PLUGIN(QPPVMPlugin, demo::QPPVMPlugin)

using namespace demo;
static const char* kLeftEE = "arm1_7";      // ref:src/QPPVMPlugin.cpp:132
static const char* kRightEE = "arm2_7";     // ref:src/QPPVMPlugin.cpp:145
static const char* kLeftElbow = "arm1_4";   // ref:src/QPPVMPlugin.cpp:157
static const char* kRightElbow = "arm2_4";  // ref:src/QPPVMPlugin.cpp:164

QPPVMPlugin::QPPVMPlugin() {}
QPPVMPlugin::~QPPVMPlugin() { if (_solver) qppvm_destroy(_solver); }

bool QPPVMPlugin::init_control_plugin(XBot::Handle::Ptr handle)
{
    _matlogger = XBot::MatLogger::getLogger("/tmp/qppvm_log");                       // :44
    _set_ref = false;                                                                // :46
    _robot = handle->getRobotInterface();                                            // :48
    // the reference hard-codes its robot's YAML (:50-51); the drop-in asks the handle, as ForceAcc.cpp:43 does
    _model = XBot::ModelInterface::getModel(handle->getPathToConfigFile());
    if (!_model) return false;
    _model->initLog(_matlogger, 30000);                                              // :54
    _model->getEffortLimits(_tau_max_const);                                         // :56-58
    _tau_min_const = -_tau_max_const;
    const int n = _model->getJointNum();
    _tau_d.setZero(n);                                                               // :61-62
    _model->computeNonlinearTerm(_h);                                                // :65-67
    _tau_max = _tau_max_const - _h;
    _tau_min = _tau_min_const - _h;
    _model->getRobotState("home", _q_home);                                          // :69-72
    _model->setJointPosition(_q_home);
    Eigen::VectorXd zero; zero.setZero(n);
    _model->setJointVelocity(zero);
    _model->update();
    _q = _q_home;                                                                    // :74-75
    _q_ref = _q;
    _k.setZero(_robot->getJointNum());                                               // :77-96: wrists stay position-controlled
    _d.setZero(_robot->getJointNum());
    Eigen::VectorXd k0, d0;
    _robot->getStiffness(k0);
    _robot->getDamping(d0);
    const char* wrist[] = {"j_arm1_5", "j_arm1_6", "j_arm1_7", "j_arm2_5", "j_arm2_6", "j_arm2_7"};
    for (const char* jn : wrist) {
        const int idx = _robot->getDofIndex(jn);
        if (idx >= 0 && idx < _k.size()) { _k[idx] = k0[idx]; _d[idx] = d0[idx]; }
    }
    // TorqueLimits (:112) + JointImpedanceCtrl K=5 D=2 (:105-118) + two CartesianImpedanceCtrl K=700 D=70, rows 0..2
    // (:129-152) stacked as ((ee_right + ee_left) / joint_task) << torque_limits (:177-179), QPOases_sot(..., 1.0) (:188)
    qppvm_desc d;
    std::memset(&d, 0, sizeof(d));
    d.kind = QPPVM_KIND_TORQUE;
    d.n_a = n;
    d.n_contacts = 2;
    if (const char* e = std::getenv("QPPVM_PLUGIN_STACK")) {
        if (std::strstr(e, "elbows")) _stack_flags |= QPPVM_FLAG_ELBOW_TASKS;        // .../(_elbow_task_left + _elbow_task_right) (:178)
        if (std::strstr(e, "joint_limits")) _stack_flags |= QPPVM_FLAG_JOINT_LIMITS; // _joint_limits (:169-171)
    }
    if (_stack_flags & QPPVM_FLAG_JOINT_LIMITS) {
        _model->getJointLimits(_q_min, _q_max);                                      // :120-123
        _k_jl.setZero(n); _d_jl.setZero(n);
        for (int j = 0; j < n; ++j) {
            const double q_range = _q_max[j] - _q_min[j];
            _q_max[j] -= 0.1 * q_range;
            _q_min[j] += 0.1 * q_range;
            _k_jl[j] = 10.0 * k0[j];                                                 // setGains(k0*10, d0*20) (:170)
            _d_jl[j] = 20.0 * d0[j];
        }
    }
    if (_stack_flags & QPPVM_FLAG_ELBOW_TASKS) {                                     // :154-166: reference = pose at construction
        _model->getPose(kLeftElbow, _ref_elbow_left);
        _model->getPose(kRightElbow, _ref_elbow_right);
    }
    d.flags = _stack_flags;
    d.eps_regularisation = 1.0;
    d.n_reg_steps = 1;
    d.max_iter = 132;
    if (qppvm_get_layout(&d, &_L) != QPPVM_OK) return false;
    if (qppvm_create(&d, &_solver) != QPPVM_OK) {
        std::fprintf(stderr, "QPPVMPlugin: %s\n", qppvm_last_error(nullptr));
        return false;
    }
    _record.assign(_L.rec_doubles, 0.0);
    _out.assign(_L.out_bytes / 8, 0.0);
    return true;
}

void QPPVMPlugin::impedance_wrench(const std::string& link, const Eigen::Affine3d& ref, const Eigen::MatrixXd& J, double K, double D, double* F) const
{
    // CartesianImpedanceCtrl: F = K [e_pos; e_ori] + D (xdot_des - J qdot), xdot_des = 0
    Eigen::Affine3d T;
    _model->getPose(link, T);
    double e[6];
    for (int k = 0; k < 3; ++k) e[k] = ref.translation()[k] - T.translation()[k];
    qppvm_plugin::orientation_error(ref.linear(), T.linear(), e + 3);
    for (int r = 0; r < 6; ++r) {
        double v = 0.0;
        for (int j = 0; j < J.cols(); ++j) v += J(r, j) * _dq[j];
        F[r] = K * e[r] - D * v;
    }
}

void QPPVMPlugin::QPPVMControl(const double time)
{
    _tau_max = _tau_max_const - _h;                                                  // :203-205 (the kernel applies the same
    _tau_min = _tau_min_const - _h;                                                  //  shift from the constants + h)
    if (_set_ref) {                                                                  // :217-223
        _ref_left = _start_pose;
        _ref_left.translation()[1] = _start_pose.translation()[1] + 0.15 * std::sin(time - _start_time);
        _ref_left.translation()[2] = _start_pose.translation()[2] + 0.15 * (1.0 - std::cos(time - _start_time));
    }
    /* _autostack->update(_q) (:226): pack what the three tasks and the torque limits are built from */
    const int n = _L.n_x;
    double* rec = _record.data();
    const char* ee[2] = {kRightEE, kLeftEE};                                         // stack order ee_right + ee_left (:177)
    const Eigen::Affine3d* ref[2] = {&_ref_right, &_ref_left};
    for (int t = 0; t < 2; ++t) {
        _model->getJacobian(ee[t], _Jtmp);
        for (int r = 0; r < 6; ++r)
            for (int j = 0; j < n; ++j) rec[_L.off_jc + (t * 6 + r) * n + j] = _Jtmp(r, j);
        impedance_wrench(ee[t], *ref[t], _Jtmp, _Kc, _Dc, rec + _L.off_fee + 6 * t);
    }
    _model->getInertiaMatrix(_M);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) rec[_L.off_M + i * (i + 1) / 2 + j] = _M(i, j);
    for (int j = 0; j < n; ++j) {
        rec[_L.off_h + j] = _h[j];
        rec[_L.off_tauj + j] = _Kj * (_q_ref[j] - _q[j]) - _Dj * _dq[j];
        rec[_L.off_taulim + j] = _tau_min_const[j];
        rec[_L.off_taulim + n + j] = _tau_max_const[j];
    }
    if (_stack_flags & QPPVM_FLAG_JOINT_LIMITS)                                      // torque::JointLimits::update
        for (int j = 0; j < n; ++j) {
            rec[_L.off_jlim + j] = _k_jl[j] * (_q_min[j] - _q[j]) - _d_jl[j] * _dq[j];
            rec[_L.off_jlim + n + j] = _k_jl[j] * (_q_max[j] - _q[j]) - _d_jl[j] * _dq[j];
        }
    if (_stack_flags & QPPVM_FLAG_ELBOW_TASKS) {
        const char* el[2] = {kLeftElbow, kRightElbow};                               // stack order elbow_left + elbow_right (:178)
        const Eigen::Affine3d* eref[2] = {&_ref_elbow_left, &_ref_elbow_right};
        for (int t = 0; t < 2; ++t) {
            _model->getJacobian(el[t], _Jtmp);
            for (int r = 0; r < 6; ++r)
                for (int j = 0; j < n; ++j) rec[_L.off_jelbow + (t * 6 + r) * n + j] = _Jtmp(r, j);
            impedance_wrench(el[t], *eref[t], _Jtmp, _Ke, _De, rec + _L.off_felbow + 6 * t);
        }
    }
    /* _solver->solve(_tau_d) (:246) */
    int rc = qppvm_solve_one(_solver, _record.data(), _out.data());
    qppvm_trailer tr;
    std::memcpy(&tr, _out.data() + _L.n_x + _L.n_a, sizeof(tr));
    _status = rc != QPPVM_OK ? -rc : tr.status;
    if (rc != QPPVM_OK || tr.status != QPPVM_STATUS_OK) {
        _tau_d.setZero(_tau_d.size());                                               // :247-248
        std::cout << "SOLVER ERROR!" << std::endl;
    } else
        for (int j = 0; j < n; ++j) _tau_d[j] = _out[j];
    _matlogger->add("tau_qp", _tau_d);                                               // :254
    _tau_d = _tau_d + _h;                                                            // :256 (== the kernel's tau output)
    _matlogger->add("tau_desired", _tau_d);                                          // :258
}

void QPPVMPlugin::on_start(double time)
{
    sense();                                                                         // :263-264
    _model->computeNonlinearTerm(_h);
    _start_time = time;                                                              // :266-269
    _robot->setStiffness(_k);
    _robot->setDamping(_d);
    _robot->move();
    _model->getPose(kLeftEE, _ref_left);                                             // :271-279: references = current poses
    _model->getPose(kRightEE, _ref_right);
    _q_ref = _q;                                                                     // :280
    _model->getPose(kLeftEE, _start_pose);                                           // :287
}

void QPPVMPlugin::control_loop(double time, double period)
{
    sense();                                                                         // :311-312
    _model->computeNonlinearTerm(_h);
    QPPVMControl(time);                                                              // :315
    _model->setJointEffort(_tau_d);                                                  // :318-320
    _robot->setReferenceFrom(*_model, XBot::Sync::Effort);
    _matlogger->add("time_matlogger", time);                                         // :322
    _model->log(_matlogger, time);                                                   // :325
    _robot->move();                                                                  // :328
    (void)period;
}

void QPPVMPlugin::sense()
{
    syncFromMotorSide(_robot, _model);                                               // :333-335
    _model->getJointPosition(_q);
    _model->getJointVelocity(_dq);
}

bool QPPVMPlugin::close()
{
    _matlogger->flush();                                                             // :341
    return true;                                                                     // (the reference forgets to return: :339-342)
}

void QPPVMPlugin::syncFromMotorSide(XBot::RobotInterface::Ptr robot, XBot::ModelInterface::Ptr model)
{
    // ref:src/QPPVMPlugin.cpp:344-353: motor-side position / velocity maps into the model, then update().
    // The JointIdMap overloads of ModelInterface are folded into syncFrom() in the shim.
    robot->getMotorPosition(_jidmap);
    robot->getMotorVelocity(_jidmap);
    model->syncFrom(*robot);
    model->update();
}

// Boundary-test accessor (not in the reference surface): 0 = last record, 1 = last raw output, 2 = status.
extern "C" const double* qppvm_plugin_debug(XBot::XBotControlPlugin* p, int what, int* n)
{
    static double st;
    auto* s = static_cast<demo::QPPVMPlugin*>(p);
    if (what == 0) { *n = (int)s->last_record().size(); return s->last_record().data(); }
    if (what == 1) { *n = (int)s->last_output().size(); return s->last_output().data(); }
    st = s->last_status(); *n = 1; return &st;
}
