/*
 * ForceAccExample drop-in: the per-tick sequence of ref:src/ForceAcc.cpp:167-253 with the OpenSoT stack update and
 * the QPOases_sot solve (ref:src/ForceAcc.cpp:184-193) replaced by one call through the C-ABI to the B200 kernel.
 * Host code only packs the record (what `_autostack->update()` assembles inside OpenSoT) and applies the outputs.
 */
#include "ForceAccPlugin.h"
#include "plugin_math.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>

REGISTER_XBOT_PLUGIN_(XBotPlugin::ForceAccExample)

using XBot::Logger;
static const std::string floating_base_name = "pelvis";            // ref:src/ForceAcc.cpp:29

bool XBotPlugin::ForceAccExample::init_control_plugin(XBot::Handle::Ptr handle)
{
    _robot = handle->getRobotInterface();                                            // :33
    _logger = XBot::MatLogger::getLogger("/tmp/opensot_force_acc_example");          // :34
    _robot->getStiffness(_k);                                                        // :36-39
    _robot->getDamping(_d);
    _k /= 16;
    _d /= 4;
    _imu = _robot->getImu().begin()->second;                                         // :41
    _model = XBot::ModelInterface::getModel(handle->getPathToConfigFile());          // :43
    if (!_model) return false;
    _model->getRobotState("home", _qhome);                                           // :45-48
    _model->setJointPosition(_qhome);
    _model->update();
    _model->initLog(_logger, 10000);                                                 // :50
    _sh_fb_pos = handle->getSharedMemory()->getSharedObject<Eigen::Vector3d>("/gazebo/floating_base_position");   // :52-55
    _sh_fb_vel = handle->getSharedMemory()->getSharedObject<Eigen::Vector3d>("/gazebo/floating_base_velocity");
    _sh_fb_pos.set(Eigen::Vector3d::Zero());
    _sh_fb_vel.set(Eigen::Vector3d::Zero());
    _contact_links = {"foot_fl", "foot_fr", "foot_hr", "foot_hl"};                   // :58
    _wrench_value.assign(_contact_links.size(), Eigen::VectorXd(6));                 // :61
    _feet_ref.resize(_contact_links.size());

    // variables "qddot"(n_v) + 3 per contact (:63-70), stack waist / (postural + feet) << dyn_feas << wrench bounds
    // (:131-133), QPOases_sot(..., 1e4) (:135-137)  ->  one solver handle for that problem shape
    qppvm_desc d;
    std::memset(&d, 0, sizeof(d));
    d.kind = QPPVM_KIND_FORCEACC;
    d.n_a = _model->getJointNum() - 6;
    d.n_contacts = (int)_contact_links.size();
    if (const char* e = std::getenv("FORCEACC_PLUGIN_STACK")) _stack_com = std::strstr(e, "com") != nullptr;
    d.flags = _stack_com ? QPPVM_FLAG_COM_TASK : 0;
    d.eps_regularisation = 1e4;
    d.n_reg_steps = 1;
    d.max_iter = 132;
    d.device = 0;
    if (qppvm_get_layout(&d, &_L) != QPPVM_OK) return false;
    if (qppvm_create(&d, &_solver) != QPPVM_OK) {
        std::fprintf(stderr, "ForceAccExample: %s\n", qppvm_last_error(nullptr));
        return false;
    }
    _record.assign(_L.rec_doubles, 0.0);
    _out.assign(_L.out_bytes / 8, 0.0);
    _x.setZero(_L.n_x);
    _qddot_value.setZero(_L.n_v);
    _tau.setZero(_L.n_v);
    _tau_c.setZero(_L.n_v);
    return true;
}

XBotPlugin::ForceAccExample::~ForceAccExample()
{
    if (_solver) qppvm_destroy(_solver);
}

bool XBotPlugin::ForceAccExample::close()
{
    _logger->flush();                                                                // ref:include/ForceAccPlugin/ForceAcc.h:43
    return true;
}

void XBotPlugin::ForceAccExample::on_start(double time)
{
    _start_time = time;                                                              // :153-154
    _model->getJointPosition(_q);
    _qdot.setZero(_q.size());
    sync_model();                                                                    // :156
    for (size_t i = 0; i < _contact_links.size(); ++i)                               // resetReference(): :158-161
        _model->getPose(_contact_links[i], _feet_ref[i].pose);
    _model->getPose(floating_base_name, _waist_ref.pose);                            // :162
    _model->getPointPosition(floating_base_name, Eigen::Vector3d::Zero(), _initial_com);   // :164
    if (_stack_com) {                                                                // the CoM task's reference: where the CoM is now
        _model->getInertiaMatrix(_M);
        Eigen::Affine3d Tb;
        _model->getPose(floating_base_name, Tb);
        const double m = _M(0, 0);
        _com_ref[0] = Tb.translation()[0] + _M(1, 5) / m;
        _com_ref[1] = Tb.translation()[1] + _M(2, 3) / m;
        _com_ref[2] = Tb.translation()[2] + _M(0, 4) / m;
    }
}

// Centroidal dynamics on the contact forces (OpenSoT tasks::force::CoM, ref:src/ForceAcc.cpp:103):
//   sum f_i = m (lambda (c_ref - c) - lambda2 cdot) + m g z,     sum (p_i - c) x f_i = -lambda2 L_c.
// Total mass, centre of mass and centroidal momentum are read from the floating-base block of the joint-space inertia
// matrix (M[0:3,0:3] = m I, M[0:3,3:6] = -m [c - p_base]x in the world-aligned convention of the model) and M[0:6] qdot.
void XBotPlugin::ForceAccExample::com_task_rows(double* A, double* b) const
{
    const int nv = _L.n_v, c = _L.n_c, nf = 3 * c;
    Eigen::Affine3d Tb, Tc;
    _model->getPose(floating_base_name, Tb);
    const double m = _M(0, 0);
    const double d[3] = {_M(1, 5) / m, _M(2, 3) / m, _M(0, 4) / m};                  // c - p_base
    double mom[6];
    for (int r = 0; r < 6; ++r) { double v = 0.0; for (int j = 0; j < nv; ++j) v += _M(r, j) * _qdot[j]; mom[r] = v; }
    const double Lc[3] = {mom[3] - (d[1] * mom[2] - d[2] * mom[1]), mom[4] - (d[2] * mom[0] - d[0] * mom[2]), mom[5] - (d[0] * mom[1] - d[1] * mom[0])};
    for (int e = 0; e < 6 * nf; ++e) A[e] = 0.0;
    for (int i = 0; i < c; ++i) {
        _model->getPose(_contact_links[i], Tc);
        double r[3];
        for (int k = 0; k < 3; ++k) r[k] = Tc.translation()[k] - (Tb.translation()[k] + d[k]);
        for (int k = 0; k < 3; ++k) A[k * nf + 3 * i + k] = 1.0;
        A[3 * nf + 3 * i + 1] = -r[2]; A[3 * nf + 3 * i + 2] = r[1];                 // [r]x
        A[4 * nf + 3 * i + 0] = r[2];  A[4 * nf + 3 * i + 2] = -r[0];
        A[5 * nf + 3 * i + 0] = -r[1]; A[5 * nf + 3 * i + 1] = r[0];
    }
    for (int k = 0; k < 3; ++k) {
        const double ck = Tb.translation()[k] + d[k];
        b[k] = m * (_lambda * (_com_ref[k] - ck) - _lambda2 * mom[k] / m) + (k == 2 ? m * 9.81 : 0.0);
        b[3 + k] = -_lambda2 * Lc[k];
    }
}

void XBotPlugin::ForceAccExample::cartesian_rhs(const std::string& link, const CartesianRef& ref,
                                                const Eigen::MatrixXd& J, double* rhs) const
{
    // acceleration::Cartesian: b + Jdot qdot = a_ref + lambda2 (v_ref - J qdot) + lambda e_pose, a_ref = v_ref = 0
    Eigen::Affine3d T;
    _model->getPose(link, T);
    double e[6];
    for (int k = 0; k < 3; ++k) e[k] = ref.pose.translation()[k] - T.translation()[k];
    qppvm_plugin::orientation_error(ref.pose.linear(), T.linear(), e + 3);
    for (int r = 0; r < 6; ++r) {
        double v = 0.0;
        for (int j = 0; j < J.cols(); ++j) v += J(r, j) * _qdot[j];
        rhs[r] = _lambda * e[r] - _lambda2 * v;
    }
}

void XBotPlugin::ForceAccExample::build_record()
{
    const int nv = _L.n_v, c = _L.n_c;
    double* rec = _record.data();
    _model->getJointPosition(_q);
    _model->getJointVelocity(_qdot);
    // waist task (:118-122) with the position reference of :181
    _model->getJacobian(floating_base_name, _Jtmp);
    for (int r = 0; r < 6; ++r)
        for (int j = 0; j < nv; ++j) rec[_L.off_jwaist + r * nv + j] = _Jtmp(r, j);
    cartesian_rhs(floating_base_name, _waist_ref, _Jtmp, rec + _L.off_rhs);
    Eigen::Vector6d jd;
    _model->computeJdotQdot(floating_base_name, Eigen::Vector3d::Zero(), jd);
    for (int r = 0; r < 6; ++r) rec[_L.off_jdqd + r] = jd[r];
    // contact-link Cartesian tasks (:83-89) and wrench bounds (:74-76, 91-95)
    for (int i = 0; i < c; ++i) {
        _model->getJacobian(_contact_links[i], _Jtmp);
        for (int r = 0; r < 6; ++r)
            for (int j = 0; j < nv; ++j) rec[_L.off_jc + (i * 6 + r) * nv + j] = _Jtmp(r, j);
        cartesian_rhs(_contact_links[i], _feet_ref[i], _Jtmp, rec + _L.off_rhs + 6 * (1 + i));
        _model->computeJdotQdot(_contact_links[i], Eigen::Vector3d::Zero(), jd);
        for (int r = 0; r < 6; ++r) rec[_L.off_jdqd + 6 * (1 + i) + r] = jd[r];
        const double lb[3] = {-1000, -1000, 10}, ub[3] = {1000, 1000, 1000};
        for (int k = 0; k < 3; ++k) { rec[_L.off_fbox + 6 * i + k] = lb[k]; rec[_L.off_fbox + 6 * i + 3 + k] = ub[k]; }
    }
    // postural (:105-107): qddot = lambda2 (0 - qdot) + lambda (q_home - q)
    for (int j = 0; j < nv; ++j) rec[_L.off_rhs + 6 * (1 + c) + j] = _lambda * (_qhome[j] - _q[j]) - _lambda2 * _qdot[j];
    // dynamic feasibility (:109-114) and the inverse-dynamics recovery (:206-219) need M and h
    _model->getInertiaMatrix(_M);
    for (int i = 0; i < nv; ++i)
        for (int j = 0; j <= i; ++j) rec[_L.off_M + i * (i + 1) / 2 + j] = _M(i, j);
    _model->computeNonlinearTerm(_h);
    for (int j = 0; j < nv; ++j) rec[_L.off_h + j] = _h[j];
    if (_stack_com) com_task_rows(rec + _L.off_com, rec + _L.off_com + 6 * 3 * c);
}

void XBotPlugin::ForceAccExample::control_loop(double time, double period)
{
    const bool enable_torque_ctrl = true;                                            // :169-170
    const bool enable_feedback = true;
    if (enable_feedback) sync_model();                                               // :172-178

    /* Set reference */                                                              // :181
    for (int k = 0; k < 3; ++k) _waist_ref.pose.translation()[k] = _initial_com[k] - (k == 2 ? 0.1 : 0.0);

    /* Update stack + solve QP: one call through the C-ABI */                        // :184-193
    build_record();
    _x.setZero(_x.size());
    int rc = qppvm_solve_one(_solver, _record.data(), _out.data());
    qppvm_trailer tr;
    std::memcpy(&tr, _out.data() + _L.n_x + _L.n_a, sizeof(tr));
    _status = rc != QPPVM_OK ? -rc : tr.status;
    if (rc != QPPVM_OK || tr.status != QPPVM_STATUS_OK) {
        Logger::error("Unable to solve!!!");                                         // :189-193: nothing is commanded
        return;
    }

    /* Retrieve values */                                                            // :196-201
    const int nv = _L.n_v;
    for (int j = 0; j < _L.n_x; ++j) _x[j] = _out[j];
    for (int j = 0; j < nv; ++j) _qddot_value[j] = _x[j];
    for (size_t i = 0; i < _contact_links.size(); ++i) {
        for (int k = 0; k < 3; ++k) { _wrench_value[i][k] = _x[nv + 3 * i + k]; _wrench_value[i][3 + k] = 0.0; }
        _logger->add(_contact_links[i] + "_wrench", _wrench_value[i]);
    }

    /* Dynamic-feasibility residual, what _dyn_feas->checkConstraint(_x) reports at :203: the base rows of
     * M qdd + h - sum J_i^T w_i, from the same record the solver saw.  Traced next to the solver diagnostics. */
    {
        Eigen::VectorXd res(6), diag(4);
        const double* M = _record.data() + _L.off_M;
        for (int r = 0; r < 6; ++r) {
            double v = _record[_L.off_h + r];
            for (int j = 0; j < nv; ++j) v += (r >= j ? M[r * (r + 1) / 2 + j] : M[j * (j + 1) / 2 + r]) * _x[j];
            for (size_t i = 0; i < _contact_links.size(); ++i)
                for (int k = 0; k < 3; ++k) v -= _record[_L.off_jc + (i * 6 + k) * nv + r] * _x[nv + 3 * i + k];
            res[r] = v;
        }
        _logger->add("dyn_feas_residual", res);
        diag[0] = tr.status; diag[1] = tr.iters & 0xffff; diag[2] = tr.iters >> 16; diag[3] = tr.kkt[0] > tr.kkt[1] ? tr.kkt[0] : tr.kkt[1];
        _logger->add("qp_status_iters_kkt", diag);
    }

    /* Torques due to contacts (:206-210) and inverse dynamics (:213-219): tau = M qdd + h - sum J^T w.
     * The kernel returns the actuated rows; the 6 base rows are the dyn-feas residual (zero). */
    _model->setJointAcceleration(_qddot_value);
    _model->update();
    for (int j = 0; j < 6; ++j) _tau[j] = 0.0;
    for (int a = 0; a < _L.n_a; ++a) _tau[6 + a] = _out[_L.n_x + a];
    _tau_c.setZero(nv);
    for (size_t i = 0; i < _contact_links.size(); ++i) {
        _model->getJacobian(_contact_links[i], _Jtmp);
        for (int j = 0; j < nv; ++j)
            for (int r = 0; r < 3; ++r) _tau_c[j] += _Jtmp(r, j) * _wrench_value[i][r];
    }
    _model->setJointEffort(_tau);

    /* Update model */                                                               // :222-230
    _model->getJointPosition(_q);
    _model->getJointVelocity(_qdot);
    _model->setJointPosition(_q);
    _model->setJointVelocity(_qdot);
    _model->update();

    _logger->add("tau", _tau);                                                       // :233-236
    _logger->add("tau_c", _tau_c);
    _logger->add("qddot_value", _qddot_value);
    _logger->add("x", _x);

    /* Send commands to robot */                                                     // :239-248
    if (enable_torque_ctrl) {
        _robot->setStiffness(_k);
        _robot->setDamping(_d);
    }
    _robot->setReferenceFrom(*_model, XBot::Sync::Position, XBot::Sync::Effort);
    _robot->move();
    _model->log(_logger, time);
    (void)period;
}

void XBotPlugin::ForceAccExample::sync_model()
{
    _model->syncFrom(*_robot);                                                       // :258
    Eigen::Affine3d w_T_fb;
    Eigen::Matrix3d w_R_fb;
    Eigen::Vector6d fb_twist;
    Eigen::Vector3d fb_vel, fb_omega, fb_pos;
    _sh_fb_pos.get(fb_pos);                                                          // :265-268
    _sh_fb_vel.get(fb_vel);
    _imu->getAngularVelocity(fb_omega);
    _imu->getOrientation(w_R_fb);
    w_T_fb.linear() = w_R_fb;                                                        // :270-272
    w_T_fb.translation() = fb_pos;
    for (int k = 0; k < 3; ++k) { fb_twist[k] = fb_vel[k]; fb_twist[3 + k] = fb_omega[k]; }
    _model->setFloatingBaseState(w_T_fb, fb_twist);                                  // :274-275
    _model->update();
    _model->getFloatingBasePose(w_T_fb);                                             // :279
}

// Boundary-test accessor (not in the reference surface): 0 = last record, 1 = last raw output, 2 = status.
extern "C" const double* qppvm_plugin_debug(XBot::XBotControlPlugin* p, int what, int* n)
{
    static double st;
    auto* s = static_cast<XBotPlugin::ForceAccExample*>(p);
    if (what == 0) { *n = (int)s->last_record().size(); return s->last_record().data(); }
    if (what == 1) { *n = (int)s->last_output().size(); return s->last_output().data(); }
    st = s->last_status(); *n = 1; return &st;
}
