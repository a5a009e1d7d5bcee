// Boundary test harness (tier T4): plays the role of XBotCore's plugin handler.  It dlopen()s a plugin
// library exactly as the loader would, builds it through its registration symbol, and drives
// init_control_plugin / on_start / control_loop / close against FAKE XBot::Handle / RobotInterface / ModelInterface
// objects that serve a recorded sequence of synthetic states.  What the plugin commanded is written to a file that
// tests/test_plugin.py compares with the CPU checker.
//
//   plugin_test <plugin.so> <factory symbol> <states.bin> <out.bin> <n_ticks> <n_v> <floating 0|1> link...
#include <XCM/XBotControlPlugin.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct LinkState { std::vector<double> J, jdqd, R, p, twist; };
struct TickState {
    std::vector<double> q, qd, home, M, h, tmax, fb_pos, fb_vel, imu_R, imu_w;
    std::vector<LinkState> links;
};

int g_nv = 0;
std::vector<std::string> g_links;
std::vector<TickState> g_ticks;
int g_cur = 0;

class FakeModel : public XBot::ModelInterface {
public:
    const TickState& s() const { return g_ticks[g_cur]; }
    int link(const std::string& name) const
    {
        for (size_t i = 0; i < g_links.size(); ++i) if (g_links[i] == name) return (int)i;
        std::fprintf(stderr, "FakeModel: unknown link %s\n", name.c_str());
        std::exit(3);
    }
    int getJointNum() const override { return g_nv; }
    bool getRobotState(const std::string&, Eigen::VectorXd& q) const override { q.resize(g_nv); for (int i = 0; i < g_nv; ++i) q[i] = s().home[i]; return true; }
    bool setJointPosition(const Eigen::VectorXd&) override { return true; }
    bool setJointVelocity(const Eigen::VectorXd&) override { return true; }
    bool setJointAcceleration(const Eigen::VectorXd& a) override { qdd = a; return true; }
    bool setJointEffort(const Eigen::VectorXd& t) override { effort = t; return true; }
    bool getJointPosition(Eigen::VectorXd& q) const override { q.resize(g_nv); for (int i = 0; i < g_nv; ++i) q[i] = s().q[i]; return true; }
    bool getJointVelocity(Eigen::VectorXd& v) const override { v.resize(g_nv); for (int i = 0; i < g_nv; ++i) v[i] = s().qd[i]; return true; }
    bool getJointEffort(Eigen::VectorXd& t) const override { t = effort; return true; }
    bool update() override { ++updates; return true; }
    bool getJacobian(const std::string& l, Eigen::MatrixXd& J) const override
    {
        const LinkState& ls = s().links[link(l)];
        J.resize(6, g_nv);
        for (int r = 0; r < 6; ++r) for (int j = 0; j < g_nv; ++j) J(r, j) = ls.J[r * g_nv + j];
        return true;
    }
    bool computeJdotQdot(const std::string& l, const Eigen::Vector3d&, Eigen::Vector6d& v) const override
    {
        for (int r = 0; r < 6; ++r) v[r] = s().links[link(l)].jdqd[r];
        return true;
    }
    void getInertiaMatrix(Eigen::MatrixXd& M) const override
    {
        M.resize(g_nv, g_nv);
        for (int i = 0; i < g_nv; ++i) for (int j = 0; j < g_nv; ++j) M(i, j) = s().M[i * g_nv + j];
    }
    void computeNonlinearTerm(Eigen::VectorXd& h) const override { h.resize(g_nv); for (int i = 0; i < g_nv; ++i) h[i] = s().h[i]; }
    void computeInverseDynamics(Eigen::VectorXd& tau) const override
    {
        tau.resize(g_nv);
        for (int i = 0; i < g_nv; ++i) { double v = s().h[i]; for (int j = 0; j < g_nv; ++j) v += s().M[i * g_nv + j] * (qdd.size() ? qdd[j] : 0.0); tau[i] = v; }
    }
    bool getPose(const std::string& l, Eigen::Affine3d& T) const override
    {
        const LinkState& ls = s().links[link(l)];
        for (int i = 0; i < 3; ++i) { T.translation()[i] = ls.p[i]; for (int j = 0; j < 3; ++j) T.linear()(i, j) = ls.R[i * 3 + j]; }
        return true;
    }
    bool getPointPosition(const std::string& l, const Eigen::Vector3d&, Eigen::Vector3d& w) const override
    {
        for (int i = 0; i < 3; ++i) w[i] = s().links[link(l)].p[i];
        return true;
    }
    bool getVelocityTwist(const std::string& l, Eigen::Vector6d& v) const override { for (int r = 0; r < 6; ++r) v[r] = s().links[link(l)].twist[r]; return true; }
    bool getEffortLimits(Eigen::VectorXd& t) const override { t.resize(g_nv); for (int i = 0; i < g_nv; ++i) t[i] = s().tmax[i]; return true; }
    bool getJointLimits(Eigen::VectorXd& lo, Eigen::VectorXd& hi) const override
    {
        lo.resize(g_nv); hi.resize(g_nv);
        for (int i = 0; i < g_nv; ++i) { lo[i] = s().home[i] - 0.35; hi[i] = s().home[i] + 0.35; }
        return true;
    }
    bool setFloatingBaseState(const Eigen::Affine3d& T, const Eigen::Vector6d& tw) override { fb_pose = T; fb_twist = tw; ++fb_sets; return true; }
    bool getFloatingBasePose(Eigen::Affine3d& T) const override { T = fb_pose; return true; }
    bool syncFrom(const XBot::RobotInterface&) override { ++syncs; return true; }
    bool getStiffness(Eigen::VectorXd& k) const override { k.setConstant(g_nv, 100.0); return true; }
    bool getDamping(Eigen::VectorXd& d) const override { d.setConstant(g_nv, 10.0); return true; }
    Eigen::VectorXd qdd, effort;
    Eigen::Affine3d fb_pose;
    Eigen::Vector6d fb_twist;
    int updates = 0, syncs = 0, fb_sets = 0;
};

std::shared_ptr<FakeModel> g_model;

class FakeRobot : public XBot::RobotInterface {
public:
    FakeRobot() { imu = std::make_shared<XBot::ImuSensor>(); }
    int getJointNum() const override { return g_nv; }
    bool getStiffness(Eigen::VectorXd& k) const override { k.setConstant(g_nv, 1600.0); return true; }
    bool getDamping(Eigen::VectorXd& d) const override { d.setConstant(g_nv, 40.0); return true; }
    bool setStiffness(const Eigen::VectorXd& k) override { stiffness = k; return true; }
    bool setDamping(const Eigen::VectorXd& d) override { damping = d; return true; }
    int getDofIndex(const std::string& j) const override { return (int)(std::hash<std::string>()(j) % (size_t)g_nv); }
    bool getMotorPosition(XBot::JointIdMap& m) const override { for (int i = 0; i < g_nv; ++i) m[i] = g_ticks[g_cur].q[i]; return true; }
    bool getMotorVelocity(XBot::JointIdMap& m) const override { for (int i = 0; i < g_nv; ++i) m[i] = g_ticks[g_cur].qd[i]; return true; }
    bool setReferenceFrom(const XBot::ModelInterface& model, XBot::Sync, XBot::Sync) override { model.getJointEffort(effort_ref); ++refs; return true; }
    bool move() override { ++moves; return true; }
    std::map<std::string, XBot::ImuSensor::ConstPtr> getImu() const override { return {{"imu_link", imu}}; }
    void load_imu()
    {
        const TickState& s = g_ticks[g_cur];
        for (int i = 0; i < 3; ++i) { imu->omega[i] = s.imu_w[i]; for (int j = 0; j < 3; ++j) imu->orientation(i, j) = s.imu_R[i * 3 + j]; }
    }
    std::shared_ptr<XBot::ImuSensor> imu;
    Eigen::VectorXd stiffness, damping, effort_ref;
    int moves = 0, refs = 0;
};

class FakeHandle : public XBot::Handle {
public:
    FakeHandle() : robot(std::make_shared<FakeRobot>()), shm(std::make_shared<XBot::SharedMemory>()) {}
    XBot::RobotInterface::Ptr getRobotInterface() override { return robot; }
    std::string getPathToConfigFile() const override { return "fake://synthetic_humanoid.yaml"; }
    XBot::SharedMemory::Ptr getSharedMemory() override { return shm; }
    std::shared_ptr<FakeRobot> robot;
    XBot::SharedMemory::Ptr shm;
};

}  // namespace

// symbols the plugin libraries resolve against the host process (as they would against libXBotInterface)
namespace XBot {
ModelInterface::Ptr ModelInterface::getModel(const std::string&) { return g_model; }
int Logger::n_errors = 0;
void Logger::error(const char* msg) { ++n_errors; std::fprintf(stderr, "[XBot::Logger::error] %s\n", msg); }
}  // namespace XBot

int main(int argc, char** argv)
{
    if (argc < 9) { std::fprintf(stderr, "usage: plugin_test lib factory states out n_ticks n_v floating link...\n"); return 2; }
    const char* lib = argv[1]; const char* factory = argv[2];
    const int n_ticks = std::atoi(argv[5]); g_nv = std::atoi(argv[6]);
    const bool floating = std::atoi(argv[7]) != 0;
    for (int i = 8; i < argc; ++i) g_links.push_back(argv[i]);
    FILE* f = std::fopen(argv[3], "rb");
    if (!f) { std::perror("states"); return 2; }
    auto rd = [&](std::vector<double>& v, size_t n) { v.resize(n); if (std::fread(v.data(), 8, n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); } };
    g_ticks.resize(n_ticks);
    for (auto& t : g_ticks) {
        rd(t.q, g_nv); rd(t.qd, g_nv); rd(t.home, g_nv); rd(t.M, (size_t)g_nv * g_nv); rd(t.h, g_nv); rd(t.tmax, g_nv);
        rd(t.fb_pos, 3); rd(t.fb_vel, 3); rd(t.imu_R, 9); rd(t.imu_w, 3);
        t.links.resize(g_links.size());
        for (auto& l : t.links) { rd(l.J, (size_t)6 * g_nv); rd(l.jdqd, 6); rd(l.R, 9); rd(l.p, 3); rd(l.twist, 6); }
    }
    std::fclose(f);

    void* h = dlopen(lib, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { std::fprintf(stderr, "dlopen: %s\n", dlerror()); return 4; }
    auto create = (XBot::XBotControlPlugin * (*)()) dlsym(h, factory);
    auto debug = (const double* (*)(XBot::XBotControlPlugin*, int, int*))dlsym(h, "qppvm_plugin_debug");
    if (!create || !debug) { std::fprintf(stderr, "missing symbol %s\n", factory); return 4; }
    g_model = std::make_shared<FakeModel>();
    auto handle = std::make_shared<FakeHandle>();
    XBot::XBotControlPlugin* plugin = create();
    g_cur = 0;
    handle->robot->load_imu();
    if (!plugin->init_control_plugin(handle)) { std::fprintf(stderr, "init_control_plugin failed\n"); return 5; }
    auto set_tick = [&](int t) {
        g_cur = t;
        handle->robot->load_imu();
        if (floating) {
            Eigen::Vector3d p(g_ticks[t].fb_pos[0], g_ticks[t].fb_pos[1], g_ticks[t].fb_pos[2]);
            Eigen::Vector3d v(g_ticks[t].fb_vel[0], g_ticks[t].fb_vel[1], g_ticks[t].fb_vel[2]);
            handle->shm->getSharedObject<Eigen::Vector3d>("/gazebo/floating_base_position").set(p);
            handle->shm->getSharedObject<Eigen::Vector3d>("/gazebo/floating_base_velocity").set(v);
        }
    };
    set_tick(0);
    plugin->on_start(0.0);
    FILE* out = std::fopen(argv[4], "wb");
    for (int t = 0; t < n_ticks; ++t) {
        set_tick(t);
        const int moves0 = handle->robot->moves;
        plugin->run(0.001 * t, 0.001);
        int n;
        const double* st = debug(plugin, 2, &n);
        double head[3] = {st[0], (double)(handle->robot->moves - moves0), (double)XBot::Logger::n_errors};
        std::fwrite(head, 8, 3, out);
        std::vector<double> eff(g_nv, 0.0);
        for (int i = 0; i < handle->robot->effort_ref.size() && i < g_nv; ++i) eff[i] = handle->robot->effort_ref[i];
        std::fwrite(eff.data(), 8, g_nv, out);
        const double* rec = debug(plugin, 0, &n); std::fwrite(rec, 8, n, out);
        const double* o = debug(plugin, 1, &n); std::fwrite(o, 8, n, out);
    }
    std::fclose(out);
    const bool closed = plugin->close();
    std::printf("ticks=%d moves=%d model_updates=%d fb_sets=%d logger_errors=%d close=%d\n", n_ticks, handle->robot->moves,
                g_model->updates, g_model->fb_sets, XBot::Logger::n_errors, (int)closed);
    delete plugin;
    return 0;
}
