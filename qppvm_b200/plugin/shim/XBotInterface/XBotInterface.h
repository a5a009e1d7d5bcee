// Minimal stand-in for the XBotInterface / XBotCore types the two plugins touch (SURVEY.md 8(b) "Inputs consumed
// through the boundary").  TEST SHIM ONLY: the real headers are not in this image.  Every method here exists with
// the same name and argument meaning in XBot::ModelInterface / RobotInterface / ImuSensor / SharedObject /
// MatLogger / Handle as the reference calls them (file:line next to each).
#pragma once
#include <Eigen/Dense>
#include "../../MatTrace.h"
#include <map>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

namespace XBot {

enum class Sync { Position, Velocity, Effort, Stiffness, Damping, All };
typedef std::map<int, double> JointIdMap;

class MatLogger {
public:
    typedef std::shared_ptr<MatLogger> Ptr;
    static Ptr getLogger(const std::string& path)                                            // ForceAcc.cpp:34
    {
        auto p = std::make_shared<MatLogger>();
        p->path_ = path;
        return p;
    }
    void add(const std::string& name, const Eigen::VectorXd& v)                              // ForceAcc.cpp:200,233-236
    {
        last_[name] = v;
        trace_.add(name, v.data(), (int)v.size());
    }
    void add(const std::string& name, double v) { trace_.add(name, v); }                     // QPPVMPlugin.cpp:322
    // the real MatLogger writes <path>.mat on flush; here the file lands next to the test outputs (QPPVM_TRACE_DIR)
    void flush()                                                                             // ForceAcc.h:43
    {
        ++flushes;
        const char* dir = getenv("QPPVM_TRACE_DIR");
        if (!dir) return;
        std::string base = path_.substr(path_.find_last_of('/') == std::string::npos ? 0 : path_.find_last_of('/') + 1);
        trace_.flush(std::string(dir) + "/" + base + ".mat");
    }
    std::map<std::string, Eigen::VectorXd> last_;
    qppvm::MatTrace trace_;
    std::string path_;
    int flushes = 0;
};

struct Logger {
    static void error(const char* msg);                                                     // ForceAcc.cpp:191
    static int n_errors;
};

class ImuSensor {
public:
    typedef std::shared_ptr<const ImuSensor> ConstPtr;
    void getAngularVelocity(Eigen::Vector3d& w) const { w = omega; }                         // ForceAcc.cpp:267
    void getOrientation(Eigen::Matrix3d& R) const { R = orientation; }                       // ForceAcc.cpp:268
    Eigen::Vector3d omega;
    Eigen::Matrix3d orientation;
};

template <typename T>
class SharedObject {
public:
    void set(const T& v) { if (v_) *v_ = v; }                                                // ForceAcc.cpp:54-55
    void get(T& v) const { if (v_) v = *v_; }                                                // ForceAcc.cpp:265-266
    std::shared_ptr<T> v_;
};

class SharedMemory {
public:
    typedef std::shared_ptr<SharedMemory> Ptr;
    template <typename T>
    SharedObject<T> getSharedObject(const std::string& name)                                 // ForceAcc.cpp:52-53
    {
        SharedObject<T> o;
        auto it = vec3_.find(name);
        if (it == vec3_.end()) it = vec3_.emplace(name, std::make_shared<T>()).first;
        o.v_ = it->second;
        return o;
    }
    std::map<std::string, std::shared_ptr<Eigen::Vector3d>> vec3_;
};

// Rigid-body model.  The fake implementation (plugin_test.cpp) serves J, M, h, Jdot*qdot and poses of a recorded
// synthetic state; the plugins only see this interface.
class ModelInterface {
public:
    typedef std::shared_ptr<ModelInterface> Ptr;
    virtual ~ModelInterface() {}
    static Ptr getModel(const std::string& path_to_config);                                  // ForceAcc.cpp:43, QPPVMPlugin.cpp:50
    virtual int getJointNum() const = 0;                                                     // ForceAcc.cpp:64
    virtual bool getRobotState(const std::string& name, Eigen::VectorXd& q) const = 0;       // "home": ForceAcc.cpp:46
    virtual bool setJointPosition(const Eigen::VectorXd& q) = 0;
    virtual bool setJointVelocity(const Eigen::VectorXd& qd) = 0;
    virtual bool setJointAcceleration(const Eigen::VectorXd& qdd) = 0;                       // ForceAcc.cpp:213
    virtual bool setJointEffort(const Eigen::VectorXd& tau) = 0;                             // ForceAcc.cpp:219
    virtual bool getJointPosition(Eigen::VectorXd& q) const = 0;
    virtual bool getJointVelocity(Eigen::VectorXd& qd) const = 0;
    virtual bool getJointEffort(Eigen::VectorXd& tau) const = 0;
    virtual bool update() = 0;
    virtual bool getJacobian(const std::string& link, Eigen::MatrixXd& J) const = 0;         // ForceAcc.cpp:208
    virtual bool computeJdotQdot(const std::string& link, const Eigen::Vector3d& p, Eigen::Vector6d& jdqd) const = 0;
    virtual void getInertiaMatrix(Eigen::MatrixXd& M) const = 0;
    virtual void computeNonlinearTerm(Eigen::VectorXd& h) const = 0;                          // QPPVMPlugin.cpp:312
    virtual void computeInverseDynamics(Eigen::VectorXd& tau) const = 0;                      // ForceAcc.cpp:218
    virtual bool getPose(const std::string& link, Eigen::Affine3d& T) const = 0;              // QPPVMPlugin.cpp:272
    virtual bool getPointPosition(const std::string& link, const Eigen::Vector3d& p, Eigen::Vector3d& w_p) const = 0;
    virtual bool getVelocityTwist(const std::string& link, Eigen::Vector6d& v) const = 0;
    virtual bool getEffortLimits(Eigen::VectorXd& tmax) const = 0;                            // QPPVMPlugin.cpp:56
    virtual bool getJointLimits(Eigen::VectorXd& qmin, Eigen::VectorXd& qmax) const = 0;      // QPPVMPlugin.cpp:120
    virtual bool setFloatingBaseState(const Eigen::Affine3d& T, const Eigen::Vector6d& twist) = 0;   // ForceAcc.cpp:274
    virtual bool getFloatingBasePose(Eigen::Affine3d& T) const = 0;                           // ForceAcc.cpp:279
    virtual bool syncFrom(const class RobotInterface& robot) = 0;                             // ForceAcc.cpp:258
    virtual bool getStiffness(Eigen::VectorXd& k) const = 0;
    virtual bool getDamping(Eigen::VectorXd& d) const = 0;
    virtual void initLog(MatLogger::Ptr, int) {}                                              // ForceAcc.cpp:50
    virtual void log(MatLogger::Ptr, double) {}                                               // ForceAcc.cpp:249
};

class RobotInterface {
public:
    typedef std::shared_ptr<RobotInterface> Ptr;
    virtual ~RobotInterface() {}
    virtual int getJointNum() const = 0;
    virtual bool getStiffness(Eigen::VectorXd& k) const = 0;                                  // ForceAcc.cpp:36
    virtual bool getDamping(Eigen::VectorXd& d) const = 0;                                    // ForceAcc.cpp:37
    virtual bool setStiffness(const Eigen::VectorXd& k) = 0;                                  // ForceAcc.cpp:240
    virtual bool setDamping(const Eigen::VectorXd& d) = 0;
    virtual int getDofIndex(const std::string& joint) const = 0;                              // QPPVMPlugin.cpp:84
    virtual bool getMotorPosition(JointIdMap& q) const = 0;                                   // QPPVMPlugin.cpp:346
    virtual bool getMotorVelocity(JointIdMap& qd) const = 0;
    virtual bool setReferenceFrom(const ModelInterface& model, Sync a, Sync b = Sync::All) = 0;   // ForceAcc.cpp:242
    virtual bool move() = 0;                                                                  // ForceAcc.cpp:248
    virtual std::map<std::string, ImuSensor::ConstPtr> getImu() const = 0;                    // ForceAcc.cpp:41
};

class Handle {
public:
    typedef std::shared_ptr<Handle> Ptr;
    virtual ~Handle() {}
    virtual RobotInterface::Ptr getRobotInterface() = 0;                                      // ForceAcc.cpp:33
    virtual std::string getPathToConfigFile() const = 0;                                      // ForceAcc.cpp:43
    virtual SharedMemory::Ptr getSharedMemory() = 0;                                          // ForceAcc.cpp:52
};

}  // namespace XBot
