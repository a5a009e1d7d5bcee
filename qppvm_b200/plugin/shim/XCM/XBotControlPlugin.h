// Stand-in for <XCM/XBotControlPlugin.h> (ref:include/ForceAccPlugin/ForceAcc.h:23,36,
// ref:include/QPPVM_RT_plugin/QPPVMPlugin.h:23,37).  TEST SHIM ONLY.
#pragma once
#include <XBotInterface/XBotInterface.h>

namespace XBot {

class XBotControlPlugin {
public:
    virtual ~XBotControlPlugin() {}
    virtual bool init_control_plugin(XBot::Handle::Ptr handle) = 0;
    virtual void on_start(double time) {}
    virtual void on_stop(double time) {}
    virtual bool close() = 0;
    // XBotCore's plugin handler calls run() once per control period, which dispatches to control_loop()
    void run(double time, double period) { control_loop(time, period); }
protected:
    virtual void control_loop(double time, double period) = 0;
};

}  // namespace XBot

// The two registration flavours the reference uses (ref:src/ForceAcc.cpp:26, ref:src/QPPVMPlugin.cpp:29).
#define REGISTER_XBOT_PLUGIN_(cls)                                                         \
    extern "C" XBot::XBotControlPlugin* create_instance() { return new cls(); }             \
    extern "C" void destroy_instance(XBot::XBotControlPlugin* p) { delete p; }
#define REGISTER_XBOT_PLUGIN(name, cls)                                                    \
    extern "C" XBot::XBotControlPlugin* name##_factory() { return new cls(); }              \
    extern "C" void name##_factory_destroy(XBot::XBotControlPlugin* p) { delete p; }
