"""Builds the plugin drop-ins against the test shim: libForceAccPlugin.so, libQPPVMPlugin.so (the reference's
library names, ref:CMakeLists.txt:48-49) and the boundary-test harness plugin_test."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
COMMON = ["-std=c++17", "-O2", "-fPIC", "-Wall", "-Wno-unused-parameter", "-I", os.path.join(HERE, "shim"), "-I", HERE]
TARGETS = {
    "libForceAccPlugin.so": ["ForceAccPlugin.cpp"],
    "libQPPVMPlugin.so": ["QPPVMPlugin.cpp"],
}


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = srcs + [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".h")]
    for dp, _, fns in os.walk(os.path.join(HERE, "shim")):
        deps += [os.path.join(dp, f) for f in fns]
    deps.append(os.path.join(PKG, "..", "include", "qppvm_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False) -> None:
    for name, srcs in TARGETS.items():
        out = os.path.join(HERE, name)
        srcs = [os.path.join(HERE, s) for s in srcs]
        if force or _stale(out, srcs):
            subprocess.check_call([CXX] + COMMON + ["-shared", "-o", out] + srcs +
                                  ["-L", PKG, "-lqppvm_b200", "-Wl,-rpath,$ORIGIN/.."])
    out = os.path.join(HERE, "plugin_test")
    src = [os.path.join(HERE, "plugin_test.cpp")]
    if force or _stale(out, src):
        subprocess.check_call([CXX] + COMMON + ["-rdynamic", "-o", out] + src + ["-ldl"])


if __name__ == "__main__":
    build(force=True)
    print("built", list(TARGETS), "plugin_test")
