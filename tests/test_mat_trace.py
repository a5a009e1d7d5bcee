"""The MatLogger-compatible trace writer (qppvm_b200/plugin/MatTrace.h; the reference's per-tick
dumps: ref:src/QPPVMPlugin.cpp:250-258, ref:src/ForceAcc.cpp:200,233-236): a Level-5 MAT-file that MATLAB / scipy read."""
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = r'''
#include "MatTrace.h"
int main(int argc, char** argv) {
    qppvm::MatTrace t;
    for (int k = 0; k < 5; ++k) {
        double v[3] = {1.0 * k, -2.5 * k, 1e-3 * k * k};
        t.add("tau_qp", v, 3);
        t.add("time_matlogger", 0.001 * k);
        double w[7] = {k + 0.5, 0, 0, 0, 0, 0, -1.0};
        t.add("a_seven_char_name_that_needs_padding", w, 7);
    }
    return t.flush(argv[1]) && t.samples("tau_qp") == 5 ? 0 : 1;
}
'''


def test_mat_trace_round_trip(tmp_path):
    import scipy.io
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(HERE, "..", "qppvm_b200", "plugin"),
                           "-o", str(exe), str(src)])
    subprocess.check_call([str(exe), str(tmp_path / "trace.mat")])
    m = scipy.io.loadmat(str(tmp_path / "trace.mat"))
    k = np.arange(5.0)
    np.testing.assert_array_equal(m["tau_qp"], np.stack([k, -2.5 * k, 1e-3 * k * k]))
    np.testing.assert_array_equal(m["time_matlogger"], 0.001 * k[None])
    assert m["a_seven_char_name_that_needs_padding"].shape == (7, 5)
    np.testing.assert_array_equal(m["a_seven_char_name_that_needs_padding"][0], k + 0.5)
