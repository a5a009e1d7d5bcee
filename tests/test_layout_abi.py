"""Host logic without a GPU: layouts agree across Python / C-ABI / oracle, the shared library loads
and exports every symbol include/qppvm_b200.h declares, algorithmic bytes match SURVEY.md 8(d)."""
import ctypes
import os
import re

import pytest

from qppvm_b200 import api, build
from qppvm_b200.layout import (CONFIGS, Desc, KIND_TORQUE, layout, FLAG_FRICTION_CONES, FLAG_TORQUE_LIMITS, FLAG_ELBOW_TASKS,
                               FLAG_JOINT_LIMITS, FLAG_COM_TASK)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return api.load_library()


DESCS = [c["desc"] for c in CONFIGS.values()] + [
    Desc(n_a=33, n_contacts=4, flags=0), Desc(n_a=39, n_contacts=4, flags=0),
    Desc(n_a=12, n_contacts=1, flags=FLAG_FRICTION_CONES), Desc(n_a=20, n_contacts=3, flags=FLAG_TORQUE_LIMITS),
    Desc(kind=KIND_TORQUE, n_a=39, n_contacts=2, flags=0, eps_regularisation=1.0),
    Desc(kind=KIND_TORQUE, n_a=29, n_contacts=2, flags=FLAG_JOINT_LIMITS, eps_regularisation=1.0),
    Desc(kind=KIND_TORQUE, n_a=29, n_contacts=2, flags=FLAG_ELBOW_TASKS, eps_regularisation=1.0),
    Desc(kind=KIND_TORQUE, n_a=33, n_contacts=2, flags=FLAG_JOINT_LIMITS | FLAG_ELBOW_TASKS, eps_regularisation=1.0),
    Desc(n_a=29, n_contacts=2, flags=FLAG_COM_TASK), Desc(n_a=33, n_contacts=4, flags=FLAG_COM_TASK | FLAG_FRICTION_CONES | FLAG_TORQUE_LIMITS),
]


@pytest.mark.parametrize("desc", DESCS)
def test_layouts_agree(lib, oracle_mod, desc):
    L = layout(desc)
    py = {f: getattr(L, f) for f in L.FIELDS}
    assert py == api.c_layout(desc)
    assert py == oracle_mod.c_layout(desc)
    assert L.rec_doubles % 2 == 0 and L.out_bytes % 8 == 0 and L.n_rows <= 128


def test_algorithmic_bytes_match_survey():
    # cfg[1] 11 616 B; cfg[0],[4] 12 240 B; cfg[2],[3] 18 448 B
    assert layout(CONFIGS[1]["desc"]).algorithmic_bytes() == 11616
    assert layout(CONFIGS[0]["desc"]).algorithmic_bytes() == 12240
    assert layout(CONFIGS[4]["desc"]).algorithmic_bytes() == 12240
    assert layout(CONFIGS[2]["desc"]).algorithmic_bytes() == 18448
    assert layout(CONFIGS[3]["desc"]).algorithmic_bytes() == 18448


def test_survey_dimensions():
    # SURVEY.md 8(a) per-config table: n_x, n_C(L0)/n_C(L1)
    for ci, nx, nc1 in ((1, 41, 24), (0, 41, 63), (2, 51, 89)):
        L = layout(CONFIGS[ci]["desc"])
        assert (L.n_x, L.n_rows) == (nx, nc1) and L.row_opt == nc1 - 6


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "qppvm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(qppvm_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(api.EXPORTS), declared ^ set(api.EXPORTS)
    raw = ctypes.CDLL(api.LIB_PATH)
    for name in declared:
        assert getattr(raw, name) is not None


def test_bad_descriptions_rejected(lib):
    for bad in (Desc(n_a=0), Desc(n_a=59), Desc(n_contacts=0), Desc(n_contacts=5), Desc(n_a=58, n_contacts=4)):
        with pytest.raises(ValueError):
            layout(bad)
        with pytest.raises(ValueError):
            api.c_layout(bad)


def test_shapes_cover_baseline_configs(lib):
    shapes = set(api.supported_shapes())
    for c in CONFIGS.values():
        d = c["desc"]
        assert (d.kind, d.n_a, d.n_contacts, d.flags) in shapes


def test_no_cpu_fallback_in_product(lib):
    """Without a CUDA device the product must fail loudly (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.QPError):
        api.Solver(CONFIGS[1]["desc"])


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "qppvm_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dp, fn)).read()
                assert "oracle" not in src.lower() or fn == "__init__.py" and False, os.path.join(dp, fn)


def test_null_handle_is_an_argument_error_everywhere():
    """No entry point dereferences a null handle (no GPU needed: each returns before touching the device)."""
    raw = ctypes.CDLL(api.LIB_PATH)
    P, I64, D = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double
    calls = {
        "qppvm_destroy": (P,), "qppvm_solve_batch": (P, P, P, I64, P), "qppvm_solve_batch_diag": (P, P, P, P, I64, P),
        "qppvm_solve_batch_host": (P, P, P, I64), "qppvm_solve_batch_host_async": (P, P, P, I64), "qppvm_host_sync": (P,),
        "qppvm_solve_one": (P, P, P), "qppvm_set_robot": (P, P), "qppvm_records_from_states": (P, P, P, I64, P),
        "qppvm_solve_states_host": (P, P, P, I64), "qppvm_solve_states_host_async": (P, P, P, I64),
        "qppvm_integrate_states": (P, P, P, D, I64, P), "qppvm_rollout_states": (P, P, P, ctypes.c_int, D, I64, P),
        "qppvm_fp64_peak": (P, P),
    }
    for name, sig in calls.items():
        fn = getattr(raw, name)
        fn.argtypes = list(sig); fn.restype = ctypes.c_int
        args = [None if t is P else (1e-3 if t is D else 1) for t in sig]
        assert fn(*args) == 1, name                      # QPPVM_ERR_ARG
    raw.qppvm_kernel_launches.argtypes = [P]; raw.qppvm_kernel_launches.restype = I64
    assert raw.qppvm_kernel_launches(None) == 0
    raw.qppvm_last_error.argtypes = [P]; raw.qppvm_last_error.restype = ctypes.c_char_p
    assert raw.qppvm_last_error(None) is not None
