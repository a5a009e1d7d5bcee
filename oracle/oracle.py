"""ctypes front end of the CPU oracle (oracle/qppvm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package qppvm_b200/.
PARITY UNPINNED (no reference binary, no golden vectors upstream) -- see the C file header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
FACTOR_CHOLESKY = 0   # H = A^T A + eps I formed, Cholesky: the reference's numerics
FACTOR_QR = 1         # Householder QR of [A; sqrt(eps) I]


class CDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_a", C.c_int32), ("n_contacts", C.c_int32), ("flags", C.c_int32),
                ("eps_regularisation", C.c_double), ("n_reg_steps", C.c_int32), ("max_iter", C.c_int32),
                ("device", C.c_int32), ("postural_actuated_only", C.c_int32), ("lambda_solver", C.c_double),
                ("task_weight", C.c_double * 3)]


LAYOUT_FIELDS = ("n_a", "n_v", "n_c", "n_x", "n_rows", "row_dyn", "row_box", "row_cone", "row_tau",
                 "row_opt", "off_jwaist", "off_jc", "off_M", "off_h", "off_jdqd", "off_rhs",
                 "off_taulim", "off_cone", "off_fbox", "off_fee", "off_tauj", "rec_doubles",
                 "out_bytes", "diag_doubles", "off_jlim", "off_jelbow", "off_felbow", "off_com")


class CLayout(C.Structure):
    _fields_ = [(f, C.c_int32) for f in LAYOUT_FIELDS]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "qppvm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None
_NATIVE_PATH = os.path.join(_HERE, "_build", "liboracle_native.so")


def use_native() -> bool:
    """Switches this process to a -march=native build made on THIS machine (bench.py's CPU legs); falls back to the
    portable library when the compiler is missing.  Returns whether the native build is in use."""
    global _lib
    try:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "native"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        return False
    _lib = None
    lib(_NATIVE_PATH)
    return True


def lib(path: str | None = None):
    global _lib
    if _lib is None:
        if path is None and not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(path or _LIB_PATH)
        _lib.oracle_solve_sequence.argtypes = [C.POINTER(CDesc), C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
        _lib.oracle_solve_batch.argtypes = [C.POINTER(CDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_longlong, C.c_int, C.c_int]
        _lib.oracle_layout.argtypes = [C.POINTER(CDesc), C.POINTER(CLayout)]
        _lib.oracle_dense_qp.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        _lib.oracle_assemble.argtypes = [C.POINTER(CDesc), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_int), C.POINTER(C.c_double)]
    return _lib


def cdesc(desc) -> CDesc:
    return CDesc(desc.kind, desc.n_a, desc.n_contacts, desc.flags, desc.eps_regularisation,
                 desc.n_reg_steps, desc.max_iter, desc.device, desc.postural_actuated_only, desc.lambda_solver,
                 (C.c_double * 3)(*desc.task_weight))


def c_layout(desc) -> dict:
    L = CLayout()
    if lib().oracle_layout(C.byref(cdesc(desc)), C.byref(L)):
        raise ValueError("oracle_layout rejected the description")
    return {f: getattr(L, f) for f in LAYOUT_FIELDS}


def num_threads() -> int:
    return lib().oracle_num_threads()


def solve_batch(desc, records: np.ndarray, mode: int = FACTOR_QR, threads: int = 0, diag: bool = False):
    """records (B, rec_doubles) float64 -> (out (B, out_bytes/8) float64 view, diag or None)."""
    L = c_layout(desc)
    records = np.ascontiguousarray(records, dtype=np.float64)
    B = records.shape[0]
    assert records.shape[1] == L["rec_doubles"]
    out = np.zeros((B, L["out_bytes"] // 8))
    dg = np.zeros((B, L["diag_doubles"])) if diag else None
    rc = lib().oracle_solve_batch(C.byref(cdesc(desc)), records.ctypes.data, out.ctypes.data,
                                  dg.ctypes.data if diag else None, B, mode, threads)
    if rc:
        raise RuntimeError("oracle_solve_batch failed: %d" % rc)
    return out, dg


def solve_sequence(desc, records: np.ndarray, warm: np.ndarray, mode: int = FACTOR_QR) -> np.ndarray:
    """Consecutive ticks on one thread, hot-started: `warm` (8 x uint32, in/out) carries the working sets."""
    L = c_layout(desc)
    records = np.ascontiguousarray(records, dtype=np.float64).reshape(-1, L["rec_doubles"])
    assert warm.dtype == np.uint32 and warm.size == 8 and warm.flags.c_contiguous
    out = np.zeros((records.shape[0], L["out_bytes"] // 8))
    rc = lib().oracle_solve_sequence(C.byref(cdesc(desc)), records.ctypes.data, out.ctypes.data, records.shape[0], mode,
                                     warm.ctypes.data)
    if rc:
        raise RuntimeError("oracle_solve_sequence failed: %d" % rc)
    return out


def dense_qp(A, b, Cm, lA, uA, eps, n_reg_steps=0, mode=FACTOR_QR, max_iter=1000):
    """min 1/2||Ax-b||^2 + eps/2||x||^2 s.t. lA <= C x <= uA -> (status, x, y, iters, kkt)."""
    A = np.ascontiguousarray(A, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    Cm = np.ascontiguousarray(Cm, dtype=np.float64).reshape(-1, A.shape[1])
    lA = np.ascontiguousarray(lA, dtype=np.float64); uA = np.ascontiguousarray(uA, dtype=np.float64)
    m, n = A.shape
    nc = Cm.shape[0]
    x = np.zeros(n); y = np.zeros(max(nc, 1))
    it = C.c_int(0); kkt = C.c_double(0)
    st = lib().oracle_dense_qp(n, m, A.ctypes.data, b.ctypes.data, nc, Cm.ctypes.data, lA.ctypes.data,
                               uA.ctypes.data, eps, n_reg_steps, mode, max_iter, x.ctypes.data,
                               y.ctypes.data, C.byref(it), C.byref(kkt))
    return st, x, y[:nc], it.value, kkt.value


def assemble(desc, record: np.ndarray, level: int, x0=None):
    """Explicit (A, b, C, lA, uA, eps) of one level of one record, as OpenSoT would hand to qpOASES."""
    L = c_layout(desc)
    n = L["n_x"]
    A = np.zeros((160, n)); b = np.zeros(160); Cm = np.zeros((128, n)); lA = np.zeros(128); uA = np.zeros(128)
    dims = (C.c_int * 2)(); eps = C.c_double(0)
    record = np.ascontiguousarray(record, dtype=np.float64)
    x0 = np.zeros(n) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
    rc = lib().oracle_assemble(C.byref(cdesc(desc)), record.ctypes.data, level, x0.ctypes.data, A.ctypes.data,
                               b.ctypes.data, Cm.ctypes.data, lA.ctypes.data, uA.ctypes.data, dims, C.byref(eps))
    if rc:
        raise RuntimeError("oracle_assemble failed")
    m, nc = dims[0], dims[1]
    return A[:m].copy(), b[:m].copy(), Cm[:nc].copy(), lA[:nc].copy(), uA[:nc].copy(), eps.value


def split_out(desc, out: np.ndarray):
    """out (B, out_doubles) -> dict(x, tau, status, iters0, iters1, active (B,4) uint32, kkt (B,2) float32)."""
    L = c_layout(desc)
    n, na = L["n_x"], L["n_a"]
    tr = np.ascontiguousarray(out[:, n + na:n + na + 4]).view(np.uint8).reshape(out.shape[0], 32)
    i32 = tr[:, 0:8].copy().view(np.int32)
    return dict(x=out[:, :n], tau=out[:, n:n + na], status=i32[:, 0], iters0=i32[:, 1] & 0xffff,
                iters1=(i32[:, 1] >> 16) & 0xffff, active=tr[:, 8:24].copy().view(np.uint32),
                kkt=tr[:, 24:32].copy().view(np.float32))
