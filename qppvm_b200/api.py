"""Host-side front end of the C-ABI (``include/qppvm_b200.h``) via ctypes.

PyTorch is used only as plumbing (device memory, streams, torch.distributed); the solve is the
hand-written sm_100a kernel in ``libqppvm_b200.so``.  There is no CPU path: constructing a
:class:`Solver` without the built library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .layout import Desc, Layout, layout

_HERE = os.path.dirname(os.path.abspath(__file__))
# QPPVM_B200_LIB: another build of the same library (kernel experiments: tools/build_variant.sh)
LIB_PATH = os.environ.get("QPPVM_B200_LIB") or os.path.join(_HERE, "libqppvm_b200.so")


class CDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_a", C.c_int32), ("n_contacts", C.c_int32), ("flags", C.c_int32),
                ("eps_regularisation", C.c_double), ("n_reg_steps", C.c_int32), ("max_iter", C.c_int32),
                ("device", C.c_int32), ("postural_actuated_only", C.c_int32), ("lambda_solver", C.c_double),
                ("task_weight", C.c_double * 3)]


class CLayout(C.Structure):
    _fields_ = [(f, C.c_int32) for f in Layout.FIELDS]


class CRobot(C.Structure):
    _fields_ = [("n_a", C.c_int32), ("parent", C.POINTER(C.c_int32)), ("axis", C.POINTER(C.c_double)),
                ("offset", C.POINTER(C.c_double)), ("mass", C.POINTER(C.c_double)), ("com", C.POINTER(C.c_double)),
                ("inertia", C.POINTER(C.c_double)), ("q_home", C.POINTER(C.c_double)), ("tau_max", C.POINTER(C.c_double)),
                ("contact_body", C.POINTER(C.c_int32))]


EXPORTS = ("qppvm_get_layout", "qppvm_create", "qppvm_destroy", "qppvm_last_error", "qppvm_solve_batch",
           "qppvm_solve_batch_diag", "qppvm_solve_batch_host", "qppvm_solve_one", "qppvm_kernel_launches",
           "qppvm_fp64_peak", "qppvm_supported_shapes", "qppvm_state_doubles", "qppvm_set_robot",
           "qppvm_records_from_states", "qppvm_solve_states_host", "qppvm_solve_batch_host_async", "qppvm_host_sync",
           "qppvm_solve_states_host_async", "qppvm_integrate_states", "qppvm_rollout_states", "qppvm_solve_batch_warm", "qppvm_reset_warm", "qppvm_tick_stamps",
           "qppvm_reserve_sms", "qppvm_kernel_timing", "qppvm_integrate_states_tracking", "qppvm_multi_create", "qppvm_multi_destroy", "qppvm_multi_last_error", "qppvm_multi_devices",
           "qppvm_multi_set_robot", "qppvm_multi_solve_batch", "qppvm_multi_solve_states", "qppvm_multi_solve_batch_host",
           "qppvm_multi_solve_states_host", "qppvm_multi_kernel_launches", "qppvm_multi_nccl_calls")

_lib = None


def load_library():
    """Loads the in-tree native library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("native library missing: %s (run `python -m qppvm_b200.build`); "
                               "qppvm_b200 has no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        P = C.c_void_p
        lib.qppvm_get_layout.argtypes = [C.POINTER(CDesc), C.POINTER(CLayout)]
        lib.qppvm_create.argtypes = [C.POINTER(CDesc), C.POINTER(P)]
        lib.qppvm_destroy.argtypes = [P]
        lib.qppvm_last_error.argtypes = [P]
        lib.qppvm_last_error.restype = C.c_char_p
        lib.qppvm_solve_batch.argtypes = [P, P, P, C.c_int64, P]
        lib.qppvm_solve_batch_diag.argtypes = [P, P, P, P, C.c_int64, P]
        lib.qppvm_solve_batch_warm.argtypes = [P, P, P, P, C.c_int64, P]
        lib.qppvm_solve_batch_host.argtypes = [P, P, P, C.c_int64]
        lib.qppvm_solve_one.argtypes = [P, P, P]
        lib.qppvm_reset_warm.argtypes = [P]
        lib.qppvm_reserve_sms.argtypes = [P, C.c_int]
        lib.qppvm_kernel_timing.argtypes = [P, C.c_int, P, P]
        lib.qppvm_multi_create.argtypes = [C.POINTER(CDesc), C.POINTER(C.c_int32), C.c_int, C.POINTER(P)]
        lib.qppvm_multi_destroy.argtypes = [P]
        lib.qppvm_multi_last_error.argtypes = [P]
        lib.qppvm_multi_last_error.restype = C.c_char_p
        lib.qppvm_multi_devices.argtypes = [P]
        lib.qppvm_multi_set_robot.argtypes = [P, C.POINTER(CRobot)]
        lib.qppvm_multi_solve_batch.argtypes = [P, P, P, C.c_int64]
        lib.qppvm_multi_solve_batch_host.argtypes = [P, P, P, C.c_int64]
        lib.qppvm_multi_solve_states.argtypes = [P, P, P, C.c_int64]
        lib.qppvm_multi_solve_states_host.argtypes = [P, P, P, C.c_int64]
        lib.qppvm_multi_kernel_launches.argtypes = [P]
        lib.qppvm_multi_kernel_launches.restype = C.c_int64
        lib.qppvm_multi_nccl_calls.argtypes = [P]
        lib.qppvm_multi_nccl_calls.restype = C.c_int64
        lib.qppvm_tick_stamps.argtypes = [P, P]
        lib.qppvm_solve_batch_host_async.argtypes = [P, P, P, C.c_int64]
        lib.qppvm_host_sync.argtypes = [P]
        lib.qppvm_solve_states_host_async.argtypes = [P, P, P, C.c_int64]
        lib.qppvm_integrate_states.argtypes = [P, P, P, C.c_double, C.c_int64, P]
        lib.qppvm_integrate_states_tracking.argtypes = [P, P, P, P, C.c_double, C.c_int64, P]
        lib.qppvm_rollout_states.argtypes = [P, P, P, C.c_int, C.c_double, C.c_int64, P]
        lib.qppvm_kernel_launches.argtypes = [P]
        lib.qppvm_kernel_launches.restype = C.c_int64
        lib.qppvm_fp64_peak.argtypes = [P, C.POINTER(C.c_double)]
        lib.qppvm_supported_shapes.argtypes = [C.POINTER(C.c_int32), C.c_int]
        lib.qppvm_state_doubles.argtypes = [C.POINTER(CDesc)]
        lib.qppvm_set_robot.argtypes = [P, C.POINTER(CRobot)]
        lib.qppvm_records_from_states.argtypes = [P, P, P, C.c_int64, P]
        lib.qppvm_solve_states_host.argtypes = [P, P, P, C.c_int64]
        _lib = lib
    return _lib


def cdesc(desc: Desc) -> CDesc:
    return CDesc(desc.kind, desc.n_a, desc.n_contacts, desc.flags, desc.eps_regularisation,
                 desc.n_reg_steps, desc.max_iter, desc.device, desc.postural_actuated_only, desc.lambda_solver,
                 (C.c_double * 3)(*desc.task_weight))


def c_layout(desc: Desc) -> dict:
    L = CLayout()
    if load_library().qppvm_get_layout(C.byref(cdesc(desc)), C.byref(L)):
        raise ValueError("qppvm_get_layout rejected the description")
    return {f: getattr(L, f) for f in Layout.FIELDS}


def supported_shapes():
    lib = load_library()
    n = lib.qppvm_supported_shapes(None, 0)
    buf = (C.c_int32 * (4 * n))()
    lib.qppvm_supported_shapes(buf, n)
    return [tuple(buf[4 * i:4 * i + 4]) for i in range(n)]


class QPError(RuntimeError):
    pass


class Solver:
    """One ``qppvm_handle``: batched / single-tick whole-body QP solves for one problem shape."""

    def __init__(self, desc: Desc):
        self.desc = desc
        self.layout = layout(desc)
        self._lib = load_library()
        self._h = C.c_void_p()
        rc = self._lib.qppvm_create(C.byref(cdesc(desc)), C.byref(self._h))
        if rc:
            raise QPError("qppvm_create failed (%d): %s" % (rc, self._lib.qppvm_last_error(None).decode()))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.qppvm_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise QPError("qppvm call failed (%d): %s" % (rc, self._lib.qppvm_last_error(self._h).decode()))

    # ---- device path (torch tensors are only the memory/stream plumbing) -------------------
    def solve_batch(self, records, out=None, diag=None, stream=None):
        """records: CUDA float64 tensor (B, rec_doubles).  Returns (out, diag): out is a CUDA float64
        tensor (B, out_bytes/8) (trailer bit-packed in the last 4 doubles); asynchronous on `stream`."""
        import torch
        L = self.layout
        assert records.is_cuda and records.dtype == torch.float64 and records.is_contiguous()
        assert records.shape[1] == L.rec_doubles
        B = records.shape[0]
        if out is None:
            out = torch.empty((B, L.out_doubles), dtype=torch.float64, device=records.device)
        if diag is False:
            diag = None
        if diag is True:
            diag = torch.empty((B, L.diag_doubles), dtype=torch.float64, device=records.device)
        st = torch.cuda.current_stream(records.device) if stream is None else stream
        self._check(self._lib.qppvm_solve_batch_diag(
            self._h, records.data_ptr(), out.data_ptr(), diag.data_ptr() if diag is not None else None,
            B, st.cuda_stream))
        return out, diag

    def solve_batch_warm(self, records, warm, out=None, stream=None):
        """Hot-started batched solve: `warm` is a CUDA int32 tensor (B, 8), the working sets of the previous tick
        (in/out; zeros = cold start)."""
        import torch
        L = self.layout
        assert records.is_cuda and records.dtype == torch.float64 and records.is_contiguous() and records.shape[1] == L.rec_doubles
        B = records.shape[0]
        assert warm.is_cuda and warm.dtype == torch.int32 and warm.is_contiguous() and warm.shape == (B, 8)
        if out is None:
            out = torch.empty((B, L.out_doubles), dtype=torch.float64, device=records.device)
        st = torch.cuda.current_stream(records.device) if stream is None else stream
        self._check(self._lib.qppvm_solve_batch_warm(self._h, records.data_ptr(), out.data_ptr(), warm.data_ptr(), B, st.cuda_stream))
        return out

    # ---- host path: the reference-facing call (host buffers in, host buffers out) ----------
    def solve_batch_host(self, records: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        L = self.layout
        assert records.dtype == np.float64 and records.flags.c_contiguous and records.shape[1] == L.rec_doubles
        B = records.shape[0]
        if out is None:
            out = np.empty((B, L.out_doubles))
        self._check(self._lib.qppvm_solve_batch_host(self._h, records.ctypes.data, out.ctypes.data, B))
        return out

    def solve_batch_host_ptr(self, rec_ptr: int, out_ptr: int, batch: int):
        self._check(self._lib.qppvm_solve_batch_host(self._h, rec_ptr, out_ptr, batch))

    def solve_batch_host_async_ptr(self, rec_ptr: int, out_ptr: int, batch: int):
        """Pipelined host path (pinned buffers): returns after enqueueing; read `out` only after host_sync()."""
        self._check(self._lib.qppvm_solve_batch_host_async(self._h, rec_ptr, out_ptr, batch))

    def host_sync(self):
        self._check(self._lib.qppvm_host_sync(self._h))

    def solve_one(self, record: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        L = self.layout
        assert record.dtype == np.float64 and record.flags.c_contiguous and record.size == L.rec_doubles
        if out is None:
            out = np.empty(L.out_doubles)
        self._check(self._lib.qppvm_solve_one(self._h, record.ctypes.data, out.ctypes.data))
        return out

    def kernel_timing(self, enable: bool):
        """(ms[3], launches[3]) of the prepare / solve / certify kernels since the previous call; switches recording on/off."""
        ms = np.zeros(3); n = np.zeros(3, dtype=np.int64)
        self._check(self._lib.qppvm_kernel_timing(self._h, int(enable), ms.ctypes.data, n.ctypes.data))
        return ms, n

    def tick_stamps(self) -> np.ndarray:
        """Device-clock stamps (ns) of the last tick through the resident chain (7 stage boundaries)."""
        a = np.zeros(7, dtype=np.uint64)
        self._check(self._lib.qppvm_tick_stamps(self._h, a.ctypes.data))
        return a

    def reset_warm(self):
        """Forget the working sets of the previous tick: the next solve_one is a cold start."""
        self._check(self._lib.qppvm_reset_warm(self._h))

    # ---- rigid-body front end: compact states instead of records -------
    def set_robot(self, robot, contact_bodies):
        """robot: qppvm_b200.gen.Robot (kinematic tree tables); contact_bodies: body index per contact."""
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        keep = dict(parent=i32(robot.parent), axis=f64(robot.axis), offset=f64(robot.offset), mass=f64(robot.mass),
                    com=f64(robot.com), inertia=f64(robot.inertia), q_home=f64(robot.q_home), tau_max=f64(robot.tau_max),
                    contact=i32(contact_bodies))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        r = CRobot(robot.n_a, ip(keep["parent"]), dp(keep["axis"]), dp(keep["offset"]), dp(keep["mass"]), dp(keep["com"]),
                   dp(keep["inertia"]), dp(keep["q_home"]), dp(keep["tau_max"]), ip(keep["contact"]))
        self._check(self._lib.qppvm_set_robot(self._h, C.byref(r)))

    @property
    def state_doubles(self) -> int:
        return self._lib.qppvm_state_doubles(C.byref(cdesc(self.desc)))

    def records_from_states(self, states, records=None, stream=None):
        import torch
        assert states.is_cuda and states.dtype == torch.float64 and states.is_contiguous()
        assert states.shape[1] == self.state_doubles
        B = states.shape[0]
        if records is None:
            records = torch.empty((B, self.layout.rec_doubles), dtype=torch.float64, device=states.device)
        st = torch.cuda.current_stream(states.device) if stream is None else stream
        self._check(self._lib.qppvm_records_from_states(self._h, states.data_ptr(), records.data_ptr(), B, st.cuda_stream))
        return records

    def integrate_states(self, states, out, dt: float, stream=None, records=None):
        """In place: advance `states` (cuda float64) by dt with the accelerations in `out` (solve_batch's block); with the
        tick's `records` the stored task errors follow the motion (closed loop in the references)."""
        import torch
        assert states.is_cuda and states.dtype == torch.float64 and states.is_contiguous() and out.is_contiguous()
        assert states.shape[1] == self.state_doubles and out.shape == (states.shape[0], self.layout.out_doubles)
        st = torch.cuda.current_stream(states.device) if stream is None else stream
        if records is None:
            self._check(self._lib.qppvm_integrate_states(self._h, states.data_ptr(), out.data_ptr(), dt, states.shape[0], st.cuda_stream))
        else:
            assert records.is_cuda and records.is_contiguous() and records.shape == (states.shape[0], self.layout.rec_doubles)
            self._check(self._lib.qppvm_integrate_states_tracking(self._h, states.data_ptr(), out.data_ptr(), records.data_ptr(),
                                                                  dt, states.shape[0], st.cuda_stream))
        return states

    def rollout_states(self, states, ticks: int, dt: float, out=None, stream=None):
        """`ticks` control periods of front end -> solve -> integrate on the device; `states` is updated in place,
        the returned block holds the last tick's solutions."""
        import torch
        assert states.is_cuda and states.dtype == torch.float64 and states.is_contiguous()
        assert states.shape[1] == self.state_doubles
        if out is None:
            out = torch.empty((states.shape[0], self.layout.out_doubles), dtype=torch.float64, device=states.device)
        st = torch.cuda.current_stream(states.device) if stream is None else stream
        self._check(self._lib.qppvm_rollout_states(self._h, states.data_ptr(), out.data_ptr(), ticks, dt, states.shape[0], st.cuda_stream))
        return out

    def solve_states_host_ptr(self, states_ptr: int, out_ptr: int, batch: int):
        self._check(self._lib.qppvm_solve_states_host(self._h, states_ptr, out_ptr, batch))

    def solve_states_host_async_ptr(self, states_ptr: int, out_ptr: int, batch: int):
        self._check(self._lib.qppvm_solve_states_host_async(self._h, states_ptr, out_ptr, batch))

    def solve_states_host(self, states: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        assert states.dtype == np.float64 and states.flags.c_contiguous and states.shape[1] == self.state_doubles
        if out is None:
            out = np.empty((states.shape[0], self.layout.out_doubles))
        self._check(self._lib.qppvm_solve_states_host(self._h, states.ctypes.data, out.ctypes.data, states.shape[0]))
        return out

    @property
    def kernel_launches(self) -> int:
        return self._lib.qppvm_kernel_launches(self._h)

    def fp64_peak_tflops(self) -> float:
        v = C.c_double(0)
        self._check(self._lib.qppvm_fp64_peak(self._h, C.byref(v)))
        return v.value


def _crobot(robot, contact_bodies):
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    keep = dict(parent=i32(robot.parent), axis=f64(robot.axis), offset=f64(robot.offset), mass=f64(robot.mass),
                com=f64(robot.com), inertia=f64(robot.inertia), q_home=f64(robot.q_home), tau_max=f64(robot.tau_max),
                contact=i32(contact_bodies))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    r = CRobot(robot.n_a, ip(keep["parent"]), dp(keep["axis"]), dp(keep["offset"]), dp(keep["mass"]), dp(keep["com"]),
               dp(keep["inertia"]), dp(keep["q_home"]), dp(keep["tau_max"]), ip(keep["contact"]))
    return r, keep


class MultiSolver:
    """One ``qppvm_multi``: the batch sharded over several GPUs of this box from ONE process (block split, NCCL
    scatter / gather pipelined against the solves; see include/qppvm_b200.h)."""

    def __init__(self, desc: Desc, devices=None):
        import torch
        self.desc, self.layout, self._lib = desc, layout(desc), load_library()
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        self.devices = list(devices)
        arr = (C.c_int32 * len(self.devices))(*self.devices)
        self._m = C.c_void_p()
        rc = self._lib.qppvm_multi_create(C.byref(cdesc(desc)), arr, len(self.devices), C.byref(self._m))
        if rc:
            raise QPError("qppvm_multi_create failed (%d): %s" % (rc, self._lib.qppvm_multi_last_error(None).decode()))

    def close(self):
        m, self._m = getattr(self, "_m", None), None
        if m:
            self._lib.qppvm_multi_destroy(m)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise QPError("qppvm_multi call failed (%d): %s" % (rc, self._lib.qppvm_multi_last_error(self._m).decode()))

    def set_robot(self, robot, contact_bodies):
        r, _keep = _crobot(robot, contact_bodies)
        self._check(self._lib.qppvm_multi_set_robot(self._m, C.byref(r)))

    def solve_batch(self, records, out=None):
        """records: float64 CUDA tensor on the ROOT device (devices[0]); synchronous."""
        import torch
        L = self.layout
        assert records.is_cuda and records.device.index == self.devices[0] and records.dtype == torch.float64
        assert records.is_contiguous() and records.shape[1] == L.rec_doubles
        if out is None:
            out = torch.empty((records.shape[0], L.out_doubles), dtype=torch.float64, device=records.device)
        torch.cuda.current_stream(records.device).synchronize()
        self._check(self._lib.qppvm_multi_solve_batch(self._m, records.data_ptr(), out.data_ptr(), records.shape[0]))
        return out

    def solve_states(self, states, out=None):
        """states: float64 CUDA tensor (batch, state_doubles) on the ROOT device; the states are what NCCL scatters,
        every GPU runs the rigid-body front end on its chunks (after set_robot); synchronous."""
        import torch
        L = self.layout
        assert states.is_cuda and states.device.index == self.devices[0] and states.dtype == torch.float64 and states.is_contiguous()
        if out is None:
            out = torch.empty((states.shape[0], L.out_doubles), dtype=torch.float64, device=states.device)
        torch.cuda.current_stream(states.device).synchronize()
        self._check(self._lib.qppvm_multi_solve_states(self._m, states.data_ptr(), out.data_ptr(), states.shape[0]))
        return out

    def solve_batch_host_ptr(self, rec_ptr: int, out_ptr: int, batch: int):
        self._check(self._lib.qppvm_multi_solve_batch_host(self._m, rec_ptr, out_ptr, batch))

    def solve_states_host_ptr(self, st_ptr: int, out_ptr: int, batch: int):
        self._check(self._lib.qppvm_multi_solve_states_host(self._m, st_ptr, out_ptr, batch))

    @property
    def kernel_launches(self) -> int:
        return self._lib.qppvm_multi_kernel_launches(self._m)

    @property
    def nccl_calls(self) -> int:
        return self._lib.qppvm_multi_nccl_calls(self._m)


def split_out(L: Layout, out: np.ndarray) -> dict:
    """out (B, out_doubles) float64 -> x, tau, status, iters0, iters1, active (B,4) u32, kkt (B,2) f32."""
    n, na = L.n_x, L.n_a
    out = np.ascontiguousarray(out)
    tr = out[:, n + na:n + na + 4].copy().view(np.uint8).reshape(out.shape[0], 32)
    i32 = tr[:, 0:8].copy().view(np.int32)
    return dict(x=out[:, :n], tau=out[:, n:n + na], status=i32[:, 0], iters0=i32[:, 1] & 0xffff,
                iters1=(i32[:, 1] >> 16) & 0xffff, active=tr[:, 8:24].copy().view(np.uint32),
                kkt=tr[:, 24:32].copy().view(np.float32))


def split_diag(L: Layout, diag: np.ndarray) -> dict:
    n, nr = L.n_x, L.n_rows
    return dict(x0=diag[:, :n], y0=diag[:, n:n + nr], y1=diag[:, n + nr:n + 2 * nr], eopt=diag[:, n + 2 * nr:])
