"""ctypes binding of libqgt_b200.so (include/qgt_b200.h) used by the tests and bench.py.

This is harness code: the product is the C-ABI shared library and the C compatibility layer in
``csrc/compat``; a reference user links those from C.  The binding fails loudly
when the library is missing or when no sm_100 device is present — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import weakref
from typing import Optional, Tuple

import numpy as np

from .circuits import CCircuit, Circuit

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqgt_b200.so")
_DP = C.POINTER(C.c_double)


class QgtError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"qgt_b200 status {status}: {msg}")
        self.status = status


class Stats(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_sweep", C.c_double), ("ms_gram", C.c_double), ("ms_other", C.c_double),
                ("sweep_bytes", C.c_double), ("gram_flops", C.c_double), ("gram_bytes", C.c_double), ("exchange_bytes", C.c_double),
                ("sweep_launches", C.c_int64), ("gram_launches", C.c_int64), ("other_launches", C.c_int64),
                ("sweep_column_passes", C.c_int64),
                ("num_runs", C.c_int32), ("resident_columns", C.c_int32), ("blocks", C.c_int32), ("tile_qubits", C.c_int32),
                ("ms_wall", C.c_double), ("ms_host_plan", C.c_double), ("tensor_flops", C.c_double),
                ("fused", C.c_int32), ("fused_launches", C.c_int32), ("ms_exchange", C.c_double)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


class NatGradConfig(C.Structure):
    _fields_ = [("regularization", C.c_double), ("condition_threshold", C.c_double),
                ("adaptive", C.c_int), ("pseudoinverse_fallback", C.c_int), ("singular_cutoff", C.c_double)]


_lib = None


def load() -> C.CDLL:
    """Load libqgt_b200.so; raises if it was not built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not built: run __graft_entry__.build()")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.qgt_b200_last_error.restype = C.c_char_p
    L.qgt_b200_error_string.restype = C.c_char_p
    L.qgt_b200_error_string.argtypes = [C.c_int]
    L.qgt_b200_create.argtypes = [C.POINTER(vp), C.c_int]
    L.qgt_b200_destroy.argtypes = [vp]
    L.qgt_b200_destroy.restype = None
    L.qgt_b200_set_workspace_limit.argtypes = [vp, C.c_size_t]
    L.qgt_b200_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.qgt_b200_state_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.qgt_b200_state_destroy.argtypes = [vp]
    L.qgt_b200_state_destroy.restype = None
    L.qgt_b200_state_init.argtypes = [vp, C.c_int]
    L.qgt_b200_state_upload.argtypes = [vp, _DP]
    L.qgt_b200_state_download.argtypes = [vp, _DP]
    L.qgt_b200_state_norm2.argtypes = [vp, _DP]
    L.qgt_b200_state_device_ptr.argtypes = [vp]
    L.qgt_b200_state_device_ptr.restype = vp
    L.qgt_b200_apply_circuit.argtypes = [vp, C.POINTER(CCircuit), _DP]
    L.qgt_b200_simulate_host.argtypes = [vp, _DP, C.c_int, C.POINTER(CCircuit), _DP]
    L.qgt_b200_qgt.argtypes = [vp, C.POINTER(CCircuit), _DP, _DP, _DP, _DP, vp]
    L.qgt_b200_gram.argtypes = [vp, vp, vp, C.c_size_t, C.c_size_t, _DP, _DP, _DP]
    L.qgt_b200_derivative.argtypes = [vp, C.POINTER(CCircuit), _DP, C.c_int, vp]
    L.qgt_b200_natural_gradient.argtypes = [vp, _DP, _DP, C.c_size_t, C.POINTER(NatGradConfig), _DP, _DP]
    L.qgt_b200_expectation_gradient.argtypes = [vp, C.POINTER(CCircuit), _DP, _DP, _DP]
    L.qgt_b200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.qgt_b200_plan_dump.argtypes = [C.POINTER(CCircuit), _DP, C.c_int, C.c_int, C.c_size_t, C.c_char_p, C.c_size_t]
    L.qgt_b200_plan_dump.restype = C.c_long
    L.qgt_b200_dist_unique_id.argtypes = [C.c_char_p]
    L.qgt_b200_dist_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.qgt_b200_dist_world.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.qgt_b200_plan_dump_sharded.argtypes = [C.POINTER(CCircuit), _DP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_char_p, C.c_size_t]
    L.qgt_b200_plan_dump_sharded.restype = C.c_long
    L.qgt_b200_dist_barrier.argtypes = [vp]
    L.qgt_b200_measure_peaks.argtypes = [vp, _DP, _DP]
    u64 = C.c_uint64
    L.qgt_b200_state_probability.argtypes = [vp, u64, u64, _DP, _DP]
    L.qgt_b200_state_expectation_z.argtypes = [vp, u64, _DP]
    L.qgt_b200_state_inner_product.argtypes = [vp, vp, _DP]
    L.qgt_b200_state_scale.argtypes = [vp, C.c_double, C.c_double]
    L.qgt_b200_state_normalize.argtypes = [vp, _DP]
    L.qgt_b200_state_collapse.argtypes = [vp, C.c_int, C.c_int, _DP]
    L.qgt_b200_state_measure.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int), _DP]
    L.qgt_b200_state_sample.argtypes = [vp, _DP, C.c_size_t, C.POINTER(u64)]
    L.qgt_b200_state_upload_c64.argtypes = [vp, C.c_void_p]
    L.qgt_b200_state_download_c64.argtypes = [vp, C.c_void_p]
    L.qgt_b200_plan_dump_fused.argtypes = [C.POINTER(CCircuit), _DP, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_char_p, C.c_size_t]
    L.qgt_b200_plan_dump_fused.restype = C.c_long
    L.qgt_b200_plan_dump_gradient.argtypes = [C.POINTER(CCircuit), _DP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    L.qgt_b200_plan_dump_gradient.restype = C.c_long
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise QgtError(rc, load().qgt_b200_last_error().decode(errors="replace"))


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_DP)


def device_count() -> int:
    return int(load().qgt_b200_device_count())


def plan_dump(circ: Circuit, theta: Optional[np.ndarray] = None, tile_qubits: int = 0, reg_qubits: int = 0,
              column_slots: int = 0) -> dict:
    """Fused-run plan (and column schedule when ``column_slots`` > 0) as a dict.  Needs no GPU."""
    L = load()
    cc = circ.to_c()
    th = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
    n = L.qgt_b200_plan_dump(C.byref(cc), _dp(th), tile_qubits, reg_qubits, column_slots, None, 0)
    if n < 0:
        _check(int(n))
    buf = C.create_string_buffer(n + 1)
    n2 = L.qgt_b200_plan_dump(C.byref(cc), _dp(th), tile_qubits, reg_qubits, column_slots, buf, n + 1)
    if n2 < 0:
        _check(int(n2))
    return json.loads(buf.value.decode())


def plan_dump_sharded(circ: Circuit, theta: Optional[np.ndarray], world: int, restore_identity: bool = True,
                      tile_qubits: int = 0, reg_qubits: int = 0, column_slots: int = 0) -> dict:
    """Plan of a state sharded over ``world`` ranks: runs, EXCHANGE pseudo-runs, qubit maps.  Needs no GPU."""
    L = load()
    cc = circ.to_c()
    th = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
    args = (C.byref(cc), _dp(th), world, 1 if restore_identity else 0, tile_qubits, reg_qubits, column_slots)
    n = L.qgt_b200_plan_dump_sharded(*args, None, 0)
    if n < 0:
        _check(int(n))
    buf = C.create_string_buffer(n + 1)
    n2 = L.qgt_b200_plan_dump_sharded(*args, buf, n + 1)
    if n2 < 0:
        _check(int(n2))
    return json.loads(buf.value.decode())


def plan_dump_fused(circ: Circuit, theta: Optional[np.ndarray], column_slots: int, world: int = 1,
                    tile_qubits: int = 0, reg_qubits: int = 0) -> dict:
    """Plan with the fused column schedule (transition matrices instead of Gram passes).  Needs no GPU."""
    L = load()
    cc = circ.to_c()
    th = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
    args = (C.byref(cc), _dp(th), world, tile_qubits, reg_qubits, column_slots)
    n = L.qgt_b200_plan_dump_fused(*args, None, 0)
    if n < 0:
        _check(int(n))
    buf = C.create_string_buffer(n + 1)
    n2 = L.qgt_b200_plan_dump_fused(*args, buf, n + 1)
    if n2 < 0:
        _check(int(n2))
    return json.loads(buf.value.decode())


def plan_dump_gradient(circ: Circuit, theta: Optional[np.ndarray], fused: bool, scratch_slots: int = 2,
                       tile_qubits: int = 0, reg_qubits: int = 0, world: int = 1) -> dict:
    """Plan of the inverse circuit with the adjoint-gradient program (fused or generic).  Needs no GPU."""
    L = load()
    cc = circ.to_c()
    th = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
    args = (C.byref(cc), _dp(th), 1 if fused else 0, scratch_slots, tile_qubits, reg_qubits, world)
    n = L.qgt_b200_plan_dump_gradient(*args, None, 0)
    if n < 0:
        _check(int(n))
    buf = C.create_string_buffer(n + 1)
    n2 = L.qgt_b200_plan_dump_gradient(*args, buf, n + 1)
    if n2 < 0:
        _check(int(n2))
    return json.loads(buf.value.decode())


class State:
    def __init__(self, ctx: "Context", num_qubits: int):
        self.ctx = ctx
        self.num_qubits = num_qubits
        self.h = C.c_void_p()
        _check(ctx.L.qgt_b200_state_create(ctx.h, num_qubits, C.byref(self.h)))
        ctx._states.add(self)

    def init(self, initial_state: int = 0) -> "State":
        _check(self.ctx.L.qgt_b200_state_init(self.h, initial_state))
        return self

    def upload(self, amps: np.ndarray) -> "State":
        a = np.ascontiguousarray(amps, dtype=np.complex128)
        _check(self.ctx.L.qgt_b200_state_upload(self.h, _dp(a)))
        return self

    def download(self, local_len: Optional[int] = None) -> np.ndarray:
        out = np.empty(local_len if local_len else (1 << self.num_qubits) // self.ctx.world, dtype=np.complex128)
        _check(self.ctx.L.qgt_b200_state_download(self.h, _dp(out)))
        return out

    def norm2(self) -> float:
        v = C.c_double(0)
        _check(self.ctx.L.qgt_b200_state_norm2(self.h, C.byref(v)))
        return v.value

    # -- reductions (device kernels) ---------------------------------------------------------------
    def probability(self, mask: int, want: int) -> float:
        v = C.c_double(0)
        _check(self.ctx.L.qgt_b200_state_probability(self.h, mask, want, C.byref(v), None))
        return v.value

    def expectation_z(self, zmask: int) -> float:
        v = C.c_double(0)
        _check(self.ctx.L.qgt_b200_state_expectation_z(self.h, zmask, C.byref(v)))
        return v.value

    def inner(self, other: "State") -> complex:
        out = np.zeros(2)
        _check(self.ctx.L.qgt_b200_state_inner_product(self.h, other.h, _dp(out)))
        return complex(out[0], out[1])

    def scale(self, z: complex) -> "State":
        _check(self.ctx.L.qgt_b200_state_scale(self.h, float(np.real(z)), float(np.imag(z))))
        return self

    def measure(self, qubit: int, uniform: float, readout_error: float = 0.0):
        o, p = C.c_int(0), C.c_double(0)
        _check(self.ctx.L.qgt_b200_state_measure(self.h, qubit, uniform, readout_error, C.byref(o), C.byref(p)))
        return o.value, p.value

    def sample(self, uniforms: np.ndarray) -> np.ndarray:
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        out = np.zeros(u.size, dtype=np.uint64)
        _check(self.ctx.L.qgt_b200_state_sample(self.h, _dp(u), u.size, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def argmax(self):
        i, p = C.c_uint64(0), C.c_double(0)
        _check(self.ctx.L.qgt_b200_state_argmax(self.h, C.byref(i), C.byref(p)))
        return int(i.value), p.value

    def upload_c64(self, amps: np.ndarray) -> "State":
        a = np.ascontiguousarray(amps, dtype=np.complex64)
        _check(self.ctx.L.qgt_b200_state_upload_c64(self.h, a.ctypes.data))
        return self

    def download_c64(self) -> np.ndarray:
        out = np.empty((1 << self.num_qubits) // self.ctx.world, dtype=np.complex64)
        _check(self.ctx.L.qgt_b200_state_download_c64(self.h, out.ctypes.data))
        return out

    def apply(self, circ: Circuit, theta: np.ndarray) -> "State":
        cc = circ.to_c()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        _check(self.ctx.L.qgt_b200_apply_circuit(self.h, C.byref(cc), _dp(th)))
        return self

    def device_ptr(self) -> int:
        return int(self.ctx.L.qgt_b200_state_device_ptr(self.h) or 0)

    def close(self) -> None:
        if self.h and self.ctx.h:
            self.ctx.L.qgt_b200_state_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One per process and GPU (qgt_b200_create)."""

    def __init__(self, device: int = 0):
        self.L = load()
        self.h = C.c_void_p()
        self.world = 1
        self.rank = 0
        self._states = weakref.WeakSet()     # states must be destroyed before their context
        _check(self.L.qgt_b200_create(C.byref(self.h), device))

    def dist_init(self, rank: int, world: int, unique_id: bytes) -> None:
        """Join the NCCL communicator (collective); ``unique_id`` comes from rank 0's ``dist_unique_id()``."""
        assert len(unique_id) == 128
        _check(self.L.qgt_b200_dist_init(self.h, rank, world, unique_id))
        self.rank, self.world = rank, world

    @staticmethod
    def dist_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(load().qgt_b200_dist_unique_id(buf))
        return buf.raw

    def barrier(self) -> None:
        _check(self.L.qgt_b200_dist_barrier(self.h))

    def set_option(self, key: str, value: float) -> None:
        _check(self.L.qgt_b200_set_option(self.h, key.encode(), float(value)))

    def set_workspace_limit(self, nbytes: int) -> None:
        _check(self.L.qgt_b200_set_workspace_limit(self.h, int(nbytes)))

    def state(self, num_qubits: int) -> State:
        return State(self, num_qubits)

    def simulate_host(self, amps: np.ndarray, circ: Circuit, theta: np.ndarray) -> np.ndarray:
        a = np.array(amps, dtype=np.complex128)
        cc = circ.to_c()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        _check(self.L.qgt_b200_simulate_host(self.h, _dp(a), circ.num_qubits, C.byref(cc), _dp(th)))
        return a

    def qgt(self, circ: Circuit, theta: np.ndarray, psi_out: Optional[State] = None) -> np.ndarray:
        """Full Q (P x P complex128); metric = Q.real, Berry curvature = Q.imag."""
        P = circ.num_params
        q = np.zeros((P, P), dtype=np.complex128)
        cc = circ.to_c()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        _check(self.L.qgt_b200_qgt(self.h, C.byref(cc), _dp(th), None, None, _dp(q), psi_out.h if psi_out else None))
        return q

    def qgt_metric_berry(self, circ: Circuit, theta: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        P = circ.num_params
        g = np.zeros((P, P))
        b = np.zeros((P, P))
        cc = circ.to_c()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        _check(self.L.qgt_b200_qgt(self.h, C.byref(cc), _dp(th), _dp(g), _dp(b), None, None))
        return g, b

    def gram(self, psi: np.ndarray, dpsi: np.ndarray) -> np.ndarray:
        P, dim = dpsi.shape
        psi = np.ascontiguousarray(psi, dtype=np.complex128)
        dpsi = np.ascontiguousarray(dpsi, dtype=np.complex128)
        q = np.zeros((P, P), dtype=np.complex128)
        _check(self.L.qgt_b200_gram(self.h, psi.ctypes.data, dpsi.ctypes.data, dim, P, None, None, _dp(q)))
        return q

    def derivative(self, circ: Circuit, theta: np.ndarray, mu: int) -> np.ndarray:
        st = self.state(circ.num_qubits)
        cc = circ.to_c()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        _check(self.L.qgt_b200_derivative(self.h, C.byref(cc), _dp(th), mu, st.h))
        out = st.download()
        st.close()
        return out

    def natural_gradient(self, metric: np.ndarray, grad: np.ndarray, cfg: Optional[NatGradConfig] = None) -> Tuple[np.ndarray, float]:
        m = np.ascontiguousarray(metric, dtype=np.float64)
        g = np.ascontiguousarray(grad, dtype=np.float64)
        out = np.zeros_like(g)
        lam = C.c_double(0)
        _check(self.L.qgt_b200_natural_gradient(self.h, _dp(m), _dp(g), g.size, C.byref(cfg) if cfg else None, _dp(out), C.byref(lam)))
        return out, lam.value

    def expectation_gradient(self, circ: Circuit, theta: np.ndarray) -> Tuple[float, np.ndarray]:
        cc = circ.to_c()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        e = C.c_double(0)
        g = np.zeros(circ.num_params)
        _check(self.L.qgt_b200_expectation_gradient(self.h, C.byref(cc), _dp(th), C.byref(e), _dp(g)))
        return e.value, g

    def natural_gradient_descent(self, circ: Circuit, theta: np.ndarray, iterations: int, learning_rate: float) -> Tuple[np.ndarray, np.ndarray]:
        """`iterations` natural-gradient steps on the device; returns (theta, energy history of length iterations + 1)."""
        cc = circ.to_c()
        th = np.array(theta, dtype=np.float64, copy=True)
        hist = np.zeros(iterations + 1)
        self.L.qgt_b200_natural_gradient_descent.argtypes = [C.c_void_p, C.POINTER(CCircuit), _DP, C.c_int, C.c_double, C.c_void_p, _DP]
        _check(self.L.qgt_b200_natural_gradient_descent(self.h, C.byref(cc), _dp(th), iterations, learning_rate, None, _dp(hist)))
        return th, hist

    def collective(self, kind: int, send: Optional[np.ndarray], recv: np.ndarray, count: int, dtype: int, op: int = 0, root: int = 0) -> None:
        """qgt_b200_dist_collective on host arrays (kind: 0 broadcast, 1 allreduce, 2 scatter, 3 gather, 4 allgather, 5 reduce_scatter)."""
        self.L.qgt_b200_dist_collective.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int]
        sp = None if send is None else send.ctypes.data_as(C.c_void_p)
        _check(self.L.qgt_b200_dist_collective(self.h, kind, sp, recv.ctypes.data_as(C.c_void_p), count, dtype, op, root))

    def measure_peaks(self) -> dict:
        """FP64 tensor-pipe peak (TFLOP/s) and D2D copy bandwidth (GB/s) measured on this device."""
        a, b = C.c_double(0), C.c_double(0)
        _check(self.L.qgt_b200_measure_peaks(self.h, C.byref(a), C.byref(b)))
        return {"dmma_tflops": a.value, "copy_gbs": b.value}

    def stats(self) -> dict:
        s = Stats()
        _check(self.L.qgt_b200_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def close(self) -> None:
        if self.h:
            for st in list(self._states):
                st.close()
            self.L.qgt_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
