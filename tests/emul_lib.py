"""Loader for the host emulation of the sweep kernel (tests/emul/emul.cpp).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from quantum_geometric_tensor_b200.circuits import CCircuit

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_DP = C.POINTER(C.c_double)
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    out = os.path.join(_HERE, "emul", "_build", "libqgt_emul.so")
    srcs = [os.path.join(_HERE, "emul", "emul.cpp"),
            os.path.join(_ROOT, "quantum_geometric_tensor_b200", "csrc", "plan.cpp")]
    deps = srcs + [os.path.join(_ROOT, "quantum_geometric_tensor_b200", "csrc", f)
                   for f in ("plan.hpp", "sweep_core.cuh", "dev_structs.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out] + srcs, check=True)
    L = C.CDLL(out)
    L.emul_sweep.argtypes = [C.POINTER(CCircuit), _DP, C.c_int, C.c_int, C.c_int, _DP, _DP, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
    L.emul_num_runs.argtypes = [C.POINTER(CCircuit), _DP, C.c_int, C.c_int]
    _lib = L
    return L


def sweep(circ, theta, K, R, run_idx, src, ovr_op=-1, dst=None, accumulate=False, extra=()):
    L = load()
    cc = circ.to_c()
    th = np.ascontiguousarray(theta if len(theta) else np.zeros(1), dtype=np.float64)
    src = np.ascontiguousarray(src, dtype=np.complex128)
    out = np.zeros_like(src) if dst is None else np.array(dst, dtype=np.complex128)
    rc = L.emul_sweep(C.byref(cc), th.ctypes.data_as(_DP), K, R, run_idx, src.ctypes.data_as(_DP),
                      out.ctypes.data_as(_DP), ovr_op, 1 if accumulate else 0, (C.c_int * max(1, len(extra)))(*extra), len(extra))
    if rc:
        raise RuntimeError(f"emul_sweep -> {rc}")
    return out
