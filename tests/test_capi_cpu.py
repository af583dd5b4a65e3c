"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/qgt_b200.h
declares, refuses to compute without a device (no CPU fallback) and its host-only parts work."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from quantum_geometric_tensor_b200 import api, circuits as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qgt_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = api.load()
    names = _declared_functions(os.path.join(ROOT, "include", "qgt_b200.h"))
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.qgt_b200_abi_version() == 3


def test_no_cpu_fallback_without_device():
    lib = api.load()
    if lib.qgt_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.qgt_b200_create(C.byref(h), 0)
    assert rc == -31 and not h.value          # QGT_B200_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.qgt_b200_last_error()
    with pytest.raises(api.QgtError):
        api.Context(0)


def test_struct_layouts_match_header(tmp_path):
    # compile a C probe against the real header and compare with the ctypes mirrors
    import subprocess
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include "qgt_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(qgt_b200_gate),sizeof(qgt_b200_edge),sizeof(qgt_b200_circuit),'
                   'sizeof(qgt_b200_natgrad_config),sizeof(qgt_b200_stats));return 0;}\n')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(K.CGate), C.sizeof(K.CEdge), C.sizeof(K.CCircuit), C.sizeof(api.NatGradConfig),
                     C.sizeof(api.Stats)]


def test_host_natural_gradient_matches_oracle(oracle):
    lib = api.load()
    rng = np.random.default_rng(1)
    A = rng.normal(size=(20, 20))
    G = A @ A.T / 20
    g = rng.normal(size=20)
    out = np.zeros(20)
    lam = C.c_double(0)
    dp = C.POINTER(C.c_double)
    rc = lib.qgt_b200_natural_gradient(None, G.ctypes.data_as(dp), g.ctypes.data_as(dp), 20, None, out.ctypes.data_as(dp), C.byref(lam))
    assert rc == 0
    x, lam_o = oracle.natural_gradient(G, g)
    assert lam.value == lam_o
    assert np.abs(out - x).max() / np.abs(x).max() < 1e-9
    # ill-conditioned metric: adaptive regularisation kicks in identically
    q, _ = np.linalg.qr(A)
    G2 = q @ np.diag(np.logspace(0, -9, 20)) @ q.T
    rc = lib.qgt_b200_natural_gradient(None, G2.ctypes.data_as(dp), g.ctypes.data_as(dp), 20, None, out.ctypes.data_as(dp), C.byref(lam))
    assert rc == 0
    x2, lam2 = oracle.natural_gradient(G2, g)
    assert lam2 > 1e-4 and abs(lam.value - lam2) / lam2 < 1e-4   # kappa of a 1e9-conditioned matrix is itself ill-determined
    assert np.abs(out - x2).max() / np.abs(x2).max() < 1e-3


def test_natural_gradient_rank_deficient_metric_keeps_a_usable_lambda():
    """a rank-deficient metric must not push the adaptive lambda to 100 (kappa = 1e16)."""
    lib = api.load()
    rng = np.random.default_rng(3)
    A = rng.normal(size=(24, 12))
    G = A @ A.T / 24 * 0.25                      # rank 12 of 24, eigenvalues <= ~0.25 like a Fubini-Study metric
    g = G @ rng.normal(size=24)                  # a gradient inside the range
    out = np.zeros(24)
    lam = C.c_double(0)
    dp = C.POINTER(C.c_double)
    rc = lib.qgt_b200_natural_gradient(None, G.ctypes.data_as(dp), g.ctypes.data_as(dp), 24, None, out.ctypes.data_as(dp), C.byref(lam))
    assert rc == 0
    w = np.linalg.eigvalsh(G)
    kappa_range = w.max() / w[w > 1e-10].min()
    assert lam.value == max(1e-4, 1e-6 * np.sqrt(kappa_range) if kappa_range > 1e8 else 1e-4)
    assert lam.value < 1e-2
    x = np.linalg.solve(G + lam.value * np.eye(24), g)
    assert np.abs(out - x).max() / np.abs(x).max() < 1e-8


def test_error_strings():
    lib = api.load()
    assert lib.qgt_b200_error_string(0) == b"success"
    assert b"CPU fallback" in lib.qgt_b200_error_string(-31)
