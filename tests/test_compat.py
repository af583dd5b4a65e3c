"""The reference-named C entry points (include/quantum_geometric/...): compiled from a C test program in the
style of the reference's own tests and run against libqgt_b200_compat.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "quantum_geometric_tensor_b200")


def _build(tmp_path):
    exe = tmp_path / "test_compat"
    subprocess.run(["gcc", "-std=gnu11", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "compat", "test_compat.c"),
                    "-o", str(exe), "-L" + PKG, "-lqgt_b200_compat", "-lqgt_b200", "-Wl,-rpath," + PKG, "-lm"], check=True)
    return str(exe)


def test_compat_library_exports_reference_symbols():
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(PKG, "libqgt_b200_compat.so")], check=True,
                         capture_output=True, text=True).stdout
    for sym in ("init_simulator_state", "cpu_sim_create_circuit", "cpu_sim_add_gate", "simulate_circuit_cpu", "cpu_sim_cleanup_circuit",
                "cpu_sim_get_error_statistics", "configure_circuit_optimization", "sim_init", "sim_add_gate", "sim_execute_circuit",
                "sim_get_statevector", "diffgeo_compute_fubini_study", "diffgeo_compute_berry_curvature",
                "create_quantum_geometric_tensor_network", "apply_quantum_gate", "compute_quantum_geometric_tensor", "compute_quantum_metric",
                "compute_berry_curvature", "geometric_compute_fubini_study_metric", "geometric_compute_berry_curvature",
                "geometric_compose_qgt", "geometric_compute_full_qgt", "compute_regularized_natural_gradient",
                "get_default_natural_gradient_config", "gpu_malloc", "qgt_gpu_free_buffer", "gpu_memcpy_host_to_device",
                "gpu_memcpy_device_to_host", "qg_gpu_init", "qg_gpu_cleanup", "qg_gpu_shutdown", "qg_gpu_get_device_count",
                "qg_gpu_get_device_info", "qg_gpu_set_device", "qg_gpu_get_last_error", "qg_gpu_get_error_string", "qg_gpu_allocate",
                "qg_gpu_allocate_pinned", "qg_gpu_free", "qg_gpu_memcpy_to_device", "qg_gpu_memcpy_to_host", "qg_gpu_create_stream",
                "qg_gpu_destroy_stream", "qg_gpu_synchronize_stream", "qg_gpu_synchronize"):
        assert f" T {sym}\n" in out, sym


def test_compat_host_only_parts(tmp_path):
    from quantum_geometric_tensor_b200 import api
    if api.device_count() > 0:
        pytest.skip("a GPU is present: the full program runs in the gpu test")
    r = subprocess.run([_build(tmp_path), "--host-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all host-only checks passed" in r.stdout


@pytest.mark.gpu
def test_compat_entry_points_on_gpu(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all compat checks passed" in r.stdout
