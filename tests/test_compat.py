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
                    "-o", str(exe), "-L" + PKG, "-lqgt_b200_compat", "-lqgt_b200", "-Wl,-rpath," + PKG, "-lm", "-lpthread"], check=True)
    return str(exe)


def _build_seams(tmp_path):
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "libcompute_ref.so")):
        pytest.skip("oracle/_ref/libcompute_ref.so (the reference's registry + CPU backend) is not built")
    exe = tmp_path / "test_seams"
    subprocess.run(["gcc", "-std=gnu11", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "compat", "test_seams.c"),
                    "-o", str(exe), "-L" + PKG, "-lqgt_b200_compat", "-lqgt_b200", "-L" + ref, "-lcompute_ref",
                    "-Wl,-rpath," + PKG, "-Wl,-rpath," + ref, "-lm"], check=True)
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
                "qg_gpu_destroy_stream", "qg_gpu_synchronize_stream", "qg_gpu_synchronize",
                # round 2: the back-end seams, gate objects, parameter shifts, measurement
                "qgt_b200_compute_backend_ops", "qgt_b200_compute_backend_info", "qgt_b200_register_compute_backend",
                "qgt_b200_compute_backend_ctx", "compute_quantum_metric_gpu", "compute_quantum_connection_gpu",
                "compute_quantum_curvature_gpu", "qgt_b200_context_init", "qgt_default_config", "qgt_error_string",
                "create_quantum_gate", "copy_quantum_gate", "update_gate_parameters", "shift_gate_parameters", "destroy_quantum_gate",
                "create_rx_gate", "create_cnot_gate", "initialize_numerical_backend", "shutdown_numerical_backend",
                "shift_parameter", "compute_shifted_states", "compute_parameter_shift_gradient",
                "compute_centered_difference_gradient", "compute_higher_order_gradient", "compute_gradient_with_error",
                "sim_measure_qubit", "sim_measure_all", "sim_get_measurement_counts", "sim_get_expectation_value",
                "sim_save_circuit", "sim_load_circuit",
                "quantum_circuit_create", "quantum_circuit_destroy", "quantum_circuit_hadamard", "quantum_circuit_rotation",
                "quantum_circuit_cnot", "quantum_circuit_execute", "quantum_circuit_measure_all", "quantum_circuit_depth",
                "init_quantum_state", "quantum_state_cleanup",
                "qaoa_create_graph", "qaoa_add_edge", "qaoa_init", "qaoa_apply_circuit", "qaoa_apply_layer", "qaoa_compute_expectation",
                "qaoa_compute_gradient", "qaoa_optimize", "qaoa_sample", "qaoa_evaluate_solution", "qaoa_destroy",
                "qgt_b200_qaoa_exact_gradient",
                "geometric_tensor_create", "geometric_tensor_destroy", "geometric_tensor_add", "geometric_tensor_multiply",
                "geometric_tensor_contract", "geometric_tensor_transpose", "geometric_tensor_conjugate", "geometric_tensor_norm",
                "geometric_tensor_validate", "geometric_tensor_initialize_random", "geometric_tensor_is_hermitian"):
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


def test_seams_host_only_parts(tmp_path):
    """Registration with the reference's registry, probe() without a device, gate objects: no GPU needed."""
    from quantum_geometric_tensor_b200 import api
    if api.device_count() > 0:
        pytest.skip("a GPU is present: the full program runs in the gpu test")
    r = subprocess.run([_build_seams(tmp_path), "--host-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all host-only seam checks passed" in r.stdout


@pytest.mark.gpu
def test_seams_on_gpu(tmp_path):
    """ComputeBackendOps against the reference's CPU backend, the GPUContext seam, the network API, parameter shifts."""
    r = subprocess.run([_build_seams(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all seam checks passed" in r.stdout


REF_TESTS = "/root/reference/tests"


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", ["test_quantum_simulator_cpu", "test_quantum_geometric_minimal", "test_quantum_geometric_tensor_gpu"])
def test_reference_tests_compile_and_link_unmodified(tmp_path, name):
    """The reference's own test programs for the path build against include/ + the two libraries without edits."""
    exe = tmp_path / name
    subprocess.run(["gcc", "-std=gnu11", "-w", "-I" + os.path.join(ROOT, "include"), os.path.join(REF_TESTS, name + ".c"), "-o", str(exe),
                    "-L" + PKG, "-lqgt_b200_compat", "-lqgt_b200", "-Wl,-rpath," + PKG, "-lm"], check=True)
    assert exe.exists()


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", ["test_quantum_geometric_tensor", "test_quantum_geometric_tensor_init"])
def test_reference_tensor_tests_pass_unmodified(tmp_path, name):
    """The reference's own tests of the generic tensor objects (host algebra, no GPU involved) build against include/ and
    pass against the compat library."""
    exe = tmp_path / name
    subprocess.run(["gcc", "-std=gnu11", "-w", "-I" + os.path.join(ROOT, "include"), os.path.join(REF_TESTS, name + ".c"), "-o", str(exe),
                    "-L" + PKG, "-lqgt_b200_compat", "-lqgt_b200", "-Wl,-rpath," + PKG, "-lm"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "All tests passed!" in r.stdout


def _declared(header):
    import re
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"#define[^\n]*(\\\n[^\n]*)*", "", src)
    names = set()
    for stmt in src.split(";"):
        s = stmt.strip()
        if "{" in s or "}" in s or s.startswith("typedef"):
            continue
        m = re.match(r"^(?:[\w\s\*]+?)\b([a-z_][a-z0-9_]*)\s*\([^()]*(\([^()]*\)[^()]*)*\)\s*$", s, flags=re.S)
        if m and "(*" not in s.split("(")[0]:
            names.add(m.group(1))
    return names


def test_every_declared_compat_function_is_exported():
    """Every function include/qgt_compat.h and include/qgt_compute_backend.h declare is defined by libqgt_b200_compat.so; the
    only exceptions are the registry / engine functions, which the reference's supercomputer/compute_backend.c provides."""
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(PKG, "libqgt_b200_compat.so")], check=True,
                         capture_output=True, text=True).stdout
    defined = {l.split()[-1] for l in out.splitlines() if " T " in l}
    compat = _declared("qgt_compat.h")
    assert len(compat) >= 160
    assert sorted(n for n in compat if n not in defined) == []
    backend = _declared("qgt_compute_backend.h")
    registry = {"compute_register_backend", "compute_get_backend_count", "compute_get_backend_info", "compute_get_backend_by_type",
                "compute_select_backend", "compute_backend_available", "compute_engine_init", "compute_engine_cleanup",
                "compute_engine_get_backend_type", "compute_engine_get_ops", "compute_engine_get_backend"}
    assert sorted(n for n in backend if n not in defined and n not in registry) == []
    assert {"qgt_b200_compute_backend_ops", "qgt_b200_register_compute_backend"} <= backend & defined
