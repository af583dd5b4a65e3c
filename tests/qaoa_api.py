"""ctypes view of the reference's QAOA driver API (algorithms/qaoa.h), usable on the compat library (the product) and on
oracle/_ref/libqaoa_ref.so (the unmodified reference, test infrastructure)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "quantum_geometric_tensor_b200", "libqgt_b200_compat.so")
REF = os.path.join(ROOT, "oracle", "_ref", "libqaoa_ref.so")


class Edge(C.Structure):
    _fields_ = [("i", C.c_size_t), ("j", C.c_size_t), ("weight", C.c_double)]


class Graph(C.Structure):
    _fields_ = [("num_vertices", C.c_size_t), ("num_edges", C.c_size_t), ("edges", C.POINTER(Edge)), ("vertex_weights", C.POINTER(C.c_double))]


class Config(C.Structure):
    _fields_ = [("p", C.c_size_t), ("problem_type", C.c_int), ("mixer_type", C.c_int), ("optimizer_type", C.c_int),
                ("initial_gamma", C.POINTER(C.c_double)), ("initial_beta", C.POINTER(C.c_double)), ("max_iterations", C.c_size_t),
                ("tolerance", C.c_double), ("learning_rate", C.c_double), ("num_shots", C.c_size_t), ("use_expectation", C.c_bool),
                ("use_gpu", C.c_bool), ("backend", C.c_void_p)]


class QuantumState(C.Structure):
    _fields_ = [("num_qubits", C.c_size_t), ("amplitudes", C.POINTER(C.c_float)), ("workspace", C.c_void_p),
                ("dimension", C.c_size_t), ("is_normalized", C.c_bool)]


class State(C.Structure):
    _fields_ = [("graph", C.POINTER(Graph)), ("cost_hamiltonian", C.c_void_p), ("mixer_hamiltonian", C.c_void_p),
                ("gamma", C.POINTER(C.c_double)), ("beta", C.POINTER(C.c_double)), ("p", C.c_size_t), ("num_qubits", C.c_size_t),
                ("qstate", C.POINTER(QuantumState)), ("current_cost", C.c_double), ("best_cost", C.c_double),
                ("best_gamma", C.POINTER(C.c_double)), ("best_beta", C.POINTER(C.c_double)), ("iteration", C.c_size_t),
                ("config", Config), ("best_solution", C.POINTER(C.c_int)), ("solution_probabilities", C.POINTER(C.c_double))]


class Result(C.Structure):
    _fields_ = [("optimal_cost", C.c_double), ("optimal_solution", C.POINTER(C.c_int)), ("optimal_gamma", C.POINTER(C.c_double)),
                ("optimal_beta", C.POINTER(C.c_double)), ("num_iterations", C.c_size_t), ("cost_history", C.POINTER(C.c_double)),
                ("history_length", C.c_size_t), ("execution_time", C.c_double), ("approximation_ratio", C.c_double)]


_DP = C.POINTER(C.c_double)


def dp(a):
    return a.ctypes.data_as(_DP)


class Qaoa:
    def __init__(self, path):
        self.lib = L = C.CDLL(path)
        L.qaoa_create_graph.restype = C.POINTER(Graph); L.qaoa_create_graph.argtypes = [C.c_size_t]
        L.qaoa_add_edge.argtypes = [C.POINTER(Graph), C.c_size_t, C.c_size_t, C.c_double]
        L.qaoa_set_vertex_weight.argtypes = [C.POINTER(Graph), C.c_size_t, C.c_double]
        L.qaoa_destroy_graph.argtypes = [C.POINTER(Graph)]
        L.qaoa_default_config.restype = Config; L.qaoa_default_config.argtypes = [C.c_size_t]
        L.qaoa_init.restype = C.POINTER(State); L.qaoa_init.argtypes = [C.POINTER(Graph), C.POINTER(Config)]
        L.qaoa_apply_circuit.argtypes = [C.POINTER(State), _DP, _DP]
        L.qaoa_apply_layer.argtypes = [C.POINTER(State), C.c_size_t]
        L.qaoa_prepare_initial_state.argtypes = [C.POINTER(State)]
        L.qaoa_compute_expectation.argtypes = [C.POINTER(State), _DP]
        L.qaoa_compute_gradient.argtypes = [C.POINTER(State), _DP, _DP]
        L.qaoa_optimize.restype = C.POINTER(Result); L.qaoa_optimize.argtypes = [C.POINTER(State)]
        L.qaoa_sample.argtypes = [C.POINTER(State), C.POINTER(C.POINTER(C.c_int)), C.c_size_t]
        L.qaoa_evaluate_solution.restype = C.c_double; L.qaoa_evaluate_solution.argtypes = [C.POINTER(Graph), C.POINTER(C.c_int)]
        L.qaoa_destroy.argtypes = [C.POINTER(State)]
        L.qaoa_destroy_result.argtypes = [C.POINTER(Result)]
        L.qaoa_estimate_optimal_p.restype = C.c_size_t; L.qaoa_estimate_optimal_p.argtypes = [C.c_size_t, C.c_size_t]

    def graph(self, n, edges, vertex_weights=None):
        g = self.lib.qaoa_create_graph(n)
        for i, j, w in edges:
            assert self.lib.qaoa_add_edge(g, i, j, w) == 0
        if vertex_weights is not None:
            for q, v in enumerate(vertex_weights):
                assert self.lib.qaoa_set_vertex_weight(g, q, float(v)) == 0
        return g

    def init(self, g, p, gamma, beta, **cfg):
        c = self.lib.qaoa_default_config(p)
        ga, be = np.ascontiguousarray(gamma, dtype=np.float64), np.ascontiguousarray(beta, dtype=np.float64)
        c.initial_gamma, c.initial_beta = dp(ga), dp(be)
        for k, v in cfg.items():
            setattr(c, k, v)
        s = self.lib.qaoa_init(g, C.byref(c))
        return s

    def amplitudes(self, s):
        n = s.contents.num_qubits
        return np.ctypeslib.as_array(s.contents.qstate.contents.amplitudes, shape=(2 << n,)).copy().view(np.complex64)
