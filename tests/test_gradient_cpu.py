"""The adjoint-gradient programs (fused and generic) interpreted on the CPU from the planner's JSON dump and compared
with the oracle's per-parameter derivative columns: pins the inverse-circuit construction, the sign and the program
structure without a GPU."""
import numpy as np
import pytest

import plan_interp
from quantum_geometric_tensor_b200 import api, circuits as K


def ring(c, seed=3):
    rng = np.random.default_rng(seed)
    n = c.num_qubits
    c.edges = [(q, (q + 1) % n, float(rng.uniform(0.5, 1.5))) for q in range(n if n > 2 else 1)]
    c.vertex_weights = [float(v) for v in rng.uniform(-1, 1, n)]
    return c


def mixed_circuit(n, seed, P=6):
    rng = np.random.default_rng(seed)
    c = K.Circuit(n)
    for _ in range(40):
        r = rng.uniform()
        t = int(rng.integers(0, n))
        o = int(rng.integers(0, n - 1)); o += o >= t
        if r < 0.4:
            c.add(int(rng.choice([K.RX, K.RY, K.RZ, K.U1, K.PHASE])), t, -1, int(rng.integers(0, P)), float(rng.uniform(-1, 1)), float(rng.uniform(-2, 2)))
        elif r < 0.6:
            c.add(int(rng.choice([K.CRX, K.CRY, K.CRZ, K.ZZ])), t, o, int(rng.integers(0, P)), float(rng.uniform(-1, 1)), float(rng.uniform(-2, 2)))
        elif r < 0.8:
            c.add(int(rng.choice([K.H, K.S, K.T, K.SX, K.SDG, K.TDG, K.X, K.Y, K.Z])), t)
        else:
            c.add(int(rng.choice([K.CNOT, K.CZ, K.SWAP, K.CH, K.CY])), t, o)
    c.num_params = P
    return ring(c, seed)


CASES = [("hea10", lambda: ring(K.hea_layers(10, 2)), True), ("hea9", lambda: ring(K.hea_layers(9, 1)), True),
         ("hea5", lambda: ring(K.hea_layers(5, 2)), False), ("qaoa6", lambda: K.qaoa_maxcut(6, 2), False),
         ("mixed10", lambda: mixed_circuit(10, 1), True), ("mixed6", lambda: mixed_circuit(6, 2), False),
         ("hea8_generic", lambda: ring(K.hea_layers(8, 1)), False), ("hea10_generic", lambda: ring(K.hea_layers(10, 1)), False)]


@pytest.mark.parametrize("name,make,fused", CASES, ids=[c[0] for c in CASES])
def test_gradient_program_matches_oracle(oracle, name, make, fused):
    c = make()
    th = K.default_angles(c.num_params, 9)
    try:
        plan = api.plan_dump_gradient(c, th, fused, scratch_slots=2)
    except api.QgtError as e:
        if fused and e.status == -7:
            pytest.skip("this circuit's inverse plan does not qualify for the fused schedule")
        raise
    assert plan["program"]["fused"] == (1 if fused else 0)
    psi = oracle.apply(c, th)
    g = plan_interp.run_gradient_program(plan, c, psi)
    eo, go = oracle.expectation_gradient(c, th)
    assert np.abs(g - go).max() < 1e-11 * max(1.0, np.abs(go).max())


def test_inverse_circuit_restores_the_initial_state(oracle):
    """The inverse plan applied to psi gives |init> back (every gate kind, including the ones whose inverse is another kind)."""
    c = mixed_circuit(7, 5)
    th = K.default_angles(c.num_params, 2)
    plan = api.plan_dump_gradient(c, th, False, scratch_slots=1)
    st = oracle.apply(c, th)
    for run in plan["runs"]:
        for op in run["ops"]:
            st = plan_interp.apply_op(st, op, c)
    assert abs(st[0] - 1.0) < 1e-12 and np.abs(st[1:]).max() < 1e-12


SHARDED = [("hea11_w2_fused", lambda: ring(K.hea_layers(11, 2)), True, 2), ("hea12_w4_fused", lambda: ring(K.hea_layers(12, 1)), True, 4),
           ("qaoa8_w2", lambda: K.qaoa_maxcut(8, 2), False, 2), ("mixed9_w4", lambda: mixed_circuit(9, 4), False, 4),
           ("hea10_w2_generic", lambda: ring(K.hea_layers(10, 2)), False, 2)]


@pytest.mark.parametrize("name,make,fused,world", SHARDED, ids=[c[0] for c in SHARDED])
def test_sharded_gradient_program_matches_oracle(oracle, name, make, fused, world):
    """The inverse circuit mapped onto `world` ranks (exchanges included) with the gradient program of adjoint.cu, all ranks
    interpreted in one process, against the oracle's gradient of the unsharded circuit."""
    c = make()
    th = K.default_angles(c.num_params, 13)
    try:
        plan = api.plan_dump_gradient(c, th, fused, scratch_slots=1, world=world)
    except api.QgtError as e:
        if fused and e.status == -7:
            pytest.skip("this circuit's sharded inverse plan does not qualify for the fused schedule")
        raise
    assert plan["nloc"] == c.num_qubits - {2: 1, 4: 2}[world]
    psi = oracle.apply(c, th)
    g = plan_interp.run_gradient_program_sharded(plan, c, psi, world)
    eo, go = oracle.expectation_gradient(c, th)
    assert np.abs(g - go).max() < 1e-11 * max(1.0, np.abs(go).max())
