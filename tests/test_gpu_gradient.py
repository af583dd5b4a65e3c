"""Adjoint energy gradient (qgt_b200_expectation_gradient): one forward circuit + one backward pass of two states.
Checked against the CPU oracle's per-parameter derivative columns (oracle/qgt_oracle.c orc_expectation_gradient, which
follows the reference's caller semantics, algorithms/qaoa.c:489-558: dE/dtheta_mu = 2 Re <d_mu psi|H|psi>), to 1e-10."""
import numpy as np
import pytest

from quantum_geometric_tensor_b200 import api, circuits as K

pytestmark = pytest.mark.gpu


def ring_observable(c, seed=3):
    rng = np.random.default_rng(seed)
    n = c.num_qubits
    c.edges = [(q, (q + 1) % n, float(rng.uniform(0.5, 1.5))) for q in range(n if n > 2 else 1)]
    c.vertex_weights = [float(v) for v in rng.uniform(-1, 1, n)]
    return c


def fusable(c, th):
    try:
        api.plan_dump_gradient(c, th, True)
        return True
    except api.QgtError:
        return False


def check(ctx, oracle, c, th, want_fused, tol=1e-10):
    assert fusable(c, th) == want_fused
    e, g = ctx.expectation_gradient(c, th)
    st = ctx.stats()
    eo, go = oracle.expectation_gradient(c, th)
    scale = max(1.0, np.abs(go).max())
    assert abs(e - eo) < tol * max(1.0, abs(eo))
    assert np.abs(g - go).max() < tol * scale
    assert st["fused"] == (1 if want_fused else 0)
    return st


@pytest.mark.parametrize("n,layers", [(9, 2), (10, 2), (12, 3), (14, 2), (16, 2)])
def test_ansatz_gradient_fused_path(ctx, oracle, n, layers):
    c = ring_observable(K.hea_layers(n, layers))
    check(ctx, oracle, c, K.default_angles(c.num_params, 11 + n), want_fused=True)


@pytest.mark.parametrize("n,layers", [(2, 2), (5, 3), (7, 1), (8, 2), (10, 3)])
def test_small_register_gradient_generic_path(ctx, oracle, n, layers):
    c = ring_observable(K.hea_layers(n, layers))
    check(ctx, oracle, c, K.default_angles(c.num_params, 5 + n), want_fused=False)


@pytest.mark.parametrize("n,p", [(8, 2), (10, 3), (12, 1)])
def test_qaoa_gradient_cost_layers(ctx, oracle, n, p):
    c = K.qaoa_maxcut(n, p)
    check(ctx, oracle, c, K.default_angles(c.num_params, 6), want_fused=False)


@pytest.mark.parametrize("n,seed", [(9, 1), (10, 2), (6, 3), (12, 4)])
def test_random_parametric_circuits(ctx, oracle, n, seed):
    """Every parametric kind (controlled rotations, U1/PHASE, ZZ), shared parameters with scales and offsets, fixed gates
    whose inverse is another kind (S, T, SX)."""
    rng = np.random.default_rng(seed)
    c = K.Circuit(n)
    P = 7
    kinds1 = [K.RX, K.RY, K.RZ, K.U1, K.PHASE]
    kinds2 = [K.CRX, K.CRY, K.CRZ, K.ZZ]
    fixed = [K.H, K.S, K.T, K.SX, K.SDG, K.TDG, K.X, K.Y, K.Z]
    for _ in range(60):
        r = rng.uniform()
        t = int(rng.integers(0, n))
        o = int(rng.integers(0, n - 1)); o += o >= t
        if r < 0.35:
            c.add(int(rng.choice(kinds1)), t, -1, int(rng.integers(0, P)), float(rng.uniform(-1, 1)), float(rng.uniform(-2, 2)))
        elif r < 0.55:
            c.add(int(rng.choice(kinds2)), t, o, int(rng.integers(0, P)), float(rng.uniform(-1, 1)), float(rng.uniform(-2, 2)))
        elif r < 0.8:
            c.add(int(rng.choice(fixed)), t)
        else:
            c.add(int(rng.choice([K.CNOT, K.CZ, K.SWAP, K.CH, K.CY])), t, o)
    c.num_params = P
    ring_observable(c, seed)
    e, g = ctx.expectation_gradient(c, K.default_angles(P, seed))
    eo, go = oracle.expectation_gradient(c, K.default_angles(P, seed))
    assert abs(e - eo) < 1e-10 and np.abs(g - go).max() < 1e-10 * max(1.0, np.abs(go).max())


def test_energy_only_and_parameter_free(ctx, oracle):
    c = ring_observable(K.Circuit(6))
    c.add(K.H, 0); c.add(K.CNOT, 1, 0); c.add(K.SX, 3)
    e, g = ctx.expectation_gradient(c, np.zeros(1))
    eo, _ = oracle.expectation_gradient(c, np.zeros(1))
    assert abs(e - eo) < 1e-12 and g.size == 0


def test_config2_gradient_by_parameter_shift_and_cost(ctx, oracle):
    """BASELINE config 2 (20 qubits, 160 parameters): every component in one call; three of them checked against the
    oracle through the exact parameter-shift rule dE = [E(theta + pi/2) - E(theta - pi/2)] / 2 (two oracle circuits each);
    device time a small multiple of one forward circuit (the per-parameter re-simulation it replaces costs > 160 of them)."""
    c = ring_observable(K.config("c2"))
    th = K.default_angles(c.num_params)
    e, g = ctx.expectation_gradient(c, th)
    st = ctx.stats()
    assert st["fused"] == 1 and g.shape == (160,)
    for mu in (0, 77, 159):
        ep = []
        for s in (+np.pi / 2, -np.pi / 2):
            t2 = th.copy(); t2[mu] += s
            psi = oracle.apply(c, t2)
            ep.append(oracle.energy(c, psi))
        assert abs(g[mu] - 0.5 * (ep[0] - ep[1])) < 1e-10
    psi = oracle.apply(c, th)
    assert abs(e - oracle.energy(c, psi)) < 1e-10
    # cost: compare with the device time of a forward circuit
    s = ctx.state(c.num_qubits).init(0)
    s.apply(c, th)
    s.apply(c, th)
    fwd = ctx.stats()["ms_total"]
    s.close()
    ctx.expectation_gradient(c, th)
    ms = ctx.stats()["ms_total"]
    print(f"adjoint gradient {ms:.3f} ms, forward circuit {fwd:.3f} ms, ratio {ms / fwd:.2f}")
    assert ms < 12 * fwd


def test_natural_gradient_descent_lowers_the_energy(ctx, oracle):
    """The optimiser loop (metric + adjoint gradient + regularised solve per step) on a 10-qubit ansatz: the first step
    equals the one assembled from the oracle's gradient and this library's metric, and the energy goes down."""
    c = ring_observable(K.hea_layers(10, 2))
    th0 = K.default_angles(c.num_params, 21)
    th, hist = ctx.natural_gradient_descent(c, th0, 6, 0.05)
    assert hist.shape == (7,) and np.all(np.diff(hist) < 1e-9) and hist[-1] < hist[0] - 1e-3
    eo, go = oracle.expectation_gradient(c, th0)
    q = ctx.qgt(c, th0)
    x, _ = oracle.natural_gradient(q.real, go)
    th1, h1 = ctx.natural_gradient_descent(c, th0, 1, 0.05)
    assert abs(h1[0] - eo) < 1e-10 and np.abs(th1 - (th0 - 0.05 * x)).max() < 1e-7
    psi = oracle.apply(c, th)
    assert abs(hist[-1] - oracle.energy(c, psi)) < 1e-9
