"""GPU parity tests: the CUDA path (through the C-ABI of include/qgt_b200.h) against the CPU oracle,
the committed golden vectors of the reference, and size-independent properties at full sizes.

Tolerances (north star): 1e-10 relative for complex-double results."""
import ctypes as C

import numpy as np
import pytest

from quantum_geometric_tensor_b200 import api, circuits as K

from helpers import golden_cases, known_answer_cases, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-10
ALL_KINDS = [K.X, K.Y, K.Z, K.H, K.S, K.T, K.SDG, K.TDG, K.SX, K.RX, K.RY, K.RZ, K.PHASE, K.CNOT, K.CY, K.CZ,
             K.CH, K.SWAP, K.CRX, K.CRY, K.CRZ, K.ZZ]


def _state(ctx, circ, theta):
    st = ctx.state(circ.num_qubits).init(circ.initial_state)
    st.apply(circ, theta)
    out = st.download()
    st.close()
    return out


# ---- gate sweeps -------------------------------------------------------------------------------------
@pytest.mark.parametrize("label,circ,expected", known_answer_cases(), ids=[c[0] for c in known_answer_cases()])
def test_reference_known_answers(ctx, label, circ, expected):
    # the magnitudes tests/test_quantum_simulator_cpu.c asserts, through the host-buffer entry point
    n = circ.num_qubits
    amps = np.zeros(1 << n, dtype=np.complex128)
    amps[0] = 1
    out = ctx.simulate_host(amps, circ, np.zeros(1))
    assert np.abs(np.abs(out) - np.array(expected)).max() < 1e-10


@pytest.mark.parametrize("name", golden_cases())
def test_golden_vectors(ctx, name):
    c, z = load_golden(name)
    th = z["theta"]
    psi = _state(ctx, c, th)
    assert np.abs(psi[: z["psi"].size] - z["psi"]).max() < 1e-13
    assert abs(np.sum(psi * np.arange(1, psi.size + 1)) - z["psi_checksum"][0]) < 1e-9
    q = ctx.qgt(c, th)
    assert rel_err(q.real, z["metric"]) < TOL
    assert rel_err(-2 * q.imag, z["curvature"]) < TOL


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 16])
def test_random_circuit_state_parity(ctx, oracle, n):
    kinds = ALL_KINDS if n > 1 else [K.X, K.Y, K.Z, K.H, K.S, K.T, K.SX, K.RX, K.RY, K.RZ, K.PHASE]
    c = K.random_circuit(n, 60, 1000 + n, kinds=kinds, share_params=True)
    th = K.default_angles(max(1, c.num_params), n)
    psi = _state(ctx, c, th)
    ref = oracle.apply(c, th)
    assert np.abs(psi - ref).max() < 1e-13


@pytest.mark.parametrize("n,tile", [(12, 8), (12, 9), (13, 10), (14, 11), (15, 11)])
def test_tile_size_option(ctx, oracle, n, tile):
    c = K.hea_layers(n, 2)
    th = K.default_angles(c.num_params)
    ctx.set_option("tile_qubits", tile)
    try:
        psi = _state(ctx, c, th)
    finally:
        ctx.set_option("tile_qubits", 11)
    assert np.abs(psi - oracle.apply(c, th)).max() < 1e-13


@pytest.mark.parametrize("opts", [{"use_mma": 0}, {"use_mma": 0, "batch_qubits": 1}, {"use_mma": 1, "reg_qubits": 2},
                                  {"use_mma": 1, "tile_qubits": 10, "max_ops_per_run": 12}])
def test_kernel_path_options(ctx, oracle, opts):
    # the register (DFMA) path, the batched variant and other tilings must agree with the tensor-pipe default
    defaults = {"use_mma": 1, "batch_qubits": 0, "reg_qubits": 3, "tile_qubits": 11, "max_ops_per_run": 160}
    c = K.random_circuit(12, 80, 4242, kinds=ALL_KINDS, share_params=True)
    th = K.default_angles(max(1, c.num_params), 1)
    c1 = K.config("c1")
    th1 = K.default_angles(c1.num_params)
    for k, v in opts.items():
        ctx.set_option(k, v)
    try:
        psi = _state(ctx, c, th)
        q = ctx.qgt(c, th)
        q1 = ctx.qgt(c1, th1)
    finally:
        for k, v in defaults.items():
            ctx.set_option(k, v)
    assert np.abs(psi - oracle.apply(c, th)).max() < 1e-13
    assert rel_err(q, oracle.qgt(c, th)) < TOL
    assert rel_err(q1, oracle.qgt(c1, th1)) < TOL


def test_every_target_every_kind(ctx, oracle):
    # each gate kind on each target (and a control on every other position) of a 13-qubit register
    n = 13
    rng = np.random.default_rng(5)
    base = K.hea_layers(n, 1)
    th = K.default_angles(base.num_params)
    for kind in (K.H, K.RX, K.RY, K.RZ, K.Y, K.T, K.CNOT, K.CZ, K.CRY, K.ZZ, K.SWAP):
        c = K.Circuit(n)
        c.gates = list(base.gates)
        c.num_params = base.num_params
        for t in range(n):
            ctl = int((t + 1 + rng.integers(n - 1)) % n) if kind in K.TWO_QUBIT else -1
            c.add(kind, t, ctl, -1, float(rng.uniform(-3, 3)))
        assert np.abs(_state(ctx, c, th) - oracle.apply(c, th)).max() < 1e-13, kind


def test_norm_and_upload_download_roundtrip(ctx):
    n = 15
    rng = np.random.default_rng(0)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    st = ctx.state(n).upload(v)
    assert np.array_equal(st.download(), v)
    assert abs(st.norm2() - np.vdot(v, v).real) / np.vdot(v, v).real < 1e-13
    st.init(K.INIT_PLUS)
    assert abs(st.norm2() - 1.0) < 1e-13
    st.close()


# ---- QGT -----------------------------------------------------------------------------------------------
def test_qgt_config1_matches_oracle(ctx, oracle):
    c = K.config("c1")                      # 12 qubits, 48 parameters: BASELINE config 1
    th = K.default_angles(c.num_params)
    psi = ctx.state(c.num_qubits)
    q = ctx.qgt(c, th, psi_out=psi)
    qo = oracle.qgt(c, th)
    assert rel_err(q, qo) < TOL
    assert np.abs(psi.download() - oracle.apply(c, th)).max() < 1e-13
    g, b = ctx.qgt_metric_berry(c, th)
    assert rel_err(g, qo.real) < TOL and rel_err(b, qo.imag) < TOL
    psi.close()


@pytest.mark.parametrize("seed", range(5))
def test_qgt_random_circuits_shared_params(ctx, oracle, seed):
    n = 4 + 2 * seed
    c = K.random_circuit(n, 50, 77 + seed, kinds=ALL_KINDS, share_params=True)
    th = K.default_angles(max(1, c.num_params), seed)
    q = ctx.qgt(c, th)
    assert rel_err(q, oracle.qgt(c, th)) < TOL


def test_qgt_qaoa_matches_oracle(ctx, oracle):
    c = K.qaoa_maxcut(10, 3)
    c.vertex_weights = list(np.linspace(-0.5, 0.5, 10))
    th = K.default_angles(c.num_params, 4)
    assert rel_err(ctx.qgt(c, th), oracle.qgt(c, th)) < TOL


@pytest.mark.parametrize("n,p,weights", [(12, 2, True), (14, 1, False), (13 + 1, 2, True)])
def test_qgt_qaoa_multi_tile(ctx, oracle, n, p, weights):
    # more qubits than a tile holds: cross-tile edges, per-tile energy tables, cost passes on the tensor-only kernel
    c = K.qaoa_maxcut(n, p)
    if weights:
        c.vertex_weights = list(np.linspace(-0.3, 0.7, n))
        c.edges = [(i, j, 0.5 + 0.25 * ((i + j) % 3)) for (i, j, *_) in c.edges]
    th = K.default_angles(c.num_params, n)
    assert rel_err(ctx.qgt(c, th), oracle.qgt(c, th)) < TOL
    st = ctx.state(n).init(1).apply(c, th)
    assert np.abs(st.download() - oracle.apply(c, th, oracle.init_state(n, 1))).max() < 1e-13
    st.close()


def test_stage_forms_both_paths(ctx, oracle):
    # one-layer stages factor as (diagonal) x (real) and take the 4-DMMA path, two fused layers are dense:
    # both must agree with the oracle, with and without the tensor pipe
    for layers in (1, 2, 3):
        c = K.hea_layers(12, layers)
        th = K.default_angles(c.num_params, layers)
        ref = oracle.qgt(c, th)
        for use_mma in (1, 0):
            ctx.set_option("use_mma", use_mma)
            try:
                assert rel_err(ctx.qgt(c, th), ref) < TOL
            finally:
                ctx.set_option("use_mma", 1)


@pytest.mark.parametrize("slots", [0, 9])
def test_runs_split_where_parameters_are_born(ctx, oracle, slots):
    # birth_cut=2 forces on a small state what 20-qubit plans do by themselves: the first run executed in pieces
    c = K.hea_layers(12, 3)
    th = K.default_angles(c.num_params, 3)
    ctx.set_option("birth_cut", 0)
    ctx.qgt(c, th)
    runs_whole = ctx.stats()["num_runs"]
    ctx.set_option("birth_cut", 2)
    ctx.set_option("max_slots", slots)
    try:
        q = ctx.qgt(c, th)
        assert ctx.stats()["num_runs"] > runs_whole
    finally:
        ctx.set_option("birth_cut", 1)
        ctx.set_option("max_slots", 0)
    assert rel_err(q, oracle.qgt(c, th)) < TOL


@pytest.mark.parametrize("slots", [5, 6, 9, 17])
def test_qgt_blocked_equals_resident(ctx, oracle, slots):
    # force the column-block schedule (what n >= 26 uses) on a size the oracle can check
    c = K.hea_layers(10, 2)
    th = K.default_angles(c.num_params)
    ctx.set_option("max_slots", slots)
    try:
        q = ctx.qgt(c, th)
        st = ctx.stats()
    finally:
        ctx.set_option("max_slots", 0)
    assert st["blocks"] > 1
    assert rel_err(q, oracle.qgt(c, th)) < TOL


def test_qgt_blocked_qaoa_long_lived_columns(ctx, oracle):
    c = K.qaoa_maxcut(9 + 1, 3)
    th = K.default_angles(c.num_params, 8)
    ctx.set_option("max_slots", 5)
    try:
        q = ctx.qgt(c, th)
    finally:
        ctx.set_option("max_slots", 0)
    assert rel_err(q, oracle.qgt(c, th)) < TOL


def test_analytic_closed_form(ctx):
    t1, t2 = 0.7, 1.3
    c = K.Circuit(1)
    c.rot(K.RY, 0, 0)
    c.rot(K.RZ, 0, 1)
    q = ctx.qgt(c, np.array([t1, t2]))
    expect = np.array([[0.25, 0.25j * np.sin(t1)], [-0.25j * np.sin(t1), 0.25 * np.sin(t1) ** 2]])
    assert np.abs(q - expect).max() < 1e-14


def test_derivative_columns_match_oracle(ctx, oracle):
    c = K.random_circuit(9, 50, 321, kinds=ALL_KINDS, share_params=True)
    th = K.default_angles(max(1, c.num_params), 2)
    for mu in range(c.num_params):
        assert np.abs(ctx.derivative(c, th, mu) - oracle.derivative(c, th, mu)).max() < 1e-13


# P = 32, 64, 96: the projection column is contracted as the strip of the diagonal tiles; 31 / 33: ragged last tile
@pytest.mark.parametrize("P,dim", [(1, 64), (5, 1 << 10), (31, 1 << 10), (32, 1 << 11), (33, 1 << 12), (64, 1 << 10), (70, 1 << 11),
                                   (96, 1 << 12), (130, 1 << 9)])
def test_gram_entry_point_matches_oracle(ctx, oracle, P, dim):
    rng = np.random.default_rng(P)
    psi = rng.normal(size=dim) + 1j * rng.normal(size=dim)
    psi /= np.linalg.norm(psi)
    J = (rng.normal(size=(P, dim)) + 1j * rng.normal(size=(P, dim))) / np.sqrt(dim)
    q = ctx.gram(psi, J)
    assert rel_err(q, oracle.qgt_from_columns(psi, J)) < TOL


# ---- full sizes: properties that need no oracle ------------------------------------------------------------
def test_config2_subblock_and_properties(ctx, oracle):
    c = K.config("c2")                      # 20 qubits, 160 parameters
    th = K.default_angles(c.num_params)
    psi = ctx.state(c.num_qubits)
    q = ctx.qgt(c, th, psi_out=psi)
    assert abs(psi.norm2() - 1.0) < 1e-12
    psi.close()
    assert np.abs(q - q.conj().T).max() < 1e-13                      # Hermitian
    assert np.linalg.eigvalsh(q.real).min() > -1e-12                  # metric is PSD
    # first layer closed forms: RY on |0> has variance 1/4; RZ after RY(t) has sin^2(t)/4
    n = c.num_qubits
    assert np.abs(np.diag(q.real)[:n] - 0.25).max() < 1e-12
    assert np.abs(np.diag(q.real)[n:2 * n] - 0.25 * np.sin(th[:n]) ** 2).max() < 1e-12
    assert np.diag(q.real).max() <= 0.25 + 1e-12                      # Var(P/2) <= 1/4
    # a sub-block against the oracle (3 columns of 2^20 amplitudes take a few seconds on the CPU)
    cols = [0, 41, 159]
    psi_o = oracle.apply(c, th)
    J = np.stack([oracle.derivative(c, th, mu) for mu in cols])
    qo = oracle.qgt_from_columns(psi_o, J)
    assert rel_err(q[np.ix_(cols, cols)], qo) < TOL


def test_large_state_round_trip_and_linearity(ctx):
    # 26 qubits (1 GiB): U^dagger U = 1 via the inverse circuit, and norm preservation
    n = 26
    c = K.hea_layers(n, 1)
    th = K.default_angles(c.num_params)
    inv = K.Circuit(n)
    for (kind, t, ctl, p, a, s) in reversed(c.gates):
        if kind == K.CNOT:
            inv.add(kind, t, ctl)
        else:
            inv.add(kind, t, ctl, -1, -th[p], 1.0)
    st = ctx.state(n).init(0)
    st.apply(c, th)
    assert abs(st.norm2() - 1.0) < 1e-12
    st.apply(inv, np.zeros(1))
    out = st.download()
    st.close()
    assert abs(out[0] - 1.0) < 1e-12
    out[0] = 0
    assert np.abs(out).max() < 1e-12


# ---- natural gradient / energy gradient ("next" rows) -----------------------------------------------------
def test_expectation_gradient_and_natural_gradient(ctx, oracle):
    c = K.qaoa_maxcut(8, 2)
    th = K.default_angles(c.num_params, 6)
    e, g = ctx.expectation_gradient(c, th)
    eo, go = oracle.expectation_gradient(c, th)
    assert abs(e - eo) < 1e-12 and np.abs(g - go).max() < 1e-12
    q = ctx.qgt(c, th)
    x, lam = ctx.natural_gradient(q.real, g)
    xo, lamo = oracle.natural_gradient(q.real, g)
    assert lam == lamo and rel_err(x, xo) < 1e-8


# ---- error behaviour ------------------------------------------------------------------------------------------
def test_invalid_inputs_are_rejected(ctx):
    c = K.Circuit(3)
    c.add(K.CNOT, 1, 1)
    with pytest.raises(api.QgtError) as ei:
        ctx.qgt(c, np.zeros(1))
    assert ei.value.status == -36
    c = K.hea_layers(4, 1)
    st = ctx.state(5)
    with pytest.raises(api.QgtError) as ei:
        st.apply(c, K.default_angles(c.num_params))
    assert ei.value.status == -3
    st.close()


def test_empty_and_parameterless_circuits(ctx, oracle):
    c = K.Circuit(6)
    st = ctx.state(6).init(0)
    st.apply(c, np.zeros(1))
    out = st.download()
    assert out[0] == 1 and np.abs(out[1:]).max() == 0
    c.add(K.H, 2)
    c.add(K.CNOT, 3, 2)
    st.apply(c, np.zeros(1))
    assert np.abs(st.download() - oracle.apply(c, np.zeros(1))).max() < 1e-15
    st.close()
    # a parameter that no gate uses gives a zero row/column
    c2 = K.Circuit(4)
    c2.rot(K.RY, 0, 0)
    c2.rot(K.RX, 1, 2)
    c2.num_params = 3
    q = ctx.qgt(c2, np.array([0.3, 0.0, 0.9]))
    assert np.abs(q[1]).max() == 0 and np.abs(q[:, 1]).max() == 0
    assert abs(q[0, 0] - 0.25) < 1e-14
