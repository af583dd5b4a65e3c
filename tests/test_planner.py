"""CPU tests of the host-side planner (fusion into runs, derivative overrides, column schedule) and of
the sweep kernel's index arithmetic through the host emulation.  No GPU needed."""
import numpy as np
import pytest

from quantum_geometric_tensor_b200 import api, circuits as K

import emul_lib
import plan_interp as pi

ALL_KINDS = [K.X, K.Y, K.Z, K.H, K.S, K.T, K.SDG, K.TDG, K.SX, K.RX, K.RY, K.RZ, K.PHASE, K.CNOT, K.CY, K.CZ,
             K.CH, K.SWAP, K.CRX, K.CRY, K.CRZ, K.ZZ]


def _run_program_emulated(plan, circ, theta, Kt, R):
    """Same as plan_interp.run_program but every sweep item goes through the emulated CUDA kernel."""
    prog = plan["program"]
    P = plan["P"]
    dim = 1 << circ.num_qubits
    slots = [np.zeros(dim, dtype=np.complex128) for _ in range(prog["slots"])]
    Cm = np.zeros((P + 1, P + 1), dtype=np.complex128)
    for ins in prog["instrs"]:
        k = ins["k"]
        if k == "init":
            slots[ins["dst"]] = pi.initial_state(circ)
        elif k == "copy":
            slots[ins["dst"]] = slots[ins["src"]].copy()
        elif k == "sweep":
            res = [(dst, emul_lib.sweep(circ, theta, Kt, R, ins["run"], slots[src], ovr, slots[dst], bool(acc), extra))
                   for (src, dst, ovr, acc, extra) in ins["cols"]]
            for dst, v in res:
                slots[dst] = v
        else:
            for sa, ia in zip(ins["a"], ins["aid"]):
                for sb, ib in zip(ins["b"], ins["bid"]):
                    val = np.vdot(slots[sa], slots[sb])
                    Cm[ia, ib] = val
                    Cm[ib, ia] = np.conj(val)
    v = Cm[:P, P]
    return Cm[:P, :P] - np.outer(v, v.conj()), slots[prog["psi"]]


@pytest.mark.parametrize("n,layers,Kt,R", [(5, 2, 4, 2), (6, 1, 5, 3), (7, 2, 11, 3), (3, 2, 11, 3), (8, 1, 6, 3)])
def test_plan_forward_matches_oracle(oracle, n, layers, Kt, R):
    c = K.hea_layers(n, layers)
    th = K.default_angles(c.num_params)
    plan = api.plan_dump(c, th, tile_qubits=Kt, reg_qubits=R)
    psi = pi.apply_plan(plan, c, pi.initial_state(c))
    assert np.abs(psi - oracle.apply(c, th)).max() < 1e-14
    # every gate is scheduled exactly once
    gates = sorted(op["gate"] for r in plan["runs"] for op in r["ops"])
    assert gates == list(range(len(c.gates)))


@pytest.mark.parametrize("seed", range(6))
def test_plan_random_circuits_all_kinds(oracle, seed):
    n = 2 + seed % 5
    c = K.random_circuit(n, 30 + 5 * seed, seed, kinds=ALL_KINDS, share_params=True)
    th = K.default_angles(max(1, c.num_params), seed + 1)
    plan = api.plan_dump(c, th, tile_qubits=max(3, n - 1), reg_qubits=2, column_slots=256)
    Q, psi, _, seen = pi.run_program(plan, c)
    assert np.abs(psi - oracle.apply(c, th)).max() < 1e-13
    assert np.abs(Q - oracle.qgt(c, th)).max() < 1e-12


@pytest.mark.parametrize("slots", [5, 6, 7, 9, 13, 40])
def test_blocked_schedule_covers_every_pair(oracle, slots):
    c = K.hea_layers(5, 2)
    th = K.default_angles(c.num_params)
    plan = api.plan_dump(c, th, tile_qubits=4, reg_qubits=2, column_slots=slots)
    Q, psi, cnt, seen = pi.run_program(plan, c)
    P = c.num_params
    assert seen[:P, :P].min() >= 1, "a (mu, nu) pair was never computed"
    assert seen[:P, P].min() >= 1, "a projection <d_mu psi|psi> was never computed"
    assert np.abs(Q - oracle.qgt(c, th)).max() < 1e-13
    assert np.abs(psi - oracle.apply(c, th)).max() < 1e-14


@pytest.mark.parametrize("slots", [5, 8, 64])
def test_shared_parameters_and_qaoa_schedule(oracle, slots):
    c = K.qaoa_maxcut(6, 3)
    th = K.default_angles(c.num_params, 11)
    plan = api.plan_dump(c, th, tile_qubits=4, reg_qubits=2, column_slots=slots)
    Q, psi, cnt, seen = pi.run_program(plan, c)
    assert np.abs(Q - oracle.qgt(c, th)).max() < 1e-12
    assert np.abs(psi - oracle.apply(c, th)).max() < 1e-13


@pytest.mark.parametrize("slots", [5, 7, 64])
def test_long_lived_shared_parameters_multi_pass(oracle, slots):
    # parameters reused far apart in the circuit: streaming columns stay alive across many runs
    n = 5
    c = K.Circuit(n)
    for rep in range(3):
        for q in range(n):
            c.add(K.RY, q, -1, q, 0.1 * rep, 1.0 + 0.5 * rep)
        for q in range(n - 1):
            c.add(K.CNOT, q + 1, q)
        for q in range(n):
            c.add(K.RZ, q, -1, n + q, 0.0, 1.0)
    th = K.default_angles(c.num_params, 5)
    plan = api.plan_dump(c, th, tile_qubits=4, reg_qubits=2, column_slots=slots)
    Q, psi, cnt, seen = pi.run_program(plan, c)
    assert seen[:c.num_params, :c.num_params].min() >= 1
    assert np.abs(Q - oracle.qgt(c, th)).max() < 1e-12


@pytest.mark.parametrize("n,Kt,R", [(1, 11, 3), (2, 11, 3), (3, 11, 3), (4, 11, 3), (6, 5, 3), (7, 6, 2), (9, 8, 3), (10, 11, 3), (12, 11, 3), (12, 9, 3)])
def test_emulated_kernel_forward(oracle, n, Kt, R):
    c = K.random_circuit(n, 40, 100 + n, kinds=ALL_KINDS if n > 1 else [K.X, K.Y, K.H, K.RX, K.RZ, K.T, K.SX])
    th = K.default_angles(max(1, c.num_params), 3)
    nruns = emul_lib.load().emul_num_runs(c.to_c(), th.ctypes.data_as(emul_lib._DP), Kt, R)
    assert nruns >= 1
    v = pi.initial_state(c)
    for r in range(nruns):
        v = emul_lib.sweep(c, th, Kt, R, r, v)
    assert np.abs(v - oracle.apply(c, th)).max() < 1e-13


@pytest.mark.parametrize("seed,slots", [(0, 64), (1, 6), (2, 9), (3, 64)])
def test_emulated_kernel_full_qgt(oracle, seed, slots):
    n = 5 + seed % 3
    kinds = ALL_KINDS
    c = K.random_circuit(n, 36, 40 + seed, kinds=kinds, share_params=(seed % 2 == 1))
    th = K.default_angles(max(1, c.num_params), seed)
    Kt, R = 4 + seed % 2, 2 + seed % 2
    plan = api.plan_dump(c, th, tile_qubits=Kt, reg_qubits=R, column_slots=slots)
    Q, psi = _run_program_emulated(plan, c, th, Kt, R)
    assert np.abs(psi - oracle.apply(c, th)).max() < 1e-13
    assert np.abs(Q - oracle.qgt(c, th)).max() < 1e-12


def test_emulated_kernel_qaoa_cost_and_derivative(oracle):
    c = K.qaoa_maxcut(6, 2)
    c.vertex_weights = [0.3, -0.2, 0.1, 0.0, 0.5, -0.4]
    th = K.default_angles(c.num_params, 9)
    plan = api.plan_dump(c, th, tile_qubits=5, reg_qubits=3, column_slots=32)
    Q, psi = _run_program_emulated(plan, c, th, 5, 3)
    assert np.abs(psi - oracle.apply(c, th)).max() < 1e-13
    assert np.abs(Q - oracle.qgt(c, th)).max() < 1e-12


def test_planner_rejects_bad_circuits():
    c = K.Circuit(3)
    c.add(K.CNOT, 1, 1)
    with pytest.raises(api.QgtError):
        api.plan_dump(c)
    c = K.Circuit(3)
    c.add(K.RX, 5, -1, 0)
    with pytest.raises(api.QgtError):
        api.plan_dump(c)
    c = K.Circuit(3)
    c.add(77, 0)
    with pytest.raises(api.QgtError):
        api.plan_dump(c)


def test_fusion_depth_on_hea():
    # the 28-qubit ansatz must fuse into a handful of sweeps per layer, not one per gate
    c = K.hea(28, 256)
    plan = api.plan_dump(c, K.default_angles(256))
    assert len(plan["runs"]) <= 24, len(plan["runs"])
    for r in plan["runs"]:
        assert len(r["tile"]) == 11 and r["tile"][:4] == [0, 1, 2, 3]


def test_stage_matrix_forms_are_detected():
    """One ansatz layer per stage factors as (diagonal) x (real); RX layers next to a cost pass (QAOA mixer) take the parity form
    (real part on even, imaginary part on odd index distance; 2), RX layers of other circuits and two fused layers are dense."""
    def forms(circ):
        d = api.plan_dump(circ, K.default_angles(max(1, circ.num_params)))
        return [f for r in d["runs"] for s in r["subs"] for f in s["forms"]]

    one_layer = forms(K.hea_layers(12, 1))
    assert one_layer and all(f == 1 for f in one_layer)
    qaoa = forms(K.qaoa_maxcut(12, 2))
    assert qaoa and all(f == 2 for f in qaoa)
    rx_only = K.Circuit(12)
    for q in range(12):
        rx_only.rot(K.RX, q, q)
    plain = forms(rx_only)
    assert plain and all(f == 0 for f in plain)   # no cost pass: the plan may run in the fused kernels, which do not know the form
    deep = forms(K.config("c2"))
    assert 0 in deep and 1 in deep              # two fused layers are dense


@pytest.mark.parametrize("slots", [0, 7])
def test_runs_split_where_parameters_are_born(oracle, slots, monkeypatch):
    """A long first run is executed in pieces (same tile, consecutive sub-pass ranges); the plan and its column
    schedule must still reproduce the oracle.  QGT_B200_BIRTH_CUT=2 forces the split on a small state."""
    c = K.hea_layers(9, 3)
    th = K.default_angles(c.num_params, 5)
    monkeypatch.setenv("QGT_B200_BIRTH_CUT", "0")
    whole = api.plan_dump(c, th, tile_qubits=7)
    monkeypatch.setenv("QGT_B200_BIRTH_CUT", "2")
    d = api.plan_dump(c, th, tile_qubits=7, column_slots=slots or c.num_params + 2)
    assert len(d["runs"]) > len(whole["runs"])
    assert sum(len(r["subs"]) for r in d["runs"]) == sum(len(r["subs"]) for r in whole["runs"])
    q, psi, _, seen = pi.run_program(d, c)
    assert (seen[:c.num_params, :c.num_params] >= 1).all()
    assert np.abs(q - oracle.qgt(c, th)).max() < 1e-12
    assert np.abs(psi - oracle.apply(c, th)).max() < 1e-13


def test_parameter_free_circuit_applies_every_run_once(oracle):
    """with P = 0 and psi requested the program applied the circuit twice (psi = U U|0>)."""
    c = K.Circuit(3).add(K.H, 1)
    plan = api.plan_dump(c, np.zeros(1), column_slots=300)
    kinds = [i["k"] for i in plan["program"]["instrs"]]
    assert kinds == ["init", "sweep"]
    Q, psi, _, _ = pi.run_program(plan, c)
    assert np.abs(psi - oracle.apply(c, np.zeros(1))).max() < 1e-15
    for seed in range(8):
        c = K.random_circuit(4, 25, 900 + seed, kinds=[K.X, K.Y, K.Z, K.H, K.S, K.T, K.CNOT, K.CZ, K.SWAP])
        assert c.num_params == 0
        plan = api.plan_dump(c, np.zeros(1), tile_qubits=3, reg_qubits=2, column_slots=16)
        _, psi, _, _ = pi.run_program(plan, c)
        assert np.abs(psi - oracle.apply(c, np.zeros(1))).max() < 1e-13


def test_blocked_fused_schedule_keeps_phi_images_in_one_column():
    """Blocked fused schedule: the executor walks a run's launches range by range over the tiles, so the images of phi after
    every transition-matrix stage share ONE column (traj_ranges = 8) and all other columns stay resident; when everything
    fits, every stage keeps its own full image and a run is one launch pair."""
    c = K.config("t30")
    th = K.default_angles(c.num_params)
    prog = api.plan_dump_fused(c, th, 9)["program"]
    # 8 ranges, or 16 when a run has more than 4 transition-matrix stages (two ranges' images at once: overlap of phi's launch)
    assert prog["fused"] == 1 and prog["traj_ranges"] in (8, 16) and len(prog["traj"]) == 1
    assert prog["resident"] == 9 - 4 and prog["blocks"] == -(-c.num_params // 5)
    c2 = K.config("c2")
    prog = api.plan_dump_fused(c2, K.default_angles(c2.num_params), 400)["program"]
    assert prog["traj_ranges"] == 1 and len(prog["traj"]) >= 2 and prog["blocks"] == 1
    # ... and a blocked schedule whose images cost only a small share of the columns keeps them whole (28 qubits: 39 columns)
    c3 = K.config("c3")
    prog = api.plan_dump_fused(c3, K.default_angles(c3.num_params), 39)["program"]
    assert prog["traj_ranges"] == 1 and prog["blocks"] > 1 and prog["resident"] == 39 - 3 - len(prog["traj"])
    # states with fewer than 8 tiles keep the one-image-per-stage layout even when blocked
    small = K.hea_layers(12, 2)
    prog = api.plan_dump_fused(small, K.default_angles(small.num_params), 12)["program"]
    assert prog["traj_ranges"] == 1


def test_blocked_gram_split_is_chosen_by_traffic():
    """Few parameters on a state that leaves room for a handful of columns (BASELINE config 3: 30-qubit QAOA, 16 parameters,
    9 columns): the planner builds the blocked Gram program for every resident / streaming split and keeps the one that moves
    the fewest bytes - 4 + 3 instead of the 5 + 2 of the fixed rule (601 instead of 810 column passes)."""
    c = K.config("c4")
    prog = api.plan_dump(c, K.default_angles(c.num_params), column_slots=9)["program"]
    assert prog["fused"] == 0 and prog["resident"] + prog["streaming"] == 9 - 2
    passes = sum(len(i["cols"]) for i in prog["instrs"] if i["k"] == "sweep")
    assert prog["streaming"] >= 3 and passes <= 601
    # every pair of parameters (and every parameter with psi) is covered exactly by the union of the Gram instructions:
    # the choice of the split must not change what is computed
    P = c.num_params
    seen = set()
    for i in prog["instrs"]:
        if i["k"] != "gram":
            continue
        for a in i["aid"]:
            for b in i["bid"]:
                seen.add((min(a, b), max(a, b)))
    want = {(a, b) for a in range(P) for b in range(a, P + 1)}
    assert want <= seen
