"""GPU parity tests of the FUSED schedule (transition matrices contracted inside the sweeps, fused.cu): the same
C-ABI call as every other QGT test, with option "fused" = 1 so the path also runs at sizes whose columns would all
fit.  Tolerance 1e-10 relative (north star, complex double)."""
import numpy as np
import pytest

from quantum_geometric_tensor_b200 import api, circuits as K

from helpers import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-10
ALL_KINDS = [K.X, K.Y, K.Z, K.H, K.S, K.T, K.SDG, K.TDG, K.SX, K.RX, K.RY, K.RZ, K.PHASE, K.CNOT, K.CY, K.CZ,
             K.CH, K.SWAP, K.CRX, K.CRY, K.CRZ, K.ZZ]


def _fused_qgt(ctx, circ, theta, slots=0, psi=None, traj=-1):
    """traj: 0 = phi recomputed next to every column, 1 = phi's tiles published once per run and fetched with
    cp.async.bulk (trajectory mode), -1 = the library's choice"""
    ctx.set_option("fused", 1)
    ctx.set_option("fused_traj", traj)
    ctx.set_option("max_slots", slots)
    try:
        q = ctx.qgt(circ, theta, psi_out=psi)
        st = ctx.stats()
    finally:
        ctx.set_option("fused", -1)
        ctx.set_option("fused_traj", -1)
        ctx.set_option("max_slots", 0)
    return q, st


@pytest.mark.parametrize("traj", [0, 1])
@pytest.mark.parametrize("n,layers,slots", [(11, 2, 0), (12, 2, 0), (13, 2, 0), (14, 2, 12), (13, 1, 9), (12, 3, 14), (16, 2, 16)])
def test_fused_hea_matches_oracle(ctx, oracle, n, layers, slots, traj):
    c = K.hea_layers(n, layers)
    th = K.default_angles(c.num_params)
    psi = ctx.state(n)
    q, st = _fused_qgt(ctx, c, th, slots, psi, traj)
    assert st["fused"] == 1 and st["fused_launches"] > 0 and st["gram_launches"] == 0
    if slots:
        assert st["blocks"] > 1
    assert rel_err(q, oracle.qgt(c, th)) < TOL
    assert np.abs(psi.download() - oracle.apply(c, th)).max() < 1e-13
    psi.close()


@pytest.mark.parametrize("seed", range(8))
def test_fused_random_circuits_shared_parameters(ctx, oracle, seed):
    n = 12 + seed % 3
    c = K.random_circuit(n, 70, 500 + seed, kinds=ALL_KINDS, share_params=(seed % 2 == 1))
    th = K.default_angles(max(1, c.num_params), seed)
    try:
        api.plan_dump_fused(c, th, 64)
    except api.QgtError:
        pytest.skip("plan does not qualify for the fused schedule (falls back to the Gram schedule)")
    q, st = _fused_qgt(ctx, c, th, [0, 5, 9][seed % 3], traj=seed % 2)
    assert st["fused"] == 1
    assert rel_err(q, oracle.qgt(c, th)) < TOL


def test_fused_pipelined_kernel_at_full_tile_size(ctx, oracle):
    # 11-qubit tiles: trajectory mode runs the lean 2 x 16-warp kernel (default), the persistent double-buffered one
    # (cp.async.bulk + mbarrier ring) or the generic 8-warp one
    c = K.hea_layers(15, 2)
    th = K.default_angles(c.num_params)
    ref = oracle.qgt(c, th)
    for slots in (0, 14):
        q, st = _fused_qgt(ctx, c, th, slots, traj=1)
        assert st["fused"] == 1
        assert rel_err(q, ref) < TOL
        for pipeline in (0, 1, 2):
            ctx.set_option("fused_pipeline", pipeline)
            try:
                q2, _ = _fused_qgt(ctx, c, th, slots, traj=1)
            finally:
                ctx.set_option("fused_pipeline", 3)
            assert rel_err(q2, ref) < TOL, pipeline


def test_fused_equals_gram_schedule_at_20_qubits(ctx):
    # both schedules on BASELINE config 2: two independent evaluations of the same tensor
    c = K.config("c2")
    th = K.default_angles(c.num_params)
    q_gram = ctx.qgt(c, th)
    assert ctx.stats()["fused"] == 0
    for traj in (0, 1):
        q_fused, st = _fused_qgt(ctx, c, th, traj=traj)
        assert st["fused"] == 1
        assert rel_err(q_fused, q_gram) < TOL
        q_blocked, st = _fused_qgt(ctx, c, th, slots=40, traj=traj)
        assert st["blocks"] > 1
        assert rel_err(q_blocked, q_gram) < TOL
