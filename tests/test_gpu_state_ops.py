"""GPU parity of the statevector reductions / measurement / sampling / ComplexFloat boundary (state_ops.cu) against
numpy restatements of the reference loops (hardware/quantum_simulator.c:563-729)."""
import numpy as np
import pytest

from quantum_geometric_tensor_b200 import circuits as K

pytestmark = pytest.mark.gpu


def _rand_state(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


@pytest.mark.parametrize("n", [1, 3, 10, 17, 22])
def test_probability_expectation_inner_scale(ctx, n):
    a, b = _rand_state(n, n), _rand_state(n, 100 + n)
    sa, sb = ctx.state(n).upload(a), ctx.state(n).upload(b)
    idx = np.arange(1 << n, dtype=np.uint64)
    p = np.abs(a) ** 2
    for q in {0, n // 2, n - 1}:
        m = 1 << q
        assert abs(sa.probability(m, m) - p[(idx & np.uint64(m)) != 0].sum()) < 1e-13      # P(qubit q = 1), sim_measure_qubit :569-575
    if n >= 3:
        m, w = 0b101, 0b100
        assert abs(sa.probability(m, w) - p[(idx & np.uint64(m)) == np.uint64(w)].sum()) < 1e-13
    par = np.zeros(1 << n, dtype=np.int64)
    for q in range(n):
        par ^= ((idx >> np.uint64(q)) & np.uint64(1)).astype(np.int64)
    assert abs(sa.expectation_z((1 << n) - 1) - ((1 - 2 * par) * p).sum()) < 1e-13      # sim_get_expectation_value "Z" :708-725
    assert abs(sa.expectation_z(1) - ((1 - 2 * (idx & np.uint64(1)).astype(np.int64)) * p).sum()) < 1e-13
    assert abs(sa.inner(sb) - np.vdot(a, b)) < 1e-13
    sa.scale(0.3 - 0.4j)
    assert np.abs(sa.download() - a * (0.3 - 0.4j)).max() < 1e-15
    sa.close(); sb.close()


@pytest.mark.parametrize("n,qubit,uniform", [(4, 0, 0.1), (4, 3, 0.9), (12, 5, 0.4), (20, 19, 0.55)])
def test_measure_collapses_like_the_reference(ctx, n, qubit, uniform):
    a = _rand_state(n, 7 * n + qubit)
    s = ctx.state(n).upload(a)
    outcome, p1 = s.measure(qubit, uniform)
    idx = np.arange(1 << n, dtype=np.uint64)
    one = (idx & np.uint64(1 << qubit)) != 0
    p1_ref = (np.abs(a[one]) ** 2).sum()
    assert abs(p1 - p1_ref) < 1e-13
    assert outcome == (1 if uniform < p1_ref else 0)
    keep = one if outcome else ~one
    ref = np.where(keep, a, 0)
    ref = ref / np.sqrt((np.abs(ref) ** 2).sum())
    assert np.abs(s.download() - ref).max() < 1e-13
    assert abs(s.norm2() - 1.0) < 1e-13
    s.close()


def test_measure_with_readout_error_mixes_the_probability(ctx):
    a = np.array([np.sqrt(0.9), np.sqrt(0.1)], dtype=np.complex128)
    s = ctx.state(1).upload(a)
    # p = 0.1 * 0.8 + 0.9 * 0.2 = 0.26 (quantum_simulator.c:577-582): uniform 0.2 -> outcome 1 although P(1) = 0.1
    outcome, p1 = s.measure(0, 0.2, readout_error=0.2)
    assert outcome == 1 and abs(p1 - 0.1) < 1e-15
    assert np.abs(s.download() - np.array([0, 1])).max() < 1e-15
    s.close()


@pytest.mark.parametrize("n", [3, 14, 15, 21])
def test_sampling_is_the_inverse_cdf(ctx, n):
    a = _rand_state(n, 5 * n)
    s = ctx.state(n).upload(a)
    rng = np.random.default_rng(n)
    u = np.concatenate([rng.uniform(0, 1, 200), [0.0, 1e-12, 0.5, 1 - 1e-12]])
    got = s.sample(u)
    cdf = np.cumsum(np.abs(a) ** 2)
    for r, i in zip(u, got):
        i = int(i)
        lo = cdf[i - 1] if i > 0 else 0.0
        assert lo - 1e-11 <= r < cdf[i] + 1e-11, (r, i)       # first index whose cumulative probability exceeds r (:650-658)
    # exact agreement away from the boundaries
    ref = np.searchsorted(cdf, u, side="right")
    gap = np.minimum(np.abs(cdf[np.minimum(ref, cdf.size - 1)] - u), np.abs(u - np.where(ref > 0, cdf[ref - 1], 0)))
    far = gap > 1e-10
    assert np.array_equal(got[far].astype(np.int64), np.minimum(ref[far], cdf.size - 1))
    s.close()


@pytest.mark.parametrize("n", [2, 13, 20])
def test_complex_float_boundary(ctx, n):
    a = _rand_state(n, n).astype(np.complex64)
    s = ctx.state(n).upload_c64(a)
    assert np.abs(s.download() - a.astype(np.complex128)).max() == 0.0            # widening is exact
    c = K.hea_layers(n, 1)
    th = K.default_angles(c.num_params)
    s.apply(c, th)
    out64 = s.download()
    out32 = s.download_c64()
    assert np.abs(out32 - out64.astype(np.complex64)).max() == 0.0               # one rounding, at the boundary
    s.close()
