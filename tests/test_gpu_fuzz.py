"""GPU parity on hostile inputs: the fast division path must hand over to the IEEE built-ins
whenever an operand leaves its comfortable range, so that even nonsense models give the oracle's
bits (NaNs compared as NaNs)."""
import numpy as np
import pytest

import oracle
import raytracerfortran_b200 as rt
from raytracerfortran_b200 import workloads

pytestmark = pytest.mark.gpu


def same(got, want):
    g, w = np.asarray(got), np.asarray(want)
    nan = np.isnan(w)
    return np.array_equal(np.isnan(g), nan) and np.array_equal(g[~nan].view(np.uint64), w[~nan].view(np.uint64))


@pytest.mark.parametrize("variant", [0, 1, 3, 4, 5])
def test_hostile_models(variant):
    rt.set_option("variant", variant)
    rng = np.random.default_rng(2025)
    B, L, S = 400, 8, 48
    v, z, nl = workloads.make_models(B, L, 31)
    so, sd = workloads.make_sources(S, 31)
    # duplicate interface depths (zero-thickness layers) and sources exactly on interfaces
    z[0:40, 3] = z[0:40, 2]
    sd[:6] = z[0, :6]
    # non-monotone interfaces, negative and zero velocities, huge and tiny scales
    z[40:80] = rng.permutation(z[40:80].T).T
    v[80:100, 2] = -v[80:100, 2]
    v[100:110, 1] = 0.0
    v[110:130] *= 1e12
    v[130:150] *= 1e-12
    z[150:170] *= 1e-9
    z[170:190] *= 1e9
    # NaN and inf entries
    v[190:200, 3] = np.nan
    z[200:210, 4] = np.nan
    v[210:220, 0] = np.inf
    z[220:230, 1] = np.inf
    # ragged layer counts, including none
    nl[230:300] = rng.integers(0, L + 1, 70)
    # sources: zero / negative offsets and depths, a NaN, an inf
    so[6], sd[7] = 0.0, 0.0
    so[8], sd[9] = -500.0, -10.0
    so[10], sd[11] = np.nan, np.nan
    so[12] = np.inf
    tobs, sigma = workloads.make_observations(np.ones(S), B, 31)
    sigma[300:310] = 0.0
    sigma[310:320] = -0.02
    sigma[320:330] = np.nan
    with np.errstate(all="ignore"):
        ref = oracle.dff_batch(v, z, nl, so, sd, tobs=tobs, sigma=sigma, want_p=True)
    got = rt.dff_batch(v, z, nl, so, sd, tobs=tobs, sigma=sigma, want_p=True)
    rt.set_option("variant", -1)
    assert same(got["timeP"], ref["timeP"])
    assert same(got["p"], ref["p"])
    # logL: identical where the terms are identical; the device log may differ in the last ulp
    w, g = ref["logL"], got["logL"]
    fin = np.isfinite(w)
    assert np.array_equal(np.isfinite(g), fin) and np.array_equal(g[~fin], w[~fin])
    with np.errstate(all="ignore"):
        tol = 1e-12 * np.maximum(np.abs(w[fin]), S * np.abs(np.log(np.abs(sigma[fin]) + 1e-300)))
    assert np.all((g[fin] == w[fin]) | (np.abs(g[fin] - w[fin]) <= tol))


def test_random_small_batches_many_shapes():
    rng = np.random.default_rng(7)
    for trial in range(25):
        B = int(rng.integers(1, 70))
        L = int(rng.integers(1, 20))
        S = int(rng.integers(1, 90))
        v, z, nl = workloads.make_models(B, L, 100 + trial)
        nl = rng.integers(1, L + 1, B).astype(np.int32)
        so, sd = workloads.make_sources(S, 100 + trial, near_critical=bool(trial % 2))
        ref = oracle.dff_batch(v, z, nl, so, sd, want_p=True)
        got = rt.dff_batch(v, z, nl, so, sd, want_p=True)
        assert same(got["timeP"], ref["timeP"]), (trial, B, L, S)
        assert same(got["p"], ref["p"]), (trial, B, L, S)


def test_extreme_shapes():
    """Many sources for one model (hundreds of source chunks), and the deepest model the packed
    ray word allows (254 interfaces)."""
    v, z, nl = workloads.make_models(1, 12, 3)
    so, sd = workloads.make_sources(70001, 3)
    got = rt.dff(v[0], z[0], so, sd)
    want, _, _ = oracle.trace_rays(v[0], z[0], so, sd)
    assert same(got, want)
    rng = np.random.default_rng(9)
    B, L = 5, 254
    vv = rng.uniform(1500, 10000, (B, L + 1))
    zz = np.sort(rng.uniform(50, 10000, (B, L)), axis=1)
    nn = np.array([254, 254, 200, 1, 77], dtype=np.int32)
    so, sd = workloads.make_sources(40, 9, near_critical=True)
    ref = oracle.dff_batch(vv, zz, nn, so, sd, want_p=True)
    out = rt.dff_batch(vv, zz, nn, so, sd, want_p=True)
    assert same(out["timeP"], ref["timeP"]) and same(out["p"], ref["p"])
    with pytest.raises(rt.RayTraceError):
        rt.dff_batch(np.ones((2, 300)), np.ones((2, 299)), np.array([10, 10], dtype=np.int32), so, sd)
