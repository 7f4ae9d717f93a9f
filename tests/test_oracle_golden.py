"""Pin the CPU oracle to the reference's own numbers.

G1  rays.dat (reference root): the reference's dff wrote, at full list-directed precision,
    the per-layer horizontal advance and thickness of 20 rays through the test_1 model
    (subroutineR-quiet.f90:157-164).  From them p = sin(theta_i)/v_i and
    T = sum sqrt(h^2+delta^2)/v follow, so p, T and every delta of the oracle are checked
    to 1e-12 relative.
G2  the rendered notebook's stored output: 5 travel times printed with 7 digits and the
    per-layer ray table with 8 digits (raytracerR-export-data-to-MCMC.nb.html:274,357).
G3  test_1_RT.txt = dff output + N(0, 0.016^2): statistical check only.
"""
import math
import os

import numpy as np
import pytest

import oracle


def _geometry(v, h, p):
    cosv = np.sqrt(1.0 - (p * p) * (v * v))
    return np.sqrt((h / cosv) ** 2 - h ** 2)


def test_g1_rays_dat_full_precision(golden):
    c = golden["config1"]
    v, z = np.array(c["vels"]), np.array(c["depths"])
    so, sd = np.array(c["src_offset_full"]), np.array(c["src_depth_full"])
    t, p, tr = oracle.trace_rays(v, z, so, sd, want_trace=True)
    assert len(c["rays_dat"]) == 20
    for k, ray in enumerate(c["rays_dat"]):
        delta, h = np.array(ray["delta"]), np.array(ray["h"])
        nl = tr[k].nl
        assert len(h) == nl == len(delta)
        # thickness row: interface spacings, last entry = source depth - last interface above
        assert abs(h.sum() - sd[k]) <= 1e-12 * sd[k]
        vv = v[:nl]
        hyp = np.sqrt(h * h + delta * delta)
        p_ref = (delta / hyp) / vv            # one estimate of p per layer; all must agree
        t_ref = float((hyp / vv).sum())
        assert np.max(np.abs(p_ref - p[k])) <= 1e-12 * p[k]
        assert abs(t_ref - t[k]) <= 1e-12 * t_ref
        d_or = _geometry(vv, h, p[k])
        assert np.max(np.abs(d_or - delta) / delta) <= 1e-12
        # the horizontal advances add up to the source offset only to the solver's 0.1 m
        # tolerance (subroutineR-quiet.f90:234) -- F3 of SURVEY.md
        assert abs(delta.sum() - so[k]) < 0.1


def test_g1_trajectories_match_survey_appendix_b3(golden):
    c = golden["config1"]
    _, _, tr = oracle.trace_rays(c["vels"], c["depths"], c["src_offset_full"],
                                 c["src_depth_full"], want_trace=True)
    # (branch, bisect iterations, Newton updates) for the first rays, SURVEY.md B.3
    expect = [("bisect", 5, 4), ("bisect", 6, 3), ("bisect", 7, 3), ("bisect", 7, 3),
              ("bisect", 11, 3), ("bisect", 3, 2), ("bisect", 9, 1), ("neg", 0, 3)]
    got = [(oracle.BRANCH_NAMES[x.branch], x.n_bisect, x.n_newton) for x in tr[:8]]
    assert got == expect
    assert all(x.conv == 1 for x in tr)
    assert max(abs(x.f_final) for x in tr) < 0.1


def test_g2_notebook_known_answers(golden):
    n = golden["notebook"]
    v, z = np.array(n["vels"]), np.array(n["depths"])
    so, sd = np.array(n["src_offset"]), np.array(n["src_depth"])
    t, p, tr = oracle.trace_rays(v, z, so, sd, want_trace=True)
    for got, want in zip(t, n["timeP_7digits"]):
        assert abs(got - want) < 0.5e-7            # printed with 7 decimals
    rows = [r for r in n["ray_table"] if r["depth"] != 0.0]  # read.rays prepends a 0 row per ray
    i = 0
    for k in range(5):
        nl = tr[k].nl
        vv = v[:nl]
        h = np.concatenate(([z[0]], np.diff(z), [0.0]))[:nl].copy()
        h[nl - 1] = sd[k] - (z[nl - 2] if nl > 1 else 0.0)
        delta = _geometry(vv, h, p[k])
        for j in range(nl):
            assert rows[i]["ray"] == k + 1
            assert abs(rows[i]["delta"] - delta[j]) < 0.6e-5   # 5 decimals printed
            assert abs(rows[i]["depth"] - h[j]) < 0.6e-5
            i += 1
    assert i == len(rows)


def test_g3_observed_times_statistics(golden):
    c = golden["config1"]
    t, _, _ = oracle.trace_rays(c["vels"], c["depths"], c["src_offset_file"], c["src_depth_file"])
    res = np.array(c["tobs"]) - t
    rms = math.sqrt(float(np.mean(res ** 2)))
    assert 0.008 < rms < 0.025                      # noise sd was 0.016 (…Rmd:89-90)
    ll = oracle.loglhood_from_times(t, c["tobs"], c["sigma_map"])
    # closed form of loglhood.f90:194-196 evaluated independently in numpy
    n = len(res)
    want = math.log(1.0 / (2.0 * math.pi) ** (n / 2.0)) - (float(np.sum(res * res)) / (2 * 0.02 ** 2)
                                                            + n * math.log(0.02))
    assert abs(ll - want) <= 1e-12 * abs(want)
    assert abs(ll - 55.12412980548659) < 1e-6       # SURVEY.md B.4 (file-precision sources)


def test_readme_example_deep_sources(golden):
    r = golden["readme_example"]
    t, _, tr = oracle.trace_rays(r["vels"], r["depths"], r["src_offset"], r["src_depth"],
                                 want_trace=True)
    assert [x.nl for x in tr] == [2, 2, 3]          # exercises nl = NLayers + 1
    want = [0.7681121105884936, 1.0784165599997246, 1.5302120321561707]
    assert np.allclose(t, want, rtol=1e-13, atol=0)


def test_rays_dat_writer_roundtrip(golden, tmp_path):
    c = golden["config1"]
    path = str(tmp_path / "rays.dat")
    oracle.trace_rays(c["vels"], c["depths"], c["src_offset_full"], c["src_depth_full"],
                      keep_delta=10, rays_path=path)
    rows = [[float(x) for x in l.split()] for l in open(path) if l.strip()]
    assert len(rows) == 40
    for k, ray in enumerate(c["rays_dat"]):
        assert np.allclose(rows[2 * k], ray["delta"], rtol=1e-12)
        assert np.allclose(rows[2 * k + 1], ray["h"], rtol=1e-15)
    # list-directed look: 3 blanks, 17 significant digits, 5 blanks
    first = open(path).readline()
    assert first.startswith("   2727.72106786829") and first.rstrip("\n").endswith("     ")


def test_interplayer_novar_sort_matches_a_plain_sort_and_is_deterministic_on_ties():
    """The oracle's restatement of QSORTC2D (quicksort.f90:66-123) used by INTERPLAYER_novar."""
    rng = np.random.default_rng(5)
    for _ in range(200):
        k = int(rng.integers(1, 31))
        dep, vp = rng.uniform(0, 9000, k), rng.uniform(1500, 9000, k)
        _, _, sd_, sv_ = oracle.loglhood_voro(dep, vp, [100.0], [500.0], [0.3], 0.02)
        order = np.argsort(dep, kind="stable")
        assert np.array_equal(sd_, dep[order]) and np.array_equal(sv_, vp[order])
    dep = np.array([0.0, 3000.0, 1000.0, 2000.0, 1000.0])
    vp = np.array([3000.0, 6000.0, 4000.0, 5000.0, 4500.0])
    _, _, sd_, sv_ = oracle.loglhood_voro(dep, vp, [100.0], [500.0], [0.3], 0.02)
    assert sd_.tolist() == [0.0, 1000.0, 1000.0, 2000.0, 3000.0]
    assert sorted(sv_[1:3].tolist()) == [4000.0, 4500.0]
