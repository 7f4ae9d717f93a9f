"""Oracle restatement of the fixed-dimension MH move (SURVEY 8f rows N1 + N2): PROPOSAL,
INTERPLAYER_novar, CHECKBOUNDS2 and EXPLORE_MH_NOVARPAR's accept test, checked rule by rule
against prjmh_temper_rf.f90:725-757,1386-1447,1681-1716 on the reference's test_1 data."""
import json
import math
import os

import numpy as np
import pytest

import oracle
from raytracerfortran_b200 import chains

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")


@pytest.fixture(scope="module")
def cfg1():
    c = json.load(open(GOLD))["config1"]
    so, sd = np.array(c["src_offset_full"]), np.array(c["src_depth_full"])
    v, z = np.array(c["vels"]), np.array(c["depths"])
    t, _, _ = oracle.trace_rays(v, z, so, sd)
    k = len(v)
    voro = np.zeros((1, 2, 10))
    voro[0, 0, 1:k] = z              # node depths: the top node sits at 0, the others at the interfaces
    voro[0, 1, :k] = v
    return {"so": so, "sd": sd, "tobs": t, "k": k, "voro": voro}


def _step(cfg, voro, logL, ivo, iwhich, cauchy, u, beta=1.0, sigma=0.02, prior=None):
    pr = chains.prior_array() if prior is None else prior
    B = voro.shape[0]
    f = lambda x, dt=np.float64: np.full(B, x, dtype=dt)
    return oracle.mh_step_batch(f(cfg["k"], np.int32), voro, logL, f(ivo, np.int32), f(iwhich, np.int32),
                                f(cauchy), f(u), f(beta), f(sigma), pr, cfg["so"], cfg["sd"], cfg["tobs"])


def test_prior_array_matches_read_input():
    pr = chains.prior_array()
    # read_input.f90:207-214 with test_1_parameter.dat (hmin 100.1, hmx 10000.1), pertsdsc = 30
    assert pr[2:7].tolist() == [100.1, 1500.0, 10000.1, 10000.0, 100.1]
    assert pr[0] == (10000.1 - 100.1) / 30.0 and pr[1] == (10000.0 - 1500.0) / 30.0


def test_proposal_and_bounds_rules(cfg1):
    voro, k = cfg1["voro"], cfg1["k"]
    ll0 = np.array([oracle.loglhood_rt(voro[0, 1, :k], voro[0, 0, 1:k], cfg1["so"], cfg1["sd"],
                                       cfg1["tobs"], 0.02)[0]])
    pr = chains.prior_array()
    # the fixed top node's depth is never proposed (:730)
    assert _step(cfg1, voro, ll0, 1, 1, 0.1, 0.5)["accept"][0] == -1
    # a tiny vp step is evaluated; u = 0 always accepts, u = 1 never (ran_uni >= EXP(.) with EXP <= 1 here)
    r = _step(cfg1, voro, ll0, 3, 2, 1e-3, 0.0)
    assert r["accept"][0] == 1
    assert r["voro"][0, 1, 2] == voro[0, 1, 2] + pr[1] * 1e-3            # :1405
    assert r["logL"][0] == r["logL_prop"][0] != ll0[0]
    r = _step(cfg1, voro, ll0, 3, 2, 1e-3, 1.0)
    assert r["accept"][0] == 0 and np.array_equal(r["voro"], voro) and r["logL"][0] == ll0[0]
    # vp pushed past maxlim(2) = 10000 -> outside, never evaluated (:1705-1712, :753-757)
    r = _step(cfg1, voro, ll0, 3, 2, 1e3, 0.0)
    assert r["accept"][0] == -1 and np.isnan(r["logL_prop"][0]) and np.array_equal(r["voro"], voro)
    # a depth step that crosses a neighbour re-sorts the nodes (INTERPLAYER_novar) ...
    step = (voro[0, 0, 3] - voro[0, 0, 2] + 150.0) / pr[0]
    r = _step(cfg1, voro, ll0, 3, 1, step, 0.0)
    assert np.all(np.diff(r["voro_prop"][0, 0, :k]) >= 0)
    assert r["voro_prop"][0, 1, 2] == voro[0, 1, 3] and r["voro_prop"][0, 1, 3] == voro[0, 1, 2]
    # ... and one that leaves a layer thinner than hmin is outside (:1693)
    step = (voro[0, 0, 3] - voro[0, 0, 2] - 50.0) / pr[0]
    assert _step(cfg1, voro, ll0, 3, 1, step, 0.0)["accept"][0] == -1
    # a negative depth is reflected (:1441-1443): node 2 at z1 proposed to -(z1 + 500) lands at z1 + 500
    z1 = voro[0, 0, 1]
    r = _step(cfg1, voro, ll0, 2, 1, -(2 * z1 + 500.0) / pr[0], 0.0)
    assert abs(r["voro_prop"][0, 0, :k] - np.sort(np.r_[0.0, z1 + 500.0, voro[0, 0, 2:k]])).max() < 1e-9
    # below the deepest allowed interface
    assert _step(cfg1, voro, ll0, k, 1, 1e3, 0.0)["accept"][0] == -1


def test_accept_rule(cfg1):
    """reject iff ran_uni >= EXP((logL_new - logL) * beta_mh)   (:744-751)."""
    voro, k = cfg1["voro"], cfg1["k"]
    ll0 = np.array([oracle.loglhood_rt(voro[0, 1, :k], voro[0, 0, 1:k], cfg1["so"], cfg1["sd"],
                                       cfg1["tobs"], 0.02)[0]])
    for beta in (1.0, 0.3):
        r = _step(cfg1, voro, ll0, 4, 2, 0.02, 0.0, beta=beta)
        thr = math.exp((r["logL_prop"][0] - ll0[0]) * beta)
        assert 0.0 < thr < 1.0                                   # the true model is the best one
        for u in (thr * 0.999, thr * 1.001):
            got = _step(cfg1, voro, ll0, 4, 2, 0.02, u, beta=beta)["accept"][0]
            assert got == (0 if u >= thr else 1)


def test_chain_recovers_the_model(cfg1):
    """A few hundred sweeps from a perturbed start: logL climbs back towards the true model's."""
    rng = np.random.default_rng(5)
    k = cfg1["k"]
    B = 8
    voro = np.repeat(cfg1["voro"], B, axis=0)
    voro[:, 1, :k] += rng.normal(0, 300.0, (B, k))
    kk = np.full(B, k, dtype=np.int32)
    sig = np.full(B, 0.02)
    ll = np.array([oracle.loglhood_rt(voro[b, 1, :k], voro[b, 0, 1:k], cfg1["so"], cfg1["sd"],
                                      cfg1["tobs"], 0.02)[0] for b in range(B)])
    start = ll.copy()
    pr = chains.prior_array()
    pr[:2] /= 20.0                                               # smaller steps: a short test
    nacc = nprop = 0
    for sweep in range(60):
        for ivo in range(1, k + 1):
            for iw in (1, 2):
                if ivo == 1 and iw == 1:
                    continue
                u = rng.random((2, B))
                r = oracle.mh_step_batch(kk, voro, ll, np.full(B, ivo, np.int32), np.full(B, iw, np.int32),
                                         np.tan(np.pi * (u[0] - 0.5)), u[1], np.ones(B), sig, pr,
                                         cfg1["so"], cfg1["sd"], cfg1["tobs"])
                voro, ll = r["voro"], r["logL"]
                nacc += int((r["accept"] == 1).sum())
                nprop += B
    assert np.all(ll > start) and 0.02 < nacc / nprop < 0.9
    assert np.all(np.diff(voro[:, 0, :k], axis=1) >= 100.1 - 1e-9)   # every state stayed inside the prior
