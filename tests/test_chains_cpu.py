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


def test_poisson_pk_matches_read_input():
    pk = chains.poisson_pk(3.01, 1, 10)
    for ik in (1, 4, 10):                                        # read_input.f90:80
        assert abs(pk[ik - 1] - math.exp(-3.01) * 3.01 ** ik / math.factorial(ik)) < 1e-15 * pk[ik - 1] * 10


def test_birth_death_rules(cfg1):
    """EXPLORE_MH_NOVARPAR :658-710 with BIRTH_FULL :997-1103 / DEATH_FULL :917-994."""
    voro, k = cfg1["voro"], cfg1["k"]
    so, sd, tobs = cfg1["so"], cfg1["sd"], cfg1["tobs"]
    ll0 = np.array([oracle.loglhood_rt(voro[0, 1, :k], voro[0, 0, 1:k], so, sd, tobs, 0.02)[0]])
    pr, pk = chains.prior_array(), chains.poisson_pk(3.01, 1, 10)
    one = lambda x, dt=np.float64: np.array([x], dtype=dt)

    def bd(u_k, idel=2, u_z=0.5, u_v=0.5, u_acc=0.0, kk=k, kmin=1, kmax=10, v=voro, pk_=pk):
        return oracle.bd_step_batch(one(kk, np.int32), v, ll0, one(u_k), one(idel, np.int32), one(u_z),
                                    one(u_v), one(u_acc), one(1.0), one(0.02), pr, pk_, kmin, kmax,
                                    so, sd, tobs)
    # move choice (:666-680)
    assert bd(0.5)["accept"][0] == 2                             # middle third: stay
    assert bd(0.2)["k_prop"][0] == k + 1 and bd(0.9)["k_prop"][0] == k - 1
    assert bd(0.9, kmin=k)["accept"][0] == 2                     # at kmin no death ...
    assert bd(0.2, kmin=k)["k_prop"][0] == k + 1                 # ... but birth with 1/3
    assert bd(0.2, kmax=k)["k_prop"][0] == k - 1                 # at kmax "birth" third is death
    assert bd(0.5, kmin=k, kmax=k)["accept"][0] == 2 and bd(0.1, kmin=k, kmax=k)["accept"][0] == 2
    # birth: node at maxpert(1)*u_z with vp minlim(2)+maxpert(2)*u_v, sorted in (:1035-1057)
    z_new, v_new = (10000.1 - 100.1) * 0.37, 1500.0 + (10000.0 - 1500.0) * 0.25
    r = bd(0.2, u_z=0.37, u_v=0.25, u_acc=0.0)
    zz = r["voro_prop"][0, 0, :k + 1]
    assert np.all(np.diff(zz) > 0) and z_new in zz and r["voro_prop"][0, 1, list(zz).index(z_new)] == v_new
    # u_acc = 0 accepts whatever the likelihood says unless outside; the state then holds k + 1 nodes
    if r["accept"][0] == 1:
        assert r["k"][0] == k + 1 and np.array_equal(r["voro"], r["voro_prop"])
    # accept threshold includes the Poisson prior ratio (:986/:1094, :689-693)
    thr = math.exp(math.log(pk[k]) - math.log(pk[k - 1]) + (r["logL_prop"][0] - ll0[0]) * 1.0)
    for u in (thr * 0.999, thr * 1.001):
        if 0 < u < 1:
            assert bd(0.2, u_z=0.37, u_v=0.25, u_acc=u)["accept"][0] == (0 if u >= thr else 1)
    # a birth closer than hmin to an interface is outside (:1650-1653)
    u_close = (voro[0, 0, 2] + 50.0) / (10000.1 - 100.1)
    assert bd(0.2, u_z=u_close)["accept"][0] == -1
    # death of node idel: the remaining nodes keep their order, the slot past k - 1 is zero
    r = bd(0.9, idel=3)
    keep = [i for i in range(k) if i != 2]
    assert r["k_prop"][0] == k - 1
    assert np.array_equal(r["voro_prop"][0, :, :k - 1], voro[0][:, keep]) and np.all(r["voro_prop"][0, :, k - 1:] == 0)
    thr = math.exp(math.log(pk[k - 2]) - math.log(pk[k - 1]) + (r["logL_prop"][0] - ll0[0]))
    assert bd(0.9, idel=3, u_acc=min(thr * 1.001, 1.0))["accept"][0] == 0
    # without the Poisson prior (IPOIPR = 0) logPr = 0
    r0 = bd(0.9, idel=3, pk_=None)
    assert r0["logL_prop"][0] == r["logL_prop"][0]


def test_sigma_move_rules(cfg1):
    """EXPLORE_MH :545-575 with PROPOSAL_SDRT :1616-1635."""
    voro, k = cfg1["voro"], cfg1["k"]
    so, sd = cfg1["so"], cfg1["sd"]
    tobs = cfg1["tobs"] + np.random.default_rng(3).normal(0, 0.016, len(so))
    sp = chains.sd_prior_array()
    assert sp.tolist() == [(0.07 - 0.001) / 10.0, 0.001, 0.07]
    ll0 = np.array([oracle.loglhood_rt(voro[0, 1, :k], voro[0, 0, 1:k], so, sd, tobs, 0.05)[0]])
    one = lambda x: np.array([x], dtype=np.float64)
    step = lambda ug, g, ua, s=0.05: oracle.sd_step_batch([k], voro, ll0, one(s), one(ug), one(g), one(ua),
                                                          one(1.0), sp, so, sd, tobs)
    assert step(0.05, -1.0, 0.0)["accept"][0] == 2               # gate: no move 10 % of the time
    r = step(0.5, -1.0, 0.0)                                     # sigma 0.05 -> 0.0431 fits noise 0.016 better
    assert r["accept"][0] == 1 and r["sigma"][0] == 0.05 + sp[0] * -1.0 and r["logL"][0] > ll0[0]
    assert r["logL"][0] == oracle.loglhood_rt(voro[0, 1, :k], voro[0, 0, 1:k], so, sd, tobs, r["sigma"][0])[0]
    assert step(0.5, 5.0, 0.0)["accept"][0] == -1                # 0.0845 > sdmx
    assert step(0.5, -8.0, 0.0)["accept"][0] == -1               # negative
    r = step(0.5, 2.0, 0.0)                                      # a worse sigma: accepted only for small u
    thr = math.exp(r["logL_prop"][0] - ll0[0])
    assert 0 < thr < 1
    assert step(0.5, 2.0, thr * 1.001)["accept"][0] == 0 and step(0.5, 2.0, thr * 0.999)["accept"][0] == 1


def test_ar_move_rules(cfg1):
    """EXPLORE_MH :583-631 with PROPOSAL_ARRT :1521-1552 (IAR = 1)."""
    voro, k = cfg1["voro"], cfg1["k"]
    so, sd = cfg1["so"], cfg1["sd"]
    tobs = cfg1["tobs"] + np.random.default_rng(4).normal(0, 0.016, len(so))
    ap = chains.ar_prior_array()
    assert ap.tolist() == [(0.9 - -0.5) / 10.0, -0.5, 0.9, 0.5]
    pred = oracle.loglhood_rt(voro[0, 1, :k], voro[0, 0, 1:k], so, sd, tobs, 0.02)[1]
    ll_of = lambda idx, a: oracle.loglhood_from_times_ar(pred, tobs, 0.02, idx, a, 0.5)
    one = lambda x, dt=np.float64: np.array([x], dtype=dt)

    def step(idx, a, uc, up, g, ua):
        return oracle.ar_step_batch([k], voro, one(ll_of(idx, a)), one(0.02), one(idx, np.int32), one(a), one(uc),
                                    one(up), one(g), one(ua), one(1.0), ap, so, sd, tobs)
    # no AR parameter yet: birth, uniform over the prior, threshold carries LOG(0.5)
    r = step(0, -1.5, 0.9, 0.25, 0.0, 0.0)
    a_new = 0.25 * (0.9 - -0.5) + -0.5
    assert r["accept"][0] == 1 and r["idxar"][0] == 1 and r["arpar"][0] == a_new
    assert r["logL"][0] == ll_of(1, a_new)
    thr = math.exp(math.log(0.5) + ll_of(1, a_new) - ll_of(0, -1.5))
    if 0 < thr < 1:
        assert step(0, -1.5, 0.9, 0.25, 0.0, thr * 1.001)["accept"][0] == 0
    # AR on: choice uniform >= 0.5 proposes death (arpar = minlim - 1, LOG(2)), below 0.5 a perturbation
    r = step(1, 0.3, 0.7, 0.0, 0.0, 0.0)
    assert r["accept"][0] == 1 and r["idxar"][0] == 0 and r["arpar"][0] == -1.5 and r["logL"][0] == ll_of(0, -1.5)
    r = step(1, 0.3, 0.2, 0.0, 1.0, 0.0)
    assert r["accept"][0] == 1 and r["idxar"][0] == 1 and r["arpar"][0] == 0.3 + ap[0] * 1.0
    assert step(1, 0.3, 0.2, 0.0, 5.0, 0.0)["accept"][0] == -1       # 1.0 > maxlimarRT
    assert step(1, 0.3, 0.2, 0.0, -6.0, 0.0)["accept"][0] == -1      # -0.54 < minlimarRT


def test_enos_rules(cfg1):
    """ENOS = 1 (even-numbered order statistics, Green 1995): PROPOSAL :1418-1431 draws the node
    uniformly between its neighbours with logPr = LOG(zjp1-zp)+LOG(zp-zjm1)-LOG(zjp1-zj)-LOG(zj-zjm1);
    DEATH_FULL :981-991 and BIRTH_FULL :1090-1098 add their order-statistics terms to logPr."""
    voro, k = cfg1["voro"], cfg1["k"]
    so, sd, tobs = cfg1["so"], cfg1["sd"], cfg1["tobs"]
    ll0 = np.array([oracle.loglhood_rt(voro[0, 1, :k], voro[0, 0, 1:k], so, sd, tobs, 0.02)[0]])
    pr, pk = chains.prior_array(), chains.poisson_pk(3.01, 1, 10)
    hmx, hmin = pr[4], pr[6]
    one = lambda x, dt=np.float64: np.array([x], dtype=dt)
    oracle.set_enos(1)
    try:
        # ---- fixed-dimension depth move of node ivo = 3 (between nodes 2 and 4)
        u = 0.3
        zj, zjm1, zjp1 = voro[0, 0, 2], voro[0, 0, 1], voro[0, 0, 3]
        zp = zjm1 + u * (zjp1 - zjm1)
        r = _step(cfg1, voro, ll0, 3, 1, u, 0.0)
        assert r["voro_prop"][0, 0, 2] == zp                          # :1428-1429
        logPr = math.log(zjp1 - zp) + math.log(zp - zjm1) - math.log(zjp1 - zj) - math.log(zj - zjm1)
        thr = math.exp(logPr + (r["logL_prop"][0] - ll0[0]))
        for uu in (thr * 0.999, thr * 1.001):
            if 0 < uu < 1:
                assert _step(cfg1, voro, ll0, 3, 1, u, uu)["accept"][0] == (0 if uu >= thr else 1)
        # the deepest node moves between its upper neighbour and hmx (:1423-1424)
        r = _step(cfg1, voro, ll0, k, 1, 0.5, 0.0)
        assert r["voro_prop"][0, 0, k - 1] == voro[0, 0, k - 2] + 0.5 * (hmx - voro[0, 0, k - 2])
        # velocity moves are Cauchy steps whatever ENOS says (:1404-1405)
        r = _step(cfg1, voro, ll0, 3, 2, 1e-3, 0.0)
        assert r["voro"][0, 1, 2] == voro[0, 1, 2] + pr[1] * 1e-3

        def bd(u_k, idel=2, u_z=0.5, u_v=0.5, u_acc=0.0, pk_=pk):
            return oracle.bd_step_batch(one(k, np.int32), voro, ll0, one(u_k), one(idel, np.int32), one(u_z),
                                        one(u_v), one(u_acc), one(1.0), one(0.02), pr, pk_, 1, 10, so, sd, tobs)
        # ---- death of node idel = 3 (:941-947, :981-991)
        zdel, zj, zjp1 = voro[0, 0, 2], voro[0, 0, 1], voro[0, 0, 3]
        r = bd(0.9, idel=3)
        enos = (2 * math.log(hmx - hmin) - math.log(2.0 * k * (2.0 * k + 1.0)) + math.log(zjp1 - zj)
                - math.log(zdel - zj) - math.log(zjp1 - zdel))
        for pk_, lp in ((pk, math.log(pk[k - 2]) - math.log(pk[k - 1]) + enos), (None, enos)):
            thr = math.exp(lp + (r["logL_prop"][0] - ll0[0]))
            for uu in (thr * 0.999, thr * 1.001):
                if 0 < uu < 1:
                    assert bd(0.9, idel=3, u_acc=uu, pk_=pk_)["accept"][0] == (0 if uu >= thr else 1)
        # ---- birth (:1061-1073, :1090-1098)
        u_z = 0.37
        znew = (hmx - hmin) * u_z
        r = bd(0.2, u_z=u_z, u_v=0.25)
        zz = list(r["voro_prop"][0, 0, :k + 1])
        i = zz.index(znew)
        zj, zjp1 = zz[i - 1], (zz[i + 1] if i + 1 <= k else hmx)
        enos = (math.log(2.0 * k + 2.0) + math.log(2.0 * k + 3.0) - 2 * math.log(hmx - hmin)
                + math.log(znew - zj) + math.log(zjp1 - znew) - math.log(zjp1 - zj))
        if r["accept"][0] != -1:
            thr = math.exp(math.log(pk[k]) - math.log(pk[k - 1]) + enos + (r["logL_prop"][0] - ll0[0]))
            for uu in (thr * 0.999, thr * 1.001):
                if 0 < uu < 1:
                    assert bd(0.2, u_z=u_z, u_v=0.25, u_acc=uu)["accept"][0] == (0 if uu >= thr else 1)
    finally:
        oracle.set_enos(0)


def test_mcmc_workspace_layout():
    """chains.mcmc_workspace_views mirrors the C layout (McmcWs, csrc/rt_internal.h): the views
    tile the workspace exactly, in the documented order, and the size is what the library reports
    (rtb200_mcmc_workspace_bytes is plain arithmetic: no GPU needed)."""
    import torch
    from raytracerfortran_b200 import _lib
    for B, M in ((1, 0), (7, 3), (100, 19)):
        n = _lib.load().rtb200_mcmc_workspace_bytes(B, M)
        assert n == (11 + 2 * M) * B * 8 + (4 + 3 * M) * B * 4
        ws = torch.zeros(n, dtype=torch.uint8)
        v = chains.mcmc_workspace_views(ws, B, M)
        order = ["u_k", "u_z", "u_v", "u_acc_bd", "u_gate", "gauss", "u_acc_sd", "u_choice", "u_prop_ar",
                 "gauss_ar", "u_acc_ar", "dev", "u_acc", "idel", "acc_bd", "acc_sd", "acc_ar", "ivo", "iwhich",
                 "acc_mh"]
        off = 0
        for name in order:
            t = v[name]
            assert t.numel() == 0 or t.data_ptr() - ws.data_ptr() == off, name
            off += t.numel() * t.element_size()
            assert t.shape == ((M, B) if name in ("dev", "u_acc", "ivo", "iwhich", "acc_mh") else (B,))
        assert off == n
    assert _lib.load().rtb200_mcmc_workspace_bytes(0, 5) == 0
