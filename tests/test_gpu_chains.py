"""SURVEY 8f rows N1 + N2 on the GPU: rtb200_mh_step_device (PROPOSAL + INTERPLAYER_novar +
CHECKBOUNDS2 + LOGLHOOD + accept, prjmh_temper_rf.f90:725-757) against the CPU oracle's
restatement on the same random numbers."""
import numpy as np
import pytest

import oracle
from raytracerfortran_b200 import chains, workloads

pytestmark = pytest.mark.gpu


def _random_chains(B, ldk, seed, hmin=100.1, hmx=10000.1):
    rng = np.random.default_rng(seed)
    k = rng.integers(1, ldk + 1, B).astype(np.int32)
    voro = np.zeros((B, 2, ldk))
    for b in range(B):
        n = int(k[b])
        # the top node at depth 0, the others at sorted depths at least hmin apart
        gaps = hmin + rng.random(n - 1) * (hmx - hmin * n) / max(n, 1) if n > 1 else np.zeros(0)
        voro[b, 0, 1:n] = np.cumsum(gaps)
        voro[b, 1, :n] = rng.uniform(1500.0, 10000.0, n)
    return k, voro


def _setup(B, ldk, nsrc, seed):
    k, voro = _random_chains(B, ldk, seed)
    so, sd = workloads.make_sources(nsrc, seed)
    tobs, sigma = workloads.make_observations(np.full(nsrc, 1.3), B, seed)
    ll = np.array([oracle.loglhood_rt(voro[b, 1, :k[b]], voro[b, 0, 1:k[b]], so, sd, tobs, sigma[b])[0]
                   for b in range(B)])
    return k, voro, so, sd, tobs, sigma, ll


def _dev(*arrays):
    import torch
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrays]


def test_mh_step_matches_oracle():
    import torch
    B, ldk, nsrc = 3000, 12, 24
    k, voro, so, sd, tobs, sigma, ll = _setup(B, ldk, nsrc, 11)
    rng = np.random.default_rng(12)
    ivo = rng.integers(1, k + 2).astype(np.int32)            # includes ivo = k + 1 (no such node)
    iwhich = rng.integers(1, 3, B).astype(np.int32)
    u = rng.random((2, B))
    cauchy = np.tan(np.pi * (u[0] - 0.5))
    cauchy[::7] *= 0.01                                      # a good share of small, acceptable steps
    beta = 1.0 / 1.4 ** rng.integers(0, 8, B)
    prior = chains.prior_array()
    want = oracle.mh_step_batch(k, voro, ll, ivo, iwhich, cauchy, u[1], beta, sigma, prior, so, sd, tobs)
    tk, tv, tl, ti, tw, tc, tu, tb, tg, ts, td, to = _dev(k, voro, ll, ivo, iwhich, cauchy, u[1], beta,
                                                          sigma, so, sd, tobs)
    acc = chains.mh_step_device(tk, tv, tl, ti, tw, tc, tu, tb, tg, prior, ts, td, to)
    torch.cuda.synchronize()
    acc = acc.cpu().numpy()
    assert np.array_equal(acc, want["accept"])
    assert (acc == 1).sum() > 100 and (acc == 0).sum() > 100 and (acc == -1).sum() > 100
    assert np.array_equal(tv.cpu().numpy().view(np.uint64), want["voro"].view(np.uint64))
    got_ll = tl.cpu().numpy()
    scale = np.maximum(np.abs(want["logL"]), nsrc * np.abs(np.log(sigma)))
    assert np.all(np.abs(got_ll - want["logL"]) <= 1e-12 * scale)
    assert np.array_equal(got_ll[acc != 1], ll[acc != 1])    # rejected chains keep their logL bits


def test_mh_trajectory_matches_oracle():
    """Forty moves in a row: the same decisions and bit-identical chain states at the end."""
    import torch
    B, ldk, nsrc = 512, 8, 20
    k, voro, so, sd, tobs, sigma, ll = _setup(B, ldk, nsrc, 21)
    rng = np.random.default_rng(22)
    beta = 1.0 / 1.4 ** rng.integers(0, 4, B)
    prior = chains.prior_array()
    prior[:2] /= 10.0
    tk, tv, tl, tb, tg, ts, td, to = _dev(k, voro, ll, beta, sigma, so, sd, tobs)
    cur_v, cur_l = voro, ll
    total = 0
    for step in range(40):
        ivo = (1 + step % ldk) * np.ones(B, dtype=np.int32)
        iwhich = (1 + (step // ldk) % 2) * np.ones(B, dtype=np.int32)
        u = rng.random((2, B))
        cauchy = np.tan(np.pi * (u[0] - 0.5))
        r = oracle.mh_step_batch(k, cur_v, cur_l, ivo, iwhich, cauchy, u[1], beta, sigma, prior, so, sd, tobs)
        cur_v, cur_l = r["voro"], r["logL"]
        ti, tw, tc, tu = _dev(ivo, iwhich, cauchy, u[1])
        acc = chains.mh_step_device(tk, tv, tl, ti, tw, tc, tu, tb, tg, prior, ts, td, to)
        assert np.array_equal(acc.cpu().numpy(), r["accept"]), f"step {step}"
        total += int((r["accept"] == 1).sum())
    assert total > 500
    assert np.array_equal(tv.cpu().numpy().view(np.uint64), cur_v.view(np.uint64))


def test_mh_sweep_device_runs_and_respects_the_prior():
    import torch
    B, ldk, nsrc = 2048, 10, 32
    k, voro, so, sd, tobs, sigma, ll = _setup(B, ldk, nsrc, 31)
    beta = np.ones(B)
    prior = chains.prior_array()
    tk, tv, tl, tb, tg, ts, td, to = _dev(k, voro, ll, beta, sigma, so, sd, tobs)
    gen = torch.Generator(device="cuda").manual_seed(7)
    acc_n, prop_n = chains.mh_sweep_device(tk, tv, tl, tb, tg, prior, ts, td, to, generator=gen)
    torch.cuda.synchronize()
    assert int(prop_n.sum()) == int((2 * k - 1).sum())       # every node's vp, every node's depth but the top one
    assert 0 < int(acc_n.sum()) < int(prop_n.sum())
    v = tv.cpu().numpy()
    for b in range(0, B, 37):
        n = int(k[b])
        z = v[b, 0, :n]
        assert z[0] == 0.0 and np.all(np.diff(z) >= 100.1) and (n == 1 or z[-1] <= 10000.1)
        assert np.all((v[b, 1, :n] >= 1500.0) & (v[b, 1, :n] <= 10000.0))
    # the stored logL is the likelihood of the stored state
    ref = np.array([oracle.loglhood_rt(v[b, 1, :k[b]], v[b, 0, 1:k[b]], so, sd, tobs, sigma[b])[0]
                    for b in range(0, B, 37)])
    got = tl.cpu().numpy()[::37]
    assert np.all(np.abs(got - ref) <= 1e-12 * np.maximum(np.abs(ref), nsrc * np.abs(np.log(sigma[::37]))))


def test_mh_moves_device_walks_each_chains_own_sweep():
    """mh_moves_device: chain b visits (1,2), (2,1), (2,2), ..., (k_b,2) and wraps; replayed on the
    oracle with the same random numbers the chains end bit-identical."""
    import math
    import torch
    B, ldk, nsrc, n_moves = 300, 6, 16, 9
    k, voro, so, sd, tobs, sigma, ll = _setup(B, ldk, nsrc, 41)
    beta = np.ones(B)
    prior = chains.prior_array()
    prior[:2] /= 10.0
    tk, tv, tl, tb, tg, ts, td, to = _dev(k, voro, ll, beta, sigma, so, sd, tobs)
    pos = torch.zeros(B, dtype=torch.int32, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(99)
    nacc = chains.mh_moves_device(tk, tv, tl, pos, n_moves, tb, tg, prior, ts, td, to, generator=gen)
    gen = torch.Generator(device="cuda").manual_seed(99)
    u = torch.rand((2, n_moves, B), dtype=torch.float64, device="cuda", generator=gen)
    cauchy, uacc = torch.tan(math.pi * (u[0] - 0.5)).cpu().numpy(), u[1].cpu().numpy()
    cur_v, cur_l, total = voro, ll, np.zeros(B, dtype=np.int64)
    for m in range(n_moves):
        j = m % (2 * k - 1) + 1                              # 1-based position in the chain's own sweep
        ivo, iwhich = (j // 2 + 1).astype(np.int32), (j % 2 + 1).astype(np.int32)
        assert np.all(ivo <= k) and not np.any((ivo == 1) & (iwhich == 1))
        r = oracle.mh_step_batch(k, cur_v, cur_l, ivo, iwhich, cauchy[m], uacc[m], beta, sigma, prior,
                                 so, sd, tobs)
        cur_v, cur_l = r["voro"], r["logL"]
        total += (r["accept"] == 1)
    assert np.array_equal(nacc.cpu().numpy(), total)
    assert np.array_equal(tv.cpu().numpy().view(np.uint64), cur_v.view(np.uint64))
    assert np.array_equal(pos.cpu().numpy(), n_moves % (2 * k - 1))


def test_bd_step_matches_oracle():
    """rtb200_bd_step_device (birth/death, :658-710) against the oracle: same move choice, same
    proposals, same decisions, bit-identical states; then a trajectory of mixed BD + MH moves."""
    import torch
    B, ldk, nsrc = 3000, 12, 24
    kmin, kmax = 1, ldk
    k, voro, so, sd, tobs, sigma, ll = _setup(B, ldk, nsrc, 51)
    rng = np.random.default_rng(52)
    prior, pk = chains.prior_array(), chains.poisson_pk(3.01, kmin, kmax)
    beta = 1.0 / 1.4 ** rng.integers(0, 6, B)
    tk, tv, tl, tb, tg, ts, td, to = _dev(k, voro, ll, beta, sigma, so, sd, tobs)
    cur_k, cur_v, cur_l = k.copy(), voro, ll
    seen = {1: 0, 0: 0, -1: 0, 2: 0}
    for step in range(12):
        u = rng.random((4, B))
        idel = (2 + np.floor(rng.random(B) * np.maximum(cur_k - 1, 1))).astype(np.int32)
        r = oracle.bd_step_batch(cur_k, cur_v, cur_l, u[0], idel, u[1], u[2], u[3], beta, sigma, prior,
                                 pk, kmin, kmax, so, sd, tobs)
        t_u = _dev(u[0], u[1], u[2], u[3])
        (t_idel,) = _dev(idel)
        acc = chains.bd_step_device(tk, tv, tl, t_u[0], t_idel, t_u[1], t_u[2], t_u[3], tb, tg, prior, pk,
                                    kmin, kmax, ts, td, to)
        acc = acc.cpu().numpy()
        assert np.array_equal(acc, r["accept"]), f"step {step}"
        assert np.array_equal(tk.cpu().numpy(), r["k"])
        assert np.array_equal(tv.cpu().numpy().view(np.uint64), r["voro"].view(np.uint64)), f"step {step}"
        cur_k, cur_v, cur_l = r["k"], r["voro"], r["logL"]
        for c in seen:
            seen[c] += int((acc == c).sum())
        # interleave a fixed-dimension move so both kinds act on each other's states
        ivo = np.minimum(1 + step % 4, cur_k).astype(np.int32)
        iwhich = np.full(B, 2, dtype=np.int32)
        uu = rng.random((2, B))
        cauchy = 0.05 * np.tan(np.pi * (uu[0] - 0.5))
        r2 = oracle.mh_step_batch(cur_k, cur_v, cur_l, ivo, iwhich, cauchy, uu[1], beta, sigma, prior, so, sd, tobs)
        ti, tw, tc, tu = _dev(ivo, iwhich, cauchy, uu[1])
        acc2 = chains.mh_step_device(tk, tv, tl, ti, tw, tc, tu, tb, tg, prior, ts, td, to)
        assert np.array_equal(acc2.cpu().numpy(), r2["accept"]), f"mh step {step}"
        cur_v, cur_l = r2["voro"], r2["logL"]
    assert np.array_equal(tv.cpu().numpy().view(np.uint64), cur_v.view(np.uint64))
    assert all(n > 50 for n in seen.values()), seen
    got_ll = tl.cpu().numpy()
    scale = np.maximum(np.abs(cur_l), nsrc * np.abs(np.log(sigma)))
    assert np.all(np.abs(got_ll - cur_l) <= 1e-11 * scale)
    assert cur_k.min() >= kmin and cur_k.max() <= kmax and len(np.unique(cur_k)) > 3


def test_sd_step_matches_oracle():
    """rtb200_sd_step_device (the data-error move, :545-575) against the oracle, ten moves in a row."""
    import torch
    B, ldk, nsrc = 2000, 10, 24
    k, voro, so, sd, tobs, sigma, ll = _setup(B, ldk, nsrc, 61)
    rng = np.random.default_rng(62)
    sp = chains.sd_prior_array()
    beta = 1.0 / 1.4 ** rng.integers(0, 6, B)
    tk, tv, tl, tb, tg, ts, td, to = _dev(k, voro, ll, beta, sigma, so, sd, tobs)
    cur_l, cur_s = ll, sigma
    seen = {1: 0, 0: 0, -1: 0, 2: 0}
    for step in range(10):
        u = rng.random((2, B))
        gauss = rng.standard_normal(B)
        r = oracle.sd_step_batch(k, voro, cur_l, cur_s, u[0], gauss, u[1], beta, sp, so, sd, tobs)
        tu0, tga, tu1 = _dev(u[0], gauss, u[1])
        acc = chains.sd_step_device(tk, tv, tl, tg, tu0, tga, tu1, tb, sp, ts, td, to).cpu().numpy()
        assert np.array_equal(acc, r["accept"]), f"step {step}"
        assert np.array_equal(tg.cpu().numpy().view(np.uint64), r["sigma"].view(np.uint64))
        cur_l, cur_s = r["logL"], r["sigma"]
        for c in seen:
            seen[c] += int((acc == c).sum())
    assert all(n > 50 for n in seen.values()), seen
    got = tl.cpu().numpy()
    assert np.all(np.abs(got - cur_l) <= 1e-11 * np.maximum(np.abs(cur_l), nsrc * np.abs(np.log(cur_s))))
    assert np.array_equal(tv.cpu().numpy(), voro)                # the model is untouched


def test_mcmc_step_device_samples_a_sane_posterior():
    """The whole worker-loop iteration (birth/death + sweep + sigma move) on chains started at the
    test_1 model with its noisy data: chains stay inside the prior, the stored logL is the
    likelihood of the stored state, sigma drifts towards the noise level, k spreads out."""
    import json, os, torch
    c = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")))["config1"]
    so, sd = np.array(c["src_offset_full"]), np.array(c["src_depth_full"])
    v, z = np.array(c["vels"]), np.array(c["depths"])
    t, _, _ = oracle.trace_rays(v, z, so, sd)
    rng = np.random.default_rng(71)
    tobs = t + rng.normal(0, 0.016, len(so))
    B, ldk, k0 = 256, 10, len(v)
    voro = np.zeros((B, 2, ldk))
    voro[:, 0, 1:k0] = z
    voro[:, 1, :k0] = v
    k = np.full(B, k0, dtype=np.int32)
    sigma = np.full(B, 0.05)
    ll = np.full(B, oracle.loglhood_rt(v, z, so, sd, tobs, 0.05)[0])
    beta = np.ones(B)
    tk, tv, tl, tg, tb, ts, td, to = _dev(k, voro, ll, sigma, beta, so, sd, tobs)
    prior, sp, pk = chains.prior_array(), chains.sd_prior_array(), chains.poisson_pk(3.01, 1, ldk)
    gen = torch.Generator(device="cuda").manual_seed(5)
    acc = prop = 0
    for it in range(30):
        r = chains.mcmc_step_device(tk, tv, tl, tg, tb, prior, sp, pk, 1, ldk, ts, td, to, generator=gen)
        acc, prop = acc + int(r["accepted"].sum()), prop + int(r["proposed"].sum())
    kk, vv, sg, got = tk.cpu().numpy(), tv.cpu().numpy(), tg.cpu().numpy(), tl.cpu().numpy()
    assert 0.01 < acc / prop < 0.9
    assert kk.min() >= 1 and kk.max() <= ldk and len(np.unique(kk)) >= 2
    assert np.all((sg >= 0.001) & (sg <= 0.07)) and np.median(sg) < 0.05
    for b in range(0, B, 9):
        n = int(kk[b])
        zz = vv[b, 0, :n]
        assert zz[0] == 0.0 and np.all(np.diff(zz) >= 100.1) and np.all(vv[b, :, n:] == 0)
        assert np.all((vv[b, 1, :n] >= 1500.0) & (vv[b, 1, :n] <= 10000.0))
        ref = oracle.loglhood_rt(vv[b, 1, :n], zz[1:], so, sd, tobs, sg[b])[0]
        assert abs(got[b] - ref) <= 1e-11 * max(abs(ref), 20 * abs(np.log(sg[b])))
    assert np.median(got) > ll[0]                                 # the chains moved to better-fitting states


def test_chain_moves_survive_hostile_inputs():
    """Garbage in some chains (NaN / Inf deviates, node counts and indices out of range, NaN
    states) must neither hang nor disturb the other chains, which still match the oracle."""
    import torch
    B, ldk, nsrc = 1024, 8, 16
    k, voro, so, sd, tobs, sigma, ll = _setup(B, ldk, nsrc, 81)
    rng = np.random.default_rng(82)
    ivo = rng.integers(1, k + 1).astype(np.int32)
    iwhich = rng.integers(1, 3, B).astype(np.int32)
    u = rng.random((2, B))
    cauchy = 0.1 * np.tan(np.pi * (u[0] - 0.5))
    beta = np.ones(B)
    prior = chains.prior_array()
    want = oracle.mh_step_batch(k, voro, ll, ivo, iwhich, cauchy, u[1], beta, sigma, prior, so, sd, tobs)
    bad = np.arange(B) % 4 == 0
    kb, vb, ib, wb, cb, ub, lb = k.copy(), voro.copy(), ivo.copy(), iwhich.copy(), cauchy.copy(), u[1].copy(), ll.copy()
    sel = np.flatnonzero(bad)
    kb[sel[0::6]] = 0
    kb[sel[1::6]] = ldk + 5
    ib[sel[2::6]] = -3
    wb[sel[2::6]] = 7
    cb[sel[3::6]] = np.inf
    cb[sel[4::6]] = np.nan
    vb[sel[5::6], 0, 1] = np.nan
    ub[sel[3::6]] = np.nan
    tk, tv, tl, ti, tw, tc, tu, tb, tg, ts, td, to = _dev(kb, vb, lb, ib, wb, cb, ub, beta, sigma, so, sd, tobs)
    acc = chains.mh_step_device(tk, tv, tl, ti, tw, tc, tu, tb, tg, prior, ts, td, to)
    torch.cuda.synchronize()
    acc = acc.cpu().numpy()
    assert set(np.unique(acc)) <= {-1, 0, 1}
    good = ~bad
    assert np.array_equal(acc[good], want["accept"][good])
    assert np.array_equal(tv.cpu().numpy()[good].view(np.uint64), want["voro"][good].view(np.uint64))
    # the birth/death and sigma moves on the same garbage: they must come back
    pk = chains.poisson_pk(3.01, 1, ldk)
    uu = _dev(rng.random(B), rng.random(B), rng.random(B), ub)
    (idel,) = _dev(rng.integers(-2, ldk + 3, B).astype(np.int32))
    a2 = chains.bd_step_device(tk, tv, tl, uu[0], idel, uu[1], uu[2], uu[3], tb, tg, prior, pk, 1, ldk, ts, td, to)
    a3 = chains.sd_step_device(tk, tv, tl, tg, uu[0], tc, uu[3], tb, chains.sd_prior_array(), ts, td, to)
    torch.cuda.synchronize()
    assert set(np.unique(a2.cpu().numpy())) <= {-1, 0, 1, 2} and set(np.unique(a3.cpu().numpy())) <= {-1, 0, 1, 2}
    kk = tk.cpu().numpy()
    assert np.all((kk[good] >= 1) & (kk[good] <= ldk))


def test_mh_moves_graph_replay_and_recapture():
    """rtb200_mh_moves_device replays its captured graph when called again with the same buffers
    and captures anew when they change; either way the chains follow the oracle."""
    import math
    import torch
    prior = chains.prior_array()
    prior[:2] /= 10.0
    for B, n_moves, calls in ((200, 5, 3), (333, 4, 2), (200, 5, 1)):
        ldk, nsrc = 6, 12
        k, voro, so, sd, tobs, sigma, ll = _setup(B, ldk, nsrc, 90 + B)
        beta = np.ones(B)
        tk, tv, tl, tb, tg, ts, td, to = _dev(k, voro, ll, beta, sigma, so, sd, tobs)
        pos = torch.zeros(B, dtype=torch.int32, device="cuda")
        cur_v, cur_l, p = voro, ll, np.zeros(B, dtype=np.int64)
        for c in range(calls):
            gen = torch.Generator(device="cuda").manual_seed(1000 + c)
            chains.mh_moves_device(tk, tv, tl, pos, n_moves, tb, tg, prior, ts, td, to, generator=gen)
            gen = torch.Generator(device="cuda").manual_seed(1000 + c)
            u = torch.empty((2, n_moves, B), dtype=torch.float64, device="cuda").uniform_(generator=gen)
            cauchy, uacc = torch.tan(math.pi * (u[0] - 0.5)).cpu().numpy(), u[1].cpu().numpy()
            for m in range(n_moves):
                j = (p + m) % (2 * k - 1) + 1
                r = oracle.mh_step_batch(k, cur_v, cur_l, (j // 2 + 1).astype(np.int32),
                                         (j % 2 + 1).astype(np.int32), cauchy[m], uacc[m], beta, sigma,
                                         prior, so, sd, tobs)
                cur_v, cur_l = r["voro"], r["logL"]
            p = (p + n_moves) % (2 * k - 1)
            assert np.array_equal(tv.cpu().numpy().view(np.uint64), cur_v.view(np.uint64)), (B, c)
            assert np.array_equal(pos.cpu().numpy(), p)


def test_ar_step_matches_oracle():
    """rtb200_ar_step_device (the AR(1) move, :583-631, IAR = 1) against the oracle, ten moves in a row."""
    import torch
    B, ldk, nsrc = 2000, 10, 24
    k, voro, so, sd, tobs, sigma, _ = _setup(B, ldk, nsrc, 71)
    rng = np.random.default_rng(72)
    ap = chains.ar_prior_array()
    ap[3] = 5.0                                               # a roomy armxRT: most proposals are evaluated
    beta = 1.0 / 1.4 ** rng.integers(0, 6, B)
    idxar = rng.integers(0, 2, B).astype(np.int32)
    arpar = np.where(idxar == 1, rng.uniform(-0.5, 0.9, B), -1.5)
    ll = np.empty(B)
    for b in range(B):
        pred = oracle.loglhood_rt(voro[b, 1, :k[b]], voro[b, 0, 1:k[b]], so, sd, tobs, sigma[b])[1]
        ll[b] = oracle.loglhood_from_times_ar(pred, tobs, sigma[b], int(idxar[b]), float(arpar[b]), 5.0)
    tk, tv, tl, tg, ti, ta, tb, ts, td, to = _dev(k, voro, ll, sigma, idxar, arpar, beta, so, sd, tobs)
    cur_l, cur_i, cur_a = ll, idxar, arpar
    seen = {1: 0, 0: 0, -1: 0}
    for step in range(10):
        u = rng.random((3, B))
        gauss = rng.standard_normal(B)
        r = oracle.ar_step_batch(k, voro, cur_l, sigma, cur_i, cur_a, u[0], u[1], gauss, u[2], beta, ap, so, sd, tobs)
        tu0, tu1, tga, tu2 = _dev(u[0], u[1], gauss, u[2])
        acc = chains.ar_step_device(tk, tv, tl, tg, ti, ta, tu0, tu1, tga, tu2, tb, ap, ts, td, to).cpu().numpy()
        assert np.array_equal(acc, r["accept"]), f"step {step}"
        assert np.array_equal(ti.cpu().numpy(), r["idxar"])
        assert np.array_equal(ta.cpu().numpy().view(np.uint64), r["arpar"].view(np.uint64))
        cur_l, cur_i, cur_a = r["logL"], r["idxar"], r["arpar"]
        for c in seen:
            seen[c] += int((acc == c).sum())
    assert all(n > 50 for n in seen.values()), seen
    got = tl.cpu().numpy()
    fin = np.isfinite(cur_l) & (np.abs(cur_l) < 1e300)
    assert np.all(np.abs(got[fin] - cur_l[fin]) <= 1e-11 * np.maximum(np.abs(cur_l[fin]), nsrc * np.abs(np.log(sigma[fin]))))
    assert np.array_equal(got[~fin], cur_l[~fin])


def test_all_moves_with_the_ar_likelihood():
    """IAR = 1: with the chains' AR(1) state registered, every move (fixed-k, birth/death, sigma,
    AR) evaluates the AR likelihood; six rounds of all four against the oracle."""
    import torch
    B, ldk, nsrc = 1500, 8, 20
    k, voro, so, sd, tobs, sigma, _ = _setup(B, ldk, nsrc, 101)
    rng = np.random.default_rng(102)
    prior, sp, pk = chains.prior_array(), chains.sd_prior_array(), chains.poisson_pk(3.01, 1, ldk)
    prior[:2] /= 10.0
    ap = chains.ar_prior_array()
    ap[3] = 5.0
    beta = np.ones(B)
    idxar = rng.integers(0, 2, B).astype(np.int32)
    arpar = np.where(idxar == 1, rng.uniform(-0.5, 0.9, B), -1.5)
    ll = np.empty(B)
    for b in range(B):
        pred = oracle.loglhood_rt(voro[b, 1, :k[b]], voro[b, 0, 1:k[b]], so, sd, tobs, sigma[b])[1]
        ll[b] = oracle.loglhood_from_times_ar(pred, tobs, sigma[b], int(idxar[b]), float(arpar[b]), 5.0)
    tk, tv, tl, tg, ti, ta, tb, ts, td, to = _dev(k, voro, ll, sigma, idxar, arpar, beta, so, sd, tobs)
    chains.set_chain_ar(ti, ta, 5.0)
    ck, cv, cl, cs, ci, ca = k.copy(), voro, ll, sigma, idxar, arpar
    try:
        for rnd in range(6):
            oracle.set_chain_ar(ci, ca, 5.0)
            u = rng.random((4, B))
            idel = (2 + np.floor(rng.random(B) * np.maximum(ck - 1, 1))).astype(np.int32)
            r = oracle.bd_step_batch(ck, cv, cl, u[0], idel, u[1], u[2], u[3], beta, cs, prior, pk, 1, ldk, so, sd, tobs)
            t = _dev(u[0], idel, u[1], u[2], u[3])
            acc = chains.bd_step_device(tk, tv, tl, t[0], t[1], t[2], t[3], t[4], tb, tg, prior, pk, 1, ldk, ts, td, to)
            assert np.array_equal(acc.cpu().numpy(), r["accept"]), f"bd {rnd}"
            ck, cv, cl = r["k"], r["voro"], r["logL"]
            ivo = np.minimum(1 + rnd % 3, ck).astype(np.int32)
            iwhich = np.full(B, 2, dtype=np.int32)
            uu = rng.random((2, B))
            cauchy = np.tan(np.pi * (uu[0] - 0.5))
            r = oracle.mh_step_batch(ck, cv, cl, ivo, iwhich, cauchy, uu[1], beta, cs, prior, so, sd, tobs)
            t = _dev(ivo, iwhich, cauchy, uu[1])
            acc = chains.mh_step_device(tk, tv, tl, t[0], t[1], t[2], t[3], tb, tg, prior, ts, td, to)
            assert np.array_equal(acc.cpu().numpy(), r["accept"]), f"mh {rnd}"
            cv, cl = r["voro"], r["logL"]
            us, gs = rng.random((2, B)), rng.standard_normal(B)
            r = oracle.sd_step_batch(ck, cv, cl, cs, us[0], gs, us[1], beta, sp, so, sd, tobs)
            t = _dev(us[0], gs, us[1])
            acc = chains.sd_step_device(tk, tv, tl, tg, t[0], t[1], t[2], tb, sp, ts, td, to)
            assert np.array_equal(acc.cpu().numpy(), r["accept"]), f"sd {rnd}"
            cl, cs = r["logL"], r["sigma"]
            ua, ga = rng.random((3, B)), rng.standard_normal(B)
            r = oracle.ar_step_batch(ck, cv, cl, cs, ci, ca, ua[0], ua[1], ga, ua[2], beta, ap, so, sd, tobs)
            t = _dev(ua[0], ua[1], ga, ua[2])
            acc = chains.ar_step_device(tk, tv, tl, tg, ti, ta, t[0], t[1], t[2], t[3], tb, ap, ts, td, to)
            assert np.array_equal(acc.cpu().numpy(), r["accept"]), f"ar {rnd}"
            cl, ci, ca = r["logL"], r["idxar"], r["arpar"]
        assert np.array_equal(tk.cpu().numpy(), ck)
        assert np.array_equal(tv.cpu().numpy().view(np.uint64), cv.view(np.uint64))
        assert np.array_equal(tg.cpu().numpy().view(np.uint64), cs.view(np.uint64))
        assert np.array_equal(ti.cpu().numpy(), ci)
        assert np.array_equal(ta.cpu().numpy().view(np.uint64), ca.view(np.uint64))
    finally:
        chains.set_chain_ar()
        oracle.set_chain_ar()


def test_mcmc_step_device_with_ar():
    """IAR = 1 through mcmc_step_device: after a few iterations the stored logL is the AR(1)
    likelihood of the stored (model, sigma, idxar, arpar), and AR parameters stay inside the prior."""
    import torch
    B, ldk, nsrc = 300, 8, 20
    k, voro, so, sd, tobs, sigma, _ = _setup(B, ldk, nsrc, 111)
    rng = np.random.default_rng(112)
    idxar = np.zeros(B, dtype=np.int32)
    arpar = np.full(B, -1.5)
    ll = np.array([oracle.loglhood_rt(voro[b, 1, :k[b]], voro[b, 0, 1:k[b]], so, sd, tobs, sigma[b])[0]
                   for b in range(B)])
    ap = chains.ar_prior_array()
    tk, tv, tl, tg, ti, ta, tb, ts, td, to = _dev(k, voro, ll, sigma, idxar, arpar, np.ones(B), so, sd, tobs)
    prior, sp, pk = chains.prior_array(), chains.sd_prior_array(), chains.poisson_pk(3.01, 1, ldk)
    gen = torch.Generator(device="cuda").manual_seed(3)
    seen = set()
    for it in range(6):
        r = chains.mcmc_step_device(tk, tv, tl, tg, tb, prior, sp, pk, 1, ldk, ts, td, to, generator=gen,
                                    ar=(ti, ta, ap))
        seen |= set(np.unique(r["ar"].cpu().numpy()).tolist())
    assert {0, 1} <= seen <= {-1, 0, 1}
    kk, vv, sg, ii, aa, got = (t.cpu().numpy() for t in (tk, tv, tg, ti, ta, tl))
    assert ii.max() == 1 and np.all((aa[ii == 1] >= -0.5) & (aa[ii == 1] <= 0.9)) and np.all(aa[ii == 0] == -1.5)
    for b in range(0, B, 7):
        n = int(kk[b])
        pred = oracle.loglhood_rt(vv[b, 1, :n], vv[b, 0, 1:n], so, sd, tobs, sg[b])[1]
        ref = oracle.loglhood_from_times_ar(pred, tobs, sg[b], int(ii[b]), float(aa[b]), 0.5)
        if abs(ref) < 1e300:
            assert abs(got[b] - ref) <= 1e-11 * max(abs(ref), nsrc * abs(np.log(sg[b])))
        else:
            assert got[b] == ref


def test_enos_moves_match_oracle():
    """ENOS = 1 (order-statistics prior): fixed-dimension and birth/death moves against the oracle
    over a trajectory -- same decisions, same node counts, bit-identical chain states."""
    import torch
    B, ldk, nsrc = 2000, 10, 20
    kmin, kmax = 1, ldk
    k, voro, so, sd, tobs, sigma, ll = _setup(B, ldk, nsrc, 71)
    rng = np.random.default_rng(72)
    prior, pk = chains.prior_array(), chains.poisson_pk(3.01, kmin, kmax)
    beta = 1.0 / 1.4 ** rng.integers(0, 6, B)
    tk, tv, tl, tb, tg, ts, td, to = _dev(k, voro, ll, beta, sigma, so, sd, tobs)
    cur_k, cur_v, cur_l = k.copy(), voro, ll
    seen_mh = {1: 0, 0: 0, -1: 0}
    seen_bd = {1: 0, 0: 0, -1: 0, 2: 0}
    oracle.set_enos(1)
    try:
        for step in range(10):
            # a depth or velocity move of a random existing node
            ivo = (1 + np.floor(rng.random(B) * cur_k)).astype(np.int32)
            iwhich = rng.integers(1, 3, B).astype(np.int32)
            uu = rng.random((2, B))
            dev = np.where(iwhich == 1, uu[0], 0.05 * np.tan(np.pi * (uu[0] - 0.5)))
            r = oracle.mh_step_batch(cur_k, cur_v, cur_l, ivo, iwhich, dev, uu[1], beta, sigma, prior, so, sd, tobs)
            ti, tw, tc, tu = _dev(ivo, iwhich, dev, uu[1])
            acc = chains.mh_step_device(tk, tv, tl, ti, tw, tc, tu, tb, tg, prior, ts, td, to, enos=True)
            acc = acc.cpu().numpy()
            assert np.array_equal(acc, r["accept"]), f"mh step {step}"
            assert np.array_equal(tv.cpu().numpy().view(np.uint64), r["voro"].view(np.uint64)), f"mh step {step}"
            cur_v, cur_l = r["voro"], r["logL"]
            for c in seen_mh:
                seen_mh[c] += int((acc == c).sum())
            # birth / death
            u = rng.random((4, B))
            idel = (2 + np.floor(rng.random(B) * np.maximum(cur_k - 1, 1))).astype(np.int32)
            r = oracle.bd_step_batch(cur_k, cur_v, cur_l, u[0], idel, u[1], u[2], u[3], beta, sigma, prior,
                                     pk if step % 2 == 0 else None, kmin, kmax, so, sd, tobs)
            t_u = _dev(u[0], u[1], u[2], u[3])
            (t_idel,) = _dev(idel)
            acc = chains.bd_step_device(tk, tv, tl, t_u[0], t_idel, t_u[1], t_u[2], t_u[3], tb, tg, prior,
                                        pk if step % 2 == 0 else None, kmin, kmax, ts, td, to, enos=True)
            acc = acc.cpu().numpy()
            assert np.array_equal(acc, r["accept"]), f"bd step {step}"
            assert np.array_equal(tk.cpu().numpy(), r["k"])
            assert np.array_equal(tv.cpu().numpy().view(np.uint64), r["voro"].view(np.uint64)), f"bd step {step}"
            cur_k, cur_v, cur_l = r["k"], r["voro"], r["logL"]
            for c in seen_bd:
                seen_bd[c] += int((acc == c).sum())
    finally:
        oracle.set_enos(0)
    assert all(n > 50 for n in seen_mh.values()), seen_mh
    assert all(n > 50 for n in seen_bd.values()), seen_bd
    # and the ENOS depth move differs from the Cauchy one on the same deviates
    ivo = np.minimum(2, cur_k).astype(np.int32)
    iwhich = np.ones(B, dtype=np.int32)
    uu = rng.random((2, B))
    a = oracle.mh_step_batch(cur_k, cur_v, cur_l, ivo, iwhich, uu[0], uu[1], beta, sigma, prior, so, sd, tobs)
    ti, tw, tc, tu = _dev(ivo, iwhich, uu[0], uu[1])
    acc = chains.mh_step_device(tk, tv, tl, ti, tw, tc, tu, tb, tg, prior, ts, td, to, enos=False)
    assert np.array_equal(acc.cpu().numpy(), a["accept"])


@pytest.mark.parametrize("enos", [False, True])
def test_mcmc_graph_iteration_matches_oracle(enos):
    """rtb200_mcmc_iterations_device: a whole iteration (birth/death, every chain's next moves of
    its own sweep, data-error move) as one CUDA graph with the deviates drawn on the device.  The
    deviates are read back from the workspace and the oracle replays the iteration with them:
    same outcomes of every move, same node counts, bit-identical states, iteration after iteration."""
    import torch
    B, ldk, nsrc, M = 1500, 10, 20, 7
    kmin, kmax = 1, ldk
    k, voro, so, sd, tobs, sigma, ll = _setup(B, ldk, nsrc, 91)
    rng = np.random.default_rng(92)
    prior, sd_prior, pk = chains.prior_array(), chains.sd_prior_array(), chains.poisson_pk(3.01, kmin, kmax)
    beta = 1.0 / 1.4 ** rng.integers(0, 6, B)
    tk, tv, tl, tg, tb, ts, td, to = _dev(k, voro, ll, sigma, beta, so, sd, tobs)
    g = chains.McmcGraph(tk, tv, tl, tg, tb, M, prior, sd_prior, pk, kmin, kmax, ts, td, to, seed=5, enos=enos)
    cur_k, cur_v, cur_l, cur_s = k.copy(), voro, ll, sigma.copy()
    pos = np.zeros(B, dtype=np.int64)
    first_u = None
    oracle.set_enos(1 if enos else 0)
    try:
        for it in range(3):
            g.run(1)
            torch.cuda.synchronize()
            w = {n: t.cpu().numpy().copy() for n, t in g.views.items()}
            for n in ("u_k", "u_z", "u_v", "u_acc_bd", "u_gate", "u_acc_sd", "u_acc"):
                assert ((w[n] >= 0) & (w[n] < 1)).all(), n
            assert abs(w["gauss"].mean()) < 0.1 and 0.9 < w["gauss"].std() < 1.1
            if first_u is None:
                first_u = w["u_k"].copy()
            else:
                assert not np.array_equal(first_u, w["u_k"])         # the counter moved the stream on
            assert ((w["idel"] >= 2) & (w["idel"] <= np.maximum(cur_k, 2))).all()
            r = oracle.bd_step_batch(cur_k, cur_v, cur_l, w["u_k"], w["idel"], w["u_z"], w["u_v"], w["u_acc_bd"],
                                     beta, cur_s, prior, pk, kmin, kmax, so, sd, tobs)
            assert np.array_equal(w["acc_bd"], r["accept"]), f"bd, iteration {it}"
            cur_k, cur_v, cur_l = r["k"], r["voro"], r["logL"]
            period = np.maximum(2 * cur_k.astype(np.int64) - 1, 1)
            for m in range(M):
                j = (pos + m) % period + 1
                assert np.array_equal(w["ivo"][m], j // 2 + 1) and np.array_equal(w["iwhich"][m], j % 2 + 1)
                r = oracle.mh_step_batch(cur_k, cur_v, cur_l, w["ivo"][m], w["iwhich"][m], w["dev"][m], w["u_acc"][m],
                                         beta, cur_s, prior, so, sd, tobs)
                assert np.array_equal(w["acc_mh"][m], r["accept"]), f"move {m}, iteration {it}"
                cur_v, cur_l = r["voro"], r["logL"]
            pos = (pos + M) % period
            r = oracle.sd_step_batch(cur_k, cur_v, cur_l, cur_s, w["u_gate"], w["gauss"], w["u_acc_sd"], beta,
                                     sd_prior, so, sd, tobs)
            assert np.array_equal(w["acc_sd"], r["accept"]), f"sd, iteration {it}"
            cur_l, cur_s = r["logL"], r["sigma"]
            assert np.array_equal(tk.cpu().numpy(), cur_k)
            assert np.array_equal(tv.cpu().numpy().view(np.uint64), cur_v.view(np.uint64)), f"iteration {it}"
            assert np.array_equal(tg.cpu().numpy().view(np.uint64), cur_s.view(np.uint64))
            assert np.array_equal(g.pos.cpu().numpy(), pos)
    finally:
        oracle.set_enos(0)
    assert int(g.counter.item()) == 3
    tally = g.tally.cpu().numpy()
    assert tally[1].sum() > 0 and 0 < tally[0].sum() < tally[1].sum() and tally[2].sum() > 0 and tally[3].sum() > 0
    # many iterations back to back, no host work in between
    g.run(20)
    torch.cuda.synchronize()
    assert int(g.counter.item()) == 23
    kk = tk.cpu().numpy()
    assert kk.min() >= kmin and kk.max() <= kmax and np.isfinite(tl.cpu().numpy()).all()


def test_prior_sampling_mode_matches_oracle():
    """ISMPPRIOR = 1: every proposal's likelihood is LOGLHOOD2's constant (loglhood.f90:704-716), so
    inside-the-bounds proposals are accepted by the prior ratio alone; same decisions and states as
    the oracle, through the single moves and the whole-iteration graph."""
    import torch
    import raytracerfortran_b200 as rt
    B, ldk, nsrc, M = 1200, 10, 16, 5
    k, voro, so, sd, tobs, sigma, _ = _setup(B, ldk, nsrc, 101)
    ll = np.ones(B)                                     # LOGLHOOD2 of the starting states (:256-257)
    rng = np.random.default_rng(102)
    prior, sd_prior, pk = chains.prior_array(), chains.sd_prior_array(), chains.poisson_pk(3.01, 1, ldk)
    beta = 1.0 / 1.4 ** rng.integers(0, 6, B)
    tk, tv, tl, tg, tb, ts, td, to = _dev(k, voro, ll, sigma, beta, so, sd, tobs)
    rt.set_option("ismpprior", 1)
    oracle.set_ismpprior(1)
    try:
        ivo = np.minimum(2, k).astype(np.int32)
        iwhich = rng.integers(1, 3, B).astype(np.int32)
        u = rng.random((2, B))
        cauchy = 0.3 * np.tan(np.pi * (u[0] - 0.5))
        r = oracle.mh_step_batch(k, voro, ll, ivo, iwhich, cauchy, u[1], beta, sigma, prior, so, sd, tobs)
        ti, tw, tc, tu = _dev(ivo, iwhich, cauchy, u[1])
        acc = chains.mh_step_device(tk, tv, tl, ti, tw, tc, tu, tb, tg, prior, ts, td, to).cpu().numpy()
        assert np.array_equal(acc, r["accept"]) and set(np.unique(acc)) <= {1, -1}      # never rejected inside
        assert np.array_equal(tv.cpu().numpy().view(np.uint64), r["voro"].view(np.uint64))
        assert np.all(tl.cpu().numpy() == 1.0)
        cur_k, cur_v, cur_l, cur_s = k.copy(), r["voro"], r["logL"], sigma.copy()
        g = chains.McmcGraph(tk, tv, tl, tg, tb, M, prior, sd_prior, pk, 1, ldk, ts, td, to, seed=9)
        pos = np.zeros(B, dtype=np.int64)
        for it in range(2):
            g.run(1)
            torch.cuda.synchronize()
            w = {n: t.cpu().numpy().copy() for n, t in g.views.items()}
            rr = oracle.bd_step_batch(cur_k, cur_v, cur_l, w["u_k"], w["idel"], w["u_z"], w["u_v"], w["u_acc_bd"],
                                      beta, cur_s, prior, pk, 1, ldk, so, sd, tobs)
            assert np.array_equal(w["acc_bd"], rr["accept"])
            cur_k, cur_v, cur_l = rr["k"], rr["voro"], rr["logL"]
            for m in range(M):
                rr = oracle.mh_step_batch(cur_k, cur_v, cur_l, w["ivo"][m], w["iwhich"][m], w["dev"][m], w["u_acc"][m],
                                          beta, cur_s, prior, so, sd, tobs)
                assert np.array_equal(w["acc_mh"][m], rr["accept"])
                cur_v, cur_l = rr["voro"], rr["logL"]
            rr = oracle.sd_step_batch(cur_k, cur_v, cur_l, cur_s, w["u_gate"], w["gauss"], w["u_acc_sd"], beta,
                                      sd_prior, so, sd, tobs)
            assert np.array_equal(w["acc_sd"], rr["accept"])
            cur_l, cur_s = rr["logL"], rr["sigma"]
            assert np.array_equal(tv.cpu().numpy().view(np.uint64), cur_v.view(np.uint64))
            assert np.array_equal(tg.cpu().numpy().view(np.uint64), cur_s.view(np.uint64))
    finally:
        rt.set_option("ismpprior", 0)
        oracle.set_ismpprior(0)


def test_mcmc_graph_with_the_ar_move_matches_oracle():
    """IAR = 1 through the iteration graph: every likelihood uses the chains' AR(1) state and the
    iteration ends with EXPLORE_MH's AR move (:583-631); the oracle replays the device-drawn
    deviates: same outcomes, same idxarRT / arparRT / sigma / node bits."""
    import torch
    B, ldk, nsrc, M = 1200, 8, 20, 4
    k, voro, so, sd, tobs, sigma, _ = _setup(B, ldk, nsrc, 111)
    rng = np.random.default_rng(112)
    prior, sp, pk = chains.prior_array(), chains.sd_prior_array(), chains.poisson_pk(3.01, 1, ldk)
    prior[:2] /= 10.0
    ap = chains.ar_prior_array()
    ap[3] = 5.0
    beta = 1.0 / 1.4 ** rng.integers(0, 4, B)
    idxar = rng.integers(0, 2, B).astype(np.int32)
    arpar = np.where(idxar == 1, rng.uniform(-0.5, 0.9, B), -1.5)
    ll = np.empty(B)
    for b in range(B):
        pred = oracle.loglhood_rt(voro[b, 1, :k[b]], voro[b, 0, 1:k[b]], so, sd, tobs, sigma[b])[1]
        ll[b] = oracle.loglhood_from_times_ar(pred, tobs, sigma[b], int(idxar[b]), float(arpar[b]), 5.0)
    tk, tv, tl, tg, ti, ta, tb, ts, td, to = _dev(k, voro, ll, sigma, idxar, arpar, beta, so, sd, tobs)
    g = chains.McmcGraph(tk, tv, tl, tg, tb, M, prior, sp, pk, 1, ldk, ts, td, to, seed=21, ar=(ti, ta, ap))
    ck, cv, cl, cs, ci, ca = k.copy(), voro, ll, sigma.copy(), idxar, arpar
    pos = np.zeros(B, dtype=np.int64)
    seen = {1: 0, 0: 0, -1: 0}
    try:
        for it in range(3):
            g.run(1)
            torch.cuda.synchronize()
            w = {n: t.cpu().numpy().copy() for n, t in g.views.items()}
            oracle.set_chain_ar(ci, ca, 5.0)
            r = oracle.bd_step_batch(ck, cv, cl, w["u_k"], w["idel"], w["u_z"], w["u_v"], w["u_acc_bd"], beta, cs,
                                     prior, pk, 1, ldk, so, sd, tobs)
            assert np.array_equal(w["acc_bd"], r["accept"]), f"bd {it}"
            ck, cv, cl = r["k"], r["voro"], r["logL"]
            period = np.maximum(2 * ck.astype(np.int64) - 1, 1)
            for m in range(M):
                r = oracle.mh_step_batch(ck, cv, cl, w["ivo"][m], w["iwhich"][m], w["dev"][m], w["u_acc"][m], beta, cs,
                                         prior, so, sd, tobs)
                assert np.array_equal(w["acc_mh"][m], r["accept"]), f"move {m}, iteration {it}"
                cv, cl = r["voro"], r["logL"]
            pos = (pos + M) % period
            r = oracle.sd_step_batch(ck, cv, cl, cs, w["u_gate"], w["gauss"], w["u_acc_sd"], beta, sp, so, sd, tobs)
            assert np.array_equal(w["acc_sd"], r["accept"]), f"sd {it}"
            cl, cs = r["logL"], r["sigma"]
            r = oracle.ar_step_batch(ck, cv, cl, cs, ci, ca, w["u_choice"], w["u_prop_ar"], w["gauss_ar"], w["u_acc_ar"],
                                     beta, ap, so, sd, tobs)
            assert np.array_equal(w["acc_ar"], r["accept"]), f"ar {it}"
            cl, ci, ca = r["logL"], r["idxar"], r["arpar"]
            for c_ in seen:
                seen[c_] += int((w["acc_ar"] == c_).sum())
            assert np.array_equal(tk.cpu().numpy(), ck)
            assert np.array_equal(tv.cpu().numpy().view(np.uint64), cv.view(np.uint64)), f"iteration {it}"
            assert np.array_equal(tg.cpu().numpy().view(np.uint64), cs.view(np.uint64))
            assert np.array_equal(ti.cpu().numpy(), ci)
            assert np.array_equal(ta.cpu().numpy().view(np.uint64), ca.view(np.uint64))
    finally:
        oracle.set_chain_ar()
    assert all(n > 20 for n in seen.values()), seen
    assert int(g.tally[4].sum().item()) == seen[1]
