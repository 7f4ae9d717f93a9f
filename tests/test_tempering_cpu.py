"""Multi-rank logic on CPU: model-axis sharding and the parallel-tempering swap all-gather,
world_size 2 over gloo (the GPU path uses the same code over NCCL)."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raytracerfortran_b200 import tempering


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_swap_rule_matches_tempswp_mh():
    """prjmh_temper_rf.f90:1341-1344: logratio = (beta2-beta1)*(logL1-logL2); accept if u <= exp(.)."""
    rng = np.random.default_rng(3)
    logL = rng.normal(-50, 30, 64)
    beta = tempering.temperature_ladder(64, 1.4)
    for rnd in range(20):
        pairs, acc = tempering.swap_decisions(logL, beta, seed=99, round_index=rnd)
        pairs2, u = tempering.swap_pairs(64, 99, rnd)
        assert np.array_equal(pairs, pairs2)
        assert sorted(pairs.ravel().tolist()) == list(range(64))       # a perfect matching
        for (i, j), a, uu in zip(pairs, acc, u):
            logratio = (beta[j] - beta[i]) * (logL[i] - logL[j])
            want = uu <= (math.exp(logratio) if logratio < 700 else math.inf)
            assert bool(a) == want
        new = tempering.apply_swaps(beta, pairs, acc)
        assert sorted(new.tolist()) == sorted(beta.tolist())           # betas are permuted, never lost
    # the hot chain (beta 0.1) holds the much better state: the swap that hands it to the cold
    # chain (beta 1.0) has logratio = +900 and is always accepted, whatever the pairing order
    _, a = tempering.swap_decisions([0.0, -1000.0], [0.1, 1.0], 1, 0)
    assert bool(a[0])
    # the reverse (cold chain already holds the better state): logratio = -900, never accepted
    _, a = tempering.swap_decisions([0.0, -1000.0], [1.0, 0.1], 1, 0)
    assert not bool(a[0])


def test_temperature_ladder():
    b = tempering.temperature_ladder(6, 1.4, n_cold=2)
    assert b[0] == b[1] == 1.0 and np.allclose(b[2:], 1.0 / 1.4 ** np.arange(1, 5))


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_local, seed = 4, 2024
        rng = np.random.default_rng(100)                    # same stream on both ranks
        logL_all = rng.normal(-40, 25, world * n_local)
        beta_all = tempering.temperature_ladder(world * n_local, 1.4)
        lo, hi = tempering.shard_range(world * n_local, rank, world)
        assert (lo, hi) == (rank * n_local, (rank + 1) * n_local)
        logL = torch.from_numpy(logL_all[lo:hi].copy())
        beta = torch.from_numpy(beta_all[lo:hi].copy())
        g_l, g_b = tempering.allgather_replicas(logL, beta)
        assert np.array_equal(g_l, logL_all) and np.array_equal(g_b, beta_all)
        history = []
        for rnd in range(5):
            beta, st = tempering.tempering_swap_round(logL, beta, seed, rnd)
            history.append((beta.numpy().copy(), st["accept"].copy(), st["pairs"].copy()))
        # the device-side scheme (rtb200_swap_round_device), restated by oracle/tempering_ref.py:
        # every rank applies it to what it gathered; ranks must agree on pairs, decisions, betas
        from oracle import tempering_ref
        beta_d = torch.from_numpy(beta_all[lo:hi].copy())
        dev_hist = []
        for rnd in range(5):
            g_l, g_b = tempering.allgather_replicas(logL, beta_d)
            new, pairs, acc, _ = tempering_ref.swap_round(g_l, g_b, seed, rnd)
            dev_hist.append((g_b.copy(), new.copy(), pairs[:, 0].copy(), pairs[:, 1].copy(), acc.copy()))
            beta_d = torch.from_numpy(new[lo:hi].copy())
        ret[rank] = history
        ret[("dev", rank)] = dev_hist
    finally:
        dist.destroy_process_group()


def test_swap_round_world_size_2_gloo():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()     # no fork() of a process that already runs OpenMP threads
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    h0, h1 = ret[0], ret[1]
    # serial restatement of the same five rounds
    rng = np.random.default_rng(100)
    logL_all = rng.normal(-40, 25, 8)
    beta_all = tempering.temperature_ladder(8, 1.4)
    for rnd in range(5):
        pairs, acc = tempering.swap_decisions(logL_all, beta_all, 2024, rnd)
        beta_all = tempering.apply_swaps(beta_all, pairs, acc)
        assert np.array_equal(h0[rnd][1], acc) and np.array_equal(h1[rnd][1], acc)     # same decisions
        assert np.array_equal(h0[rnd][2], pairs) and np.array_equal(h1[rnd][2], pairs)
        assert np.array_equal(np.concatenate([h0[rnd][0], h1[rnd][0]]), beta_all)      # same ladder
    assert sum(int(h0[r][1].sum()) for r in range(5)) > 0                              # something swapped
    # device variant: both ranks derived the same pairs and decisions, betas stay a permutation,
    # and the decisions follow TEMPSWP_MH's rule on the gathered state
    d0, d1 = ret[("dev", 0)], ret[("dev", 1)]
    ladder = tempering.temperature_ladder(8, 1.4)
    for rnd in range(5):
        before, after, i0, j0, a0 = d0[rnd]
        b1, n1, i1, j1, a1 = d1[rnd]
        assert np.array_equal(i0, i1) and np.array_equal(j0, j1) and np.array_equal(a0, a1)
        assert np.array_equal(before, b1) and np.array_equal(after, n1)
        assert sorted(after.tolist()) == sorted(ladder.tolist())
        for i, j, a in zip(i0, j0, a0):
            if a:
                assert after[i] == before[j] and after[j] == before[i]
            else:
                assert after[i] == before[i] and after[j] == before[j]


def test_mh_accept_rule_matches_explore_mh_novarpar():
    """prjmh_temper_rf.f90:739-757: reject iff ran_uni >= EXP(logPr + (logL_new - logL)*beta_mh),
    and always when the proposal left the prior bounds."""
    g = torch.Generator().manual_seed(5)
    n = 4096
    cur = torch.randn(n, generator=g, dtype=torch.float64) * 20 - 50
    new = cur + torch.randn(n, generator=g, dtype=torch.float64) * 3
    beta = torch.rand(n, generator=g, dtype=torch.float64)
    u = torch.rand(n, generator=g, dtype=torch.float64)
    lpr = torch.randn(n, generator=g, dtype=torch.float64) * 0.1
    out = torch.rand(n, generator=g) < 0.1
    acc = tempering.mh_accept(cur, new, beta, u, logPr_new=lpr, outside=out)
    for i in range(0, n, 7):
        want = (not bool(out[i])) and not (float(u[i]) >= math.exp(float(lpr[i]) + (float(new[i]) - float(cur[i])) * float(beta[i])))
        assert bool(acc[i]) == want
    assert 0.2 < acc.double().mean() < 0.9


def test_ladder_adapter_follows_the_burn_in_rule():
    """prjmh_temper_rf.f90:363-383, :1373-1379: rate latched after more than `window` proposals;
    dTlog * 1.02 below 0.2, * 0.98 above 0.5, untouched in between or after burn-in."""
    ad = tempering.LadderAdapter(6, 1.4, acceptance_window=10)
    assert ad.update(np.zeros(10, bool)) is None and ad.acceptance_rate == 0.25      # window not exceeded yet
    lad = ad.update(np.zeros(1, bool))                                               # 11th proposal: rate 0
    assert ad.acceptance_rate == 0.0 and abs(ad.dTlog - 1.4 * 1.02) < 1e-15
    assert np.allclose(lad, 1.0 / (1.4 * 1.02) ** np.arange(6)) and (ad.ncswap, ad.ncswapprop) == (0, 0)
    lad = ad.update(np.ones(11, bool))                                               # rate 1 -> shrink
    assert ad.acceptance_rate == 1.0 and abs(ad.dTlog - 1.4 * 1.02 * 0.98) < 1e-15 and lad is not None
    acc = np.zeros(11, bool)
    acc[:4] = True                                                                   # 4/11 = 0.36: keep
    d = ad.dTlog
    assert ad.update(acc) is None and ad.dTlog == d and abs(ad.acceptance_rate - 4 / 11) < 1e-15
    assert ad.update(np.zeros(11, bool), burn_in=False) is None and ad.dTlog == d    # after burn-in: only the rate
    assert ad.acceptance_rate == 0.0


def test_philox4x32_10_known_answers():
    """Random123's known-answer vectors for philox4x32-10 pin the generator both the oracle and the
    swap kernel implement."""
    from oracle import tempering_ref as t
    h = lambda x: [int(np.asarray(v).reshape(-1)[0]) for v in x]
    assert h(t.philox4x32_10((0, 0, 0, 0), (0, 0))) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert h(t.philox4x32_10((0xffffffff,) * 4, (0xffffffff, 0xffffffff))) == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert h(t.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_device_swap_scheme_reference():
    """The counter-derived pairing is a perfect matching for any n, changes with the round, and the
    decisions follow TEMPSWP_MH's rule (prjmh_temper_rf.f90:1339-1344)."""
    from oracle import tempering_ref as t
    rng = np.random.default_rng(8)
    for n in (2, 3, 5, 64, 1000, 4097):
        p = t.swap_perm(n, 11, 3)
        assert sorted(p.tolist()) == list(range(n))
        assert np.array_equal(p[:7], t.swap_perm(n, 11, 3, idx=np.arange(min(7, n))))
        logL = rng.normal(-50, 30, n)
        beta = tempering.temperature_ladder(n, 1.02)
        new, pairs, acc, u = t.swap_round(logL, beta, 11, 3)
        assert pairs.shape == (n // 2, 2) and len(set(pairs.ravel().tolist())) == 2 * (n // 2)
        assert sorted(new.tolist()) == sorted(beta.tolist())
        assert ((u >= 0) & (u < 1)).all()
        for (i, j), a, uu in zip(pairs[:50], acc[:50], u[:50]):
            lr = (beta[j] - beta[i]) * (logL[i] - logL[j])
            assert bool(a) == (uu <= (math.exp(lr) if lr < 700 else math.inf))
            assert (new[i], new[j]) == ((beta[j], beta[i]) if a else (beta[i], beta[j]))
    assert not np.array_equal(t.swap_perm(64, 11, 3), t.swap_perm(64, 11, 4))
    assert not np.array_equal(t.swap_perm(64, 11, 3), t.swap_perm(64, 12, 3))


def test_assign_ladder_keeps_the_temperature_order():
    """A new ladder goes to the chains by the rank of the beta they currently hold
    (prjmh_temper_rf.f90:373-383 reassigns by slot; here swaps moved the betas, not the states)."""
    cur = np.array([0.5, 1.0, 0.25, 0.7])
    lad = tempering.temperature_ladder(4, 1.5)                     # 1, 1/1.5, 1/2.25, 1/3.375
    out = tempering.assign_ladder(cur, lad)
    assert np.array_equal(out, np.array([lad[2], lad[0], lad[3], lad[1]]))
    out_t = tempering.assign_ladder(torch.from_numpy(cur), lad)
    assert np.array_equal(out_t.numpy(), out)
