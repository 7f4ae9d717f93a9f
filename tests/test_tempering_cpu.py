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
        # the all-device variant: same rule, device-side pairing; ranks must agree
        beta_d = torch.from_numpy(beta_all[lo:hi].copy())
        dev_hist = []
        for rnd in range(5):
            before = beta_d.clone()
            beta_d, info = tempering.tempering_swap_round_device(logL, beta_d, seed, rnd)
            dev_hist.append((before.numpy().copy(), beta_d.numpy().copy(), info["pairs_i"].numpy().copy(),
                             info["pairs_j"].numpy().copy(), info["accept"].numpy().copy()))
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
        b0, n0, i0, j0, a0 = d0[rnd]
        b1, n1, i1, j1, a1 = d1[rnd]
        assert np.array_equal(i0, i1) and np.array_equal(j0, j1) and np.array_equal(a0, a1)
        before, after = np.concatenate([b0, b1]), np.concatenate([n0, n1])
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
