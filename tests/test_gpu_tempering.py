"""The device-side parallel-tempering swap round (rtb200_swap_pack_device /
rtb200_swap_round_device) against its numpy restatement (oracle/tempering_ref.py), bit for bit:
pairs, uniforms' decisions, exchanged betas, for every slice a rank may own."""
import numpy as np
import pytest
import torch

from oracle import tempering_ref
from raytracerfortran_b200 import tempering

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [2, 3, 64, 1001, 65536])
def test_swap_round_kernel_matches_the_reference(n):
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(n)
    logL = rng.normal(-60, 40, n)
    beta = rng.permutation(tempering.temperature_ladder(n, 1.0 + 3.0 / n))
    tl, tb = torch.from_numpy(logL).to(dev), torch.from_numpy(beta).to(dev)
    for rnd in (0, 1, 12345678901):
        want, pairs, acc, _ = tempering_ref.swap_round(logL, beta, 2026, rnd)
        got, info = tempering.tempering_swap_round_device(tl, tb, 2026, rnd)
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy().view(np.uint64), want.view(np.uint64))
        assert np.array_equal(info["accept"].cpu().numpy()[:n // 2].astype(bool), acc)
        partner = info["partner"].cpu().numpy()
        for (i, j), a in zip(pairs[:200], acc[:200]):
            assert partner[i] == (j if a else -1 - j) and partner[j] == (i if a else -1 - i)
        assert 0 < acc.sum() or n < 8


def test_swap_round_kernel_slices_agree_with_the_whole():
    """A rank writes only its own chains; the slices of all ranks tile the full result."""
    from raytracerfortran_b200 import _lib
    from raytracerfortran_b200.device import _ensure_device
    dev = torch.device("cuda", 0)
    _ensure_device(0)
    n, world = 4096, 8
    rng = np.random.default_rng(1)
    logL = rng.normal(-60, 40, n)
    beta = rng.permutation(tempering.temperature_ladder(n, 1.001))
    want, _, _, _ = tempering_ref.swap_round(logL, beta, 7, 3)
    allr = torch.from_numpy(np.stack([logL, beta], axis=1).copy()).to(dev)
    lib = _lib.load()
    out = []
    for r in range(world):
        nl = n // world
        b = torch.full((nl,), -1.0, dtype=torch.float64, device=dev)
        _lib.check(lib.rtb200_swap_round_device(allr.data_ptr(), n, r * nl, nl, 7, 3, b.data_ptr(), None, None, None))
        out.append(b.cpu().numpy())
    assert np.array_equal(np.concatenate(out).view(np.uint64), want.view(np.uint64))
    # bad arguments are refused
    assert lib.rtb200_swap_round_device(allr.data_ptr(), n, n - 3, 8, 7, 3, b.data_ptr(), None, None, None) != 0


def test_swap_round_overlaps_with_other_work():
    """launch() returns at once and wait() orders the caller's stream behind the new betas."""
    dev = torch.device("cuda", 0)
    n = 8192
    rng = np.random.default_rng(3)
    logL = torch.from_numpy(rng.normal(-60, 40, n)).to(dev)
    beta = torch.from_numpy(rng.permutation(tempering.temperature_ladder(n, 1.001))).to(dev)
    ref, _, _, _ = tempering_ref.swap_round(logL.cpu().numpy(), beta.cpu().numpy(), 5, 9)
    sr = tempering.SwapRound(n, dev)
    sr.launch(logL, beta, 5, 9)                       # in place
    busy = torch.randn(1 << 20, device=dev).cumsum(0)  # unrelated work on the caller's stream
    sr.wait()
    got = beta.clone()
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy().view(np.uint64), ref.view(np.uint64))
    assert busy.numel() == 1 << 20


def test_overlapped_swap_rounds_equal_the_sequential_schedule():
    """The bench's config-4 step: MH move of every chain, then a swap round launched on a side
    stream under the NEXT step's proposal and likelihood kernels (only the accept test waits for the
    new betas).  The chain states after ten such steps equal, bit for bit, those of the same steps
    run one after the other with a synchronise in between."""
    from raytracerfortran_b200 import chains, device, workloads
    dev = torch.device("cuda", 0)
    B, ldk, nsrc, steps = 4096, 12, 32, 10
    k, vp, zi = workloads.make_transd_models(B, ldk, 77)
    voro = np.zeros((B, 2, ldk))
    voro[:, 1, :] = vp
    voro[:, 0, 1:] = zi
    so, sd = workloads.make_sources(nsrc, 77)
    tobs, sigma = workloads.make_observations(np.full(nsrc, 1.3), B, 77)
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ts, td, to, tg = f(so), f(sd), f(tobs), f(sigma)
    prior = chains.prior_array()
    gen = torch.Generator(device=dev).manual_seed(3)
    tk0 = f(k)
    period = (2 * tk0 - 1).to(torch.int64)
    j = torch.arange(steps, device=dev, dtype=torch.int64)[:, None] % period[None, :] + 1
    ivo = (torch.div(j, 2, rounding_mode="floor") + 1).to(torch.int32).contiguous()
    iwh = (j % 2 + 1).to(torch.int32).contiguous()
    u = torch.rand((2, steps, B), dtype=torch.float64, device=dev, generator=gen)
    cauchy, uacc = (0.2 * chains.cauchy_deviates(u[0])).contiguous(), u[1].contiguous()
    ladder = np.repeat(tempering.temperature_ladder(64, 1.3), B // 64)

    def run(overlap):
        tv, beta = f(voro), f(ladder)
        tl = device.dff_batch_device(tv[:, 1, :].contiguous(), tv[:, 0, 1:].contiguous(), tk0, ts, td,
                                     tobs=to, sigma=tg, kmode=True)["logL"]
        sr = tempering.SwapRound(B, dev)
        acc = torch.empty(B, dtype=torch.int32, device=dev)
        for i in range(steps):
            chains.mh_step_device(tk0, tv, tl, ivo[i], iwh[i], cauchy[i], uacc[i], beta, tg, prior, ts, td, to,
                                  accept=acc, beta_ready=sr.done if (overlap and i > 0) else None)
            sr.launch(tl, beta, 11, i)
            if not overlap:
                sr.wait()
                torch.cuda.synchronize()
        sr.wait()
        torch.cuda.synchronize()
        return tv.cpu().numpy(), tl.cpu().numpy(), beta.cpu().numpy()

    a, b = run(False), run(True)
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint64), y.view(np.uint64))
    assert not np.array_equal(a[2], ladder)              # swaps did happen
