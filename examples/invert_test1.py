#!/usr/bin/env python
"""The reference's test_1 inversion run the B200 way: a population of parallel-tempering chains on
one GPU, every MCMC iteration (birth/death, the sweep over the nodes, the data-error move, a swap
round) on the device, the T = 1 chains written to `<base>_voro_sample.txt` in the sampler's own
format and re-evaluated by the replica sweep.

    python examples/invert_test1.py [--chains 256] [--iters 300]

Needs a B200 (there is no CPU path).  Data: the test_1 model and sources (tests/golden), travel
times from the forward model plus N(0, 0.016^2) noise as in the reference's Rmd (:89-90)."""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=256)
    ap.add_argument("--temps", type=int, default=8)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--thin", type=int, default=5)           # ICHAINTHIN of test_1_parameter.dat
    args = ap.parse_args()
    import torch
    import raytracerfortran_b200 as rt
    from raytracerfortran_b200 import chains, samplefile, tempering

    c = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")))["config1"]
    so, sd = np.array(c["src_offset_full"]), np.array(c["src_depth_full"])
    v_true, z_true = np.array(c["vels"]), np.array(c["depths"])
    rng = np.random.default_rng(12)
    tobs = rt.dff(v_true, z_true, so, sd) + rng.normal(0.0, 0.016, len(so))

    NLMX, B = 10, args.chains
    dev = torch.device("cuda:0")
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    # every chain starts from a one-node half-space drawn from the prior
    k = np.ones(B, dtype=np.int32)
    voro = np.zeros((B, 2, NLMX))
    voro[:, 1, 0] = rng.uniform(1500.0, 10000.0, B)
    sigma = rng.uniform(0.001, 0.07, B)
    beta = np.tile(tempering.temperature_ladder(args.temps, 1.4), B // args.temps + 1)[:B]
    tk, tv, tg, tb, ts, td, to = f(k), f(voro), f(sigma), f(beta), f(so), f(sd), f(tobs)
    from raytracerfortran_b200 import device
    tl = device.dff_batch_device(tv[:, 1, :].contiguous(), tv[:, 0, 1:].contiguous(), tk, ts, td,
                                 tobs=to, sigma=tg, kmode=True)["logL"]
    prior, sp, pk = chains.prior_array(), chains.sd_prior_array(), chains.poisson_pk(3.01, 1, NLMX)
    # one CUDA graph per iteration (birth/death, 2*NLMX-1 moves of every chain's own sweep, sigma
    # move; deviates drawn on the device), the swap round on a side stream in between
    graph = chains.McmcGraph(tk, tv, tl, tg, tb, 2 * NLMX - 1, prior, sp, pk, 1, NLMX, ts, td, to, seed=12)
    swap = tempering.SwapRound(B, dev)
    rows = []
    for it in range(args.iters):
        graph.run(1)
        swap.launch(tl, tb, 12, it)                    # betas exchanged in place
        swap.wait()
        if it >= args.iters // 2 and it % args.thin == 0:        # keep the T = 1 chains after burn-in
            cold = (tb == 1.0).nonzero().squeeze(1)
            rows.append(samplefile.pack_rows(tl[cold].cpu().numpy(), np.zeros(len(cold)), np.zeros(len(cold)),
                                             tk[cold].cpu().numpy(), tv[cold].cpu().numpy(),
                                             tg[cold].cpu().numpy()))
    rows = np.concatenate(rows)
    path = os.path.join(tempfile.mkdtemp(), "test_1_voro_sample.txt")
    samplefile.write_samples(path, rows)
    smp, logL, pred = samplefile.replica_sweep(path, NLMX, so, sd, tobs)
    rms = np.sqrt(np.mean((pred - tobs) ** 2, axis=1))
    ll_true = rt.loglhood_batch([len(v_true)], v_true[None, :], z_true[None, :], so, sd, tobs, [0.016])[0][0]
    print(json.dumps({"chains": B, "temperatures": args.temps, "iterations": args.iters,
                      "kept_samples": int(len(logL)), "sample_file": path,
                      "logL_true_model": float(ll_true), "logL_median": float(np.median(logL)),
                      "rms_residual_median_s": float(np.median(rms)), "noise_sd_s": 0.016,
                      "k_mean": float(smp["k"].mean()), "sigma_median": float(np.median(smp["sdparRT"]))}))


if __name__ == "__main__":
    main()
