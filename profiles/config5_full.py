#!/usr/bin/env python
"""BASELINE.json config 5 at FULL size on one GPU: 16M models x 50 interfaces x 1024 near-critical
sources (1.7e10 rays), generated on the device, logL fused (no travel-time store: that would be
131 GB).  Checks: a sample of models re-evaluated alone gives the same bits (batch invariance) and
matches the CPU oracle bit for bit in travel time."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import oracle
    import raytracerfortran_b200 as rt
    from raytracerfortran_b200 import device, workloads

    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
    L, S = 50, 1024
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    v = 1500.0 + 8500.0 * torch.rand((B, L + 1), dtype=torch.float64, device=dev, generator=g)
    z = torch.empty((B, L), dtype=torch.float64, device=dev)
    step = 2_000_000
    for lo in range(0, B, step):           # sort in slices to bound the temporary memory
        u = torch.rand((min(step, B - lo), L), dtype=torch.float64, device=dev, generator=g)
        z[lo:lo + step] = 50.0 + 9950.0 * torch.sort(u, dim=1).values
    nl = torch.full((B,), L, dtype=torch.int32, device=dev)
    so, sd = workloads.make_sources(S, 5, near_critical=True)
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ts, td = f(so), f(sd)
    tobs, sigma = f(np.full(S, 2.0)), torch.full((B,), 0.02, dtype=torch.float64, device=dev)
    ll = torch.empty(B, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    device.dff_batch_device(v, z, nl, ts, td, tobs=tobs, sigma=sigma, logL=ll)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # batch invariance + oracle on a sample (travel times of 12 models spread over the batch)
    pick = torch.linspace(0, B - 1, 12, device=dev).long()
    out = device.dff_batch_device(v[pick].contiguous(), z[pick].contiguous(), nl[pick].contiguous(),
                                  ts, td, tobs=tobs, sigma=sigma[pick].contiguous(), want_times=True)
    torch.cuda.synchronize()
    same_ll = bool(torch.equal(out["logL"], ll[pick]))
    ref = oracle.dff_batch(v[pick].cpu().numpy(), z[pick].cpu().numpy(), nl[pick].cpu().numpy(), so, sd)
    same_t = bool(np.array_equal(out["timeP"].cpu().numpy().view(np.uint64), ref["timeP"].view(np.uint64)))
    print(json.dumps({"workload": f"config5 full: {B} models x {L} interfaces x {S} near-critical sources",
                      "rays": B * S, "seconds": ms / 1e3, "evals_per_s": B * S / (ms / 1e3),
                      "logL_all_minus_inf_like_reference": bool(torch.isneginf(ll).all()),
                      "batch_invariant_logL": same_ll, "sample_travel_times_bit_identical_to_oracle": same_t,
                      "variant": int(rt.get_stat("variant")), "tile_models": int(rt.get_stat("tile_models")),
                      "grid": int(rt.get_stat("grid"))}))


if __name__ == "__main__":
    main()
