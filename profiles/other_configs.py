#!/usr/bin/env python
"""Device-resident throughput on the shapes of BASELINE.json configs 3, 4 and 5 (bench.py keeps
to config 2, the metric's configuration).  One JSON line per shape: evals/s, logL/s, W_min from
the instrumented oracle on a sample of the same inputs, FP64 roofline fraction.

    python profiles/other_configs.py [--steps K] [--warmup W] [--only NAME]

config 5 is measured on a slice of its 16M models (the full sweep is ~1.6e10 rays); the slice is
as deep (50 interfaces), as wide (1024 sources) and as near-critical as the full configuration.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()

    import torch
    import oracle
    import raytracerfortran_b200 as rt
    from raytracerfortran_b200 import device, workloads

    for kv in args.opt:
        name, val = kv.split("=")
        rt.set_option(name, float(val))
    dev = torch.device("cuda:0")
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    peak = rt.fp64_peak_tflops(3)

    shapes = {}
    c3 = workloads.CONFIGS["config3"]
    k, vp, zi = workloads.make_transd_models(c3["B"], c3["kmax"], c3["seed"])
    shapes["config3 (4096 states/step, k=1..30 Poisson(3.01), 256 sources)"] = dict(
        kmode=True, v=vp, z=zi, n=k, src=workloads.make_sources(c3["nsrc"], c3["seed"]))
    k, vp, zi = workloads.make_transd_models(c3["B"], c3["kmax"], c3["seed"] + 1, uniform_k=True)
    shapes["config3u (4096 states/step, k uniform 1..30, 256 sources)"] = dict(
        kmode=True, v=vp, z=zi, n=k, src=workloads.make_sources(c3["nsrc"], c3["seed"]))
    c4 = workloads.CONFIGS["config4"]
    k, vp, zi = workloads.make_transd_models(8 * c4["proposals"], c4["kmax"], c4["seed"])
    shapes["config4 per GPU (8 replicas x 1024 proposals, 256 sources)"] = dict(
        kmode=True, v=vp, z=zi, n=k, src=workloads.make_sources(c4["nsrc"], c4["seed"]))
    c5 = workloads.CONFIGS["config5"]
    v, z, nl = workloads.make_models(16384, c5["nlayers"], c5["seed"], min_thickness=False)
    shapes["config5 slice (16384 of 16M models, 50 interfaces, 1024 near-critical sources)"] = dict(
        kmode=False, v=v, z=z, n=nl, src=workloads.make_sources(c5["nsrc"], c5["seed"], near_critical=True))

    for name, s in shapes.items():
        if "config5" in name and args.opt and not args.only:
            continue
        if args.only and args.only not in name:
            continue
        so, sd = s["src"]
        B, S = len(s["v"]), len(so)
        tobs, sigma = workloads.make_observations(np.full(S, 1.5), B, 1)
        tv, tz, tn = f(s["v"]), f(s["z"]), f(s["n"].astype(np.int32))
        ts, td, to, tg = f(so), f(sd), f(tobs), f(sigma)
        ll = torch.empty(B, dtype=torch.float64, device=dev)
        step = lambda: device.dff_batch_device(tv, tz, tn, ts, td, tobs=to, sigma=tg, logL=ll,
                                               kmode=s["kmode"])
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        # work per evaluation from the oracle on a sample of the same models
        nb = min(B, max(8, 60000 // S))
        if s["kmode"]:
            kk = s["n"][:nb]
            nlay = np.where(kk > 1, kk - 1, 1).astype(np.int32)
            vv = s["v"][:nb].copy()
            zz = np.concatenate([s["z"][:nb], np.zeros((nb, 1))], axis=1)[:, :max(s["z"].shape[1], 1)].copy()
            one = kk <= 1
            vv[one, 1] = vv[one, 0]
            zz[one, 0] = 9999.9
            st = oracle.batch_stats(vv, zz, nlay, so, sd)
        else:
            st = oracle.batch_stats(s["v"][:nb], s["z"][:nb], s["n"][:nb], so, sd)
        w_min = st["flops_min"] / st["rays"]
        evals = B * S / (ms * 1e-3)
        tf = w_min * evals / 1e12
        print(json.dumps({
            "shape": name, "models": B, "sources": S, "ms_per_step": ms, "evals_per_s": evals,
            "logL_per_s": B / (ms * 1e-3), "flops_per_eval_min": w_min,
            "flops_per_eval_reference": st["flops_ref"] / st["rays"],
            "mean_layers_above_source": st["sum_nl"] / st["rays"],
            "bisect_share": st["bisect"] / st["rays"],
            "fp64_tflops_achieved": tf, "fp64_tflops_peak": peak, "roofline_frac": tf / peak,
            "tile_models": int(rt.get_stat("tile_models")), "tile_sources": int(rt.get_stat("tile_sources")),
            "grid": int(rt.get_stat("grid")), "ctas_per_sm": int(rt.get_stat("ctas_per_sm")),
            "smem_bytes": int(rt.get_stat("smem_bytes"))}))


if __name__ == "__main__":
    main()
