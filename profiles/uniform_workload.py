#!/usr/bin/env python
"""Diagnostic: how much of the lost issue bandwidth is ray-to-ray variance (drain, barrier skew)?
Runs config-2-sized batches where (a) every ray is the same ray, (b) every model is the same but the
64 sources differ, (c) the real random workload, and prints ms/step and the oracle's mean solver
passes per ray, so the cost per solver pass can be compared."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import oracle
import raytracerfortran_b200 as rt
from raytracerfortran_b200 import device, workloads

dev = torch.device("cuda:0")
f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
B, S = 400_000, 64
v, z, nl = workloads.make_models(B, 10, 2)
so, sd = workloads.make_sources(S, 2)
cases = {"random (config 2 inputs)": (v, z, so, sd)}
# pick a typical bisection ray: model 0, a mid-depth source
vv, zz = np.repeat(v[:1], B, 0), np.repeat(z[:1], B, 0)
cases["same model, 64 different sources"] = (vv, zz, so, sd)
cases["one ray repeated"] = (vv, zz, np.full(S, so[3]), np.full(S, sd[3]))
for name, (a, b, o, d) in cases.items():
    st = oracle.batch_stats(a[:300], b[:300], nl[:300], o, d)
    passes = (st["n_ffp_min"] + st["n_f_min"]) / st["rays"] + 1.0
    tv, tz, tn, ts, td = f(a), f(b), f(nl), f(o), f(d)
    to, tg = f(np.ones(S)), f(np.full(B, 0.02))
    ll = torch.empty(B, dtype=torch.float64, device=dev)
    step = lambda: device.dff_batch_device(tv, tz, tn, ts, td, tobs=to, sigma=tg, logL=ll)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    layers = st["sum_nl"] / st["rays"]
    print(json.dumps({"case": name, "ms_per_step": ms, "evals_per_s": B * S / ms * 1e3,
                      "mean_layers": layers, "solver_passes_per_ray": passes,
                      "ns_per_pass_layer_pair": ms * 1e6 / (B * S * passes * (layers / 2 + 0.5))}))
