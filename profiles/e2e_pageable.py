#!/usr/bin/env python
"""Config 2 through dff_batch from pageable numpy arrays: the pinned-ring staging on and off,
against the pinned-buffer figure.  One JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import raytracerfortran_b200 as rt
from raytracerfortran_b200 import workloads

c = workloads.CONFIGS["config2"]
v, z, nl = workloads.make_models(c["B"], c["nlayers"], c["seed"])
so, sd = workloads.make_sources(c["nsrc"], c["seed"])
tobs, sigma = workloads.make_observations(np.full(c["nsrc"], 1.2), c["B"], c["seed"])
out = np.empty(c["B"])
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
hv, hz, hn, hg, hl = pin(v), pin(z), pin(nl), pin(sigma), pin(out)
res = {}
def run(name, args, stage):
    rt.set_option("stage_pageable", stage)
    ts = []
    for i in range(8):
        t0 = time.perf_counter(); rt.dff_batch(*args[:3], so, sd, tobs=tobs, sigma=args[3], want_times=False, out_logL=args[4]); ts.append(time.perf_counter() - t0)
    res[name] = {"first_call_ms": 1e3 * ts[0], "ms_per_call": 1e3 * float(np.median(ts[3:])), "gpu_span_ms": rt.get_stat("total_ms")}
run("pinned", (hv, hz, hn, hg, hl), -1)
run("pageable_direct", (v, z, nl, sigma, out), 0)
run("pageable_ring", (v, z, nl, sigma, out), -1)
t0 = time.perf_counter(); a = hv.copy(); res["host_memcpy_GBps_one_thread"] = hv.nbytes / (time.perf_counter() - t0) / 1e9
res["cores"] = os.cpu_count()
print(json.dumps(res))
