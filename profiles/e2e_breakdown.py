#!/usr/bin/env python
"""Where the end-to-end time of one dff_batch call goes (config 2, pinned host buffers):
python wall time, the library's GPU span (first H2D to last D2H), python-side validation."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import raytracerfortran_b200 as rt
from raytracerfortran_b200 import workloads

B = 1 << 20
v, z, nl = workloads.make_models(B, 10, 2)
so, sd = workloads.make_sources(64, 2)
tobs, sigma = workloads.make_observations(np.full(64, 1.3), B, 2)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
hv, hz, hn, hg = pin(v), pin(z), pin(nl), pin(sigma)
h_ll = torch.empty(B, dtype=torch.float64).pin_memory().numpy()
lib = rt._lib.load()
for k, val in [a.split("=") for a in sys.argv[1:]]:
    rt.set_option(k, float(val))
res = []
for i in range(8):
    t0 = time.perf_counter()
    rt.dff_batch(hv, hz, hn, so, sd, tobs=tobs, sigma=hg, want_times=False, out_logL=h_ll)
    t1 = time.perf_counter()
    res.append((1e3 * (t1 - t0), rt.get_stat("total_ms"), rt.get_stat("kernel_ms")))
t0 = time.perf_counter(); m = hn.max(); t1 = time.perf_counter()
print(json.dumps({"wall_ms": [round(r[0], 3) for r in res], "gpu_span_ms": [round(r[1], 3) for r in res],
                  "sum_kernel_ms": [round(r[2], 3) for r in res], "np_max_ms": 1e3 * (t1 - t0)}))
