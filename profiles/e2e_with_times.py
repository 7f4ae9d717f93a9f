#!/usr/bin/env python
"""Config 2 end to end through dff_batch with the travel times returned to the host as well
(512 MB of D2H per call on top of the 8 MB of logL): pinned host buffers, chunked pipeline."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import raytracerfortran_b200 as rt
from raytracerfortran_b200 import workloads

B, S = 1_000_000, 64
v, z, nl = workloads.make_models(B, 10, 2)
so, sd = workloads.make_sources(S, 2)
tobs, sigma = workloads.make_observations(np.full(S, 1.3), B, 2)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
hv, hz, hn, hg = pin(v), pin(z), pin(nl), pin(sigma)
h_ll = torch.empty(B, dtype=torch.float64).pin_memory().numpy()
h_t = torch.empty((B, S), dtype=torch.float64).pin_memory().numpy()
out = {}
for name, kw in (("logL only", dict(want_times=False)), ("logL + travel times", dict(out_times=h_t)),
                 ("logL + travel times + ray parameters", dict(out_times=h_t, want_p=True,
                                                              out_p=torch.empty((B, S), dtype=torch.float64).pin_memory().numpy()))):
    for _ in range(2):
        rt.dff_batch(hv, hz, hn, so, sd, tobs=tobs, sigma=hg, out_logL=h_ll, **kw)
    t0 = time.perf_counter()
    for _ in range(5):
        rt.dff_batch(hv, hz, hn, so, sd, tobs=tobs, sigma=hg, out_logL=h_ll, **kw)
    dt = (time.perf_counter() - t0) / 5
    out[name] = {"ms_per_call": 1e3 * dt, "evals_per_s": B * S / dt}
print(json.dumps(out))
