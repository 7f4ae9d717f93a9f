#!/usr/bin/env python
"""replica.f90 as one batched call: write a small synthetic `_voro_sample.txt` in the sampler's
format, read it back with burn-in and thinning, and re-evaluate every kept state's likelihood."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracerfortran_b200 import samplefile, workloads  # noqa: E402

B, NLMX, NSRC = 2000, 10, 20
k, vp, zi = workloads.make_transd_models(B, NLMX, 1)
voro = np.zeros((B, 2, NLMX))
voro[:, 1, :] = vp
voro[:, 0, 1:] = zi
rng = np.random.default_rng(1)
sigma = rng.uniform(0.001, 0.07, B)
rows = samplefile.pack_rows(np.zeros(B), np.zeros(B), np.zeros(B), k, voro, sigma)
path = os.path.join(tempfile.mkdtemp(), "demo_voro_sample.txt")
samplefile.write_samples(path, rows)
so, sd = workloads.make_sources(NSRC, 1)
tobs = 1.0 + 0.5 * rng.random(NSRC)
smp, logL, pred = samplefile.replica_sweep(path, NLMX, so, sd, tobs, burnin=500, thin=5)
print(f"{len(logL)} kept samples re-evaluated; logL range [{logL.min():.2f}, {logL.max():.2f}]; "
      f"DpredRT {pred.shape}")
