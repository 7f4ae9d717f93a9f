#!/usr/bin/env python
"""The reference's own example (rayTracerR.R:21-37, raytracerR-export-data-to-MCMC.Rmd:60-90) on
the B200 path: the test_1 layered model, its 20 sources, travel times through `dff_`, the ray
geometry file `rays.dat` (keep_delta > 0) and the log-likelihood of the shipped observations."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracerfortran_b200 as rt  # noqa: E402

g = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")))["config1"]
v, d = np.array(g["vels"]), np.array(g["depths"])
srco, srcd = np.array(g["src_offset_file"]), np.array(g["src_depth_file"])

timeP = rt.dff(v, d, srco, srcd, keep_delta=10)          # .Fortran("dff", ..., keep_delta = 10)
print("travel times [s]:", np.round(timeP, 7))
print("rays.dat lines  :", sum(1 for _ in open("rays.dat")))

# one chain state through LOGLHOOD (k = 7 nodes), sigma as in test_1_map.dat
k = np.array([len(v)], dtype=np.int32)
logL, pred = rt.loglhood_batch(k, v[None, :], d[None, :], srco, srcd, np.array(g["tobs"]),
                               np.array([g["sigma_map"]]), want_pred=True)
print("logL of the MAP model:", logL[0])
