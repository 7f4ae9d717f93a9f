#!/usr/bin/env python
"""Latency of the reference's own entry point on its own example (BASELINE.json configs[0]):
dff_ on the test_1 model x 20 sources, one model per call, host buffers in and out -- what
R's .Fortran("dff", ...) pays per call -- next to the CPU oracle for the same call."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import oracle
    import raytracerfortran_b200 as rt
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")))["config1"]
    v, z = np.array(g["vels"]), np.array(g["depths"])
    so, sd = np.array(g["src_offset_full"]), np.array(g["src_depth_full"])
    for _ in range(20):
        t = rt.dff(v, z, so, sd)
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n):
        t = rt.dff(v, z, so, sd)
    gpu_us = (time.perf_counter() - t0) / n * 1e6
    t0 = time.perf_counter()
    for _ in range(n):
        to, _, _ = oracle.trace_rays(v, z, so, sd)
    cpu_us = (time.perf_counter() - t0) / n * 1e6
    tmp = "/tmp/rays_latency.dat"
    t0 = time.perf_counter()
    for _ in range(n):
        oracle.trace_rays(v, z, so, sd, keep_delta=-1, rays_path=tmp)   # with the per-call truncate
    cpu_file_us = (time.perf_counter() - t0) / n * 1e6
    assert np.array_equal(t.view(np.uint64), to.view(np.uint64))
    print(json.dumps({"workload": "config1: dff_ on test_1 (6 interfaces, 20 sources), one call",
                      "gpu_us_per_call": gpu_us, "kernel_ms_last": rt.get_stat("kernel_ms"),
                      "cpu_oracle_us_per_call": cpu_us,
                      "cpu_oracle_with_rays_dat_truncate_us_per_call": cpu_file_us,
                      "note": "python ctypes overhead included on both sides"}))


if __name__ == "__main__":
    main()
