#!/usr/bin/env python
"""Generate tests/golden/reference_golden.json from the reference's own shipped artefacts.

Run in the build container (needs /root/reference; the GPU box does not have it, which is
why the output is committed):

    python tests/golden/make_golden.py

Sources of truth (paths relative to /root/reference):
  * rays.dat:1-40           full-precision per-layer ray geometry written by the reference's own
                            dff (keep_delta > 0) for the test_1 model and its 20 sources
                            (writer: subroutineR-quiet.f90:157-164).
  * raytracerR-export-data-to-MCMC.nb.html:274,357   stored notebook output: 5 travel times
                            (7 digits) and the per-layer ray table (8 digits) for NSrc <- 5.
  * test_1/test_1_map.dat, test_1/test_1_src_data.txt, test_1/test_1_RT.txt   the example's
                            model, sources (15 digits) and noisy observations.
  * raytracerR-export-data-to-MCMC.Rmd:44-47   how the sources were drawn:
                            set.seed(12); z=runif(n,1050,4200); x=runif(n,10,6500); y=runif(n,10,4500).
    R's Mersenne-Twister seeding is re-implemented below so the *full precision* sources that
    produced rays.dat can be regenerated (the .txt file keeps only 15 digits).
  * eqdf.csv + rayTracerR.R:22-27   the README example (v=3100,4270,6000; z=2000,4000).

No reference SOURCE code is copied; only data the reference ships or prints.
"""
import json
import math
import os
import re

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.json")


def r_set_seed_uniforms(seed, n):
    """First n values of R's unif_rand() after set.seed(seed) (Mersenne-Twister, default kind).

    R: RNG_Init scrambles the seed 50 times with the LCG 69069*s+1 (mod 2^32), then fills the
    625-word MT seed vector with the next 625 LCG outputs; FixupSeeds forces word 0 (the
    position) to 624.  MT_genrand() returns genrand_int32() * 2.3283064365386963e-10."""
    s = seed & 0xFFFFFFFF
    for _ in range(50):
        s = (69069 * s + 1) & 0xFFFFFFFF
    words = []
    for _ in range(625):
        s = (69069 * s + 1) & 0xFFFFFFFF
        words.append(s)
    key = np.array(words[1:], dtype=np.uint32)
    bg = np.random.MT19937()
    bg.state = {"bit_generator": "MT19937", "state": {"key": key, "pos": 624}}
    raw = bg.random_raw(n)
    return [float(int(r)) * 2.3283064365386963e-10 for r in raw]


def r_sources(n, seed=12):
    u = r_set_seed_uniforms(seed, 3 * n)
    z = [1050.0 + (4200.0 - 1050.0) * u[i] for i in range(n)]
    x = [10.0 + (6500.0 - 10.0) * u[n + i] for i in range(n)]
    y = [10.0 + (4500.0 - 10.0) * u[2 * n + i] for i in range(n)]
    off = [math.sqrt(x[i] * x[i] + y[i] * y[i]) for i in range(n)]
    return off, z


def floats(line):
    return [float(t) for t in line.split()]


def main():
    g = {}

    # --- config 1: the shipped test_1 example -------------------------------------------------
    m = floats(open(f"{REF}/test_1/test_1_map.dat").read())
    k = int(m[0])
    nodes = m[1:1 + 2 * k]
    zn, vp = nodes[0::2], nodes[1::2]
    src = [floats(l) for l in open(f"{REF}/test_1/test_1_src_data.txt") if l.strip()]
    tobs = [float(l) for l in open(f"{REF}/test_1/test_1_RT.txt") if l.strip()]
    off20, dep20 = r_sources(20)
    # the regenerated sources must be the file's sources at the file's precision
    for (fo, fd), o, d in zip(src, off20, dep20):
        assert abs(fo - o) <= 1e-11 * abs(o) and abs(fd - d) <= 1e-11 * abs(d), (fo, o, fd, d)
    rays = [floats(l) for l in open(f"{REF}/rays.dat") if l.strip()]
    assert len(rays) == 40
    g["config1"] = {
        "cite": "test_1/test_1_map.dat; test_1/test_1_src_data.txt; test_1/test_1_RT.txt; rays.dat:1-40",
        "vels": vp,
        "depths": zn[1:],
        "sigma_map": m[21],  # 1-based index NFPMX+2 = 22 (prjmh_temper_rf.f90:152-189)
        "src_offset_file": [s[0] for s in src],
        "src_depth_file": [s[1] for s in src],
        "src_offset_full": off20,
        "src_depth_full": dep20,
        "tobs": tobs,
        "rays_dat": [{"delta": rays[2 * i], "h": rays[2 * i + 1]} for i in range(20)],
    }

    # --- notebook known-answer test ------------------------------------------------------------
    html = open(f"{REF}/raytracerR-export-data-to-MCMC.nb.html").read().splitlines()
    assert "NSrc &lt;- 5" in html[222]
    times = floats(re.search(r"\[1\]([^<]*)<", html[273]).group(1))
    table = json.loads(html[356])["data"]
    rows = [{"delta": float(r["1"]), "depth": float(r["2"]), "ray": int(r["3"])} for r in table]
    off5, dep5 = r_sources(5)
    g["notebook"] = {
        "cite": "raytracerR-export-data-to-MCMC.nb.html:223,274,357",
        "vels": vp,
        "depths": zn[1:],
        "src_offset": off5,
        "src_depth": dep5,
        "timeP_7digits": times,
        "ray_table": rows,
    }

    # --- README / rayTracerR.R example ----------------------------------------------------------
    eq = [l.strip().split(",") for l in open(f"{REF}/eqdf.csv") if l.strip()]
    hdr = [h.strip('"') for h in eq[0]]
    cols = {h: [float(r[i]) for r in eq[1:]] for i, h in enumerate(hdr) if h in ("x", "y", "z")}
    g["readme_example"] = {
        "cite": "rayTracerR.R:22-33; eqdf.csv",
        "vels": [3100.0, 4270.0, 6000.0],
        "depths": [2000.0, 4000.0],
        "src_offset": [math.sqrt(x * x + y * y) for x, y in zip(cols["x"], cols["y"])],
        "src_depth": cols["z"],
    }

    with open(OUT, "w") as fh:
        json.dump(g, fh, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
