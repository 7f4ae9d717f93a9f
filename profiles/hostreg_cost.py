#!/usr/bin/env python
"""Cost of cudaHostRegister / cudaHostUnregister on a 92 MB pageable array (per-call pinning as an
alternative to the staging ring), and the ring with different copy-thread counts."""
import ctypes, json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "ring":
    import raytracerfortran_b200 as rt
    from raytracerfortran_b200 import workloads
    c = workloads.CONFIGS["config2"]
    v, z, nl = workloads.make_models(c["B"], c["nlayers"], c["seed"])
    so, sd = workloads.make_sources(c["nsrc"], c["seed"])
    tobs, sigma = workloads.make_observations(np.full(c["nsrc"], 1.2), c["B"], c["seed"])
    out = np.empty(c["B"])
    ts = []
    for i in range(8):
        t0 = time.perf_counter(); rt.dff_batch(v, z, nl, so, sd, tobs=tobs, sigma=sigma, want_times=False, out_logL=out); ts.append(time.perf_counter() - t0)
    print(json.dumps({"copy_threads": os.environ.get("RTB200_COPY_THREADS"), "ms_per_call": 1e3 * float(np.median(ts[3:]))}))
    sys.exit(0)
import torch
torch.zeros(1, device="cuda")
rt_ = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL(torch.__path__[0] + "/lib/libcudart.so.12")
a = np.random.default_rng(0).random(11_500_000)
res = []
for i in range(4):
    t0 = time.perf_counter(); rc = rt_.cudaHostRegister(ctypes.c_void_p(a.ctypes.data), ctypes.c_size_t(a.nbytes), 0); t1 = time.perf_counter()
    rc2 = rt_.cudaHostUnregister(ctypes.c_void_p(a.ctypes.data)); t2 = time.perf_counter()
    res.append((rc, rc2, 1e3 * (t1 - t0), 1e3 * (t2 - t1)))
print(json.dumps({"bytes": a.nbytes, "register_unregister_ms": res}))
for n in (4, 8, 12, 16):
    env = dict(os.environ, RTB200_COPY_THREADS=str(n))
    print(subprocess.run([sys.executable, __file__, "ring"], env=env, capture_output=True, text=True).stdout.strip())
