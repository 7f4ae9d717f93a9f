#!/usr/bin/env python
"""Solver passes by number of active lanes (profiling build -DRTB_HIST only):
    RTB200_LIB=build/ab/lib_hist.so python profiles/lane_hist.py [models]
config-2 models (10 interfaces x 64 sources); bins 0..32 while the warp's list segment still has
rays, 33..65 after it is exhausted (the drain)."""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import raytracerfortran_b200 as rt
from raytracerfortran_b200 import _lib, device, workloads

B = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
cfg = workloads.CONFIGS["config2"]
v, z, nl = workloads.make_models(B, cfg["nlayers"], cfg["seed"])
so, sd = workloads.make_sources(cfg["nsrc"], cfg["seed"])
dev = torch.device("cuda:0")
f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
tobs, sigma = workloads.make_observations(np.full(len(so), 1.5), B, 1)
args = (f(v), f(z), f(nl.astype(np.int32)), f(so), f(sd))
lib = _lib.load()
lib.rtb200_debug_hist.argtypes = [ctypes.c_void_p, ctypes.c_int]
h = (ctypes.c_ulonglong * 66)()
device.dff_batch_device(*args, tobs=f(tobs), sigma=f(sigma))
lib.rtb200_debug_hist(h, 1)
device.dff_batch_device(*args, tobs=f(tobs), sigma=f(sigma))
lib.rtb200_debug_hist(h, 1)
h = np.array(list(h), dtype=np.float64)
tot = h.sum()
live, drain = h[:33], h[33:]
lanes = np.arange(33)
print(json.dumps({"variant": int(rt.get_stat("variant")), "passes": tot, "rays": B * len(so),
                  "passes_live_share": live.sum() / tot, "passes_drain_share": drain.sum() / tot,
                  "mean_active_live": float((live * lanes).sum() / max(live.sum(), 1)),
                  "mean_active_drain": float((drain * lanes).sum() / max(drain.sum(), 1)),
                  "idle_lane_share_live": float((live * (32 - lanes)).sum() / (32 * tot)),
                  "idle_lane_share_drain": float((drain * (32 - lanes)).sum() / (32 * tot)),
                  "drain_passes_le4_share": float(drain[:5].sum() / tot), "drain_passes_le8_share": float(drain[:9].sum() / tot),
                  "drain_hist": [int(x) for x in drain], "live_hist": [int(x) for x in live]}))
