#!/usr/bin/env python
"""Pinned host -> device copy bandwidth of the box, by copy size (what bounds bench.py's e2e leg)."""
import json
import torch
dev = torch.device("cuda:0")
out = {}
for mb in (0.3, 1, 6, 24, 176):
    n = int(mb * 1e6) // 8
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    d = torch.empty(n, dtype=torch.float64, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, int(400 / mb))
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    out[f"h2d_{mb}MB_GBps"] = n * 8 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    e0.record()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    out[f"d2h_{mb}MB_GBps"] = n * 8 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
print(json.dumps(out))
