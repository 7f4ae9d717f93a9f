#!/usr/bin/env python
"""Multi-GPU check of the device-side swap round over NCCL (run under torchrun): every rank's new
betas must equal the numpy restatement applied to the all-gathered (logL, beta), for several
rounds, and all ranks must agree on the accept mask.  Rank 0 prints one JSON line."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from oracle import tempering_ref
from raytracerfortran_b200 import tempering

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ["RTB200_DEVICE"] = str(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
n_local = 8192
n = n_local * world
rng = np.random.default_rng(5)                       # the same global state on every rank
logL_all = rng.normal(-60, 40, n)
beta_all = rng.permutation(tempering.temperature_ladder(n, 1.0005))
lo = rank * n_local
logL = torch.from_numpy(logL_all[lo:lo + n_local].copy()).to(dev)
beta = torch.from_numpy(beta_all[lo:lo + n_local].copy()).to(dev)
sr = tempering.SwapRound(n_local, dev, want_info=True)
ok, accepted = True, 0
for rnd in range(6):
    want, pairs, acc, _ = tempering_ref.swap_round(logL_all, beta_all, 99, rnd)
    sr.launch(logL, beta, 99, rnd)
    sr.wait()
    torch.cuda.synchronize()
    ok = ok and np.array_equal(beta.cpu().numpy().view(np.uint64), want[lo:lo + n_local].view(np.uint64))
    ok = ok and np.array_equal(sr.accept.cpu().numpy().astype(bool), acc)
    beta_all = want
    accepted += int(acc.sum())
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "chains": n, "rounds": 6, "all_ranks_match_reference": bool(flag.item()),
                      "swaps_accepted": accepted, "allgather_bytes_per_rank": 16 * n_local}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
