#!/usr/bin/env python
"""BASELINE.json config 4: 64 tempering replicas x 1024 proposals per step, sharded over the ranks
(8 replicas per GPU at 8 GPUs), 256 sources, logL fused, one swap round per step whose only
inter-GPU traffic is the all-gather of (logL, beta) per replica over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P profiles/config4_tempering.py [--steps K]

Every rank evaluates its replicas' proposals in one launch, applies the reference's MH rule per
replica to the first proposal (EXPLORE_MH_NOVARPAR, one proposal per chain per step) and takes part
in the swap round.  Rank 0 prints one JSON line; every rank checks that the swap decisions it
derived are identical to rank 0's."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--replicas", type=int, default=64)
    ap.add_argument("--proposals", type=int, default=1024)
    ap.add_argument("--host-swap", action="store_true", help="swap decisions on the host (numpy Philox)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from raytracerfortran_b200 import tempering, workloads

    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ["RTB200_DEVICE"] = str(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    R = args.replicas // world
    P, kmax, nsrc = args.proposals, 30, 256
    k, vp, zi = workloads.make_transd_models(R * P, kmax, 4 + rank)
    so, sd = workloads.make_sources(nsrc, 4)
    tobs, sigma = workloads.make_observations(np.full(nsrc, 1.3), R * P, 4 + rank)
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tk, tv, tz = f(k).reshape(R, P), f(vp).reshape(R, P, kmax), f(zi).reshape(R, P, kmax - 1)
    ts, td, to, tg = f(so), f(sd), f(tobs), f(sigma).reshape(R, P)
    beta = f(tempering.temperature_ladder(args.replicas, 1.4)[rank * R:(rank + 1) * R])
    logL = torch.empty((R, P), dtype=torch.float64, device=dev)
    cur = None
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    timers, nacc = [], []

    def one_step(step):
        nonlocal cur, beta
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        tempering.evaluate_replicas(tk, tv, tz, ts, td, to, tg, logL=logL)
        new = logL[:, step % P]
        if cur is None:
            cur = new.clone()
        else:
            u = torch.rand(R, dtype=torch.float64, device=dev, generator=gen)
            acc = tempering.mh_accept(cur, new, beta, u)
            cur = torch.where(acc, new, cur)
        e1.record()
        if args.host_swap:
            beta, st = tempering.tempering_swap_round(cur, beta, seed=2026, round_index=step)
            acc = torch.from_numpy(st["accept"].astype(np.int64)).to(dev)
        else:
            beta, st = tempering.tempering_swap_round_device(cur, beta, seed=2026, round_index=step)
            acc = st["accept"].to(torch.int64)
        e2.record()
        timers.append((e0, e1, e2))
        nacc.append(acc.sum())
        return acc

    for w in range(3):
        one_step(w)
    torch.cuda.synchronize()
    timers.clear()
    nacc.clear()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    last = None
    for s in range(args.steps):
        last = one_step(3 + s)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t_eval = sum(a.elapsed_time(b) for a, b, _ in timers)
    t_swap = sum(b.elapsed_time(c) for _, b, c in timers)
    swaps = int(torch.stack(nacc).sum().item())
    # every rank derived the same decisions
    if world > 1:
        mine = last.clone()
        ref = mine.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(mine, ref), "swap decisions differ between ranks"
        tt = torch.tensor([wall, t_eval, t_swap], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        wall, t_eval, t_swap = tt.tolist()
    if rank == 0:
        evals = args.replicas * P * nsrc * args.steps
        print(json.dumps({
            "workload": f"config4: {args.replicas} replicas x {P} proposals/step x {nsrc} sources over {world} GPU(s)",
            "steps": args.steps, "ms_per_step_wall": 1e3 * wall / args.steps,
            "ms_per_step_evaluate": t_eval / args.steps, "ms_per_step_swap_round": t_swap / args.steps,
            "evals_per_s": evals / wall, "loglhood_per_s": args.replicas * P * args.steps / wall,
            "swap_allgather_bytes_per_rank": 16 * R, "swaps_accepted": swaps,
            "swap_pairs_per_round": args.replicas // 2,
            "swap_decisions": "host (numpy Philox)" if args.host_swap else "device"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
