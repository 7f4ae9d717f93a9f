#!/usr/bin/env python
"""Config 4 as the sampler would run it: R tempering replicas x C independent chains per replica
per GPU (default 8 x 1024 = 8192 chains), 256 sources, every chain doing real fixed-dimension MH
moves on the device (rtb200_mh_step_device: PROPOSAL + INTERPLAYER_novar + CHECKBOUNDS2 + LOGLHOOD
+ accept), one tempering swap round per sweep (all-gather over NCCL when launched with torchrun).

    python profiles/config4_chains.py [--sweeps K]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P profiles/config4_chains.py

Chain states never leave HBM.  Rank 0 prints one JSON line."""
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
    ap.add_argument("--sweeps", type=int, default=5, help="rounds of --moves moves + one swap round")
    ap.add_argument("--moves", type=int, default=20, help="MH moves per chain between swap rounds")
    ap.add_argument("--replicas", type=int, default=8, help="tempering replicas per GPU")
    ap.add_argument("--chains", type=int, default=1024, help="independent chains per replica")
    ap.add_argument("--eager", action="store_true", help="round 1's path: host-side deviates, one graph per run of moves")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import raytracerfortran_b200 as rt
    from raytracerfortran_b200 import chains, tempering, workloads

    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ["RTB200_DEVICE"] = str(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    R, C, ldk, nsrc = args.replicas, args.chains, 30, 256
    B = R * C
    rng = np.random.default_rng(40 + rank)
    k = np.clip(rng.poisson(3.01, B), 1, ldk).astype(np.int32)
    voro = np.zeros((B, 2, ldk))
    for b in range(B):
        n = int(k[b])
        voro[b, 0, 1:n] = np.cumsum(100.1 + rng.random(n - 1) * (9000.0 / max(n, 1)))
        voro[b, 1, :n] = rng.uniform(1500.0, 10000.0, n)
    so, sd = workloads.make_sources(nsrc, 4)
    tobs, sigma = workloads.make_observations(np.full(nsrc, 1.3), B, 4 + rank)
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tk, tv, ts, td, to, tg = f(k), f(voro), f(so), f(sd), f(tobs), f(sigma)
    ladder = tempering.temperature_ladder(R * world, 1.4)[rank * R:(rank + 1) * R]
    beta = f(np.repeat(ladder, C))
    # logL of the starting states through the same kernels
    from raytracerfortran_b200 import device
    tl = device.dff_batch_device(tv[:, 1, :].contiguous(), tv[:, 0, 1:].contiguous(), tk, ts, td,
                                 tobs=to, sigma=tg, kmode=True)["logL"]
    prior = chains.prior_array()
    gen = torch.Generator(device=dev).manual_seed(100 + rank)

    pos = torch.zeros(B, dtype=torch.int32, device=dev)
    M = args.moves

    pk = chains.poisson_pk(3.01, 1, ldk)

    sd_prior = chains.sd_prior_array()
    graph = chains.McmcGraph(tk, tv, tl, tg, beta, M, prior, sd_prior, pk, 1, ldk, ts, td, to, seed=2026 + rank)
    swapper = tempering.SwapRound(B, dev)

    def sweep_graph(i):
        # one captured iteration (birth/death + M moves of every chain's own sweep + sigma move, all
        # deviates drawn on the device), then the swap round on a side stream; the next iteration's
        # graph waits for the new betas (the graph reads beta in its first accept test)
        swapper.wait()
        graph.run(1)
        swapper.launch(tl, beta, 2026, i)
        return None

    def sweep(i):
        # the birth/death move that opens EXPLORE_MH_NOVARPAR (:658-710), then the fixed-k moves
        u = torch.rand((5, B), dtype=torch.float64, device=dev, generator=gen)
        idel = (2 + torch.floor(u[4] * (tk - 1).clamp(min=1))).to(torch.int32)
        chains.bd_step_device(tk, tv, tl, u[0].contiguous(), idel, u[1].contiguous(), u[2].contiguous(),
                              u[3].contiguous(), beta, tg, prior, pk, 1, ldk, ts, td, to)
        acc = chains.mh_moves_device(tk, tv, tl, pos, M, beta, tg, prior, ts, td, to, generator=gen)
        # swap round between replicas
        nb, _ = tempering.tempering_swap_round_device(tl, beta, seed=2026, round_index=i)
        beta.copy_(nb)          # same buffer every round: the captured graph of moves stays valid
        return acc

    use_graph = not args.eager
    swapper.launch(tl, beta, 2026, 0)
    (sweep_graph if use_graph else sweep)(0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = rt.get_stat("launches")
    graph.tally.zero_()
    t0 = time.perf_counter()
    acc_t = 0
    for i in range(args.sweeps):
        if use_graph:
            sweep_graph(1 + i)
        else:
            acc_t = acc_t + sweep(1 + i).sum()
    if use_graph:
        swapper.wait()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if use_graph:
        acc_t = graph.tally[0].sum() + graph.tally[2].sum()
    moves = B * (M + (2 if use_graph else 1)) * args.sweeps
    tt = torch.tensor([wall, float(moves), float(acc_t.item())], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        wall = float(mx[0])
    if rank == 0:
        moves_all, acc_all = float(tt[1]), float(tt[2])
        print(json.dumps({
            "workload": f"config4 as chains: {R * world} replicas x {C} chains, k ~ Poisson(3.01) in 1..30, "
                        f"{nsrc} sources, {world} GPU(s)",
            "rounds": args.sweeps, "moves_per_round": M, "seconds": wall, "mh_moves": moves_all,
            "mh_moves_per_s": moves_all / wall, "evals_per_s": moves_all * nsrc / wall,
            "acceptance": acc_all / moves_all,
            "mode": "eager (torch deviates, graph of the fixed-k moves only)" if args.eager else
                    "one CUDA graph per iteration, deviates drawn on the device, swap round on a side stream",
            "kernel_launches_per_move": 3, "birth_death_moves_per_round": 1, "final_mean_k": float(tk.double().mean().item()), "library_launches": rt.get_stat("launches") - launches0,
            "max_k": int(k.max())}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
