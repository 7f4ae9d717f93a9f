"""Sharding across GPUs and the parallel-tempering swap step.

The hot path shards on the model axis with no data-path collective: every (model, source) ray is
independent and a model's logL reduces over its own sources only, so each rank (one process per
GPU) evaluates a contiguous slice of the models (`shard_range`) or, for tempering, its own
replicas and their proposals.

The only inter-GPU traffic is the swap step of parallel tempering.  The reference does it through
a master rank that receives two whole chain states (1.7 KB each, MPI_RECV/MPI_SEND,
prjmh_temper_rf.f90:331-349) and swaps the STATES with probability
min(1, exp((beta2 - beta1) (logL1 - logL2)))            (TEMPSWP_MH, prjmh_temper_rf.f90:1329-1384).
Here every rank all-gathers (logL, beta) of all replicas -- 16 bytes per replica -- derives the
same pairing and the same accept/reject decisions from a shared counter-based generator, and
swaps the BETAS of accepted pairs locally, which is the same Markov kernel with zero state
traffic.  `torch.distributed` is the plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist

from .raymod import shard_range  # noqa: F401  (re-exported)


def swap_pairs(n_replicas, seed, round_index):
    """Disjoint random pairs of replicas for one swap round, identical on every rank.

    The reference pairs whichever two chains report to the master first
    (prjmh_temper_rf.f90:326-336), i.e. arbitrary temperatures; a seeded random perfect matching
    is the deterministic analogue.  Returns (pairs [n//2, 2] int64, u [n//2] uniforms)."""
    rng = np.random.Generator(np.random.Philox(key=int(seed) & 0xFFFFFFFFFFFFFFFF,
                                               counter=[int(round_index), 0, 0, 0]))
    perm = rng.permutation(n_replicas)
    npair = n_replicas // 2
    pairs = perm[:2 * npair].reshape(npair, 2)
    return pairs, rng.random(npair)


def swap_decisions(logL, beta, seed, round_index):
    """Accept mask of TEMPSWP_MH for every pair of this round.

    logratio = (beta2 - beta1) * (logL1 - logL2); accept if u <= exp(logratio)
    (prjmh_temper_rf.f90:1341-1344).  Returns (pairs, accept [npair] bool)."""
    logL = np.asarray(logL, dtype=np.float64)
    beta = np.asarray(beta, dtype=np.float64)
    pairs, u = swap_pairs(len(logL), seed, round_index)
    i, j = pairs[:, 0], pairs[:, 1]
    logratio = (beta[j] - beta[i]) * (logL[i] - logL[j])
    with np.errstate(over="ignore"):
        accept = u <= np.exp(logratio)
    return pairs, accept


def apply_swaps(beta, pairs, accept):
    """Exchange the betas of the accepted pairs (states stay where they are)."""
    beta = np.array(beta, dtype=np.float64, copy=True)
    for (i, j), ok in zip(pairs, accept):
        if ok:
            beta[i], beta[j] = beta[j], beta[i]
    return beta


def allgather_replicas(logL_local, beta_local, group=None):
    """The one collective of the path: all-gather (logL, beta) of every replica.

    logL_local, beta_local: 1-D float64 tensors of this rank's replicas (same count on every
    rank), on the device the process group communicates from (CUDA for NCCL, CPU for gloo).
    Returns two numpy arrays of all replicas, ordered by rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return logL_local.detach().cpu().numpy().copy(), beta_local.detach().cpu().numpy().copy()
    world = dist.get_world_size(group)
    mine = torch.stack([logL_local.to(torch.float64), beta_local.to(torch.float64)], dim=1).contiguous()
    out = torch.empty((world * mine.shape[0], 2), dtype=torch.float64, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    flat = out.cpu().numpy()
    return flat[:, 0].copy(), flat[:, 1].copy()


def tempering_swap_round(logL_local, beta_local, seed, round_index, group=None):
    """One swap round.  Returns (new beta_local tensor, stats dict); every rank computes the same
    global decision and keeps its own slice."""
    n_local = logL_local.numel()
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    logL_all, beta_all = allgather_replicas(logL_local, beta_local, group)
    pairs, accept = swap_decisions(logL_all, beta_all, seed, round_index)
    new_all = apply_swaps(beta_all, pairs, accept)
    mine = new_all[rank * n_local:(rank + 1) * n_local]
    out = torch.from_numpy(mine.copy()).to(beta_local.device)
    return out, {"pairs": pairs, "accept": accept, "accepted": int(accept.sum()),
                 "bytes_per_rank": 16 * n_local}


def temperature_ladder(n_replicas, dTlog, n_cold=1):
    """beta = 1/T with T = dTlog**j, the first n_cold chains at T = 1
    (NPTCHAINS1 / dTlog of the parameter file, prjmh_temper_rf.f90:116-134)."""
    j = np.maximum(np.arange(n_replicas) - (n_cold - 1), 0)
    return 1.0 / np.power(float(dTlog), j)


def evaluate_replicas(k, voro_vp, ziface, src_offset, src_depth, tobs, sigma, logL=None):
    """logL of every proposal of this rank's replicas in one launch.

    k [R, P] int32, voro_vp [R, P, kmax], ziface [R, P, kmax-1], sigma [R, P]: CUDA tensors of the
    R local replicas x P proposals per step (config 4: 8 x 1024 per GPU).  Returns logL [R, P]."""
    from . import device
    R, P = k.shape
    out = device.dff_batch_device(voro_vp.reshape(R * P, -1), ziface.reshape(R * P, -1),
                                  k.reshape(R * P), src_offset, src_depth, tobs=tobs,
                                  sigma=sigma.reshape(R * P),
                                  logL=None if logL is None else logL.reshape(R * P), kmode=True)
    return out["logL"].reshape(R, P)


def mh_accept(logL_cur, logL_new, beta, u, logPr_new=None, outside=None):
    """The Metropolis-Hastings decision of EXPLORE_MH_NOVARPAR (prjmh_temper_rf.f90:739-751) for a
    batch of independent chains, one proposal each:

        logPLratio = logPr_new + (logL_new - logL_cur) * beta;   reject iff u >= exp(logPLratio)

    and proposals that left the prior bounds (CHECKBOUNDS2, `ioutside`) are rejected outright
    (:753-757).  Tensors of any matching shape on any device; returns a bool tensor (True =
    accept).  This is the rule only -- the reference's sampler applies it one proposal at a time
    per chain; batching proposals across chains keeps each chain's kernel unchanged."""
    ratio = (logL_new - logL_cur) * beta
    if logPr_new is not None:
        ratio = logPr_new + ratio
    accept = ~(u >= torch.exp(ratio))
    if outside is not None:
        accept = accept & ~outside.to(torch.bool)
    return accept


def tempering_swap_round_device(logL_local, beta_local, seed, round_index, group=None):
    """The swap round without leaving the device: all-gather, pairing, TEMPSWP_MH's accept rule
    (prjmh_temper_rf.f90:1341-1344) and the beta exchange are tensor operations on the device the
    replicas live on, so an MCMC step never synchronises with the host.  Every rank seeds an
    identical device generator from (seed, round_index) and therefore derives the same pairs and
    the same decisions.  Returns (new beta_local, info) where info holds device tensors
    `pairs_i`, `pairs_j`, `accept`."""
    dev = logL_local.device
    n_local = logL_local.numel()
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if distributed else 0
    mine = torch.stack([logL_local.to(torch.float64), beta_local.to(torch.float64)], dim=1).contiguous()
    if distributed:
        allr = torch.empty((dist.get_world_size(group) * n_local, 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine, group=group)
    else:
        allr = mine
    logL, beta = allr[:, 0], allr[:, 1]
    n = logL.numel()
    g = torch.Generator(device=dev)
    g.manual_seed((int(seed) * 1_000_003 + int(round_index)) & 0x7FFFFFFFFFFFFFFF)
    perm = torch.randperm(n, generator=g, device=dev)
    npair = n // 2
    i, j = perm[:npair], perm[npair:2 * npair]
    u = torch.rand(npair, generator=g, device=dev, dtype=torch.float64)
    accept = u <= torch.exp((beta[j] - beta[i]) * (logL[i] - logL[j]))
    new = beta.clone()
    new[i] = torch.where(accept, beta[j], beta[i])
    new[j] = torch.where(accept, beta[i], beta[j])
    return new[rank * n_local:(rank + 1) * n_local].clone(), {"pairs_i": i, "pairs_j": j, "accept": accept}


class LadderAdapter:
    """The burn-in adaptation of the temperature ladder (prjmh_temper_rf.f90:363-383 with the
    rolling window of TEMPSWP_MH, :1373-1379): swap proposals and acceptances are counted; once
    more than `acceptance_window` proposals have been seen the acceptance rate is latched and the
    counters reset; at that moment, during burn-in, dTlog grows by 2 % if the rate is below 0.2 and
    shrinks by 2 % if it is above 0.5, and the betas are reassigned as 1/dTlog**(it-1).

    `update(accept)` takes the accept mask of one swap round (any array-like of booleans) and
    returns the new ladder (numpy array, T = 1 chains first) when it changed, else None."""

    def __init__(self, n_replicas, dTlog, n_cold=1, acceptance_window=150):   # rjmcmc_com.f90:159
        self.n, self.dTlog, self.n_cold = int(n_replicas), float(dTlog), int(n_cold)
        self.window = int(acceptance_window)
        self.ncswap = self.ncswapprop = 0
        self.acceptance_rate = 0.25                         # :317

    def update(self, accept, burn_in=True):
        changed = False
        for a in np.asarray(accept, dtype=bool).ravel():
            self.ncswap += int(a)
            self.ncswapprop += 1
            if self.ncswapprop > self.window:               # :1375-1379
                self.acceptance_rate = self.ncswap / self.ncswapprop
                self.ncswap = self.ncswapprop = 0
                if burn_in:                                 # :363-372
                    if self.acceptance_rate < 0.2:
                        self.dTlog *= 1.02
                        changed = True
                    if self.acceptance_rate > 0.5:
                        self.dTlog *= 0.98
                        changed = True
        return temperature_ladder(self.n, self.dTlog, self.n_cold) if changed else None
