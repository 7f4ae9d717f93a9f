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
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if logL_all.size != world * n_local:
        raise ValueError("tempering_swap_round needs the same number of replicas on every rank")
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


class SwapRound:
    """The swap round of parallel tempering with the decisions taken on the device
    (rtb200_swap_pack_device -> all-gather of 16 bytes per chain -> rtb200_swap_round_device;
    TEMPSWP_MH, prjmh_temper_rf.f90:1329-1384).

    One object per set of local chains: it owns the packed / gathered buffers and a side stream.
    `launch` enqueues the round on the side stream behind whatever the caller's stream has done so
    far and returns at once; `wait` makes the caller's stream wait for the new betas.  Between the
    two the caller can enqueue work that does not read beta -- the next step's proposal and
    likelihood kernels -- so the collective and the swap kernel run under it
    ("overlap it with the next step's proposal generation").  Every rank derives the
    same pairing and decisions from (seed, round_index); CUDA tensors only (no CPU path)."""

    def __init__(self, n_local, device, group=None, want_info=False):
        self.group = group
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.n_local, self.n = int(n_local), int(n_local) * self.world
        if self.distributed:                       # every rank must hold the same number of chains
            cnt = torch.tensor([self.n_local], dtype=torch.int64, device=device)
            lo, hi = cnt.clone(), cnt.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
            if int(lo) != int(hi):
                raise ValueError("SwapRound needs the same number of chains on every rank")
        f64 = torch.float64
        self.mine = torch.empty((self.n_local, 2), dtype=f64, device=device)
        self.all = torch.empty((self.n, 2), dtype=f64, device=device) if self.distributed else self.mine
        self.accept = torch.empty((max(self.n // 2, 1),), dtype=torch.int32, device=device) if want_info else None
        self.partner = torch.empty((self.n_local,), dtype=torch.int32, device=device) if want_info else None
        self.side = torch.cuda.Stream(device=device)
        self.ready = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.device = device

    def launch(self, logL_local, beta_local, seed, round_index, beta_out=None, stream=None):
        """Enqueue one round.  beta_out (default: beta_local, in place) receives the new betas."""
        from . import _lib
        from .device import _ensure_device, _ptr
        dev = self.device
        _ensure_device(dev.index if dev.index is not None else torch.cuda.current_device())
        f64 = torch.float64
        if beta_out is None:
            beta_out = beta_local
        cur = stream if stream is not None else torch.cuda.current_stream(dev)
        self.ready.record(cur)
        lib = _lib.load()
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ready)
            h = self.side.cuda_stream
            _lib.check(lib.rtb200_swap_pack_device(_ptr(logL_local, f64), _ptr(beta_local, f64),
                                                   self.n_local, _ptr(self.mine, f64), h))
            if self.distributed:
                dist.all_gather_into_tensor(self.all, self.mine, group=self.group)
            _lib.check(lib.rtb200_swap_round_device(
                _ptr(self.all, f64), self.n, self.rank * self.n_local, self.n_local,
                int(seed) & 0xFFFFFFFFFFFFFFFF, int(round_index) & 0xFFFFFFFFFFFFFFFF,
                _ptr(beta_out, f64), _ptr(self.accept, torch.int32), _ptr(self.partner, torch.int32), h))
            self.done.record(self.side)
        return self.done

    def wait(self, stream=None):
        (stream if stream is not None else torch.cuda.current_stream(self.device)).wait_event(self.done)


_swap_rounds = {}


def tempering_swap_round_device(logL_local, beta_local, seed, round_index, group=None):
    """One swap round without leaving the device, in stream order (launch + wait of a cached
    SwapRound).  Returns (new beta_local, info) with info = {"accept": [n/2] i32 per pair,
    "partner": [n_local] i32 (see rtb200_swap_round_device)} as device tensors."""
    if not logL_local.is_cuda:
        raise ValueError("tempering_swap_round_device needs CUDA tensors (there is no CPU path); "
                         "tempering_swap_round takes the decisions on the host")
    key = (logL_local.device, logL_local.numel(), id(group))
    sr = _swap_rounds.get(key)
    if sr is None:
        sr = _swap_rounds[key] = SwapRound(logL_local.numel(), logL_local.device, group, want_info=True)
    out = torch.empty_like(beta_local)
    sr.launch(logL_local, beta_local, seed, round_index, beta_out=out)
    sr.wait()
    return out, {"accept": sr.accept, "partner": sr.partner}


def assign_ladder(beta_current, ladder):
    """Hand a new ladder (LadderAdapter.update: ordered T = 1 first) to chains whose betas have
    been exchanged by swap rounds: the chain holding the r-th largest beta gets the r-th largest
    new one, so every state keeps its place in the temperature order -- the reference reassigns
    beta_pt by chain slot and its swaps move states, not temperatures
    (prjmh_temper_rf.f90:373-383, :1351-1357).  Works on numpy arrays or torch tensors."""
    if isinstance(beta_current, torch.Tensor):
        lad = torch.as_tensor(np.sort(np.asarray(ladder, dtype=np.float64))[::-1].copy(),
                              device=beta_current.device)
        order = torch.argsort(beta_current, descending=True, stable=True)
        out = torch.empty_like(beta_current)
        out[order] = lad
        return out
    cur = np.asarray(beta_current, dtype=np.float64)
    order = np.argsort(-cur, kind="stable")
    out = np.empty_like(cur)
    out[order] = np.sort(np.asarray(ladder, dtype=np.float64))[::-1]
    return out


class LadderAdapter:
    """The burn-in adaptation of the temperature ladder (prjmh_temper_rf.f90:363-383 with the
    rolling window of TEMPSWP_MH, :1373-1379): swap proposals and acceptances are counted; once
    more than `acceptance_window` proposals have been seen the acceptance rate is latched and the
    counters reset; at that moment, during burn-in, dTlog grows by 2 % if the rate is below 0.2 and
    shrinks by 2 % if it is above 0.5, and the betas are reassigned as 1/dTlog**(it-1).

    `update(accept)` takes the accept mask of one swap round (any array-like of booleans) and
    returns the new ladder (numpy array, T = 1 chains first: ordered by ladder index, NOT by
    replica -- hand it to the replicas with `assign_ladder`) when it changed, else None."""

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
