"""CPU ORACLE -- test infrastructure, NOT product code.

ctypes binding of oracle/liboracle_raymod.so (the plain-C restatement of the reference's
raymod + LOGLHOOD_RT; see raymod_oracle.h).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_raymod.so")
_lib = None

BRANCH_NAMES = {0: "top", 1: "neg", 2: "safe", 3: "bisect"}


class Trace(C.Structure):
    _fields_ = [("p", C.c_double), ("f_final", C.c_double), ("nl", C.c_int), ("branch", C.c_int),
                ("n_halve", C.c_int), ("n_bisect", C.c_int), ("n_newton", C.c_int),
                ("n_clamp", C.c_int), ("conv", C.c_int), ("n_f_ref", C.c_int),
                ("n_fp_ref", C.c_int), ("n_ffp_min", C.c_int), ("n_f_min", C.c_int)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in
                ("rays", "top", "neg", "safe", "bisect", "sum_nl", "n_halve", "n_bisect",
                 "n_newton", "n_clamp", "not_conv", "n_f_ref", "n_fp_ref", "n_ffp_min", "n_f_min")] \
        + [("flops_ref", C.c_double), ("flops_min", C.c_double)] \
        + [(n, C.c_longlong) for n in ("sqrt_ref", "div_ref", "sqrt_min", "div_min")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    """Compile the oracle with oracle/Makefile (gcc -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "raymod_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        _lib.orc_which_layer.restype = C.c_int
        _lib.orc_which_layer.argtypes = [dp, C.c_int, C.c_double]
        _lib.orc_ray_time.restype = C.c_double
        _lib.orc_ray_time.argtypes = [dp, dp, C.c_int, C.c_double, C.c_double, C.POINTER(Trace)]
        _lib.orc_trace_rays.restype = None
        _lib.orc_trace_rays.argtypes = [dp, dp, C.c_int, dp, dp, C.c_int, dp, dp,
                                        C.POINTER(Trace), C.c_int, C.c_char_p]
        _lib.orc_loglhood_from_times.restype = C.c_double
        _lib.orc_loglhood_from_times.argtypes = [dp, dp, C.c_int, C.c_double]
        _lib.orc_loglhood_rt.restype = C.c_double
        _lib.orc_loglhood_rt.argtypes = [C.c_int, dp, dp, dp, dp, C.c_int, dp, C.c_double, dp]
        _lib.orc_dff_batch.restype = C.c_int
        _lib.orc_dff_batch.argtypes = [dp, dp, ip, C.c_int, C.c_int, C.c_int, dp, dp, C.c_int,
                                       dp, dp, dp, dp, dp, C.c_int]
        _lib.orc_dff_batch_faithful.restype = C.c_int
        _lib.orc_dff_batch_faithful.argtypes = [dp, dp, ip, C.c_int, C.c_int, C.c_int, dp, dp,
                                                C.c_int, dp, C.c_char_p]
        _lib.orc_loglhood_voro.restype = C.c_double
        _lib.orc_loglhood_voro.argtypes = [C.c_int, dp, dp, dp, dp, C.c_int, dp, C.c_double, dp, dp, dp]
        _lib.orc_loglhood_from_times_ar.restype = C.c_double
        _lib.orc_loglhood_from_times_ar.argtypes = [dp, dp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double]
        _lib.orc_set_ismpprior.restype = None
        _lib.orc_set_ismpprior.argtypes = [C.c_int]
        _lib.orc_set_enos.restype = None
        _lib.orc_set_enos.argtypes = [C.c_int]
        _lib.orc_mh_step_batch.restype = None
        _lib.orc_mh_step_batch.argtypes = [ip, dp, dp, C.c_int, C.c_int, ip, ip, dp, dp, dp, dp, dp,
                                           dp, dp, C.c_int, dp, ip, dp, dp]
        _lib.orc_bd_step_batch.restype = None
        _lib.orc_bd_step_batch.argtypes = [ip, dp, dp, C.c_int, C.c_int, dp, ip, dp, dp, dp, dp, dp, dp,
                                           dp, C.c_int, C.c_int, dp, dp, C.c_int, dp, ip, ip, dp, dp]
        _lib.orc_sd_step_batch.restype = None
        _lib.orc_sd_step_batch.argtypes = [ip, dp, dp, dp, C.c_int, C.c_int, dp, dp, dp, dp, dp,
                                           dp, dp, C.c_int, dp, ip, dp]
        _lib.orc_ar_step_batch.restype = None
        _lib.orc_ar_step_batch.argtypes = [ip, dp, dp, dp, ip, dp, C.c_int, C.c_int, dp, dp, dp, dp, dp,
                                           dp, dp, dp, C.c_int, dp, ip, dp]
        _lib.orc_set_chain_ar.restype = None
        _lib.orc_set_chain_ar.argtypes = [ip, dp, C.c_double]
        _lib.orc_batch_stats.restype = None
        _lib.orc_batch_stats.argtypes = [dp, dp, ip, C.c_int, C.c_int, C.c_int, dp, dp, C.c_int,
                                         C.POINTER(Stats)]
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def which_layer(depths, dph):
    z = _d(depths)
    return lib().orc_which_layer(_p(z), z.size, float(dph))


def trace_rays(vels, depths, src_offset, src_depth, keep_delta=-1, rays_path=None, want_trace=False):
    """dff / TraceRays for one model.  Returns (timeP, p, traces or None)."""
    v, z, so, sd = _d(vels), _d(depths), _d(src_offset), _d(src_depth)
    assert v.size == z.size + 1 and so.size == sd.size
    n = so.size
    t, p = np.empty(n), np.empty(n)
    tr = (Trace * max(n, 1))()
    lib().orc_trace_rays(_p(v), _p(z), z.size, _p(so), _p(sd), n, _p(t), _p(p), tr,
                         int(keep_delta), rays_path.encode() if rays_path else None)
    return t, p, (list(tr)[:n] if want_trace else None)


def loglhood_from_times(tpred, tobs, sigma):
    a, b = _d(tpred), _d(tobs)
    return lib().orc_loglhood_from_times(_p(a), _p(b), a.size, float(sigma))


def loglhood_from_times_ar(tpred, tobs, sigma, idxar, arpar, armx=0.5):
    a, b = _d(tpred), _d(tobs)
    return lib().orc_loglhood_from_times_ar(_p(a), _p(b), a.size, float(sigma), int(idxar), float(arpar), float(armx))


def loglhood_rt(vp, ziface, src_offset, src_depth, tobs, sigma):
    """LOGLHOOD_RT for one objstruc-like model: k = len(vp) nodes.  Returns (logL, DpredRT)."""
    v, z = _d(vp), _d(ziface)
    so, sd, ob = _d(src_offset), _d(src_depth), _d(tobs)
    pred = np.empty(so.size)
    ll = lib().orc_loglhood_rt(v.size, _p(v), _p(z) if z.size else None, _p(so), _p(sd), so.size,
                               _p(ob), float(sigma), _p(pred))
    return ll, pred


def mh_step_batch(k, voro, logL, ivo, iwhich, cauchy, u_acc, beta, sigma, prior,
                  src_offset, src_depth, tobs):
    """One fixed-dimension MH move of B independent chains (orc_mh_step_batch).  voro [B, 2, ldk]
    and logL [B] are the current states; returns a dict with the updated copies, `accept` [B]
    (1 / 0 / -1 outside), the sorted proposals `voro_prop` and their `logL_prop` (NaN if outside)."""
    kk = np.ascontiguousarray(k, dtype=np.int32)
    vo = np.array(voro, dtype=np.float64, order="C", copy=True)
    ll = np.array(logL, dtype=np.float64, copy=True)
    B, two, ldk = vo.shape
    assert two == 2
    iv = np.ascontiguousarray(ivo, dtype=np.int32)
    iw = np.ascontiguousarray(iwhich, dtype=np.int32)
    so, sd, ob = _d(src_offset), _d(src_depth), _d(tobs)
    acc = np.zeros(B, dtype=np.int32)
    prop = vo.copy()
    llp = np.full(B, np.nan)
    I = C.POINTER(C.c_int)
    lib().orc_mh_step_batch(kk.ctypes.data_as(I), _p(vo), _p(ll), B, ldk, iv.ctypes.data_as(I),
                            iw.ctypes.data_as(I), _p(_d(cauchy)), _p(_d(u_acc)), _p(_d(beta)),
                            _p(_d(sigma)), _p(_d(prior)), _p(so), _p(sd), so.size, _p(ob),
                            acc.ctypes.data_as(I), _p(prop), _p(llp))
    return {"voro": vo, "logL": ll, "accept": acc, "voro_prop": prop, "logL_prop": llp}


def bd_step_batch(k, voro, logL, u_k, idel, u_z, u_v, u_acc, beta, sigma, prior, pk, kmin, kmax,
                  src_offset, src_depth, tobs):
    """The birth/death move of B independent chains (orc_bd_step_batch).  Returns a dict with the
    updated copies of k, voro, logL, `accept` [B] (1 / 0 / -1 outside / 2 no move), the proposals
    (`k_prop`, `voro_prop`, `logL_prop`)."""
    kk = np.array(k, dtype=np.int32, copy=True)
    vo = np.array(voro, dtype=np.float64, order="C", copy=True)
    ll = np.array(logL, dtype=np.float64, copy=True)
    B, two, ldk = vo.shape
    assert two == 2
    so, sd, ob = _d(src_offset), _d(src_depth), _d(tobs)
    acc = np.zeros(B, dtype=np.int32)
    kp = kk.copy()
    prop = vo.copy()
    llp = np.full(B, np.nan)
    I = C.POINTER(C.c_int)
    idl = np.ascontiguousarray(idel, dtype=np.int32)
    pkp = None if pk is None else _p(_d(pk))
    lib().orc_bd_step_batch(kk.ctypes.data_as(I), _p(vo), _p(ll), B, ldk, _p(_d(u_k)),
                            idl.ctypes.data_as(I), _p(_d(u_z)), _p(_d(u_v)), _p(_d(u_acc)),
                            _p(_d(beta)), _p(_d(sigma)), _p(_d(prior)), pkp, int(kmin), int(kmax),
                            _p(so), _p(sd), so.size, _p(ob), acc.ctypes.data_as(I),
                            kp.ctypes.data_as(I), _p(prop), _p(llp))
    return {"k": kk, "voro": vo, "logL": ll, "accept": acc, "k_prop": kp, "voro_prop": prop,
            "logL_prop": llp}


def sd_step_batch(k, voro, logL, sigma, u_gate, gauss, u_acc, beta, sd_prior, src_offset, src_depth, tobs):
    """The data-error move of B independent chains (orc_sd_step_batch).  Returns a dict with the
    updated copies of logL and sigma, `accept` [B] (1 / 0 / -1 outside / 2 no move), `logL_prop`."""
    kk = np.ascontiguousarray(k, dtype=np.int32)
    vo = _d(voro)
    ll = np.array(logL, dtype=np.float64, copy=True)
    sg = np.array(sigma, dtype=np.float64, copy=True)
    B, two, ldk = vo.shape
    so, sd, ob = _d(src_offset), _d(src_depth), _d(tobs)
    acc = np.zeros(B, dtype=np.int32)
    llp = np.full(B, np.nan)
    I = C.POINTER(C.c_int)
    lib().orc_sd_step_batch(kk.ctypes.data_as(I), _p(vo), _p(ll), _p(sg), B, ldk, _p(_d(u_gate)),
                            _p(_d(gauss)), _p(_d(u_acc)), _p(_d(beta)), _p(_d(sd_prior)), _p(so), _p(sd),
                            so.size, _p(ob), acc.ctypes.data_as(I), _p(llp))
    return {"logL": ll, "sigma": sg, "accept": acc, "logL_prop": llp}


def ar_step_batch(k, voro, logL, sigma, idxar, arpar, u_choice, u_prop, gauss, u_acc, beta, ar_prior,
                  src_offset, src_depth, tobs):
    """The AR(1) move of B independent chains (orc_ar_step_batch).  Returns a dict with the updated
    copies of logL, idxar, arpar, `accept` [B] (1 / 0 / -1 outside) and `logL_prop`."""
    kk = np.ascontiguousarray(k, dtype=np.int32)
    vo = _d(voro)
    ll = np.array(logL, dtype=np.float64, copy=True)
    ia = np.array(idxar, dtype=np.int32, copy=True)
    ap = np.array(arpar, dtype=np.float64, copy=True)
    B, two, ldk = vo.shape
    so, sd, ob = _d(src_offset), _d(src_depth), _d(tobs)
    acc = np.zeros(B, dtype=np.int32)
    llp = np.full(B, np.nan)
    I = C.POINTER(C.c_int)
    lib().orc_ar_step_batch(kk.ctypes.data_as(I), _p(vo), _p(ll), _p(_d(sigma)), ia.ctypes.data_as(I), _p(ap),
                            B, ldk, _p(_d(u_choice)), _p(_d(u_prop)), _p(_d(gauss)), _p(_d(u_acc)),
                            _p(_d(beta)), _p(_d(ar_prior)), _p(so), _p(sd), so.size, _p(ob),
                            acc.ctypes.data_as(I), _p(llp))
    return {"logL": ll, "idxar": ia, "arpar": ap, "accept": acc, "logL_prop": llp}


_chain_ar_keep = None


def set_chain_ar(idxar=None, arpar=None, armx=0.5):
    """IAR = 1 for the *_step_batch functions: the chains' idxarRT / arparRT (copied and kept alive
    here) enter every likelihood evaluation; call without arguments to return to IAR = 0."""
    global _chain_ar_keep
    if idxar is None:
        _chain_ar_keep = None
        lib().orc_set_chain_ar(None, None, float(armx))
        return
    ia = np.ascontiguousarray(idxar, dtype=np.int32).copy()
    ap = _d(arpar).copy()
    _chain_ar_keep = (ia, ap)
    lib().orc_set_chain_ar(ia.ctypes.data_as(C.POINTER(C.c_int)), _p(ap), float(armx))


def loglhood_voro(node_depth, node_vp, src_offset, src_depth, tobs, sigma):
    """INTERPLAYER_novar + LOGLHOOD_RT on unsorted Voronoi nodes.
    Returns (logL, DpredRT, sorted_depth, sorted_vp)."""
    d, v = _d(node_depth), _d(node_vp)
    so, sd, ob = _d(src_offset), _d(src_depth), _d(tobs)
    pred, sdp, svp = np.empty(so.size), np.empty(d.size), np.empty(d.size)
    ll = lib().orc_loglhood_voro(d.size, _p(d), _p(v), _p(so), _p(sd), so.size, _p(ob), float(sigma),
                                 _p(pred), _p(sdp), _p(svp))
    return ll, pred, sdp, svp


def dff_batch(vels, depths, nlayers, src_offset, src_depth, tobs=None, sigma=None,
              want_times=True, want_p=False, nthreads=0):
    """Batched sweep.  vels[B, ldv], depths[B, ldz], nlayers[B].  Returns dict + threads used."""
    v, z = _d(vels), _d(depths)
    nl = np.ascontiguousarray(nlayers, dtype=np.int32)
    so, sd = _d(src_offset), _d(src_depth)
    B, n = v.shape[0], so.size
    t = np.empty((B, n)) if want_times else None
    p = np.empty((B, n)) if want_p else None
    ll = ob = sg = None
    if tobs is not None:
        ob, sg, ll = _d(tobs), _d(sigma), np.empty(B)
    used = lib().orc_dff_batch(_p(v), _p(z), nl.ctypes.data_as(C.POINTER(C.c_int)), B,
                               v.shape[1], z.shape[1], _p(so), _p(sd), n, _p(t), _p(ob), _p(sg),
                               _p(ll), _p(p), int(nthreads))
    return {"timeP": t, "p": p, "logL": ll, "threads": used}


def dff_batch_faithful(vels, depths, nlayers, src_offset, src_depth, rays_path):
    v, z = _d(vels), _d(depths)
    nl = np.ascontiguousarray(nlayers, dtype=np.int32)
    so, sd = _d(src_offset), _d(src_depth)
    B, n = v.shape[0], so.size
    t = np.empty((B, n))
    lib().orc_dff_batch_faithful(_p(v), _p(z), nl.ctypes.data_as(C.POINTER(C.c_int)), B,
                                 v.shape[1], z.shape[1], _p(so), _p(sd), n, _p(t),
                                 rays_path.encode())
    return t


def batch_stats(vels, depths, nlayers, src_offset, src_depth):
    v, z = _d(vels), _d(depths)
    nl = np.ascontiguousarray(nlayers, dtype=np.int32)
    so, sd = _d(src_offset), _d(src_depth)
    st = Stats()
    lib().orc_batch_stats(_p(v), _p(z), nl.ctypes.data_as(C.POINTER(C.c_int)), v.shape[0],
                          v.shape[1], z.shape[1], _p(so), _p(sd), so.size, C.byref(st))
    return st.as_dict()


def set_enos(enos):
    """ENOS switch of the move oracles (orc_set_enos): 1 = even-numbered order statistics prior."""
    lib().orc_set_enos(1 if enos else 0)


def set_ismpprior(on):
    """ISMPPRIOR switch of the move oracles (orc_set_ismpprior): 1 = LOGLHOOD2, logL = 1."""
    lib().orc_set_ismpprior(1 if on else 0)
