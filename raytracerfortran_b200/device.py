"""The ray-tracing path on tensors that already live in HBM.

torch is used for device memory and streams only; the work is done by the sm_100a kernels in
libraytrace_b200.so through rtb200_dff_batch_device (include/raytrace_b200.h).
"""
import torch

from . import _lib

_inited_device = None


def _ensure_device(index):
    global _inited_device
    if _inited_device != index:
        rc = _lib.load().rtb200_init(int(index))
        _lib.check(rc)
        _inited_device = index


def _ptr(t, dtype):
    if t is None:
        return None
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"expected a contiguous CUDA tensor of {dtype}, got {t.dtype} on {t.device}")
    return t.data_ptr()


def dff_batch_device(vels, depths, nlayers, src_offset, src_depth, tobs=None, sigma=None,
                     want_times=False, want_p=False, timeP=None, logL=None, p_out=None,
                     kmode=False, stream=None):
    """B models x NSrc sources on the GPU; every argument is a CUDA tensor.

    vels [B, ldv] f64, depths [B, ldz] f64, nlayers [B] i32, src_offset/src_depth [NSrc] f64,
    tobs [NSrc] f64 and sigma [B] f64 (both needed for logL).  Asynchronous on `stream`
    (default: torch's current stream).  Returns {"timeP", "p", "logL"} tensors (or None)."""
    dev = vels.device
    if not vels.is_cuda:
        raise ValueError("dff_batch_device needs CUDA tensors (there is no CPU path)")
    _ensure_device(dev.index if dev.index is not None else torch.cuda.current_device())
    B, ldv = vels.shape
    ldz = depths.shape[1] if depths.dim() == 2 else 0
    nsrc = src_offset.numel()
    f64 = torch.float64
    if timeP is None and want_times:
        timeP = torch.empty((B, nsrc), dtype=f64, device=dev)
    if p_out is None and want_p:
        p_out = torch.empty((B, nsrc), dtype=f64, device=dev)
    if tobs is not None and logL is None:
        logL = torch.empty((B,), dtype=f64, device=dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    rc = _lib.load().rtb200_dff_batch_device(
        _ptr(vels, f64), _ptr(depths, f64) if ldz else None, _ptr(nlayers, torch.int32), B, ldv, ldz,
        _ptr(src_offset, f64), _ptr(src_depth, f64), nsrc, _ptr(timeP, f64), _ptr(tobs, f64),
        _ptr(sigma, f64), _ptr(logL, f64), _ptr(p_out, f64), 1 if kmode else 0,
        st.cuda_stream if st.cuda_stream != 0 else _legacy_stream_handle())
    _lib.check(rc)
    return {"timeP": timeP, "p": p_out, "logL": logL}


def _legacy_stream_handle():
    # torch's default stream is the legacy stream (handle 0).  The C ABI reads NULL as "use the
    # library's own stream and synchronise", so name the legacy stream explicitly instead.
    return 1  # cudaStreamLegacy
