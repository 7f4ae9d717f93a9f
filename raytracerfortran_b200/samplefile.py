"""The sampler's on-disk sample format and the posterior re-evaluation sweep ("next" row N3).

The reference writes one row per kept chain state to `<base>_voro_sample.txt` with
`FORMAT(500ES18.8)` (prjmh_temper_rf.f90:1880-1912, `207 FORMAT(500ES18.8)`):

    logL, logPr, tcmp, k, tmpvoro(NLMX*NPL), sdparRT(NMODE), arpar(NMODE), arparRT(NMODE),
    acceptance rate, iaccept_bd, ireject_bd, iaccept_bds, ic, rank

where tmpvoro holds the Voronoi nodes node-major, (depth, vp) per node, `-100` for parameters
that are switched off and `0` beyond node k (prjmh_temper_rf.f90:1859-1872).  `replica.f90:173-232`
reads such a file back, thins it, rebuilds each state (INTERPLAYER_novar) and re-evaluates
LOGLHOOD one state at a time; `reduce_sample.f90` drops the burn-in and keeps every NSUB-th row.

Here the file is parsed into a batch and the whole sweep is one `loglhood_batch_voro` call.
"""
import numpy as np

from . import raymod

_LEAD = 4      # logL, logPr, tcmp, k
_TRAIL = 6     # acceptance rate, iaccept_bd, ireject_bd, iaccept_bds, ic, rank


def row_width(NLMX, NPL=2, NMODE=1):
    return _LEAD + NLMX * NPL + 3 * NMODE + _TRAIL


def es18_8(x):
    """One value in Fortran's ES18.8 edit descriptor (exponents beyond two digits drop the E)."""
    s = f"{float(x):.8E}"
    mant, exp = s.split("E")
    e = int(exp)
    if abs(e) > 99:
        s = f"{mant}{'+' if e >= 0 else '-'}{abs(e):03d}"
    return s.rjust(18)


def format_row(values):
    """A row as WRITE(usample, '(500ES18.8)') prints it."""
    return "".join(es18_8(v) for v in values)


def pack_rows(logL, logPr, tcmp, k, voro, sdparRT, arpar=None, arparRT=None, acc=None,
              counters=None, ic=None, rank=None):
    """Assemble sample rows.  voro [B, 2, NLMX] (depth row, vp row); entries beyond k are 0."""
    voro = np.asarray(voro, dtype=np.float64)
    B, npl, nlmx = voro.shape
    sd = np.asarray(sdparRT, dtype=np.float64).reshape(B, -1)
    nmode = sd.shape[1]
    z = lambda a, w: np.zeros((B, w)) if a is None else np.asarray(a, dtype=np.float64).reshape(B, w)
    kk = np.asarray(k).reshape(B)
    tmpvoro = np.transpose(voro, (0, 2, 1)).copy()             # node-major (depth, vp) pairs
    tmpvoro[np.arange(nlmx)[None, :] >= kk[:, None]] = 0.0     # tmpvoro = 0 beyond node k
    return np.concatenate([
        np.asarray(logL, dtype=np.float64).reshape(B, 1), z(logPr, 1), z(tcmp, 1),
        kk.astype(np.float64).reshape(B, 1), tmpvoro.reshape(B, nlmx * npl), sd,
        z(arpar, nmode), z(arparRT, nmode), z(acc, 1), z(counters, 3), z(ic, 1), z(rank, 1)], axis=1)


def write_samples(path, rows, append=False):
    with open(path, "a" if append else "w") as fh:
        for r in np.atleast_2d(rows):
            fh.write(format_row(r) + "\n")


def read_samples(path, NLMX, NPL=2, NMODE=1, burnin=0, thin=1):
    """Parse a sample file.  `burnin` rows are dropped and every `thin`-th row kept, as
    reduce_sample.f90 / replica.f90:173-181 do.  Returns a dict of arrays; `voro` is
    [B, NPL, NLMX] with switched-off parameters (-100) reported in `voroidx` == 0."""
    width = row_width(NLMX, NPL, NMODE)
    rows = []
    with open(path) as fh:
        for line in fh:
            if not line.strip():
                continue
            body = line.rstrip("\n")
            vals = [_parse(body[i:i + 18]) for i in range(0, len(body), 18)]
            if len(vals) != width:
                raise ValueError(f"{path}: expected {width} columns, found {len(vals)}")
            rows.append(vals)
    a = np.array(rows, dtype=np.float64).reshape(-1, width)[burnin::max(1, thin)]
    o = _LEAD
    tmpvoro = a[:, o:o + NLMX * NPL].reshape(-1, NLMX, NPL)
    o += NLMX * NPL
    out = {
        "logL": a[:, 0], "logPr": a[:, 1], "tcmp": a[:, 2], "k": a[:, 3].astype(np.int32),
        "voro": np.transpose(tmpvoro, (0, 2, 1)).copy(),
        "voroidx": (np.transpose(tmpvoro, (0, 2, 1)) >= -99.0).astype(np.int32),   # replica.f90:196-202
        "sdparRT": a[:, o:o + NMODE], "arpar": a[:, o + NMODE:o + 2 * NMODE],
        "arparRT": a[:, o + 2 * NMODE:o + 3 * NMODE],
    }
    o += 3 * NMODE
    out.update({"acc": a[:, o], "iaccept_bd": a[:, o + 1], "ireject_bd": a[:, o + 2],
                "iaccept_bds": a[:, o + 3], "ic": a[:, o + 4], "rank": a[:, o + 5], "rows": a})
    return out


def _parse(tok):
    """One 18-column field; `1.00000000+100` is Fortran's three-digit-exponent form without E."""
    t = tok.strip().replace("D", "E")
    try:
        return float(t)
    except ValueError:
        for i in range(len(t) - 1, 0, -1):
            if t[i] in "+-":
                return float(t[:i] + "E" + t[i:])
        raise


def replica_sweep(path, NLMX, src_offset, src_depth, DobsRT, NPL=2, NMODE=1, burnin=0, thin=1):
    """replica.f90:173-232 as one batched call: every kept sample is rebuilt (INTERPLAYER_novar,
    sort by depth on the device) and LOGLHOOD re-evaluated with the sample's own sigma.
    Returns (samples dict, recomputed logL [B], DpredRT [B, NSRC])."""
    smp = read_samples(path, NLMX, NPL, NMODE, burnin, thin)
    if NPL != 2:
        raise ValueError("the travel-time path uses NPL = 2 (depth, vp) nodes")
    logL, pred, _ = raymod.loglhood_batch_voro(smp["k"], smp["voro"], src_offset, src_depth, DobsRT,
                                               smp["sdparRT"][:, 0], want_pred=True)
    return smp, logL, pred
