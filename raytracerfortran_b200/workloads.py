"""Seeded synthetic workloads of the benchmark shapes.

Velocities follow the reference's prior U[1500, 10000] m/s (read_input.f90:207-208); sources
follow the generator of raytracerR-export-data-to-MCMC.Rmd:46 (depth U[1050, 4200],
offset = sqrt(x^2 + y^2), x ~ U[10, 6500], y ~ U[10, 4500]); interfaces are sorted uniforms
with the minimum thickness hmin = 100.1 of test_1_parameter.dat:27-28 enforced by
construction.  numpy's PCG64 (default_rng) with the given seed.
"""
import numpy as np

HMIN, HMX = 100.1, 10000.1
VMIN, VMAX = 1500.0, 10000.0


def make_sources(n, seed, near_critical=False):
    rng = np.random.default_rng(seed + 7919)
    if near_critical:   # config-5 style: deep sources, long offsets, p*v -> 1
        d = rng.uniform(1050.0, 9900.0, n)
        x, y = rng.uniform(10.0, 30000.0, n), rng.uniform(10.0, 30000.0, n)
    else:
        d = rng.uniform(1050.0, 4200.0, n)
        x, y = rng.uniform(10.0, 6500.0, n), rng.uniform(10.0, 4500.0, n)
    return np.sqrt(x * x + y * y), d


def make_models(B, nlayers, seed, min_thickness=True):
    """B models with `nlayers` interfaces each.  Returns vels[B, nlayers+1], depths[B, nlayers],
    nl[B] (int32)."""
    rng = np.random.default_rng(seed)
    v = rng.uniform(VMIN, VMAX, (B, nlayers + 1))
    u = np.sort(rng.random((B, nlayers)), axis=1)
    if min_thickness:
        free = HMX - HMIN * (nlayers + 1)
        assert free > 0
        z = HMIN * np.arange(1, nlayers + 1)[None, :] + free * u
    else:
        z = 50.0 + (10000.0 - 50.0) * u
    return v, z, np.full(B, nlayers, dtype=np.int32)


def make_transd_models(B, kmax, seed, lam=3.01, uniform_k=False):
    """Trans-dimensional chain states (config 3): k nodes per model, 1 <= k <= kmax, drawn
    from a truncated Poisson(lam) (test_1_parameter.dat:26) or uniformly.  Returns
    k[B] int32, voro_vp[B, kmax], ziface[B, kmax-1]; entries beyond k are zero."""
    rng = np.random.default_rng(seed)
    if uniform_k:
        k = rng.integers(1, kmax + 1, B)
    else:
        k = rng.poisson(lam, B)
        bad = (k < 1) | (k > kmax)
        while bad.any():
            k[bad] = rng.poisson(lam, int(bad.sum()))
            bad = (k < 1) | (k > kmax)
    vp = rng.uniform(VMIN, VMAX, (B, kmax))
    u = rng.random((B, max(kmax - 1, 1)))
    col = np.arange(max(kmax - 1, 1))[None, :]
    valid = col < (k[:, None] - 1)
    u = np.where(valid, u, np.inf)
    u = np.sort(u, axis=1)
    free = HMX - HMIN * kmax
    z = HMIN * (col + 1) + free * u
    z = np.where(valid, z, 0.0)
    vp = np.where(np.arange(kmax)[None, :] < k[:, None], vp, 0.0)
    return k.astype(np.int32), vp, z[:, :max(kmax - 1, 1)]


def make_observations(t_true, B, seed, noise_sd=0.016, sd_range=(0.001, 0.07)):
    """tobs = T(model 0) + N(0, 0.016^2) (Rmd:89-90); sigma ~ U[sdmn, sdmx] per model
    (test_1_parameter.dat:30-31)."""
    rng = np.random.default_rng(seed + 104729)
    tobs = np.asarray(t_true, dtype=np.float64) + rng.normal(0.0, noise_sd, len(t_true))
    sigma = rng.uniform(sd_range[0], sd_range[1], B)
    return tobs, sigma


CONFIGS = {
    # name: (models, interfaces, sources, near_critical, seed)
    "config2": dict(B=1_000_000, nlayers=10, nsrc=64, near_critical=False, seed=2),
    "config3": dict(B=4096, kmax=30, nsrc=256, near_critical=False, seed=3),
    "config4": dict(replicas=64, proposals=1024, kmax=30, nsrc=256, near_critical=False, seed=4),
    "config5": dict(B=16_000_000, nlayers=50, nsrc=1024, near_critical=True, seed=5),
}
