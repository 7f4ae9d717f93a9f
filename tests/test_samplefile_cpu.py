"""The sampler's sample-file format (prjmh_temper_rf.f90:1880-1912, FORMAT(500ES18.8)) -- "next"
row N3.  No sample file ships with the reference (git-ignored there), so the format is pinned by
the edit descriptor's rules and by a write/read round trip."""
import numpy as np

from raytracerfortran_b200 import samplefile, workloads


def test_es18_8_edit_descriptor():
    assert samplefile.es18_8(1.0) == "    1.00000000E+00"
    assert samplefile.es18_8(-55.12412980548659) == "   -5.51241298E+01"
    assert samplefile.es18_8(0.0) == "    0.00000000E+00"
    assert samplefile.es18_8(0.016) == "    1.60000000E-02"
    assert samplefile.es18_8(-100.0) == "   -1.00000000E+02"
    assert samplefile.es18_8(-1.7976931348623157e308) == "   -1.79769313+308"   # -HUGE: E is dropped
    assert all(len(samplefile.es18_8(v)) == 18 for v in (1e-300, 3.3e99, -2.5e-100, 7.0))


def test_row_layout_of_the_shipped_example():
    # test_1_parameter.dat: NLMX = 10, NPL = 2, NMODE = 1  ->  4 + 20 + 3 + 6 columns
    assert samplefile.row_width(10, 2, 1) == 33


def test_write_read_round_trip(tmp_path):
    B, nlmx = 40, 10
    k, vp, zi = workloads.make_transd_models(B, nlmx, 12, uniform_k=True)
    voro = np.zeros((B, 2, nlmx))
    voro[:, 1, :] = vp
    voro[:, 0, 1:] = zi
    rng = np.random.default_rng(1)
    logL, sig = rng.normal(50, 5, B), rng.uniform(0.001, 0.07, B)
    rows = samplefile.pack_rows(logL, rng.normal(0, 1, B), rng.uniform(0, 1e-3, B), k, voro, sig,
                                acc=rng.random(B), counters=rng.integers(0, 99, (B, 3)),
                                ic=np.ones(B), rank=np.arange(B) % 4)
    path = tmp_path / "x_voro_sample.txt"
    samplefile.write_samples(path, rows[:25])
    samplefile.write_samples(path, rows[25:], append=True)
    text = open(path).read().splitlines()
    assert len(text) == B and all(len(l) == 18 * 33 for l in text)
    smp = samplefile.read_samples(path, nlmx)
    assert np.array_equal(smp["k"], k)
    assert np.allclose(smp["rows"], rows, rtol=5e-9, atol=0)            # nine significant digits
    for b in range(B):
        assert np.allclose(smp["voro"][b, :, :k[b]], voro[b, :, :k[b]], rtol=5e-9)
        assert np.all(smp["voro"][b, :, k[b]:] == 0.0)
    thinned = samplefile.read_samples(path, nlmx, burnin=10, thin=3)
    assert np.array_equal(thinned["rank"], smp["rank"][10::3])
