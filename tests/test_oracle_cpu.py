"""CPU tests: the oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py), and the Philox mirror's samplers against their target laws."""
import json
import os

import numpy as np
import pytest
from scipy import stats

from oracle import ref_numpy

SONY = dict(K=2.2881136684755243, g_scale=6.4508722699636545, sat=15583, ratio=208.9766365993794)
PARAMS = (SONY['K'], SONY['g_scale'], SONY['sat'], SONY['ratio'])


@pytest.fixture(scope='module')
def kat(golden_dir):
    with open(os.path.join(golden_dir, 'noise_kat.json')) as f:
        return json.load(f)


def test_sample_params_golden(kat):
    """noise.py:201-225 - RNG call order and arithmetic, all five cameras, seed 0."""
    for i, cam in enumerate(ref_numpy.CAMERAS):
        nm = ref_numpy.NoiseModelRef('P+g', include=i)
        np.random.seed(0)
        K, g, sat, ratio = nm._sample_params()
        gold = kat['sample_params_seed0'][cam]
        assert K == gold['K'] and g == gold['g_scale'] and sat == gold['sat'] and ratio == gold['ratio']


def test_sample_params_stream_golden(kat):
    nm = ref_numpy.NoiseModelRef('P+g', include=4)
    np.random.seed(2018)
    got = [[float(v) for v in nm._sample_params()] for _ in range(5)]
    assert got == kat['sample_params_seed2018_x5']
    nm = ref_numpy.NoiseModelRef('g')
    np.random.seed(11)
    got = [[float(v) for v in nm._sample_params()] for _ in range(4)]
    assert got == kat['sample_params_allcams_seed11_x4']


@pytest.mark.parametrize('model', ['P+g', 'p+g', 'g', 'P', 'p', 'Pg'])
def test_call_golden(kat, model):
    """noise.py:149-170 with numpy's own RNG: the restatement reproduces the reference bit for bit."""
    y = (np.arange(256).reshape(4, 8, 8).astype(np.float32)) / 256
    nm = ref_numpy.NoiseModelRef(model, include=4)
    np.random.seed(123)
    z = np.asarray(nm(y, params=PARAMS))
    gold = kat['call_seed123'][model]
    assert str(z.dtype) == gold['dtype']          # float64 under NumPy 2
    assert np.array_equal(z.astype(np.float64).ravel(), np.asarray(gold['z']))


def test_pack_golden(golden_dir, oracle):
    with open(os.path.join(golden_dir, 'pack_kat.json')) as f:
        g = json.load(f)
    m = np.asarray(g['mosaic_8x6'])
    assert np.array_equal(ref_numpy.pack_raw_bayer(m), np.asarray(g['packed'], np.float32))
    assert np.array_equal(ref_numpy.unpack_raw_bayer(ref_numpy.pack_raw_bayer(m)), m)
    assert np.array_equal(oracle.pack_bayer(m.astype(np.float32)), np.asarray(g['packed'], np.float32))
    m2 = np.asarray(g['mosaic_16x12'], np.uint16)
    assert np.array_equal(oracle.pack_bayer(m2), np.asarray(g['packed_16x12'], np.float32))
    # KAT quoted in SURVEY 8a row a-K
    p = ref_numpy.pack_raw_bayer(m)
    assert p[0, 0].tolist() == [0, 2, 4] and p[1, 0].tolist() == [1, 3, 5]
    assert p[2, 0].tolist() == [7, 9, 11] and p[3, 0].tolist() == [6, 8, 10]


def test_philox_kat(oracle):
    """Random123 known-answer vectors for Philox4x32-10."""
    assert oracle.philox([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize('lam', [0.05, 0.7, 3.0, 9.99, 10.0, 17.3, 33.0, 150.0, 1558.0])
def test_poisson_law(oracle, lam):
    """Philox-driven Poisson sampler vs the exact Poisson CDF (KS, 1% level) and moments."""
    n = 300000
    x = oracle.poisson_stream(lam, 7, 3, 0, n).astype(np.float64)
    assert (x == np.floor(x)).all() and x.min() >= 0
    ks = np.arange(0, x.max() + 2)
    emp = np.searchsorted(np.sort(x), ks, side='right') / n
    d = np.abs(emp - stats.poisson.cdf(ks, lam)).max()
    assert d < 1.63 / np.sqrt(n), (lam, d)
    assert abs(x.mean() - lam) < 5 * np.sqrt(lam / n)
    assert abs(x.var() - lam) < 6 * lam * np.sqrt(2.0 / n) + 6 * np.sqrt(lam / n)


def test_poisson_matches_numpy_distribution(oracle):
    """Two-sample check against numpy's own sampler (what the reference calls, noise.py:159)."""
    rs = np.random.RandomState(5)
    for lam in [2.5, 40.0]:
        a = oracle.poisson_stream(lam, 11, 0, 0, 200000)
        b = rs.poisson(lam, 200000)
        assert stats.ks_2samp(a, b).pvalue > 1e-3


def test_normal_law(oracle):
    x = oracle.normal_stream(1, 0, 0, 1, 0, 1000000).astype(np.float64)
    assert abs(x.mean()) < 5e-3 and abs(x.std() - 1) < 5e-3
    assert stats.kstest(x, 'norm').pvalue > 1e-3
    assert abs(stats.kurtosis(x)) < 0.02 and abs(stats.skew(x)) < 0.01
    # neighbouring pixels and different draw slots are uncorrelated
    y = oracle.normal_stream(1, 0, 0, 0, 0, 1000000).astype(np.float64)
    assert abs(np.corrcoef(x, y)[0, 1]) < 5e-3
    assert abs(np.corrcoef(x[:-1], x[1:])[0, 1]) < 5e-3


@pytest.mark.parametrize('lam', [-0.14285714, 0.0, 0.12857143])
def test_tukey_law(oracle, lam):
    """Tukey-lambda read noise vs scipy.stats.tukeylambda (what the paper's model uses)."""
    x = oracle.tukey_stream(lam, 3, 1, 0, 400000).astype(np.float64)
    assert stats.kstest(x, lambda v: stats.tukeylambda.cdf(v, lam)).pvalue > 1e-3


def test_formation_moments(oracle):
    """E[z] = y and Var[z] = (ratio/sat)^2 (K y_DN + g^2) for P+g and p+g (SURVEY 8c pin 4)."""
    n = 64
    y = np.full((n, 4, 16, 16), 0.35, np.float32)
    for mask in (0x05, 0x06):
        z = oracle.noise_packed(y, [PARAMS] * n, mask, 9, 100, False).astype(np.float64)
        ydn = 0.35 * SONY['sat'] / SONY['ratio']
        var = (SONY['ratio'] / SONY['sat']) ** 2 * (SONY['K'] * ydn + SONY['g_scale'] ** 2)
        assert abs(z.mean() - 0.35) < 5 * np.sqrt(var / z.size)
        assert abs(z.var() / var - 1) < 0.03


def test_formation_vs_reference_distribution(oracle):
    """Philox mirror vs the numpy-driven restatement of noise.py:149-170 on the same clean frame:
    per-pixel two-sample KS over many realisations collapses to a pooled KS on standardised residuals."""
    rs = np.random.RandomState(1)
    y = rs.rand(4, 32, 32).astype(np.float32)
    for model, mask in (('P+g', 0x05), ('p+g', 0x06), ('g', 0x04)):
        nm = ref_numpy.NoiseModelRef(model, include=4)
        np.random.seed(77)
        ref = np.stack([np.asarray(nm(y, params=PARAMS), np.float64) for _ in range(24)])
        mine = oracle.noise_packed(np.repeat(y[None], 24, 0), [PARAMS] * 24, mask, 5, 0, False).astype(np.float64)
        assert stats.ks_2samp((ref - y).ravel(), (mine - y).ravel()).pvalue > 1e-3
        assert abs(ref.var() / mine.var() - 1) < 0.03


def test_oracle_deterministic_and_frame_keyed(oracle):
    y = np.random.RandomState(0).rand(3, 4, 8, 12).astype(np.float32)
    a = oracle.noise_packed(y, [PARAMS] * 3, 0x05, 42, 10, True)
    b = oracle.noise_packed(y, [PARAMS] * 3, 0x05, 42, 10, True)
    assert np.array_equal(a, b)
    # frame 11 computed alone == frame index 1 of the batch starting at 10 (sharding invariance)
    c = oracle.noise_packed(y[1:2], [PARAMS], 0x05, 42, 11, True)
    assert np.array_equal(a[1:2], c)
    assert not np.array_equal(a[0], oracle.noise_packed(y[:1], [PARAMS], 0x05, 43, 10, True)[0])


def test_row_noise_structure(oracle):
    """Row noise: planes 0,1 share the even sensor row draw, planes 3,2 the odd one (noise.py:16-19)."""
    p = dict(K=1.0, ratio=100.0, saturation=100.0, R_scale=2.0)
    y = np.zeros((1, 4, 6, 8), np.float32)
    z = oracle.noise_packed(y, [p], 0x20, 1, 0, False)[0]
    assert np.allclose(z[0], z[0][:, :1]) and np.array_equal(z[0], z[1]) and np.array_equal(z[2], z[3])
    assert not np.array_equal(z[0], z[2])
    assert len(np.unique(z[0][:, 0])) == 6


def test_psnr_restatement():
    a = np.full((4, 8, 8), 0.5)
    b = a + 1.0 / 255
    assert abs(ref_numpy.psnr255(a, b) - 10 * np.log10(255 ** 2)) < 1e-9
