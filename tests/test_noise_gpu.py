"""GPU parity tests of the fused noise kernel (through the C ABI) against the CPU oracle.

Stated tolerances
  * Bayer pack / de-quantise (integer index map): BIT-EXACT.
  * Gaussian-type terms: |z_gpu - z_oracle| <= ATOL_SIGMA * sigma_pixel + 1e-6, where sigma_pixel is
    the total noise std of that pixel in output units.  The kernel uses MUFU approximations
    (lg2/sin/cos/sqrt.approx) where the oracle uses libm; ATOL_SIGMA = 1e-3 bounds the worst
    case (log of a uniform within 1e-6 of 1), the MEAN abs error must be <= 2e-5 sigma.
  * Poisson: integer-valued draws; a 1-ulp transcendental difference can flip an accept/reject
    decision, so at most MISMATCH_FRAC = 2e-4 of the pixels may differ from the oracle (each of
    those still being a valid integer count); all others obey the Gaussian-type tolerance.
"""
import json
import os

import numpy as np
import pytest

from tests.conftest import has_gpu

pytestmark = pytest.mark.gpu

ATOL_SIGMA = 1e-3
MEAN_SIGMA = 2e-5
MISMATCH_FRAC = 2e-4
SONY = (2.2881136684755243, 6.4508722699636545, 15583, 208.9766365993794)
FULL = dict(K=2.2881136684755243, g_scale=6.4508722699636545, G_scale=3.1, G_lambda=-0.0857, R_scale=0.9,
            q_step=1.0, saturation=15583.0, ratio=208.9766365993794, color_bias=(0.5, -1.25, 2.0, 0.25))


@pytest.fixture(scope='module')
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no GPU')
    return torch


def _nm(model):
    from eld_b200.noise import NoiseModel
    return NoiseModel(model, include=4, verbose=False, seed=1234)


def _sigma(y, p, mask):
    if isinstance(p, dict):
        K, g, sat, ratio = p['K'], p['g_scale'], p['saturation'], p['ratio']
        extra = (p['G_scale'] * 2.0) ** 2 * bool(mask & 8) + p['R_scale'] ** 2 * bool(mask & 0x20) + \
            p['q_step'] ** 2 / 12 * bool(mask & 0x40)
    else:
        K, g, sat, ratio = p
        extra = 0.0
    ydn = np.maximum(y, 0) * sat / ratio
    var = (K * ydn if mask & 3 else 0) + (g * g if mask & 4 else 0) + extra
    return np.broadcast_to(np.sqrt(var + 1e-12) * ratio / sat, y.shape)


def _compare(gpu, ref, y, p, mask):
    sig = _sigma(y, p, mask)
    err = np.abs(gpu.astype(np.float64) - ref.astype(np.float64))
    bad = err > ATOL_SIGMA * sig + 1e-6
    if mask & 1:
        assert bad.mean() <= MISMATCH_FRAC, 'Poisson mismatch fraction %g' % bad.mean()
        ok = ~bad
    else:
        assert not bad.any(), 'max err/sigma %g' % (err / (sig + 1e-12)).max()
        ok = slice(None)
    assert (err[ok] / (sig[ok] + 1e-9)).mean() <= MEAN_SIGMA


MODELS = [('g', 0x04), ('p+g', 0x06), ('P+g', 0x05), ('P', 0x01), ('p', 0x02),
          ('ELD:P+G+R+U', 0x69), ('ELD:P+G+B+R+U', 0x79), ('ELD:g+R', 0x24), ('ELD:G', 0x08), ('ELD:U', 0x40)]


@pytest.mark.parametrize('model,mask', MODELS)
@pytest.mark.parametrize('shape', [(1, 4, 8, 8), (3, 4, 64, 64), (2, 4, 6, 10), (1, 4, 5, 7), (2, 4, 33, 12)])
def test_packed_parity(torch, oracle, model, mask, shape):
    """CUDA kernel vs oracle on the same seeded inputs: aligned (w%4==0) and generic shapes."""
    from eld_b200._lib import model_mask
    assert model_mask(model) == mask
    rs = np.random.RandomState(hash((model, shape)) % 2 ** 31)
    y = rs.rand(*shape).astype(np.float32)
    y[0, 0, 0, :2] = 0.0                                   # lambda == 0 branch
    p = FULL if mask & 0x78 else SONY
    nm = _nm(model)
    z = nm.batch_gpu(torch.from_numpy(y).cuda(), params=p, frame_id0=77, clip=False).cpu().numpy()
    ref = oracle.noise_packed(y, [p] * shape[0], mask, 1234, 77, False)
    _compare(z, ref, y, p, mask)
    zc = nm.batch_gpu(torch.from_numpy(y).cuda(), params=p, frame_id0=77, clip=True).cpu().numpy()
    assert zc.min() >= 0 and zc.max() <= 1
    assert np.array_equal(zc, np.clip(z, 0, 1))            # sid_dataset.py:277


def test_golden_ramp_inputs(torch, oracle, golden_dir):
    """Same clean frame and parameters as the reference KAT (noise_kat.json): the reference output
    is one realisation; ours must be a realisation of the same law -> standardised residuals."""
    kat = json.load(open(os.path.join(golden_dir, 'noise_kat.json')))
    y = (np.arange(256).reshape(1, 4, 8, 8).astype(np.float32)) / 256
    for model, mask in (('P+g', 5), ('p+g', 6), ('g', 4)):
        ref = np.asarray(kat['call_seed123'][model]['z']).reshape(4, 8, 8)
        z = _nm(model).batch_gpu(torch.from_numpy(y).cuda(), params=SONY, frame_id0=0, clip=False).cpu().numpy()[0]
        sig = _sigma(y[0], SONY, mask)
        for r in ((ref - y[0]) / sig, (z - y[0]) / sig):
            assert abs(r.mean()) < 0.25 and 0.75 < r.std() < 1.25 and np.abs(r).max() < 5.5


def test_reference_call_signature(torch, oracle):
    """noise_maker(y, params) numpy->numpy, one frame, unclipped; reproducible under np.random.seed."""
    nm = _nm('p+g')
    y = np.random.RandomState(0).rand(4, 16, 16).astype(np.float32)
    np.random.seed(5)
    a = nm(y)
    np.random.seed(5)
    b = nm(y)
    assert a.dtype == np.float32 and a.shape == y.shape and np.array_equal(a, b)
    np.random.seed(5)
    K, g, sat, ratio = nm._sample_params()
    fid = int(np.random.randint(0, 2 ** 62))
    ref = oracle.noise_packed(y[None], [(K, g, sat, ratio)], 0x06, 1234, fid, False)[0]
    _compare(a, ref, y, (K, g, sat, ratio), 0x06)


def test_sharding_invariance_and_chunking(torch):
    """Philox key = GLOBAL frame id: any split of the batch gives bit-identical frames;
    also crosses the 48-frames-per-launch chunk boundary."""
    nm = _nm('P+g')
    rs = np.random.RandomState(3)
    y = torch.from_numpy(rs.rand(100, 4, 8, 16).astype(np.float32)).cuda()
    plist = [(1.0 + 0.01 * i, 2.0, 15583, 150.0 + i) for i in range(100)]
    whole = nm.batch_gpu(y, params=plist, frame_id0=1000)
    parts = torch.cat([nm.batch_gpu(y[a:b], params=plist[a:b], frame_id0=1000 + a)
                       for a, b in ((0, 13), (13, 50), (50, 51), (51, 100))])
    assert torch.equal(whole, parts)
    again = nm.batch_gpu(y, params=plist, frame_id0=1000)
    assert torch.equal(whole, again)
    other = nm.batch_gpu(y, params=plist, frame_id0=1001)
    assert not torch.equal(whole, other)


def test_in_place_and_empty(torch):
    nm = _nm('p+g')
    y = torch.rand(2, 4, 32, 32, device='cuda')
    ref = nm.batch_gpu(y, params=SONY, frame_id0=5)
    y2 = y.clone()
    out = nm.batch_gpu(y2, params=SONY, frame_id0=5, out=y2)
    assert out.data_ptr() == y2.data_ptr() and torch.equal(out, ref)
    e = nm.batch_gpu(torch.empty(0, 4, 8, 8, device='cuda'), params=[], frame_id0=0)
    assert e.shape == (0, 4, 8, 8)
    e = nm.batch_gpu(torch.empty(2, 4, 0, 8, device='cuda'), params=SONY, frame_id0=0)
    assert e.numel() == 0


def test_bad_arguments_raise(torch):
    from eld_b200 import _lib
    nm = _nm('g')
    with pytest.raises(_lib.EldError):
        nm.batch_gpu(torch.rand(1, 4, 8, 8, device='cuda'), params=(0.0, 1.0, 15583, 100.0), frame_id0=0)   # K == 0
    lib = _lib.load()
    rc = lib.eld_noise_packed(_lib.ctx(0), None, None, 1, 8, 8, None, 4, 0, 0, 0, None)
    assert rc == -1 and b'NULL' in lib.eld_last_error()
    rc = lib.eld_noise_packed(_lib.ctx(0), None, None, 1, 8, 8, None, 0x80, 0, 0, 0, None)
    assert rc == -1


@pytest.mark.parametrize('dtype', ['u16', 'f32'])
@pytest.mark.parametrize('HW', [(16, 16), (12, 20), (64, 48), (10, 6)])
def test_mosaic_pack_bit_exact(torch, oracle, golden_dir, dtype, HW):
    """Fused Bayer pack with the noise terms off == RawPacker.pack_raw_bayer: BIT-EXACT index map."""
    from oracle import ref_numpy
    H, W = HW
    rs = np.random.RandomState(H * 100 + W)
    m = rs.randint(0, 16384, size=(2, H, W)).astype(np.uint16)
    one = (1.0, 0.0, 1.0, 1.0)                       # sat = ratio = 1 -> scale_in = scale_out = 1
    nm = _nm('')
    if dtype == 'u16':
        t = torch.from_numpy(m.view(np.int16)).cuda()
    else:
        t = torch.from_numpy(m.astype(np.float32)).cuda()
    noisy, clean = nm.mosaic_gpu(t, black=0.0, white=1.0, params=one, frame_id0=0, clip=False)
    want = np.stack([ref_numpy.pack_raw_bayer(m[i]) for i in range(2)])
    assert np.array_equal(clean.cpu().numpy(), want)
    assert np.array_equal(noisy.cpu().numpy(), want)
    assert np.array_equal(want[0], oracle.pack_bayer(m[0]))


def test_mosaic_golden_pack(torch, golden_dir):
    g = json.load(open(os.path.join(golden_dir, 'pack_kat.json')))
    m = np.asarray(g['mosaic_16x12'], np.uint16)
    t = torch.from_numpy(m.view(np.int16)[None]).cuda()
    _, clean = _nm('').mosaic_gpu(t, black=0.0, white=1.0, params=(1.0, 0.0, 1.0, 1.0), frame_id0=0, clip=False)
    assert np.array_equal(clean.cpu().numpy()[0], np.asarray(g['packed_16x12'], np.float32))
    m = np.asarray(g['mosaic_8x6'], np.float32)
    _, clean = _nm('').mosaic_gpu(torch.from_numpy(m[None]).cuda(), black=0.0, white=1.0,
                                  params=(1.0, 0.0, 1.0, 1.0), frame_id0=0, clip=False)
    assert np.array_equal(clean.cpu().numpy()[0], np.asarray(g['packed'], np.float32))


@pytest.mark.parametrize('model,mask', [('P+g', 5), ('p+g', 6), ('ELD:P+G+B+R+U', 0x79)])
def test_mosaic_noise_parity(torch, oracle, model, mask):
    """uint16 LMDB-style mosaic -> dequantise (x/65535, lmdb_dataset.py:38) -> pack -> noise, vs oracle;
    and mosaic path == packed path on the packed clean frame (same random stream)."""
    rs = np.random.RandomState(9)
    m = rs.randint(0, 65536, size=(2, 32, 48)).astype(np.uint16)
    p = FULL if mask & 0x78 else SONY
    nm = _nm(model)
    noisy, clean = nm.mosaic_gpu(torch.from_numpy(m.view(np.int16)).cuda(), black=0.0, white=65535.0,
                                 params=p, frame_id0=40, clip=True)
    rn, rc = oracle.noise_mosaic(m, 0.0, 65535.0, [p] * 2, mask, 1234, 40, True)
    assert np.array_equal(clean.cpu().numpy(), rc)
    got = noisy.cpu().numpy()
    sig = _sigma(rc, p, mask)
    err = np.abs(got.astype(np.float64) - rn)
    assert (err > ATOL_SIGMA * sig + 1e-6).mean() <= (MISMATCH_FRAC if mask & 1 else 0)
    via_packed = nm.batch_gpu(clean, params=p, frame_id0=40, clip=True)
    assert torch.equal(via_packed, noisy)


@pytest.mark.parametrize('model,mask', [('P+g', 5), ('p+g', 6), ('g', 4)])
def test_full_size_moments(torch, model, mask):
    """BASELINE size (4x512x512): size-independent properties - E[z]=y, Var = (ratio/sat)^2 (K y_DN + g^2)
    per intensity bin, Poisson part integer-valued, no NaN/Inf, determinism."""
    nm = _nm(model)
    levels = torch.linspace(0.02, 0.98, 8, device='cuda')
    y = levels.view(8, 1, 1, 1).expand(8, 4, 512, 512).contiguous()
    z = nm.batch_gpu(y, params=SONY, frame_id0=0, clip=False)
    assert torch.isfinite(z).all()
    K, g, sat, ratio = SONY
    for i in range(8):
        yi = float(levels[i])
        zi = z[i].double()
        var = (ratio / sat) ** 2 * ((K * yi * sat / ratio if mask & 3 else 0) + g * g)
        n = zi.numel()
        assert abs(zi.mean().item() - yi) < 6 * (var / n) ** 0.5
        assert abs(zi.var().item() / var - 1) < 0.01
    assert torch.equal(z, nm.batch_gpu(y, params=SONY, frame_id0=0, clip=False))
    if mask == 5:
        cnt = _nm('P').batch_gpu(y, params=SONY, frame_id0=0, clip=False)
        k = cnt.double() * sat / ratio / K
        assert (k - k.round()).abs().max() < 2e-3     # integer photon counts times K


def test_full_size_poisson_law_against_scipy(torch):
    """4x512x512 constant frame -> 1M Poisson draws per lambda on the GPU vs the exact CDF."""
    from scipy import stats
    nm = _nm('P')
    K, g, sat, ratio = 1.0, 0.0, 1000.0, 100.0
    for lam in (0.6, 7.0, 12.0, 250.0):
        y = torch.full((1, 4, 512, 512), lam / (sat / ratio), device='cuda')
        k = (nm.batch_gpu(y, params=(K, g, sat, ratio), frame_id0=3, clip=False).double() * sat / ratio).round()
        x = np.sort(k.cpu().numpy().ravel())
        ks = np.arange(0, x.max() + 2)
        emp = np.searchsorted(x, ks, side='right') / x.size
        lam_eff = float(np.float32(np.float32(lam / (sat / ratio)) * np.float32(sat / ratio)))
        d = np.abs(emp - stats.poisson.cdf(ks, lam_eff)).max()
        assert d < 1.63 / np.sqrt(x.size), (lam, d)


@pytest.mark.parametrize('model,mask', [('p+g', 6), ('P+g', 5)])
def test_lmdb_u16_ingest(torch, oracle, model, mask):
    """LMDB wire format (packed uint16, lmdb_dataset.py:38-39) -> de-quantise -> noise == the f32 path on the
    de-quantised frame (same random stream), and the de-quantised target is bit-exact clip(v/65535,0,1)."""
    rs = np.random.RandomState(21)
    v = rs.randint(0, 65536, size=(3, 4, 16, 24)).astype(np.uint16)
    nm = _nm(model)
    noisy, clean = nm.lmdb_gpu(torch.from_numpy(v.view(np.int16)).cuda(), params=SONY, frame_id0=9)
    want_clean = np.clip(v.astype(np.float32) * np.float32(1.0 / 65535.0), 0, 1).astype(np.float32)
    assert np.array_equal(clean.cpu().numpy(), want_clean)
    via_f32 = nm.batch_gpu(clean, params=SONY, frame_id0=9, clip=True)
    assert torch.equal(via_f32, noisy)
    ref = oracle.noise_packed(want_clean, [SONY] * 3, mask, 1234, 9, True)
    sig = _sigma(want_clean, SONY, mask)
    bad = np.abs(noisy.cpu().numpy().astype(np.float64) - ref) > ATOL_SIGMA * sig + 1e-6
    assert bad.mean() <= (MISMATCH_FRAC if mask & 1 else 0)


def _np_augment(x, flags):
    """ELDTrainDataset.__getitem__ (dataset/sid_dataset.py:344-352) on one [4,h,w] frame."""
    if flags & 1:
        x = np.flip(x, axis=1)
    if flags & 2:
        x = np.flip(x, axis=2)
    if flags & 4:
        x = np.transpose(x, (0, 2, 1))
    return np.ascontiguousarray(x)


@pytest.mark.gpu
@pytest.mark.parametrize('model', ['p+g', 'P+g'])
def test_fused_augmentation_is_an_exact_index_map(torch, model):
    """noise + flips/transpose/clip in one pass == numpy's flips applied to the un-augmented kernel
    output (bit for bit), for all 8 flag combinations; the target is the same map of the clean frame."""
    from eld_b200.noise import NoiseModel
    nm = NoiseModel(model, include=4, verbose=False, seed=31)
    rs = np.random.RandomState(5)
    y = rs.rand(8, 4, 64, 64).astype(np.float32)
    clean = torch.from_numpy(y).cuda()
    flags = np.arange(8, dtype=np.uint8)
    plain = nm.batch_gpu(clean, params=SONY, frame_id0=100, clip=True).cpu().numpy()
    noisy, target = nm.batch_gpu_augmented(clean, aug=flags, params=SONY, frame_id0=100, clip=True)
    noisy, target = noisy.cpu().numpy(), target.cpu().numpy()
    for f in range(8):
        assert np.array_equal(noisy[f], _np_augment(plain[f], int(flags[f]))), 'input, flags %d' % flags[f]
        assert np.array_equal(target[f], _np_augment(y[f], int(flags[f]))), 'target, flags %d' % flags[f]
    # non-square frames: flips only; transpose is refused
    y2 = rs.rand(2, 4, 32, 48).astype(np.float32)
    c2 = torch.from_numpy(y2).cuda()
    p2 = nm.batch_gpu(c2, params=SONY, frame_id0=7).cpu().numpy()
    n2, t2 = nm.batch_gpu_augmented(c2, aug=np.array([3, 1], dtype=np.uint8), params=SONY, frame_id0=7)
    assert np.array_equal(n2.cpu().numpy()[0], _np_augment(p2[0], 3)) and np.array_equal(t2.cpu().numpy()[1], _np_augment(y2[1], 1))
    from eld_b200._lib import EldError
    with pytest.raises(EldError):
        nm.batch_gpu_augmented(c2, aug=np.array([4, 0], dtype=np.uint8), params=SONY, frame_id0=7)


def test_config3_full_eld_batch32_structure(torch):
    """BASELINE configs[3]: full ELD noise (P + Tukey-lambda read + row + quantisation [+ colour bias]), batch 32 of
    4x512x512, sharded 4 frames per rank over 8 ranks.  Size-independent properties: the frames do not depend on the
    sharding (global frame ids), row noise is ONE draw per sensor row (planes 0,1 share row 2i; planes 3,2 share row
    2i+1, noise.py:16-19) with the calibrated sigma, quantisation noise stays within half a step."""
    from eld_b200.noise import NoiseModel
    nm = NoiseModel('ELD:P+G+B+R+U', include=4, verbose=False, seed=99)
    y = torch.rand(32, 4, 512, 512, device='cuda')
    whole = nm.batch_gpu(y, params=FULL, frame_id0=640, clip=True)
    for rank in range(8):
        part = nm.batch_gpu(y[4 * rank:4 * rank + 4], params=FULL, frame_id0=640 + 4 * rank, clip=True)
        assert torch.equal(part, whole[4 * rank:4 * rank + 4]), 'rank %d' % rank
    assert torch.isfinite(whole).all() and whole.min() >= 0 and whole.max() <= 1
    sat, ratio = FULL['saturation'], FULL['ratio']
    flat = torch.full((2, 4, 512, 512), 0.5, device='cuda')
    r = (NoiseModel('ELD:R', include=4, verbose=False, seed=5).batch_gpu(flat, params=FULL, frame_id0=0, clip=False) - flat) * sat / ratio
    assert (r - r[..., :1]).abs().max() < 1e-3                       # constant along a packed row
    assert torch.allclose(r[:, 0], r[:, 1], atol=1e-3) and torch.allclose(r[:, 3], r[:, 2], atol=1e-3)
    assert not torch.allclose(r[:, 0], r[:, 2], atol=1e-3)           # even and odd sensor rows are independent draws
    rows = torch.cat([r[:, 0, :, 0], r[:, 2, :, 0]]).double()       # 2 frames x 512 x 2 independent normals
    assert abs(rows.std().item() / FULL['R_scale'] - 1) < 0.06 and abs(rows.mean().item()) < 0.1
    u = (NoiseModel('ELD:U', include=4, verbose=False, seed=5).batch_gpu(flat, params=FULL, frame_id0=0, clip=False) - flat) * sat / ratio
    assert u.abs().max() <= 0.5 * FULL['q_step'] + 1e-3
    assert abs(u.double().var().item() / (FULL['q_step'] ** 2 / 12) - 1) < 0.01


def test_config4_full_frame_camera_sweep(torch):
    """BASELINE configs[4]: 4-camera parameter sweep on full 4256x2848 frames (packed 4x1424x2128), frames sharded
    round-robin.  One launch over the four frames == four single-frame launches; moments per camera as calibrated."""
    from eld_b200.noise import NoiseModel
    frames, plist = [], []
    for cam in range(1, 5):                                  # include = 1..4 (noise.py:179-182)
        nm_c = NoiseModel('p+g', include=cam, verbose=False, seed=7)
        np.random.seed(100 + cam)
        plist.append(nm_c._sample_params())
    nm = NoiseModel('p+g', include=4, verbose=False, seed=7)
    y = torch.full((4, 4, 1424, 2128), 0.25, device='cuda')
    z = nm.batch_gpu(y, params=plist, frame_id0=1000, clip=False)
    for f in range(4):
        one = nm.batch_gpu(y[f:f + 1], params=[plist[f]], frame_id0=1000 + f, clip=False)
        assert torch.equal(one[0], z[f])
        K, g, sat, ratio = plist[f]
        var = (ratio / sat) ** 2 * (K * 0.25 * sat / ratio + g * g)
        zf = z[f].double()
        assert abs(zf.mean().item() - 0.25) < 6 * (var / zf.numel()) ** 0.5
        assert abs(zf.var().item() / var - 1) < 0.01
    # the mosaic entry point on the same full frames: bit-equal to packing first (noise.py:10-20)
    m = torch.rand(1, 2848, 4256, device='cuda')
    packed = torch.stack([m[:, 0::2, 0::2], m[:, 0::2, 1::2], m[:, 1::2, 1::2], m[:, 1::2, 0::2]], dim=1).contiguous()
    a, clean = nm.mosaic_gpu(m, black=0.0, white=1.0, params=[plist[0]], frame_id0=5, clip=True)
    b = nm.batch_gpu(packed, params=[plist[0]], frame_id0=5, clip=True)
    assert torch.equal(clean, packed) and torch.equal(a, b)
