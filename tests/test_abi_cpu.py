"""CPU tests of the boundary: libeld_b200.so loads, exports every symbol include/*.h declares,
fails loudly without a GPU, and the host-side mirror of the reference interface behaves like
the reference (noise.py:175-225) - no compute calls here."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from tests.conftest import REPO, has_gpu


def _declared_functions():
    names = []
    inc = os.path.join(REPO, 'include')
    for fn in sorted(os.listdir(inc)):
        src = open(os.path.join(inc, fn)).read()
        src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
        src = re.sub(r'//[^\n]*', '', src)
        src = re.sub(r'^\s*#.*$', '', src, flags=re.M)
        for m in re.finditer(r'\b(eld_[a-z0-9_]+)\s*\(', src):
            if m.group(1) not in names:
                names.append(m.group(1))
    return names


def test_library_exports_every_declared_symbol():
    from eld_b200 import _lib
    lib = _lib.load()
    names = _declared_functions()
    assert 'eld_noise_packed' in names and 'eld_ctx_create' in names
    for n in names:
        assert hasattr(lib, n), 'libeld_b200.so does not export %s' % n
    assert lib.eld_abi_version() == 1


@pytest.mark.skipif(has_gpu(), reason='checks the no-GPU failure mode')
def test_no_gpu_fails_loudly():
    from eld_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.eld_ctx_create(0, ctypes.byref(h))
    assert rc != 0 and h.value is None
    assert b'no CPU fallback' in lib.eld_last_error() or b'CUDA' in lib.eld_last_error()
    from eld_b200.noise import NoiseModel
    nm = NoiseModel('g', include=4, verbose=False)
    with pytest.raises(_lib.EldError):
        nm(np.zeros((4, 8, 8), np.float32), params=(1.0, 1.0, 15583, 100.0))


def test_params_struct_layout():
    from eld_b200 import _lib
    assert ctypes.sizeof(_lib.NoiseParams) == 48


def test_model_mask_substring_semantics():
    """noise.py:158-166: 'P' in model wins over 'p'; '+' and order are irrelevant."""
    from eld_b200._lib import model_mask
    assert model_mask('P+g') == 0x05 and model_mask('Pg') == 0x05 and model_mask('g+P') == 0x05
    assert model_mask('p+g') == 0x06 and model_mask('g') == 0x04 and model_mask('P') == 0x01
    assert model_mask('Pp') == 0x01
    assert model_mask('ELD:P+G+B+R+U') == 0x79
    # reference strings keep their reference meaning (README.md:46 names the baselines G / G+P / G+P*): without the
    # explicit 'ELD:' prefix the letters G, B, R, U select nothing, exactly like noise.py:158-166
    assert model_mask('G+P') == 0x01 and model_mask('G') == 0x00 and model_mask('G+P*') == 0x01


def test_noise_model_host_side_matches_reference_golden(golden_dir):
    """_sample_params: same numpy-RNG call order and values as the reference (golden KAT)."""
    from eld_b200.noise import NoiseModel
    kat = json.load(open(os.path.join(golden_dir, 'noise_kat.json')))
    cams = ['CanonEOS5D4', 'CanonEOS70D', 'CanonEOS700D', 'NikonD850', 'SonyA7S2']
    for i, cam in enumerate(cams):
        nm = NoiseModel('P+g', include=i, verbose=False)
        assert nm.cameras == [cam]
        np.random.seed(0)
        K, g, sat, ratio = nm._sample_params()
        gold = kat['sample_params_seed0'][cam]
        assert (K, g, sat, ratio) == (gold['K'], gold['g_scale'], gold['sat'], gold['ratio'])
    nm = NoiseModel('g', verbose=False)
    np.random.seed(11)
    got = [[float(v) for v in nm._sample_params()] for _ in range(4)]
    assert got == kat['sample_params_allcams_seed11_x4']


def test_noise_model_ctor_contract():
    from eld_b200.noise import NoiseModel
    with pytest.raises(AssertionError):
        NoiseModel('g', include=1, exclude=2, verbose=False)      # noise.py:178
    with pytest.raises(AssertionError):
        NoiseModel('g', cfa='foveon', verbose=False)              # noise.py:177
    nm = NoiseModel('g', exclude=0, verbose=False)
    assert 'CanonEOS5D4' not in nm.cameras and len(nm.cameras) == 4
    nm.model = 'P+g'                                              # .model is a mutable str attribute
    full = NoiseModel('ELD:P+G+B+R+U', include=4, verbose=False)
    np.random.seed(3)
    p = full._sample_params_full()
    assert set(p) >= {'K', 'g_scale', 'G_scale', 'G_lambda', 'R_scale', 'color_bias', 'ratio', 'q_step'}
    assert 100 <= p['ratio'] <= 300 and 0.1 <= p['K'] <= 30 and len(p['color_bias']) == 4


def test_sample_augment_follows_the_reference_rng_order():
    """Three randint(2) draws per frame in the reference's order (sid_dataset.py:344,347,350) - host logic, no GPU."""
    from eld_b200.noise import NoiseModel
    np.random.seed(77)
    got = NoiseModel.sample_augment(5)
    np.random.seed(77)
    want = [sum(b for b in (1, 2, 4) if np.random.randint(2, size=1)[0] == 1) for _ in range(5)]
    assert got.tolist() == want and got.dtype == np.uint8


def test_read_emor_and_load_crf(tmp_path):
    """util/process.py:147-175 restated: EMoR text blocks and the per-channel CRF table (synthetic files, same format)."""
    from eld_b200 import process
    E = np.linspace(0, 1, 1024)
    with open(tmp_path / 'emor.txt', 'w') as f:
        for name, arr in (('E', E), ('f0', E ** 0.5), ('h( 1)', np.sin(E))):
            f.write('%s =\n' % name)
            for i in range(0, 1024, 4):
                f.write(' '.join('%.6e' % v for v in arr[i:i + 4]) + '\n')
    np.savetxt(tmp_path / 'CRF_SonyA7S2_5.txt', np.stack([E ** 0.4, E ** 0.45, E ** 0.5]))
    e, f0, H = process.read_emor(str(tmp_path / 'emor.txt'))
    assert e.shape == (1024,) and np.allclose(e, E, atol=1e-6) and np.allclose(f0, E ** 0.5, atol=1e-6) and H.shape == (1, 1024)
    Erep, fs = process.load_CRF(str(tmp_path))
    assert Erep.shape == (3, 1024) and fs.shape == (3, 1024) and np.allclose(Erep[2], E, atol=1e-6)


def test_frame_params_are_keyed_by_global_frame_id():
    """per-frame (K, g_scale, ratio) and flip flags used by ELDModel.set_input come from a generator keyed by
    (seed, global frame id) - W ranks draw W*B different tuples, frame f gets the same tuple at any world size, numpy's
    global stream is left untouched, and each draw follows noise.py:201-225's distributions."""
    from eld_b200.noise import NoiseModel
    nm = NoiseModel('P+g', include=4, verbose=False, seed=2018)
    np.random.seed(5)
    before = np.random.get_state()[1].copy()
    B = 4
    one = nm.frame_params(0, 2 * B)                                   # world 1: frames 0..7 in one batch
    two = [nm.frame_params(r * B, B) for r in range(2)]               # world 2: rank r owns [r*B, (r+1)*B)
    assert np.array_equal(np.random.get_state()[1], before)          # global RNG not consumed
    assert one == two[0] + two[1]
    assert len({p[0] for p in one}) == 2 * B                          # 8 distinct K, not 2 copies of 4
    assert all(0.1 <= p[0] <= 30 and 100 <= p[3] <= 300 and p[2] == 15583 for p in one)
    assert nm.frame_params(3, 1)[0] == one[3]                         # any sharding
    a1, a2 = nm.frame_augment(0, 8), np.concatenate([nm.frame_augment(0, 4), nm.frame_augment(4, 4)])
    assert np.array_equal(a1, a2) and a1.max() <= 7
    # same call order as the reference on the per-frame RandomState: reproduce frame 3 by hand
    rng = nm._frame_rng(3)
    np.random.set_state(rng.get_state())
    want = nm._sample_params()
    assert want == one[3]
    full = NoiseModel('ELD:P+G+B+R+U', include=4, verbose=False, seed=1)
    f = full.frame_params(10, 2)
    assert f[0] != f[1] and set(f[0]) >= {'K', 'G_scale', 'R_scale', 'G_lambda', 'color_bias'}
    assert NoiseModel('G+P', include=4, verbose=False).frame_params(0, 1)[0][2] == 15583   # reference string: 4-tuple


def test_grad_buckets_partition_the_flat_gradient():
    """the data-parallel buckets are contiguous ranges of the state_dict-ordered flat gradient, listed in
    backward-completion order (decoder first, encoder last), and together cover every parameter exactly once."""
    import ctypes as c
    from eld_b200 import _lib
    lib = _lib.load()
    arr = (c.c_size_t * 16)()
    k = lib.eld_unet_grad_buckets(arr, 16)
    b = [(int(arr[2 * i]), int(arr[2 * i + 1])) for i in range(k)]
    n = lib.eld_unet_param_count()
    assert k == 4 and sum(cnt for _, cnt in b) == n == 7760484
    spans = sorted(b)
    assert spans[0][0] == 0 and all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(k - 1)) and spans[-1][0] + spans[-1][1] == n
    off, cnt = c.c_size_t(), c.c_size_t()
    lib.eld_unet_param_offset(b'upv6', 0, c.byref(off), c.byref(cnt))
    assert b[0][0] == off.value                        # decoder bucket starts at upv6.weight ...
    lib.eld_unet_param_offset(b'conv5_1', 0, c.byref(off), c.byref(cnt))
    assert b[1][0] == off.value and b[3][0] == 0       # ... bottleneck at conv5_1.weight, the last bucket at conv1_1.weight
    lib.eld_unet_param_offset(b'conv2_1', 0, c.byref(off), c.byref(cnt))
    assert b[2][0] == off.value and b[3][1] == off.value   # encoder bucket conv2_1..conv4_2; conv1_1 + conv1_2 travel last


def test_burst_frames_share_parameters():
    """SynDataset (sid_dataset.py:269-275): one _sample_params() per burst of k frames.  frame_params(burst=k) gives the
    frames f // k == const the same tuple, at any sharding of the burst over calls."""
    from eld_b200.noise import NoiseModel
    nm = NoiseModel('P+g', include=4, verbose=False, seed=7)
    p = nm.frame_params(0, 8, burst=4)
    assert p[0] == p[1] == p[2] == p[3] and p[4] == p[7] and p[0] != p[4]
    assert nm.frame_params(2, 4, burst=4) == p[2:6]
    assert nm.frame_params(0, 4) != p[:4]                     # burst = 1: every frame its own draw
