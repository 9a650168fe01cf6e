"""The eval-path oracle (oracle/eval_ref.py) against the golden produced by the UNMODIFIED reference
(models/ELD_model.py imported with stub modules, tests/golden/make_golden.py eval)."""
import os

import numpy as np


def _kat(golden_dir):
    return np.load(os.path.join(golden_dir, 'eval_kat.npz'))


def test_illuminance_correct_tensor2im_psnr_match_the_reference(golden_dir):
    from oracle import eval_ref
    k = _kat(golden_dir)
    corr = eval_ref.illuminance_correct(k['pred'], k['target'])
    assert np.allclose(corr, k['corrected'], rtol=2e-6, atol=1e-7)
    for i in range(2):
        a = eval_ref.psnr(eval_ref.tensor2im(corr[i:i + 1]), eval_ref.tensor2im(k['target'][i:i + 1]))
        b = eval_ref.psnr(eval_ref.tensor2im(k['pred'][i:i + 1]), eval_ref.tensor2im(k['target'][i:i + 1]))
        assert abs(a - k['psnr_corrected'][i]) <= 1e-5 and abs(b - k['psnr_raw'][i]) <= 1e-5
    # the saturated elements (target == 1) really are excluded: including them changes the gain
    p, s = np.clip(k['pred'][0], 0, 1), k['target'][0]
    assert abs(np.dot(p.ravel(), s.ravel()) / np.dot(p.ravel(), p.ravel()) - (corr[0] / np.maximum(p, 1e-12)).max()) > 1e-5


def test_crop_center_and_forward_chop_match_the_reference(golden_dir):
    import torch
    from oracle import eval_ref
    from oracle.unet_ref import UNetSeeInDarkRef
    k = _kat(golden_dir)
    g = torch.Generator().manual_seed(2018)
    torch.rand(2, 4, 24, 20, generator=g); torch.rand(2, 4, 24, 20, generator=g)        # replay the generator's stream
    big = torch.rand(1, 4, 600, 540, generator=g)
    crop = eval_ref.crop_center(big, 512, 512)
    assert crop.shape == (1, 4, 512, 512) and abs(crop.double().sum().item() - float(k['crop_sum'])) < 1e-6
    assert np.array_equal(crop[0, 0, 0, :4].numpy(), k['crop_first'])
    torch.manual_seed(2018)
    net = UNetSeeInDarkRef(4, 4).eval()
    with torch.no_grad():
        out = eval_ref.forward_chop(net, torch.from_numpy(k['chop_in']))
    assert np.allclose(out.numpy(), k['chop_out'], atol=1e-6)
