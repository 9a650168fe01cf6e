"""CPU: the torch restatement of UNetSeeInDark against the golden KAT produced by the reference module."""
import os

import numpy as np
import torch

from oracle.unet_ref import UNetSeeInDarkRef


def test_unet_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'unet_kat.npz'))
    torch.manual_seed(2018)
    net = UNetSeeInDarkRef(4, 4)
    assert sum(p.numel() for p in net.parameters()) == int(g['nparam']) == 7760484
    names = [str(n) for n in g['names']]
    assert list(net.state_dict().keys()) == names                       # checkpoint key compatibility
    for k, v in net.state_dict().items():
        assert np.array_equal(v.reshape(-1)[:16].numpy(), g['head_' + k]), k
        assert abs(v.double().sum().item() - float(g['sum_' + k])) < 1e-9
    torch.set_num_threads(1)
    x, t = torch.from_numpy(g['x']), torch.from_numpy(g['target'])
    out = net(x)
    assert np.allclose(out.detach().numpy(), g['out'], rtol=0, atol=1e-6)
    loss = torch.nn.functional.l1_loss(out, t)
    assert abs(loss.item() - float(g['loss'])) < 1e-6
    loss.backward()
    for k, p in net.named_parameters():
        assert abs(p.grad.double().sum().item() - float(g['gsum_' + k])) < 1e-6 + 1e-4 * abs(float(g['gabs_' + k]))
        assert abs(p.grad.double().abs().sum().item() - float(g['gabs_' + k])) < 1e-4 * abs(float(g['gabs_' + k])) + 1e-9
    assert float(g['deconv_identity_err']) < 1e-6                        # deconv == 1x1 conv + pixel-shuffle
