"""GPU tests of the drop-in ELDModel / Engine seams (reference models/ELD_model.py, engine.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no GPU')
    return torch


def _batch(torch, n, seed, h=128, w=256):
    g = torch.Generator().manual_seed(seed)
    return {'target': torch.rand(n, 4, h, w, generator=g)}


def test_engine_train_loop_matches_reference_step(torch, tmp_path):
    """Engine.train over 3 batches with GPU noise == oracle loop fed the SAME noisy inputs:
    losses within 1e-2 relative, parameters after 3 Adam steps within 2e-3 of the fp32 oracle's."""
    from eld_b200 import models
    from eld_b200.engine import Engine
    from eld_b200.noise import NoiseModel
    from oracle.unet_ref import UNetSeeInDarkRef, l1_train_step
    torch.manual_seed(2018)
    np.random.seed(2018)
    opt = models.default_opt(name='t', checkpoints_dir=str(tmp_path), noise='p+g', noise_on_gpu=True, lr=1e-3)
    nm = NoiseModel('p+g', include=4, verbose=False, seed=7)
    eng = Engine(opt, noise_maker=nm)
    torch.manual_seed(2018)
    ref = UNetSeeInDarkRef(4, 4)
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=0)
    eng.set_learning_rate(1e-3)
    assert eng.model.optimizers[0].param_groups[0]['lr'] == 1e-3
    batches = [_batch(torch, 2, s) for s in range(3)]
    for b in batches:
        eng.model.set_input(b, 'train')
        x = eng.model.input.cpu()
        assert x.min() >= 0 and x.max() <= 1 and not torch.equal(x, b['target'])       # noisy, clipped
        eng.model.optimize_parameters()
        err = eng.model.get_current_errors()
        _, loss_ref = l1_train_step(ref, ropt, x, b['target'])
        assert abs(err['Pixel'] - loss_ref.item()) <= 1e-2 * loss_ref.item()
    for (k, p), (_, q) in zip(ref.named_parameters(), eng.model.netG.named_parameters()):
        d = (p.detach() - q.detach().cpu()).abs().max().item()
        assert d <= 2e-3 + 0.2 * 3e-3, (k, d)        # 3 Adam steps of lr 1e-3 move a weight by <= 3e-3
    avg = eng.train(batches)
    assert eng.epoch == 1 and eng.iterations == 3 and avg['Pixel'] > 0


def test_checkpoint_roundtrip_reference_format(torch, tmp_path):
    from eld_b200 import models
    from oracle.unet_ref import UNetSeeInDarkRef
    opt = models.default_opt(name='ck', checkpoints_dir=str(tmp_path))
    m = models.eld_model()
    m.initialize(opt)
    m.set_input({'input': torch.rand(1, 4, 128, 256), 'target': torch.rand(1, 4, 128, 256)}, 'train')
    m.optimize_parameters()
    m.epoch, m.iterations = 3, 17
    m.save(label='latest')
    sd = torch.load(os.path.join(str(tmp_path), 'ck', 'model_latest.pt'), weights_only=False)
    # ELD_model.py:516-523's four keys (the reference's load reads exactly these) + the running global frame count
    assert set(sd) == {'netG', 'opt_g', 'epoch', 'iterations', 'frames_seen'}
    UNetSeeInDarkRef(4, 4).load_state_dict(sd['netG'])                          # loads into the reference-shaped module
    ref_adam = torch.optim.Adam(UNetSeeInDarkRef(4, 4).parameters())
    ref_adam.load_state_dict(sd['opt_g'])                                       # torch.optim.Adam accepts 'opt_g'
    m2 = models.eld_model()
    opt2 = models.default_opt(name='ck', checkpoints_dir=str(tmp_path), resume=True)
    m2.initialize(opt2)
    assert m2.epoch == 3 and m2.iterations == 17
    assert torch.equal(m2.netG.flat_params, m.netG.flat_params)
    assert torch.equal(m2.optimizer_G.m, m.optimizer_G.m) and m2.optimizer_G.t == 1


def test_eval_and_chop(torch, tmp_path):
    from eld_b200 import models
    opt = models.default_opt(name='ev', checkpoints_dir=str(tmp_path))
    m = models.eld_model()
    m.initialize(opt)
    d = {'input': torch.rand(1, 4, 160, 272), 'target': torch.rand(1, 4, 160, 272), 'fn': ['x']}
    r = m.eval(d)
    assert np.isfinite(r['PSNR'])
    m.set_input(d, 'eval')
    full = m._padded_forward(m.input)
    chop = m.forward_chop(m.input)
    assert chop.shape == full.shape == (1, 4, 160, 272)
    # the quadrant interiors agree with the full-frame forward away from the replicate-padded borders
    assert (chop[:, :, 16:64, 16:120] - full[:, :, 16:64, 16:120]).abs().max().item() < 5e-2


def test_eval_kernels_match_the_reference_golden(torch, tmp_path, golden_dir):
    """csrc/eval.cu against tests/golden/eval_kat.npz = IlluminanceCorrect / tensor2im / PSNR run by the UNMODIFIED
    reference module: corrected frames to 2e-6, PSNR to 1e-4 dB; and forward_chop (ELD_model.py:434-467) through the
    engine against the reference's own forward_chop output (bf16 tiles: rel-L2 <= 2e-2)."""
    from eld_b200 import models
    from oracle.unet_ref import UNetSeeInDarkRef
    k = np.load(os.path.join(golden_dir, 'eval_kat.npz'))
    m = models.eld_model()
    m.initialize(models.default_opt(name='ev2', checkpoints_dir=str(tmp_path), chop=True))
    pred, tgt = torch.from_numpy(k['pred']).cuda(), torch.from_numpy(k['target']).cuda()
    out, psnr, gain = m.eval_metrics(pred, tgt, correct=True)
    assert np.allclose(out.cpu().numpy(), k['corrected'], rtol=2e-6, atol=1e-7)
    assert np.abs(psnr.cpu().numpy() - k['psnr_corrected']).max() <= 1e-4
    _, psnr_raw, g1 = m.eval_metrics(pred, tgt, correct=False)
    assert np.abs(psnr_raw.cpu().numpy() - k['psnr_raw']).max() <= 1e-4 and torch.all(g1 == 1)
    # forward_chop with the reference network's weights (seed 2018 default init)
    torch.manual_seed(2018)
    m.netG.load_state_dict(UNetSeeInDarkRef(4, 4).state_dict())
    m._eval()
    with torch.no_grad():
        chop = m.forward_chop(torch.from_numpy(k['chop_in']).cuda()).cpu().numpy()
    rel = np.linalg.norm(chop - k['chop_out']) / np.linalg.norm(k['chop_out'])
    assert rel <= 2e-2, rel


def test_eval_end_to_end(torch, tmp_path):
    """ELDModel.eval(correct=True, crop=True) == the oracle restatement applied to the engine's own network output."""
    from eld_b200 import models
    from oracle import eval_ref
    m = models.eld_model()
    m.initialize(models.default_opt(name='ev3', checkpoints_dir=str(tmp_path)))
    g = torch.Generator().manual_seed(4)
    t = torch.rand(1, 4, 544, 576, generator=g)
    t[0, 0, 100:108, 100:108] = 1.0
    d = {'input': (t * 0.5 + 0.02 * torch.randn(1, 4, 544, 576, generator=g)).clamp(0, 1), 'target': t, 'fn': ['x']}
    r = m.eval(d, correct=True, crop=True)
    xc, tc = eval_ref.crop_center(d['input'], 512, 512), eval_ref.crop_center(t, 512, 512)
    raw = m._padded_forward(xc.contiguous().cuda()).cpu().numpy()
    corr = eval_ref.illuminance_correct(raw, tc.numpy())
    want = eval_ref.psnr(eval_ref.tensor2im(corr), eval_ref.tensor2im(tc.numpy()))
    want_in = eval_ref.psnr(eval_ref.tensor2im(xc.numpy()), eval_ref.tensor2im(tc.numpy()))
    assert abs(r['PSNR'] - want) < 2e-3 and abs(r['PSNR_input'] - want_in) < 1e-3, (r, want, want_in)


def test_set_input_noise_stream_is_invariant_to_the_gpu_count(torch, tmp_path):
    """drive set_input with params=None (the model draws (K, g_scale, ratio) itself) as ONE rank with batch 4
    and as ranks 0/1 of a 2-rank job with batch 2 each: the union of the noisy inputs is bit-identical, frame for
    frame - pixels AND per-frame parameters - and the running frame count survives a checkpoint."""
    from eld_b200 import models
    from eld_b200.noise import NoiseModel

    def make(world, rank):
        m = models.eld_model()
        m.initialize(models.default_opt(name='w', checkpoints_dir=str(tmp_path), noise='P+g', noise_on_gpu=True),
                     noise_maker=NoiseModel('P+g', include=4, verbose=False, seed=31))
        m.world, m.rank = world, rank
        return m
    g = torch.Generator().manual_seed(1)
    frames = torch.rand(8, 4, 128, 256, generator=g)          # 2 steps x global batch 4
    one = make(1, 0)
    got1 = []
    for s in range(2):
        one.set_input({'target': frames[4 * s:4 * s + 4]}, 'train')
        got1.append(one.input.cpu())
    got2 = [[None, None], [None, None]]
    for r in range(2):
        m = make(2, r)
        for s in range(2):
            m.set_input({'target': frames[4 * s + 2 * r:4 * s + 2 * r + 2]}, 'train')
            got2[s][r] = m.input.cpu()
    for s in range(2):
        assert torch.equal(got1[s], torch.cat(got2[s]))
    # frames of one batch differ in their noise LEVEL (different K per frame), not only in their Philox stream
    flat = torch.full((4, 4, 128, 256), 0.3)
    lv = make(1, 0)
    lv.set_input({'target': flat}, 'train')
    sd = (lv.input.cpu() - flat).flatten(1).std(1)
    assert sd.max() / sd.min() > 1.05, sd
    assert one._frames_seen == 8 and one.state_dict()['frames_seen'] == 8


def test_prefetched_input_equals_the_serial_one(torch, tmp_path):
    """Engine.train's look-ahead (ELDModel.prefetch_input on a side stream) must hand set_input exactly the tensors the
    serial path makes - same global frame ids, same per-frame parameters - and fall back when the dict differs."""
    from eld_b200 import models
    from eld_b200.noise import NoiseModel

    def make():
        m = models.eld_model()
        m.initialize(models.default_opt(name='pf', checkpoints_dir=str(tmp_path), noise='P+g', noise_on_gpu=True, augment_on_gpu=True),
                     noise_maker=NoiseModel('P+g', include=4, verbose=False, seed=5))
        return m
    batches = [_batch(torch, 2, s, 256, 256) for s in range(3)]
    a, b = make(), make()
    serial = []
    for d in batches:
        a.set_input(d, 'train')
        serial.append((a.input.clone(), a.target.clone()))
    b.set_input(batches[0], 'train')
    got = [(b.input.clone(), b.target.clone())]
    for i in (1, 2):
        b.prefetch_input(batches[i])
        b.set_input(batches[i], 'train')
        got.append((b.input.clone(), b.target.clone()))
    torch.cuda.synchronize()
    for (x0, t0), (x1, t1) in zip(serial, got):
        assert torch.equal(x0, x1) and torch.equal(t0, t1)
    b.prefetch_input(batches[0])
    b.set_input(batches[1], 'train')                      # a different dict: the prefetched tensors are dropped, not misused
    assert b.input.shape == (2, 4, 256, 256) and b._prefetched is None


def test_stage_in_srgb_is_wired(torch, tmp_path):
    """train_syn.py:55-58 / ELD_model.py:377-389: --stage_in srgb builds a 3 -> 4 network; with on-GPU synthesis the input
    is ISPDataset's chain (noise -> clip -> raw2rgb_v2(wb, ccm) -> clip, sid_dataset.py:306-312) made on the stream."""
    from eld_b200 import models, process
    from eld_b200.noise import NoiseModel
    m = models.eld_model()
    nm = NoiseModel('p+g', include=4, verbose=False, seed=3)
    m.initialize(models.default_opt(name='srgb', checkpoints_dir=str(tmp_path), stage_in='srgb', noise='p+g', noise_on_gpu=True),
                 noise_maker=nm)
    assert m.netG.conv1_1.weight.shape == (32, 3, 3, 3) and m.netG.conv10_1.weight.shape == (4, 32, 1, 1)
    g = torch.Generator().manual_seed(2)
    t = torch.rand(2, 4, 128, 256, generator=g)
    wb = torch.tensor([[2.0, 1.0, 1.6, 1.0], [1.8, 1.0, 2.1, 1.0]])
    ccm = torch.eye(3).repeat(2, 1, 1) * 1.2
    m.set_input({'target': t, 'wb': wb, 'ccm': ccm}, 'train')
    assert m.input.shape == (2, 3, 128, 256) and m.target.shape == (2, 4, 128, 256)
    noisy = nm.batch_gpu(t.cuda(), params=nm.frame_params(0, 2), frame_id0=0, clip=True)
    want = process.isp_dataset_item(noisy, wb, ccm)
    assert torch.equal(m.input, want)
    m.optimize_parameters()
    assert np.isfinite(m.get_current_errors()['Pixel'])
