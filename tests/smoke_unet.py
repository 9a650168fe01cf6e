"""U-Net leg of __graft_entry__.smoke(): one tiny forward + L1 + backward + Adam on cuda:0, checked
against the CPU oracle (oracle/unet_ref.py)."""


def run():
    import torch
    from eld_b200 import arch
    from oracle.unet_ref import UNetSeeInDarkRef
    torch.manual_seed(2018)
    net = arch.unet(4, 4).cuda()
    torch.manual_seed(2018)
    ref = UNetSeeInDarkRef(4, 4)
    torch.manual_seed(1)
    x, t = torch.rand(1, 4, 128, 256), torch.rand(1, 4, 128, 256)
    out, loss = net.train_step(x.cuda(), t.cuda())
    want = ref(x)
    lref = torch.nn.functional.l1_loss(want, t)
    rel = ((out.cpu() - want.detach()).norm() / want.detach().norm()).item()
    assert rel < 2e-2, rel
    assert abs(loss.item() - lref.item()) < 1e-2 * lref.item()
    lref.backward()
    g, h = ref.conv5_2.weight.grad.reshape(-1), net.conv5_2.weight.grad.cpu().reshape(-1)
    cos = (g @ h / (g.norm() * h.norm())).item()
    assert cos > 0.99, cos
    arch.FusedAdam(net, lr=1e-4).step()
    torch.cuda.synchronize()
