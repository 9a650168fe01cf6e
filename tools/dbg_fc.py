import torch, sys
sys.path.insert(0, '.')
from eld_b200 import arch
torch.manual_seed(0)
net = arch.unet(4, 4).cuda().eval()
x = torch.rand(1, 4, 64, 64, device='cuda')
with torch.no_grad():
    y = net(x)
torch.cuda.synchronize()
print('ok', float(y.abs().mean()))
