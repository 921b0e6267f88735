import sys; sys.path.insert(0,'/root/repo')
import torch, torch.nn as nn
from clip_ppo_b200.policy import NatureCNN
torch.manual_seed(0)
C, mb = 3, 2048
seq = nn.Sequential(nn.Conv2d(C, 32, 8, stride=4), nn.ReLU(), nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(),
                    nn.Conv2d(64, 64, 3, stride=1), nn.ReLU(), nn.Flatten(), nn.Linear(64 * 7 * 7, 512), nn.ReLU()).cuda()
net = NatureCNN.from_sequential(seq)
x = torch.rand(mb, C, 84, 84, device="cuda"); gh = torch.randn(mb, 512, device="cuda")
for _ in range(2):
    net.zero_grad(); h = net(x); h.backward(gh)
torch.cuda.synchronize()
