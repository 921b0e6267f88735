import sys; sys.path.insert(0,'/root/repo')
import torch, torch.nn as nn
from clip_ppo_b200.policy import NatureCNN
torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False
def net(C,seed):
    torch.manual_seed(seed)
    seq = nn.Sequential(nn.Conv2d(C, 32, 8, stride=4), nn.ReLU(), nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(),
                        nn.Conv2d(64, 64, 3, stride=1), nn.ReLU(), nn.Flatten(), nn.Linear(64 * 7 * 7, 512), nn.ReLU())
    for m in seq:
        if isinstance(m,(nn.Conv2d,nn.Linear)): nn.init.orthogonal_(m.weight, 2**0.5); nn.init.normal_(m.bias,std=0.1)
    return seq.cuda()
for mb,C in ((37,4),(37,3),(256,4),(64,4),(48,4),(33,3)):
    seq=net(C,mb+C); n=NatureCNN.from_sequential(seq)
    g=torch.Generator(device='cuda').manual_seed(mb)
    obs=torch.randint(0,256,(mb,C,84,84),device='cuda',generator=g).float()
    gh=torch.randn(mb,512,device='cuda',generator=g)
    h_ref=seq(obs/255.0); (h_ref*gh).sum().backward()
    h=n(obs,in_scale=1/255.0); (h*gh).sum().backward()
    rel=lambda a,b:((a-b).abs().max()/b.abs().max()).item()
    print(mb,C,'h',f'{rel(h,h_ref):.1e}',' '.join(f'{k}:{rel(p.grad,q.grad):.1e}' for (k,p),(_,q) in zip(n.named_parameters(),seq.named_parameters())))
# also in float64 reference for (37,4)
mb,C=37,4
seq=net(C,mb+C); n=NatureCNN.from_sequential(seq)
g=torch.Generator(device='cuda').manual_seed(mb)
obs=torch.randint(0,256,(mb,C,84,84),device='cuda',generator=g).float(); gh=torch.randn(mb,512,device='cuda',generator=g)
import copy
s64=copy.deepcopy(seq).double()
h64=s64(obs.double()/255.0); (h64*gh.double()).sum().backward()
h_ref=seq(obs/255.0); (h_ref*gh).sum().backward()
h=n(obs,in_scale=1/255.0); (h*gh).sum().backward()
rel=lambda a,b:((a.double()-b).abs().max()/b.abs().max()).item()
print('vs fp64: torch32', ' '.join(f'{rel(q.grad,r.grad):.1e}' for q,r in zip(seq.parameters(),s64.parameters())))
print('vs fp64: native ', ' '.join(f'{rel(p.grad,r.grad):.1e}' for p,r in zip(n.parameters(),s64.parameters())))
