import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import __graft_entry__ as G
G.build()
import fedfr_b200
from fedfr_b200 import _native as N
B, C, E = 64, 3000, 256
g = torch.Generator().manual_seed(99)
w = torch.randn(C, E, generator=g) * 0.01
x = torch.nn.functional.normalize(torch.randn(B, E, generator=g))
y = torch.randint(0, C, (B,), generator=g)
print("dup labels:", B - len(set(y.tolist())))
dev = torch.device("cuda:0")
def make():
    h = fedfr_b200.PartialFC(0, 0, 1, B, False, fedfr_b200.CosFace(64.0, 0.4), C, 1.0, E, "/tmp")
    h.weight.copy_(w.to(dev))
    return h
hs = [make() for _ in range(2)]
os_ = [torch.optim.SGD([{"params": h.parameters()}], lr=0.05, momentum=0.9, weight_decay=5e-4) for h in hs]
res = []
for rep in range(3):
    for h, o in zip(hs, os_):
        h.sub_weight.grad = None
        xg, loss = h.forward_backward(y.to(dev), x.to(dev), o)
        res.append((xg.clone(), h.sub_weight.grad.clone(), float(loss)))
torch.cuda.synchronize()
for i in range(1, len(res)):
    print(i, "dx equal", torch.equal(res[0][0], res[i][0]), "dw equal", torch.equal(res[0][1], res[i][1]), "maxdiff dw", float((res[0][1] - res[i][1]).abs().max()), res[i][2])
# now step: a = torch, b = fused
hs[0].sub_weight.grad = res[-2][1].clone(); hs[1].sub_weight.grad = res[-1][1].clone()
os_[0].step(); hs[0].update(); hs[1].step(os_[1])
print("weights equal after step", torch.equal(hs[0].weight, hs[1].weight))
for h, o in zip(hs, os_):
    h.sub_weight.grad = None
r2 = []
for h, o in zip(hs, os_):
    xg, loss = h.forward_backward(y.to(dev), x.to(dev), o)
    r2.append((xg.clone(), h.sub_weight.grad.clone(), float(loss), h._norm[0].clone(), h._norm[1].clone()))
print("after step: w_hat equal", torch.equal(r2[0][3], r2[1][3]), "inv equal", torch.equal(r2[0][4], r2[1][4]), "dx equal", torch.equal(r2[0][0], r2[1][0]), "dw equal", torch.equal(r2[0][1], r2[1][1]),
      float((r2[0][1] - r2[1][1]).abs().max()), r2[0][2], r2[1][2])
