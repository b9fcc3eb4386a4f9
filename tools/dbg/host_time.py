import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import __graft_entry__ as G
G.build()
import fedfr_b200
B, C, E = 512, 1_000_000, 512
dev = torch.device("cuda:0")
head = fedfr_b200.PartialFC(0, 0, 1, B, False, fedfr_b200.CosFace(64.0, 0.4), C, 1.0, E, "/tmp")
opt = torch.optim.SGD([{"params": head.parameters()}], lr=0.1, momentum=0.9, weight_decay=5e-4)
x = torch.nn.functional.normalize(torch.randn(B, E, device=dev)); y = torch.randint(0, C, (B,), device=dev)
for _ in range(5):
    head.sub_weight.grad = None
    head.forward_backward(y, x, opt)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    head.sub_weight.grad = None
    t0 = time.perf_counter()
    head.forward_backward(y, x, opt)
    ts.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
print("host time per forward_backward (us):", sorted(int(t * 1e6) for t in ts))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    head.sub_weight.grad = None
    head.forward_backward(y, x, opt)
    torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
