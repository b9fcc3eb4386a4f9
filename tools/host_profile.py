"""Developer tool: where does the HOST time of one PartialFC.forward_backward go?  cProfile over N steps at a small GPU shape
(so that the step is host bound), top functions by cumulative and own time.   python tools/host_profile.py [sample_rate] [world_shape_B]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as G
G.build()
import fedfr_b200

sr = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
C, E = 250_000, 512
dev = torch.device("cuda:0")
head = fedfr_b200.PartialFC(0, 0, 1, B, False, fedfr_b200.CosFace(64.0, 0.4), C, sample_rate=sr, embedding_size=E, prefix="/tmp")
opt = torch.optim.SGD([{"params": head.parameters()}], lr=0.1, momentum=0.9, weight_decay=5e-4)
x = torch.nn.functional.normalize(torch.randn(B, E, device=dev))
y = torch.randint(0, C, (B,), device=dev)


def step():
    head.sub_weight.grad = None
    return head.forward_backward(y, x, opt)


for _ in range(20):
    step()
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n):
    step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"sample_rate {sr}: host enqueue {1e3 * t_host / n:.3f} ms/step, with final sync {1e3 * t_all / n:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
st.sort_stats("tottime").print_stats(22)
