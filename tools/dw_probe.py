"""Developer probe: cycle breakdown of CTA 0 of the dw kernel (needs a B200)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as G
G.build()
from fedfr_b200 import _native as N
from fedfr_b200.ops_cuda import CudaOps

dev = torch.device("cuda:0")
B, Cc, E = 512, 262144, 512
ops = CudaOps(dev)
x = torch.nn.functional.normalize(torch.randn(B, E, device=dev))
w = torch.randn(Cc, E, device=dev) * 0.01
label = torch.randint(0, Cc, (B,), device=dev)
w_hat, inv = ops.normalize(w)
x_hat = ops.cast_features(x)
stats = ops.fwd_stats(x_hat, w_hat, label, 64.0, 0.4)
rm, rs, loss = ops.finalize(stats.unsqueeze(0))
dw = torch.empty_like(w)
dbg = torch.zeros(16, dtype=torch.int64, device=dev)
N.lib.pfc_set_debug_buffer.argtypes = [C.c_void_p]
for cs in (1, 2, 4):
    N.lib.pfc_set_clusters(cs, cs)
    for _ in range(2):
        ops.bwd(x_hat, w_hat, inv, label, rm, rs, 64.0, 0.4, 1.0 / B, dw, False)
    N.lib.pfc_set_debug_buffer(dbg.data_ptr())
    ops.bwd(x_hat, w_hat, inv, label, rm, rs, 64.0, 0.4, 1.0 / B, dw, False)
    torch.cuda.synchronize()
    N.lib.pfc_set_debug_buffer(None)
    d = dbg.tolist()
    items = 2 * ((Cc // 128 + 147) // 148)
    print(f"cluster {cs}: per item (~{items} items, cycles): total {d[2]/items:.0f} | MMA warp: wait tmem_empty {d[0]/items:.0f} wait stage {d[1]/items:.0f} | "
          f"epilogue warp0: wait tmem_full {d[3]/items:.0f} wait w_hat {d[4]/items:.0f} work {d[5]/items:.0f}")
