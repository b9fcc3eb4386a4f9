"""Developer probe: cycle breakdown of CTA 0 of the dw kernel (needs a B200)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as G
G.build()
from fedfr_b200 import _native as N
from fedfr_b200.ops_cuda import CudaOps

dev = torch.device("cuda:0")
B, Cc, E = 512, int(os.environ.get("PROBE_CLASSES", "1000000")), 512
ops = CudaOps(dev)
x = torch.nn.functional.normalize(torch.randn(B, E, device=dev))
w = torch.randn(Cc, E, device=dev) * 0.01
label = torch.randint(0, Cc, (B,), device=dev)
x_hat = ops.cast_features(x)
dw = torch.empty_like(w)
dbg = torch.zeros(16, dtype=torch.int64, device=dev)
N.lib.pfc_set_debug_buffer.argtypes = [C.c_void_p]
if os.environ.get("PROBE_DX_SMS"):          # SM budget of the dx kernel; dw gets the rest (set PROBE_CLUSTERS = (148 - dx_sms) / 2)
    N.lib.pfc_set_prob_split(int(os.environ["PROBE_DX_SMS"]), 0.0, -1)


def step():
    w_hat, inv, stats = ops.normalize_fwd_stats(w, x_hat, label, 64.0, 0.4)
    rm, rs, loss = ops.finalize(stats.unsqueeze(0))
    ops.bwd(x_hat, w_hat, inv, label, rm, rs, 64.0, 0.4, 1.0 / B, dw, False)


for _ in range(2):
    step()
N.lib.pfc_set_debug_buffer(dbg.data_ptr())
step()
torch.cuda.synchronize()
N.lib.pfc_set_debug_buffer(None)
d = dbg.tolist()
n_ct = (Cc + 127) // 128
ncl = int(os.environ.get("PROBE_CLUSTERS", "52"))
items = int(os.environ.get("PROBE_ITEMS", "0")) or (n_ct + ncl - 1) // ncl
print(f"per item (~{items} items, cycles): total {d[2]/items:.0f} | MMA warp: wait tmem_empty {d[0]/items:.0f} wait stage {d[1]/items:.0f} | "
      f"epilogue warp0: wait tmem_full {d[3]/items:.0f} wait w_hat {d[4]/items:.0f} work {d[5]/items:.0f} (cumulative: pass1 {d[6]/items:.0f} exchange {d[7]/items:.0f} pass2 {d[8]/items:.0f})")
