#!/usr/bin/env python
"""SpreadOut_Module.forward + backward (server.py:48-63): fused kernels vs the reference formulation in stock PyTorch
eager fp32 on the same GPU.   usage: python tools/spreadout_bench.py [N ...]   (default 6000 40000)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.nn.functional as F
    import __graft_entry__ as G
    G.build()
    import fedfr_b200
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [6000, 40000]
    dev = torch.device("cuda:0")

    def timed(fn, n=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    for n in sizes:
        torch.manual_seed(n)
        fc = torch.randn(n, 512, device=dev) * 0.01
        ours = fedfr_b200.SpreadOut_Module(fc.clone(), margin=0.4, mode="sum")
        ref_fc = fc.clone().requires_grad_(True)
        eye = ~torch.eye(n, dtype=torch.bool, device=dev)

        def step_ours():
            ours.FC.grad = None
            ours().backward()

        def step_ref():
            ref_fc.grad = None
            w = F.normalize(ref_fc)
            (F.relu(torch.matmul(w, w.t()).masked_select(eye) - 0.4) ** 2).sum().backward()

        row = {"N": n, "E": 512, "fused_ms": round(timed(step_ours), 3)}
        try:
            row["torch_eager_fp32_ms"] = round(timed(step_ref), 3)
        except torch.cuda.OutOfMemoryError:
            row["torch_eager_fp32_ms"] = "OOM"
        print(json.dumps(row), flush=True)
        del ours, ref_fc, eye
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
