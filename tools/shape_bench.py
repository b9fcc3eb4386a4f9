#!/usr/bin/env python
"""Kernel-side step time at the per-GPU shapes of a W-rank job, measured on ONE GPU (no collectives).

A rank of a W-GPU c3 job sees Bt = 512*W gathered rows against Cs = 1M/W classes: the FLOPs per GPU do not change
with W but the tile geometry does.  This script times normalize+fwd, finalize and bwd at those shapes through the
same CudaOps calls PartialFC.forward_backward makes, so the scaling bench can be read as "kernel geometry" vs
"collectives".     usage: python tools/shape_bench.py [W ...]      (default 1 2 4 8)
"""
import ctypes as CT
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import __graft_entry__ as G
    G.build()
    from fedfr_b200 import _native as N
    from fedfr_b200.ops_cuda import CudaOps

    worlds = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 2, 4, 8]
    classes = int(os.environ.get("SHAPE_CLASSES", "1000000"))
    B, E, s, m = 512, 512, 64.0, 0.4
    dev = torch.device("cuda:0")
    ops = CudaOps(dev)
    if os.environ.get("SHAPE_FWD_CHUNKS"):            # forward class chunks (default 4), normaliser blocks per SM
        N.lib.pfc_set_fwd_overlap(int(os.environ["SHAPE_FWD_CHUNKS"]), int(os.environ.get("SHAPE_NORM_BLOCKS", "2")))
    for W in worlds:
        bt, cs = B * W, classes // W
        torch.manual_seed(100)
        w = torch.randn(cs, E, device=dev) * 0.01
        dw = torch.empty_like(w)
        x = torch.nn.functional.normalize(torch.randn(bt, E, device=dev))
        lab = torch.randint(0, classes, (bt,), device=dev)
        lab = torch.where(lab < cs, lab, torch.full_like(lab, -1))        # labels owned by "this rank"
        x_hat = ops.cast_features(x)

        def step():
            w_hat, inv, stats = ops.normalize_fwd_stats(w, x_hat, lab, s, m, 0)
            row_max, row_sum, loss = ops.finalize(stats.unsqueeze(0))
            ops.bwd(x_hat, w_hat, inv, lab, row_max, row_sum, s, m, 1.0 / bt, dw, False, 0)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        N.lib.pfc_profile_enable(1)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        pm = (CT.c_float * 5)()
        pc = (CT.c_int * 5)()
        N.lib.pfc_profile_collect(pm, pc)
        N.lib.pfc_profile_enable(0)
        names = ["normalize", "fwd", "grad", "dx", "dw"]
        print(json.dumps({"W": W, "Bt": bt, "Cs": cs, "ms_per_step": round(ms, 4), "tflops": round(6.0 * bt * cs * E / ms / 1e9, 1),
                          "phase_ms": {k: round(pm[i] / 3, 4) for i, k in enumerate(names)}}), flush=True)
        del w, dw, x, x_hat
        ops._ws.clear()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
