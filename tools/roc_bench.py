"""Time the ROC histogram kernel at the reference's launch shape (roc_cuda.py:40-51: a batch of 800 sub rows against
N features, E = 512) and at a full triangle; prints one JSON line per shape.  Unit: pair-dimensions per second
(one fp32 multiply + fp64 add each)."""
import json
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from fedfr_b200 import _native as N  # noqa: E402
from fedfr_b200.roc import calc_ROC  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    mode = int(sys.argv[sys.argv.index("--mode") + 1]) if "--mode" in sys.argv else 0     # 1 = two-tier kernel (opt-in)
    N.check(N.lib.pfc_set_roc_mode(mode), "pfc_set_roc_mode")
    for n, t in [(100_000, 800), (20_000, 20_000)]:
        f = torch.nn.functional.normalize(torch.randn(n, 512, device=dev, generator=g))
        l = torch.randint(0, 1000, (n,), device=dev, generator=g).int()
        out = torch.zeros(4002, dtype=torch.int64, device=dev)
        calc_ROC(f, l, f[:t], l[:t], out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        a.record()
        for _ in range(reps):
            calc_ROC(f, l, f[:t], l[:t], out)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        pairs = t * (t - 1) // 2 + t * (n - t)
        assert int(out.sum()) == pairs * (reps + 1)
        print(json.dumps({"tool": "roc_bench", "mode": mode, "n": n, "sub_rows": t, "emb": 512, "pairs": pairs, "ms": round(ms, 3),
                          "pair_dims_per_s": round(pairs * 512 / (ms * 1e-3), 0), "gpairs_per_s": round(pairs / ms / 1e6, 3)}))


if __name__ == "__main__":
    main()
