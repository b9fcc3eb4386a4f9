"""NVTX ranges for the host-side phases (tracing row of SURVEY section 5): ``FEDFR_NVTX=1`` turns them on, together with the
ranges the C library puts around its compute entries; off (the default) they cost one attribute test."""
import os

NVTX = os.environ.get("FEDFR_NVTX", "0") not in ("", "0")


class nvtx_range:
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_pop()
        return False
