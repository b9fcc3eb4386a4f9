"""Run the SIMT kernels of csrc/roc.cu and csrc/bce_head.cu on the CPU under a small CUDA execution-model emulator
(tests/emu/cuda_emu.h: one fiber per CUDA thread, rendezvous at __syncthreads / __shfl) and compare with the oracles.

This checks the kernels' indexing, tiling, barrier placement and reductions without a GPU (the GPU parity tests remain
the authority on the real hardware path).  The kernel source is taken verbatim from the .cu file up to the end of its
``namespace pfc`` block; only the ``#include "common.cuh"`` line is replaced by the emulator header."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

HARNESS = {
    "roc": r'''
extern "C" void emu_roc(const float* feature, const int32_t* label, int64_t n, const float* sub, const int32_t* sublabel,
                        int64_t n_sub, int64_t sub_offset, int emb, unsigned long long* hist, int grid) {
  emu_launch((unsigned)grid, pfc::kRocThreads, [=]() {
    pfc::roc_hist_kernel(feature, label, n, sub, sublabel, n_sub, sub_offset, emb, hist);
  });
}
''',
    "hardneg": r'''
extern "C" void emu_similar(const float* a, int64_t n_a, const float* b, int64_t n_b, int emb, float thr, unsigned char* hit, int grid) {
  emu_launch((unsigned)grid, pfc::kHnThreads, [=]() { pfc::similar_columns_kernel(a, n_a, b, n_b, emb, thr, hit); });
}
''',
    "bce_head": r'''
extern "C" void emu_bce_fwd(const float* feat, const float* weight, const float* bias, const int64_t* label, int64_t n_rows,
                            int64_t n_classes, int emb, float m, float r, float t, float* logits, unsigned char* gt,
                            float* cosine, float* inv_nf, float* inv_nw, int grid) {
  emu_launch((unsigned)grid, 256, [=]() {
    pfc::bce_head_fwd_kernel(feat, weight, bias, label, n_rows, n_classes, emb, m, r, t, logits, gt, cosine, inv_nf, inv_nw);
  });
}
extern "C" void emu_bce_bwd(const float* feat, const float* weight, const float* cosine, const float* inv_nf, const float* inv_nw,
                            const float* dlogits, int64_t n_rows, int64_t n_classes, int emb, float r, float t, float* dfeat,
                            float* dweight, float* dbias) {
  emu_launch((unsigned)(n_rows + n_classes), pfc::kBceThreads, [=]() {
    pfc::bce_head_bwd_kernel(feat, weight, cosine, inv_nf, inv_nw, dlogits, n_rows, n_classes, emb, r, t, dfeat, dweight, dbias);
  });
}
''',
}


def _build(name, tmp):
    src = open(os.path.join(ROOT, "fedfr_b200", "csrc", name + ".cu")).read()
    end = src.index("}  // namespace pfc") + len("}  // namespace pfc")
    body = src[:end].replace('#include "common.cuh"', '#include "cuda_emu.h"')
    assert "cuda_emu.h" in body
    cpp = os.path.join(tmp, name + "_emu.cpp")
    with open(cpp, "w") as f:
        f.write(body + "\n" + HARNESS[name])
    so = os.path.join(tmp, name + "_emu.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                           "-I", os.path.join(HERE, "emu"), cpp, "-o", so])
    return C.CDLL(so)


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    import __graft_entry__ as g
    g.build()                                          # the C oracle
    tmp = str(tmp_path_factory.mktemp("emu"))
    return {n: _build(n, tmp) for n in HARNESS}


@pytest.mark.parametrize("n,emb,t,off,grid", [(150, 40, 100, 0, 3), (131, 33, 131, 0, 2), (70, 7, 20, 35, 1), (64, 32, 64, 0, 5)])
def test_roc_kernel_under_emulation(libs, n, emb, t, off, grid):
    from oracle import roc_oracle as R
    rng = np.random.default_rng(n)
    f = rng.standard_normal((n, emb)).astype(np.float32)
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    f[3] = f[1]
    f[n - 2] = -f[0]
    l = rng.integers(0, 5, n).astype(np.int32)
    sub, subl = np.ascontiguousarray(f[off:off + t]), np.ascontiguousarray(l[off:off + t])
    want = R.roc_histogram(f, l, sub, subl, off)
    hist = np.zeros(4002, dtype=np.uint64)
    libs["roc"].emu_roc(_p(f), _p(l), C.c_int64(n), _p(sub), _p(subl), C.c_int64(sub.shape[0]), C.c_int64(off), emb, _p(hist), grid)
    assert np.array_equal(hist.astype(np.int64), want)


@pytest.mark.parametrize("B,Cn,E,t,grid", [(9, 5, 40, 3, 2), (3, 300, 33, 2, 7), (20, 1, 130, 4, 1), (2, 2, 1024, 3, 1)])
def test_bce_head_kernels_under_emulation(libs, B, Cn, E, t, grid):
    from oracle import bce_head_oracle as O
    g = torch.Generator().manual_seed(B * 1000 + Cn)
    feat = (torch.randn(B, E, generator=g) * 2).numpy()
    weight = (torch.randn(Cn, E, generator=g) * 0.01).numpy()
    bias = (torch.randn(Cn, generator=g) * 0.1).numpy()
    label = torch.randint(-1, Cn + 3, (B,), generator=g).numpy()
    lo, gto, cso, nf, nw = O.forward(torch.from_numpy(feat).double(), torch.from_numpy(weight).double(),
                                     torch.from_numpy(bias).double(), torch.from_numpy(label), 0.4, 30.0, t)
    logits, gt = np.zeros((B, Cn), np.float32), np.zeros((B, Cn), np.uint8)
    cosine, inv_nf, inv_nw = np.zeros((B, Cn), np.float32), np.zeros(B, np.float32), np.zeros(Cn, np.float32)
    libs["bce_head"].emu_bce_fwd(_p(feat), _p(weight), _p(bias), _p(label), C.c_int64(B), C.c_int64(Cn), E, C.c_float(0.4),
                                 C.c_float(30.0), C.c_float(t), _p(logits), _p(gt), _p(cosine), _p(inv_nf), _p(inv_nw), grid)
    assert np.array_equal(gt.astype(bool), gto.numpy())
    np.testing.assert_allclose(logits, lo.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(cosine, cso.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(inv_nf, 1 / nf.numpy(), rtol=1e-5)
    np.testing.assert_allclose(inv_nw, 1 / nw.numpy(), rtol=1e-5)

    dlogits = torch.randn(B, Cn, generator=g).numpy()
    dfo, dwo, dbo = O.backward(torch.from_numpy(feat).double(), torch.from_numpy(weight).double(), cso, nf, nw,
                               torch.from_numpy(dlogits).double(), 30.0, t)
    dfeat, dweight, dbias = np.zeros_like(feat), np.zeros_like(weight), np.zeros_like(bias)
    libs["bce_head"].emu_bce_bwd(_p(feat), _p(weight), _p(cosine), _p(inv_nf), _p(inv_nw), _p(dlogits), C.c_int64(B),
                                 C.c_int64(Cn), E, C.c_float(30.0), C.c_float(t), _p(dfeat), _p(dweight), _p(dbias))
    for got, want in ((dfeat, dfo), (dweight, dwo), (dbias, dbo)):
        want = want.numpy()
        assert np.linalg.norm(got - want) <= 1e-5 * np.linalg.norm(want) + 1e-9

    # dfeat == NULL (detached features) and no bias: the class-role CTAs still produce dweight
    dweight2 = np.zeros_like(weight)
    libs["bce_head"].emu_bce_bwd(_p(feat), _p(weight), _p(cosine), _p(inv_nf), _p(inv_nw), _p(dlogits), C.c_int64(B),
                                 C.c_int64(Cn), E, C.c_float(30.0), C.c_float(t), None, _p(dweight2), None)
    assert np.array_equal(dweight2, dweight)


@pytest.mark.parametrize("na,nb,emb,thr,grid", [(70, 130, 40, 0.1, 3), (5, 64, 33, 0.0, 1), (130, 65, 7, 0.3, 4), (64, 64, 64, -2.0, 2)])
def test_similar_columns_kernel_under_emulation(libs, na, nb, emb, thr, grid):
    from oracle import hardneg_oracle as O
    rng = np.random.default_rng(na * 7 + nb)
    a = rng.standard_normal((na, emb)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = (rng.standard_normal((nb, emb)) + 0.8 * a[rng.integers(0, na, nb)]).astype(np.float32)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    hit = np.zeros(nb, dtype=np.uint8)
    libs["hardneg"].emu_similar(_p(a), C.c_int64(na), _p(b), C.c_int64(nb), emb, C.c_float(thr), _p(hit), grid)
    certain, amb = O.similar_columns(a, b, thr)
    got = set(np.nonzero(hit)[0].tolist())
    assert set(certain.tolist()) <= got <= set(certain.tolist()) | set(amb.tolist())
    assert 0 < len(got) <= nb
