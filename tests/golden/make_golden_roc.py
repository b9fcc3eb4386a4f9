"""Generate tests/golden/roc.npz by running the UNMODIFIED reference kernel ``roc_cuda.calc_ROC``
(roc_cuda.py:14-28) under numba's CUDA simulator, and ``roc_cuda.plot_ROC`` (roc_cuda.py:55-88) on its output.

Run in the build container only (needs /root/reference and numba):

    NUMBA_ENABLE_CUDASIM=1 python tests/golden/make_golden_roc.py

Types.  numba compiles the kernel with ``tmp`` a float64 and ``subfeature[i, k] * feature[j, k]`` a float32 product:
``check_compiled_arithmetic`` compiles the unmodified kernel to PTX (no GPU needed) and asserts the chain
``mul.f32 -> cvt.f64.f32 -> add.f64``, then ``add.f64 1.0``, ``mul.f64 1000.0``, ``cvt.rzi.s64.f64``, no fma.  The
simulator instead executes the body with numpy scalars, and numpy >= 2 (NEP 50) makes ``0. + float32`` a float32 --
NOT what the compiled kernel does.  So the features are handed to the simulator as object arrays of ``_F32`` scalars
whose product is the float32-rounded product and whose sum with a Python float is a Python float (a double): the
compiled kernel's types, executed by the reference's own source.  The launches use the grid/block geometry and the
slicing of ``gpu_job_consumer`` (roc_cuda.py:40-52): block (32, 32), ``feature[start:]`` against its first
``len(index)`` rows, int32 labels, float64 ``out`` cast to int64.  No reference source is copied; only inputs and
outputs are stored.
"""
import os
import re
import sys
import tempfile

os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")
import numpy as np

sys.path.insert(0, "/root/reference")
import roc_cuda  # noqa: E402
from numba import cuda  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class _F32:
    """A float32 value with the compiled kernel's promotion: f32 * f32 -> f32 (rounded), double + f32 -> double."""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = np.float32(v)

    def __mul__(self, other):
        return _F32(self.v * other.v)

    def __radd__(self, acc):
        return float(acc) + float(self.v)


def _boxed(a):
    out = np.empty(a.shape, dtype=object)
    for idx in np.ndindex(a.shape):
        out[idx] = _F32(a[idx])
    return out


_PTX_CHECK = r"""
import re, sys
sys.path.insert(0, "/root/reference")
import roc_cuda
from numba import cuda, types as T
sig = (T.float32[:, :], T.int32[:], T.float32[:, :], T.int32[:], T.float64[:])
ptx, _ = cuda.compile_ptx(roc_cuda.calc_ROC.py_func, sig, cc=(9, 0))
n_mul, n_cvt = len(re.findall(r"mul\.f32", ptx)), len(re.findall(r"cvt\.f64\.f32", ptx))
assert n_mul == n_cvt >= 1 and "fma" not in ptx and "add.f32" not in ptx
assert len(re.findall(r"add\.f64", ptx)) >= n_mul + 1
assert "0d408F400000000000" in ptx and "cvt.rzi.s64.f64" in ptx      # * 1000.0, truncation
print("compiled kernel: mul.f32 -> cvt.f64.f32 -> add.f64 chain confirmed (%d unrolled steps)" % n_mul)
"""


def check_compiled_arithmetic():
    """Needs the real compiler, so it runs in a child process without the simulator switch."""
    import subprocess
    env = {k: v for k, v in os.environ.items() if k != "NUMBA_ENABLE_CUDASIM"}
    subprocess.check_call([sys.executable, "-c", _PTX_CHECK], env=env, cwd=tempfile.gettempdir())


def run_reference(feature, label, batch_size, target_size):
    out_sum = np.zeros(2001 * 2, dtype=np.int64)
    per_batch = []
    for start in range(0, target_size, batch_size):
        index = np.arange(start, min(start + batch_size, target_size))
        f = cuda.to_device(_boxed(feature[start:, :].astype(np.float32)))
        l = cuda.to_device(label[start:].astype(np.int32))
        sf = cuda.to_device(_boxed(feature[index, :].astype(np.float32)))
        sl = cuda.to_device(label[index].astype(np.int32))
        out = cuda.to_device(np.zeros(2001 * 2, dtype=np.float64))
        grid = ((len(index) + 31) // 32, (f.shape[0] + 31) // 32)
        roc_cuda.calc_ROC[grid, (32, 32)](f, l, sf, sl, out)
        h = out.copy_to_host().astype(np.int64)
        per_batch.append(h)
        out_sum += h
    return out_sum, np.stack(per_batch)


def reference_tpr(hist, target_label):
    with tempfile.TemporaryDirectory() as d:
        roc_cuda.plot_ROC(hist.reshape([-1, 2]).copy(), d, 0, target_label)
        line = open(os.path.join(d, "local_log.txt")).read().splitlines()[1]
    return np.array([float(x) for x in re.search(r"\[(.*)\]", line).group(1).split(",")])


def make_case(rng, n, emb, n_ids, target_size, plant):
    centres = rng.standard_normal((n_ids, emb))
    label = np.sort(rng.integers(0, n_ids, n)).astype(np.int32)      # the reference's caller puts target IDs first
    f = centres[label] + 0.8 * rng.standard_normal((n, emb))
    f = (f / np.linalg.norm(f, axis=1, keepdims=True)).astype(np.float32)
    if plant:                                                         # exact duplicate / exact opposite rows: bins 2000 / 0
        f[5] = f[2]
        label[5] = label[2]
        f[n - 3] = -f[1]
        f[n - 7] = f[3]
    return f, label


def main():
    check_compiled_arithmetic()
    rng = np.random.default_rng(20211)
    cases = {
        "a": dict(n=80, emb=512, n_ids=6, target_size=48, batch_size=32, plant=True),     # two batches, second ragged
        "b": dict(n=70, emb=20, n_ids=4, target_size=70, batch_size=64, plant=False),     # full triangle, emb % 32 != 0
        "c": dict(n=33, emb=7, n_ids=3, target_size=1, batch_size=800, plant=False),      # a single sub row
    }
    blob = {}
    for name, c in cases.items():
        f, label = make_case(rng, c["n"], c["emb"], c["n_ids"], c["target_size"], c["plant"])
        hist, per_batch = run_reference(f, label, c["batch_size"], c["target_size"])
        n, t = c["n"], c["target_size"]
        assert hist.sum() == t * (t - 1) // 2 + t * (n - t)
        blob[f"{name}/feature"], blob[f"{name}/label"] = f, label
        blob[f"{name}/hist"], blob[f"{name}/per_batch"] = hist, per_batch
        blob[f"{name}/batch_size"], blob[f"{name}/target_size"] = np.int64(c["batch_size"]), np.int64(t)
        if name == "a":
            blob[f"{name}/tpr"] = reference_tpr(hist, list(range(int(label[0]), int(label[t - 1]) + 1)))
        print(name, "pairs", int(hist.sum()), "same", int(hist[0::2].sum()), "bins", int((hist > 0).sum()))
    np.savez_compressed(os.path.join(OUT, "roc.npz"), **blob)


if __name__ == "__main__":
    main()
