"""__graft_entry__.build() is called by every rank of a torchrun job at once: the build step must be content-hashed
(file times do not survive the snapshot to the GPU box), locked and atomic.  Exercised here with a tiny C library."""
import ctypes
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(args):
    target, src, digest = args
    import __graft_entry__ as g
    g._build_locked(target, target + ".srchash", digest, ["gcc", "-O0", "-shared", "-fPIC", src], os.path.dirname(src), False)
    return ctypes.CDLL(target).answer()


def test_concurrent_builds_produce_one_valid_library(tmp_path):
    import __graft_entry__ as g
    src = tmp_path / "tiny.c"
    src.write_text("int answer(void) { return 42; }\n")
    target = str(tmp_path / "lib" / "libtiny.so")
    digest = g._digest([str(src)], "flags")
    with mp.get_context("spawn").Pool(4) as pool:
        assert pool.map(_worker, [(target, str(src), digest)] * 4) == [42] * 4
    assert open(target + ".srchash").read().strip() == digest
    assert not [f for f in os.listdir(os.path.dirname(target)) if ".tmp." in f]
    # fresh: a second round does not rebuild; a changed source changes the digest
    mtime = os.path.getmtime(target)
    assert _worker((target, str(src), digest)) == 42 and os.path.getmtime(target) == mtime
    src.write_text("int answer(void) { return 43; }\n")
    assert g._digest([str(src)], "flags") != digest
