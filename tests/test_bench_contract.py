"""bench.py contract pieces that run without a GPU: the reference arm prints exactly one JSON line on stdout (library
chatter goes to stderr), with the keys the driver reads; ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, env=env, timeout=600)


# what a torchrun worker inherits when --nproc-per-node > 1: the agent-store switch (turns every TCP rendezvous into a client
# of the agent's store -- the reference's private 1-rank group must not depend on it) and one OpenMP thread per worker
TORCHRUN_ENV = {"TORCHELASTIC_USE_AGENT_STORE": "True", "TORCHELASTIC_RESTART_COUNT": "0", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "1",
                "LOCAL_RANK": "0", "OMP_NUM_THREADS": "1"}


def test_reference_arm_prints_one_json_line():
    r = _run(dict(TORCHRUN_ENV, RANK="0", WORLD_SIZE="2"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True and d["n_gpus"] == 2
    # "reference": the unmodified class staged in baseline/_ref (or /root/reference) ran; "port": those files are absent
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "1000000 classes" in d["cpu_baseline"]["sample"] and "full size" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "c3" in d["config"]["workload"] and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)          # rank 0 has the box to itself: OMP_NUM_THREADS=1 is overridden


def test_reference_arm_other_ranks_are_silent():
    r = _run(dict(TORCHRUN_ENV, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""
