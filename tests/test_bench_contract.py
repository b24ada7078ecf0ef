"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, the SAME
`config` dict our arm prints (the driver compares them), the steps / warm-ups it really timed, and uses every host thread
even when launched the way torchrun launches workers (OMP_NUM_THREADS=1 in the environment)."""
import json
import os
import subprocess
import sys

from util import ROOT


def test_reference_arm_line():
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["config"] == bench.CONFIG and "workload" in d["config"]
    assert (d["steps"], d["warmup"], d["n_gpus"]) == (2, 1, 2)
    assert d["scaling"] == "strong" and d["dtype"] == "f32" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == len(os.sched_getaffinity(0)) and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] * 1e9 - (3 * 32768 * 16384 * 4 + 8 + 16384 * 4)) < 1e7


def test_other_ranks_of_the_reference_arm_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
