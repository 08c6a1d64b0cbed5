"""bench.py's output contract, checked without a GPU: the GPU arm runs against a stand-in library
(tests/support/fake_bench.py) so that every key the driver reads is present and well-formed; the
reference arm runs for real on a tiny sample."""
import json
import os
import subprocess
import sys

import pytest

import mpsort_oracle as O
from conftest import ROOT

BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"]


def run(args):
    rc = subprocess.run([sys.executable] + args, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert rc.returncode == 0, rc.stderr.decode()[-2000:]
    lines = [l for l in rc.stdout.decode().splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    return json.loads(lines[0])


@pytest.mark.parametrize("extra", [[], ["--pinfail"]])
def test_gpu_arm_line_has_every_contract_key(extra):
    d = run([os.path.join(ROOT, "tests", "support", "fake_bench.py")] + extra)
    for k in BASE_KEYS + ["clocks", "roofline"]:
        assert k in d, k
    assert d["metric"] == "sorted records/s" and d["unit"] == "records/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "u64" and d["data"] == "synthetic"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] == d["config"]["records_per_gpu"] * d["config"]["elsize"]
    assert ("pageable" if extra else "pinned") in d["e2e"]["api"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_reference_arm_line():
    d = run([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-records", "32768"])
    for k in BASE_KEYS:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "sorted records/s" and d["unit"] == "records/s"
    assert d["e2e"] == {"value": d["value"], "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] == d["value"] and "sample" in c
    assert d["config"]["workload"].startswith("uniform16: 2^28")
    assert d["same_workload_as_gpu_arm"] is False and d["config"]["records_per_step"] == 32768


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_reference_arm_sorts_the_gpu_arms_workload_at_one_gpu():
    """--gpus 1 without --ref-records: every sort is the GPU arm's full workload (here 2^16 records), the value is the
    MEAN over the timed sorts and `config` is the GPU arm's `config`, key for key"""
    d = run([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "2", "--log2n", "16"])
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert d["config"] == b.make_config("uniform16", 16, 1)
    assert d["same_workload_as_gpu_arm"] is True and d["sorts_timed"] == 4 and len(d["seconds_per_sort"]) == 5
    mean = sum(d["seconds_per_sort"][1:]) / 4
    assert abs(d["ms_per_step"] - mean * 1e3) < 1e-3 * d["ms_per_step"] + 1e-3
    assert abs(d["value"] - (1 << 16) / mean) < 2e-3 * d["value"]


def test_reference_arm_on_other_ranks_is_silent():
    env = dict(os.environ, RANK="3", WORLD_SIZE="8")
    rc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8"], env=env,
                        stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert rc.returncode == 0 and rc.stdout.strip() == b""


def test_gpu_arm_line_at_eight_ranks():
    """rank 0 of 8 (the other ranks' answers are copies of its own): whole-job value and e2e byte counts"""
    d = run([os.path.join(ROOT, "tests", "support", "fake_bench.py"), "--size", "8"])
    assert d["n_gpus"] == 8 and d["cpu_baseline"] is None
    per_gpu = d["config"]["records_per_gpu"]
    assert abs(d["value"] - 8 * per_gpu / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * per_gpu * d["config"]["elsize"]
    assert "NVLink" in d["transport"] and d["config"]["baseline_config"].startswith("configs[2]")
    assert set(d["workloads"]) == {"mostly_sorted16", "particles48"}
    for w in d["workloads"].values():
        assert w["value"] > 0 and "kernels" in w and "exchange" in w and "config" in w
    assert d["exchange"]["gb_per_s_per_gpu"] > 0
