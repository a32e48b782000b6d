"""bench.py pieces that run without a GPU: the reference arm's JSON line (the CPU oracle port timed through bench.py) and
the workload helpers."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_prints_one_contract_line():
    out = _run(["--impl", "reference", "--steps", "3", "--warmup", "1", "--users", "20000", "--items", "5000", "--dim", "32",
                "--batch", "512"])
    lines = [l for l in out.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "apr_train_triples_per_s" and j["unit"] == "triples/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["value"] > 0 and j["vs_baseline"] is None and j["higher_is_better"] is True
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in j["config"] and "model" not in j["config"]


def test_reference_arm_other_ranks_exit_quietly():
    out = _run(["--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert out.strip() == ""


def test_zipf_items_is_bounded_and_heavy_headed():
    sys.path.insert(0, ROOT)
    import bench
    rng = np.random.default_rng(0)
    x = bench.zipf_items(rng, (4, 50000), 100000)
    assert x.dtype == np.int32 and x.shape == (4, 50000)
    assert x.min() >= 0 and x.max() < 100000
    share0 = float((x == 0).mean())
    h = (np.arange(1, 100001, dtype=np.float64) ** -1.05).sum()
    assert abs(share0 - 1.0 / h) < 0.01          # p(0) = 1 / H(N, 1.05)
    assert (x < 100).mean() > 0.3                 # the head carries a large share of the mass


def test_measured_peaks_are_used_when_present():
    sys.path.insert(0, ROOT)
    import bench
    hbm, tc, kind = bench.peaks()
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        m = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        assert kind == "measured" and hbm == m["hbm_gbs"] and tc == m["bf16_tflops"]
        assert bench.sustained_tc_peak(tc) == m.get("bf16_tflops_sustained", tc)
    else:
        assert kind == "fallback"
