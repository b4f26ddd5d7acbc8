"""CPU tests of bench.py's output contract: the reference arm runs here (it is the reference's CPU
path on the host cores), and the committed B200 bench line carries every key the driver reads."""
import glob
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--queries", "2000000"],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["metric"] == "k-mer lookups/sec" and d["unit"] == "lookups/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e5 and "workload" in d["config"]
    import bench
    assert d["config"] == bench.make_config(2000000, 1)      # the two arms print the same config


def test_committed_b200_line_has_every_contract_key():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9]_bench_v*.json")), key=os.path.getmtime)
    line = [l for l in open(files[-1]) if l.startswith("{")][-1]
    d = json.loads(line)
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline"} <= set(d), sorted(set(d))
    assert d["gpu_launches"] >= d["steps"] > 0 and d["dtype"] == "u64" and d["vs_baseline"] is None
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    if d["n_gpus"] == 1:
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
