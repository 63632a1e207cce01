"""CPU checks of the bench.py contract: the reference arm (the reference's CPU custom ops + the oracle port on the
host cores) prints ONE JSON line with the keys the driver reads, and the bench lines committed under profiles/ --
what the last GPU run of the round printed -- carry the roofline / cpu_baseline / e2e / clocks objects."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _lines(path):
    return [json.loads(x) for x in open(path).read().splitlines() if x.strip().startswith("{")]


def test_reference_arm_prints_one_contract_line(O):
    if not O.ref_cpu_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--workload", "1080p-stab-files"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(out) == 1, r.stdout           # stdout carries only the JSON line
    d = json.loads(out[0])
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "stabilized_frames_per_sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d.get("gpu_launches", 0) == 0


@pytest.mark.parametrize("name", ["bench_r1_final_1080p.json", "bench_r1_final_4k.jsonl"])
def test_committed_bench_lines_carry_the_contract(name):
    path = os.path.join(ROOT, "profiles", name)
    rows = _lines(path)
    assert rows, path
    for d in rows:
        assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
        assert d["metric"] == "stabilized_frames_per_sec" and d["dtype"] == "f32" and d["data"] == "synthetic"
        assert d["steps"] >= 1 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
        assert d["gpu_launches"] > 0
        e = d["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
        assert e["value"] != d["value"]      # the end-to-end number is measured, not copied
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert r["traffic"] is None or r["traffic"] > 0
        assert 0 < r["fused_stage_a"]["frac"] < 1.2 and 0.9 < r["unblocked_sweep"]["frac"] < 1.5
        c = d["clocks"]
        assert c["sm_mhz"] and c["sm_max_mhz"] and not (set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                             "sw_thermal_slowdown"})
        assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    if name.endswith("1080p.json"):
        cb = rows[0]["cpu_baseline"]
        assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["unit"] == "frames/s"
