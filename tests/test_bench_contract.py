"""CPU: the bench line the driver parses.  The committed end-of-round line (profiles/r01_bench_c2_head.json,
printed by `python bench.py` on a B200) must carry every key of the contract with sane types, and
`bench.py --impl reference` (the CPU arm; it needs no GPU) must print the same shape here."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check_line(d, reference=False):
    for k, t in (("metric", str), ("value", (int, float)), ("unit", str), ("n_gpus", int), ("steps", int), ("warmup", int),
                 ("ms_per_step", (int, float)), ("higher_is_better", bool), ("scaling", str), ("dtype", str), ("data", str), ("config", dict)):
        assert k in d and isinstance(d[k], t), k
    assert "vs_baseline" in d and d["vs_baseline"] is None          # BASELINE.md publishes no number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["scaling"] == "weak" and d["higher_is_better"] is True and d["dtype"] == "u8" and d["data"] == "synthetic"
    e = d["e2e"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(e) and e["unit"] == d["unit"]
    cb = d["cpu_baseline"]
    if cb is not None:
        assert set(("value", "unit", "cores", "kind", "sample")) <= set(cb) and cb["kind"] in ("port", "reference")
    if reference:
        assert d["impl"] == "reference" and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0 and e["value"] == d["value"]
        return
    r = d["roofline"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6 and 0 < r["frac"] < 1.05
    assert d["gpu_launches"] >= d["steps"] > 0 and d["warmup"] >= 3
    c = d["clocks"]
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(c)
    assert not (set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"})
    assert e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]            # a real end-to-end number, not the kernel figure again


def test_committed_bench_line_follows_the_contract():
    d = json.loads(open(os.path.join(ROOT, "profiles", "r01_bench_c2_head.json")).read().strip().splitlines()[-1])
    check_line(d)
    assert d["n_gpus"] == 1 and "C2" in d["config"]["workload"]
    d2 = json.loads(open(os.path.join(ROOT, "profiles", "r01_bench_c2_n2_head.json")).read().strip().splitlines()[-1])
    check_line(d2)
    assert d2["n_gpus"] == 2 and d2["value"] > 1.7 * d["value"] * 0.9           # weak scaling, no collective


@pytest.mark.timeout(600)
def test_reference_arm_prints_the_same_shape_on_cpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--object-bytes", str(64 << 20), "--cpu-budget", "1"], capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-800:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    check_line(d, reference=True)
