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
    # one config builder for both arms (the driver compares the two dicts), and the reference's own configuration -
    # ONE worker (nginx.conf:2) - is what `value` reports; the all-core rate sits beside it
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.c2_config(64 << 20, 64, 1)
    assert d["cpu_baseline"]["cores"] == 1 and d["all_cores"]["cores"] >= 1 and d["all_cores"]["value"] > 0


def test_round2_default_line_carries_every_config():
    """The committed round-2 default line (one B200): C2 headline + C1 / C3 / C5 / HTTP sections; the N=2 line adds the span
    section (C4) at the named 125 GiB per GPU and HTTP through `nzproxy --devices`."""
    d = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_default_a.json")).read().strip().splitlines()[-1])
    check_line(d)
    assert d["config"]["request_seed"] == 22 and d["cpu_baseline"]["cores"] == 1
    assert d["c1"]["object_bytes"] == 169_001_437 and d["c1"]["reference_published_mbs"] == 323 and d["c1"]["c_abi_host_mbs"] > 323
    assert d["http"]["c1"]["gpu"]["status_2xx"] > 0 and d["http"]["gpu"]["errors"] == 0
    assert "C3" in d["c3"]["config"]["workload"] and d["c3"]["value"] > 150 and d["c3"]["c3"]["single_stream_ms"] < 3
    assert "C5" in d["c5"]["config"]["workload"] and len(d["c5"]["sweep"]) == 6
    d2 = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_default_n2.json")).read().strip().splitlines()[-1])
    check_line(d2)
    sp = d2["span"]
    assert d2["n_gpus"] == 2 and sp["config"]["arena_gib_per_gpu"] == 125.0 and 0.5 < sp["ingress_frac_of_peak"] < 1.05
    assert len(sp["per_rank"]) == 2 and sp["nccl_allgather"]["ms"] > 0
    assert d2["http"]["gpu"]["devices"] == 2 and d2["http"]["gpu"]["errors"] == 0
