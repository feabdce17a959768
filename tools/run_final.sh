#!/bin/bash
# end-of-round validation on one B200: suite, smoke, default bench line, ncu launch list + full capture of the gather, sanitizer on the new paths
mkdir -p gpurun_out
timeout 200 ./nezha_b200/_build/probe_scatter_write 32 only32 2>&1 | grep "copy \|bulk \|ahead" | tee gpurun_out/r2_probe_scatter_ahead.txt
(timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -6) | tee gpurun_out/r2_gpu_suite_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 800 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default_final.json 2> gpurun_out/r2_bench_default_final.err; tail -c 300 gpurun_out/r2_bench_default_final.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-http --no-sections > gpurun_out/r2_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_gather_tma -s 4 -c 1 -o gpurun_out/r02_k_gather_tma python bench.py --steps 2 --warmup 3 --no-cpu --no-http --no-sections --no-e2e > gpurun_out/r2_ncu_gather.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 9 python -m pytest tests/test_gpu_span.py tests/test_gpu_parity.py tests/test_gpu_inflate.py -m gpu -q -x -k "striped or packed or recycled or crc32c_device or coalesced_identical" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck.log; tail -4 gpurun_out/r2_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_span.py tests/test_gpu_parity.py -m gpu -q -x -k "striped_object_single_gpu or crc32c_device" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck.log; tail -4 gpurun_out/r2_sanitizer_racecheck.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_default_final.json").read().strip().splitlines()[-1])
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "http", d["http"]["gpu"]["gbs"], d["http"]["gpu"]["ttfb_us_p50"], d["http"]["gpu"]["ttfbody_us_p50"])
print("c1", d["c1"]["c_abi_host_mbs"], "c3", d["c3"].get("value"), (d["c3"].get("c3") or {}).get("single_stream_ms"), "c5 fill", (d["c5"].get("fill") or {}).get("gbs"))
PY
