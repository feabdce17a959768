#!/bin/bash
# One 8-GPU lease: what bounds the e2e leg, the span tile order at 8 GPUs, the default bench line, the front end.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(cat /sys/fs/cgroup/cpu.max; nproc; lscpu | grep -i "model name\|socket(s)\|numa node"; free -g | head -2) > gpurun_out/r2_n8_host.txt 2>&1
cat gpurun_out/r2_n8_host.txt
timeout 200 $TR --nproc-per-node 8 --master-port 29601 tools/probe_d2h.py 2>/dev/null | grep "^{" > gpurun_out/r2_probe_d2h_n8.txt
timeout 200 $TR --nproc-per-node 4 --master-port 29602 tools/probe_d2h.py 2>/dev/null | grep "^{" > gpurun_out/r2_probe_d2h_n4.txt
cat gpurun_out/r2_probe_d2h_n8.txt gpurun_out/r2_probe_d2h_n4.txt | cut -c1-260
for cfg in "1 64" "1 256" "0 64"; do
  set -- $cfg
  NZ_SPAN_DYN=$1 NZ_SPAN_SPREAD=$2 timeout 300 $TR --nproc-per-node 8 --master-port 29603 bench.py --workload c4 --c4-gib-per-gpu 24 --steps 10 --warmup 3 > gpurun_out/r2_c4_n8_dyn$1_sp$2.json 2> gpurun_out/r2_c4_n8.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_c4_n8_dyn$1_sp$2.json").read().strip().splitlines()[-1]); s=d["span"]
    print("dyn $1 spread $2: value", round(d["value"],1), "busiest ingress", round(s["busiest_rank_ingress_gbs"],1), "frac", round(s["ingress_frac_of_peak"],3), "max", round(s["max_rank_ingress_gbs"],1), [round(p["gather_ms"],2) for p in s["per_rank"]], "nccl", round(s["nccl_allgather"].get("ingress_gbs_per_gpu",0),1))
except Exception as e:
    print("dyn $1 spread $2 failed", e); print(open("gpurun_out/r2_c4_n8.err").read()[-600:])
PY
done 2>&1 | tee gpurun_out/r2_c4_n8_ab.txt
NZ_BENCH_HTTP_GPU_ONLY=1 timeout 600 $TR --nproc-per-node 8 --master-port 29604 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -c 500 gpurun_out/r2_bench_n8.err
NZ_PROXY_IO_THREADS=16 NZ_BENCH_HTTP_GPU_ONLY=1 timeout 400 $TR --nproc-per-node 8 --master-port 29605 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --no-span --no-sections > gpurun_out/r2_bench_n8_epoll16.json 2> gpurun_out/r2_bench_n8_epoll16.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_n8.json", "gpurun_out/r2_bench_n8_epoll16.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "no line", e); continue
    sp = d.get("span") or {}
    print(f, "value", round(d["value"], 1), "frac", round(d["roofline"]["frac"], 4), "e2e", d.get("e2e") and round(d["e2e"]["value"], 1))
    print("  span", {k: sp.get(k) for k in ("value", "busiest_rank_ingress_gbs", "max_rank_ingress_gbs", "ingress_frac_of_peak", "fill_s", "error")})
    h = d.get("http") or {}
    print("  http", {k: (h.get("gpu") or {}).get(k) for k in ("gbs", "ttfb_us_p50", "ttfbody_us_p50", "ttfbody_us_p99", "errors", "conns", "devices")}, h.get("error"))
    print("  c5_fill", d.get("c5_fill"))
PY
