#!/bin/bash
# Quick GPU iteration: parity tests, then bench + ncu for the named workloads.
# Usage: bash scripts/gpu_quick.sh TAG "workload ..." [ncu-kernel-regex]
TAG=$1; WLS=${2:-"multi1000"}; KRE=${3:-k_ac}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -5 $O/${TAG}_pytest_gpu.log
for wl in $WLS; do
  timeout 600 python bench.py --workload $wl --steps 30 --no-cpu > $O/${TAG}_bench_$wl.json 2> $O/${TAG}_bench_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench_$wl.json")); r=d["roofline"]
    print("$wl", "value %.0f"%d["value"], "kernel_ms %.3f"%r["kernel_ms"], "achieved %.0f"%r["achieved"], "frac %.3f"%r["frac"], "matches", d["matches"], "e2e %.1f"%d["e2e"]["value"], d["config"]["filter"])
except Exception as e:
    print("$wl FAILED", e); print(open("$O/${TAG}_bench_$wl.err").read()[-2000:])
PY
done
if [ -n "$KRE" ]; then
  wl=$(echo $WLS | awk '{print $1}')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -c 1 -o $O/${TAG}_${wl}_full -f \
     python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu > $O/${TAG}_${wl}_ncu_full.log 2>&1
fi
