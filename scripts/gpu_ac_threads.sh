#!/bin/bash
# tri4 kernel under both CTA sizes: parity tests, bench, full ncu capture.
TAG=$1; O=gpurun_out; mkdir -p $O
for th in 640 768; do
  export KREP_B200_AC_THREADS=$th
  timeout 300 python -m pytest tests -m gpu -x -q -k "aho or multi or shard" > $O/${TAG}_pytest_$th.log 2>&1; echo "$th pytest rc=$? $(tail -1 $O/${TAG}_pytest_$th.log)"
  timeout 300 python bench.py --workload multi1000 --steps 30 --no-cpu --no-e2e > $O/${TAG}_bench_$th.json 2> $O/${TAG}_bench_$th.err
  python -c "
import json; d=json.load(open('$O/${TAG}_bench_$th.json')); r=d['roofline']; print('$th kernel_ms %.3f achieved %.0f frac %.3f matches %d'%(r['kernel_ms'],r['achieved'],r['frac'],d['matches']))" || tail -5 $O/${TAG}_bench_$th.err
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ac -c 1 -o $O/${TAG}_${th}_full -f \
     python bench.py --workload multi1000 --steps 1 --warmup 3 --no-e2e --no-cpu > $O/${TAG}_${th}_ncu.log 2>&1
done
