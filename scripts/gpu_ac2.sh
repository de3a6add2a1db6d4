#!/bin/bash
# multi-pattern regimes: parity tests + bench for the three pattern-set workloads + one ncu capture
TAG=$1; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "aho or multi or shard or fixtures or vectors" > $O/${TAG}_pytest_ac.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${TAG}_pytest_ac.log)"
for wl in multi1000 multi1000_8to14 multi1000_i; do
  timeout 400 python bench.py --workload $wl --steps 30 --no-cpu --no-e2e > $O/${TAG}_bench_$wl.json 2> $O/${TAG}_bench_$wl.err
  python -c "
import json; d=json.load(open('$O/${TAG}_bench_$wl.json')); r=d['roofline']; print('$wl value %.0f kernel_ms %.3f achieved %.0f frac %.3f matches %d'%(d['value'],r['kernel_ms'],r['achieved'],r['frac'],d['matches']), d['config']['filter'])" || tail -5 $O/${TAG}_bench_$wl.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ac -c 1 -o $O/${TAG}_8to14_full -f \
   python bench.py --workload multi1000_8to14 --steps 1 --warmup 3 --no-e2e --no-cpu > $O/${TAG}_8to14_ncu.log 2>&1
