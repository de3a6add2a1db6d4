#!/bin/bash
# AC-kernel iteration: multi-pattern parity tests, bench, ncu capture.  Usage: bash scripts/gpu_ac.sh TAG
TAG=$1; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "aho or multi or shard or fixtures or vectors" > $O/${TAG}_pytest_ac.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_ac.log
tail -3 $O/${TAG}_pytest_ac.log
timeout 600 python bench.py --workload multi1000 --steps 30 --no-cpu --no-e2e > $O/${TAG}_bench_multi1000.json 2> $O/${TAG}_bench_multi1000.err
python -c "
import json; d=json.load(open('$O/${TAG}_bench_multi1000.json')); r=d['roofline']; print('multi1000 value %.0f kernel_ms %.3f achieved %.0f frac %.3f matches %d'%(d['value'],r['kernel_ms'],r['achieved'],r['frac'],d['matches']), d['config']['filter'])" || tail -20 $O/${TAG}_bench_multi1000.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ac -c 1 -o $O/${TAG}_multi1000_full -f \
   python bench.py --workload multi1000 --steps 1 --warmup 3 --no-e2e --no-cpu > $O/${TAG}_multi1000_ncu_full.log 2>&1
