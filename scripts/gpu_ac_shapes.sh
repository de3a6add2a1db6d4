#!/bin/bash
# Compare CTA shapes of the tri4 kernel: parity tests under each shape, bench, one ncu capture per shape.
TAG=$1; O=gpurun_out; mkdir -p $O
for shape in 640x2 640x1 768x1; do
  export KREP_B200_AC_SHAPE=$shape
  timeout 300 python -m pytest tests -m gpu -x -q -k "aho or multi" > $O/${TAG}_pytest_$shape.log 2>&1; echo "$shape pytest rc=$? $(tail -1 $O/${TAG}_pytest_$shape.log)"
  timeout 300 python bench.py --workload multi1000 --steps 30 --no-cpu --no-e2e > $O/${TAG}_bench_$shape.json 2> $O/${TAG}_bench_$shape.err
  python -c "
import json; d=json.load(open('$O/${TAG}_bench_$shape.json')); r=d['roofline']; print('$shape kernel_ms %.3f achieved %.0f frac %.3f matches %d'%(r['kernel_ms'],r['achieved'],r['frac'],d['matches']))" || tail -5 $O/${TAG}_bench_$shape.err
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ac -c 1 -o $O/${TAG}_${shape}_full -f \
     python bench.py --workload multi1000 --steps 1 --warmup 3 --no-e2e --no-cpu > $O/${TAG}_${shape}_ncu.log 2>&1
done
