#!/bin/bash
TAG=${1:-r2d}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -4 $O/${TAG}_pytest_gpu.log
for wl in literal8 icase4 multi1000; do
  timeout 300 python bench.py --workload $wl --gib 10 --no-side --no-e2e --no-cpu --steps 50 > $O/${TAG}_q_$wl.json 2> $O/${TAG}_q_$wl.err
  python scripts/bench_summary.py $O/${TAG}_q_$wl.json | head -1 || tail -20 $O/${TAG}_q_$wl.err
done
SECONDS=0; timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$? wall ${SECONDS}s"
python scripts/bench_summary.py $O/${TAG}_bench.json || tail -30 $O/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_literal8_launches.csv \
   python bench.py --workload literal8 --steps 5 --warmup 3 --no-e2e --no-cpu --no-side > $O/${TAG}_literal8_ncu_bench.log 2>&1
grep -E "k_finish|k_lit" $O/${TAG}_literal8_launches.csv | tail -4 | cut -c1-250
