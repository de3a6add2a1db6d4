#!/bin/bash
TAG=${1:-r2g}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -3 $O/${TAG}_pytest_gpu.log
SECONDS=0; timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$? wall ${SECONDS}s"
python scripts/bench_summary.py $O/${TAG}_bench.json || tail -30 $O/${TAG}_bench.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/${TAG}_literal8_launches.csv \
   python bench.py --workload literal8 --steps 3 --warmup 3 --no-e2e --no-cpu --no-side > $O/${TAG}_literal8_ncu_bench.log 2>&1
grep -E "k_finish" $O/${TAG}_literal8_launches.csv | tail -2 | cut -c150-260
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_count_lines -c 1 -o $O/${TAG}_the_1k_c_k_count_lines_full -f \
   python bench.py --workload the_1k_c --steps 1 --warmup 3 --no-e2e --no-cpu --no-side > $O/${TAG}_count_ncu_full.log 2>&1
