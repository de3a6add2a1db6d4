#!/bin/bash
# One GPU-box session: parity tests, smoke, bench lines for every workload, ncu launch lists + full captures.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> $O/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench_literal8.json 2> $O/${TAG}_bench_literal8.err
timeout 600 python bench.py --workload multi1000 --steps 20 > $O/${TAG}_bench_multi1000.json 2> $O/${TAG}_bench_multi1000.err
timeout 600 python bench.py --workload icase4 --steps 50 > $O/${TAG}_bench_icase4.json 2> $O/${TAG}_bench_icase4.err
timeout 600 python bench.py --workload word16 --steps 50 > $O/${TAG}_bench_word16.json 2> $O/${TAG}_bench_word16.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
for wl in literal8 multi1000 icase4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_${wl}_launches.csv \
     python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu > $O/${TAG}_${wl}_ncu_bench.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lit_aligned4 -c 1 -o $O/${TAG}_literal8_full -f \
   python bench.py --workload literal8 --steps 1 --warmup 3 --no-e2e --no-cpu > $O/${TAG}_literal8_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ac -c 1 -o $O/${TAG}_multi1000_full -f \
   python bench.py --workload multi1000 --steps 1 --warmup 3 --no-e2e --no-cpu > $O/${TAG}_multi1000_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lit_window4 -c 1 -o $O/${TAG}_icase4_full -f \
   python bench.py --workload icase4 --steps 1 --warmup 3 --no-e2e --no-cpu > $O/${TAG}_icase4_ncu_full.log 2>&1
ls -la $O
