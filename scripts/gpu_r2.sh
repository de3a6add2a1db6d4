#!/bin/bash
# One single-GPU box session of round 2: parity tests, smoke, the default bench line (headline + side workloads),
# reference arm, CLI timing.  Usage (under gpurun): bash scripts/gpu_r2.sh [tag] [steps: tests|bench|cli|ncu ...]
TAG=${1:-r2a}; shift; WHAT=${@:-tests bench cli}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> $O/${TAG}_gpu.txt; free -g | head -2 >> $O/${TAG}_gpu.txt
for w in $WHAT; do case $w in
tests)
  timeout 1800 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
  tail -15 $O/${TAG}_pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log; tail -2 $O/${TAG}_smoke.log ;;
bench)
  SECONDS=0; timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$? wall ${SECONDS}s"
  python scripts/bench_summary.py $O/${TAG}_bench.json || tail -30 $O/${TAG}_bench.err
  timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
  head -c 600 $O/${TAG}_bench_reference.json; echo ;;
cli)
  bash scripts/gpu_cli_timing.sh 8 ${TAG}_cli ;;
sanitize)
  bash scripts/gpu_sanitize.sh ;;
ncu)
  for wl in literal8 icase4 multi1000; do
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_${wl}_launches.csv \
       python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu --no-side > $O/${TAG}_${wl}_ncu_bench.log 2>&1
  done
  for spec in "literal8 k_lit_aligned4" "icase4 k_lit_window4" "multi1000 k_ac"; do set -- $spec
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -o $O/${TAG}_$1_full -f \
       python bench.py --workload $1 --steps 1 --warmup 3 --no-e2e --no-cpu --no-side > $O/${TAG}_$1_ncu_full.log 2>&1
  done ;;
esac; done
ls $O | grep "^${TAG}" | head -40
