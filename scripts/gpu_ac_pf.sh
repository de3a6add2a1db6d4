#!/bin/bash
# Sweep the L2-prefetch knobs of the tri4 kernel; per config: kernel ms + DRAM bytes / L2 hit rate from a light ncu pass.
TAG=$1; O=gpurun_out; mkdir -p $O
for cfg in "0 0" "2 0" "4 0" "8 0" "2 1" "4 1" "8 1"; do
  set -- $cfg; export KREP_B200_AC_PF=$1 KREP_B200_AC_PFMODE=$2
  ms=$(timeout 300 python bench.py --workload multi1000 --steps 20 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('%.3f'%d['roofline']['kernel_ms'])")
  timeout 300 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:k_ac -c 1 --csv --log-file $O/${TAG}_pf_$1_$2.csv \
     python bench.py --workload multi1000 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
  echo "PF=$1 mode=$2 kernel_ms=$ms $(grep -E 'dram__bytes_read|hit_rate' $O/${TAG}_pf_$1_$2.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | tr '\n' ' ')"
done
