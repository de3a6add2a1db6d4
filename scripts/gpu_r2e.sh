#!/bin/bash
TAG=${1:-r2e}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -4 $O/${TAG}_pytest_gpu.log
one() { # name lib workload
  KREP_B200_LIB=$2 timeout 300 python bench.py --workload $3 --gib 10 --no-side --no-e2e --no-cpu --steps 30 > $O/${TAG}_v_$1_$3.json 2> $O/${TAG}_v_$1_$3.err
  echo -n "variant $1 $3: "; python scripts/bench_summary.py $O/${TAG}_v_$1_$3.json | head -1 || tail -20 $O/${TAG}_v_$1_$3.err
}
for wl in literal8 icase4 multi1000 the_1k the_1k_c the_64_c; do one default "" $wl; done
one count_minb2 build/variants/libkrep_b200_count_minb2.so the_1k_c
one count_minb2 build/variants/libkrep_b200_count_minb2.so the_64_c
SECONDS=0; timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$? wall ${SECONDS}s"
python scripts/bench_summary.py $O/${TAG}_bench.json || tail -30 $O/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_literal8_launches.csv \
   python bench.py --workload literal8 --steps 5 --warmup 3 --no-e2e --no-cpu --no-side > $O/${TAG}_literal8_ncu_bench.log 2>&1
grep -E "k_finish|k_lit" $O/${TAG}_literal8_launches.csv | tail -4 | cut -c90-260
