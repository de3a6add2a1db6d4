#!/bin/bash
# r2c: A/B of the window-kernel variants, full tests, default bench line, launch list, CLI timing.
TAG=${1:-r2c}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -4 $O/${TAG}_pytest_gpu.log
one() { # name lib workload
  KREP_B200_LIB=$2 timeout 300 python bench.py --workload $3 --gib 10 --no-side --no-e2e --no-cpu --steps 30 > $O/${TAG}_v_$1_$3.json 2> $O/${TAG}_v_$1_$3.err
  python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_v_$1_$3.json")); r = d["roofline"]
    print("variant $1 $3: kernel_ms %.4f achieved %.0f frac %.3f value %.0f ms/step %.4f matches %d" % (r["kernel_ms"], r["achieved"], r["frac"], d["value"], d["ms_per_step"], d["matches"]))
except Exception as e:
    print("variant $1 $3 FAILED", e); print(open("$O/${TAG}_v_$1_$3.err").read()[-1500:])
PY
}
one default "" icase4
for v in nx0 nx1 nx0_lane_emit nx2_lane_emit; do one $v build/variants/libkrep_b200_$v.so icase4; done
one default "" literal8
one nx0_lane_emit build/variants/libkrep_b200_nx0_lane_emit.so literal8
one default "" the_1k_c
one default "" the_64_c
one default "" the_1k
SECONDS=0; timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$? wall ${SECONDS}s"
python scripts/bench_summary.py $O/${TAG}_bench.json || tail -30 $O/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_literal8_launches.csv \
   python bench.py --workload literal8 --steps 5 --warmup 3 --no-e2e --no-cpu --no-side > $O/${TAG}_literal8_ncu_bench.log 2>&1
grep -E "k_finish|k_lit" $O/${TAG}_literal8_launches.csv | tail -6
bash scripts/gpu_cli_timing.sh 8 ${TAG}_cli > /dev/null 2>&1; cat $O/${TAG}_cli_timing.txt | cut -c1-220
