#!/bin/bash
# The round's final single-GPU session: parity tests, smoke, default bench line + reference arm, launch lists and
# ncu --set full captures of the hot kernels, short CLI timing.  Usage (under gpurun): bash scripts/gpu_r2_final.sh [tag]
TAG=${1:-r2f}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1; nproc >> $O/${TAG}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -3 $O/${TAG}_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log; tail -2 $O/${TAG}_smoke.log
SECONDS=0; timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$? wall ${SECONDS}s"
python scripts/bench_summary.py $O/${TAG}_bench.json || tail -30 $O/${TAG}_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
for wl in literal8 icase4 multi1000; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/${TAG}_${wl}_launches.csv \
     python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu --no-side > $O/${TAG}_${wl}_ncu_bench.log 2>&1
done
for spec in "literal8 k_lit_aligned4" "icase4 k_lit_window4" "multi1000 k_ac_tri4" "the_1k_c k_count_lines" "literal8 k_finish"; do set -- $spec
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -o $O/${TAG}_$1_$2_full -f \
     python bench.py --workload $1 --steps 1 --warmup 3 --no-e2e --no-cpu --no-side > $O/${TAG}_$1_$2_ncu_full.log 2>&1
done
# CLI: one cold run of each command, the trace, then the same with the driver kept warm
python - <<PY
import ctypes as C, sys, torch
sys.path.insert(0, ".")
import bench
from krep_b200 import lib
L = lib.load(); assert L.krep_b200_init(0) == 0
n = 8 << 30
spec = lib.make_spec(bench.SEED, bench.PLANT_SEED, 1 << 20, b"qzXv9Kpw", 0)
t = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
L.krep_b200_corpus_generate(C.byref(spec), t.data_ptr(), 0, n, None)
t[:n].cpu().numpy().tofile("/dev/shm/krep_cli_corpus.txt")
open("/dev/shm/krep_cli_pats.txt", "wb").write(b"\n".join(bench.multi_patterns(1000, b"kqzvxjwpy")) + b"\n")
PY
F=/dev/shm/krep_cli_corpus.txt; P=/dev/shm/krep_cli_pats.txt
run() { python - "$@" <<'PY'
import subprocess, sys, time
t0 = time.perf_counter(); r = subprocess.run(sys.argv[1:], capture_output=True); dt = time.perf_counter() - t0
print(f"{dt:.3f} s rc={r.returncode} out={r.stdout.decode()[-40:].strip()}")
PY
}
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1) x $(nvidia-smi -L | wc -l), $(nproc) host threads, 8 GiB corpus in /dev/shm; CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-unset}"
build/cuinit_probe 1
for args in "-c qzXv9Kpw" "-c -i QzXv" "-c -w needleneedle0016" "-c -o -f $P" "-c the"; do
  echo "cold: stock $args : $(run oracle/_ref/krep $args $F)"
  echo "cold: gpu   $args : $(run build/krep_gpu/krep $args $F)"
done
KREP_B200_TRACE=1 build/krep_gpu/krep -c qzXv9Kpw $F 2>&1 | tail -14
python -c "import torch, time; torch.zeros(1, device='cuda'); time.sleep(70)" &
HOLD=$!; sleep 15
build/cuinit_probe 1
for args in "-c qzXv9Kpw" "-c -i QzXv" "-c -w needleneedle0016" "-c -o -f $P" "-c the"; do
  echo "warm driver: stock $args : $(run oracle/_ref/krep $args $F)"
  echo "warm driver: gpu   $args : $(run build/krep_gpu/krep $args $F) | $(run build/krep_gpu/krep $args $F)"
done
KREP_B200_TRACE=1 build/krep_gpu/krep -c qzXv9Kpw $F 2>&1 | tail -14
kill $HOLD
} 2>&1 | tee $O/${TAG}_cli_timing.txt | cut -c1-200
rm -f $F $P
ls $O | grep "^${TAG}" | wc -l
