#!/bin/bash
# One multi-GPU box session: usage (under gpurun --gpus N): bash scripts/gpu_r2_multi.sh N [tag] [steps]
N=${1:-2}; TAG=${2:-r2m$N}; STEPS=${3:-100}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; free -g | head -2 >> $O/${TAG}_gpu.txt; nvidia-smi topo -m >> $O/${TAG}_gpu.txt 2>&1
build/cuinit_probe $N > $O/${TAG}_cuinit.txt 2>&1; CUDA_VISIBLE_DEVICES=0 build/cuinit_probe 1 >> $O/${TAG}_cuinit.txt 2>&1; cat $O/${TAG}_cuinit.txt
# parity of everything that needs more than one GPU in one process
timeout 900 python -m pytest tests -m gpu -q -k "several_devices or search_shards or pattern_set_shards or cut_into_ranges" > $O/${TAG}_pytest_multi.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_multi.log; tail -5 $O/${TAG}_pytest_multi.log
# the bench line at N ranks (what the driver runs)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps $STEPS --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python scripts/bench_summary.py $O/${TAG}_bench.json || tail -40 $O/${TAG}_bench.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
   bench.py --gpus $N --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
head -c 400 $O/${TAG}_bench_reference.json; echo
# the drop-in CLI over N GPUs: 16 GiB file, stock vs GPU-backed with 1 and N devices
python - <<PY
import ctypes as C, sys, torch
sys.path.insert(0, ".")
import bench
from krep_b200 import lib
L = lib.load(); assert L.krep_b200_init(0) == 0
n = 16 << 30
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
print(f"{dt:.3f} s rc={r.returncode} out={r.stdout.decode()[-60:].strip()}")
PY
}
{
ARGS=("-c qzXv9Kpw" "-c -i QzXv" "-c -o -f $P" "-c the"); [ "$N" -ge 4 ] && ARGS=("-c qzXv9Kpw" "-c -o -f $P")
for args in "${ARGS[@]}"; do
  echo "stock            $args : $(run oracle/_ref/krep $args $F) | $(run oracle/_ref/krep $args $F)"
  for D in 1 $N; do
    export KREP_B200_DEVICES=$D
    echo "gpu devices=$D    $args : $(run build/krep_gpu/krep $args $F) | $(run build/krep_gpu/krep $args $F)"
  done
  unset KREP_B200_DEVICES
done
a=$(oracle/_ref/krep -t 1 -o qzXv9Kpw $F | md5sum); b=$(KREP_B200_DEVICES=$N build/krep_gpu/krep -o qzXv9Kpw $F | md5sum); echo "identical -o output on $N devices: $([ "$a" = "$b" ] && echo yes || echo NO)"
a=$(oracle/_ref/krep -t 1 -o -f $P $F | md5sum); b=$(KREP_B200_DEVICES=$N build/krep_gpu/krep -o -f $P $F | md5sum); echo "identical -o -f output on $N devices: $([ "$a" = "$b" ] && echo yes || echo NO)"
echo "# one device, the other GPUs hidden from the driver (KREP_B200_LIMIT_VISIBLE=1)"
for args in "-c qzXv9Kpw"; do
  echo "gpu limit-visible $args : $(KREP_B200_LIMIT_VISIBLE=1 run build/krep_gpu/krep $args $F) | $(KREP_B200_LIMIT_VISIBLE=1 run build/krep_gpu/krep $args $F)"
done
KREP_B200_LIMIT_VISIBLE=1 KREP_B200_TRACE=1 build/krep_gpu/krep -c qzXv9Kpw $F 2>&1 | tail -25
echo "# phase trace, $N devices"
KREP_B200_DEVICES=$N KREP_B200_TRACE=1 build/krep_gpu/krep -c qzXv9Kpw $F 2>&1 | tail -30
} 2>&1 | tee $O/${TAG}_cli_timing.txt
rm -f $F $P
ls $O | grep "^${TAG}"
