#!/bin/bash
# What a krep user sees: the stock CLI vs the same CLI relinked against libkrep_b200.so, whole-process wall time,
# on a corpus file in /dev/shm (page cache).  Usage: bash scripts/gpu_cli_timing.sh [GiB] [tag]
G=${1:-8}; TAG=${2:-cli}; O=gpurun_out; mkdir -p $O
python - <<PY
import ctypes as C, sys, torch
sys.path.insert(0, ".")
import bench
from krep_b200 import lib
L = lib.load(); assert L.krep_b200_init(0) == 0
n = int($G * (1 << 30))
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
print(f"{dt:.3f} s rc={r.returncode} out={r.stdout.decode()[-80:].strip()}")
PY
}
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1) x $(nvidia-smi -L | wc -l), $(nproc) host threads, corpus $G GiB in /dev/shm"
build/cuinit_probe 1
for args in "-c qzXv9Kpw" "-c -o qzXv9Kpw" "-c -i QzXv" "-c -w needleneedle0016" "-c -o -f $P" "-c the"; do
  for bin in oracle/_ref/krep build/krep_gpu/krep; do
    run $bin $args $F > /dev/null   # warm
    echo "$bin $args : $(run $bin $args $F) | $(run $bin $args $F)"
  done
done
echo "# BASELINE configs[0]: krep --algo=bm -t 1 -c the on 100 000 000 bytes of the synthetic corpus (plumbing check)"
head -c 100000000 $F > /dev/shm/krep_cfg1.txt
a=$(oracle/_ref/krep --algo=bm -t 1 -c the /dev/shm/krep_cfg1.txt); b=$(build/krep_gpu/krep --algo=bm -c the /dev/shm/krep_cfg1.txt)
echo "stock: $a   gpu-backed: $b   identical: $([ "${a##*:}" = "${b##*:}" ] && echo yes || echo NO)"
echo "stock  --algo=bm -t 1 -c the : $(run oracle/_ref/krep --algo=bm -t 1 -c the /dev/shm/krep_cfg1.txt)"
echo "gpu    --algo=bm      -c the : $(run build/krep_gpu/krep --algo=bm -c the /dev/shm/krep_cfg1.txt)"
rm -f /dev/shm/krep_cfg1.txt
echo "# phase trace of one GPU-backed run (KREP_B200_TRACE=1)"
KREP_B200_TRACE=1 build/krep_gpu/krep -c qzXv9Kpw $F 2>&1 | tail -40
KREP_B200_TRACE=1 build/krep_gpu/krep -c the $F 2>&1 | tail -12
echo "# variants of the GPU-backed run: krep maps with MAP_POPULATE again / 16 staging threads / pages pre-faulted during start-up"
for v in "KREP_B200_MAP_POPULATE=1" "KREP_B200_COPY_THREADS=16" "KREP_B200_PREFAULT=1"; do
  echo "$v gpu -c qzXv9Kpw : $(env $v python - build/krep_gpu/krep -c qzXv9Kpw $F <<'PY'
import subprocess, sys, time
t0 = time.perf_counter(); r = subprocess.run(sys.argv[1:], capture_output=True); dt = time.perf_counter() - t0
print(f"{dt:.3f} s rc={r.returncode}")
PY
)"
done
echo "# the same with the driver kept initialised by another process (what nvidia-persistenced / any resident CUDA client gives)"
python -c "import torch, time; torch.zeros(1, device='cuda'); time.sleep(120)" &
HOLD=$!; sleep 20
build/cuinit_probe 1
for args in "-c qzXv9Kpw" "-c -i QzXv" "-c -w needleneedle0016" "-c -o -f $P" "-c the"; do
  echo "warm driver: stock $args : $(run oracle/_ref/krep $args $F)"
  echo "warm driver: gpu   $args : $(run build/krep_gpu/krep $args $F) | $(run build/krep_gpu/krep $args $F)"
done
KREP_B200_TRACE=1 build/krep_gpu/krep -c qzXv9Kpw $F 2>&1 | tail -16
KREP_B200_MAP_POPULATE=1 KREP_B200_TRACE=1 build/krep_gpu/krep -c qzXv9Kpw $F 2>&1 | tail -12
echo "warm driver, MAP_POPULATE: gpu -c qzXv9Kpw : $(KREP_B200_MAP_POPULATE=1 run build/krep_gpu/krep -c qzXv9Kpw $F) | 16 threads: $(KREP_B200_COPY_THREADS=16 run build/krep_gpu/krep -c qzXv9Kpw $F)"
kill $HOLD
} 2>&1 | tee $O/${TAG}_timing.txt
rm -f $F $P
