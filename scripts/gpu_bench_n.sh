#!/bin/bash
# bench.py at N ranks, as the driver launches it.  Usage (under gpurun --gpus N): bash scripts/gpu_bench_n.sh N TAG [steps]
N=$1; TAG=$2; STEPS=${3:-100}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader > $O/${TAG}_gpu.txt 2>&1; free -g | head -2 >> $O/${TAG}_gpu.txt
SECONDS=0
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 --steps $STEPS --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --gpus $N --steps $STEPS --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
fi
echo "bench rc=$? wall ${SECONDS}s"
python scripts/bench_summary.py $O/${TAG}_bench.json || tail -40 $O/${TAG}_bench.err
