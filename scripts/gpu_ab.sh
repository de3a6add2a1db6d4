#!/bin/bash
# A/B of kernel-variant builds (scripts/build_variants.py) on one GPU: the same bench line against each library.
# Usage (under gpurun): bash scripts/gpu_ab.sh TAG "workload ..." "variant ..."      (variant "default" = the in-tree build)
TAG=${1:-ab}; WLS=${2:-"icase4"}; VARS=${3:-"default nx2 nx0_warp_emit"}
O=gpurun_out; mkdir -p $O
for wl in $WLS; do for v in $VARS; do
  libp=""; [ "$v" != "default" ] && libp=build/variants/libkrep_b200_$v.so
  KREP_B200_LIB=$libp timeout 300 python bench.py --workload $wl --gib 10 --no-side --no-e2e --no-cpu --steps 30 > $O/${TAG}_v_${v}_$wl.json 2> $O/${TAG}_v_${v}_$wl.err
  echo -n "variant $v $wl: "; python scripts/bench_summary.py $O/${TAG}_v_${v}_$wl.json | head -1 || tail -20 $O/${TAG}_v_${v}_$wl.err
done; done
