#!/bin/bash
# A/B of two builds of libsvb.so on one box: alternates svdag-compression_b200/build/libsvb_<name>.so (copied over libsvb.so) and
# runs tools/gpu_ab_inproc.py on each.   usage: tools/gpu_ab_libs.sh nameA nameB [rounds]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=svdag-compression_b200
cp $P/libsvb.so /tmp/libsvb_orig.so
for r in $(seq 1 ${3:-2}); do
  for n in $1 $2; do
    cp $P/build/libsvb_$n.so $P/libsvb.so
    echo "== $n (round $r)"
    AB_WARMUP=1 AB_REPS=3 timeout 300 python tools/gpu_ab_inproc.py "$n:" 2>/dev/null | tail -1
  done
done | tee gpurun_out/ab_libs.log
cp /tmp/libsvb_orig.so $P/libsvb.so
