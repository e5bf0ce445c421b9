#!/bin/bash
# ncu --set full with source correlation of selected kernels of the 16K^3 city build; dumps the raw page and the per-line
# source page (CSV) to gpurun_out/.   usage: tools/gpu_ncu_source.sh TAG "kernel-regex" SKIP COUNT
TAG=${1:-src}; RX=${2:-k_classify_filtered}; SKIP=${3:-28}; CNT=${4:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/ncu_city.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
pkg = g._pkg()
tris = pkg.meshgen.city(256)
v = tris.reshape(-1, 3)
bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
t = pkg.GeomOctree(tris)
st = t.build(14, 4, bbox=bbox)
print(st["nTotalVoxels"], st["msTotal"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -o gpurun_out/tmp_$TAG python /tmp/ncu_city.py > gpurun_out/ncu_${TAG}.log 2>&1
ncu -i gpurun_out/tmp_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/tmp_$TAG.ncu-rep --page source --csv > gpurun_out/ncu_${TAG}_source.csv 2>/dev/null
rm -f gpurun_out/tmp_$TAG.ncu-rep
tail -2 gpurun_out/ncu_${TAG}.log
ls -la gpurun_out/ncu_${TAG}_*.csv
