#!/bin/bash
# scratch helper (not part of the product)
cd /root/repo
mkdir -p gpurun_out
python bench.py --workload city_small --steps 2 --warmup 1 2>&1 | tail -3
python bench.py --impl reference --workload city_small --steps 1 --warmup 0 2>&1 | tail -2
python bench.py 2>&1 | tail -2 | tee gpurun_out/bench_city16k.json
