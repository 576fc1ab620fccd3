#!/bin/bash
mkdir -p gpurun_out
W="--pairs 250000 --levels 1000000 --stages 0 --cpu-pairs 400"
HLALA_LANES=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_extend_lean -s 2 -c 2 -o gpurun_out/r02b_lean -f python bench.py $W --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/r02b_ncu.log 2>&1
tail -5 gpurun_out/r02b_ncu.log
