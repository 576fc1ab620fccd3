#!/bin/bash
mkdir -p gpurun_out
W="--stages 0 --cpu-pairs 400"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
HLALA_DP_TRACE=1 HLALA_LANES=1 timeout 900 python bench.py $W --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/r02f_trace.json 2> gpurun_out/r02f_trace.err
grep "dp-trace" gpurun_out/r02f_trace.err | head -5
timeout 900 python bench.py $W --steps 5 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
grep "\[bench\]" gpurun_out/r02f_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r02f_bench.json')); print(d['check']); print(d['roofline']['single_lane_step']['per_kernel_ms'])
"
HLALA_LANES=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_extend_lean -s 2 -c 1 -o gpurun_out/r02f_lean -f python bench.py $W --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/r02f_ncu.log 2>&1
