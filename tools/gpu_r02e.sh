#!/bin/bash
mkdir -p gpurun_out
W="--stages 0 --cpu-pairs 400"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
HLALA_DP_TRACE=1 HLALA_LANES=1 timeout 900 python bench.py $W --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/r02e_trace.json 2> gpurun_out/r02e_trace.err
grep "dp-trace" gpurun_out/r02e_trace.err | head -5
timeout 900 python bench.py $W --steps 5 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
grep "\[bench\]" gpurun_out/r02e_bench.err
python -c "
import json
for f in ('gpurun_out/r02e_bench.json',):
    d=json.load(open(f)); print(f, d['check']); print(d['roofline']['single_lane_step']['per_kernel_ms'])
"
