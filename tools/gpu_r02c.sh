#!/bin/bash
mkdir -p gpurun_out
W="--pairs 250000 --levels 1000000 --stages 0 --cpu-pairs 400"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
HLALA_DP_TRACE=1 HLALA_LANES=1 timeout 600 python bench.py $W --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/r02c_trace.json 2> gpurun_out/r02c_trace.err
grep "dp-trace" gpurun_out/r02c_trace.err | head -6
timeout 600 python bench.py $W --steps 3 --warmup 3 > gpurun_out/r02c_bench_lean.json 2> gpurun_out/r02c_bench_lean.err
grep "\[bench\]" gpurun_out/r02c_bench_lean.err
python -c "
import json
for f in ('gpurun_out/r02c_bench_lean.json','gpurun_out/r02a_bench_group.json'):
    try: print(f, json.load(open(f))['check'])
    except Exception as e: print(f, e)
"
