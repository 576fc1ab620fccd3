#!/bin/bash
mkdir -p gpurun_out
W="--stages 0 --cpu-pairs 400"
HLALA_DP_TRACE=1 HLALA_LANES=1 timeout 900 python bench.py $W --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/r02d_trace.json 2> gpurun_out/r02d_trace.err
grep "dp-trace" gpurun_out/r02d_trace.err | head -10
HLALA_LEAN_BIG=1 HLALA_DP_TRACE=1 HLALA_LANES=1 timeout 900 python bench.py $W --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/r02d_trace_big.json 2> gpurun_out/r02d_trace_big.err
grep "dp-trace" gpurun_out/r02d_trace_big.err | head -6
HLALA_NO_LEAN_DP=1 HLALA_DP_TRACE=1 HLALA_LANES=1 timeout 900 python bench.py $W --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/r02d_trace_old.json 2> gpurun_out/r02d_trace_old.err
grep "dp-trace" gpurun_out/r02d_trace_old.err | head -5
timeout 900 python bench.py $W --steps 5 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
grep "\[bench\]" gpurun_out/r02d_bench.err
python -c "
import json
for f in ('gpurun_out/r02d_bench.json','gpurun_out/r02d_trace_old.json'):
    try: print(f, json.load(open(f))['check'])
    except Exception as e: print(f, e)
"
