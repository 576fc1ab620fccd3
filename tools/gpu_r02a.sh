#!/bin/bash
# round 2, first GPU session: parity suite, then lean-vs-group A/B on a dev-size workload with per-tier trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02a_pytest.log
W="--pairs 250000 --levels 1000000 --stages 0 --cpu-pairs 400"
timeout 600 python bench.py $W --steps 3 --warmup 3 > gpurun_out/r02a_bench_lean.json 2> gpurun_out/r02a_bench_lean.err
HLALA_NO_LEAN_DP=1 timeout 600 python bench.py $W --steps 3 --warmup 3 > gpurun_out/r02a_bench_group.json 2> gpurun_out/r02a_bench_group.err
HLALA_DP_TRACE=1 HLALA_LANES=1 timeout 600 python bench.py $W --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/r02a_trace.json 2> gpurun_out/r02a_trace.err
tail -3 gpurun_out/r02a_pytest.log; grep "\[bench\]" gpurun_out/r02a_bench_lean.err gpurun_out/r02a_bench_group.err; grep "dp-trace" gpurun_out/r02a_trace.err | head -24
