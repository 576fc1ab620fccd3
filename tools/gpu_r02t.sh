#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_long_reads.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -5 | cut -c1-300
timeout 900 python bench.py --stages 0 --cpu-pairs 400 --steps 5 --warmup 3 > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; grep "\[bench\]" gpurun_out/r02t_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02t_bench.json').read().strip().splitlines()[-1]); print(d['check']); print(d['roofline']['single_lane_step']['per_kernel_ms'])
"
timeout 600 python tools/long_reads_probe.py 20000 8000 60 2>/dev/null | tail -1 | cut -c1-600
