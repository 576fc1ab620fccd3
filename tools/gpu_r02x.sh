#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02x_pytest.log; tail -5 gpurun_out/r02x_pytest.log | cut -c1-300
W="--stages 0 --cpu-pairs 400 --steps 4 --warmup 2"
timeout 900 python bench.py $W > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; grep '\[bench\]' gpurun_out/r02x_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02x_bench.json').read().strip().splitlines()[-1]); print(d['check'])"
