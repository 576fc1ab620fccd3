#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_long_reads.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -4 | cut -c1-300
W="--stages 0 --pipeline 0 --cpu-pairs 400 --e2e-steps 0 --steps 3 --warmup 2"
for v in f g f g; do
  HLALA_B200_LIB=$PWD/ab/lib_$v.so timeout 900 python bench.py $W > gpurun_out/r02t_$v.json 2> gpurun_out/r02t_$v.err; echo "$v: $(grep '\[bench\] resident' gpurun_out/r02t_$v.err | cut -c1-160)"
  python -c "
import json
d=json.loads(open('gpurun_out/r02t_$v.json').read().strip().splitlines()[-1]); print(d['check'])"
done
