#!/bin/bash
mkdir -p gpurun_out
W="--stages 0 --cpu-pairs 400 --e2e-steps 0 --steps 3 --warmup 2"
for v in c f c f; do
  HLALA_B200_LIB=$PWD/ab/lib_$v.so timeout 900 python bench.py $W > gpurun_out/r02y_$v.json 2> gpurun_out/r02y_$v.err; echo "$v: $(grep '\[bench\] resident' gpurun_out/r02y_$v.err | cut -c1-60) $(grep 'extension ms' gpurun_out/r02y_$v.err)"
  python -c "
import json
d=json.loads(open('gpurun_out/r02y_$v.json').read().strip().splitlines()[-1]); print(d['check'])"
done
