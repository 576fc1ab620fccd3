#!/bin/bash
mkdir -p gpurun_out
W="--stages 0 --cpu-pairs 400 --e2e-steps 0 --steps 4 --warmup 2"
for mode in clip level clip level; do
  HLALA_DP_SORT=$mode timeout 900 python bench.py $W > gpurun_out/r02v_$mode.json 2> gpurun_out/r02v_$mode.err; echo "$mode: $(grep '\[bench\] resident' gpurun_out/r02v_$mode.err) $(grep 'extension ms' gpurun_out/r02v_$mode.err)"
  python -c "
import json
d=json.loads(open('gpurun_out/r02v_$mode.json').read().strip().splitlines()[-1]); print(d['check'])"
done
