#!/bin/bash
mkdir -p gpurun_out
W="--stages 0 --cpu-pairs 400 --steps 3 --warmup 2 --e2e-steps 1"
for cfg in "24000000000 0" "24000000000 1" "64000000000 0" "64000000000 1"; do
  set -- $cfg
  if [ "$2" = "1" ]; then export HLALA_LEAN_BIG=1; else unset HLALA_LEAN_BIG; fi
  HLALA_WAVE_BYTES=$1 timeout 900 python bench.py $W > gpurun_out/r02h_$1_$2.json 2> gpurun_out/r02h_$1_$2.err
  echo "wave $1 big $2"; grep "\[bench\]" gpurun_out/r02h_$1_$2.err | head -3
  python -c "
import json
d=json.load(open('gpurun_out/r02h_$1_$2.json')); print(d['check']['edge_checksum'], d['roofline']['single_lane_step']['per_kernel_ms'])
"
done
