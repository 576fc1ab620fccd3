#!/bin/bash
mkdir -p gpurun_out
W="--stages 0 --cpu-pairs 400 --e2e-steps 0 --steps 3 --warmup 2"
for b in 2 4 6 10 16 24; do
  HLALA_LN_BATCH=$b timeout 900 python bench.py $W > gpurun_out/r02w_$b.json 2> gpurun_out/r02w_$b.err; echo "batch $b: $(grep '\[bench\] resident' gpurun_out/r02w_$b.err | cut -c1-60) $(grep 'extension ms' gpurun_out/r02w_$b.err)"
done
