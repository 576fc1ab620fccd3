#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_long_reads.py -m gpu -x -q 2>&1 | tail -3
# full default bench (driver's command) with the typing host profile on stderr
HLALA_TYPING_PROFILE=1 timeout 1500 python bench.py > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc $?"
grep -E "\[bench\]|typing-profile" gpurun_out/r02s_bench.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r02s_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']); print(json.dumps(d['stages'].get('long_reads'))[:1500]); print(json.dumps(d['stages'].get('typing'))[:600])
"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02s_ref.json 2> gpurun_out/r02s_ref.err; tail -c 600 gpurun_out/r02s_ref.json
# launch list of one step (cold-cache, serialised: shares only)
W="--stages 0 --cpu-pairs 400 --e2e-steps 0"
HLALA_LANES=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s_launches.csv python bench.py $W --steps 1 --warmup 1 > gpurun_out/r02s_launch.log 2>&1; echo "launch list rc $?"
# one full capture of the dominant kernel at the bench's launch size (one wave = whole batch)
HLALA_LANES=1 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_extend_lean -s 1 -c 1 -o gpurun_out/r02s_lean -f python bench.py $W --steps 1 --warmup 1 > gpurun_out/r02s_ncu.log 2>&1; echo "ncu rc $?"
ls -la gpurun_out/r02s_*
