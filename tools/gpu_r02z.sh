#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02f_pytest.log; tail -4 gpurun_out/r02f_pytest.log | cut -c1-300
HLALA_TYPING_PROFILE=1 timeout 1500 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc $?"
grep -E "\[bench\]" gpurun_out/r02f_bench.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']); print(json.dumps(d['pipeline_config2'])[:900]); print(json.dumps(d['stages'].get('long_reads'))[:900]); print(d['roofline']['frac'], d['roofline']['traffic'])
"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_ref.json 2> gpurun_out/r02f_ref.err; tail -c 300 gpurun_out/r02f_ref.json
W="--stages 0 --pipeline 0 --cpu-pairs 400 --e2e-steps 0"
HLALA_LANES=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py $W --steps 1 --warmup 1 > gpurun_out/r02f_launch.log 2>&1; echo "launch list rc $?"
HLALA_LANES=1 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_extend_lean -s 1 -c 1 -o gpurun_out/r02f_lean -f python bench.py $W --steps 1 --warmup 1 > gpurun_out/r02f_ncu.log 2>&1; echo "ncu rc $?"
HLALA_LANES=1 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_chain_seed -s 2 -c 1 -o gpurun_out/r02f_seed -f python bench.py $W --steps 1 --warmup 1 > gpurun_out/r02f_ncu2.log 2>&1; echo "ncu2 rc $?"

python -c "
